// Batched 'pde'-type forward models of ces/utils.py: Lorenz 63 (ces/utils.py:124-229) and the two-scale Lorenz 96
// family (:231-447).  The reference integrates one particle per Python call (enka.G_pde, ces/calibrate.py:132-154:
// scipy odeint / solve_ivp with adaptive steps, then pandas/numpy window statistics); here the whole ensemble is
// integrated in one launch with a fixed-step classical Runge-Kutta scheme (`substeps` RK4 steps per output interval),
// the window statistics are accumulated on the fly and nothing but the (n_obs x J) statistics and the final states
// (the next iteration's initial conditions, ces/calibrate.py:390-396) is ever written.
//
// Parity: these systems are chaotic and the reference's integrators are adaptive, so trajectories agree with the
// reference only over short horizons (tests/test_gpu_lorenz.py pins that against golden vectors made with the real
// ces.utils classes); bit-level parity is against the numpy restatement of exactly this scheme (oracle/lorenz_oracle.py).
#include "kernels.h"

namespace ces {

// ------------------------------------------------------------------------------------------------ Lorenz 63
struct L63Rhs {
    double sigma, r, b;
    __device__ __forceinline__ void operator()(double x, double y, double z, double& dx, double& dy, double& dz) const {
        dx = sigma * (y - x);               // ces/utils.py:164-166
        dy = r * x - y - x * z;
        dz = x * y - b * z;
    }
};

// One thread per particle.  Statistics (ces/utils.py:181-194): means of (x, y, z, x^2, y^2, z^2, xy, xz, yz) over the
// last `window` output samples of t[1:].
__global__ void __launch_bounds__(128) lorenz63_kernel(const double* __restrict__ U, long long ldu, int p, long long cols,
                                                       int log_params, const double* __restrict__ W0, long long ldw0,
                                                       long long n_out, double dt_out, int substeps, long long window,
                                                       double* __restrict__ G, long long ldg, double* __restrict__ Wend,
                                                       long long ldwe, double* __restrict__ traj, long long ldt) {
    const long long j = (long long)blockIdx.x * 128 + threadIdx.x;
    if (j >= cols) return;
    L63Rhs f;
    f.sigma = 10.0;                                                      // :155
    f.r = p > 0 ? U[j] : (log_params ? log(28.0) : 28.0);                // args = k[:p] fills (r, b) in order (:150, :145)
    f.b = p > 1 ? U[ldu + j] : (log_params ? log(8.0 / 3.0) : 8.0 / 3.0);
    if (log_params) { f.r = exp(f.r); f.b = exp(f.b); }                  // lorenz63_log, :213-214
    double x = W0[j], y = W0[ldw0 + j], z = W0[2 * ldw0 + j];
    const double h = dt_out / (double)substeps, h2 = 0.5 * h, h6 = h / 6.0;
    double s[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long first = n_out - window;                              // samples first .. n_out-1 form the last window
    if (traj) { traj[j] = x; traj[ldt + j] = y; traj[2 * ldt + j] = z; }
    for (long long i = 1; i < n_out; ++i) {
        for (int q = 0; q < substeps; ++q) {
            double k1x, k1y, k1z, k2x, k2y, k2z, k3x, k3y, k3z, k4x, k4y, k4z;
            f(x, y, z, k1x, k1y, k1z);
            f(x + h2 * k1x, y + h2 * k1y, z + h2 * k1z, k2x, k2y, k2z);
            f(x + h2 * k2x, y + h2 * k2y, z + h2 * k2z, k3x, k3y, k3z);
            f(x + h * k3x, y + h * k3y, z + h * k3z, k4x, k4y, k4z);
            x += h6 * ((k1x + k4x) + 2.0 * (k2x + k3x));
            y += h6 * ((k1y + k4y) + 2.0 * (k2y + k3y));
            z += h6 * ((k1z + k4z) + 2.0 * (k2z + k3z));
        }
        if (traj) { traj[(3 * i) * ldt + j] = x; traj[(3 * i + 1) * ldt + j] = y; traj[(3 * i + 2) * ldt + j] = z; }
        if (i >= first) {
            s[0] += x; s[1] += y; s[2] += z;
            s[3] += x * x; s[4] += y * y; s[5] += z * z;
            s[6] += x * y; s[7] += x * z; s[8] += y * z;
        }
    }
    const double inv = 1.0 / (double)window;
#pragma unroll
    for (int m = 0; m < 9; ++m) G[(size_t)m * ldg + j] = s[m] * inv;
    if (Wend) {
        Wend[j] = x;
        Wend[ldwe + j] = y;
        Wend[2 * ldwe + j] = z;
    }
}

int lorenz63_forward(cudaStream_t st, int log_params, const double* U, int64_t ldu, int64_t p, int64_t cols, const double* W0,
                     int64_t ldw0, int64_t n_out, double dt_out, int substeps, int64_t window, double* G, int64_t ldg,
                     double* Wend, int64_t ldwe, double* traj, int64_t ldt) {
    if (cols == 0) return CES_OK;
    lorenz63_kernel<<<(unsigned)ceil_div(cols, 128), 128, 0, st>>>(U, ldu, (int)p, cols, log_params, W0, ldw0, n_out, dt_out,
                                                                   substeps, window, G, ldg, Wend, ldwe, traj, ldt);
    CES_LAUNCHED(1);
    return CES_OK;
}

// ------------------------------------------------------------------------------------------------ Lorenz 96
// One CTA per particle, one thread per state variable (n_slow slow X_k followed by n_slow * n_fast fast Y_j,
// ces/utils.py:289-308).  The stage states of the RK4 step alternate between two shared-memory buffers, so a step costs
// four barriers.  Statistics (ces/utils.py:332-342), accumulated by the slow threads over the last window of the samples
// after `skip`: mean X_k, mean X_k^2, mean of Ybar_k, mean of (Y^2)bar_k, mean of X_k Ybar_k  (bar = mean over the n_fast
// fast variables of slow variable k).
struct L96Params { double h, F, c, b; };

__device__ __forceinline__ double l96_rhs(const double* __restrict__ w, int v, int ns, int nf, const L96Params& pr) {
    if (v < ns) {
        const int k = v;
        const double xm1 = w[k == 0 ? ns - 1 : k - 1], xm2 = w[k < 2 ? ns + k - 2 : k - 2], xp1 = w[k == ns - 1 ? 0 : k + 1];
        const double* y = w + ns + k * nf;
        double ya = 0.0, yb = 0.0;                      // two chains: the slow threads are the critical path of a stage
        int l = 0;
        for (; l + 1 < nf; l += 2) { ya += y[l]; yb += y[l + 1]; }
        if (l < nf) ya += y[l];
        const double ybar = (ya + yb) / (double)nf;
        return -xm1 * (xm2 - xp1) - w[k] + pr.F - (pr.h * pr.c) * ybar;                      // :299-301
    }
    const int n = ns * nf, j = v - ns;
    const double* y = w + ns;
    const double yp1 = y[j + 1 == n ? 0 : j + 1], yp2 = y[j + 2 >= n ? j + 2 - n : j + 2], ym1 = y[j == 0 ? n - 1 : j - 1];
    return -pr.c * pr.b * yp1 * (yp2 - ym1) - pr.c * y[j] + ((pr.h * pr.c) / (double)nf) * w[j / nf];   // :303-305
}

__global__ void __launch_bounds__(512) lorenz96_kernel(const double* __restrict__ U, long long ldu, int p, int4 slot,
                                                       int ns, int nf, const double* __restrict__ W0, long long ldw0,
                                                       long long n_out, double dt_out, int substeps, long long skip,
                                                       long long window, int out_mode, int out_col,
                                                       double* __restrict__ G, long long ldg, double* __restrict__ Wend,
                                                       long long ldwe, double* __restrict__ traj, long long ldt) {
    extern __shared__ double smem[];
    const int nstate = ns * (nf + 1);
    double* buf0 = smem;
    double* buf1 = smem + nstate;
    double* stats = buf1 + nstate;              // 5 x ns, written once at the end
    const long long j = blockIdx.x;
    const int v = threadIdx.x;
    const bool live = v < nstate;
    L96Params pr = {1.0, 10.0, log(10.0), 10.0};                         // defaults of model(), :289
    {
        const int slots[4] = {slot.x, slot.y, slot.z, slot.w};            // U row i -> parameter slot (0 h, 1 F, 2 log c, 3 b)
        for (int i = 0; i < p && i < 4; ++i) {
            const double val = U[(size_t)i * ldu + j];
            if (slots[i] == 0) pr.h = val;
            else if (slots[i] == 1) pr.F = val;
            else if (slots[i] == 2) pr.c = val;
            else if (slots[i] == 3) pr.b = val;
        }
        pr.c = exp(pr.c);                                                 // :293
    }
    double w = live ? W0[(size_t)v * ldw0 + j] : 0.0;
    const double h = dt_out / (double)substeps, h2 = 0.5 * h, h6 = h / 6.0;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
    // the last window of the samples skip .. n_out-1 (reshape(n_state, -1, window)[..., -1])
    const long long first = n_out - window;
    if (live) buf0[v] = w;
    if (traj && live) traj[(size_t)v * ldt + j] = w;
    __syncthreads();
    for (long long i = 1; i < n_out; ++i) {
        for (int q = 0; q < substeps; ++q) {
            // buf0 holds the current state on entry and on exit
            double k1 = 0, k2 = 0, k3 = 0, k4 = 0;
            if (live) { k1 = l96_rhs(buf0, v, ns, nf, pr); buf1[v] = w + h2 * k1; }
            __syncthreads();
            if (live) { k2 = l96_rhs(buf1, v, ns, nf, pr); buf0[v] = w + h2 * k2; }
            __syncthreads();
            if (live) { k3 = l96_rhs(buf0, v, ns, nf, pr); buf1[v] = w + h * k3; }
            __syncthreads();
            if (live) { k4 = l96_rhs(buf1, v, ns, nf, pr); w += h6 * ((k1 + k4) + 2.0 * (k2 + k3)); buf0[v] = w; }
            __syncthreads();
        }
        if (traj && live) traj[((size_t)i * nstate + v) * ldt + j] = w;
        if (i >= first && v < ns) {
            const double* y = buf0 + ns + v * nf;
            double yb = 0.0, y2b = 0.0;
            for (int l = 0; l < nf; ++l) { yb += y[l]; y2b += y[l] * y[l]; }
            yb /= (double)nf;
            y2b /= (double)nf;
            s0 += w; s1 += w * w; s2 += yb; s3 += y2b; s4 += w * yb;
        }
    }
    (void)skip;
    const double inv = 1.0 / (double)window;
    if (v < ns) {
        stats[0 * ns + v] = s0 * inv; stats[1 * ns + v] = s1 * inv; stats[2 * ns + v] = s2 * inv;
        stats[3 * ns + v] = s3 * inv; stats[4 * ns + v] = s4 * inv;
    }
    __syncthreads();
    if (out_mode == 0) {                       // lorenz96.statistics: all 5 n_slow values
        for (int m = v; m < 5 * ns; m += blockDim.x) G[(size_t)m * ldg + j] = stats[m];
    } else if (v < 5) {                        // lorenz96_hom.statistics: mean over k (mode 1) or column out_col (mode 2)
        double a;
        if (out_mode == 1) {
            a = 0.0;
            for (int k = 0; k < ns; ++k) a += stats[v * ns + k];
            a /= (double)ns;
        } else {
            a = stats[v * ns + out_col];
        }
        G[(size_t)v * ldg + j] = a;
    }
    if (Wend && live) Wend[(size_t)v * ldwe + j] = w;
}

int lorenz96_forward(cudaStream_t st, const int* slots, int n_slow, int n_fast, const double* U, int64_t ldu, int64_t p,
                     int64_t cols, const double* W0, int64_t ldw0, int64_t n_out, double dt_out, int substeps, int64_t skip,
                     int64_t window, int out_mode, int out_col, double* G, int64_t ldg, double* Wend, int64_t ldwe,
                     double* traj, int64_t ldt) {
    if (cols == 0) return CES_OK;
    const int nstate = n_slow * (n_fast + 1);
    if (nstate > 512) return fail(CES_ERR_INVALID, "lorenz96: n_slow (n_fast + 1) must not exceed %s%lld", "", 512);
    const size_t smem = ((size_t)2 * nstate + 5 * n_slow) * sizeof(double);
    lorenz96_kernel<<<(unsigned)cols, (unsigned)round_up(nstate, 32), smem, st>>>(
        U, ldu, (int)p, make_int4(slots[0], slots[1], slots[2], slots[3]), n_slow, n_fast, W0, ldw0, n_out, dt_out, substeps,
        skip, window, out_mode, out_col, G, ldg, Wend, ldwe, traj, ldt);
    CES_LAUNCHED(1);
    return CES_OK;
}

}  // namespace ces

using namespace ces;

extern "C" {

int ces_lorenz63_forward(void* stream, int log_params, const double* U_dev, int64_t ldu, int64_t p, int64_t cols,
                         const double* W0_dev, int64_t ldw0, int64_t n_out, double dt_out, int substeps, int64_t window,
                         double* G_dev, int64_t ldg, double* Wend_dev, int64_t ldwe, double* traj_dev, int64_t ldt) {
    if (!U_dev || !W0_dev || !G_dev || p < 0 || p > 2 || cols < 0 || ldu < cols || ldw0 < cols || ldg < cols ||
        (Wend_dev && ldwe < cols) || (traj_dev && ldt < cols))
        return fail(CES_ERR_INVALID, "ces_lorenz63_forward: bad argument%s", "");
    if (n_out < 2 || !(dt_out > 0.0) || substeps < 1 || window < 1 || (n_out - 1) % window != 0)
        return fail(CES_ERR_INVALID, "ces_lorenz63_forward: needs n_out >= 2, dt > 0, substeps >= 1 and (n_out - 1) a multiple of "
                                     "the window (%s%lld samples)", "", (long long)window);
    return lorenz63_forward(static_cast<cudaStream_t>(stream), log_params, U_dev, ldu, p, cols, W0_dev, ldw0, n_out, dt_out,
                            substeps, window, G_dev, ldg, Wend_dev, ldwe, traj_dev, ldt);
}

int ces_lorenz96_forward(void* stream, const int* param_slots, int n_slow, int n_fast, const double* U_dev, int64_t ldu,
                         int64_t p, int64_t cols, const double* W0_dev, int64_t ldw0, int64_t n_out, double dt_out,
                         int substeps, int64_t skip, int64_t window, int out_mode, int out_col, double* G_dev, int64_t ldg,
                         double* Wend_dev, int64_t ldwe, double* traj_dev, int64_t ldt) {
    if (!param_slots || !U_dev || !W0_dev || !G_dev || p < 0 || p > 4 || cols < 0 || ldu < cols || ldw0 < cols || ldg < cols ||
        (Wend_dev && ldwe < cols) || (traj_dev && ldt < cols) || n_slow < 3 || n_fast < 1 || out_mode < 0 || out_mode > 2 || out_col < 0 || out_col >= n_slow)
        return fail(CES_ERR_INVALID, "ces_lorenz96_forward: bad argument%s", "");
    for (int i = 0; i < 4; ++i)
        if (param_slots[i] < -1 || param_slots[i] > 3) return fail(CES_ERR_INVALID, "ces_lorenz96_forward: bad parameter slot%s", "");
    if (n_out < 2 || !(dt_out > 0.0) || substeps < 1 || window < 1 || skip < 1 || skip >= n_out || (n_out - skip) % window != 0)
        return fail(CES_ERR_INVALID, "ces_lorenz96_forward: the samples after the spin-up must be a whole number of windows "
                                     "(window = %s%lld samples)", "", (long long)window);
    return lorenz96_forward(static_cast<cudaStream_t>(stream), param_slots, n_slow, n_fast, U_dev, ldu, p, cols, W0_dev, ldw0,
                            n_out, dt_out, substeps, skip, window, out_mode, out_col, G_dev, ldg, Wend_dev, ldwe, traj_dev, ldt);
}

}  // extern "C"
