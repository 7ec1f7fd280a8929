// Largest eigenvalue of the interaction matrix D = (1/J) E^T Gamma^-1 R  (time_step='spectral',
// ces/calibrate.py:249-251: radspec = eigvals(D).real.max(), hk = 1 / radspec).
//
// The reference hands the non-symmetric J x J matrix to LAPACK's dgeev.  Here D is never eigen-decomposed: the non-zero
// eigenvalues of D = X Y (X = E^T, Y = Gamma^-1 R / J) are those of Y X = Gamma^-1 R E^T / J, and R E^T = E E^T because
// R = E + (gbar - y) 1^T and E 1 = 0.  So spec(D) \ {0} = spec(Gamma^-1 C^pp) with C^pp = E E^T / J the k x k matrix the
// K11 phase already forms (and all-reduces when the ensemble is sharded); all of them are real and >= 0, and the rest of
// D's spectrum is 0, so eigvals(D).real.max() = lambda_max(Gamma^-1 C^pp).
//
// B = C^pp Gamma^-1 is self-adjoint in the inner product <x, y> = x^T Gamma^-1 y.  Lanczos in that inner product with
// full re-orthogonalisation (classical Gram-Schmidt applied twice) needs one product with C^pp and one with Gamma^-1
// per step and two stored bases, V and Z = Gamma^-1 V (so every inner product is a plain dot product against Z); no
// factor of Gamma, no transposes.  The recurrence coefficients stay on the device; lambda_max of the tridiagonal matrix
// is found by 256-way multisection of the Sturm count every few steps, and the host stops the iteration when two
// successive estimates agree to 2e-15 (or the Krylov space is exhausted).
#include "kernels.h"

namespace ces {

// y[i] = sum_c M[i, c] x[c]  (rows x cols, row-major): one warp per row.
__global__ void __launch_bounds__(256) gemv_rows_kernel(const double* __restrict__ M, long long ld, int rows, int cols,
                                                        const double* __restrict__ x, double* __restrict__ y, int accumulate) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    double a = 0.0;
    for (int c = lane; c < cols; c += 32) a += M[(size_t)row * ld + c] * x[c];
    a = warp_sum(a);
    if (lane == 0) y[row] = accumulate ? y[row] + a : a;
}

// w[c] -= sum_i coef[i] V[i, c]
__global__ void __launch_bounds__(256) orth_update_kernel(const double* __restrict__ V, long long ld, int rows, int cols,
                                                          const double* __restrict__ coef, double* __restrict__ w) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= cols) return;
    double a = 0.0;
    for (int i = 0; i < rows; ++i) a += coef[i] * V[(size_t)i * ld + c];
    w[c] -= a;
}

// state: [0] largest |alpha| seen (breakdown scale), [1] 1 once an invariant subspace was found, [2] steps completed
// One CTA.  beta = sqrt(w . zw); V[j+1] = w / beta, Z[j+1] = zw / beta; alpha[j] = c1[j] + c2[j].
__global__ void __launch_bounds__(1024) lanczos_finish_kernel(const double* __restrict__ w, const double* __restrict__ zw, int n,
                                                              const double* __restrict__ c1, const double* __restrict__ c2, int j,
                                                              double* __restrict__ alpha, double* __restrict__ beta,
                                                              double* __restrict__ vnext, double* __restrict__ znext,
                                                              double* __restrict__ state) {
    __shared__ double scratch[32];
    __shared__ double bcast;
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) a += w[i] * zw[i];
    a = block_sum(a, scratch);
    if (threadIdx.x == 0) {
        if (state[1] == 0.0) {
            const double al = c1[j] + c2[j];
            alpha[j] = al;
            const double scale = fmax(state[0], fabs(al));
            state[0] = scale;
            const double b = a > 0.0 ? sqrt(a) : 0.0;
            state[2] = (double)(j + 1);
            if (!(b > 1e-14 * scale)) {      // the Krylov space is invariant: T_{j+1} holds exact eigenvalues
                beta[j] = 0.0;
                state[1] = 1.0;
                bcast = 0.0;
            } else {
                beta[j] = b;
                bcast = 1.0 / b;
            }
        } else {
            bcast = 0.0;
        }
    }
    __syncthreads();
    const double s = bcast;
    if (vnext)
        for (int i = threadIdx.x; i < n; i += 1024) {
            vnext[i] = w[i] * s;
            znext[i] = zw[i] * s;
        }
}

// Deterministic start vector (a fixed hash of the index, in [0.5, 1.5)), then normalised like any Lanczos vector.
__global__ void lanczos_start_kernel(double* __restrict__ w, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned x = (unsigned)i * 2654435761u + 12345u;
    x ^= x >> 16; x *= 2246822519u; x ^= x >> 13; x *= 3266489917u; x ^= x >> 16;
    w[i] = 0.5 + (double)x / 4294967296.0;
}

__global__ void scale_rows_kernel(const double* __restrict__ d, const double* __restrict__ x, double* __restrict__ y, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = d[i] * x[i];
}

// Largest eigenvalue of the symmetric tridiagonal matrix (alpha[0..m), beta[0..m-1)) by multisection of the Sturm count:
// 256 abscissae per round, each thread counts the eigenvalues below its own.  m = state[2].
__global__ void __launch_bounds__(256) tridiag_lmax_kernel(const double* __restrict__ alpha, const double* __restrict__ beta,
                                                           const double* __restrict__ state, double* __restrict__ out) {
    __shared__ double lo_s, hi_s;
    __shared__ int first;
    const int m = (int)state[2];
    if (m <= 0) { if (threadIdx.x == 0) *out = 0.0; return; }
    if (threadIdx.x == 0) {
        double lo = 1e300, hi = -1e300;                 // Gershgorin
        for (int i = 0; i < m; ++i) {
            const double rad = (i > 0 ? fabs(beta[i - 1]) : 0.0) + (i < m - 1 ? fabs(beta[i]) : 0.0);
            lo = fmin(lo, alpha[i] - rad);
            hi = fmax(hi, alpha[i] + rad);
        }
        const double pad = 1e-15 * fmax(fabs(lo), fabs(hi)) + 1e-300;
        lo_s = lo - pad;
        hi_s = hi + pad;
    }
    __syncthreads();
    for (int round = 0; round < 16; ++round) {
        const double lo = lo_s, hi = hi_s;
        if (threadIdx.x == 0) first = 256;
        __syncthreads();
        const double x = lo + (hi - lo) * ((double)(threadIdx.x + 1) / 257.0);
        // number of eigenvalues < x
        int count = 0;
        double dprev = 1.0;
        for (int i = 0; i < m; ++i) {
            const double b = i > 0 ? beta[i - 1] : 0.0;
            double dcur = alpha[i] - x - (i > 0 ? b * b / dprev : 0.0);
            if (dcur == 0.0) dcur = -1e-300;
            count += dcur < 0.0;
            dprev = dcur;
        }
        if (count == m) atomicMin(&first, (int)threadIdx.x);     // all eigenvalues below x: lambda_max < x
        __syncthreads();
        const int f = first;
        __syncthreads();
        if (threadIdx.x == 0) {
            const double nlo = f == 0 ? lo : lo + (hi - lo) * ((double)f / 257.0);
            const double nhi = f == 256 ? hi : lo + (hi - lo) * ((double)(f + 1) / 257.0);
            lo_s = nlo;
            hi_s = nhi;
        }
        __syncthreads();
        if (hi_s - lo_s <= 4e-16 * fmax(fabs(lo_s), fabs(hi_s))) break;
    }
    if (threadIdx.x == 0) *out = 0.5 * (lo_s + hi_s);
}

// lambda_max(Ginv C) for symmetric positive semi-definite C (n x n, ld) and Gamma^-1 given either as a dense matrix
// (Ginv, ld) or as its diagonal (ginv_diag).  `work` holds >= (2 * (mmax + 1) + 2) * n + 4 * mmax + 8 doubles.
int spectral_radius(cudaStream_t st, const double* C, int64_t ld, int64_t n, const double* Ginv, const double* ginv_diag,
                    double* work, int64_t mmax, double* lambda_host, int* steps_host) {
    if (n < 1 || mmax < 1 || (!Ginv && !ginv_diag)) return fail(CES_ERR_INVALID, "spectral_radius: bad argument%s", "");
    if (mmax > n) mmax = n;
    double* V = work;                                   // (mmax + 1) x n
    double* Z = V + (mmax + 1) * n;                     // (mmax + 1) x n
    double* w = Z + (mmax + 1) * n;
    double* zw = w + n;
    double* alpha = zw + n;
    double* beta = alpha + mmax;
    double* c1 = beta + mmax;
    double* c2 = c1 + mmax;
    double* state = c2 + mmax;                          // 3 doubles + 1 output
    double* lam = state + 3;
    const unsigned gv = (unsigned)ceil_div(n, 256);
    auto apply_ginv = [&](const double* x, double* y) -> int {
        if (ginv_diag) {
            scale_rows_kernel<<<gv, 256, 0, st>>>(ginv_diag, x, y, (int)n);
        } else {
            gemv_rows_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, st>>>(Ginv, ld, (int)n, (int)n, x, y, 0);
        }
        CES_LAUNCHED(1);
        return CES_OK;
    };
    CES_CUDA(cudaMemsetAsync(state, 0, 4 * sizeof(double), st));
    CES_CUDA(cudaMemsetAsync(c1, 0, 2 * mmax * sizeof(double), st));
    // v_0: start vector normalised in the Gamma^-1 inner product (the finish kernel with j = -1 semantics is spelled out)
    lanczos_start_kernel<<<gv, 256, 0, st>>>(w, (int)n);
    CES_LAUNCHED(1);
    CES_TRY(apply_ginv(w, zw));
    {
        // reuse the finish kernel: alpha[mmax-1] is scratch here and overwritten later; state[2] is reset below
        lanczos_finish_kernel<<<1, 1024, 0, st>>>(w, zw, (int)n, c1, c2, (int)mmax - 1, alpha, beta, V, Z, state);
        CES_LAUNCHED(1);
        CES_CUDA(cudaMemsetAsync(state, 0, 4 * sizeof(double), st));
    }
    double prev = -1.0, cur = 0.0;
    int agree = 0, steps = 0;
    bool converged = (mmax == n);                                   // a full Krylov space is exact
    const int check_every = 4;
    for (int64_t j = 0; j < mmax; ++j) {
        const double* zj = Z + j * n;
        // w = C z_j ; orthogonalise against v_0..v_j in the Gamma^-1 inner product (coefficients = Z w), twice
        gemv_rows_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, st>>>(C, ld, (int)n, (int)n, zj, w, 0);
        gemv_rows_kernel<<<(unsigned)ceil_div(j + 1, 8), 256, 0, st>>>(Z, n, (int)(j + 1), (int)n, w, c1, 0);
        orth_update_kernel<<<gv, 256, 0, st>>>(V, n, (int)(j + 1), (int)n, c1, w);
        gemv_rows_kernel<<<(unsigned)ceil_div(j + 1, 8), 256, 0, st>>>(Z, n, (int)(j + 1), (int)n, w, c2, 0);
        orth_update_kernel<<<gv, 256, 0, st>>>(V, n, (int)(j + 1), (int)n, c2, w);
        CES_LAUNCHED(5);
        CES_TRY(apply_ginv(w, zw));
        const bool last = (j + 1 == mmax);
        lanczos_finish_kernel<<<1, 1024, 0, st>>>(w, zw, (int)n, c1, c2, (int)j, alpha, beta, last ? nullptr : V + (j + 1) * n,
                                                   last ? nullptr : Z + (j + 1) * n, state);
        CES_LAUNCHED(1);
        steps = (int)(j + 1);
        if ((j + 1) % check_every == 0 || last) {
            tridiag_lmax_kernel<<<1, 256, 0, st>>>(alpha, beta, state, lam);
            CES_LAUNCHED(1);
            double host[4];
            CES_CUDA(cudaMemcpyAsync(host, state, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
            CES_CUDA(cudaStreamSynchronize(st));
            cur = host[3];
            if (host[1] != 0.0) { converged = true; break; }        // invariant subspace: exact
            if (fabs(cur - prev) <= 2e-15 * fabs(cur)) { if (++agree >= 2) { converged = true; break; } } else agree = 0;
            prev = cur;
        }
    }
    if (lambda_host) *lambda_host = cur;
    // a NEGATIVE step count reports that the iteration stopped at the cap without two agreeing estimates: the value
    // is then a lower bound of lambda_max (Ritz values grow monotonically), i.e. hk = 1 / lambda may be too large
    if (steps_host) *steps_host = converged ? steps : -steps;
    return CES_OK;
}

}  // namespace ces
