// Batched map-type forward models of ces/utils.py, one particle per column (enka.G_ens,
// ces/calibrate.py:106-130 evaluates them one Python call per particle).
#include "kernels.h"

namespace ces {

// out = exp(X) elementwise (lineal_log, ces/utils.py:39-42), zero in the padding columns.
__global__ void __launch_bounds__(256) exp_kernel(const double* __restrict__ X, long long ldx, long long cols,
                                                  double* __restrict__ out, long long ldo) {
    const long long j = (long long)blockIdx.x * 256 + threadIdx.x;
    if (j >= ldo) return;
    out[(size_t)blockIdx.y * ldo + j] = j < cols ? exp(X[(size_t)blockIdx.y * ldx + j]) : 0.0;
}
int exp_map(cudaStream_t st, const double* X, int64_t ldx, int64_t rows, int64_t cols, double* out, int64_t ldo) {
    dim3 grid((unsigned)ceil_div(ldo, 256), (unsigned)rows);
    exp_kernel<<<grid, 256, 0, st>>>(X, ldx, cols, out, ldo);
    CES_LAUNCHED(1);
    return CES_OK;
}

// elliptic (ces/utils.py:72-89): p(x) = u2 x + exp(-u1) (x - x^2)/2 at x1, x2.
__global__ void __launch_bounds__(256) elliptic_kernel(const double* __restrict__ U, long long ldu, long long cols,
                                                       double x1, double x2, double* __restrict__ G, long long ldg) {
    const long long j = (long long)blockIdx.x * 256 + threadIdx.x;
    if (j >= cols) return;
    const double u1 = U[j], u2 = U[ldu + j];
    const double e = exp(-u1);
    G[j] = (u2 * x1) + (e * (-x1 * x1 + x1) * 0.5);
    G[ldg + j] = (u2 * x2) + (e * (-x2 * x2 + x2) * 0.5);
}
// banana (ces/utils.py:116-122): (a u1, u2/a - b (u1^2 + a^2)).
__global__ void __launch_bounds__(256) banana_kernel(const double* __restrict__ U, long long ldu, long long cols,
                                                     double a, double b, double* __restrict__ G, long long ldg) {
    const long long j = (long long)blockIdx.x * 256 + threadIdx.x;
    if (j >= cols) return;
    const double u1 = U[j], u2 = U[ldu + j];
    G[j] = u1 * a;
    G[ldg + j] = u2 / a - b * (u1 * u1 + a * a);
}
int elliptic_map(cudaStream_t st, const double* U, int64_t ldu, int64_t cols, double x1, double x2, double* G, int64_t ldg) {
    elliptic_kernel<<<(unsigned)ceil_div(cols, 256), 256, 0, st>>>(U, ldu, cols, x1, x2, G, ldg);
    CES_LAUNCHED(1);
    return CES_OK;
}
int banana_map(cudaStream_t st, const double* U, int64_t ldu, int64_t cols, double a, double b, double* G, int64_t ldg) {
    banana_kernel<<<(unsigned)ceil_div(cols, 256), 256, 0, st>>>(U, ldu, cols, a, b, G, ldg);
    CES_LAUNCHED(1);
    return CES_OK;
}

}  // namespace ces
