// Blocked Cholesky (right-looking) and blocked triangular solves for the d x d / k x k SPD systems
// of the update: chol(C^uu) every step (ces/calibrate.py:446,478,487,526), Gamma^-1 and Sigma0^-1 once,
// (Sigma0 + h C^uu)^-1 for the semi-implicit rule (:443-445).
//
// Diagonal CHOL_NB x CHOL_NB blocks are factorised and inverted inside one CTA in shared memory; every
// off-diagonal operation (panel solve, trailing update, block substitution) is a DMMA GEMM call, with
// the explicit inverse of the diagonal block standing in for the small triangular solve.
#include "kernels.h"

namespace ces {

constexpr int NB = CHOL_NB;
constexpr int NBP = NB + 1;   // padded shared-memory pitch

// One warp factors the 32 x 32 SPD block at a[o..o+31][o..o+31] (shared memory, pitch NBP: a column of the block is
// conflict-free across lanes, a row entry is a broadcast) and inverts the factor into x[o..][o..], with warp-level
// synchronisation only.  Cholesky-Crout: column c of L is finished by one dot product per lane over the columns already
// done (two loads and one FMA per term, no store in the inner loop, four independent accumulators), then one sqrt, one
// division and one store -- ~60 cycles of fixed latency per column instead of two block barriers.  The inverse follows by
// forward substitution, lane t owning column t.  `piv0` is the global index of the block's first pivot, `nvalid` the
// number of real (non-padding) pivots; the first non-positive pivot is reported like LAPACK's info.
__device__ __forceinline__ void warp_chol_inv32(double (*a)[NBP], double (*x)[NBP], int o, int lane, int piv0, int nvalid,
                                                int* info) {
    const int i = o + lane;
    for (int c = 0; c < 32; ++c) {
        const int cc = o + c;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int m = o;
        for (; m + 3 < cc; m += 4) {
            s0 = fma(a[i][m], a[cc][m], s0);
            s1 = fma(a[i][m + 1], a[cc][m + 1], s1);
            s2 = fma(a[i][m + 2], a[cc][m + 2], s2);
            s3 = fma(a[i][m + 3], a[cc][m + 3], s3);
        }
        for (; m < cc; ++m) s0 = fma(a[i][m], a[cc][m], s0);
        const double v = a[i][cc] - ((s0 + s1) + (s2 + s3));
        const double d = __shfl_sync(0xffffffffu, v, c);
        if (lane == 0 && c < nvalid && !(d > 0.0)) atomicCAS(info, 0, piv0 + c + 1);
        const double l = sqrt(d);
        if (lane == c) a[i][cc] = l;
        else if (lane > c) a[i][cc] = v / l;
        __syncwarp();
    }
    // X = L^-1: x[r][t] = -(sum_{m = t}^{r-1} L[r][m] x[m][t]) / L[r][r], lane t owns column t
    const int t = o + lane;
    for (int r = 0; r < 32; ++r) {
        const int rr = o + r;
        double s0 = 0.0, s1 = 0.0;
        int m = t;
        for (; m + 1 < rr; m += 2) {
            s0 = fma(a[rr][m], x[m][t], s0);
            s1 = fma(a[rr][m + 1], x[m + 1][t], s1);
        }
        if (m < rr) s0 = fma(a[rr][m], x[m][t], s0);
        const double lrr = a[rr][rr];
        x[rr][t] = (lane == r) ? 1.0 / lrr : (lane < r ? -(s0 + s1) / lrr : 0.0);
        __syncwarp();
    }
}

// Diagonal block of the blocked factorisation: L (in place) and L^-1 of a 64 x 64 SPD block, as a 2 x 2 recursion on
// 32 x 32 blocks -- warp 0 factors and inverts the two diagonal 32-blocks (warp_chol_inv32), the whole CTA
// does the three small products in between from shared memory:
//   L11 = chol(A11), X11 = L11^-1;  L21 = A21 X11^T;  A22 -= L21 L21^T;  L22 = chol(A22), X22 = L22^-1;  X21 = -X22 L21 X11.
// (The first version ran an unblocked right-looking loop with two block barriers per column and a serial substitution:
// 77 us per block, 16 blocks on the critical path of chol(C^uu) at d = 1024.)
__global__ void __launch_bounds__(256) potrf_diag_kernel(double* __restrict__ A, long long ld, int j0, int nb,
                                                         double* __restrict__ Linv, long long ldinv, int* info) {
    extern __shared__ double potrf_smem[];
    double (*a)[NBP] = reinterpret_cast<double (*)[NBP]>(potrf_smem);
    double (*x)[NBP] = reinterpret_cast<double (*)[NBP]>(potrf_smem + NB * NBP);
    const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6, lane = tid & 31, warp = tid >> 5;
    double* blk = A + (size_t)j0 * ld + j0;
    for (int r = ty; r < NB; r += 4) {
        a[r][tx] = (r < nb && tx < nb) ? blk[(size_t)r * ld + tx] : (r == tx ? 1.0 : 0.0);    // identity padding
        x[r][tx] = 0.0;
    }
    __syncthreads();
    if (warp == 0) warp_chol_inv32(a, x, 0, lane, j0, nb, info);          // L11, X11
    __syncthreads();
    // L21 = A21 X11^T : L21[r][c] = sum_{m <= c} A21[r][m] X11[c][m]
    {
        double out[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = tid + 256 * e, r = 32 + (idx >> 5), c = idx & 31;
            // X11 is lower triangular (zeros stored above the diagonal): a fixed trip count unrolls into independent loads
            double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 8
            for (int m = 0; m < 32; m += 2) { acc0 = fma(a[r][m], x[c][m], acc0); acc1 = fma(a[r][m + 1], x[c][m + 1], acc1); }
            out[e] = acc0 + acc1;
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = tid + 256 * e;
            a[32 + (idx >> 5)][idx & 31] = out[e];
        }
    }
    __syncthreads();
    // A22 -= L21 L21^T (lower triangle)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int idx = tid + 256 * e, r = idx >> 5, c = idx & 31;
        if (c <= r) {
            double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 8
            for (int m = 0; m < 32; m += 2) { acc0 = fma(a[32 + r][m], a[32 + c][m], acc0); acc1 = fma(a[32 + r][m + 1], a[32 + c][m + 1], acc1); }
            a[32 + r][32 + c] -= acc0 + acc1;
        }
    }
    __syncthreads();
    if (warp == 0) warp_chol_inv32(a, x, 32, lane, j0 + 32, nb - 32, info);       // L22, X22
    __syncthreads();
    // X21 = -X22 (L21 X11): T = L21 X11 into the (unused) upper-right quadrant of a, then X21 = -X22 T
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int idx = tid + 256 * e, r = idx >> 5, c = idx & 31;
        double acc0 = 0.0, acc1 = 0.0;                  // X11 is lower triangular: terms with m < c are exact zeros
#pragma unroll 8
        for (int m = 0; m < 32; m += 2) { acc0 = fma(a[32 + r][m], x[m][c], acc0); acc1 = fma(a[32 + r][m + 1], x[m + 1][c], acc1); }
        a[r][32 + c] = acc0 + acc1;
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int idx = tid + 256 * e, r = idx >> 5, c = idx & 31;
        double acc0 = 0.0, acc1 = 0.0;                  // X22 lower triangular: terms with m > r are exact zeros
#pragma unroll 8
        for (int m = 0; m < 32; m += 2) { acc0 = fma(x[32 + r][32 + m], a[m][32 + c], acc0); acc1 = fma(x[32 + r][32 + m + 1], a[m + 1][32 + c], acc1); }
        x[32 + r][c] = -(acc0 + acc1);
    }
    __syncthreads();
    double* inv = Linv + (size_t)j0 * ldinv;
    for (int r = ty; r < nb; r += 4) {
        if (tx < nb) {
            blk[(size_t)r * ld + tx] = (tx <= r) ? a[r][tx] : 0.0;
            inv[(size_t)r * ldinv + tx] = (tx <= r) ? x[r][tx] : 0.0;
        }
    }
}

__global__ void __launch_bounds__(256) zero_upper_kernel(double* __restrict__ A, long long ld, int n) {
    const int j = blockIdx.x * 256 + threadIdx.x, i = blockIdx.y;
    if (j < n && j > i) A[(size_t)i * ld + j] = 0.0;
}

__global__ void __launch_bounds__(256) set_identity_kernel(double* __restrict__ A, long long ld, int n) {
    const int j = blockIdx.x * 256 + threadIdx.x, i = blockIdx.y;
    if (j < n) A[(size_t)i * ld + j] = (i == j) ? 1.0 : 0.0;
}

int set_identity(cudaStream_t st, double* A, int64_t ld, int64_t n) {
    dim3 grid((unsigned)ceil_div(n, 256), (unsigned)n);
    set_identity_kernel<<<grid, 256, 0, st>>>(A, ld, (int)n);
    CES_LAUNCHED(1);
    return CES_OK;
}

int potrf_lower(cudaStream_t st, double* A, int64_t ld, int64_t n, double* Linv, int64_t ldinv, int* info_dev) {
    if (n < 1) return CES_OK;
    constexpr int kDiagSmem = 2 * NB * NBP * (int)sizeof(double);
    static bool attr_set[kMaxDevices] = {};
    const int slot = device_slot();
    if (!attr_set[slot]) {
        CES_CUDA(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDiagSmem));
        attr_set[slot] = true;
    }
    for (int64_t j0 = 0; j0 < n; j0 += NB) {
        const int nb = (int)((n - j0) < NB ? (n - j0) : NB);
        potrf_diag_kernel<<<1, 256, kDiagSmem, st>>>(A, ld, (int)j0, nb, Linv, ldinv, info_dev);
        CES_LAUNCHED(1);
        const int64_t n2 = n - j0 - nb;
        if (n2 <= 0) break;
        double* A21 = A + (size_t)(j0 + nb) * ld + j0;
        // panel: L21 = A21 * inv(L11)^T   (in place; one tile column, so each CTA reads its rows fully before storing)
        GemmCall pc;
        pc.a_mode = A_MK; pc.b_mode = B_NK;
        pc.M = (int)n2; pc.N = nb; pc.K = nb;
        pc.A = A21; pc.lda = ld;
        pc.B = Linv + (size_t)j0 * ldinv; pc.ldb = ldinv;
        pc.C = A21; pc.ldc = ld;
        CES_TRY(gemm(st, pc));
        // trailing update: A22 -= L21 L21^T  (lower tiles only)
        GemmCall tc;
        tc.a_mode = A_MK; tc.b_mode = B_NK;
        tc.M = (int)n2; tc.N = (int)n2; tc.K = nb;
        tc.A = A21; tc.lda = ld;
        tc.B = A21; tc.ldb = ld;
        tc.C = A + (size_t)(j0 + nb) * ld + (j0 + nb); tc.ldc = ld;
        tc.alpha = -1.0; tc.beta = 1.0;
        tc.flags = GEMM_C_LOWER_ONLY;
        CES_TRY(gemm(st, tc));
    }
    dim3 grid((unsigned)ceil_div(n, 256), (unsigned)n);
    zero_upper_kernel<<<grid, 256, 0, st>>>(A, ld, (int)n);
    CES_LAUNCHED(1);
    return CES_OK;
}

int trsm_lower(cudaStream_t st, const double* L, int64_t ld, int64_t n, const double* Linv, int64_t ldinv,
               double* B, int64_t ldb, int64_t nrhs, bool transpose) {
    const int64_t nblk = ceil_div(n, NB);
    for (int64_t bi = 0; bi < nblk; ++bi) {
        const int64_t b = transpose ? nblk - 1 - bi : bi;
        const int64_t i0 = b * NB;
        const int nb = (int)((n - i0) < NB ? (n - i0) : NB);
        double* Bi = B + (size_t)i0 * ldb;
        GemmCall u;   // subtract the already-solved blocks
        u.N = (int)nrhs; u.M = nb;
        u.C = Bi; u.ldc = ldb; u.alpha = -1.0; u.beta = 1.0;
        u.b_mode = B_KN; u.ldb = ldb;
        bool have = false;
        if (!transpose && i0 > 0) {
            u.a_mode = A_MK; u.A = L + (size_t)i0 * ld; u.lda = ld; u.K = (int)i0; u.B = B;
            have = true;
        } else if (transpose && i0 + nb < n) {
            const int64_t i1 = i0 + nb;
            u.a_mode = A_KM; u.A = L + (size_t)i1 * ld + i0; u.lda = ld; u.K = (int)(n - i1); u.B = B + (size_t)i1 * ldb;
            have = true;
        }
        if (have) CES_TRY(gemm(st, u));
        GemmCall d;   // multiply by the inverse of the diagonal block (in place, single tile row)
        d.a_mode = transpose ? A_KM : A_MK;
        d.b_mode = B_KN;
        d.M = nb; d.N = (int)nrhs; d.K = nb;
        d.A = Linv + (size_t)i0 * ldinv; d.lda = ldinv;
        d.B = Bi; d.ldb = ldb;
        d.C = Bi; d.ldc = ldb;
        CES_TRY(gemm(st, d));
    }
    return CES_OK;
}

int spd_inverse_from_factor(cudaStream_t st, const double* L, int64_t ld, int64_t n, const double* Linv, int64_t ldinv,
                            double* Ainv, int64_t ldo) {
    CES_TRY(set_identity(st, Ainv, ldo, n));
    CES_TRY(trsm_lower(st, L, ld, n, Linv, ldinv, Ainv, ldo, n, false));
    CES_TRY(trsm_lower(st, L, ld, n, Linv, ldinv, Ainv, ldo, n, true));
    return CES_OK;
}

}  // namespace ces
