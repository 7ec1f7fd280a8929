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

__global__ void __launch_bounds__(256) potrf_diag_kernel(double* __restrict__ A, long long ld, int j0, int nb,
                                                         double* __restrict__ Linv, long long ldinv, int* info) {
    extern __shared__ double potrf_smem[];
    double (*a)[NBP] = reinterpret_cast<double (*)[NBP]>(potrf_smem);
    double (*x)[NBP] = reinterpret_cast<double (*)[NBP]>(potrf_smem + NB * NBP);
    double* dl = potrf_smem + 2 * NB * NBP;
    const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
    double* blk = A + (size_t)j0 * ld + j0;
    for (int r = ty; r < NB; r += 4) {
        a[r][tx] = (r < nb && tx < nb) ? blk[(size_t)r * ld + tx] : (r == tx ? 1.0 : 0.0);
        x[r][tx] = 0.0;
    }
    for (int c = 0; c < nb; ++c) {
        __syncthreads();
        const double d = a[c][c];
        if (!(d > 0.0)) {
            if (tid == 0) atomicCAS(info, 0, j0 + c + 1);
        }
        const double l = sqrt(d);
        if (tid == c) dl[c] = l;
        if (tid > c && tid < nb) a[tid][c] = a[tid][c] / l;
        __syncthreads();
        if (tx > c && tx < nb) {
            const double lc = a[tx][c];
            for (int r = ty; r < nb; r += 4)
                if (r >= tx) a[r][tx] -= a[r][c] * lc;
        }
    }
    __syncthreads();
    if (tid < nb) a[tid][tid] = dl[tid];
    __syncthreads();
    // Inverse of the lower-triangular block by forward substitution, column t of X = L^-1 owned by the four lanes
    // 4t .. 4t+3 of one warp: they split the inner product over m, combine with two shuffles, lane 0 stores.
    {
        const int t = tid >> 2, part = tid & 3;
        for (int i = 0; i < nb; ++i) {
            double s = 0.0;
            if (t < nb && i > t)
                for (int m = t + part; m < i; m += 4) s -= a[i][m] * x[m][t];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (part == 0 && t < nb) x[i][t] = (i > t) ? s / a[i][i] : (i == t ? 1.0 / a[i][i] : 0.0);
            __syncwarp();
        }
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
    constexpr int kDiagSmem = (2 * NB * NBP + NB) * (int)sizeof(double);
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
