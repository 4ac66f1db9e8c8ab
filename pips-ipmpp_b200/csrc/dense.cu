// Dense FP64 kernels of the root stage and of the Schur-term accumulation (sm_100a).
//
//  dgemm_nt_kernel   C (-)= A * B^T on the FP64 tensor pipe (mma.sync m8n8k4 f64 = DMMA; tcgen05 has no f64 kind).
//                    Operands are K-contiguous row-major matrices described by TMA tensor maps
//                    (cuTensorMapEncodeTiled, box 16 x 128 doubles, SWIZZLE_128B); one elected thread keeps a 4-stage
//                    ring of cp.async.bulk.tensor.2d loads in flight (full / empty mbarriers), 8 warps consume the
//                    stages with conflict-light swizzled fragment loads; 128 x 128 CTA tile, 64 x 32 warp tile.
//                    Epilogues: trailing update of the root LDL^T (lower tiles, read-modify-write), panel
//                    W21 = A21 X^T / L21 = W21 D^-1, and the Schur-term update  SC[map i][map j] -= P Q^T.
//  ldlt_diag_kernel  static-pivot LDL^T of one 128 x 128 diagonal block in registers (one CTA, 1024 threads), together
//                    with X = L11^-1 so that the panel below becomes a GEMM.
//  dense_trisolve_kernel  forward / D^-1 / backward solve with the blocked factors, one persistent CTA per 64-row block,
//                    data-flow synchronised through epoch flags in global memory.
// Replaces DeSymIndefSolver::matrixChanged / solve (SOL/DenseSymmetricIndefinitSolver/DeSymIndefSolver.C:56-130:
// LAPACK dsytrf / dsytrs) for quasi-definite Schur complements; the Bunch-Kaufman fallback stays in kernels.cu.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <string>

#include "common.h"
#include "kernels.cuh"

namespace pipsgpu {

// =====================================================================================================
// tensor maps
// =====================================================================================================
static_assert(sizeof(TMap) == sizeof(CUtensorMap), "TMap must mirror CUtensorMap");

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
            throw Error(PIPSGPU_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

// rows x cols doubles, row-major with leading dimension ld (doubles, even); box = 16 columns (128 bytes) x GEMM_TILE rows
void make_tmap(TMap* out, const double* base, int64_t rows, int64_t cols, int64_t ld) {
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld & 1))
        throw Error(PIPSGPU_ERR_INVALID, "make_tmap: base must be 16-byte aligned and the leading dimension even");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)GEMM_KS, (cuuint32_t)GEMM_TILE};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode_tiled_fn()(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2,
                                         const_cast<double*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(PIPSGPU_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
}

// =====================================================================================================
// PTX helpers
// =====================================================================================================
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 2-D tiled TMA load global -> shared (coordinates: c0 = column / innermost, c1 = row), completion on an mbarrier
__device__ __forceinline__ void tma_load_2d(void* dst, const TMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

}  // namespace

// =====================================================================================================
// C (-)= A * B^T, A: M x K, B: N x K (both row-major, K contiguous), FP64 tensor pipe
// =====================================================================================================
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_STAGE_BYTES = 2 * GEMM_TILE * GEMM_KS * 8;      // A tile + B tile = 32 KB
constexpr int GEMM_SMEM = GEMM_STAGES * GEMM_STAGE_BYTES + 1024;   // + slack for the 1024-byte alignment of the swizzle

template <int EPI>
__global__ void __launch_bounds__(256, 1) dgemm_nt_kernel(const __grid_constant__ TMap tmA, const __grid_constant__ TMap tmB,
                                                          const GemmArgs g) {
    extern __shared__ uint8_t gemm_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[GEMM_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[GEMM_STAGES];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gemm_smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // ---- tile of this CTA ----
    int bi, bj;
    if (g.lower) {      // linear index over the lower triangle of an m_tiles x m_tiles tile grid, row by row
        const int t = blockIdx.x;
        bi = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        while (bi * (bi + 1) / 2 > t) --bi;
        while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
        bj = t - bi * (bi + 1) / 2;
    } else {
        bi = blockIdx.x / g.n_tiles;
        bj = blockIdx.x - bi * g.n_tiles;
    }
    const int arow = g.a_row0 + bi * GEMM_TILE, brow = g.b_row0 + bj * GEMM_TILE;
    const int nk = (g.K + GEMM_KS - 1) / GEMM_KS;
    if (tid == 0) {
        for (int s = 0; s < GEMM_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int kt) {      // thread 0 only
        const int s = kt % GEMM_STAGES;
        uint8_t* st = smem + (size_t)s * GEMM_STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], GEMM_STAGE_BYTES);
        tma_load_2d(st, &tmA, g.a_col0 + kt * GEMM_KS, arow, &full_bar[s]);
        tma_load_2d(st + GEMM_TILE * GEMM_KS * 8, &tmB, g.b_col0 + kt * GEMM_KS, brow, &full_bar[s]);
    };
    if (tid == 0)
        for (int kt = 0; kt < GEMM_STAGES && kt < nk; ++kt) issue(kt);

    const int wr = (warp >> 2) * 64, wc = (warp & 3) * 32;      // warp tile 64 x 32
    const int lr = lane >> 2, lk = lane & 3;
    double acc[8][4][2];
#pragma unroll
    for (int x = 0; x < 8; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
    // swizzled element (r, k) of a 128 x 16 tile: r * 128 + (((k >> 1) ^ (r & 7)) << 4) + (k & 1) * 8 ; r & 7 == lr here
    const uint32_t a_base = (uint32_t)((wr + lr) * 128 + (lk & 1) * 8);
    const uint32_t b_base = (uint32_t)(GEMM_TILE * GEMM_KS * 8 + (wc + lr) * 128 + (lk & 1) * 8);
    for (int kt = 0; kt < nk; ++kt) {
        const int s = kt % GEMM_STAGES;
        mbar_wait(&full_bar[s], (uint32_t)((kt / GEMM_STAGES) & 1));
        const uint8_t* st = smem + (size_t)s * GEMM_STAGE_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const uint32_t ch = (uint32_t)((((kk << 1) | (lk >> 1)) ^ lr) << 4);
            double af[8], bf[4];
#pragma unroll
            for (int x = 0; x < 8; ++x) af[x] = *reinterpret_cast<const double*>(st + a_base + ch + x * 1024);
#pragma unroll
            for (int y = 0; y < 4; ++y) bf[y] = *reinterpret_cast<const double*>(st + b_base + ch + y * 1024);
#pragma unroll
            for (int x = 0; x < 8; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) dmma(acc[x][y][0], acc[x][y][1], af[x], bf[y]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
        // refill with a lag of one stage: the stage consumed in the PREVIOUS iteration has almost surely been released by
        // every warp, so the elected thread does not stall its own warp's tensor work on the slowest warp
        if (tid == 0 && kt >= 1 && kt - 1 + GEMM_STAGES < nk) {
            const int ps = (kt - 1) % GEMM_STAGES;
            mbar_wait(&empty_bar[ps], (uint32_t)(((kt - 1) / GEMM_STAGES) & 1));
            issue(kt - 1 + GEMM_STAGES);
        }
    }
    // ---- epilogue ----
    const int ci0 = g.c_row0 + bi * GEMM_TILE + wr + lr;       // + x * 8
    const int cj0 = g.c_col0 + bj * GEMM_TILE + wc + lk * 2;   // + y * 8 (+ e)
    if (EPI == GEMM_EPI_TRAIL) {
        // C(i,j) -= acc for j <= i (lower triangle of the trailing matrix), row-major read-modify-write
#pragma unroll
        for (int x = 0; x < 8; ++x) {
            const int gi = ci0 + x * 8;
            if (gi >= g.c_rows) continue;
            double* crow = g.C + (int64_t)gi * g.ldc;
            double2 v[4];
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const int gj = cj0 + y * 8;
                if (gj + 1 <= gi) v[y] = *reinterpret_cast<const double2*>(crow + gj);
                else if (gj <= gi) { v[y].x = crow[gj]; v[y].y = 0.0; }
            }
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const int gj = cj0 + y * 8;
                if (gj + 1 <= gi) *reinterpret_cast<double2*>(crow + gj) = make_double2(v[y].x - acc[x][y][0], v[y].y - acc[x][y][1]);
                else if (gj <= gi) crow[gj] = v[y].x - acc[x][y][0];
            }
        }
    } else if (EPI == GEMM_EPI_PANEL) {
        // W21 = acc -> C2 (row-major, ldc2, column index local to the panel); L21 = W21 * Dinv -> C (the factor, in place)
#pragma unroll
        for (int x = 0; x < 8; ++x) {
            const int gi = ci0 + x * 8;
            if (gi >= g.c_rows) continue;
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const int jl = bj * GEMM_TILE + wc + lk * 2 + y * 8;      // column inside the panel
                if (jl + 1 < g.c_cols) {
                    *reinterpret_cast<double2*>(g.C2 + (int64_t)gi * g.ldc2 + jl) = make_double2(acc[x][y][0], acc[x][y][1]);
                    const double2 sc = *reinterpret_cast<const double2*>(g.scale + jl);
                    *reinterpret_cast<double2*>(g.C + (int64_t)gi * g.ldc + g.c_col0 + jl) =
                        make_double2(acc[x][y][0] * sc.x, acc[x][y][1] * sc.y);
                } else if (jl < g.c_cols) {
                    g.C2[(int64_t)gi * g.ldc2 + jl] = acc[x][y][0];
                    g.C[(int64_t)gi * g.ldc + g.c_col0 + jl] = acc[x][y][0] * g.scale[jl];
                }
            }
        }
    } else {
        // Schur term: SC[map[i]][map[j]] -= acc for packed rows j <= i (both orientations land in the lower triangle of SC)
#pragma unroll
        for (int x = 0; x < 8; ++x) {
            const int pi = ci0 + x * 8;
            if (pi >= g.c_rows) continue;
            const int mi = g.map[pi];
#pragma unroll
            for (int y = 0; y < 4; ++y)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int pj = cj0 + y * 8 + e;
                    if (pj > pi) continue;
                    const double v = acc[x][y][e];
                    if (v == 0.0) continue;
                    const int mj = g.map[pj];
                    const int64_t off = (mi >= mj) ? (int64_t)mi * g.ldc + mj : (int64_t)mj * g.ldc + mi;
                    atomicAdd(g.C + off, -v);
                }
        }
    }
}

cudaError_t init_dense_attributes();

void launch_dgemm_nt(int epi, const TMap& tmA, const TMap& tmB, const GemmArgs& g, cudaStream_t s) {
    const int ntiles = g.lower ? g.m_tiles * (g.m_tiles + 1) / 2 : g.m_tiles * g.n_tiles;
    if (ntiles <= 0 || g.K <= 0) return;
    if (epi == GEMM_EPI_TRAIL) dgemm_nt_kernel<GEMM_EPI_TRAIL><<<ntiles, 256, GEMM_SMEM, s>>>(tmA, tmB, g);
    else if (epi == GEMM_EPI_PANEL) dgemm_nt_kernel<GEMM_EPI_PANEL><<<ntiles, 256, GEMM_SMEM, s>>>(tmA, tmB, g);
    else dgemm_nt_kernel<GEMM_EPI_SCHUR><<<ntiles, 256, GEMM_SMEM, s>>>(tmA, tmB, g);
    count_launch("dgemm_nt_kernel");
}

// =====================================================================================================
// 128 x 128 diagonal block: static-pivot LDL^T and X = L11^-1 in ONE CTA of 8 warps, blocked 4 x 4 by 32.
//   The block lives in shared memory as ten 32 x 32 lower blocks (leading dimension 33). Step j = 0..3:
//     (a) warp 0 factorises the diagonal block (j,j): lane = row, the row in registers, column k broadcast through a
//         double-buffered shared-memory line (one __syncwarp per column, no CTA barrier on the 128-step chain);
//     (b) rows below: W = A L_jj^-T by forward substitution, one thread per row (L_jj reads are warp broadcasts),
//         L = W D^-1; W kept in a scratch tile;
//     (c) trailing blocks (i,k), j < k <= i:  A_ik -= L_ij W_kj^T, all 256 threads, 2 x 2 register blocks.
//   Then X = L11^-1: the four diagonal blocks by four warps (lane = column, forward substitution with broadcast reads),
//   the off-diagonal blocks level by level (distance 1, 2, 3 from the diagonal):  X_ij = -X_ii sum_k L_ik X_kj.
//   Pivots with the wrong sign or below eps_rel * (largest entry of the pivot's original row) are replaced by +-eps
//   (the static order is x0 first: expected signs + ... + - ... -); inertia counts the COMPUTED pivots.
//   The previous version kept the whole block in the registers of 1024 threads and paid two CTA barriers plus an FP64
//   division per column on a 255-step chain (230 us per panel); this one has 128 warp-level steps and ~25 barriers.
// =====================================================================================================
constexpr int DB = 32;                       // sub-block size
constexpr int DBL = DB + 1;                  // leading dimension of a sub-block in shared memory
constexpr int DNB = DENSE_PANEL / DB;        // 4
constexpr int DBLK = DB * DBL;               // doubles per sub-block
constexpr int DNLOW = DNB * (DNB + 1) / 2;   // 10 lower blocks
__device__ __forceinline__ int dblk(int i, int j) { return (i * (i + 1) / 2 + j) * DBLK; }

__global__ void __launch_bounds__(256, 1) ldlt_diag_kernel(double* __restrict__ A, int64_t ld, int kb, int bw, int n0, double eps_rel,
                                                           const double* __restrict__ scale, double* __restrict__ D,
                                                           double* __restrict__ Dinv, double* __restrict__ X, int* __restrict__ stats) {
    extern __shared__ double dg_smem[];
    double* sL = dg_smem;                         // A -> L (unit lower; W on the way), lower blocks
    double* sX = sL + DNLOW * DBLK;               // X = L^-1, lower blocks
    double* sW = sX + DNLOW * DBLK;               // W = L D of the rows below the current diagonal block, [96][33]
    __shared__ double s_col[2][DB];
    __shared__ double s_dinv[DENSE_PANEL], s_d[DENSE_PANEL], s_eps[DENSE_PANEL];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < DENSE_PANEL) s_eps[tid] = (tid < bw) ? eps_rel * fmax(scale[kb + tid], 1e-300) : 0.0;
    // ---- load the lower blocks (rows / columns beyond bw: identity). All 40 loads of a thread are issued before the first
    // one is consumed: the block comes from HBM / L2 and sits on the critical path of the whole root factorisation ----
    {
        double v[DNLOW * 4];
#pragma unroll
        for (int b = 0; b < DNLOW; ++b) {
            const int bi = (b >= 6) ? 3 : (b >= 3) ? 2 : (b >= 1) ? 1 : 0;
            const int bj = b - bi * (bi + 1) / 2;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int idx = tid + q * 256;
                const int r = idx >> 5, c = idx & 31;
                const int gi = bi * DB + r, gj = bj * DB + c;
                v[b * 4 + q] = (gi < bw && gj <= gi) ? A[(int64_t)(kb + gi) * ld + kb + gj] : ((gi == gj) ? 1.0 : 0.0);
            }
        }
#pragma unroll
        for (int b = 0; b < DNLOW; ++b)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int idx = tid + q * 256;
                sL[b * DBLK + (idx >> 5) * DBL + (idx & 31)] = v[b * 4 + q];
            }
    }
    __syncthreads();
    int npos = 0, nneg = 0, nzero = 0, npert = 0;      // warp 0, uniform over its lanes
    for (int j = 0; j < DNB; ++j) {
        double* Ljj = sL + dblk(j, j);
        // ---------------- (a) diagonal block, warp 0 ----------------
        if (warp == 0) {
            double a[DB];
#pragma unroll
            for (int c = 0; c < DB; ++c) a[c] = Ljj[lane * DBL + c];
#pragma unroll
            for (int k = 0; k < DB; ++k) {
                const int p = k & 1;
                const double colk = a[k];                 // A(lane, k) after the updates of columns < k
                s_col[p][lane] = colk;
                __syncwarp();
                double d = s_col[p][k];
                const int gk = j * DB + k;
                if (gk < bw) {
                    const double sg = (kb + gk < n0) ? 1.0 : -1.0;
                    const double eps = s_eps[gk];
                    if (d >= eps) ++npos; else if (d <= -eps) ++nneg; else ++nzero;      // inertia of the computed pivot
                    if (!(sg * d >= eps)) { d = sg * eps; ++npert; }
                }
                const double dinv = __drcp_rn(d);        // correctly rounded reciprocal, shorter than the general division
                const double l = colk * dinv;
                if (lane == k) { s_d[gk] = d; s_dinv[gk] = dinv; }
#pragma unroll
                for (int c = k + 1; c < DB; ++c) {
                    const double wc = s_col[p][c];        // A(c, k) = d_k L(c, k)
                    if (lane >= c) a[c] = fma(-l, wc, a[c]);
                }
                if (lane > k) a[k] = l;
            }
#pragma unroll
            for (int c = 0; c < DB; ++c) Ljj[lane * DBL + c] = (c < lane) ? a[c] : ((c == lane) ? 1.0 : 0.0);
        }
        __syncthreads();
        const int nbelow = (DNB - 1 - j) * DB;
        if (nbelow == 0) break;
        // ---------------- (b) rows below: W L_jj^T = A,  L = W D^-1 ----------------
        if (tid < nbelow) {
            const int i = j + 1 + (tid >> 5);
            double* Aij = sL + dblk(i, j) + (tid & 31) * DBL;
            double w[DB];
#pragma unroll
            for (int c = 0; c < DB; ++c) w[c] = Aij[c];
#pragma unroll
            for (int c = 1; c < DB; ++c) {
                double acc = w[c];
#pragma unroll
                for (int q = 0; q < c; ++q) acc = fma(-w[q], Ljj[c * DBL + q], acc);
                w[c] = acc;
            }
#pragma unroll
            for (int c = 0; c < DB; ++c) {
                sW[tid * DBL + c] = w[c];
                Aij[c] = w[c] * s_dinv[j * DB + c];
            }
        }
        __syncthreads();
        // ---------------- (c) trailing blocks (i, k), j < k <= i:  A_ik -= L_ij W_kj^T ----------------
        {
            const int r2 = (tid >> 4) * 2, c2 = (tid & 15) * 2;
            for (int i = j + 1; i < DNB; ++i)
                for (int k = j + 1; k <= i; ++k) {
                    const double* Lij = sL + dblk(i, j);
                    const double* Wk = sW + (k - j - 1) * DB * DBL;
                    double* Aik = sL + dblk(i, k);
                    double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
#pragma unroll 8
                    for (int q = 0; q < DB; ++q) {
                        const double l0 = Lij[r2 * DBL + q], l1 = Lij[(r2 + 1) * DBL + q];
                        const double w0 = Wk[c2 * DBL + q], w1 = Wk[(c2 + 1) * DBL + q];
                        c00 = fma(l0, w0, c00); c01 = fma(l0, w1, c01);
                        c10 = fma(l1, w0, c10); c11 = fma(l1, w1, c11);
                    }
                    Aik[r2 * DBL + c2] -= c00; Aik[r2 * DBL + c2 + 1] -= c01;
                    Aik[(r2 + 1) * DBL + c2] -= c10; Aik[(r2 + 1) * DBL + c2 + 1] -= c11;
                }
        }
        __syncthreads();
    }
    // ---------------- outputs of the factorisation; X diagonal blocks ----------------
    if (tid == 0 && stats) {
        atomicAdd(stats + 0, npos); atomicAdd(stats + 1, nneg); atomicAdd(stats + 2, nzero); atomicAdd(stats + 3, npert);
    }
    if (tid < bw) { Dinv[kb + tid] = s_dinv[tid]; D[kb + tid] = s_d[tid]; }
    if (warp < DNB) {
        // X_jj = L_jj^-1, lane = column c of X: x(r) = -sum_{c <= k < r} L(r,k) x(k), x(c) = 1 (L reads are broadcasts)
        const double* Ljj = sL + dblk(warp, warp);
        double x[DB];
#pragma unroll
        for (int r = 0; r < DB; ++r) {
            double acc = (r == lane) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < r; ++k) acc = fma(-Ljj[r * DBL + k], x[k], acc);     // x(k) = 0 for k < lane
            x[r] = acc;
        }
        double* Xjj = sX + dblk(warp, warp);
#pragma unroll
        for (int r = 0; r < DB; ++r) Xjj[r * DBL + lane] = x[r];
    } else {
        // the strictly lower part of L11 goes back into the factor
        for (int b = 0; b < DNLOW; ++b) {
            int bi = 0;
            while ((bi + 1) * (bi + 2) / 2 <= b) ++bi;
            const int bj = b - bi * (bi + 1) / 2;
            for (int idx = tid - DNB * 32; idx < DB * DB; idx += 256 - DNB * 32) {
                const int r = idx >> 5, c = idx & 31;
                const int gi = bi * DB + r, gj = bj * DB + c;
                if (gi < bw && gj < gi) A[(int64_t)(kb + gi) * ld + kb + gj] = sL[b * DBLK + r * DBL + c];
            }
        }
    }
    __syncthreads();
    // ---------------- off-diagonal blocks of X, by distance from the diagonal ----------------
    {
        const int r2 = (tid >> 4) * 2, c2 = (tid & 15) * 2;
#pragma unroll
        for (int dist = 1; dist < DNB; ++dist) {
            // T_ij = sum_{k = j}^{i-1} L_ik X_kj  into the X slot (i,j), j = i - dist
            for (int i = dist; i < DNB; ++i) {
                const int j = i - dist;
                double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
                for (int k = j; k < i; ++k) {
                    const double* Lik = sL + dblk(i, k);
                    const double* Xkj = sX + dblk(k, j);
#pragma unroll 8
                    for (int q = 0; q < DB; ++q) {
                        const double l0 = Lik[r2 * DBL + q], l1 = Lik[(r2 + 1) * DBL + q];
                        const double x0 = Xkj[q * DBL + c2], x1 = Xkj[q * DBL + c2 + 1];
                        c00 = fma(l0, x0, c00); c01 = fma(l0, x1, c01);
                        c10 = fma(l1, x0, c10); c11 = fma(l1, x1, c11);
                    }
                }
                double* T = sX + dblk(i, j);
                T[r2 * DBL + c2] = c00; T[r2 * DBL + c2 + 1] = c01;
                T[(r2 + 1) * DBL + c2] = c10; T[(r2 + 1) * DBL + c2 + 1] = c11;
            }
            __syncthreads();
            // X_ij = -X_ii T_ij (in place through registers)
            double o[DNB][4];
#pragma unroll
            for (int i = dist; i < DNB; ++i) {
                const double* Xii = sX + dblk(i, i);
                const double* T = sX + dblk(i, i - dist);
                double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
#pragma unroll 8
                for (int q = 0; q < DB; ++q) {
                    const double a0 = Xii[r2 * DBL + q], a1 = Xii[(r2 + 1) * DBL + q];
                    const double t0 = T[q * DBL + c2], t1 = T[q * DBL + c2 + 1];
                    c00 = fma(a0, t0, c00); c01 = fma(a0, t1, c01);
                    c10 = fma(a1, t0, c10); c11 = fma(a1, t1, c11);
                }
                o[i][0] = -c00; o[i][1] = -c01; o[i][2] = -c10; o[i][3] = -c11;
            }
            __syncthreads();
#pragma unroll
            for (int i = dist; i < DNB; ++i) {
                double* Xo = sX + dblk(i, i - dist);
                Xo[r2 * DBL + c2] = o[i][0]; Xo[r2 * DBL + c2 + 1] = o[i][1];
                Xo[(r2 + 1) * DBL + c2] = o[i][2]; Xo[(r2 + 1) * DBL + c2 + 1] = o[i][3];
            }
            __syncthreads();
        }
    }
    // ---------------- X to global (row-major 128 x 128, zero upper part) ----------------
    for (int idx = tid; idx < DENSE_PANEL * DENSE_PANEL; idx += 256) {
        const int gi = idx >> 7, gj = idx & 127;
        double v = 0.0;
        if (gj <= gi) v = sX[dblk(gi >> 5, gj >> 5) + (gi & 31) * DBL + (gj & 31)];
        X[idx] = v;
    }
}

constexpr int DIAG_SMEM = (2 * DNLOW * DBLK + (DNB - 1) * DB * DBL) * 8;
void launch_ldlt_diag(double* A, int64_t ld, int kb, int bw, int n0, double eps_rel, const double* scale, double* D, double* Dinv,
                      double* X, int* stats, cudaStream_t s) {
    ldlt_diag_kernel<<<1, 256, DIAG_SMEM, s>>>(A, ld, kb, bw, n0, eps_rel, scale, D, Dinv, X, stats);
    count_launch("ldlt_diag_kernel");
}

// =====================================================================================================
// blocked triangular solves with the static-order factors  A = L D L^T  (unit lower L in the strict lower triangle of
// the row-major factor, X_b = inverse of every 128 x 128 diagonal block of L, D^-1 diagonal). One persistent launch per
// right-hand side: CTA c owns the 64-row blocks c, c + G, ... ; block i of the forward sweep needs y_j of every j < i,
// published through flagF[j] (= epoch when done), the backward sweep x_j of every j > i through flagB[j]. The last
// link of each chain (L(i,i-1) resp. L(i+1,i), and the 64 x 64 diagonal block of X) is prefetched into shared memory
// before the wait, so that a chain step costs one flag round trip plus two small shared-memory products.
// =====================================================================================================
constexpr int TS_B = 64;
constexpr int TS_SMEM = 2 * TS_B * (TS_B + 1) * 8;
__global__ void __launch_bounds__(256, 1) dense_trisolve_kernel(const double* __restrict__ F, int64_t ld, int n,
                                                                const double* __restrict__ Xall, const double* __restrict__ Dinv,
                                                                double* __restrict__ x, double* __restrict__ y, int* flagF,
                                                                int* flagB, int epoch) {
    extern __shared__ double ts_smem[];
    double (*sL)[TS_B + 1] = reinterpret_cast<double (*)[TS_B + 1]>(ts_smem);
    double (*sX)[TS_B + 1] = sL + TS_B;
    __shared__ double sv[TS_B];       // y_j / x_j of the block being applied
    __shared__ double sacc[TS_B];
    __shared__ double sred[4][TS_B];
    __shared__ int s_ready[2];
    const int tid = threadIdx.x;
    const int nb = (n + TS_B - 1) / TS_B;
    const int G = gridDim.x;
    auto load_X = [&](int bi) {      // 64 x 64 diagonal sub-block of the 128 x 128 inverse block
        const double* Xb = Xall + (int64_t)(bi >> 1) * DENSE_PANEL * DENSE_PANEL + (int64_t)(bi & 1) * (TS_B * DENSE_PANEL + TS_B);
        for (int idx = tid; idx < TS_B * TS_B; idx += 256) {
            const int r = idx >> 6, c = idx & 63;
            sX[r][c] = Xb[(int64_t)r * DENSE_PANEL + c];
        }
    };
    // ---------------- forward: L y = b ----------------
    for (int bi = blockIdx.x; bi < nb; bi += G) {
        const int r0 = bi * TS_B, nr = min(TS_B, n - r0);
        __syncthreads();
        load_X(bi);
        if (bi > 0)
            for (int idx = tid; idx < TS_B * TS_B; idx += 256) {
                const int r = idx >> 6, c = idx & 63;
                sL[r][c] = (r < nr) ? F[(int64_t)(r0 + r) * ld + (r0 - TS_B) + c] : 0.0;
            }
        // rows of this block against y_0 .. y_{bi-2}: thread = (row, quarter of the 64 columns)
        const int row = tid >> 2, q = tid & 3;
        double ps = 0.0;
        int ready = 0, pp = 0;       // blocks [0, ready) are known to be done (block j done implies every block below it)
        for (int j = 0; j + 1 < bi; ++j) {
            if (j >= ready) {
                if (tid == 0) {
                    int r = j + 1;
                    if (ld_acquire(flagF + bi - 2) == epoch) r = bi - 1;
                    else while (ld_acquire(flagF + j) != epoch) __nanosleep(32);
                    s_ready[pp] = r;
                }
                __syncthreads();
                ready = s_ready[pp];
                pp ^= 1;
            }
            if (row < nr) {
                const double* lp = F + (int64_t)(r0 + row) * ld + j * TS_B + q * 16;
                const double* yp = y + j * TS_B + q * 16;
#pragma unroll
                for (int c = 0; c < 16; c += 2) {
                    const double2 l2 = *reinterpret_cast<const double2*>(lp + c);
                    const double2 y2 = __ldcg(reinterpret_cast<const double2*>(yp + c));
                    ps = fma(l2.x, y2.x, ps);
                    ps = fma(l2.y, y2.y, ps);
                }
            }
        }
        if (bi > 0) {
            if (tid == 0) while (ld_acquire(flagF + bi - 1) != epoch) __nanosleep(32);
            __syncthreads();
            if (tid < TS_B) sv[tid] = __ldcg(y + (bi - 1) * TS_B + tid);
            __syncthreads();
#pragma unroll
            for (int c = 0; c < 16; ++c) ps = fma(sL[row][q * 16 + c], sv[q * 16 + c], ps);
        }
        ps += __shfl_xor_sync(0xffffffffu, ps, 1);
        ps += __shfl_xor_sync(0xffffffffu, ps, 2);
        if (q == 0) sacc[row] = (row < nr) ? x[r0 + row] - ps : 0.0;
        __syncthreads();
        // y_i = X_ii * acc (lower triangular), z = D^-1 y kept in x for the backward sweep
        {
            double t = 0.0;
            for (int c = q; c <= row; c += 4) t = fma(sX[row][c], sacc[c], t);
            t += __shfl_xor_sync(0xffffffffu, t, 1);
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            if (q == 0 && row < nr) {
                y[r0 + row] = t;
                x[r0 + row] = t * Dinv[r0 + row];
            }
        }
        __syncthreads();
        if (tid == 0) { __threadfence(); st_release(flagF + bi, epoch); }
    }
    // ---------------- backward: L^T x = z ----------------
    int last = -1;
    for (int bi = blockIdx.x; bi < nb; bi += G) last = bi;
    for (int bi = last; bi >= 0; bi -= G) {
        const int r0 = bi * TS_B, nr = min(TS_B, n - r0);
        __syncthreads();
        load_X(bi);
        if (bi + 1 < nb)
            for (int idx = tid; idx < TS_B * TS_B; idx += 256) {
                const int r = idx >> 6, c = idx & 63;
                sL[r][c] = (r0 + TS_B + r < n && c < nr) ? F[(int64_t)(r0 + TS_B + r) * ld + r0 + c] : 0.0;
            }
        // columns of this block against x_j of every j > bi + 1 (descending): thread = (column, row group of 4)
        const int col = tid & 63, rg = tid >> 6;
        double ps = 0.0;
        int ready = nb, pp = 0;      // blocks [ready, nb) are known to be done (block j done implies every block above it)
        for (int j = nb - 1; j > bi + 1; --j) {
            if (j < ready) {
                if (tid == 0) {
                    int r = j;
                    if (ld_acquire(flagB + bi + 2) == epoch) r = bi + 2;
                    else while (ld_acquire(flagB + j) != epoch) __nanosleep(32);
                    s_ready[pp] = r;
                }
                __syncthreads();
                ready = s_ready[pp];
                pp ^= 1;
            }
            const int jr0 = j * TS_B, jn = min(TS_B, n - jr0);
            if (col < nr)
                for (int r = rg; r < jn; r += 4) ps = fma(F[(int64_t)(jr0 + r) * ld + r0 + col], __ldcg(x + jr0 + r), ps);
        }
        if (bi + 1 < nb) {
            if (tid == 0) while (ld_acquire(flagB + bi + 1) != epoch) __nanosleep(32);
            __syncthreads();
            if (tid < TS_B) sv[tid] = (r0 + TS_B + tid < n) ? __ldcg(x + r0 + TS_B + tid) : 0.0;
            __syncthreads();
            for (int r = rg; r < TS_B; r += 4) ps = fma(sL[r][col], sv[r], ps);
        }
        sred[rg][col] = ps;
        __syncthreads();
        if (tid < TS_B) sacc[tid] = (tid < nr) ? x[r0 + tid] - (sred[0][tid] + sred[1][tid] + sred[2][tid] + sred[3][tid]) : 0.0;
        __syncthreads();
        // x_i = X_ii^T acc : x(c) = sum_{r >= c} X(r, c) acc(r)
        {
            double t = 0.0;
            for (int r = col + rg; r < TS_B; r += 4) t = fma(sX[r][col], sacc[r], t);
            sred[rg][col] = t;
        }
        __syncthreads();
        if (tid < nr) x[r0 + tid] = sred[0][tid] + sred[1][tid] + sred[2][tid] + sred[3][tid];
        __syncthreads();
        if (tid == 0) { __threadfence(); st_release(flagB + bi, epoch); }
    }
}

void launch_dense_trisolve(const double* F, int64_t ld, int n, const double* Xall, const double* Dinv, double* x, double* y,
                           int* flagF, int* flagB, int epoch, int sm_count, cudaStream_t s) {
    if (n <= 0) return;
    const int nb = (n + TS_B - 1) / TS_B;
    const int grid = nb < sm_count ? nb : sm_count;
    dense_trisolve_kernel<<<grid, 256, TS_SMEM, s>>>(F, ld, n, Xall, Dinv, x, y, flagF, flagB, epoch);
    count_launch("dense_trisolve_kernel");
}

cudaError_t init_dense_attributes() {
    cudaError_t e = cudaFuncSetAttribute(dgemm_nt_kernel<GEMM_EPI_TRAIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(dgemm_nt_kernel<GEMM_EPI_PANEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(dgemm_nt_kernel<GEMM_EPI_SCHUR>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ldlt_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(dense_trisolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM);
}

// =====================================================================================================
// measured FP64 GEMM rate of dgemm_nt_kernel itself (n x n x n, trailing-update epilogue on a scratch matrix)
// =====================================================================================================

}  // namespace pipsgpu
