// Direct streaming solve (PIPSGPU_SOLVE=direct) - candidate successor of psolve_stream_kernel, same contract:
// ONE launch applies K_i^-1 (forward, D^-1, backward) for every scenario block of the GPU.
//
// STATUS: written after the round-1 GPU budget was spent; it compiles for sm_100a but has NOT run on a B200 yet. It is
// off by default and nothing in the default configuration touches it.
//
// Why (profiles/README.md, items 7 and 10): the staged kernel keeps two 9.25 KB TMA buffers and a 7 KB right-hand-side
// slice per warp, i.e. 8 warps per SM, and is bound by instruction issue / latency at that occupancy (27 % of the HBM
// roofline), not by bytes. This version stages nothing:
//  * forward sweep  : lane = row. The factor panel itself (column-major, ld = m) is k-major for this access, so the
//    K-part of L21 is read straight from the factor pool, coalesced, 8 independent loads per lane and batch; the pivot
//    block is applied as  y = X y  with  X^T = L11^-T  kept k-major next to X in the XY scratch (pivot_kernel).
//  * backward sweep : lane = (row group, column). It needs L21 row-major: pack_direct_kernel writes one transposed copy
//    BL[r * wp + k] (wp = padded width) after every factorisation - half the bytes of the staged solve pool, no headers,
//    no odd leading dimension. The pivot block is applied as  x = X^T t  on the row-major X of the XY scratch.
//  * per warp only the right-hand-side slice of the unit (7 KB) + two 1.25 KB vectors live in shared memory:
//    24 warps per SM (3 CTAs of 8 warps), <= 85 registers.
// Work units, tickets, dependency counters / flags, slots and the self-cleaning accumulator are exactly those of
// psolve.cu (same SolveUnit array, same host construction).
#include "kernels.cuh"

namespace pipsgpu {

namespace {

__device__ __forceinline__ int ld_relaxed(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(int* p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed_add(int* p, int v) {
    asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

constexpr int PD_WARPS = 8;                                   // warps per CTA; 3 CTAs per SM
constexpr int PD_WARP_DOUBLES = PS_SMAX + PS_NOUT + 32 + PS_XA;   // rhs slice | rows above the unit | y | x of ancestor rows

// s += sum_r B[r * 32 + lane] * v[rg + r * nrg]  over the passes r = 0 .. npass-1 (one pass = 32 consecutive doubles =
// nrg = 32 / wp rows of a row-major, wp-padded block); rows >= nrows contribute nothing. Two accumulators.
__device__ __forceinline__ double colsum_rowmajor(const double* __restrict__ B, const double* vs, int nrows, int rg, int nrg,
                                                  int lane) {
    const int npass = (nrows + nrg - 1) / nrg;      // nrg is a power of two: the compiler shifts
    double s0 = 0.0, s1 = 0.0;
    const double* base = B + lane;
    for (int p0 = 0; p0 < npass; p0 += 8) {
        double v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int r = rg + (p0 + q) * nrg;
            v[q] = (r < nrows) ? base[(size_t)(p0 + q) * 32] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 8; q += 2) {
            const int ra = rg + (p0 + q) * nrg, rb = rg + (p0 + q + 1) * nrg;
            s0 = fma(v[q], vs[ra < nrows ? ra : 0], s0);
            s1 = fma(v[q + 1], vs[rb < nrows ? rb : 0], s1);
        }
    }
    return s0 + s1;
}

}  // namespace

// -----------------------------------------------------------------------------------------------------
// pack: BL[r * wp + k] = L21(r, k) for the K-part rows (ancestors) of every front, zero for k >= w. One warp per front.
// -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pack_direct_kernel(const DirectFront* __restrict__ df, int nfronts,
                                                          const double* __restrict__ L, double* __restrict__ BL) {
    __shared__ double tile[4][32 * 33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fi = blockIdx.x * 4 + warp;
    if (fi >= nfronts) return;
    const DirectFront f = df[fi];
    const int w = f.w, na = f.na, m = f.m, wp = f.wp;
    const int sh = (wp == 8) ? 3 : (wp == 16 ? 4 : 5);
    const double* A = L + f.Loff + w;               // L21(r, k) = A[k * m + r]
    double* out = BL + f.blobB;
    double* T = tile[warp];
    for (int s0 = 0; s0 < na; s0 += 32) {
        const int nr = min(32, na - s0);
        __syncwarp();
        for (int k = 0; k < w; ++k) T[lane * 33 + k] = (lane < nr) ? A[(size_t)k * m + s0 + lane] : 0.0;
        __syncwarp();
        double* o = out + (size_t)s0 * wp;
        for (int idx = lane; idx < nr * wp; idx += 32) {
            const int r = idx >> sh, k = idx & (wp - 1);
            o[idx] = (k < w) ? T[r * 33 + k] : 0.0;
        }
    }
}

void launch_pack_direct(const DirectFront* df, int nfronts, const double* L, double* BL, cudaStream_t s) {
    if (nfronts <= 0) return;
    pack_direct_kernel<<<(nfronts + 3) / 4, 128, 0, s>>>(df, nfronts, L, BL);
    count_launch();
}

// -----------------------------------------------------------------------------------------------------
// the persistent direct solve
// -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PD_WARPS * 32, 3) psolve_direct_kernel(PDirectArgs a) {
    extern __shared__ __align__(16) double pd_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    double* xs = pd_smem + (size_t)warp * PD_WARP_DOUBLES;      // [PS_SMAX unit columns | PS_NOUT rows above the unit]
    double* sy = xs + PS_SMAX + PS_NOUT;                        // 32
    double* sxa = sy + 32;                                      // PS_XA
    const int nu = a.nunits, nt = 2 * nu;
    const int ep = a.epoch;

    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(a.ticket, 1);
        t = __shfl_sync(full, t, 0);
        if (t >= nt) break;
        const bool fwd = t < nu;
        const int ui = fwd ? t : nt - 1 - t;
        const SolveUnit un = a.units[ui];
        const int ncols = un.ncols, nout = un.nout, pbeg = un.pbeg;
        const int* __restrict__ outrows = a.rowg + un.out_rows_off;
        // ---------------- unit prologue (as in psolve_stream_kernel) ----------------
        if (fwd) {
            if (un.nchild > 0) {
                if (lane == 0) {
                    const int target = un.nchild * (ep + 1);
                    while (ld_relaxed(a.ucnt + ui) < target) __nanosleep(20);
                }
                __syncwarp();
                for (int r = lane; r < ncols; r += 32) {
                    const int p = pbeg + r;
                    xs[r] = a.in[a.gidx[p]] + __ldcg(a.acc + p);
                    a.acc[p] = 0.0;                       // self-cleaning: ready for the next solve
                }
            } else {
                for (int r = lane; r < ncols; r += 32) xs[r] = a.in[a.gidx[pbeg + r]];
            }
            for (int j = lane; j < nout; j += 32) xs[PS_SMAX + j] = 0.0;
        } else {
            if (lane == 0) {
                const int* fl = a.uflag + ((un.parent >= 0) ? un.parent : ui);
                const int target = (un.parent >= 0) ? 2 * ep + 2 : 2 * ep + 1;
                while (ld_relaxed(fl) < target) __nanosleep(20);
            }
            __syncwarp();
            for (int r = lane; r < ncols; r += 32) xs[r] = __ldcg(a.x + pbeg + r);
            for (int j = lane; j < nout; j += 32) xs[PS_SMAX + j] = __ldcg(a.x + outrows[j]);
        }
        __syncwarp();
        // ---------------- fronts of the unit ----------------
        DirectFront f = a.df[fwd ? un.first : un.first + un.nfronts - 1];
        for (int fr = 0; fr < un.nfronts; ++fr) {
            DirectFront nf = f;                     // the next front's descriptor travels while this front is processed
            if (fr + 1 < un.nfronts) nf = a.df[fwd ? un.first + fr + 1 : un.first + un.nfronts - 2 - fr];
            const int w = f.w, na = f.na, m = f.m, wp = f.wp;
            const int crel = f.col0 - pbeg;
            const int* __restrict__ rowg = a.rowg + f.rows_off;
            const uint16_t* __restrict__ slot = a.slots + f.rows_off;
            const double* __restrict__ X = a.XY + f.XYoff;              // X(k, j)  at X[k * wp + j]   (row-major)
            const double* __restrict__ XT = X + (size_t)wp * wp;       // X(r, k)  at XT[k * wp + r]  (k-major)
            if (fwd) {
                // ---- y = X y : lane = row ----
                double y = (lane < w) ? xs[crel + lane] : 0.0;
                sy[lane] = y;
                // D^-1 coefficients of this lane's column (1x1 and 2x2 pivots; pairs never straddle a front)
                double dd = 0.0, dn = 0.0, dp = 0.0;
                if (lane < w) {
                    const int p = f.col0 + lane;
                    const int pt = a.ptype[p];
                    dd = a.Did[p];
                    if (pt == 1) dn = a.Die[p];
                    else if (pt == 2) dp = a.Die[p - 1];
                }
                __syncwarp();
                {
                    double t0 = 0.0, t1 = 0.0;
                    for (int k0 = 0; k0 < w; k0 += 8) {
                        double v[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] = (k0 + q < w && lane < w) ? XT[(size_t)(k0 + q) * wp + lane] : 0.0;
#pragma unroll
                        for (int q = 0; q < 8; q += 2) {
                            t0 = fma(v[q], sy[(k0 + q) & 31], t0);
                            t1 = fma(v[q + 1], sy[(k0 + q + 1) & 31], t1);
                        }
                    }
                    __syncwarp();
                    y = (lane < w) ? t0 + t1 : 0.0;
                }
                sy[lane] = y;
                const double ynext = __shfl_down_sync(full, y, 1);
                const double yprev = __shfl_up_sync(full, y, 1);
                const double z = fma(dd, y, fma(dn, ynext, dp * yprev));
                if (lane < w) xs[crel + lane] = z;
                __syncwarp();
                // ---- rows above -= L21 y : lane = row, straight from the factor panel (k-major for this access) ----
                const double* __restrict__ A = a.L + f.Loff + w;     // L21(i, k) = A[k * m + i]
                for (int i0 = 0; i0 < na; i0 += 32) {
                    const int i = i0 + lane;
                    const bool in = i < na;
                    const int sl = in ? (int)slot[i] : 0xFFFF;
                    double s0 = 0.0, s1 = 0.0;
                    for (int k0 = 0; k0 < w; k0 += 8) {
                        double v[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] = (k0 + q < w && in) ? A[(size_t)(k0 + q) * m + i] : 0.0;
#pragma unroll
                        for (int q = 0; q < 8; q += 2) {
                            s0 = fma(v[q], sy[(k0 + q) & 31], s0);
                            s1 = fma(v[q + 1], sy[(k0 + q + 1) & 31], s1);
                        }
                    }
                    if (in) {
                        const double s = s0 + s1;
                        if (sl != 0xFFFF) xs[sl] -= s;                      // distinct rows -> distinct slots
                        else atomicAdd(a.acc + rowg[i], -s);               // row list longer than the staging area
                    }
                }
                __syncwarp();
            } else {
                // ---- t = x_own - L21^T x_anc : lane = (row group, column), row-major transposed copy ----
                double tt = (lane < w) ? xs[crel + lane] : 0.0;
                const int sh = (wp == 8) ? 3 : (wp == 16 ? 4 : 5);
                const int rg = lane >> sh, nrg = 32 >> sh;
                const double* __restrict__ BL = a.BL + f.blobB;
                double s = 0.0;
                for (int r0 = 0; r0 < na; r0 += PS_XA) {
                    const int nr = min(PS_XA, na - r0);
                    for (int i = lane; i < nr; i += 32) {
                        const int sl = slot[r0 + i];
                        sxa[i] = (sl != 0xFFFF) ? xs[sl] : __ldcg(a.x + rowg[r0 + i]);
                    }
                    __syncwarp();
                    s += colsum_rowmajor(BL + (size_t)r0 * wp, sxa, nr, rg, nrg, lane);
                    __syncwarp();
                }
                for (int o = wp; o < 32; o <<= 1) s += __shfl_xor_sync(full, s, o);
                tt -= s;
                if (lane >= w) tt = 0.0;
                // ---- x = X^T t : the same column sum over the w pivot rows of the row-major X ----
                sy[lane] = tt;
                __syncwarp();
                double xv = colsum_rowmajor(X, sy, w, rg, nrg, lane);
                for (int o = wp; o < 32; o <<= 1) xv += __shfl_xor_sync(full, xv, o);
                __syncwarp();
                if (lane < w) xs[crel + lane] = xv;
                __syncwarp();
            }
            f = nf;
        }
        // ---------------- unit epilogue (as in psolve_stream_kernel) ----------------
        if (fwd) {
            for (int r = lane; r < ncols; r += 32) a.x[pbeg + r] = xs[r];
            for (int j = lane; j < nout; j += 32) {
                const double v = xs[PS_SMAX + j];
                if (v != 0.0) atomicAdd(a.acc + outrows[j], v);
            }
            __syncwarp();
            if (lane == 0) {
                fence_acq_rel_gpu();
                if (un.parent >= 0) red_relaxed_add(a.ucnt + un.parent, 1);
                else st_relaxed(a.uflag + ui, 2 * ep + 1);      // a root publishes "forward sweep of my tree is done"
            }
        } else {
            if (a.out_mode == 0) {
                for (int r = lane; r < ncols; r += 32) a.out[a.gidx[pbeg + r]] = xs[r];
            } else {
                for (int r = lane; r < ncols; r += 32) a.out[a.gidx[pbeg + r]] -= xs[r];
            }
            if (un.nchild > 0) {
                for (int r = lane; r < ncols; r += 32) a.x[pbeg + r] = xs[r];
                __syncwarp();
                if (lane == 0) {
                    fence_acq_rel_gpu();
                    st_relaxed(a.uflag + ui, 2 * ep + 2);
                }
            }
        }
        __syncwarp();
    }
    // the last warp to leave re-arms the ticket counter for the next launch
    if (lane == 0) {
        const int total = gridDim.x * PD_WARPS;
        const int done = atomicAdd(a.ticket + 1, 1);
        if (done == total - 1) {
            a.ticket[1] = 0;
            __threadfence();
            a.ticket[0] = 0;
        }
    }
}

// The grid must be co-resident (units spin on units that hold earlier tickets): as many CTAs as the occupancy calculator
// guarantees, never more than the tickets need.
void launch_psolve_direct(const PDirectArgs& a, int sm_count, cudaStream_t s) {
    if (a.nunits <= 0) return;
    static int ctas_per_sm = 0;
    const int smem = PD_WARPS * PD_WARP_DOUBLES * (int)sizeof(double);
    if (ctas_per_sm == 0) {
        cudaFuncSetAttribute(psolve_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, psolve_direct_kernel, PD_WARPS * 32, smem) != cudaSuccess ||
            ctas_per_sm < 1)
            ctas_per_sm = 1;
    }
    int ctas = sm_count * ctas_per_sm;
    const int need = (2 * a.nunits + PD_WARPS - 1) / PD_WARPS;
    if (ctas > need) ctas = need;
    psolve_direct_kernel<<<ctas, PD_WARPS * 32, smem, s>>>(a);
    count_launch();
}

}  // namespace pipsgpu
