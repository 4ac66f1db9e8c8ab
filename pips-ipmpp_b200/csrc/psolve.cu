// Streaming persistent supernodal solve (sm_100a): ONE launch applies K_i^-1 for every scenario block of the GPU.
//
// Replaces the per-right-hand-side triangular solves the reference delegates to its sparse plugins
// (PardisoSolver.C:207-228 phase 33, Ma27Solver.C:127-138 ma27cd, Ma57Solver.C:108-186 ma57cd) as they are used by
// Lsolve / Ltsolve (sLinsysRootAug.C:323-365, DistributedLeafLinearSystem.C:153-212).
//
// Design (measured history in profiles/README.md):
//  * The assembly forest of every block is cut into WORK UNITS: maximal bottom subtrees with at most `solve_subtree`
//    columns (their columns are a contiguous range of the postorder), and single "top" fronts above the cut. One warp
//    owns a unit. Inside a subtree nothing is shared with other warps: the slice of the right-hand side lives in shared
//    memory, updates to ancestor rows inside the subtree are shared-memory read-modify-writes, updates to rows above the
//    cut are accumulated in shared memory too (they all belong to the row structure of the subtree's root front) and
//    flushed with one red.global.add.f64 per row at the end. Only units synchronise through global memory: a
//    child-unit counter (forward) and a completion flag (backward) per unit, monotone in the solve epoch so that nothing
//    is cleared between launches.
//  * Every factor entry is read once per sweep. pack_solve_kernel (after every numeric factorisation) copies the K-part
//    of each front into the "solve pool": per front a sequence of self-contained chunks [ D^-1 header | rows, row-major
//    with odd leading dimension | shared-memory slot of every row ], chunks of one unit contiguous and in sweep order,
//    the last chunk of a front cut to its own rows. With PIPSGPU_XFORM (default) the pivot-block rows hold X = L11^-1
//    instead of L11 and are applied as matrix-vector products instead of substitution chains. A warp keeps a TMA bulk
//    copy (cp.async.bulk global -> shared, completion on an mbarrier) in flight for the next chunk while it works on the
//    current one (two buffers per warp), independent of any dependency wait.
//  * Measured (profiles/README.md): 0.90 ms per launch at C2 = 27 % of the HBM roofline on algorithmic bytes, 57 % on
//    the 3.3 GB it really streams; at 8 warps per SM (209 KB of shared memory per CTA) it is bound by instruction issue
//    and latency, not by bytes - rows-per-chunk and chunk counts therefore travel in SolveFront (no device-side integer
//    division) and the next front's descriptor is prefetched. psolve_direct.cu is the no-staging successor candidate.
//  * Tickets: forward sweep = [subtree units, largest first | top units by tree level, deepest first], backward sweep
//    = the reverse. Warps draw tickets from a global counter, one unit ahead of the one they process (so the bulk-copy
//    pipeline never drains). A unit's prerequisites always hold EARLIER tickets and a warp finishes its tickets in
//    order, so the spin-waits only ever wait on warps that are resident and running: no deadlock as long as the grid
//    is co-resident (1 CTA per SM).
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
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
// 1-D TMA bulk copy global -> shared; dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace

// chunk geometry of a front (host + device): every chunk = [3*32 header doubles | R*ldr row doubles | R/4 slot doubles]
__host__ __device__ int ps_rows_per_chunk(int ldr) {
    int r = (4 * (PS_BUF - PS_HDR - 1)) / (4 * ldr + 1);
    r &= ~3;                       // slots: 4 uint16 per double
    return r < PS_XA ? r : PS_XA;
}
__host__ __device__ int ps_chunk_doubles(int ldr) {
    const int R = ps_rows_per_chunk(ldr);
    return (PS_HDR + R * ldr + R / 4 + 1) & ~1;      // even: bulk copies need 16-byte addresses and sizes
}
// Rows stored in chunk c of a front with `rows` rows: every chunk but the last holds R rows and occupies
// ps_chunk_doubles(ldr) doubles (so chunk c starts at c * ps_chunk_doubles); the last one is cut to its own rows
// (rounded up to the 4 slots of one double) - most fronts are a single short chunk.
__host__ __device__ int ps_chunk_rows(int ldr, int rows, int c) {
    const int R = ps_rows_per_chunk(ldr);
    const int rem = rows - c * R;
    return rem < R ? ((rem + 3) & ~3) : R;
}
__host__ __device__ int ps_chunk_doubles_rows(int ldr, int cr) { return (PS_HDR + cr * ldr + cr / 4 + 1) & ~1; }
// the same with R known (SolveFront::R): no integer division on the device
__device__ __forceinline__ int ps_chunk_rows_R(int R, int rows, int c) {
    const int rem = rows - c * R;
    return rem < R ? ((rem + 3) & ~3) : R;
}
__host__ __device__ int64_t ps_front_doubles(int ldr, int rows) {
    const int R = ps_rows_per_chunk(ldr);
    const int nch = (rows + R - 1) / R;
    return (int64_t)(nch - 1) * ps_chunk_doubles(ldr) + ps_chunk_doubles_rows(ldr, ps_chunk_rows(ldr, rows, nch - 1));
}

// -----------------------------------------------------------------------------------------------------
// pack: panel (column-major, ld = m, rows [pivot | ancestors | border]) -> chunks of the solve pool. One warp per chunk.
// -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pack_solve_kernel(const SolveFront* __restrict__ sf, const int* __restrict__ chunk_front,
                                                         const int* __restrict__ chunk_idx, int nchunks,
                                                         const FrontDesc* __restrict__ fronts, const double* __restrict__ L,
                                                         const double* __restrict__ Did, const double* __restrict__ Die,
                                                         const int8_t* __restrict__ ptype, const uint16_t* __restrict__ slots,
                                                         double* __restrict__ SL, const double* __restrict__ XY) {
    __shared__ double tile[4][32 * 33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * 4 + warp;
    if (q >= nchunks) return;
    const SolveFront f = sf[chunk_front[q]];
    const int c = chunk_idx[q];
    const FrontDesc fd = fronts[f.front];
    const int w = f.w, rows = f.w + f.na, m = fd.w + fd.na + fd.nb, ldr = f.ldr;
    const int R = f.R, CR = ps_chunk_rows_R(R, rows, c);
    const int r0 = c * R, r1 = min(rows, r0 + R);
    double* dst = SL + f.blob + (size_t)c * ps_chunk_doubles_rows(ldr, R);
    // header: z = dd*y + dn*y_next + dp*y_prev  (1x1 and 2x2 pivots without a type switch in the solve)
    if (c == 0) {
        const int p = fd.col0 + lane;
        double dd = 0.0, dn = 0.0, dp = 0.0;
        if (lane < w) {
            const int t = ptype[p];
            dd = Did[p];
            if (t == 1) dn = Die[p];
            else if (t == 2) dp = Die[p - 1];
        }
        dst[lane] = dd; dst[32 + lane] = dn; dst[64 + lane] = dp;
    }
    const double* F = L + fd.Loff;
    double* T = tile[warp];
    double* rowsdst = dst + PS_HDR;
    for (int s0 = r0; s0 < r1; s0 += 32) {
        const int nr = min(32, r1 - s0);
        __syncwarp();
        for (int k = 0; k < w; ++k) T[lane * 33 + k] = (lane < nr) ? F[(size_t)k * m + s0 + lane] : 0.0;
        if (XY != nullptr && s0 == 0) {
            // pivot-block rows (first slab of chunk 0): X = L11^-1 replaces L11 below the diagonal. X is a contiguous
            // wp x wp row-major block: read it coalesced and scatter it into the tile.
            const int wp = padded_width(w), sh = (wp == 8) ? 3 : (wp == 16 ? 4 : 5);
            const double* Xf = XY + fd.XYoff;
            __syncwarp();
            for (int idx = lane; idx < w * wp; idx += 32) {
                const int xr = idx >> sh, xk = idx & (wp - 1);
                if (xk < xr) T[xr * 33 + xk] = Xf[idx];
            }
        }
        __syncwarp();
        double* d2 = rowsdst + (size_t)(s0 - r0) * ldr;
        const int total = nr * ldr;
        for (int idx = lane; idx < total; idx += 32) {
            const int r = idx / ldr, k = idx - r * ldr;
            double v = (k < w) ? T[r * 33 + k] : 0.0;
            if (s0 + r < w && k > s0 + r) v = 0.0;      // pivot-block row (L11 or X): unit diagonal, zeros above it
            d2[idx] = v;
        }
    }
    // shared-memory slot of every L21 row of this chunk
    uint16_t* sd = reinterpret_cast<uint16_t*>(rowsdst + (size_t)CR * ldr);
    for (int i = r0 + lane; i < r0 + CR; i += 32)
        sd[i - r0] = (i >= w && i < r1) ? slots[f.rows_off + (i - w)] : (uint16_t)0xFFFF;
}

void launch_pack_solve(const SolveFront* sf, const int* chunk_front, const int* chunk_idx, int nchunks, const FrontDesc* fronts,
                       const double* L, const double* Did, const double* Die, const int8_t* ptype, const uint16_t* slots,
                       double* SL, const double* XY, cudaStream_t s) {
    if (nchunks <= 0) return;
    pack_solve_kernel<<<(nchunks + 3) / 4, 128, 0, s>>>(sf, chunk_front, chunk_idx, nchunks, fronts, L, Did, Die, ptype, slots, SL, XY);
    count_launch("pack_solve_kernel");
}

// -----------------------------------------------------------------------------------------------------
// the persistent streaming solve
// -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PS_WARPS * 32, 1) psolve_stream_kernel(PStreamArgs a) {
    extern __shared__ __align__(128) double smem[];
    __shared__ uint64_t s_bar[PS_WARPS][2];
    __shared__ double s_y[PS_WARPS][32];
    __shared__ double s_xa[PS_WARPS][PS_XA];
    __shared__ SolveFront s_desc[PS_WARPS][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    double* buf0 = smem + (size_t)warp * PS_WARP_DOUBLES;
    double* xs = buf0 + 2 * PS_BUF;              // [PS_SMAX unit columns | PS_NOUT rows above the unit]
    uint64_t* bar = s_bar[warp];
    double* sy = s_y[warp];
    double* sxa = s_xa[warp];
    if (lane == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int nu = a.nunits, nt = 2 * nu;
    const int ep = a.epoch;

    auto draw = [&]() {      // next ticket (all lanes get the same value)
        int t = 0;
        if (lane == 0) t = atomicAdd(a.ticket, 1);
        return __shfl_sync(full, t, 0);
    };
    auto unit_of = [&](int t) { return (t < nu) ? t : nt - 1 - t; };

    // ---- issue side: cursor over (ticket, front, chunk) of the next bulk copy; runs at most two chunks ahead ----
    int i_t = draw();            // ticket being issued
    SolveUnit i_unit = {};
    SolveFront i_desc = {}, n_desc = {};
    int n_fi = -1;               // front index held in n_desc, or -1
    int i_fr = 0, i_ch = 0, i_nch = 0;     // front position within the unit (processing order), chunk position, #chunks
    auto issue_load_front = [&]() {
        const bool fwd = i_t < nu;
        const int fi = fwd ? i_unit.first + i_fr : i_unit.first + i_unit.nfronts - 1 - i_fr;
        // the descriptor of the following front of the unit was requested one front ago (n_desc): its load latency
        // hides behind the processing of a whole front
        if (n_fi == fi) i_desc = n_desc;
        else i_desc = a.sf[fi];
        if (i_fr + 1 < i_unit.nfronts) {
            n_fi = fwd ? fi + 1 : fi - 1;
            n_desc = a.sf[n_fi];
        } else {
            n_fi = -1;
        }
        i_nch = i_desc.nch;
        i_ch = 0;
    };
    auto issue_load_unit = [&]() {
        if (i_t < nt) {
            i_unit = a.units[unit_of(i_t)];
            i_fr = 0;
            issue_load_front();
        }
    };
    auto issue = [&](int slot) {
        if (i_t >= nt) return;
        const bool fwd = i_t < nu;
        const int c = fwd ? i_ch : i_nch - 1 - i_ch;      // backward: last rows first, the L11 chunk last
        if (lane == 0) {
            const int cd = ps_chunk_doubles_rows(i_desc.ldr, i_desc.R);
            const uint32_t bytes = (uint32_t)ps_chunk_doubles_rows(i_desc.ldr, ps_chunk_rows_R(i_desc.R, i_desc.w + i_desc.na, c)) * 8u;
            s_desc[warp][slot] = i_desc;
            mbar_expect_tx(&bar[slot], bytes);
            bulk_g2s(buf0 + (size_t)slot * PS_BUF, a.SL + i_desc.blob + (int64_t)c * cd, bytes, &bar[slot]);
        }
        if (++i_ch >= i_nch) {
            if (++i_fr >= i_unit.nfronts) {
                // next unit: tickets are drawn in order, one ahead of the consume side
                i_t = draw();
                issue_load_unit();
            } else {
                issue_load_front();
            }
        }
        __syncwarp();
    };
    // The consume side must see the same ticket sequence as the issue side: tickets are queued in a 4-entry ring.
    // (The issue side is at most 2 chunks, hence at most 2 units, ahead.)
    int tq0 = i_t, tq1 = -1, tq2 = -1, tq3 = -1;       // tickets drawn-but-not-consumed, in order
    auto push_ticket = [&](int t) { if (tq1 < 0) tq1 = t; else if (tq2 < 0) tq2 = t; else tq3 = t; };

    issue_load_unit();
    {
        const int before = i_t;
        issue(0);
        if (i_t != before) push_ticket(i_t);
    }
    {
        const int before = i_t;
        issue(1);
        if (i_t != before) push_ticket(i_t);
    }

    uint32_t phase0 = 0, phase1 = 0;
    int slot = 0;
    for (;;) {
        const int t = tq0;
        if (t >= nt) break;
        const bool fwd = t < nu;
        const int ui = unit_of(t);
        const SolveUnit un = a.units[ui];
        const int ncols = un.ncols, nout = un.nout, pbeg = un.pbeg;
        const int* __restrict__ outrows = a.rowg + un.out_rows_off;      // global positions of the rows above the unit
        // ---------------- unit prologue ----------------
        if (fwd) {
            if (un.nchild > 0) {
                if (lane == 0) {
                    const int target = un.nchild * (ep + 1);
                    while (ld_relaxed(a.ucnt + ui) < target) __nanosleep(20);
                    fence_acq_rel_gpu();      // acquire side of the producers' fence + red: acc[] / x[] reads below are ordered
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
            // a unit whose root front has a parent waits for the parent's unit (which transitively implies that the
            // forward sweep of the whole tree is complete); a root unit waits for its own forward step
            if (lane == 0) {
                const int* fl = a.uflag + ((un.parent >= 0) ? un.parent : ui);
                const int target = (un.parent >= 0) ? 2 * ep + 2 : 2 * ep + 1;
                while (ld_relaxed(fl) < target) __nanosleep(20);
                fence_acq_rel_gpu();
            }
            __syncwarp();
            for (int r = lane; r < ncols; r += 32) xs[r] = __ldcg(a.x + pbeg + r);
            for (int j = lane; j < nout; j += 32) xs[PS_SMAX + j] = __ldcg(a.x + outrows[j]);
        }
        __syncwarp();
        // ---------------- fronts of the unit ----------------
        for (int fr = 0; fr < un.nfronts; ++fr) {
            // the descriptor travels with the first chunk of the front (written by the issue side)
            mbar_wait(&bar[slot], slot ? phase1 : phase0);
            const SolveFront f = s_desc[warp][slot];
            const int w = f.w, na = f.na, ldr = f.ldr, rows = w + na;
            const int R = f.R;
            const int nc = f.nch;
            const int crel = f.col0 - pbeg;
            const int* __restrict__ rowg = a.rowg + f.rows_off;
            if (fwd) {
                double y = (lane < w) ? xs[crel + lane] : 0.0;
                for (int c = 0; c < nc; ++c) {
                    const int r0 = c * R, r1 = min(rows, r0 + R);
                    const double* H = buf0 + (size_t)slot * PS_BUF;
                    const double* B = H + PS_HDR;
                    const uint16_t* SLT = reinterpret_cast<const uint16_t*>(B + (size_t)ps_chunk_rows_R(R, rows, c) * ldr);
                    if (c > 0) mbar_wait(&bar[slot], slot ? phase1 : phase0);
                    if (slot) phase1 ^= 1; else phase0 ^= 1;
                    int rb = r0;
                    if (c == 0) {
                        // y = L11^-1 y : lane i owns row i
                        if (a.xform) {
                            // the pool holds X = L11^-1: one small matrix-vector product, no substitution chain
                            sy[lane] = y;
                            __syncwarp();
                            const double* xr = B + lane * ldr;      // lanes >= w read rows below the pivot block: discarded
                            double t0 = 0.0, t1 = 0.0;
                            int k = 0;
                            for (; k + 2 <= w; k += 2) {
                                t0 = fma(xr[k], sy[k], t0);
                                t1 = fma(xr[k + 1], sy[k + 1], t1);
                            }
                            if (k < w) t0 = fma(xr[k], sy[k], t0);
                            __syncwarp();
                            y = (lane < w) ? t0 + t1 : 0.0;
                        } else {
                            for (int k = 0; k < w; ++k) {
                                const double yk = __shfl_sync(full, y, k);
                                if (lane > k && lane < w) y = fma(-B[lane * ldr + k], yk, y);
                            }
                        }
                        sy[lane] = y;
                        // x = D^-1 y (1x1 and 2x2 pivots; pairs never straddle a front)
                        const double ynext = __shfl_down_sync(full, y, 1);
                        const double yprev = __shfl_up_sync(full, y, 1);
                        const double z = fma(H[lane], y, fma(H[32 + lane], ynext, H[64 + lane] * yprev));
                        if (lane < w) xs[crel + lane] = z;
                        __syncwarp();
                        rb = w;
                    }
                    // rows above -= L21 y : lane per row, four independent partial sums
                    for (int i0 = rb; i0 < r1; i0 += 32) {
                        const int i = i0 + lane;
                        if (i < r1) {
                            const double* row = B + (size_t)(i - r0) * ldr;
                            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                            int k = 0;
                            for (; k + 4 <= w; k += 4) {
                                s0 = fma(row[k], sy[k], s0);
                                s1 = fma(row[k + 1], sy[k + 1], s1);
                                s2 = fma(row[k + 2], sy[k + 2], s2);
                                s3 = fma(row[k + 3], sy[k + 3], s3);
                            }
                            for (; k < w; ++k) s0 = fma(row[k], sy[k], s0);
                            const double s = (s0 + s1) + (s2 + s3);
                            const int sl = SLT[i - r0];
                            if (sl != 0xFFFF) xs[sl] -= s;                  // distinct rows -> distinct slots
                            else atomicAdd(a.acc + rowg[i - w], -s);       // row list longer than the staging area
                        }
                        __syncwarp();
                    }
                    __syncwarp();
                    {
                        const int before = i_t;
                        issue(slot);
                        if (i_t != before) push_ticket(i_t);
                    }
                    slot ^= 1;
                }
            } else {
                double tt = (lane < w) ? xs[crel + lane] : 0.0;
                const int wp = padded_width(w);
                const int kq = lane & (wp - 1), grp = lane / wp, ngrp = 32 / wp;
                const int k = (kq < w) ? kq : 0;     // padding columns redo column 0 (result discarded)
                double s0 = 0.0, s1 = 0.0;
                for (int cc = 0; cc < nc; ++cc) {
                    const int c = nc - 1 - cc;
                    const int r0 = c * R, r1 = min(rows, r0 + R);
                    const int rb = max(r0, w);
                    const double* H = buf0 + (size_t)slot * PS_BUF;
                    const double* B = H + PS_HDR;
                    const uint16_t* SLT = reinterpret_cast<const uint16_t*>(B + (size_t)ps_chunk_rows_R(R, rows, c) * ldr);
                    if (cc > 0) mbar_wait(&bar[slot], slot ? phase1 : phase0);
                    if (slot) phase1 ^= 1; else phase0 ^= 1;
                    // x of the ancestor rows of this chunk
                    for (int i = rb + lane; i < r1; i += 32) {
                        const int sl = SLT[i - r0];
                        sxa[i - rb] = (sl != 0xFFFF) ? xs[sl] : __ldcg(a.x + rowg[i - w]);
                    }
                    __syncwarp();
                    // s_k += sum_i L21(i,k) x_anc(i) : lane = (row group, column k)
                    {
                        int i = rb + grp;
                        for (; i + ngrp < r1; i += 2 * ngrp) {
                            s0 = fma(B[(size_t)(i - r0) * ldr + k], sxa[i - rb], s0);
                            s1 = fma(B[(size_t)(i + ngrp - r0) * ldr + k], sxa[i + ngrp - rb], s1);
                        }
                        if (i < r1) s0 = fma(B[(size_t)(i - r0) * ldr + k], sxa[i - rb], s0);
                    }
                    if (c == 0) {
                        double s = s0 + s1;
                        for (int o = wp; o < 32; o <<= 1) s += __shfl_xor_sync(full, s, o);
                        tt -= s;
                        if (lane >= w) tt = 0.0;
                        // x = L11^-T t : lane j owns column j
                        if (a.xform) {
                            // x_j = sum_k X(k,j) t_k (rows of X have zeros right of the diagonal); lanes >= w: discarded
                            sy[lane] = tt;
                            __syncwarp();
                            double t0 = 0.0, t1 = 0.0;
                            int kk = 0;
                            for (; kk + 2 <= w; kk += 2) {
                                t0 = fma(B[kk * ldr + lane], sy[kk], t0);
                                t1 = fma(B[(kk + 1) * ldr + lane], sy[kk + 1], t1);
                            }
                            if (kk < w) t0 = fma(B[kk * ldr + lane], sy[kk], t0);
                            __syncwarp();
                            tt = (lane < w) ? t0 + t1 : 0.0;
                        } else {
                            for (int kk = w - 1; kk > 0; --kk) {
                                const double xk = __shfl_sync(full, tt, kk);
                                if (lane < kk) tt = fma(-B[kk * ldr + lane], xk, tt);
                            }
                        }
                        if (lane < w) xs[crel + lane] = tt;
                    }
                    __syncwarp();
                    {
                        const int before = i_t;
                        issue(slot);
                        if (i_t != before) push_ticket(i_t);
                    }
                    slot ^= 1;
                }
            }
        }
        // ---------------- unit epilogue ----------------
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
        tq0 = tq1; tq1 = tq2; tq2 = tq3; tq3 = -1;
        if (tq0 < 0) break;      // cannot happen: the issue side always holds a ticket (possibly >= nt)
    }
    // the last warp to leave re-arms the ticket counter for the next launch
    if (lane == 0) {
        const int total = gridDim.x * PS_WARPS;
        const int done = atomicAdd(a.ticket + 1, 1);
        if (done == total - 1) {
            a.ticket[1] = 0;
            __threadfence();
            a.ticket[0] = 0;
        }
    }
}

int psolve_stream_smem_bytes() { return PS_WARPS * PS_WARP_DOUBLES * (int)sizeof(double); }

// function attributes are per device: called from pipsgpu_ctx_create with the context's device current
cudaError_t init_psolve_attributes() {
    return cudaFuncSetAttribute(psolve_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, psolve_stream_smem_bytes());
}

void launch_psolve_stream(const PStreamArgs& a, int sm_count, cudaStream_t s) {
    if (a.nunits <= 0) return;
    const int smem = psolve_stream_smem_bytes();
    int ctas = sm_count;
    const int need = (2 * a.nunits + PS_WARPS - 1) / PS_WARPS;
    if (ctas > need) ctas = need;
    psolve_stream_kernel<<<ctas, PS_WARPS * 32, smem, s>>>(a);
    count_launch("psolve_stream_kernel");
}

}  // namespace pipsgpu
