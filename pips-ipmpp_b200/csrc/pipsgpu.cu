// C-ABI implementation (include/pipsgpu.h): context / leaf / Schur-complement / kkt objects on top of kernels.cu.
// Host orchestration mirrors the reference's call stack: factor2 -> leaf LDL^T + Schur terms ->
// reduceKKT -> finalizeKKT -> factorizeKKT, and solveCompressed = Lsolve; Dsolve; Ltsolve - but batched over all
// scenario blocks of the GPU and stream-ordered, with the operands resident in HBM.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pipsgpu.h"
#include "common.h"
#include "kernels.cuh"
#include "nccl_dyn.h"
#include "symbolic.h"

using namespace pipsgpu;

#define CUDA_CHECK(expr)                                                                                          \
    do {                                                                                                          \
        cudaError_t _e = (expr);                                                                                  \
        if (_e != cudaSuccess)                                                                                    \
            throw pipsgpu::Error(PIPSGPU_ERR_CUDA, std::string(#expr) + " failed: " + cudaGetErrorString(_e));    \
    } while (0)

namespace {

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count) {
        release();
        n = count;
        if (count > 0) CUDA_CHECK(cudaMalloc(&p, sizeof(T) * count));
    }
    void ensure(size_t count) {
        if (count > n) alloc(count);
    }
    void upload(const std::vector<T>& v, cudaStream_t s) {
        alloc(v.size());
        if (!v.empty()) CUDA_CHECK(cudaMemcpyAsync(p, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, s));
    }
};

struct CsrDev {
    int m = 0, n = 0;
    int64_t nnz = 0;
    DevBuf<int> krow, jcol;
    DevBuf<double> val;
    void upload(int m_, int n_, const int* kr, const int* jc, const double* v, cudaStream_t s) {
        m = m_; n = n_;
        std::vector<int> k(m + 1, 0);
        if (kr && m > 0) std::copy(kr, kr + m + 1, k.begin());
        nnz = k[m];
        krow.upload(k, s);
        std::vector<int> j(jc ? jc : nullptr, jc ? jc + nnz : nullptr);
        std::vector<double> vv(v ? v : nullptr, v ? v + nnz : nullptr);
        if ((int64_t)j.size() != nnz || (int64_t)vv.size() != nnz) { j.assign(nnz, 0); vv.assign(nnz, 0.0); }
        jcol.upload(j, s);
        val.upload(vv, s);
        CUDA_CHECK(cudaStreamSynchronize(s));
    }
};

}  // namespace

// ============================================================================================================
// objects
// ============================================================================================================
struct pipsgpu_leaf {
    pipsgpu_ctx* ctx = nullptr;
    int index = -1;
    int n = 0, nx = 0, ny = 0, nz = 0;
    std::vector<int> krow, jcol;
    std::vector<double> val;
    std::vector<int> diag_entry;     // per row: index of the diagonal entry in val
    // border
    int n0 = 0, my0 = 0, myl = 0, mzl = 0, nB = 0;
    std::vector<int> brow, bcol;
    std::vector<double> bval;
    bool border_set = false;
    // symbolic
    LeafSymbolic S;
    BorderGroups G;                  // grouping of the big fronts (consumed by pipsgpu_ctx_commit)
    bool analyzed = false;
    double knorm = 1.0;
    // placement in the context pools (valid after commit)
    int pos_off = 0;
    int64_t Lbase = 0, Lsize = 0, Cbase = 0, Csize = 0, kval_off = 0, bval_off = 0, upd_base = 0, upd_size = 0;
    int front_base = 0;
    std::vector<int> level_begin, level_count;   // per global level: sub-range of level_fronts owned by this leaf
    std::vector<int> fl_begin, fl_count;         // per (global level, small|big): sub-range of fl_fronts owned by this leaf
    std::vector<int64_t> tile_begin, tile_count; // per global level: sub-range of the update tiles of this leaf's big fronts
    std::vector<int64_t> gtile_begin, gtile_count;   // per global level: sub-range of the group tiles whose group root sits on that level
    int chunk_begin = 0, chunk_count = 0;        // sub-range of the by-leaf chunk list (solve-pool pack of this leaf only)
    int batch = 0;                               // factorisation batch of this leaf
    double reg[3] = {0.0, 0.0, 0.0};             // this leaf's own (regP, regDy, regDz), added to every diagonal refresh
    // persistent work buffers of the plug-in multi-RHS solve (pipsgpu_leaf_solve), allocated on first use
    DevBuf<double> sv_b, sv_x, sv_u, sv_r, sv_r2;
    DevBuf<int> sv_perm;
    int sv_chunk = 0;
    int64_t stats8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double flops = 0;
    double flops4[4] = {0, 0, 0, 0};   // canonical F1 | Schur-term flops (lower triangle of L_B D L_B^T) | update flops | panel flops
    int inertia[3] = {0, 0, 0};
    int n_perturbed = 0;
    bool factorized = false;
};

struct pipsgpu_sc {
    pipsgpu_ctx* ctx = nullptr;
    int n0 = 0, my0 = 0, myl = 0, mzl = 0, n = 0;
    DevBuf<double> A;        // assembled matrix, n x n row-major, lower triangle used (owned by the linear system)
    DevBuf<double> Fm;       // the solver's own copy, overwritten by the factors (DeSymIndefSolver.C:65 copies too); n x ldf
    int64_t ldf = 0;         // leading dimension of Fm: n rounded up to 16 doubles (TMA: 16-byte aligned rows, 128-byte here)
    DevBuf<double> D, De, Dinv, Dinv_e, Wbuf, Wc, pack, rhs, swork, scale, Xall, ybuf;
    TMap tmF, tmW[2], tmX;   // tensor maps of the factor, the two W21 panel buffers and the inverse diagonal blocks
    DevBuf<int> stats, perm, tflags;
    int ts_epoch = 0;        // epoch of the data-flow flags of dense_trisolve_kernel
    int* h_stats = nullptr;  // pinned: inertia / perturbation counters of the last static factorisation
    ~pipsgpu_sc();
    DevBuf<int8_t> ptype;
    DevBuf<BkState> bk;
    // sparse mode (pipsgpu_sc_set_sparse_pattern; sLinsysRootAug with hasSparseKkt): the root matrix restricted to a
    // precomputed lower CSR pattern is factorised by the sparse multifrontal LDL^T in a private context on the same stream
    pipsgpu_ctx* sp_ctx = nullptr;
    pipsgpu_leaf* sp_leaf = nullptr;
    DevBuf<int64_t> sp_idx;    // per pattern entry: its position in the dense assembly buffer A
    DevBuf<double> sp_tmp;
    int64_t sp_nnz = 0;
    bool used_sparse = false;  // last factorisation went through the sparse solver
    bool factorized = false;
    bool used_bk = false;      // last factorisation needed the Bunch-Kaufman path
    int inertia[3] = {0, 0, 0};
    int n_perturbed = 0;
};

struct pipsgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::vector<pipsgpu_leaf*> leaves;
    bool committed = false;
    NcclComm comm = nullptr;
    int rank = 0, nranks = 1;

    // pools
    int64_t npos = 0, nfronts = 0, Ltot = 0, Ctot = 0, Wtot = 0, updtot = 0, nnzK = 0, nnzB = 0;
    int nlevels = 0;
    // The numeric factorisation runs over BATCHES of consecutive leaves: the contribution pool C, the panel scratch W and
    // the X = L11^-1 scratch are sized for the largest batch and reused (C alone would be 160 GB for the 512 blocks of C3);
    // the factor panels L and the solve pool stay resident for every leaf. batch_ptr[b] .. batch_ptr[b+1] = leaves of batch b.
    std::vector<int> batch_ptr;
    // Group (Schur-term) SYRKs: one launch per (batch, tree level) = the groups whose root front sits on that level, issued on
    // stream_grp as soon as the level is done. ev_gt brackets every launch (2 events per (batch, level), gt_rec = recorded in
    // the current factorisation), ev_grp_done[b] marks the end of batch b's launches.
    std::vector<cudaEvent_t> ev_gt, ev_grp_done;
    std::vector<char> gt_rec;
    std::vector<char> grp_pending;      // per batch: group kernels run on stream_grp and have not been joined yet
    cudaEvent_t ev_lvl = nullptr;
    std::vector<int> level_ptr;     // nlevels*NCLS+1 into level_fronts; group g = level_order*NCLS + width class, deepest level first
    std::vector<int64_t> tile_ptr;  // nlevels+1 into tiles
    DevBuf<TileDesc> tiles;
    // deferred border x border updates (kernels.cuh: GroupDesc)
    DevBuf<GroupDesc> groups;
    DevBuf<GroupMember> gmembers;
    DevBuf<int> ginv;
    DevBuf<int8_t> gflags;
    bool group_by_level = true;     // PIPSGPU_GROUP_BY_LEVEL=0: all group tiles of a batch in one launch after its last level
    int bp_threads = 256;           // PIPSGPU_BP_THREADS=128: smaller CTAs for the big-front panels
    DevBuf<GTile> gtiles;
    int64_t ngtiles = 0;
    bool grouping = true;           // PIPSGPU_GROUP=0: every big front flushes its own border x border update
    double group_frac = 0.5;        // a front joins its parent's group if it has at least this share of the root's border rows
    int big_u = BIG_U;              // PIPSGPU_BIG_U: fronts with at least this many off-diagonal rows take the CTA-parallel path
    // side stream: the small fronts of a level run next to the big-front panel / tile kernels of the same level, and the
    // border-group SYRKs next to pack_solve (PIPSGPU_OVERLAP=0 keeps everything on one stream)
    bool overlap = true;
    // PIPSGPU_XFORM=1: the solve pool carries X = L11^-1 instead of L11, the streaming solve applies the pivot blocks as
    // two small matrix-vector products (no serial substitution chain); X scratch is then persistent per front
    bool xform = true;
    bool panel_staged = false;      // PIPSGPU_PANEL=staged
    bool group_pipe = false;        // PIPSGPU_GROUP_PIPE=1
    int group_occ = 4;              // PIPSGPU_GROUP_OCC
    cudaStream_t stream2 = nullptr;
    cudaStream_t stream_grp = nullptr;  // group (Schur-term) SYRKs of batch b run here next to the tree levels of batch b + 1
    cudaStream_t stream_hi = nullptr;   // high-priority side stream: look-ahead panels of the dense root LDL^T
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    DevBuf<double> XY;
    int64_t XYtot = 0;
    DevBuf<double> L, C, W, Dd, De, Did, Die, Kval, Bval, x, upd, dvals, work;
    std::vector<int> fl_ptr;        // nlevels*2+1 into fl_fronts; group = level_order*2 + (0 small | 1 big)
    std::vector<int> fl_umax;       // per level (deepest first): most off-diagonal rows of a big front (grid of big_panel_kernel)
    DevBuf<int> fl_fronts;
    DevBuf<int> rowg, relidx, invmap, child_list, level_fronts, gidx, bg_row, bg_sc, bg_b0, stats;
    DevBuf<int8_t> ptype, sgn;
    DevBuf<FrontDesc> fronts;
    DevBuf<LeafDev> leafdev;
    DevBuf<int64_t> kmap, bmap, diagpos;
    DevBuf<int> kcoo_row, kcoo_col;    // lower COO of every K_i in global (original-order) positions: device KKT matvec
    // iterative refinement of the leaf solves, applied like PARDISO's iparm[7] = 2 (PardisoProjectSolver.C:72) only when the
    // factorisation perturbed pivots: up to refine_steps steps  z += K^-1 (b - K z)  with the resident matrix values.
    // PIPSGPU_REFINE=<steps> (0 disables), PIPSGPU_REFINE_ALWAYS=1 refines after every factorisation.
    // The refinement belongs to the plug-in solve (DoubleLinearSolver::solve = pipsgpu_leaf_solve), as in the reference. Inside the
    // batched solveCompressed it is OFF by default (PIPSGPU_REFINE_KKT=1 enables it): the Schur complement was formed with the
    // perturbed factors, so unrefined leaf solves make solveCompressed the exact solve of ONE consistent (perturbed) KKT
    // system, while refined leaf solves mix K and K + E and measure worse (tests/test_gpu_outer_solve.py); the outer
    // BiCGStab (pipsgpu_kkt_solve) removes the perturbation error either way.
    int refine_steps = 2;
    bool refine_always = false;
    bool refine_kkt = false;
    bool any_perturbed = false;
    // streaming persistent solve (psolve.cu): solve pool + per-front descriptors in sweep order
    DevBuf<SolveFront> sfront;
    DevBuf<SolveUnit> sunits;
    DevBuf<double> SL, acc;
    DevBuf<int> ucnt, uflag, ticket, chunk_front_l, chunk_idx_l;   // pack chunks (front, chunk index) grouped by leaf
    DevBuf<uint16_t> slots;
    int nchunks = 0, nunits = 0, nsub_units = 0;
    int64_t SLtot = 0;
    int ps_epoch = 0, ps_epoch_max = 1 << 16;
    int sm_count = 148;
    double* h_dvals = nullptr;      // pinned staging of all leaf diagonals (npos)
    bool own_stream = true;

    FactorArgs factor_args(pipsgpu_sc* sc) {
        FactorArgs a;
        a.fronts = fronts.p; a.level_fronts = level_fronts.p; a.fl_fronts = fl_fronts.p; a.child_list = child_list.p;
        a.rowg = rowg.p; a.relidx = relidx.p; a.invmap = invmap.p; a.ptype = ptype.p; a.sgn = sgn.p; a.leaves = leafdev.p; a.L = L.p; a.C = C.p;
        a.W = W.p; a.XY = XY.p;
        a.Dd = Dd.p; a.De = De.p; a.Did = Did.p; a.Die = Die.p;
        a.SC = sc ? sc->A.p : nullptr; a.ldsc = sc ? sc->n : 0; a.stats = stats.p; a.tiles = tiles.p;
        a.groups = groups.p; a.gmembers = gmembers.p; a.ginv = ginv.p; a.gflags = gflags.p; a.gtiles = gtiles.p; a.panel_staged = panel_staged ? 1 : 0; a.group_pipe = group_pipe ? 1 : 0; a.group_occ = group_occ;
        a.bp_threads = bp_threads;
        return a;
    }
    // one launch = out (=|-=) K^-1 in for every leaf of the context (original ordering on both sides)
    void psolve(const double* in, double* out, int out_mode) {
        if (nunits <= 0) return;
        if (solve_by_level) { psolve_by_level(in, out, out_mode); return; }
        if (ps_epoch >= ps_epoch_max) {   // counters / flags are monotone in the epoch: re-arm long before they wrap
            cudaMemsetAsync(ucnt.p, 0, sizeof(int) * (size_t)nunits, stream);
            cudaMemsetAsync(uflag.p, 0, sizeof(int) * (size_t)nunits, stream);
            ps_epoch = 0;
        }
        PStreamArgs a;
        a.sf = sfront.p; a.units = sunits.p; a.nunits = nunits; a.SL = SL.p; a.rowg = rowg.p; a.gidx = gidx.p;
        a.in = in; a.out = out; a.x = x.p; a.acc = acc.p; a.ucnt = ucnt.p; a.uflag = uflag.p; a.ticket = ticket.p;
        a.epoch = ps_epoch; a.out_mode = out_mode; a.xform = xform ? 1 : 0;
        launch_psolve_stream(a, sm_count, stream);
        ps_epoch += 1;
    }
    void psolve_by_level(const double* in, double* out, int out_mode);   // diagnostic path, PIPSGPU_SOLVE=level
    bool solve_by_level = false;
    // factor panels -> solve pool (after every numeric factorisation)
    // chunks [begin, begin + count) of the by-leaf chunk list (a batch or a single leaf repacks its own chunks only)
    void pack_solve(int begin, int count) {
        launch_pack_solve(sfront.p, chunk_front_l.p + begin, chunk_idx_l.p + begin, count, fronts.p, L.p, Did.p, Die.p, ptype.p,
                          slots.p, SL.p, xform ? XY.p : nullptr, stream);
    }
    SolveArgs solve_args(double* xbuf, int64_t xstride, double* updbuf, int64_t updstride) {
        SolveArgs a;
        a.fronts = fronts.p; a.level_fronts = level_fronts.p; a.child_list = child_list.p; a.rowg = rowg.p;
        a.relidx = relidx.p; a.ptype = ptype.p; a.L = L.p; a.Did = Did.p; a.Die = Die.p;
        a.x = xbuf; a.upd = updbuf; a.x_stride = xstride; a.upd_stride = updstride;
        return a;
    }
};

struct pipsgpu_kkt {
    pipsgpu_ctx* ctx = nullptr;
    int n0 = 0, my0 = 0, mz0 = 0, myl = 0, mzl = 0;
    CsrDev A0, C0, F0, G0;
    std::vector<pipsgpu_leaf*> leaves;
    std::unique_ptr<pipsgpu_sc> sc;
    bool ready = false;
    DevBuf<double> dd0, zdiag0, zdiagl, b0, r, t3, bl, tl, b0_in, bl_in, wl;
    DevBuf<double> dd_in, nz_in;        // the caller's concatenated leaf diagonals of the current factorisation
    DevBuf<LeafDiag> ldiag;             // per leaf: where its diagonals sit (launch_build_dvals)
    std::vector<LeafDiag> h_ldiag;
    double* h_root = nullptr;           // pinned staging of the root diagonals (n0 + mz0 + mzl)
    // device BiCGStab (pipsgpu_kkt_solve): whole-tree vectors [root part (nb0, padded) | leaf part (N)], allocated on first use
    DevBuf<double> kv[9], kc0, kpartial, kscal;
    double* h_scal = nullptr;       // pinned
    int64_t kv_root = 0;            // padded length of the root part
    ~pipsgpu_kkt() { if (h_scal) cudaFreeHost(h_scal); if (h_root) cudaFreeHost(h_root); }
    bool have_diag = false, have_rhs = false;
    double last_regDy = 0.0;
    int64_t sum_nx = 0, sum_nz = 0, sum_n = 0;
    double times[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaEvent_t ev[16] = {};
    double ktimes[4] = {0, 0, 0, 0};
    bool ev_init = false;
};

namespace {

void require_device(pipsgpu_ctx* ctx) {
    if (!ctx) throw Error(PIPSGPU_ERR_INVALID, "null context");
    CUDA_CHECK(cudaSetDevice(ctx->device));
}

void analyze_one(pipsgpu_leaf* lf) {
    if (lf->analyzed) return;
    SymbolicOptions opt;
    opt.max_width = MAXW;
    analyze_leaf(lf->n, lf->nx, lf->ny, lf->nz, lf->krow.data(), lf->jcol.data(), lf->val.data(), lf->nB,
                 (int64_t)lf->brow.size(), lf->brow.data(), lf->bcol.data(), opt, lf->S);
    double kn = 0.0;
    for (int i = 0; i < lf->n; ++i)
        for (int p = lf->krow[i]; p < lf->krow[i + 1]; ++p)
            if (lf->jcol[p] != i) kn = std::max(kn, std::fabs(lf->val[p]));
    lf->knorm = std::max(kn, 1.0);
    build_border_groups(lf->S, lf->ctx ? lf->ctx->big_u : BIG_U, lf->ctx ? lf->ctx->group_frac : 0.6, lf->G);
    lf->analyzed = true;
}

// ---- batched leaf solve on the context buffers: x (permuted order) <- K^-1 x for the leaves in [range of levels] ----
void run_leaf_solve(pipsgpu_ctx* ctx, const SolveArgs& sa, int nrhs, pipsgpu_leaf* only, int64_t npos_range_begin,
                    int64_t npos_range_count) {
    cudaStream_t s = ctx->stream;
    const int ngroups = ctx->nlevels * NCLS;
    for (int g = 0; g < ngroups; ++g) {   // group 0 = deepest level
        int begin = ctx->level_ptr[g], count = ctx->level_ptr[g + 1] - ctx->level_ptr[g];
        if (only) { begin = only->level_begin[g]; count = only->level_count[g]; }
        launch_fwd_level(sa, g % NCLS, begin, count, nrhs, s);
    }
    {
        SolveArgs d = sa;
        // diag kernel indexes positions from 0: shift the base pointers to the requested range
        d.x = sa.x + npos_range_begin;
        d.ptype = sa.ptype + npos_range_begin;
        d.Did = sa.Did + npos_range_begin;
        d.Die = sa.Die + npos_range_begin;
        launch_diag_solve(d, npos_range_count, nrhs, s);
    }
    for (int g = ngroups - 1; g >= 0; --g) {
        int begin = ctx->level_ptr[g], count = ctx->level_ptr[g + 1] - ctx->level_ptr[g];
        if (only) { begin = only->level_begin[g]; count = only->level_count[g]; }
        launch_bwd_level(sa, g % NCLS, begin, count, nrhs, s);
    }
}

// All levels of the numeric factorisation (deepest first) of the leaves [lo, hi) - one batch, or a single leaf for the
// plug-in entry points -, then their deferred border-group updates and their solve-pool pack. Fronts, tiles, group tiles
// and pack chunks are all stored grouped by leaf inside a level, so a leaf range is a contiguous sub-range of every list.
// the main stream waits for the group kernel of batch b (no-op if it ran on the main stream or was joined already)
void join_group(pipsgpu_ctx* ctx, int b) {
    if (b < 0 || b >= (int)ctx->grp_pending.size() || !ctx->grp_pending[(size_t)b]) return;
    CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->ev_grp_done[(size_t)b], 0));
    ctx->grp_pending[(size_t)b] = 0;
}

void run_factor_range(pipsgpu_ctx* ctx, pipsgpu_sc* sc, int lo, int hi, cudaEvent_t e_pack0, cudaEvent_t e_pack1,
                      bool defer_join = false) {
    if (lo >= hi) return;
    cudaStream_t s = ctx->stream;
    cudaStream_t s2 = ctx->overlap ? ctx->stream2 : nullptr;
    FactorArgs fa = ctx->factor_args(sc);
    const pipsgpu_leaf* first = ctx->leaves[(size_t)lo];
    const pipsgpu_leaf* last = ctx->leaves[(size_t)hi - 1];
    const bool all = lo == 0 && hi == (int)ctx->leaves.size();
    // border x border updates of the grouped big fronts (they feed nothing inside the tree, only the Schur complement):
    // read-only on the factor panels and on the batch's persistent W21 scratch. The groups whose root sits on level lv are
    // complete once lv is done; they are launched on their own stream right then and run next to the upper levels of this
    // batch, its solve-pool pack and - when the caller defers the join - the levels of the NEXT batch. The tree kernels are
    // latency-bound and the group SYRK keeps the FP64 tensor pipe about a third busy, so the two overlap well.
    const int b = first->batch;
    cudaStream_t sg = (sc && ctx->overlap && ctx->stream_grp) ? ctx->stream_grp : s;
    bool grp_any = false;
    auto launch_groups = [&](int lv_lo, int lv_hi) {      // group tiles of the levels [lv_lo, lv_hi): contiguous per leaf range
        if (!sc || ctx->ngtiles == 0) return;
        for (int lv = lv_lo; lv < lv_hi; ++lv) {
            const int64_t gb = first->gtile_begin[(size_t)lv];
            const int64_t gc = last->gtile_begin[(size_t)lv] + last->gtile_count[(size_t)lv] - gb;
            if (gc <= 0) continue;
            if (sg != s) {
                CUDA_CHECK(cudaEventRecord(ctx->ev_lvl, s));
                CUDA_CHECK(cudaStreamWaitEvent(sg, ctx->ev_lvl, 0));
            }
            const size_t e = ((size_t)b * ctx->nlevels + lv) * 2;
            CUDA_CHECK(cudaEventRecord(ctx->ev_gt[e], sg));
            launch_border_groups(fa, gb, gc, sg);
            CUDA_CHECK(cudaEventRecord(ctx->ev_gt[e + 1], sg));
            ctx->gt_rec[e / 2] = 1;
            grp_any = true;
        }
    };
    for (int lv = 0; lv < ctx->nlevels; ++lv) {   // lv 0 = deepest level; children always sit in earlier launches
        for (int c = 0; c < NCLS; ++c) {      // one launch per width class (shared-memory tiles sized by the class)
            const int g = lv * NCLS + c;
            const int begin = all ? ctx->level_ptr[g] : first->level_begin[g];
            const int count = all ? ctx->level_ptr[g + 1] - begin : last->level_begin[g] + last->level_count[g] - begin;
            launch_pivot(fa, begin, count, c, s);
        }
        int begin[2], count[2];
        for (int kind = 0; kind < 2; ++kind) {
            const int g = lv * 2 + kind;
            begin[kind] = first->fl_begin[g];
            count[kind] = last->fl_begin[g] + last->fl_count[g] - begin[kind];
        }
        const int64_t tb = first->tile_begin[lv], tc = last->tile_begin[lv] + last->tile_count[lv] - tb;
        // small and big fronts of one level are independent of each other (disjoint panels, scratch and contribution
        // blocks): the latency-bound one-warp-per-front kernel runs next to the CTA-parallel panel + tile kernels
        const bool fork = s2 && count[0] > 0 && count[1] > 0;
        cudaStream_t sb = fork ? s2 : s;
        if (fork) {
            CUDA_CHECK(cudaEventRecord(ctx->ev_fork, s));
            CUDA_CHECK(cudaStreamWaitEvent(s2, ctx->ev_fork, 0));
        }
        launch_small_fronts(fa, begin[0], count[0], s);
        launch_big_panel(fa, begin[1], count[1], ctx->fl_umax[(size_t)lv], sb);
        launch_front_update_tiles(fa, tb, tc, sb);
        if (fork) {
            CUDA_CHECK(cudaEventRecord(ctx->ev_join, s2));
            CUDA_CHECK(cudaStreamWaitEvent(s, ctx->ev_join, 0));
        }
        if (ctx->group_by_level) launch_groups(lv, lv + 1);
    }
    if (!ctx->group_by_level) launch_groups(0, ctx->nlevels);
    if (grp_any) {
        if (sg != s) CUDA_CHECK(cudaEventRecord(ctx->ev_grp_done[(size_t)b], sg));
        ctx->grp_pending[(size_t)b] = (sg != s);
    }
    if (e_pack0) CUDA_CHECK(cudaEventRecord(e_pack0, s));
    ctx->pack_solve(first->chunk_begin, last->chunk_begin + last->chunk_count - first->chunk_begin);
    if (e_pack1) CUDA_CHECK(cudaEventRecord(e_pack1, s));
    if (!defer_join) join_group(ctx, b);
}

// numeric factorisation of one leaf (plug-in entry points) or of every leaf of the context, batch by batch.
// e_pack0 / e_pack1 (optional) bracket the pack kernel of the LAST batch on the main stream.
void run_factor_levels(pipsgpu_ctx* ctx, pipsgpu_sc* sc, pipsgpu_leaf* only, cudaEvent_t e_pack0 = nullptr,
                       cudaEvent_t e_pack1 = nullptr) {
    if (only) { run_factor_range(ctx, sc, only->index, only->index + 1, e_pack0, e_pack1); return; }
    // The persistent W21 scratch of the grouped fronts is double-buffered by batch parity (pipsgpu_ctx_commit): the group
    // kernel of batch b may still read it while batch b + 1 factorises; batch b + 2 reuses it and waits for that kernel.
    const int nb = (int)ctx->batch_ptr.size() - 1;
    for (int b = 0; b < nb; ++b) {
        const bool lastb = b + 1 == nb;
        join_group(ctx, b - 2);
        run_factor_range(ctx, sc, ctx->batch_ptr[(size_t)b], ctx->batch_ptr[(size_t)b + 1], lastb ? e_pack0 : nullptr,
                         lastb ? e_pack1 : nullptr, true);
    }
    join_group(ctx, nb - 2);
    join_group(ctx, nb - 1);
}

}  // namespace

// Diagnostic only (PIPSGPU_SOLVE=level): the same K^-1 application through the per-level kernels that pipsgpu_leaf_solve
// uses, to tell a defect of the streaming solve kernel from a defect of the factors on a GPU box.
void pipsgpu_ctx::psolve_by_level(const double* in, double* out, int out_mode) {
    launch_gather_perm(gidx.p, in, x.p, npos, 1, npos, stream);
    SolveArgs sa = solve_args(x.p, npos, upd.p, std::max<int64_t>(updtot, 1));
    run_leaf_solve(this, sa, 1, nullptr, 0, npos);
    if (out_mode == 0) launch_scatter_perm(gidx.p, x.p, out, npos, 1, npos, stream);
    else {
        launch_scatter_perm(gidx.p, x.p, work.p, npos, 1, npos, stream);
        launch_axpy(-1.0, work.p, out, npos, stream);
    }
}

extern "C" {
static void factor_prepare(pipsgpu_ctx* ctx, pipsgpu_leaf* only);      // defined with the leaf entry points below
}

namespace {

void fetch_leaf_stats(pipsgpu_ctx* ctx) {
    std::vector<int> h(4 * ctx->leaves.size());
    if (h.empty()) return;
    CUDA_CHECK(cudaMemcpyAsync(h.data(), ctx->stats.p, sizeof(int) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < ctx->leaves.size(); ++i) {
        pipsgpu_leaf* lf = ctx->leaves[i];
        lf->inertia[0] = h[4 * i]; lf->inertia[1] = h[4 * i + 1]; lf->inertia[2] = h[4 * i + 2];
        lf->n_perturbed = h[4 * i + 3];
    }
    bool any = false;
    for (auto* lf : ctx->leaves) any = any || lf->n_perturbed > 0;
    ctx->any_perturbed = any;
}

// static-order blocked LDL^T (x0 block first; exact for quasi-definite Schur complements): right-looking with panels of
// 128 columns. Per panel: ldlt_diag_kernel (diagonal block + its inverse), one GEMM for the panel below it
// (W21 = A21 X^T, L21 = W21 D^-1) and the trailing update A22 -= L21 W21^T on the FP64 tensor pipe, all three fed by TMA.
// Look-ahead: the first block column of the trailing update (c1) is done first; the next panel's diagonal block and panel
// GEMM then run on the high-priority side stream next to the rest of the trailing update (c2).
void sc_factorize_static(pipsgpu_sc* sc) {
    pipsgpu_ctx* ctx = sc->ctx;
    cudaStream_t s = ctx->stream, s2 = ctx->stream_hi;
    const int n = sc->n;
    const int64_t ld = sc->ldf;
    CUDA_CHECK(cudaMemsetAsync(sc->stats.p, 0, sizeof(int) * 4, s));
    CUDA_CHECK(cudaMemsetAsync(sc->ptype.p, 0, sizeof(int8_t) * (size_t)(n + 1), s));
    CUDA_CHECK(cudaMemsetAsync(sc->Dinv_e.p, 0, sizeof(double) * (size_t)(n + 1), s));
    launch_iota(sc->perm.p, n, s);
    // pivots below 1e-12 * (largest entry of their own original row) are perturbed, in the spirit of PARDISO's pivot
    // perturbation (iparm[9], PardisoProjectIndefSolver.C:34-61) but scale-aware; the oracle applies the same rule
    const double eps_rel = 1e-12;
    launch_row_absmax(sc->A.p, n, n, sc->scale.p, s);
    CUDA_CHECK(cudaMemcpy2DAsync(sc->Fm.p, sizeof(double) * (size_t)ld, sc->A.p, sizeof(double) * (size_t)n, sizeof(double) * (size_t)n,
                                 (size_t)n, cudaMemcpyDeviceToDevice, s));
    const int NB = DENSE_PANEL;
    const int npanels = (n + NB - 1) / NB;
    auto diag_and_panel = [&](int k, cudaStream_t st) {
        const int kb = k * NB, bw = std::min(NB, n - kb);
        launch_ldlt_diag(sc->Fm.p, ld, kb, bw, sc->n0, eps_rel, sc->scale.p, sc->D.p, sc->Dinv.p,
                         sc->Xall.p + (size_t)k * NB * NB, sc->stats.p, st);
        const int rows = n - kb - bw;
        if (rows <= 0) return;
        GemmArgs g{};
        g.a_row0 = kb + bw; g.a_col0 = kb; g.b_row0 = k * NB; g.b_col0 = 0; g.K = NB;
        g.m_tiles = (rows + GEMM_TILE - 1) / GEMM_TILE; g.n_tiles = 1; g.lower = 0;
        g.C = sc->Fm.p; g.ldc = ld; g.c_row0 = kb + bw; g.c_col0 = kb; g.c_rows = n; g.c_cols = bw;
        g.C2 = sc->Wbuf.p + (size_t)(k & 1) * (size_t)n * NB; g.ldc2 = NB; g.scale = sc->Dinv.p + kb; g.map = nullptr;
        launch_dgemm_nt(GEMM_EPI_PANEL, sc->tmF, sc->tmX, g, st);
    };
    auto trailing = [&](int k, bool first_column, cudaStream_t st) {
        const int kb = k * NB;
        const int t0 = kb + NB + (first_column ? 0 : NB);       // c1: block column t0 .. t0+127; c2: everything right of it
        const int rows = n - t0;
        if (rows <= 0) return;
        GemmArgs g{};
        g.a_row0 = t0; g.a_col0 = kb; g.b_row0 = first_column ? t0 : t0; g.b_col0 = 0; g.K = NB;
        g.m_tiles = (rows + GEMM_TILE - 1) / GEMM_TILE; g.n_tiles = 1; g.lower = first_column ? 0 : 1;
        g.C = sc->Fm.p; g.ldc = ld; g.c_row0 = t0; g.c_col0 = t0; g.c_rows = n; g.c_cols = n;
        g.C2 = nullptr; g.ldc2 = 0; g.scale = nullptr; g.map = nullptr;
        launch_dgemm_nt(GEMM_EPI_TRAIL, sc->tmF, sc->tmW[k & 1], g, st);
    };
    diag_and_panel(0, s);
    for (int k = 0; k + 1 < npanels; ++k) {
        trailing(k, true, s);
        const bool ahead = s2 != nullptr && (n - (k + 2) * NB) > 0;     // something is left for c2 to overlap with
        if (ahead) {
            CUDA_CHECK(cudaEventRecord(ctx->ev_fork, s));
            CUDA_CHECK(cudaStreamWaitEvent(s2, ctx->ev_fork, 0));
            diag_and_panel(k + 1, s2);
            CUDA_CHECK(cudaEventRecord(ctx->ev_join, s2));
            trailing(k, false, s);
            CUDA_CHECK(cudaStreamWaitEvent(s, ctx->ev_join, 0));
        } else {
            trailing(k, false, s);
            diag_and_panel(k + 1, s);
        }
    }
}

// Bunch-Kaufman LDL^T (dsytf2 restated, what the reference's dsytrf does): any symmetric indefinite matrix
void sc_factorize_bk(pipsgpu_sc* sc) {
    pipsgpu_ctx* ctx = sc->ctx;
    cudaStream_t s = ctx->stream;
    const int n = sc->n;
    CUDA_CHECK(cudaMemsetAsync(sc->bk.p, 0, sizeof(BkState), s));
    CUDA_CHECK(cudaMemsetAsync(sc->ptype.p, 0, sizeof(int8_t) * (size_t)(n + 1), s));
    CUDA_CHECK(cudaMemsetAsync(sc->Dinv_e.p, 0, sizeof(double) * (size_t)(n + 1), s));
    launch_iota(sc->perm.p, n, s);
    const int ld = (int)sc->ldf;
    CUDA_CHECK(cudaMemcpy2DAsync(sc->Fm.p, sizeof(double) * (size_t)ld, sc->A.p, sizeof(double) * (size_t)n, sizeof(double) * (size_t)n,
                                 (size_t)n, cudaMemcpyDeviceToDevice, s));
    for (int step = 0; step < n; ++step)
        launch_bk_step(sc->Fm.p, ld, n, step, sc->perm.p, sc->ptype.p, sc->D.p, sc->De.p, sc->Dinv.p, sc->Dinv_e.p, sc->Wc.p,
                       sc->bk.p, 1e-12, sc->scale.p, s);
}

// factorize: fast static path first; if a pivot had to be perturbed or the inertia is not (n0, n-n0, 0) the matrix is
// not quasi-definite in this order -> redo with Bunch-Kaufman. Leaves sc->inertia / n_perturbed up to date.
// sparse root: gather the pattern values from the assembly buffer into the private context's CSR values and run the
// multifrontal LDL^T used for the leaves (static 1x1 / 2x2 pivots: x0 rows positive, y0 rows paired with x0 columns, link
// rows negative). Returns false (caller falls back to the dense solver, which can pivot) if a pivot had to be perturbed or
// the inertia is not (n0, n - n0, 0).
bool sc_factorize_sparse(pipsgpu_sc* sc) {
    pipsgpu_ctx* sp = sc->sp_ctx;
    cudaStream_t s = sp->stream;
    launch_gather_idx(sc->A.p, sc->sp_idx.p, sc->sp_nnz, sp->Kval.p, s);
    launch_absmax_eps(sp->Kval.p, sc->sp_nnz, 1e-12, &sp->leafdev.p->eps, s);
    factor_prepare(sp, nullptr);
    run_factor_levels(sp, nullptr, nullptr);
    fetch_leaf_stats(sp);
    pipsgpu_leaf* lf = sc->sp_leaf;
    lf->factorized = true;
    sc->inertia[0] = lf->inertia[0]; sc->inertia[1] = lf->inertia[1]; sc->inertia[2] = lf->inertia[2];
    sc->n_perturbed = lf->n_perturbed;
    return lf->n_perturbed == 0 && lf->inertia[0] == sc->n0 && lf->inertia[2] == 0;
}

void sc_factorize_impl(pipsgpu_sc* sc) {
    pipsgpu_ctx* ctx = sc->ctx;
    sc->used_sparse = false;
    if (sc->sp_ctx && sc_factorize_sparse(sc)) {
        sc->used_sparse = true;
        sc->used_bk = false;
        sc->factorized = true;
        return;
    }
    sc_factorize_static(sc);
    int* h = sc->h_stats;
    CUDA_CHECK(cudaMemcpyAsync(h, sc->stats.p, sizeof(int) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    sc->inertia[0] = h[0]; sc->inertia[1] = h[1]; sc->inertia[2] = h[2]; sc->n_perturbed = h[3];
    sc->used_bk = false;
    if (h[3] != 0 || h[0] != sc->n0 || h[2] != 0) {
        sc_factorize_bk(sc);
        BkState st;
        CUDA_CHECK(cudaMemcpyAsync(&st, sc->bk.p, sizeof(BkState), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        sc->inertia[0] = st.npos; sc->inertia[1] = st.nneg; sc->inertia[2] = 0; sc->n_perturbed = st.npert;
        sc->used_bk = true;
    }
    sc->factorized = true;
}

// x <- SC^-1 x for nrhs device-resident right-hand sides (stride xstride): blocked multi-CTA solve with the static-order
// factors, the single-CTA permuted solve after a Bunch-Kaufman factorisation
void sc_solve_device(pipsgpu_sc* sc, double* x, int nrhs, int64_t xstride) {
    pipsgpu_ctx* ctx = sc->ctx;
    cudaStream_t s = ctx->stream;
    if (sc->used_sparse) {
        sc->sp_tmp.ensure((size_t)sc->n + 1);
        for (int r = 0; r < nrhs; ++r) {
            double* xr = x + (size_t)r * xstride;
            sc->sp_ctx->psolve(xr, sc->sp_tmp.p, 0);
            CUDA_CHECK(cudaMemcpyAsync(xr, sc->sp_tmp.p, sizeof(double) * (size_t)sc->n, cudaMemcpyDeviceToDevice, s));
        }
        return;
    }
    if (sc->used_bk) {
        launch_dense_solve(sc->Fm.p, (int)sc->ldf, sc->n, sc->perm.p, sc->ptype.p, sc->Dinv.p, sc->Dinv_e.p, x, nrhs, xstride,
                           sc->swork.p, s);
        return;
    }
    const int nb = sc->n / 64 + 2;
    for (int r = 0; r < nrhs; ++r) {
        if (sc->ts_epoch >= (1 << 30)) {
            CUDA_CHECK(cudaMemsetAsync(sc->tflags.p, 0, sizeof(int) * sc->tflags.n, s));
            sc->ts_epoch = 0;
        }
        sc->ts_epoch += 1;
        launch_dense_trisolve(sc->Fm.p, sc->ldf, sc->n, sc->Xall.p, sc->Dinv.p, x + (int64_t)r * xstride, sc->ybuf.p, sc->tflags.p,
                              sc->tflags.p + nb, sc->ts_epoch, ctx->sm_count, s);
    }
}

void sc_allreduce_impl(pipsgpu_sc* sc) {
    pipsgpu_ctx* ctx = sc->ctx;
    if (!ctx->comm || ctx->nranks <= 1) return;
    if (sc->sp_ctx) {
        // reduceKKTsparse (LS/DistributedRootLinearSystem.C:885-925): only the pattern values travel
        sc->pack.ensure((size_t)sc->sp_nnz);
        launch_gather_idx(sc->A.p, sc->sp_idx.p, sc->sp_nnz, sc->pack.p, ctx->stream);
        NcclApi& api = NcclApi::get();
        const int rc = api.AllReduce(sc->pack.p, sc->pack.p, (size_t)sc->sp_nnz, kNcclFloat64, kNcclSum, ctx->comm, ctx->stream);
        if (rc != 0) throw Error(PIPSGPU_ERR_NCCL, std::string("ncclAllReduce(sparse SC): ") + api.GetErrorString(rc));
        launch_scatter_idx(sc->A.p, sc->sp_idx.p, sc->sp_nnz, sc->pack.p, ctx->stream);
        return;
    }
    const int nl = sc->myl + sc->mzl;
    const int64_t cnt = sc_pack_size(sc->n0, nl);
    sc->pack.ensure((size_t)cnt);
    launch_sc_pack(sc->A.p, sc->n, sc->n0, sc->my0, nl, sc->pack.p, false, nullptr, ctx->stream);
    NcclApi& api = NcclApi::get();
    const int rc = api.AllReduce(sc->pack.p, sc->pack.p, (size_t)cnt, kNcclFloat64, kNcclSum, ctx->comm, ctx->stream);
    if (rc != 0) throw Error(PIPSGPU_ERR_NCCL, std::string("ncclAllReduce: ") + api.GetErrorString(rc));
    launch_sc_pack(nullptr, sc->n, sc->n0, sc->my0, nl, sc->pack.p, true, sc->A.p, ctx->stream);
}

}  // namespace

// ============================================================================================================
// C ABI
// ============================================================================================================
extern "C" {

int pipsgpu_ctx_create(int device, pipsgpu_ctx** out) {
    PIPSGPU_TRY
    if (!out) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_ctx_create: null output");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        throw Error(PIPSGPU_ERR_NO_DEVICE, "pipsgpu_ctx_create: no CUDA device available (there is no CPU fallback)");
    if (device < 0 || device >= ndev) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_ctx_create: invalid device index");
    CUDA_CHECK(cudaSetDevice(device));
    auto ctx = std::make_unique<pipsgpu_ctx>();
    ctx->device = device;
    CUDA_CHECK(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
    CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    // function attributes (dynamic shared-memory opt-in) are per device: set them for every context's device
    CUDA_CHECK(init_psolve_attributes());
    CUDA_CHECK(init_kernel_attributes());
    CUDA_CHECK(init_dense_attributes());
    if (const char* e = std::getenv("PIPSGPU_SOLVE")) {
        ctx->solve_by_level = std::string(e) == "level";
    }
    if (const char* e = std::getenv("PIPSGPU_REFINE")) ctx->refine_steps = std::max(0, std::min(8, std::atoi(e)));
    if (const char* e = std::getenv("PIPSGPU_REFINE_ALWAYS")) ctx->refine_always = std::string(e) != "0";
    if (const char* e = std::getenv("PIPSGPU_REFINE_KKT")) ctx->refine_kkt = std::string(e) != "0";
    if (const char* e = std::getenv("PIPSGPU_GROUP")) ctx->grouping = std::string(e) != "0";
    if (const char* e = std::getenv("PIPSGPU_GROUP_FRAC")) ctx->group_frac = std::min(1.0, std::max(0.05, std::atof(e)));
    if (const char* e = std::getenv("PIPSGPU_BIG_U")) ctx->big_u = std::min(1 << 20, std::max(32, std::atoi(e)));
    if (const char* e = std::getenv("PIPSGPU_OVERLAP")) ctx->overlap = std::string(e) != "0";
    if (const char* e = std::getenv("PIPSGPU_XFORM")) ctx->xform = std::string(e) != "0";
    if (const char* e = std::getenv("PIPSGPU_PANEL")) ctx->panel_staged = std::string(e) == "staged";
    if (const char* e = std::getenv("PIPSGPU_GROUP_PIPE")) ctx->group_pipe = std::string(e) != "0";
    if (const char* e = std::getenv("PIPSGPU_GROUP_OCC")) ctx->group_occ = std::max(4, std::min(6, std::atoi(e)));
    if (const char* e = std::getenv("PIPSGPU_GROUP_BY_LEVEL")) ctx->group_by_level = std::string(e) != "0";
    if (const char* e = std::getenv("PIPSGPU_BP_THREADS")) ctx->bp_threads = (std::atoi(e) == 128) ? 128 : 256;
    CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->stream_grp, cudaStreamNonBlocking));
    {
        int lo_prio = 0, hi_prio = 0;
        CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
        CUDA_CHECK(cudaStreamCreateWithPriority(&ctx->stream_hi, cudaStreamNonBlocking, hi_prio));
    }
    CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    *out = ctx.release();
    PIPSGPU_CATCH
}

int pipsgpu_ctx_destroy(pipsgpu_ctx* ctx) {
    PIPSGPU_TRY
    if (!ctx) return PIPSGPU_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->comm) NcclApi::get().CommDestroy(ctx->comm);
    for (auto* lf : ctx->leaves) delete lf;
    if (ctx->h_dvals) cudaFreeHost(ctx->h_dvals);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    if (ctx->stream2) { cudaStreamSynchronize(ctx->stream2); cudaStreamDestroy(ctx->stream2); }
    if (ctx->stream_grp) { cudaStreamSynchronize(ctx->stream_grp); cudaStreamDestroy(ctx->stream_grp); }
    if (ctx->stream_hi) { cudaStreamSynchronize(ctx->stream_hi); cudaStreamDestroy(ctx->stream_hi); }
    for (auto& e : ctx->ev_gt) cudaEventDestroy(e);
    for (auto& e : ctx->ev_grp_done) cudaEventDestroy(e);
    if (ctx->ev_lvl) cudaEventDestroy(ctx->ev_lvl);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    delete ctx;
    PIPSGPU_CATCH
}

int pipsgpu_ctx_set_stream(pipsgpu_ctx* ctx, void* stream) {
    PIPSGPU_TRY
    require_device(ctx);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)stream;
    ctx->own_stream = false;
    PIPSGPU_CATCH
}

void* pipsgpu_ctx_stream(pipsgpu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int pipsgpu_ctx_device(pipsgpu_ctx* ctx) { return ctx ? ctx->device : -1; }

int pipsgpu_ctx_synchronize(pipsgpu_ctx* ctx) {
    PIPSGPU_TRY
    require_device(ctx);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    PIPSGPU_CATCH
}

int64_t pipsgpu_ctx_launch_count(pipsgpu_ctx*) { return kernel_launch_count(); }

int pipsgpu_bench_dmma_peak(pipsgpu_ctx* ctx, double* tflops) {
    PIPSGPU_TRY
    require_device(ctx);
    if (!tflops) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_bench_dmma_peak: null output");
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, ctx->device));
    DevBuf<double> out;
    out.alloc(8);
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0));
    CUDA_CHECK(cudaEventCreate(&e1));
    const int ctas = prop.multiProcessorCount * 4, iters = 20000;
    launch_dmma_peak(out.p, 2000, ctas, ctx->stream);   // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        CUDA_CHECK(cudaEventRecord(e0, ctx->stream));
        const double fl = launch_dmma_peak(out.p, iters, ctas, ctx->stream);
        CUDA_CHECK(cudaEventRecord(e1, ctx->stream));
        CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    PIPSGPU_CATCH
}

int pipsgpu_bench_dgemm(pipsgpu_ctx* ctx, int n, double* tflops) {
    PIPSGPU_TRY
    require_device(ctx);
    if (!tflops || n < GEMM_TILE) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_bench_dgemm: invalid argument");
    n = (n / GEMM_TILE) * GEMM_TILE;
    DevBuf<double> A, C;
    A.alloc((size_t)n * n); C.alloc((size_t)n * n);
    CUDA_CHECK(cudaMemsetAsync(A.p, 0, sizeof(double) * (size_t)n * n, ctx->stream));
    CUDA_CHECK(cudaMemsetAsync(C.p, 0, sizeof(double) * (size_t)n * n, ctx->stream));
    TMap tm;
    make_tmap(&tm, A.p, n, n, n);
    GemmArgs g{};
    g.K = n; g.m_tiles = n / GEMM_TILE; g.n_tiles = g.m_tiles; g.lower = 1;
    g.C = C.p; g.ldc = n; g.c_rows = n; g.c_cols = n;
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0));
    CUDA_CHECK(cudaEventCreate(&e1));
    launch_dgemm_nt(GEMM_EPI_TRAIL, tm, tm, g, ctx->stream);      // warm-up
    double best = 0.0;
    const double fl = 2.0 * (double)(g.m_tiles * (g.m_tiles + 1) / 2) * GEMM_TILE * GEMM_TILE * (double)n;
    for (int rep = 0; rep < 3; ++rep) {
        CUDA_CHECK(cudaEventRecord(e0, ctx->stream));
        launch_dgemm_nt(GEMM_EPI_TRAIL, tm, tm, g, ctx->stream);
        CUDA_CHECK(cudaEventRecord(e1, ctx->stream));
        CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    PIPSGPU_CATCH
}

int pipsgpu_nccl_unique_id(char id[128]) {
    PIPSGPU_TRY
    NcclApi& api = NcclApi::get();
    if (!api.ok) throw Error(PIPSGPU_ERR_NCCL, api.err);
    NcclUniqueId uid;
    const int rc = api.GetUniqueId(&uid);
    if (rc != 0) throw Error(PIPSGPU_ERR_NCCL, std::string("ncclGetUniqueId: ") + api.GetErrorString(rc));
    std::memcpy(id, uid.internal, 128);
    PIPSGPU_CATCH
}

int pipsgpu_ctx_init_comm(pipsgpu_ctx* ctx, const char id[128], int rank, int nranks) {
    PIPSGPU_TRY
    require_device(ctx);
    NcclApi& api = NcclApi::get();
    if (!api.ok) throw Error(PIPSGPU_ERR_NCCL, api.err);
    NcclUniqueId uid;
    std::memcpy(uid.internal, id, 128);
    const int rc = api.CommInitRank(&ctx->comm, nranks, uid, rank);
    if (rc != 0) throw Error(PIPSGPU_ERR_NCCL, std::string("ncclCommInitRank: ") + api.GetErrorString(rc));
    ctx->rank = rank;
    ctx->nranks = nranks;
    PIPSGPU_CATCH
}

// ------------------------------------------------------------------------------------------------------------
// leaf
// ------------------------------------------------------------------------------------------------------------
int pipsgpu_leaf_create(pipsgpu_ctx* ctx, int n, int nx, int ny, int nz, const int* krow, const int* jcol, const double* val,
                        pipsgpu_leaf** out) {
    PIPSGPU_TRY
    if (!ctx || !krow || !jcol || !out) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_leaf_create: null argument");
    if (ctx->committed) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_leaf_create: context already committed");
    if (n < 0 || nx + ny + nz != n) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_leaf_create: nx+ny+nz != n");
    auto lf = std::make_unique<pipsgpu_leaf>();
    lf->ctx = ctx; lf->n = n; lf->nx = nx; lf->ny = ny; lf->nz = nz;
    lf->krow.assign(krow, krow + n + 1);
    const int64_t nnz = krow[n];
    lf->jcol.assign(jcol, jcol + nnz);
    if (val) lf->val.assign(val, val + nnz); else lf->val.assign(nnz, 0.0);
    lf->diag_entry.assign(n, -1);
    for (int i = 0; i < n; ++i)
        for (int p = krow[i]; p < krow[i + 1]; ++p) {
            if (jcol[p] > i) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_leaf_create: matrix must be lower triangular CSR");
            if (jcol[p] == i) lf->diag_entry[i] = p;
        }
    for (int i = 0; i < n; ++i)
        if (lf->diag_entry[i] < 0) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_leaf_create: every diagonal entry must be stored");
    lf->index = (int)ctx->leaves.size();
    ctx->leaves.push_back(lf.get());
    *out = lf.release();
    PIPSGPU_CATCH
}

int pipsgpu_leaf_set_border(pipsgpu_leaf* lf, int n0, int my0, int myl, int mzl, const int* A_krow, const int* A_jcol,
                            const double* A_val, const int* C_krow, const int* C_jcol, const double* C_val, const int* F_krow,
                            const int* F_jcol, const double* F_val, const int* G_krow, const int* G_jcol, const double* G_val) {
    PIPSGPU_TRY
    if (!lf) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_leaf_set_border: null leaf");
    if (lf->analyzed) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_leaf_set_border: leaf already analyzed");
    lf->n0 = n0; lf->my0 = my0; lf->myl = myl; lf->mzl = mzl;
    lf->nB = n0 + myl + mzl;
    lf->brow.clear(); lf->bcol.clear(); lf->bval.clear();
    auto add = [&](int r, int b, double v) { lf->brow.push_back(r); lf->bcol.push_back(b); lf->bval.push_back(v); };
    if (A_krow)   // rows of A live in the y-part of K, columns are linking variables
        for (int r = 0; r < lf->ny; ++r)
            for (int p = A_krow[r]; p < A_krow[r + 1]; ++p) {
                if (A_jcol[p] < 0 || A_jcol[p] >= n0) throw Error(PIPSGPU_ERR_INVALID, "border A: column out of range");
                add(lf->nx + r, A_jcol[p], A_val[p]);
            }
    if (C_krow)
        for (int r = 0; r < lf->nz; ++r)
            for (int p = C_krow[r]; p < C_krow[r + 1]; ++p) {
                if (C_jcol[p] < 0 || C_jcol[p] >= n0) throw Error(PIPSGPU_ERR_INVALID, "border C: column out of range");
                add(lf->nx + lf->ny + r, C_jcol[p], C_val[p]);
            }
    if (F_krow)   // F^T: border column n0+j has the entries of row j of F, living in the x-part of K
        for (int j = 0; j < myl; ++j)
            for (int p = F_krow[j]; p < F_krow[j + 1]; ++p) {
                if (F_jcol[p] < 0 || F_jcol[p] >= lf->nx) throw Error(PIPSGPU_ERR_INVALID, "border F: column out of range");
                add(F_jcol[p], n0 + j, F_val[p]);
            }
    if (G_krow)
        for (int j = 0; j < mzl; ++j)
            for (int p = G_krow[j]; p < G_krow[j + 1]; ++p) {
                if (G_jcol[p] < 0 || G_jcol[p] >= lf->nx) throw Error(PIPSGPU_ERR_INVALID, "border G: column out of range");
                add(G_jcol[p], n0 + myl + j, G_val[p]);
            }
    lf->border_set = true;
    PIPSGPU_CATCH
}

int pipsgpu_leaf_analyze(pipsgpu_leaf* lf) {
    PIPSGPU_TRY
    if (!lf) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_leaf_analyze: null leaf");
    analyze_one(lf);
    PIPSGPU_CATCH
}

int pipsgpu_ctx_commit(pipsgpu_ctx* ctx, int threads) {
    PIPSGPU_TRY
    require_device(ctx);
    if (ctx->committed) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_ctx_commit: already committed");
    const int nl = (int)ctx->leaves.size();
    // ---- symbolic analysis of all leaves on host threads ----
    {
        int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
        nt = std::max(1, std::min(nt, nl));
        std::atomic<int> next{0};
        std::mutex emu;
        std::string err;
        auto work = [&]() {
            for (;;) {
                const int i = next.fetch_add(1);
                if (i >= nl) break;
                try {
                    analyze_one(ctx->leaves[i]);
                } catch (const std::exception& e) {
                    std::lock_guard<std::mutex> g(emu);
                    err = e.what();
                }
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
        if (!err.empty()) throw Error(PIPSGPU_ERR_INVALID, "symbolic analysis failed: " + err);
    }
    // ---- placement ----
    int64_t npos = 0, nfr = 0, Ltot = 0, Ctot = 0, updtot = 0, nnzK = 0, nnzB = 0, nrows = 0, nchild = 0;
    int nlev = 0;
    // batches: consecutive leaves whose contribution blocks fit PIPSGPU_BATCH_GB (default 16 GB)
    double batch_gb = 16.0;
    if (const char* e = std::getenv("PIPSGPU_BATCH_GB")) batch_gb = std::max(0.001, std::atof(e));
    const int64_t c_cap = (int64_t)(batch_gb * 1e9 / 8.0);
    // ... and at least PIPSGPU_MIN_BATCHES batches when there are enough leaves (default 1: measured on C2 / C3 / C4, every
    // additional batch costs more in under-filled top-of-tree launches than the overlap of its group SYRKs with the next
    // batch's tree levels gains - profiles/README.md, round 2 session 8)
    int min_batches = 1;
    if (const char* e = std::getenv("PIPSGPU_MIN_BATCHES")) min_batches = std::max(1, std::atoi(e));
    const int leaf_cap = (nl >= 2 * min_batches) ? (nl + min_batches - 1) / min_batches : std::max(nl, 1);
    int leaves_in_batch = 0;
    int64_t cfill = 0;
    ctx->batch_ptr.assign(1, 0);
    std::vector<char> batch_start((size_t)nl + 1, 0);
    for (auto* lf : ctx->leaves) {
        const LeafSymbolic& S = lf->S;
        lf->pos_off = (int)npos; lf->front_base = (int)nfr;
        lf->Lbase = Ltot; lf->Lsize = S.Loff.empty() ? 0 : S.Loff.back();
        lf->Csize = S.Coff.empty() ? 0 : S.Coff.back();
        if (cfill > 0 && (cfill + lf->Csize > c_cap || leaves_in_batch >= leaf_cap)) {
            ctx->batch_ptr.push_back(lf->index);
            batch_start[(size_t)lf->index] = 1;
            cfill = 0;
            leaves_in_batch = 0;
        }
        ++leaves_in_batch;
        lf->batch = (int)ctx->batch_ptr.size() - 1;
        lf->Cbase = cfill;
        cfill += lf->Csize;
        Ctot = std::max(Ctot, cfill);
        lf->kval_off = nnzK; lf->bval_off = nnzB; lf->upd_base = updtot;
        int64_t us = 0;
        for (int s = 0; s < S.ns; ++s) us += S.sn_na[s];
        lf->upd_size = us;
        npos += lf->n; nfr += S.ns; Ltot += lf->Lsize; updtot += us;
        nnzK += (int64_t)lf->val.size(); nnzB += (int64_t)lf->bval.size();
        nrows += (int64_t)S.rowidx.size(); nchild += (int64_t)S.child_idx.size();
        nlev = std::max(nlev, S.nlevels);
        if (npos > 2000000000LL) throw Error(PIPSGPU_ERR_INVALID, "too many KKT rows for one context (int32 positions)");
    }
    ctx->batch_ptr.push_back(nl);
    ctx->grp_pending.assign(ctx->batch_ptr.size() - 1, 0);
    ctx->ev_grp_done.resize(ctx->batch_ptr.size() - 1);
    for (auto& e : ctx->ev_grp_done) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (!ctx->ev_lvl) CUDA_CHECK(cudaEventCreateWithFlags(&ctx->ev_lvl, cudaEventDisableTiming));
    ctx->npos = npos; ctx->nfronts = nfr; ctx->Ltot = Ltot; ctx->Ctot = Ctot; ctx->updtot = updtot;
    ctx->nnzK = nnzK; ctx->nnzB = nnzB; ctx->nlevels = nlev;

    std::vector<FrontDesc> fronts((size_t)nfr);
    std::vector<int> rowg((size_t)nrows), relidx((size_t)nrows), child_list((size_t)nchild), gidx((size_t)npos);
    std::vector<int8_t> ptype((size_t)npos), sgn((size_t)npos);
    std::vector<int64_t> kmap((size_t)nnzK), bmap((size_t)nnzB), diagpos((size_t)npos);
    std::vector<int> kcoo_row((size_t)nnzK), kcoo_col((size_t)nnzK);
    std::vector<double> kval((size_t)nnzK), bval((size_t)nnzB);
    std::vector<int> bg_row((size_t)nnzB), bg_sc((size_t)nnzB), bg_b0((size_t)nnzB);
    std::vector<LeafDev> leafdev((size_t)nl);
    // level lists: order index 0 = deepest level; within a level fronts are grouped by leaf
    const int ngroups = nlev * NCLS;
    auto group_of = [&](const LeafSymbolic& S, int s) {
        return (nlev - 1 - S.sn_level[s]) * NCLS + width_class(S.sn_col[s + 1] - S.sn_col[s]);
    };
    std::vector<int> level_count(ngroups, 0);
    for (auto* lf : ctx->leaves)
        for (int s = 0; s < lf->S.ns; ++s) ++level_count[group_of(lf->S, s)];
    ctx->level_ptr.assign(ngroups + 1, 0);
    for (int l = 0; l < ngroups; ++l) ctx->level_ptr[l + 1] = ctx->level_ptr[l] + level_count[l];
    std::vector<int> level_fronts((size_t)nfr), lfill(ctx->level_ptr.begin(), ctx->level_ptr.end() - 1);
    std::vector<int64_t> wfill(nlev, 0), xyfill(nlev, 0);
    int64_t wlev_max = 0, xy_max = 0, xy_run = 0;      // maxima over the batches
    auto close_batch_scratch = [&]() {
        for (int l = 0; l < nlev; ++l) { wlev_max = std::max(wlev_max, wfill[l]); xy_max = std::max(xy_max, xyfill[l]); wfill[l] = 0; xyfill[l] = 0; }
        xy_max = std::max(xy_max, xy_run);
        xy_run = 0;
    };

    int64_t rows_off = 0, child_off = 0;
    for (auto* lf : ctx->leaves) {
        const LeafSymbolic& S = lf->S;
        if (batch_start[(size_t)lf->index]) close_batch_scratch();
        lf->level_begin.assign(ngroups, 0); lf->level_count.assign(ngroups, 0);
        for (int l = 0; l < ngroups; ++l) lf->level_begin[l] = lfill[l];
        auto b2sc = [&](int b) { return b < lf->n0 ? b : b + lf->my0; };
        int64_t upd = lf->upd_base;
        for (int s = 0; s < S.ns; ++s) {
            FrontDesc& fd = fronts[(size_t)lf->front_base + s];
            fd.w = S.sn_col[s + 1] - S.sn_col[s]; fd.na = S.sn_na[s]; fd.nb = S.sn_nb[s];
            fd.Loff = lf->Lbase + S.Loff[s]; fd.Coff = lf->Cbase + S.Coff[s];
            fd.rows_off = rows_off + S.sn_rowptr[s];
            fd.col0 = lf->pos_off + S.sn_col[s];
            fd.child_begin = (int)(child_off + S.child_ptr[s]); fd.child_end = (int)(child_off + S.child_ptr[s + 1]);
            fd.leaf = lf->index;
            fd.parent = S.sn_parent[s] < 0 ? -1 : lf->front_base + S.sn_parent[s];
            fd.upd_off = upd; upd += fd.na;
            const int ol = nlev - 1 - S.sn_level[s];
            const int grp = group_of(S, s);
            level_fronts[(size_t)lfill[grp]++] = lf->front_base + s;
            ++lf->level_count[grp];
            fd.Woff = wfill[ol];
            int64_t wsz = (int64_t)(fd.na + fd.nb) * padded_width(fd.w);
            wfill[ol] += (wsz + 1) & ~int64_t(1);
            if (ctx->xform) {      // persistent within a batch: pack_solve reads X after the batch's last level
                fd.XYoff = xy_run;
                xy_run += (int64_t)padded_width(fd.w) * padded_width(fd.w);
            } else {
                fd.XYoff = xyfill[ol];
                xyfill[ol] += (int64_t)padded_width(fd.w) * padded_width(fd.w);
            }
        }
        for (size_t t = 0; t < S.rowidx.size(); ++t) {
            const int r = S.rowidx[t];
            rowg[(size_t)rows_off + t] = (r < lf->n) ? lf->pos_off + r : b2sc(r - lf->n);
            relidx[(size_t)rows_off + t] = S.relidx[t];
        }
        for (size_t t = 0; t < S.child_idx.size(); ++t) child_list[(size_t)child_off + t] = lf->front_base + S.child_idx[t];
        for (int p = 0; p < lf->n; ++p) {
            gidx[(size_t)lf->pos_off + p] = lf->pos_off + S.perm[p];
            ptype[(size_t)lf->pos_off + p] = S.ptype[p];
            sgn[(size_t)lf->pos_off + p] = S.sgn[p];
            diagpos[(size_t)lf->pos_off + p] = lf->kval_off + lf->diag_entry[p];
        }
        for (size_t e = 0; e < lf->val.size(); ++e) {
            kmap[(size_t)lf->kval_off + e] = lf->Lbase + S.kmap[e];
            kval[(size_t)lf->kval_off + e] = lf->val[e];
        }
        for (int i = 0; i < lf->n; ++i)
            for (int p = lf->krow[i]; p < lf->krow[i + 1]; ++p) {
                kcoo_row[(size_t)lf->kval_off + p] = lf->pos_off + i;
                kcoo_col[(size_t)lf->kval_off + p] = lf->pos_off + lf->jcol[p];
            }
        const int b0tail = lf->n0 + lf->my0;   // b0 layout [x0;y0;z0;yl;zl]: filled in by the kkt (needs mz0)
        (void)b0tail;
        for (size_t e = 0; e < lf->bval.size(); ++e) {
            bmap[(size_t)lf->bval_off + e] = lf->Lbase + S.bmap[e];
            bval[(size_t)lf->bval_off + e] = lf->bval[e];
            bg_row[(size_t)lf->bval_off + e] = lf->pos_off + lf->brow[e];
            bg_sc[(size_t)lf->bval_off + e] = b2sc(lf->bcol[e]);
            bg_b0[(size_t)lf->bval_off + e] = lf->bcol[e];   // border index; the kkt remaps to the b0 layout
        }
        leafdev[(size_t)lf->index].eps = 1e-8 * lf->knorm;
        leafdev[(size_t)lf->index].pos_off = lf->pos_off;
        leafdev[(size_t)lf->index].n = lf->n;
        // statistics kept on the host
        int64_t stored = lf->Lsize;
        lf->stats8[0] = S.nnzL; lf->stats8[1] = S.nnzL_border; lf->stats8[2] = stored; lf->stats8[3] = S.ns;
        lf->stats8[4] = S.nlevels; lf->stats8[5] = S.n_pairs; lf->stats8[6] = S.max_m; lf->stats8[7] = lf->Csize;
        lf->flops = S.flops_factor;
        lf->flops4[0] = S.flops_factor; lf->flops4[1] = lf->flops4[2] = lf->flops4[3] = 0.0;
        for (int s = 0; s < S.ns; ++s) {
            const double w = S.sn_col[s + 1] - S.sn_col[s], na = S.sn_na[s], nb = S.sn_nb[s], u = na + nb;
            lf->flops4[1] += w * nb * (nb + 1.0);      // 2 * w * nb (nb + 1) / 2
            lf->flops4[2] += w * u * (u + 1.0);        // every update entry (ancestor columns and border x border), lower
            lf->flops4[3] += 2.0 * w * w * u;          // W21 = A21 X^T
        }
        rows_off += (int64_t)S.rowidx.size();
        child_off += (int64_t)S.child_idx.size();
        // the symbolic arrays are now on their way to the device; drop the big host copies
        LeafSymbolic& Sm = lf->S;
        std::vector<int>().swap(Sm.rowidx); std::vector<int>().swap(Sm.relidx);
        std::vector<int64_t>().swap(Sm.kmap); std::vector<int64_t>().swap(Sm.bmap);
    }
    // inverse relative indices: for every front with a parent, parent local row -> row of this front's C (or -1)
    std::vector<int> invmap;
    {
        int64_t tot = 0;
        for (int64_t f = 0; f < nfr; ++f) {
            FrontDesc& fd = fronts[(size_t)f];
            fd.inv_off = tot;
            if (fd.parent >= 0) { const FrontDesc& pd = fronts[(size_t)fd.parent]; tot += pd.w + pd.na + pd.nb; }
        }
        invmap.assign((size_t)tot, -1);
        for (int64_t f = 0; f < nfr; ++f) {
            const FrontDesc& fd = fronts[(size_t)f];
            if (fd.parent < 0) continue;
            for (int t = 0; t < fd.na + fd.nb; ++t) invmap[(size_t)fd.inv_off + relidx[(size_t)fd.rows_off + t]] = t;
        }
    }
    close_batch_scratch();
    ctx->Wtot = wlev_max;
    ctx->XYtot = xy_max;
    // ---- factorisation launch lists: fronts grouped by (level, small|big), by leaf inside a group ----
    std::vector<int> fl_fronts((size_t)nfr);
    {
        auto kind_of = [&](const FrontDesc& fd) { return (fd.na + fd.nb >= ctx->big_u) ? 1 : 0; };
        std::vector<int> cnt2(nlev * 2 + 1, 0);
        for (auto* lf : ctx->leaves)
            for (int s2 = 0; s2 < lf->S.ns; ++s2)
                ++cnt2[(nlev - 1 - lf->S.sn_level[s2]) * 2 + kind_of(fronts[(size_t)lf->front_base + s2]) + 1];
        ctx->fl_ptr.assign(nlev * 2 + 1, 0);
        for (int g = 0; g < nlev * 2; ++g) ctx->fl_ptr[g + 1] = ctx->fl_ptr[g] + cnt2[g + 1];
        std::vector<int> fill(ctx->fl_ptr.begin(), ctx->fl_ptr.end() - 1);
        ctx->fl_umax.assign((size_t)nlev, 0);
        for (auto* lf : ctx->leaves)
            for (int s2 = 0; s2 < lf->S.ns; ++s2) {
                const FrontDesc& fd = fronts[(size_t)lf->front_base + s2];
                int& mx = ctx->fl_umax[(size_t)(nlev - 1 - lf->S.sn_level[s2])];
                if (kind_of(fd)) mx = std::max(mx, fd.na + fd.nb);
            }
        for (auto* lf : ctx->leaves) {
            lf->fl_begin.assign(nlev * 2, 0); lf->fl_count.assign(nlev * 2, 0);
            for (int g = 0; g < nlev * 2; ++g) lf->fl_begin[g] = fill[g];
            for (int s2 = 0; s2 < lf->S.ns; ++s2) {
                const int g = (nlev - 1 - lf->S.sn_level[s2]) * 2 + kind_of(fronts[(size_t)lf->front_base + s2]);
                fl_fronts[(size_t)fill[g]++] = lf->front_base + s2;
                ++lf->fl_count[g];
            }
        }
    }
    // ---- border groups: which big fronts leave their border x border update to a group SYRK (kernels.cuh) ----
    // (grouping per leaf: build_border_groups in symbolic.cpp, run with the symbolic analysis on the host threads)
    std::vector<GroupDesc> groups;
    std::vector<GroupMember> gmembers;
    std::vector<int> ginv;
    std::vector<int8_t> gflags;
    std::vector<GTile> gtiles;
    std::vector<std::vector<GTile>> lev_gtiles((size_t)nlev);      // group tiles by level of the group root, by leaf inside a level
    for (auto* lf : ctx->leaves) { lf->gtile_begin.assign((size_t)nlev, 0); lf->gtile_count.assign((size_t)nlev, 0); }
    std::vector<char> grouped((size_t)nfr, 0);
    if (ctx->grouping) {
        // persistent (per batch) W21 scratch of the grouped fronts, behind the per-level area; two copies, used by even / odd
        // batches: the group kernel of a batch still reads its copy while the next batch factorises (run_factor_levels)
        int64_t gw_run = 0, gw_max = 0;
        for (auto* lf : ctx->leaves) {
            if (batch_start[(size_t)lf->index]) { gw_max = std::max(gw_max, gw_run); gw_run = 0; }
            for (int s2 = 0; s2 < lf->S.ns; ++s2) {
                if (!lf->G.grouped[(size_t)s2]) continue;
                const FrontDesc& fd = fronts[(size_t)lf->front_base + s2];
                gw_run += ((int64_t)(fd.na + fd.nb) * padded_width(fd.w) + 1) & ~int64_t(1);
            }
        }
        gw_max = std::max(gw_max, gw_run);
        gw_run = 0;
        for (auto* lf : ctx->leaves) {
            if (batch_start[(size_t)lf->index]) gw_run = 0;
            const int64_t gw_base = wlev_max + (int64_t)(lf->batch & 1) * gw_max;
            const BorderGroups& G = lf->G;
            const int ng = (int)G.root.size();
            for (int s2 = 0; s2 < lf->S.ns; ++s2) grouped[(size_t)lf->front_base + s2] = G.grouped[(size_t)s2];
            // heavy groups first: their tiles loop over the most members
            std::vector<int> order((size_t)ng);
            std::iota(order.begin(), order.end(), 0);
            std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
                return G.mem_ptr[(size_t)x + 1] - G.mem_ptr[(size_t)x] > G.mem_ptr[(size_t)y + 1] - G.mem_ptr[(size_t)y];
            });
            for (int oi : order) {
                GroupDesc gd;
                gd.N = G.N[(size_t)oi]; gd.nblk = (gd.N + GROUP_TILE - 1) / GROUP_TILE;
                gd.root = lf->front_base + G.root[(size_t)oi]; gd.pad = 0;
                gd.mem_begin = (int)gmembers.size(); gd.inv_base = (int64_t)ginv.size();
                for (int q = G.mem_ptr[(size_t)oi]; q < G.mem_ptr[(size_t)oi + 1]; ++q) {
                    const int f = lf->front_base + G.mem_front[(size_t)q];
                    const FrontDesc& fd = fronts[(size_t)f];
                    GroupMember mb;
                    // the member's W21 = L21 D scratch moves from the per-level area to a persistent one: the group SYRK
                    // runs after the last level and reads (L_B D) from there instead of rebuilding it from L and D
                    FrontDesc& fdm = fronts[(size_t)f];
                    fdm.Woff = gw_base + gw_run;
                    const int64_t wsz = (int64_t)(fd.na + fd.nb) * padded_width(fd.w);
                    gw_run += (wsz + 1) & ~int64_t(1);
                    mb.Lb_off = fd.Loff + fd.w + fd.na; mb.Wb_off = fdm.Woff + fd.na;
                    mb.m = fd.w + fd.na + fd.nb; mb.u = fd.na + fd.nb; mb.w = fd.w; mb.col0 = fd.col0; mb.front = f; mb.pad = 0;
                    gmembers.push_back(mb);
                    const size_t base = ginv.size();
                    ginv.resize(base + (size_t)gd.nblk * GROUP_TILE, -1);
                    const int* pf = G.pos.data() + G.pos_ptr[(size_t)q];
                    if (G.pos_ptr[(size_t)q + 1] - G.pos_ptr[(size_t)q] != fd.nb) throw Error(PIPSGPU_ERR_INVALID, "border group: position map size mismatch");
                    for (int t = 0; t < fd.nb; ++t) ginv[base + (size_t)pf[t]] = t;
                    gflags.resize(base / GROUP_TILE + (size_t)gd.nblk, 0);
                    for (int t = 0; t < fd.nb; ++t) gflags[base / GROUP_TILE + (size_t)pf[t] / GROUP_TILE] = 1;
                }
                gd.mem_end = (int)gmembers.size();
                const int gi = (int)groups.size();
                groups.push_back(gd);
                const int ol = nlev - 1 - lf->S.sn_level[(size_t)G.root[(size_t)oi]];      // level order of the group root
                const FrontDesc& rootd = fronts[(size_t)gd.root];
                for (int bi = 0; bi < gd.nblk; ++bi)
                    for (int bj = 0; bj <= bi; ++bj)
                        lev_gtiles[(size_t)ol].push_back(GTile{gd.inv_base, rootd.rows_off + rootd.na, gd.mem_begin,
                                                               gd.mem_end - gd.mem_begin, gd.N, gd.nblk, bi, bj});
                lf->gtile_count[(size_t)ol] += (int64_t)gd.nblk * (gd.nblk + 1) / 2;
            }
            lf->G = BorderGroups();
        }
        ctx->Wtot = wlev_max + 2 * gw_max;
        int64_t run = 0;
        for (int l = 0; l < nlev; ++l) {
            for (auto* lf : ctx->leaves) { lf->gtile_begin[(size_t)l] = run; run += lf->gtile_count[(size_t)l]; }
            gtiles.insert(gtiles.end(), lev_gtiles[(size_t)l].begin(), lev_gtiles[(size_t)l].end());
            if ((int64_t)gtiles.size() != run) throw Error(PIPSGPU_ERR_INVALID, "group tile count mismatch");
            std::vector<GTile>().swap(lev_gtiles[(size_t)l]);
        }
    }
    ctx->ngtiles = (int64_t)gtiles.size();
    for (auto& e : ctx->ev_gt) cudaEventDestroy(e);
    ctx->ev_gt.resize((size_t)2 * (ctx->batch_ptr.size() - 1) * nlev);
    for (auto& e : ctx->ev_gt) CUDA_CHECK(cudaEventCreate(&e));
    ctx->gt_rec.assign(ctx->ev_gt.size() / 2, 0);
    // ---- update tiles of the big fronts, grouped by (level, leaf); grouped fronts only keep the ancestor columns ----
    std::vector<TileDesc> tiles;
    {
        for (auto* lf : ctx->leaves) { lf->tile_begin.assign(nlev, 0); lf->tile_count.assign(nlev, 0); }
        auto ntiles_of = [&](int64_t f) -> int64_t {
            const FrontDesc& fd = fronts[(size_t)f];
            const int u = fd.na + fd.nb;
            if (u < ctx->big_u) return 0;
            const int64_t nt = (u + TILE - 1) / TILE;
            if (!grouped[(size_t)f]) return nt * (nt + 1) / 2;
            const int64_t nj = (fd.na + TILE - 1) / TILE;          // tile columns that hold ancestor columns
            return nj * nt - nj * (nj - 1) / 2;
        };
        for (auto* lf : ctx->leaves)
            for (int s2 = 0; s2 < lf->S.ns; ++s2)
                lf->tile_count[nlev - 1 - lf->S.sn_level[s2]] += ntiles_of(lf->front_base + s2);
        ctx->tile_ptr.assign(nlev + 1, 0);
        for (int l = 0; l < nlev; ++l) {
            int64_t run = ctx->tile_ptr[l];
            for (auto* lf : ctx->leaves) { lf->tile_begin[l] = run; run += lf->tile_count[l]; }
            ctx->tile_ptr[l + 1] = run;
        }
        tiles.resize((size_t)ctx->tile_ptr[nlev]);
        for (auto* lf : ctx->leaves) {
            std::vector<int64_t> fill(lf->tile_begin);
            for (int s2 = 0; s2 < lf->S.ns; ++s2) {
                const int f = lf->front_base + s2;
                const FrontDesc& fd = fronts[(size_t)f];
                const int u = fd.na + fd.nb;
                if (u < ctx->big_u) continue;
                const int ol = nlev - 1 - lf->S.sn_level[s2];
                const int nt = (u + TILE - 1) / TILE;
                const int g = grouped[(size_t)f];
                for (int bi = 0; bi < nt; ++bi)
                    for (int bj = 0; bj <= bi; ++bj) {
                        if (g && bj * TILE >= fd.na) continue;
                        tiles[(size_t)fill[ol]++] = TileDesc{f, bi, bj, g};
                    }
            }
            for (int l = 0; l < nlev; ++l)
                if (fill[l] != lf->tile_begin[l] + lf->tile_count[l]) throw Error(PIPSGPU_ERR_INVALID, "update tile count mismatch");
        }
    }

    cudaStream_t s = ctx->stream;
    ctx->fronts.upload(fronts, s); ctx->rowg.upload(rowg, s); ctx->relidx.upload(relidx, s); ctx->invmap.upload(invmap, s);
    ctx->child_list.upload(child_list, s); ctx->level_fronts.upload(level_fronts, s); ctx->gidx.upload(gidx, s);
    ctx->ptype.upload(ptype, s); ctx->sgn.upload(sgn, s); ctx->kmap.upload(kmap, s); ctx->bmap.upload(bmap, s);
    ctx->diagpos.upload(diagpos, s); ctx->Kval.upload(kval, s); ctx->Bval.upload(bval, s);
    ctx->kcoo_row.upload(kcoo_row, s); ctx->kcoo_col.upload(kcoo_col, s);
    ctx->bg_row.upload(bg_row, s); ctx->bg_sc.upload(bg_sc, s); ctx->bg_b0.upload(bg_b0, s);
    ctx->leafdev.upload(leafdev, s);
    ctx->tiles.upload(tiles, s);
    ctx->groups.upload(groups, s); ctx->gmembers.upload(gmembers, s); ctx->ginv.upload(ginv, s); ctx->gflags.upload(gflags, s); ctx->gtiles.upload(gtiles, s);
    ctx->fl_fronts.upload(fl_fronts, s);
    ctx->L.alloc((size_t)Ltot + 2); ctx->C.alloc((size_t)Ctot + 2); ctx->W.alloc((size_t)ctx->Wtot + 2); ctx->XY.alloc((size_t)ctx->XYtot + 2);
    ctx->Dd.alloc((size_t)npos + 1); ctx->De.alloc((size_t)npos + 1); ctx->Did.alloc((size_t)npos + 1);
    ctx->Die.alloc((size_t)npos + 1);
    ctx->x.alloc((size_t)npos + 1); ctx->upd.alloc((size_t)updtot + 1); ctx->dvals.alloc((size_t)npos + 1);
    ctx->work.alloc((size_t)npos + 1);
    ctx->stats.alloc((size_t)4 * std::max(nl, 1));
    {
        // ---- work units of the streaming solve (psolve.cu): bottom subtrees with <= PS_SMAX columns + single fronts above ----
        // fronts of a leaf are numbered in postorder, so a subtree is a contiguous range of fronts and of positions
        if (ps_rows_per_chunk(MAXW | 1) < MAXW) throw Error(PIPSGPU_ERR_INVALID, "PS_BUF too small: L11 must fit the first chunk");
        std::vector<int> sub_nf((size_t)nfr, 1), sub_cols((size_t)nfr), sub_minf((size_t)nfr), sub_minc((size_t)nfr);
        std::vector<int64_t> sub_wt((size_t)nfr);
        for (int64_t f = 0; f < nfr; ++f) {
            const FrontDesc& fd = fronts[(size_t)f];
            sub_cols[(size_t)f] = fd.w; sub_minf[(size_t)f] = (int)f; sub_minc[(size_t)f] = fd.col0;
            sub_wt[(size_t)f] = (int64_t)(fd.w + fd.na) * (fd.w | 1);
        }
        for (int64_t f = 0; f < nfr; ++f) {      // children precede their parent
            const int p = fronts[(size_t)f].parent;
            if (p < 0) continue;
            if (p <= f) throw Error(PIPSGPU_ERR_INVALID, "front tree is not in postorder");
            sub_nf[(size_t)p] += sub_nf[(size_t)f];
            sub_cols[(size_t)p] = std::min<int64_t>((int64_t)sub_cols[(size_t)p] + sub_cols[(size_t)f], 1 << 30);
            sub_minf[(size_t)p] = std::min(sub_minf[(size_t)p], sub_minf[(size_t)f]);
            sub_minc[(size_t)p] = std::min(sub_minc[(size_t)p], sub_minc[(size_t)f]);
            sub_wt[(size_t)p] += sub_wt[(size_t)f];
        }
        auto eligible = [&](int64_t f) {
            const FrontDesc& fd = fronts[(size_t)f];
            return sub_cols[(size_t)f] <= PS_SMAX && sub_minf[(size_t)f] == (int)f - sub_nf[(size_t)f] + 1 &&
                   fd.col0 + fd.w - sub_minc[(size_t)f] == sub_cols[(size_t)f];
        };
        struct U { int root, first_front, nfronts; int64_t wt; };
        std::vector<U> subs, tops;
        std::vector<int> unit_of_front((size_t)nfr, -1);
        std::vector<char> is_top((size_t)nfr, 0);
        for (int64_t f = nfr - 1; f >= 0; --f) {
            if (unit_of_front[(size_t)f] != -1) continue;
            if (eligible(f)) {
                const int f0 = (int)f - sub_nf[(size_t)f] + 1;
                for (int g = f0; g <= (int)f; ++g) unit_of_front[(size_t)g] = -2;      // member of a subtree unit (id below)
                subs.push_back(U{(int)f, f0, sub_nf[(size_t)f], sub_wt[(size_t)f]});
            } else {
                unit_of_front[(size_t)f] = -3;
                is_top[(size_t)f] = 1;
            }
        }
        std::stable_sort(subs.begin(), subs.end(), [](const U& a, const U& b) { return a.wt > b.wt; });
        for (int64_t t = 0; t < nfr; ++t) {      // level_fronts: deepest level first
            const int f = level_fronts[(size_t)t];
            if (is_top[(size_t)f]) tops.push_back(U{f, f, 1, 0});
        }
        const int nunits = (int)(subs.size() + tops.size());
        std::vector<SolveUnit> units((size_t)nunits);
        std::vector<SolveFront> sf((size_t)nfr);
        std::vector<int> chunk_front, chunk_idx;
        std::vector<uint16_t> slots((size_t)nrows, (uint16_t)0xFFFF);
        int64_t off = 0;
        int sfi = 0;
        for (int u = 0; u < nunits; ++u) {
            const U& uu = (u < (int)subs.size()) ? subs[(size_t)u] : tops[(size_t)u - subs.size()];
            for (int g = uu.first_front; g < uu.first_front + uu.nfronts; ++g) unit_of_front[(size_t)g] = u;
        }
        int maxchild = 1;
        for (int u = 0; u < nunits; ++u) {
            const U& uu = (u < (int)subs.size()) ? subs[(size_t)u] : tops[(size_t)u - subs.size()];
            const FrontDesc& rd = fronts[(size_t)uu.root];
            SolveUnit& un = units[(size_t)u];
            un.first = sfi; un.nfronts = uu.nfronts;
            un.pbeg = sub_minc[(size_t)uu.root]; un.ncols = rd.col0 + rd.w - un.pbeg;
            if (uu.nfronts == 1) { un.pbeg = rd.col0; un.ncols = rd.w; }
            if (rd.rows_off > 2000000000LL) throw Error(PIPSGPU_ERR_INVALID, "row index pool too large for one context");
            un.out_rows_off = (int)rd.rows_off; un.nout = std::min(rd.na, PS_NOUT);
            un.parent = rd.parent >= 0 ? unit_of_front[(size_t)rd.parent] : -1;
            un.nchild = 0;
            const int* orow = rowg.data() + rd.rows_off;      // ancestor rows of the unit's root front: ascending positions
            for (int g = uu.first_front; g < uu.first_front + uu.nfronts; ++g) {
                const FrontDesc& fd = fronts[(size_t)g];
                SolveFront& f = sf[(size_t)sfi];
                f.blob = off; f.col0 = fd.col0; f.rows_off = (int)fd.rows_off; f.w = fd.w; f.na = fd.na; f.ldr = fd.w | 1;
                f.front = g;
                const int rows = fd.w + fd.na, R = ps_rows_per_chunk(f.ldr);
                const int nch = (rows + R - 1) / R;
                f.R = R; f.nch = nch;
                for (int c = 0; c < nch; ++c) { chunk_front.push_back(sfi); chunk_idx.push_back(c); }
                off += ps_front_doubles(f.ldr, rows);
                for (int t = 0; t < fd.na; ++t) {
                    const int p = rowg[(size_t)fd.rows_off + t];
                    uint16_t sl = 0xFFFF;
                    if (p >= un.pbeg && p < un.pbeg + un.ncols) sl = (uint16_t)(p - un.pbeg);
                    else {
                        const int* it = std::lower_bound(orow, orow + rd.na, p);
                        if (it != orow + rd.na && *it == p && (it - orow) < un.nout) sl = (uint16_t)(PS_SMAX + (it - orow));
                    }
                    slots[(size_t)fd.rows_off + t] = sl;
                }
                ++sfi;
            }
        }
        for (int u = 0; u < nunits; ++u)
            if (units[(size_t)u].parent >= 0) maxchild = std::max(maxchild, ++units[(size_t)units[(size_t)u].parent].nchild);
        ctx->SLtot = off;
        ctx->nunits = nunits;
        ctx->nsub_units = (int)subs.size();
        ctx->ps_epoch_max = (1 << 30) / maxchild - 2;   // the child counters are cumulative: nchild * epoch must not wrap
        ctx->nchunks = (int)chunk_front.size();
        ctx->sfront.upload(sf, s); ctx->sunits.upload(units, s); ctx->slots.upload(slots, s);
        {   // the same chunk list grouped by leaf
            std::vector<int> ord(chunk_front.size());
            std::iota(ord.begin(), ord.end(), 0);
            auto leaf_of = [&](int q) { return fronts[(size_t)sf[(size_t)chunk_front[(size_t)q]].front].leaf; };
            std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return leaf_of(x) < leaf_of(y); });
            std::vector<int> cf(ord.size()), ci(ord.size());
            for (auto* lf : ctx->leaves) lf->chunk_count = 0;
            for (size_t t = 0; t < ord.size(); ++t) {
                cf[t] = chunk_front[(size_t)ord[t]]; ci[t] = chunk_idx[(size_t)ord[t]];
                ++ctx->leaves[(size_t)leaf_of(ord[t])]->chunk_count;
            }
            int run = 0;
            for (auto* lf : ctx->leaves) { lf->chunk_begin = run; run += lf->chunk_count; }
            ctx->chunk_front_l.upload(cf, s); ctx->chunk_idx_l.upload(ci, s);
        }
        ctx->SL.alloc((size_t)off + 2);
        ctx->acc.alloc((size_t)npos + 1);
        ctx->ucnt.alloc((size_t)nunits + 1); ctx->uflag.alloc((size_t)nunits + 1); ctx->ticket.alloc(2);
        CUDA_CHECK(cudaMemsetAsync(ctx->acc.p, 0, sizeof(double) * ((size_t)npos + 1), s));
        CUDA_CHECK(cudaMemsetAsync(ctx->ucnt.p, 0, sizeof(int) * ((size_t)nunits + 1), s));
        CUDA_CHECK(cudaMemsetAsync(ctx->uflag.p, 0, sizeof(int) * ((size_t)nunits + 1), s));
        CUDA_CHECK(cudaMemsetAsync(ctx->ticket.p, 0, sizeof(int) * 2, s));
        CUDA_CHECK(cudaStreamSynchronize(s));     // the host vectors above go out of scope
    }
    CUDA_CHECK(cudaMemsetAsync(ctx->Die.p, 0, sizeof(double) * ((size_t)npos + 1), s));
    CUDA_CHECK(cudaMemsetAsync(ctx->De.p, 0, sizeof(double) * ((size_t)npos + 1), s));
    CUDA_CHECK(cudaMallocHost(&ctx->h_dvals, sizeof(double) * ((size_t)npos + 1)));
    for (auto* lf : ctx->leaves)
        for (int i = 0; i < lf->n; ++i) ctx->h_dvals[(size_t)lf->pos_off + i] = lf->val[lf->diag_entry[i]];
    CUDA_CHECK(cudaStreamSynchronize(s));
    ctx->committed = true;
    PIPSGPU_CATCH
}

static void leaf_fill_dvals(pipsgpu_leaf* lf, const double* dd, const double* nomegainv, double regP, double regDy, double regDz) {
    double* d = lf->ctx->h_dvals + lf->pos_off;
    regP += lf->reg[0]; regDy += lf->reg[1]; regDz += lf->reg[2];
    for (int i = 0; i < lf->nx; ++i) d[i] = (dd ? dd[i] : 0.0) + regP;
    for (int i = 0; i < lf->ny; ++i) d[lf->nx + i] = -regDy;
    for (int i = 0; i < lf->nz; ++i) d[lf->nx + lf->ny + i] = (nomegainv ? nomegainv[i] : 0.0) - regDz;
}

static void leaf_push_dvals(pipsgpu_ctx* ctx, int64_t pos_begin, int64_t count) {
    if (count <= 0) return;
    CUDA_CHECK(cudaMemcpyAsync(ctx->dvals.p + pos_begin, ctx->h_dvals + pos_begin, sizeof(double) * (size_t)count,
                               cudaMemcpyHostToDevice, ctx->stream));
    launch_set_diag(ctx->diagpos.p + pos_begin, ctx->dvals.p + pos_begin, count, ctx->Kval.p, ctx->stream);
}

int pipsgpu_leaf_set_diagonal(pipsgpu_leaf* lf, const double* dd, const double* nomegainv, double regP, double regDy, double regDz) {
    PIPSGPU_TRY
    if (!lf) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_leaf_set_diagonal: null leaf");
    pipsgpu_ctx* ctx = lf->ctx;
    if (!ctx->committed) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_leaf_set_diagonal: context not committed");
    require_device(ctx);
    leaf_fill_dvals(lf, dd, nomegainv, regP, regDy, regDz);
    leaf_push_dvals(ctx, lf->pos_off, lf->n);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    lf->factorized = false;
    PIPSGPU_CATCH
}

int pipsgpu_leaf_set_regularization(pipsgpu_leaf* lf, double regP, double regDy, double regDz) {
    PIPSGPU_TRY
    if (!lf) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_leaf_set_regularization: null leaf");
    lf->reg[0] = regP; lf->reg[1] = regDy; lf->reg[2] = regDz;
    PIPSGPU_CATCH
}

int pipsgpu_leaf_set_values(pipsgpu_leaf* lf, const double* val) {
    PIPSGPU_TRY
    if (!lf || !val) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_leaf_set_values: null argument");
    pipsgpu_ctx* ctx = lf->ctx;
    if (!ctx->committed) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_leaf_set_values: context not committed");
    require_device(ctx);
    std::copy(val, val + lf->val.size(), lf->val.begin());
    for (int i = 0; i < lf->n; ++i) ctx->h_dvals[(size_t)lf->pos_off + i] = lf->val[lf->diag_entry[i]];
    CUDA_CHECK(cudaMemcpyAsync(ctx->Kval.p + lf->kval_off, lf->val.data(), sizeof(double) * lf->val.size(), cudaMemcpyHostToDevice,
                               ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    lf->factorized = false;
    PIPSGPU_CATCH
}

static void factor_prepare(pipsgpu_ctx* ctx, pipsgpu_leaf* only) {
    cudaStream_t s = ctx->stream;
    if (only) {
        launch_vec_zero(ctx->L.p + only->Lbase, only->Lsize, s);
        launch_scatter_values(ctx->kmap.p + only->kval_off, ctx->Kval.p + only->kval_off, (int64_t)only->val.size(), ctx->L.p, s);
        launch_scatter_values(ctx->bmap.p + only->bval_off, ctx->Bval.p + only->bval_off, (int64_t)only->bval.size(), ctx->L.p, s);
        CUDA_CHECK(cudaMemsetAsync(ctx->stats.p + 4 * only->index, 0, sizeof(int) * 4, s));
    } else {
        launch_vec_zero(ctx->L.p, ctx->Ltot, s);
        launch_scatter_values(ctx->kmap.p, ctx->Kval.p, ctx->nnzK, ctx->L.p, s);
        launch_scatter_values(ctx->bmap.p, ctx->Bval.p, ctx->nnzB, ctx->L.p, s);
        CUDA_CHECK(cudaMemsetAsync(ctx->stats.p, 0, sizeof(int) * 4 * ctx->leaves.size(), s));
    }
}

int pipsgpu_leaf_factorize(pipsgpu_leaf* lf, int inertia[3]) {
    PIPSGPU_TRY
    if (!lf) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_leaf_factorize: null leaf");
    pipsgpu_ctx* ctx = lf->ctx;
    if (!ctx->committed) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_leaf_factorize: context not committed");
    require_device(ctx);
    factor_prepare(ctx, lf);
    run_factor_levels(ctx, nullptr, lf);
    fetch_leaf_stats(ctx);
    lf->factorized = true;
    if (inertia) { inertia[0] = lf->inertia[0]; inertia[1] = lf->inertia[1]; inertia[2] = lf->inertia[2]; }
    PIPSGPU_CATCH
}

int pipsgpu_ctx_factorize_leaves(pipsgpu_ctx* ctx, pipsgpu_sc* sc) {
    PIPSGPU_TRY
    require_device(ctx);
    if (!ctx->committed) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_ctx_factorize_leaves: context not committed");
    factor_prepare(ctx, nullptr);
    run_factor_levels(ctx, sc, nullptr);
    fetch_leaf_stats(ctx);
    for (auto* lf : ctx->leaves) lf->factorized = true;
    PIPSGPU_CATCH
}

int pipsgpu_leaf_add_schur(pipsgpu_leaf* lf, pipsgpu_sc* sc) {
    PIPSGPU_TRY
    if (!lf || !sc) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_leaf_add_schur: null argument");
    pipsgpu_ctx* ctx = lf->ctx;
    require_device(ctx);
    if (!ctx->committed) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_leaf_add_schur: context not committed");
    // the Schur term is a by-product of the elimination: redo this leaf's numeric phase with the SC attached
    factor_prepare(ctx, lf);
    run_factor_levels(ctx, sc, lf);
    fetch_leaf_stats(ctx);
    lf->factorized = true;
    PIPSGPU_CATCH
}

int pipsgpu_leaf_solve(pipsgpu_leaf* lf, int nrhs, double* rhs) {
    PIPSGPU_TRY
    if (!lf || !rhs) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_leaf_solve: null argument");
    pipsgpu_ctx* ctx = lf->ctx;
    require_device(ctx);
    if (!lf->factorized) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_leaf_solve: leaf not factorized");
    cudaStream_t s = ctx->stream;
    const int64_t n = lf->n;
    const int chunk = 64;
    if (lf->sv_chunk == 0) {   // work buffers and the leaf-local permutation live on the leaf (allocated once)
        lf->sv_b.alloc((size_t)n * chunk); lf->sv_x.alloc((size_t)n * chunk);
        lf->sv_u.alloc((size_t)std::max<int64_t>(lf->upd_size, 1) * chunk);
        std::vector<int> h((size_t)n);
        for (int64_t p = 0; p < n; ++p) h[(size_t)p] = lf->S.perm[(size_t)p];
        lf->sv_perm.upload(h, s);
        CUDA_CHECK(cudaStreamSynchronize(s));
        lf->sv_chunk = chunk;
    }
    DevBuf<double>&b = lf->sv_b, &xp = lf->sv_x, &up = lf->sv_u;
    DevBuf<int>& lg = lf->sv_perm;
    const bool refine = ctx->refine_steps > 0 && (lf->n_perturbed > 0 || ctx->refine_always);
    for (int r0 = 0; r0 < nrhs; r0 += chunk) {
        const int nr = std::min(chunk, nrhs - r0);
        CUDA_CHECK(cudaMemcpyAsync(b.p, rhs + (size_t)r0 * n, sizeof(double) * (size_t)n * nr, cudaMemcpyHostToDevice, s));
        launch_gather_perm(lg.p, b.p, xp.p, n, nr, n, s);
        // kernels index x by global position and upd by global offset: shift the bases so that this leaf's range
        // lands at the start of the private buffers
        SolveArgs sa = ctx->solve_args(xp.p - lf->pos_off, n, up.p - lf->upd_base, std::max<int64_t>(lf->upd_size, 1));
        run_leaf_solve(ctx, sa, nr, lf, lf->pos_off, n);
        if (refine) {
            // iterative refinement as in the reference plug-ins (PARDISO iparm[7] = 2, PardisoProjectSolver.C:72; MA27 own
            // loop, Ma27Solver.C:184-249): x += K^-1 (b - K x) with the resident matrix values, per right-hand side
            DevBuf<double>& xs = lf->sv_r;      // solution in original order, nr x n
            xs.ensure((size_t)n * chunk); lf->sv_r2.ensure((size_t)n * chunk);
            launch_scatter_perm(lg.p, xp.p, xs.p, n, nr, n, s);
            for (int it = 0; it < ctx->refine_steps; ++it) {
                double* res = lf->sv_r2.p;
                CUDA_CHECK(cudaMemcpyAsync(res, b.p, sizeof(double) * (size_t)n * nr, cudaMemcpyDeviceToDevice, s));
                for (int q = 0; q < nr; ++q)      // res_q = b_q - K x_q (positions shifted to the leaf's private buffers)
                    launch_sym_spmv((int64_t)lf->val.size(), ctx->kcoo_row.p + lf->kval_off, ctx->kcoo_col.p + lf->kval_off,
                                    ctx->Kval.p + lf->kval_off, -1.0, xs.p + (size_t)q * n - lf->pos_off,
                                    res + (size_t)q * n - lf->pos_off, s);
                launch_gather_perm(lg.p, res, xp.p, n, nr, n, s);
                run_leaf_solve(ctx, sa, nr, lf, lf->pos_off, n);
                launch_scatter_perm(lg.p, xp.p, res, n, nr, n, s);
                launch_axpy(1.0, res, xs.p, (int64_t)n * nr, s);
            }
            CUDA_CHECK(cudaMemcpyAsync(rhs + (size_t)r0 * n, xs.p, sizeof(double) * (size_t)n * nr, cudaMemcpyDeviceToHost, s));
            CUDA_CHECK(cudaStreamSynchronize(s));
            continue;
        }
        launch_scatter_perm(lg.p, xp.p, b.p, n, nr, n, s);
        CUDA_CHECK(cudaMemcpyAsync(rhs + (size_t)r0 * n, b.p, sizeof(double) * (size_t)n * nr, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
    }
    PIPSGPU_CATCH
}

int pipsgpu_leaf_get_inertia(pipsgpu_leaf* lf, int inertia[3], int* n_perturbed) {
    PIPSGPU_TRY
    if (!lf) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_leaf_get_inertia: null leaf");
    if (inertia) { inertia[0] = lf->inertia[0]; inertia[1] = lf->inertia[1]; inertia[2] = lf->inertia[2]; }
    if (n_perturbed) *n_perturbed = lf->n_perturbed;
    PIPSGPU_CATCH
}

int pipsgpu_leaf_get_stats(pipsgpu_leaf* lf, int64_t out[8], double* flops) {
    PIPSGPU_TRY
    if (!lf || !out) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_leaf_get_stats: null argument");
    if (!lf->analyzed) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_leaf_get_stats: leaf not analyzed");
    if (lf->ctx->committed) {
        for (int i = 0; i < 8; ++i) out[i] = lf->stats8[i];
        if (flops) *flops = lf->flops;
    } else {
        const LeafSymbolic& S = lf->S;
        out[0] = S.nnzL; out[1] = S.nnzL_border; out[2] = S.Loff.empty() ? 0 : S.Loff.back(); out[3] = S.ns;
        out[4] = S.nlevels; out[5] = S.n_pairs; out[6] = S.max_m; out[7] = S.Coff.empty() ? 0 : S.Coff.back();
        if (flops) *flops = S.flops_factor;
    }
    PIPSGPU_CATCH
}

int pipsgpu_leaf_get_flops(pipsgpu_leaf* lf, double out[4]) {
    PIPSGPU_TRY
    if (!lf || !out) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_leaf_get_flops: null argument");
    if (!lf->ctx->committed) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_leaf_get_flops: context not committed");
    for (int i = 0; i < 4; ++i) out[i] = lf->flops4[i];
    PIPSGPU_CATCH
}

int pipsgpu_ctx_num_batches(pipsgpu_ctx* ctx) { return ctx ? (int)ctx->batch_ptr.size() - 1 : -1; }

int pipsgpu_leaf_download_factor(pipsgpu_leaf* lf, double* L, double* dinv_d, double* dinv_e) {
    PIPSGPU_TRY
    if (!lf) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_leaf_download_factor: null leaf");
    pipsgpu_ctx* ctx = lf->ctx;
    require_device(ctx);
    if (!lf->factorized) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_leaf_download_factor: leaf not factorized");
    cudaStream_t s = ctx->stream;
    if (L && lf->Lsize > 0)
        CUDA_CHECK(cudaMemcpyAsync(L, ctx->L.p + lf->Lbase, sizeof(double) * (size_t)lf->Lsize, cudaMemcpyDeviceToHost, s));
    if (dinv_d) CUDA_CHECK(cudaMemcpyAsync(dinv_d, ctx->Did.p + lf->pos_off, sizeof(double) * (size_t)lf->n, cudaMemcpyDeviceToHost, s));
    if (dinv_e) CUDA_CHECK(cudaMemcpyAsync(dinv_e, ctx->Die.p + lf->pos_off, sizeof(double) * (size_t)lf->n, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    PIPSGPU_CATCH
}

int pipsgpu_leaf_destroy(pipsgpu_leaf*) {
    // leaves are owned by their context (device pools are shared); destroyed with pipsgpu_ctx_destroy
    return PIPSGPU_OK;
}

// ------------------------------------------------------------------------------------------------------------
// dense Schur complement
// ------------------------------------------------------------------------------------------------------------
int pipsgpu_sc_create(pipsgpu_ctx* ctx, int n0, int my0, int myl, int mzl, pipsgpu_sc** out) {
    PIPSGPU_TRY
    require_device(ctx);
    if (!out || n0 < 0 || my0 < 0 || myl < 0 || mzl < 0) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_sc_create: invalid argument");
    auto sc = std::make_unique<pipsgpu_sc>();
    sc->ctx = ctx; sc->n0 = n0; sc->my0 = my0; sc->myl = myl; sc->mzl = mzl;
    sc->n = n0 + my0 + myl + mzl;
    const size_t n = (size_t)sc->n;
    sc->ldf = (int64_t)((n + 15) & ~size_t(15));
    if (sc->ldf == 0) sc->ldf = 16;
    const size_t ldf = (size_t)sc->ldf;
    const size_t npanels = (n + DENSE_PANEL - 1) / DENSE_PANEL;
    sc->A.alloc(n * n + 1); sc->Fm.alloc(std::max<size_t>(n, 1) * ldf); sc->D.alloc(n + 1); sc->Dinv.alloc(n + DENSE_PANEL + 1);
    sc->De.alloc(n + 1); sc->Dinv_e.alloc(n + 1); sc->Wc.alloc(2 * n + 2); sc->perm.alloc(n + 1); sc->ptype.alloc(n + 1);
    sc->bk.alloc(1); sc->swork.alloc(n + 1); sc->scale.alloc(n + 1);
    sc->Wbuf.alloc(2 * std::max<size_t>(n, 1) * DENSE_PANEL); sc->stats.alloc(4); sc->rhs.alloc(n + 1);
    sc->Xall.alloc(std::max<size_t>(npanels, 1) * DENSE_PANEL * DENSE_PANEL); sc->ybuf.alloc(n + 64);
    sc->tflags.alloc(2 * (n / 64 + 2));
    CUDA_CHECK(cudaMemsetAsync(sc->tflags.p, 0, sizeof(int) * sc->tflags.n, ctx->stream));
    CUDA_CHECK(cudaMemsetAsync(sc->Dinv.p, 0, sizeof(double) * sc->Dinv.n, ctx->stream));
    CUDA_CHECK(cudaMemsetAsync(sc->Fm.p, 0, sizeof(double) * sc->Fm.n, ctx->stream));
    CUDA_CHECK(cudaMallocHost(&sc->h_stats, sizeof(int) * 4));
    if (n > 0) {
        make_tmap(&sc->tmF, sc->Fm.p, (int64_t)n, (int64_t)n, sc->ldf);
        make_tmap(&sc->tmW[0], sc->Wbuf.p, (int64_t)n, DENSE_PANEL, DENSE_PANEL);
        make_tmap(&sc->tmW[1], sc->Wbuf.p + n * DENSE_PANEL, (int64_t)n, DENSE_PANEL, DENSE_PANEL);
        make_tmap(&sc->tmX, sc->Xall.p, (int64_t)(npanels * DENSE_PANEL), DENSE_PANEL, DENSE_PANEL);
    }
    CUDA_CHECK(cudaMemsetAsync(sc->A.p, 0, sizeof(double) * (n * n + 1), ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    *out = sc.release();
    PIPSGPU_CATCH
}

pipsgpu_sc::~pipsgpu_sc() {
    if (h_stats) cudaFreeHost(h_stats);
    if (sp_leaf) pipsgpu_leaf_destroy(sp_leaf);
    if (sp_ctx) pipsgpu_ctx_destroy(sp_ctx);
}

int pipsgpu_sc_set_sparse_pattern(pipsgpu_sc* sc, const int* krow, const int* jcol, int threads) {
    PIPSGPU_TRY
    if (!sc || !krow || !jcol) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_sc_set_sparse_pattern: null argument");
    pipsgpu_ctx* ctx = sc->ctx;
    require_device(ctx);
    if (sc->sp_ctx) {
        // a new pattern replaces the old one (DoubleLinearSolver::matrixRebuild, SOL/DoubleLinearSolver.h:39: the sparsified
        // Schur complement of the reference's preconditioner changes its pattern): the symbolic analysis is redone
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        if (sc->sp_leaf) pipsgpu_leaf_destroy(sc->sp_leaf);
        pipsgpu_ctx_destroy(sc->sp_ctx);
        sc->sp_leaf = nullptr; sc->sp_ctx = nullptr; sc->sp_nnz = 0; sc->used_sparse = false;
    }
    const int n = sc->n;
    const int64_t nnz = krow[n];
    std::vector<int64_t> idx((size_t)nnz);
    std::vector<double> ones((size_t)nnz, 1.0);
    for (int i = 0; i < n; ++i)
        for (int p = krow[i]; p < krow[i + 1]; ++p) {
            if (jcol[p] < 0 || jcol[p] > i) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_sc_set_sparse_pattern: pattern must be lower triangular CSR");
            idx[(size_t)p] = (int64_t)i * n + jcol[p];
            // symbolic analysis only sees magnitudes (2x2 pivot matching prefers heavy edges): expected sign on the diagonal
            if (jcol[p] == i) ones[(size_t)p] = (i < sc->n0) ? 1.0 : ((i < sc->n0 + sc->my0) ? 0.0 : -1.0);
        }
    pipsgpu_ctx* sp = nullptr;
    int rc = pipsgpu_ctx_create(ctx->device, &sp);
    if (rc != 0) return rc;
    sc->sp_ctx = sp;
    rc = pipsgpu_ctx_set_stream(sp, (void*)ctx->stream);
    if (rc != 0) return rc;
    // SC index order [x0; y0; yl; zl] = leaf order [x; y; z]: x0 rows are positive pivots, y0 rows (zero diagonal) are
    // paired with x0 columns, the link rows carry the negative (semi)definite block
    rc = pipsgpu_leaf_create(sp, n, sc->n0, sc->my0, sc->myl + sc->mzl, krow, jcol, ones.data(), &sc->sp_leaf);
    if (rc != 0) return rc;
    rc = pipsgpu_ctx_commit(sp, threads);
    if (rc != 0) return rc;
    sc->sp_nnz = nnz;
    sc->sp_idx.upload(idx, ctx->stream);
    CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    sc->factorized = false;
    PIPSGPU_CATCH
}

int pipsgpu_sc_uses_sparse_solver(pipsgpu_sc* sc) { return (sc && sc->used_sparse) ? 1 : 0; }

int pipsgpu_sc_destroy(pipsgpu_sc* sc) {
    PIPSGPU_TRY
    if (sc) {
        cudaSetDevice(sc->ctx->device);
        cudaStreamSynchronize(sc->ctx->stream);
        delete sc;
    }
    PIPSGPU_CATCH
}

int pipsgpu_sc_dim(pipsgpu_sc* sc) { return sc ? sc->n : -1; }

int pipsgpu_sc_zero(pipsgpu_sc* sc) {
    PIPSGPU_TRY
    if (!sc) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_sc_zero: null");
    require_device(sc->ctx);
    CUDA_CHECK(cudaMemsetAsync(sc->A.p, 0, sizeof(double) * (size_t)sc->n * sc->n, sc->ctx->stream));
    sc->factorized = false;
    PIPSGPU_CATCH
}

int pipsgpu_sc_allreduce(pipsgpu_sc* sc) {
    PIPSGPU_TRY
    if (!sc) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_sc_allreduce: null");
    require_device(sc->ctx);
    sc_allreduce_impl(sc);
    PIPSGPU_CATCH
}

int pipsgpu_sc_add_sparse(pipsgpu_sc* sc, int row0, int col0, int m, const int* krow, const int* jcol, const double* val, double alpha) {
    PIPSGPU_TRY
    if (!sc) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_sc_add_sparse: null");
    require_device(sc->ctx);
    if (m <= 0 || !krow) return PIPSGPU_OK;
    CsrDev M;
    int maxc = 0;
    for (int p = 0; p < krow[m]; ++p) maxc = std::max(maxc, jcol[p] + 1);
    if (row0 < 0 || col0 < 0 || row0 + m > sc->n || col0 + maxc > sc->n) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_sc_add_sparse: block out of range");
    M.upload(m, maxc, krow, jcol, val, sc->ctx->stream);
    launch_sc_add_sparse(sc->A.p, sc->n, row0, col0, m, M.krow.p, M.jcol.p, M.val.p, alpha, sc->ctx->stream);
    CUDA_CHECK(cudaStreamSynchronize(sc->ctx->stream));
    PIPSGPU_CATCH
}

int pipsgpu_sc_add_diag(pipsgpu_sc* sc, int offset, int n, const double* d, double alpha) {
    PIPSGPU_TRY
    if (!sc) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_sc_add_diag: null");
    require_device(sc->ctx);
    if (n <= 0) return PIPSGPU_OK;
    if (offset < 0 || offset + n > sc->n) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_sc_add_diag: out of range");
    DevBuf<double> dv;
    if (d) {
        dv.alloc((size_t)n);
        CUDA_CHECK(cudaMemcpyAsync(dv.p, d, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, sc->ctx->stream));
    }
    launch_sc_add_diag(sc->A.p, sc->n, offset, n, dv.p, alpha, sc->ctx->stream);
    CUDA_CHECK(cudaStreamSynchronize(sc->ctx->stream));
    PIPSGPU_CATCH
}

int pipsgpu_sc_factorize(pipsgpu_sc* sc, int inertia[3]) {
    PIPSGPU_TRY
    if (!sc) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_sc_factorize: null");
    require_device(sc->ctx);
    sc_factorize_impl(sc);
    if (inertia) { inertia[0] = sc->inertia[0]; inertia[1] = sc->inertia[1]; inertia[2] = sc->inertia[2]; }
    PIPSGPU_CATCH
}

int pipsgpu_sc_solve(pipsgpu_sc* sc, int nrhs, double* rhs) {
    PIPSGPU_TRY
    if (!sc || !rhs) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_sc_solve: null");
    require_device(sc->ctx);
    if (!sc->factorized) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_sc_solve: not factorized");
    cudaStream_t s = sc->ctx->stream;
    const size_t n = (size_t)sc->n;
    DevBuf<double> x;
    x.alloc(n * (size_t)std::max(nrhs, 1));
    CUDA_CHECK(cudaMemcpyAsync(x.p, rhs, sizeof(double) * n * nrhs, cudaMemcpyHostToDevice, s));
    sc->swork.ensure(n * (size_t)std::max(nrhs, 1));
    sc_solve_device(sc, x.p, nrhs, (int64_t)n);
    CUDA_CHECK(cudaMemcpyAsync(rhs, x.p, sizeof(double) * n * nrhs, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    PIPSGPU_CATCH
}

int pipsgpu_sc_download(pipsgpu_sc* sc, double* out) {
    PIPSGPU_TRY
    if (!sc || !out) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_sc_download: null");
    require_device(sc->ctx);
    const size_t n = (size_t)sc->n;
    CUDA_CHECK(cudaMemcpyAsync(out, sc->A.p, sizeof(double) * n * n, cudaMemcpyDeviceToHost, sc->ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(sc->ctx->stream));
    for (size_t i = 0; i < n; ++i)
        for (size_t j = i + 1; j < n; ++j) out[i * n + j] = 0.0;
    PIPSGPU_CATCH
}

int pipsgpu_sc_upload(pipsgpu_sc* sc, const double* in) {
    PIPSGPU_TRY
    if (!sc || !in) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_sc_upload: null");
    require_device(sc->ctx);
    const size_t n = (size_t)sc->n;
    CUDA_CHECK(cudaMemcpyAsync(sc->A.p, in, sizeof(double) * n * n, cudaMemcpyHostToDevice, sc->ctx->stream));
    CUDA_CHECK(cudaStreamSynchronize(sc->ctx->stream));
    sc->factorized = false;
    PIPSGPU_CATCH
}

// ------------------------------------------------------------------------------------------------------------
// kkt (sLinsysRootAug)
// ------------------------------------------------------------------------------------------------------------
int pipsgpu_kkt_create(pipsgpu_ctx* ctx, int n0, int my0, int mz0, int myl, int mzl, const int* A0_krow, const int* A0_jcol,
                       const double* A0_val, const int* C0_krow, const int* C0_jcol, const double* C0_val, const int* F0_krow,
                       const int* F0_jcol, const double* F0_val, const int* G0_krow, const int* G0_jcol, const double* G0_val,
                       pipsgpu_kkt** out) {
    PIPSGPU_TRY
    require_device(ctx);
    if (!out) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_kkt_create: null output");
    auto k = std::make_unique<pipsgpu_kkt>();
    k->ctx = ctx; k->n0 = n0; k->my0 = my0; k->mz0 = mz0; k->myl = myl; k->mzl = mzl;
    cudaStream_t s = ctx->stream;
    k->A0.upload(my0, n0, A0_krow, A0_jcol, A0_val, s);
    k->C0.upload(mz0, n0, C0_krow, C0_jcol, C0_val, s);
    k->F0.upload(myl, n0, F0_krow, F0_jcol, F0_val, s);
    k->G0.upload(mzl, n0, G0_krow, G0_jcol, G0_val, s);
    pipsgpu_sc* sc = nullptr;
    const int rc = pipsgpu_sc_create(ctx, n0, my0, myl, mzl, &sc);
    if (rc != 0) return rc;
    k->sc.reset(sc);
    const size_t nb0 = (size_t)n0 + my0 + mz0 + myl + mzl;
    k->dd0.alloc((size_t)n0 + 1); k->zdiag0.alloc((size_t)mz0 + 1); k->zdiagl.alloc((size_t)mzl + 1);
    k->b0.alloc(nb0 + 1); k->r.alloc((size_t)sc->n + 1); k->t3.alloc((size_t)mz0 + 1);
    *out = k.release();
    PIPSGPU_CATCH
}

int pipsgpu_kkt_destroy(pipsgpu_kkt* k) {
    PIPSGPU_TRY
    if (k) {
        cudaSetDevice(k->ctx->device);
        cudaStreamSynchronize(k->ctx->stream);
        if (k->ev_init) for (auto& e : k->ev) cudaEventDestroy(e);
        delete k;
    }
    PIPSGPU_CATCH
}

int pipsgpu_kkt_add_leaf(pipsgpu_kkt* k, pipsgpu_leaf* lf) {
    PIPSGPU_TRY
    if (!k || !lf) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_kkt_add_leaf: null");
    if (lf->ctx != k->ctx) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_kkt_add_leaf: leaf belongs to another context");
    if (k->ready) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_kkt_add_leaf: kkt already set up");
    if (lf->border_set && (lf->n0 != k->n0 || lf->my0 != k->my0 || lf->myl != k->myl || lf->mzl != k->mzl))
        throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_kkt_add_leaf: border dimensions do not match the root");
    if ((int)k->leaves.size() != lf->index) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_kkt_add_leaf: leaves must be added in creation order");
    k->leaves.push_back(lf);
    PIPSGPU_CATCH
}

int pipsgpu_kkt_setup(pipsgpu_kkt* k, int threads) {
    PIPSGPU_TRY
    if (!k) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_kkt_setup: null");
    pipsgpu_ctx* ctx = k->ctx;
    require_device(ctx);
    if (k->leaves.size() != ctx->leaves.size()) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_kkt_setup: every leaf of the context must be added");
    if (!ctx->committed) {
        const int rc = pipsgpu_ctx_commit(ctx, threads);
        if (rc != 0) return rc;
    }
    k->sum_nx = k->sum_nz = k->sum_n = 0;
    for (auto* lf : k->leaves) { k->sum_nx += lf->nx; k->sum_nz += lf->nz; k->sum_n += lf->n; }
    // remap the border indices to the b0 layout [x0; y0; z0; yl; zl]
    if (ctx->nnzB > 0) {
        std::vector<int> h((size_t)ctx->nnzB);
        CUDA_CHECK(cudaMemcpy(h.data(), ctx->bg_b0.p, sizeof(int) * h.size(), cudaMemcpyDeviceToHost));
        for (auto& b : h) b = (b < k->n0) ? b : b + k->my0 + k->mz0;
        CUDA_CHECK(cudaMemcpy(ctx->bg_b0.p, h.data(), sizeof(int) * h.size(), cudaMemcpyHostToDevice));
    }
    k->bl.alloc((size_t)k->sum_n + 1);
    k->tl.alloc((size_t)k->sum_n + 1);
    k->wl.alloc((size_t)k->sum_n + 1);
    k->bl_in.alloc((size_t)k->sum_n + 1);
    k->b0_in.alloc((size_t)k->n0 + k->my0 + k->mz0 + k->myl + k->mzl + 1);
    if (!k->ev_init) {
        for (auto& e : k->ev) CUDA_CHECK(cudaEventCreate(&e));
        k->ev_init = true;
    }
    k->ready = true;
    PIPSGPU_CATCH
}

static void kkt_factorize_device(pipsgpu_kkt* k, double regDy, int inertia_root[3]) {
    pipsgpu_ctx* ctx = k->ctx;
    cudaStream_t s = ctx->stream;
    pipsgpu_sc* sc = k->sc.get();
    // ---- initializeKKT + leaf LDL^T + Schur terms ----
    CUDA_CHECK(cudaMemsetAsync(sc->A.p, 0, sizeof(double) * (size_t)sc->n * sc->n, s));
    launch_set_diag(ctx->diagpos.p, ctx->dvals.p, ctx->npos, ctx->Kval.p, s);
    factor_prepare(ctx, nullptr);
    CUDA_CHECK(cudaEventRecord(k->ev[1], s));
    run_factor_levels(ctx, sc, nullptr, k->ev[12], k->ev[13]);
    CUDA_CHECK(cudaEventRecord(k->ev[2], s));
    // ---- reduceKKT ----
    sc_allreduce_impl(sc);
    CUDA_CHECK(cudaEventRecord(k->ev[3], s));
    // ---- finalizeKKTdense ----
    const int n0 = k->n0, my0 = k->my0, myl = k->myl, mzl = k->mzl;
    launch_sc_add_diag(sc->A.p, sc->n, 0, n0, k->dd0.p, 1.0, s);
    launch_sc_add_ctdc(sc->A.p, sc->n, k->mz0, k->C0.krow.p, k->C0.jcol.p, k->C0.val.p, k->zdiag0.p, s);
    launch_sc_add_sparse(sc->A.p, sc->n, n0, 0, my0, k->A0.krow.p, k->A0.jcol.p, k->A0.val.p, 1.0, s);
    if (regDy != 0.0) launch_sc_add_diag(sc->A.p, sc->n, n0, my0 + myl, nullptr, -regDy, s);
    launch_sc_add_sparse(sc->A.p, sc->n, n0 + my0, 0, myl, k->F0.krow.p, k->F0.jcol.p, k->F0.val.p, 1.0, s);
    launch_sc_add_sparse(sc->A.p, sc->n, n0 + my0 + myl, 0, mzl, k->G0.krow.p, k->G0.jcol.p, k->G0.val.p, 1.0, s);
    launch_sc_add_diag(sc->A.p, sc->n, n0 + my0 + myl, mzl, k->zdiagl.p, 1.0, s);
    CUDA_CHECK(cudaEventRecord(k->ev[4], s));
    sc_factorize_impl(sc);
    CUDA_CHECK(cudaEventRecord(k->ev[5], s));
    fetch_leaf_stats(ctx);
    for (auto* lf : ctx->leaves) lf->factorized = true;
    if (inertia_root) { inertia_root[0] = sc->inertia[0]; inertia_root[1] = sc->inertia[1]; inertia_root[2] = sc->inertia[2]; }
    float ms;
    for (int i = 0; i < 5; ++i) {
        CUDA_CHECK(cudaEventElapsedTime(&ms, k->ev[i], k->ev[i + 1]));
        k->times[i] = ms;
    }
    CUDA_CHECK(cudaEventElapsedTime(&ms, k->ev[12], k->ev[13]));
    k->ktimes[1] = ms;
    k->ktimes[3] = 0.0;
    for (size_t e = 0; e < ctx->gt_rec.size(); ++e) {      // group launches are serialised on their stream: sum of kernel times
        if (!ctx->gt_rec[e]) continue;
        if (cudaEventElapsedTime(&ms, ctx->ev_gt[2 * e], ctx->ev_gt[2 * e + 1]) == cudaSuccess) k->ktimes[3] += ms;
        else (void)cudaGetLastError();
        ctx->gt_rec[e] = 0;
    }
}

int pipsgpu_kkt_factorize(pipsgpu_kkt* k, const double* dd0, const double* nomegainv0, const double* nomegainv_link,
                          const double* dd_leaves, const double* nomegainv_leaves, const double reg[3], int inertia_root[3]) {
    PIPSGPU_TRY
    if (!k || !k->ready) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_kkt_factorize: kkt not set up");
    pipsgpu_ctx* ctx = k->ctx;
    require_device(ctx);
    cudaStream_t s = ctx->stream;
    const double regP = reg ? reg[0] : 0.0, regDy = reg ? reg[1] : 0.0, regDz = reg ? reg[2] : 0.0;
    CUDA_CHECK(cudaEventRecord(k->ev[0], s));
    // ---- diagonals down (the only per-iteration H2D payload of the factorisation): the caller's vectors go to the device as
    // they are (no host staging loop, the equality rows' constant diagonal is not transferred) and are assembled there ----
    {
        const size_t nl = k->leaves.size();
        if (k->h_ldiag.size() != nl || !k->h_root) {
            if (k->h_root) { cudaFreeHost(k->h_root); k->h_root = nullptr; }
            k->h_ldiag.resize(nl);
            int64_t ddo = 0, nzo = 0;
            for (size_t i = 0; i < nl; ++i) {
                const pipsgpu_leaf* lf = k->leaves[i];
                LeafDiag& t = k->h_ldiag[i];
                t.pos_off = lf->pos_off; t.dd_off = ddo; t.nz_off = nzo; t.nx = lf->nx; t.ny = lf->ny; t.nz = lf->nz; t.pad = 0;
                ddo += lf->nx; nzo += lf->nz;
            }
            k->dd_in.alloc((size_t)k->sum_nx + 1); k->nz_in.alloc((size_t)k->sum_nz + 1);
            CUDA_CHECK(cudaMallocHost(&k->h_root, sizeof(double) * ((size_t)k->n0 + k->mz0 + k->mzl + 1)));
        }
        for (size_t i = 0; i < nl; ++i) { const pipsgpu_leaf* lf = k->leaves[i]; for (int q = 0; q < 3; ++q) k->h_ldiag[i].reg[q] = lf->reg[q]; }
        k->ldiag.ensure(nl);
        if (nl) CUDA_CHECK(cudaMemcpyAsync(k->ldiag.p, k->h_ldiag.data(), sizeof(LeafDiag) * nl, cudaMemcpyHostToDevice, s));
        if (dd_leaves && k->sum_nx > 0)
            CUDA_CHECK(cudaMemcpyAsync(k->dd_in.p, dd_leaves, sizeof(double) * (size_t)k->sum_nx, cudaMemcpyHostToDevice, s));
        if (nomegainv_leaves && k->sum_nz > 0)
            CUDA_CHECK(cudaMemcpyAsync(k->nz_in.p, nomegainv_leaves, sizeof(double) * (size_t)k->sum_nz, cudaMemcpyHostToDevice, s));
        launch_build_dvals(ctx->npos, k->ldiag.p, (int)nl, dd_leaves ? k->dd_in.p : nullptr, nomegainv_leaves ? k->nz_in.p : nullptr,
                           regP, regDy, regDz, ctx->dvals.p, s);
        double* h0 = k->h_root; double* hz = h0 + k->n0; double* hl = hz + k->mz0;
        for (int i = 0; i < k->n0; ++i) h0[i] = (dd0 ? dd0[i] : 0.0) + regP;
        for (int i = 0; i < k->mz0; ++i) hz[i] = (nomegainv0 ? nomegainv0[i] : 0.0) - regDz;
        for (int i = 0; i < k->mzl; ++i) hl[i] = (nomegainv_link ? nomegainv_link[i] : 0.0) - regDz;
        if (k->n0) CUDA_CHECK(cudaMemcpyAsync(k->dd0.p, h0, sizeof(double) * (size_t)k->n0, cudaMemcpyHostToDevice, s));
        if (k->mz0) CUDA_CHECK(cudaMemcpyAsync(k->zdiag0.p, hz, sizeof(double) * (size_t)k->mz0, cudaMemcpyHostToDevice, s));
        if (k->mzl) CUDA_CHECK(cudaMemcpyAsync(k->zdiagl.p, hl, sizeof(double) * (size_t)k->mzl, cudaMemcpyHostToDevice, s));
        // (kkt_factorize_device ends with a stream synchronisation - the inertia read-back -, so the caller's vectors and the
        // pinned staging are free again when this call returns)
    }
    k->have_diag = true;
    k->last_regDy = regDy;
    kkt_factorize_device(k, regDy, inertia_root);
    PIPSGPU_CATCH
}

int pipsgpu_kkt_factorize_resident(pipsgpu_kkt* k, int inertia_root[3]) {
    PIPSGPU_TRY
    if (!k || !k->ready) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_kkt_factorize_resident: kkt not set up");
    if (!k->have_diag) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_kkt_factorize_resident: no diagonals on the device yet");
    require_device(k->ctx);
    CUDA_CHECK(cudaEventRecord(k->ev[0], k->ctx->stream));
    kkt_factorize_device(k, k->last_regDy, inertia_root);
    PIPSGPU_CATCH
}

// out = K^-1 in over all leaves of the context with iterative refinement (see pipsgpu_ctx::refine_steps):
//   z = K^-1 b ; repeat: r = K z - b ; z -= K^-1 r.   `res` is scratch of npos doubles. ev0 / ev1 bracket the first psolve.
static void leaf_solve_refined(pipsgpu_ctx* ctx, const double* in, double* out, double* res, cudaEvent_t ev0 = nullptr,
                               cudaEvent_t ev1 = nullptr) {
    cudaStream_t s = ctx->stream;
    if (ev0) CUDA_CHECK(cudaEventRecord(ev0, s));
    ctx->psolve(in, out, 0);
    if (ev1) CUDA_CHECK(cudaEventRecord(ev1, s));
    if (!ctx->refine_kkt || ctx->refine_steps <= 0 || !(ctx->any_perturbed || ctx->refine_always)) return;
    for (int it = 0; it < ctx->refine_steps; ++it) {
        launch_axpby(ctx->npos, -1.0, in, 0.0, res, s);                                              // res = -b
        launch_sym_spmv(ctx->nnzK, ctx->kcoo_row.p, ctx->kcoo_col.p, ctx->Kval.p, 1.0, out, res, s);   // res += K z
        ctx->psolve(res, out, 1);                                                                    // z -= K^-1 res
    }
}

// solveCompressed on device-resident vectors: (in0 | inl) -> (out0 | outl); root parts have nb0 entries, leaf parts npos.
// out0 may alias in0, outl may alias inl. k->r, k->t3, k->tl, k->wl are scratch.
static void kkt_solve_device(pipsgpu_kkt* k, const double* in0, const double* inl, double* out0, double* outl) {
    pipsgpu_ctx* ctx = k->ctx;
    pipsgpu_sc* sc = k->sc.get();
    cudaStream_t s = ctx->stream;
    const int n0 = k->n0, my0 = k->my0, mz0 = k->mz0, myl = k->myl, mzl = k->mzl;
    const int nb0 = n0 + my0 + mz0 + myl + mzl;
    const int64_t N = ctx->npos;
    if (ctx->rank > 0) launch_vec_zero(out0, nb0, s);   // sLinsysRootAug::Lsolve: ranks > 0 start from zero
    else if (out0 != in0) launch_vec_copy(out0, in0, nb0, s);
    // ---- Lsolve: z_i = K_i^-1 b_i ; b0 -= Bord_i^T z_i ; allreduce ----
    leaf_solve_refined(ctx, inl, outl, k->tl.p, k->ev[10], k->ev[11]);
    launch_border_tmult(ctx->bg_row.p, ctx->bg_b0.p, ctx->Bval.p, ctx->nnzB, outl, out0, s);
    if (ctx->comm && ctx->nranks > 1) {
        NcclApi& api = NcclApi::get();
        const int rc = api.AllReduce(out0, out0, (size_t)nb0, kNcclFloat64, kNcclSum, ctx->comm, s);
        if (rc != 0) throw Error(PIPSGPU_ERR_NCCL, std::string("ncclAllReduce(b0): ") + api.GetErrorString(rc));
    }
    CUDA_CHECK(cudaEventRecord(k->ev[7], s));
    // ---- Dsolve: solveReducedLinkCons ----
    double* b = out0;
    double* r = k->r.p;
    launch_vec_copy(r, b, n0 + my0, s);
    launch_vec_copy(r + n0 + my0, b + n0 + my0 + mz0, myl + mzl, s);
    if (mz0 > 0) {
        launch_vec_copy(k->t3.p, b + n0 + my0, mz0, s);
        launch_vec_div(k->t3.p, k->zdiag0.p, mz0, s);
        launch_csr_tmult(mz0, k->C0.krow.p, k->C0.jcol.p, k->C0.val.p, -1.0, k->t3.p, r, s);
    }
    CUDA_CHECK(cudaEventRecord(k->ev[14], s));
    sc_solve_device(sc, r, 1, sc->n);
    CUDA_CHECK(cudaEventRecord(k->ev[15], s));
    launch_vec_copy(b, r, n0 + my0, s);
    if (mz0 > 0) {
        launch_csr_mult(mz0, k->C0.krow.p, k->C0.jcol.p, k->C0.val.p, -1.0, r, b + n0 + my0, s);
        launch_vec_div(b + n0 + my0, k->zdiag0.p, mz0, s);
    }
    launch_vec_copy(b + n0 + my0 + mz0, r + n0 + my0, myl + mzl, s);
    CUDA_CHECK(cudaEventRecord(k->ev[8], s));
    // ---- Ltsolve: b_i -= K_i^-1 (Bord_i u) ----
    if (N > 0) {
        launch_vec_zero(k->tl.p, N, s);
        launch_border_mult(ctx->bg_row.p, ctx->bg_b0.p, ctx->Bval.p, ctx->nnzB, b, k->tl.p, s);
        if (ctx->refine_kkt && ctx->refine_steps > 0 && (ctx->any_perturbed || ctx->refine_always)) {
            // w = K^-1 (Bord u) refined, then b_i -= w (residual scratch: ctx->work)
            leaf_solve_refined(ctx, k->tl.p, k->wl.p, ctx->work.p);
            launch_axpy(-1.0, k->wl.p, outl, N, s);
        } else {
            ctx->psolve(k->tl.p, outl, 1);
        }
    }
}
static void kkt_solve_resident(pipsgpu_kkt* k) { kkt_solve_device(k, k->b0_in.p, k->bl_in.p, k->b0.p, k->bl.p); }

static void kkt_solve_times(pipsgpu_kkt* k) {
    float ms;
    for (int i = 0; i < 3; ++i) {
        CUDA_CHECK(cudaEventElapsedTime(&ms, k->ev[6 + i], k->ev[7 + i]));
        k->times[5 + i] = ms;
    }
    CUDA_CHECK(cudaEventElapsedTime(&ms, k->ev[10], k->ev[11]));
    k->ktimes[0] = ms;
    CUDA_CHECK(cudaEventElapsedTime(&ms, k->ev[14], k->ev[15]));
    k->ktimes[2] = ms;
}

int pipsgpu_kkt_solve_compressed(pipsgpu_kkt* k, double* b0, double* b_leaves) {
    PIPSGPU_TRY
    if (!k || !k->ready) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_kkt_solve_compressed: kkt not set up");
    if (!b0) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_kkt_solve_compressed: null b0");
    pipsgpu_ctx* ctx = k->ctx;
    require_device(ctx);
    if (!k->sc->factorized) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_kkt_solve_compressed: not factorized");
    cudaStream_t s = ctx->stream;
    const int nb0 = k->n0 + k->my0 + k->mz0 + k->myl + k->mzl;
    const int64_t N = ctx->npos;
    CUDA_CHECK(cudaEventRecord(k->ev[6], s));
    if (N > 0) {
        if (!b_leaves) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_kkt_solve_compressed: null b_leaves");
        CUDA_CHECK(cudaMemcpyAsync(k->bl_in.p, b_leaves, sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, s));
    }
    CUDA_CHECK(cudaMemcpyAsync(k->b0_in.p, b0, sizeof(double) * (size_t)nb0, cudaMemcpyHostToDevice, s));
    k->have_rhs = true;
    kkt_solve_resident(k);
    if (N > 0) CUDA_CHECK(cudaMemcpyAsync(b_leaves, k->bl.p, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaMemcpyAsync(b0, k->b0.p, sizeof(double) * (size_t)nb0, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaEventRecord(k->ev[9], s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    kkt_solve_times(k);
    PIPSGPU_CATCH
}

int pipsgpu_kkt_solve_compressed_resident(pipsgpu_kkt* k) {
    PIPSGPU_TRY
    if (!k || !k->ready) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_kkt_solve_compressed_resident: kkt not set up");
    if (!k->have_rhs) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_kkt_solve_compressed_resident: no right-hand side on the device yet");
    require_device(k->ctx);
    if (!k->sc->factorized) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_kkt_solve_compressed_resident: not factorized");
    CUDA_CHECK(cudaEventRecord(k->ev[6], k->ctx->stream));
    kkt_solve_resident(k);
    CUDA_CHECK(cudaEventRecord(k->ev[9], k->ctx->stream));
    PIPSGPU_CATCH
}

// ------------------------------------------------------------------------------------------------------------
// device-resident outer solve: LinearSystem::solveCompressedBiCGStab (LS/LinearSystem.C:550-798) preconditioned by
// solveCompressed, with the full-KKT matvec LinearSystem::system_mult (:809-844) in the node layout of
// DistributedVector (root part [x0;y0;z0;yl;zl] replicated on every rank, leaf parts distributed). The right-hand side
// goes down once and the solution comes back once; only scalars (dot products) cross PCIe in between.
// ------------------------------------------------------------------------------------------------------------
namespace {

struct Krylov {
    pipsgpu_kkt* k;
    pipsgpu_ctx* ctx;
    cudaStream_t s;
    int nb0;
    int64_t N, R;      // leaf length, padded root length
    double* root(int v) { return k->kv[v].p; }
    double* leaf(int v) { return k->kv[v].p + R; }
    void copy(int dst, int src) { CUDA_CHECK(cudaMemcpyAsync(k->kv[dst].p, k->kv[src].p, sizeof(double) * (size_t)(R + N), cudaMemcpyDeviceToDevice, s)); }
    void axpby(double a, int x, double b, int y) { launch_axpby(R + N, a, k->kv[x].p, b, k->kv[y].p, s); }
    // <x, y>: the replicated root part counts once, the leaf parts are summed over the ranks
    double dot(int x, int y) {
        launch_dot(nb0, root(x), root(y), k->kpartial.p, k->kscal.p, s);
        launch_dot(N, leaf(x), leaf(y), k->kpartial.p, k->kscal.p + 1, s);
        if (ctx->comm && ctx->nranks > 1) {
            NcclApi& api = NcclApi::get();
            const int rc = api.AllReduce(k->kscal.p + 1, k->kscal.p + 1, 1, kNcclFloat64, kNcclSum, ctx->comm, s);
            if (rc != 0) throw Error(PIPSGPU_ERR_NCCL, std::string("ncclAllReduce(dot): ") + api.GetErrorString(rc));
        }
        CUDA_CHECK(cudaMemcpyAsync(k->h_scal, k->kscal.p, sizeof(double) * 2, cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        return k->h_scal[0] + k->h_scal[1];
    }
    double nrm2(int x) { return std::sqrt(std::max(0.0, dot(x, x))); }
    // out = K in  (system_mult with beta = 0, alpha = 1 on the regularised system that was factorised)
    void matvec(int out, int in) {
        const int n0 = k->n0, my0 = k->my0, mz0 = k->mz0, myl = k->myl, mzl = k->mzl;
        double* y0 = root(out); double* yl = leaf(out);
        const double* x0 = root(in); const double* xl = leaf(in);
        double* c0 = k->kc0.p;
        launch_vec_zero(yl, N, s);
        launch_vec_zero(c0, nb0, s);
        // leaves: y_i = K_i x_i + Bord_i x_0 ; root contribution sum_i Bord_i^T x_i
        launch_sym_spmv(ctx->nnzK, ctx->kcoo_row.p, ctx->kcoo_col.p, ctx->Kval.p, 1.0, xl, yl, s);
        launch_border_axpy(ctx->bg_row.p, ctx->bg_b0.p, ctx->Bval.p, ctx->nnzB, 1.0, x0, yl, false, s);
        launch_border_axpy(ctx->bg_row.p, ctx->bg_b0.p, ctx->Bval.p, ctx->nnzB, 1.0, xl, c0, true, s);
        if (ctx->comm && ctx->nranks > 1) {
            NcclApi& api = NcclApi::get();
            const int rc = api.AllReduce(c0, c0, (size_t)nb0, kNcclFloat64, kNcclSum, ctx->comm, s);
            if (rc != 0) throw Error(PIPSGPU_ERR_NCCL, std::string("ncclAllReduce(matvec): ") + api.GetErrorString(rc));
        }
        // root rows: y_0 = c0 + K_root x_0
        launch_vec_copy(y0, c0, nb0, s);
        const double* ux = x0; const double* uy = x0 + n0; const double* uz = x0 + n0 + my0;
        const double* uyl = x0 + n0 + my0 + mz0; const double* uzl = uyl + myl;
        double* rx = y0; double* ry = y0 + n0; double* rz = y0 + n0 + my0; double* ryl = y0 + n0 + my0 + mz0; double* rzl = ryl + myl;
        launch_diag_axpy(n0, 1.0, k->dd0.p, ux, rx, s);
        launch_csr_tmult(my0, k->A0.krow.p, k->A0.jcol.p, k->A0.val.p, 1.0, uy, rx, s);
        launch_csr_tmult(mz0, k->C0.krow.p, k->C0.jcol.p, k->C0.val.p, 1.0, uz, rx, s);
        launch_csr_tmult(myl, k->F0.krow.p, k->F0.jcol.p, k->F0.val.p, 1.0, uyl, rx, s);
        launch_csr_tmult(mzl, k->G0.krow.p, k->G0.jcol.p, k->G0.val.p, 1.0, uzl, rx, s);
        launch_csr_mult(my0, k->A0.krow.p, k->A0.jcol.p, k->A0.val.p, 1.0, ux, ry, s);
        launch_csr_mult(mz0, k->C0.krow.p, k->C0.jcol.p, k->C0.val.p, 1.0, ux, rz, s);
        launch_csr_mult(myl, k->F0.krow.p, k->F0.jcol.p, k->F0.val.p, 1.0, ux, ryl, s);
        launch_csr_mult(mzl, k->G0.krow.p, k->G0.jcol.p, k->G0.val.p, 1.0, ux, rzl, s);
        if (k->last_regDy != 0.0) {
            launch_diag_axpy(my0, -k->last_regDy, nullptr, uy, ry, s);
            launch_diag_axpy(myl, -k->last_regDy, nullptr, uyl, ryl, s);
        }
        launch_diag_axpy(mz0, 1.0, k->zdiag0.p, uz, rz, s);
        launch_diag_axpy(mzl, 1.0, k->zdiagl.p, uzl, rzl, s);
    }
    void precond(int out, int in) { kkt_solve_device(k, root(in), leaf(in), root(out), leaf(out)); }
    // res = b - K x ; returns ||res||_2
    double residual(int res, int b, int x) {
        matvec(res, x);
        axpby(1.0, b, -1.0, res);
        return nrm2(res);
    }
};

}  // namespace

int pipsgpu_kkt_solve(pipsgpu_kkt* k, double* b0, double* b_leaves, double tol, int max_iter, int* status, int* n_iter,
                      double* rel_residual) {
    PIPSGPU_TRY
    if (!k || !k->ready) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_kkt_solve: kkt not set up");
    if (!b0) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_kkt_solve: null b0");
    pipsgpu_ctx* ctx = k->ctx;
    require_device(ctx);
    if (!k->sc->factorized) throw Error(PIPSGPU_ERR_STATE, "pipsgpu_kkt_solve: not factorized");
    cudaStream_t s = ctx->stream;
    const int nb0 = k->n0 + k->my0 + k->mz0 + k->myl + k->mzl;
    const int64_t N = ctx->npos;
    if (N > 0 && !b_leaves) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_kkt_solve: null b_leaves");
    Krylov K{k, ctx, s, nb0, N, (int64_t)((nb0 + 1) & ~1)};
    if (k->kv[0].n == 0) {
        for (auto& v : k->kv) { v.alloc((size_t)(K.R + N) + 2); CUDA_CHECK(cudaMemsetAsync(v.p, 0, sizeof(double) * v.n, s)); }
        k->kc0.alloc((size_t)nb0 + 2); k->kpartial.alloc(DOT_SCRATCH); k->kscal.alloc(4);
        CUDA_CHECK(cudaMallocHost(&k->h_scal, sizeof(double) * 4));
    }
    k->kv_root = K.R;
    // vectors (aliases of LinearSystem.C:561-562): b, x, r, r0, dx, best_x, v, t, p
    enum { B = 0, X, Rr, R0, DX, BEST, V, T, P };
    CUDA_CHECK(cudaMemcpyAsync(K.root(B), b0, sizeof(double) * (size_t)nb0, cudaMemcpyHostToDevice, s));
    if (N > 0) CUDA_CHECK(cudaMemcpyAsync(K.leaf(B), b_leaves, sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, s));
    const double eps = 1e-15;                     // OUTER_BICG_EPSILON
    const int max_stagnations = 4, max_normr_div = 4;      // OUTER_BICG_MAX_STAGNATIONS / _MAX_NORMR_DIVERGENCES
    auto is_zero = [](double v) { return std::fabs(v) < 1e-13; };      // PIPSisZero, pips_eps
    enum { NOT_CONVERGED_MAX_ITERATIONS = -1, CONVERGED = 0, SKIPPED = 1, STAGNATION = 3, BREAKDOWN = 4, DIVERGED = 5 };
    const double bnorm = K.nrm2(B);
    const double tolb = std::max(bnorm * tol, eps);
    int flag = NOT_CONVERGED_MAX_ITERATIONS, it = 0;
    // starting point: x = K~^-1 b
    K.precond(X, B);
    double rnorm = K.residual(Rr, B, X);
    double min_rnorm = rnorm, resnorm = rnorm;
    K.copy(BEST, X);
    if (rnorm <= tolb) {
        flag = SKIPPED;
    } else {
        K.axpby(1.0 / rnorm, Rr, 0.0, R0);
        int normr_div = 0, nstag = 0;
        double rho = 1.0, omega = 1.0, alpha = 1.0;
        auto stagnation = [&](double step, double step_norm) {
            if (std::fabs(step) * step_norm <= eps * K.nrm2(X)) ++nstag; else nstag = 0;
        };
        for (it = 0; it < max_iter; ++it) {
            // ---- first half ----
            const double rho_last = rho;
            rho = K.dot(R0, Rr);
            if (is_zero(rho)) { flag = BREAKDOWN; break; }
            if (it == 0) K.copy(P, Rr);
            else {
                const double beta = (rho / rho_last) * (alpha / omega);
                if (is_zero(beta)) { flag = BREAKDOWN; break; }
                K.axpby(-omega, V, 1.0, P);       // p = p - omega v
                K.axpby(1.0, Rr, beta, P);        // p = r + beta p
            }
            K.precond(DX, P);
            K.matvec(V, DX);
            const double rtv = K.dot(R0, V);
            if (is_zero(rtv)) { flag = BREAKDOWN; break; }
            alpha = rho / rtv;
            stagnation(alpha, K.nrm2(DX));
            K.axpby(alpha, DX, 1.0, X);
            K.axpby(-alpha, V, 1.0, Rr);
            rnorm = K.nrm2(Rr);
            if (rnorm <= tolb) {
                resnorm = K.residual(Rr, B, X);
                if (resnorm <= tolb) { flag = CONVERGED; break; }
                min_rnorm = K.residual(DX, B, BEST);
                rnorm = resnorm;
            }
            if (rnorm < min_rnorm) { min_rnorm = rnorm; K.copy(BEST, X); }
            // ---- second half ----
            K.precond(DX, Rr);
            K.matvec(T, DX);
            const double tt = K.dot(T, T);
            if (is_zero(tt)) { flag = BREAKDOWN; break; }
            omega = K.dot(T, Rr) / tt;
            stagnation(omega, K.nrm2(DX));
            K.axpby(omega, DX, 1.0, X);
            K.axpby(-omega, T, 1.0, Rr);
            rnorm = K.nrm2(Rr);
            if (rnorm <= tolb || nstag >= max_stagnations) {
                resnorm = K.residual(Rr, B, X);
                if (resnorm <= tolb) { flag = CONVERGED; break; }
                min_rnorm = K.residual(DX, B, BEST);
                rnorm = resnorm;
            } else {
                if (rnorm >= min_rnorm) ++normr_div; else normr_div = 0;
                if (normr_div > max_normr_div) {
                    K.copy(X, BEST);
                    rnorm = min_rnorm;
                    resnorm = K.residual(DX, B, X);
                    flag = DIVERGED;
                    break;
                }
            }
            if (rnorm < min_rnorm) { min_rnorm = rnorm; K.copy(BEST, X); }
            if (nstag >= max_stagnations) {
                if (min_rnorm < rnorm) { rnorm = min_rnorm; K.copy(X, BEST); }
                resnorm = K.residual(DX, B, X);
                flag = STAGNATION;
                break;
            }
            if (is_zero(omega)) { flag = BREAKDOWN; break; }
        }
        if (flag == NOT_CONVERGED_MAX_ITERATIONS || flag == BREAKDOWN) resnorm = K.residual(DX, B, X);
    }
    CUDA_CHECK(cudaMemcpyAsync(b0, K.root(X), sizeof(double) * (size_t)nb0, cudaMemcpyDeviceToHost, s));
    if (N > 0) CUDA_CHECK(cudaMemcpyAsync(b_leaves, K.leaf(X), sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    k->have_rhs = true;
    if (status) *status = flag;
    if (n_iter) *n_iter = (flag == SKIPPED) ? 0 : std::min(it + 1, max_iter);      // iterations started
    if (rel_residual) *rel_residual = bnorm > 0 ? resnorm / bnorm : 0.0;
    PIPSGPU_CATCH
}

pipsgpu_sc* pipsgpu_kkt_sc(pipsgpu_kkt* k) { return k ? k->sc.get() : nullptr; }

int pipsgpu_kkt_stage_times(pipsgpu_kkt* k, double out[8]) {
    PIPSGPU_TRY
    if (!k || !out) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_kkt_stage_times: null");
    if (k->have_rhs) {
        require_device(k->ctx);
        CUDA_CHECK(cudaStreamSynchronize(k->ctx->stream));
        kkt_solve_times(k);
    }
    for (int i = 0; i < 8; ++i) out[i] = k->times[i];
    PIPSGPU_CATCH
}

int pipsgpu_kkt_kernel_times(pipsgpu_kkt* k, double out[4]) {
    PIPSGPU_TRY
    if (!k || !out) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_kkt_kernel_times: null");
    if (k->have_rhs) {
        require_device(k->ctx);
        CUDA_CHECK(cudaStreamSynchronize(k->ctx->stream));
        kkt_solve_times(k);
    }
    for (int i = 0; i < 4; ++i) out[i] = k->ktimes[i];
    PIPSGPU_CATCH
}

int pipsgpu_ctx_pool_sizes(pipsgpu_ctx* ctx, int64_t out[8]) {
    PIPSGPU_TRY
    if (!ctx || !out) throw Error(PIPSGPU_ERR_INVALID, "pipsgpu_ctx_pool_sizes: null");
    out[0] = ctx->Ltot; out[1] = ctx->Ctot; out[2] = ctx->SLtot; out[3] = ctx->nfronts; out[4] = ctx->npos;
    out[5] = ctx->nlevels; out[6] = ctx->nnzK; out[7] = ctx->nnzB;
    PIPSGPU_CATCH
}

}  // extern "C"
