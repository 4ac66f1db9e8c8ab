// Device data layout and kernel launchers of the KKT hot path (sm_100a). See DESIGN.md for the layout rationale.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace pipsgpu {

constexpr int MAXW = 32;         // maximum front width (pivot columns); symbolic analysis splits wider supernodes
constexpr int DENSE_NB = 32;     // row-block size of the single-CTA dense solve (Bunch-Kaufman path)
constexpr int DENSE_PANEL = 128; // panel width of the blocked dense root LDL^T (dense.cu)
constexpr int GEMM_TILE = 128;   // CTA tile of dgemm_nt_kernel (rows of a TMA box)
constexpr int GEMM_KS = 16;      // K extent of one pipeline stage = one 128-byte swizzle row of doubles
constexpr int BIG_U = 64;        // default big-front threshold (PIPSGPU_BIG_U): fronts with at least this many off-diagonal rows take the CTA-parallel path
constexpr int TILE = 64;         // output tile of the tiled update

struct TileDesc { int front, bi, bj, grouped; };   // grouped: the border x border part of this front is left to its group

// ---- deferred border x border updates (border_group_kernel) ----
// A group is a connected piece of a block's front tree whose members' border row sets are nested in the group root's
// (border rows are never eliminated, so border(child) is a subset of border(parent)). Its share of
// SC -= Bord^T K^-1 Bord  is one SYRK over the concatenated border panels of the members in the root's index space,
// accumulated in registers and flushed once per 64x64 tile instead of once per front.
constexpr int GROUP_TILE = 64;
struct GroupMember {             // border rows of a member: Lb(t,k) = L[Lb_off + k*m + t],  (L_B D)(t,k) = W[Wb_off + k*u + t]
    int64_t Lb_off, Wb_off;      // Wb_off: the member's W21 scratch is persistent (not the per-level one), see pipsgpu_ctx_commit
    int m, u, w, col0, front, pad;
};
struct GroupDesc {
    int64_t inv_base;   // ginv[inv_base + (member*nblk + b)*64 + r] = member's border row at root position 64b+r, or -1
    int mem_begin, mem_end;
    int N, nblk;        // N = border rows of the root front, nblk = ceil(N / 64)
    int root, pad;
};
// one 64 x 64 output tile of a group; self-contained (carries what it needs of its GroupDesc and the position of the root's
// Schur-complement row indices), so that the kernel's prologue and flush have no dependent descriptor loads
struct GTile {
    int64_t inv_base;   // as GroupDesc
    int64_t rowb_off;   // rowg[rowb_off + p] = Schur-complement index of the root's border row p
    int mem_begin, nmem;
    int N, nblk;
    int bi, bj;
};

// One front (supernode) of one scenario block. All offsets index context-wide pools.
struct FrontDesc {
    int64_t Loff;      // panel: m x w column-major (ld = m), rows = [pivot cols | ancestor rows | border rows]
    int64_t Coff;      // contribution block: (na+nb) x na column-major (ld = na+nb), lower trapezoid
    int64_t rows_off;  // into rowg / relidx (na+nb entries)
    int64_t Woff;      // per-level scratch: W21 = L21*D, (na+nb) x wp(w), k-major: W21(j,k) = W[k*(na+nb) + j]
    int64_t XYoff;     // per-level scratch: X = L11^-1, wp x wp row-major
    int64_t upd_off;   // solve-phase update vector (na entries)
    int64_t inv_off;   // into invmap: for every local row of the PARENT front the row of this front's C or -1
    int w, na, nb;
    int col0;          // global position of the first pivot column
    int child_begin, child_end;   // into child_list (global front ids)
    int leaf;
    int parent;
};

struct LeafDev {
    double eps;       // absolute pivot perturbation threshold (1e-8 * ||K||)
    int pos_off;      // first global position of this leaf
    int n;
};

__host__ __device__ inline int padded_width(int w) { return w <= 8 ? 8 : (w <= 16 ? 16 : 32); }
__host__ __device__ inline int width_class(int w) { return w <= 8 ? 0 : (w <= 16 ? 1 : 2); }
constexpr int NCLS = 3;          // launches are grouped by (tree level, width class)

struct FactorArgs {
    const FrontDesc* fronts;
    const int* level_fronts;     // front ids grouped by level, deepest level first
    const int* fl_fronts;        // front ids grouped by (level, small|big) for the factorisation launches
    const int* child_list;
    const int* rowg;             // per front row: global position (non-border) or Schur-complement index (border)
    const int* relidx;           // per front row: local row index in the parent front
    const int* invmap;           // per front: inverse of relidx over the parent's local rows (see FrontDesc::inv_off)
    const int8_t* ptype;
    const int8_t* sgn;
    const LeafDev* leaves;
    double* L;
    double* C;
    double* W;                   // per-level scratch
    double* XY;                  // per-level scratch
    double* Dd; double* De;      // D      (diagonal / coupling to the next position for 2x2 pivots)
    double* Did; double* Die;    // D^-1
    double* SC;                  // dense Schur complement, lower, row-major (may be null)
    int ldsc;
    int* stats;                  // per leaf: {positive, negative, zero, perturbed}
    const TileDesc* tiles;       // update tiles of the big fronts, grouped by level
    const GroupDesc* groups;
    const GroupMember* gmembers;
    const int* ginv;
    int bp_threads;              // CTA size of big_panel_kernel (128 or 256)
    int group_pipe;              // PIPSGPU_GROUP_PIPE=1: software-pipelined member staging in border_group_kernel
    int group_occ;               // PIPSGPU_GROUP_OCC: CTAs per SM border_group_kernel is compiled for (4, 5 or 6)
    int panel_staged;            // PIPSGPU_PANEL=staged: panel chunks staged in shared memory (kernels.cu: panel_chunk_staged)
    const int8_t* gflags;        // per (group member, 64-row block of the root's border index space): has rows there
    const GTile* gtiles;
};

struct SolveArgs {
    const FrontDesc* fronts;
    const int* level_fronts;
    const int* child_list;
    const int* rowg;
    const int* relidx;
    const int8_t* ptype;
    const double* L;
    const double* Did; const double* Die;
    double* x;        // permuted right-hand sides, nrhs x x_stride
    double* upd;      // update vectors, nrhs x upd_stride
    int64_t x_stride, upd_stride;
};

// ---- streaming persistent solve (psolve.cu): one launch = K^-1 for every block of the GPU ----
constexpr int PS_WARPS = 9;      // warps per CTA, one CTA per SM (9 x 24 KB of staging + 12 KB static = 224 KB)
constexpr int PS_BUF = 1184;     // doubles per staging buffer (9.25 KB, holds 32 rows of a 32-wide front), two per warp
constexpr int PS_HDR = 96;       // chunk header: three 32-wide rows of D^-1 coefficients
constexpr int PS_XA = 128;       // max rows per chunk
constexpr int PS_SMAX = 512;     // max columns of a subtree unit (its right-hand-side slice lives in shared memory)
constexpr int PS_NOUT = 128;     // rows above a unit that are accumulated in shared memory (the rest: global atomics); fronts have <= ~80 ancestor rows
constexpr int PS_WARP_DOUBLES = 2 * PS_BUF + PS_SMAX + PS_NOUT;
struct SolveFront {              // one per front, fronts of a unit contiguous and in postorder
    int64_t blob;                // offset (doubles, even) of the front's chunks in the solve pool
    int col0;                    // global position of the first pivot column
    int rows_off;                // into rowg / slots (first na entries = ancestor rows)
    int w, na, ldr;              // ldr = w | 1 (odd leading dimension)
    int R;                       // rows per full chunk = ps_rows_per_chunk(ldr)   (kept here: the kernel must not divide)
    int nch;                     // number of chunks = ceil((w + na) / R)
    int front;                   // global front id (FrontDesc index)
};
struct SolveUnit {               // a bottom subtree (<= PS_SMAX columns) or one front above the cut; one warp per unit
    int first, nfronts;          // range of SolveFront
    int pbeg, ncols;             // contiguous range of global positions eliminated by the unit
    int out_rows_off, nout;      // rowg range of the root front's ancestor rows kept in shared memory (<= PS_NOUT)
    int parent;                  // unit of the root front's parent or -1
    int nchild;                  // number of child units
};
struct PStreamArgs {
    const SolveFront* sf;
    const SolveUnit* units;      // forward ticket order: subtree units (largest first), then top units deepest level first
    int nunits;
    const double* SL;            // solve pool
    const int* rowg;
    const int* gidx;             // per position: index of that row in the caller's (unpermuted, concatenated) vectors
    const double* in;            // right-hand side, original order
    double* out;                 // solution, original order
    double* x;                   // work: permuted order
    double* acc;                 // per position: updates pushed by descendant units (zero between solves)
    int* ucnt;                   // per unit: completed child units, cumulative over solves
    int* uflag;                  // per unit: 2*epoch+1 = forward sweep of the tree done (roots), 2*epoch+2 = backward step done
    int* ticket;                 // [0] next ticket, [1] warps that have left (re-arms [0] for the next launch)
    int epoch;
    int out_mode;                // 0: out = K^-1 in, 1: out -= K^-1 in
    int xform;                   // 1: the pool's pivot-block rows hold X = L11^-1 (unit lower) instead of L11
};
__host__ __device__ int ps_rows_per_chunk(int ldr);
__host__ __device__ int ps_chunk_doubles(int ldr);
__host__ __device__ int ps_chunk_rows(int ldr, int rows, int c);
__host__ __device__ int ps_chunk_doubles_rows(int ldr, int cr);
__host__ __device__ int64_t ps_front_doubles(int ldr, int rows);
void launch_psolve_stream(const PStreamArgs& a, int sm_count, cudaStream_t s);
void launch_pack_solve(const SolveFront* sf, const int* chunk_front, const int* chunk_idx, int nchunks, const FrontDesc* fronts,
                       const double* L, const double* Did, const double* Die, const int8_t* ptype, const uint16_t* slots,
                       double* SL, const double* XY, cudaStream_t s);   // XY != null: pivot-block rows = X = L11^-1
void count_launch(const char* what);    // counts the launch and throws on cudaGetLastError() != cudaSuccess
cudaError_t init_psolve_attributes();     // per-device function attributes (dynamic shared memory opt-in)
cudaError_t init_kernel_attributes();

// ---- leaf path ----
void launch_scatter_values(const int64_t* map, const double* val, int64_t nnz, double* L, cudaStream_t s);
void launch_set_diag(const int64_t* diagpos, const double* dvals, int64_t n, double* kval, cudaStream_t s);
void launch_pivot(const FactorArgs& a, int begin, int nfronts, int cls, cudaStream_t s);    // fronts level_fronts[begin..) of width class cls
void launch_small_fronts(const FactorArgs& a, int begin, int nfronts, cudaStream_t s);      // fronts fl_fronts[begin..)
void launch_big_panel(const FactorArgs& a, int begin, int nfronts, int max_u, cudaStream_t s);   // max_u: tallest front of the range
void launch_front_update_tiles(const FactorArgs& a, int64_t tile_begin, int64_t ntiles, cudaStream_t s);
void launch_border_groups(const FactorArgs& a, int64_t gtile_begin, int64_t ngtiles, cudaStream_t s);
void launch_gather_perm(const int* gidx, const double* src, double* dst, int64_t n, int nrhs, int64_t stride, cudaStream_t s);
void launch_scatter_perm(const int* gidx, const double* src, double* dst, int64_t n, int nrhs, int64_t stride, cudaStream_t s);
void launch_fwd_level(const SolveArgs& a, int cls, int level_begin, int nfronts, int nrhs, cudaStream_t s);
void launch_bwd_level(const SolveArgs& a, int cls, int level_begin, int nfronts, int nrhs, cudaStream_t s);
void launch_diag_solve(const SolveArgs& a, int64_t npos, int nrhs, cudaStream_t s);
// b0[bidx[k]] -= val[k] * z[brow[k]]   /   t[brow[k]] += val[k] * u[bidx[k]]
void launch_border_tmult(const int* brow, const int* bidx, const double* val, int64_t nnz, const double* z, double* b0, cudaStream_t s);
void launch_border_mult(const int* brow, const int* bidx, const double* val, int64_t nnz, const double* u, double* t, cudaStream_t s);
void launch_axpy(double alpha, const double* x, double* y, int64_t n, cudaStream_t s);

// ---- dense FP64 GEMM on the tensor pipe with TMA-staged operands (dense.cu) ----
struct alignas(64) TMap { unsigned char bytes[128]; };      // CUtensorMap (cuda.h) without the driver header
void make_tmap(TMap* out, const double* base, int64_t rows, int64_t cols, int64_t ld);   // row-major, box 16 cols x 128 rows
enum { GEMM_EPI_TRAIL = 0, GEMM_EPI_PANEL = 1, GEMM_EPI_SCHUR = 2 };
struct GemmArgs {
    int a_row0, a_col0;      // TMA coordinates of the A tile of (bi, kt): row a_row0 + 128 bi, column a_col0 + 16 kt
    int b_row0, b_col0;      // the same for B (tile column bj)
    int K;
    int m_tiles, n_tiles;
    int lower;               // 1: triangular tile grid (bj <= bi) over m_tiles x m_tiles
    double* C; int64_t ldc;  // TRAIL / PANEL: the factor; SCHUR: the dense Schur complement
    int c_row0, c_col0;      // global row / column of tile (0,0) (PANEL: c_col0 = first column of the panel)
    int c_rows, c_cols;      // TRAIL / PANEL: row limit (n); PANEL: panel width; SCHUR: number of packed rows
    double* C2; int64_t ldc2;    // PANEL: W21 buffer (indexed by global row)
    const double* scale;     // PANEL: D^-1 of the panel's pivots
    const int* map;          // SCHUR: Schur-complement index of every packed row
};
void launch_dgemm_nt(int epi, const TMap& tmA, const TMap& tmB, const GemmArgs& g, cudaStream_t s);
void launch_ldlt_diag(double* A, int64_t ld, int kb, int bw, int n0, double eps_rel, const double* scale, double* D, double* Dinv,
                      double* X, int* stats, cudaStream_t s);
void launch_dense_trisolve(const double* F, int64_t ld, int n, const double* Xall, const double* Dinv, double* x, double* y,
                           int* flagF, int* flagB, int epoch, int sm_count, cudaStream_t s);
cudaError_t init_dense_attributes();

// ---- dense root path (lower triangle, row-major) ----
void launch_sc_add_sparse(double* SC, int ld, int row0, int col0, int m, const int* krow, const int* jcol, const double* val, double alpha, cudaStream_t s);
void launch_sc_add_diag(double* SC, int ld, int off, int n, const double* d, double alpha, cudaStream_t s);
// SC[0:n0,0:n0] (lower) -= C^T diag(1/zdiag) C   for CSR C (mz x n0)
void launch_sc_add_ctdc(double* SC, int ld, int mz, const int* krow, const int* jcol, const double* val, const double* zdiag, cudaStream_t s);
void launch_sc_pack(const double* SC, int ld, int n0, int my0, int nl, double* buf, bool unpack, double* SCout, cudaStream_t s);
int64_t sc_pack_size(int n0, int nl);
void launch_row_absmax(const double* A, int ld, int n, double* out, cudaStream_t s);
void launch_dense_solve(const double* A, int ld, int n, const int* perm, const int8_t* ptype, const double* Dinv,
                        const double* Dinv_e, double* x, int nrhs, int64_t xstride, double* work, cudaStream_t s);
// Bunch-Kaufman fallback (dsytf2 restated): device-side step counter
struct BkState { int k, kcur, s, npos, nneg, npert; };
void launch_bk_step(double* A, int ld, int n, int step, int* perm, int8_t* ptype, double* Dd, double* De, double* Did,
                    double* Die, double* Wc, BkState* st, double eps_rel, const double* scale, cudaStream_t s);
void launch_iota(int* p, int n, cudaStream_t s);

// ---- small vector / CSR helpers for the root-level solve ----
void launch_csr_mult(int m, const int* krow, const int* jcol, const double* val, double alpha, const double* x, double* y, cudaStream_t s);    // y += alpha*A x
void launch_csr_tmult(int m, const int* krow, const int* jcol, const double* val, double alpha, const double* x, double* y, cudaStream_t s);   // y += alpha*A^T x
void launch_vec_div(double* x, const double* d, int n, cudaStream_t s);   // x /= d
void launch_vec_copy(double* dst, const double* src, int n, cudaStream_t s);
void launch_vec_zero(double* dst, int64_t n, cudaStream_t s);

// ---- outer solve on the device (krylov.cu) ----
void launch_sym_spmv(int64_t nnz, const int* row, const int* col, const double* val, double alpha, const double* x, double* y,
                     cudaStream_t s);       // y += alpha * K x, K symmetric, lower COO in global positions
void launch_border_axpy(const int* brow, const int* bidx, const double* val, int64_t nnz, double alpha, const double* u, double* t,
                        bool transpose, cudaStream_t s);
void launch_axpby(int64_t n, double a, const double* x, double b, double* y, cudaStream_t s);       // y = a x + b y
void launch_diag_axpy(int n, double alpha, const double* d, const double* x, double* y, cudaStream_t s);   // y += alpha d .* x
void launch_dot(int64_t n, const double* x, const double* y, double* partial, double* out, cudaStream_t s);   // deterministic
constexpr int DOT_SCRATCH = 1024;
struct LeafDiag { int64_t pos_off, dd_off, nz_off; int nx, ny, nz, pad; double reg[3]; };
void launch_build_dvals(int64_t npos, const LeafDiag* tab, int nl, const double* dd, const double* nz, double regP, double regDy,
                        double regDz, double* dvals, cudaStream_t s);
void launch_gather_idx(const double* A, const int64_t* idx, int64_t n, double* out, cudaStream_t s);    // out[e] = A[idx[e]]
void launch_scatter_idx(double* A, const int64_t* idx, int64_t n, const double* in, cudaStream_t s);    // A[idx[e]] = in[e]
void launch_absmax_eps(const double* v, int64_t n, double rel, double* eps_out, cudaStream_t s);        // *eps_out = rel * max|v|

int64_t kernel_launch_count();
// FP64 tensor-pipe peak microbenchmark (register-resident DMMA chains); returns the flops of the launch
double launch_dmma_peak(double* out, int iters, int ctas, cudaStream_t s);

}  // namespace pipsgpu
