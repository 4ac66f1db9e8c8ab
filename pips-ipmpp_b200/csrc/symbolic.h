// Host-side symbolic analysis of one scenario block's KKT matrix (done once, reused across IPM iterations).
//
// Replaces the analysis phase the reference delegates to its sparse-solver plugins
// (PardisoSolver.C:141-205 phase 11 with METIS iparm[1]=2 and weighted matching iparm[12];
//  Ma57Solver.C:56-78 ma57ad with METIS icntl[5]=4). Design (B200-first):
//   * structural 2x2 pivots: every equality-row node (zero diagonal) is matched to a primal column and the pair is
//     eliminated as one static 2x2 block, so no dynamic pivoting is needed on the device;
//   * nested dissection (METIS) on the pair-compressed graph, etree postorder, relaxed supernodes of bounded width;
//   * the border Bord_i (linking columns) is carried as extra, never-eliminated rows of every front, so the
//     multifrontal elimination itself produces  -Bord_i^T K_i^-1 Bord_i  (what PARDISO's iparm[37] Schur option does,
//     PardisoProjectSchurSolver.C:246-280);
//   * fronts are grouped by tree depth: one batched kernel launch per level over all blocks of a GPU.
#pragma once
#include <cstdint>
#include <vector>

namespace pipsgpu {

struct LeafSymbolic {
    int n = 0;            // n_K = n_x + n_y + n_z
    int nx = 0, ny = 0, nz = 0;
    int nB = 0;           // border index space (row ids n .. n+nB-1 in front row lists)
    std::vector<int> perm;      // perm[p]  = original index eliminated at position p
    std::vector<int> iperm;     // iperm[i] = position of original index i
    std::vector<int8_t> ptype;  // per position: 0 = 1x1 pivot, 1 = first of a 2x2 pivot, 2 = second
    std::vector<int8_t> sgn;    // per position: expected pivot sign (+1 primal, -1 dual)

    int ns = 0;                          // number of supernodes (= fronts)
    std::vector<int> sn_col;             // ns+1, first position of each supernode
    std::vector<int64_t> sn_rowptr;      // ns+1, into rowidx
    std::vector<int> rowidx;             // off-diagonal rows of each front, ascending; border rows are n + b
    std::vector<int> sn_na, sn_nb;       // #non-border / #border rows
    std::vector<int> sn_parent;          // parent front or -1
    std::vector<int> sn_level;           // depth in the assembly forest (roots = 0)
    std::vector<int> child_ptr, child_idx;
    std::vector<int> relidx;             // aligned with rowidx: local row index inside the parent front (w_p offset incl.)
    std::vector<int64_t> Loff;           // ns+1: panel offsets; panel = m x w column-major (ld = m), m = w+na+nb
    std::vector<int64_t> Coff;           // ns+1: contribution block offsets; (na+nb) x na column-major (ld = na+nb)
    std::vector<int64_t> kmap;           // per lower-CSR entry of K: offset in the L pool
    std::vector<int64_t> bmap;           // per border triplet: offset in the L pool
    int nlevels = 0;
    int max_w = 0, max_m = 0;

    // statistics (SURVEY.md 8(d) canonical work counts)
    int64_t nnzL = 0;        // eta(L) + n   (scalar factor entries, K part only)
    int64_t nnzL_border = 0; // entries of the border rows  L_{border,:}
    double flops_factor = 0; // sum_j eta_j (eta_j + 2)
    int n_pairs = 0;
};

struct SymbolicOptions {
    int max_width = 32;      // supernodes wider than this are split into chains
    int use_metis = 1;
    int relax = 1;
    double relax_small = 0.8;  // admissible fraction of explicit zeros when merging into a supernode of <= 16 columns
    double relax_mid = 0.25;   // ... of <= 48 columns (0.1 -> 0.25: -2 % on the C2 / C3 / C4 step, profiles/README.md round 2)
    int subtree_width = 32;  // bottom subtrees with at most this many columns are merged into one front (0 = off)
};

// krow/jcol/val: lower CSR (0-based, full diagonal present) of the n x n leaf KKT matrix, block sizes nx/ny/nz.
// border triplets: (brow[k] in [0,n), bcol[k] in [0,nB)).
void analyze_leaf(int n, int nx, int ny, int nz, const int* krow, const int* jcol, const double* val, int nB,
                  int64_t nnzB, const int* brow, const int* bcol, const SymbolicOptions& opt, LeafSymbolic& out);

// Grouping of the big fronts (>= big_u off-diagonal rows, at least one border row) for the deferred border x border
// updates (kernels.cuh: GroupDesc). Border rows are never eliminated, so border(child) is a subset of border(parent):
// a front joins its parent's group if the parent is grouped and the front has at least frac * (border rows of the
// group root) border rows, otherwise it opens a group of its own. pos maps every border row of a member to the index
// of the same border row in the root's list (strictly ascending).
struct BorderGroups {
    std::vector<int> root, N;          // per group: root front, its number of border rows
    std::vector<int> mem_ptr;          // ngroups+1 into mem_front
    std::vector<int> mem_front;        // members of each group, root first, then descendants (parents before children)
    std::vector<int64_t> pos_ptr;      // per member (+1): into pos
    std::vector<int> pos;
    std::vector<char> grouped;         // per front
};
void build_border_groups(const LeafSymbolic& S, int big_u, double frac, BorderGroups& out);

}  // namespace pipsgpu
