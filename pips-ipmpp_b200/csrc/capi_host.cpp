// Host-only part of the C ABI (include/pipsgpu.h): symbolic analysis inspection, rank partitioning, error string.
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pipsgpu.h"
#include "common.h"
#include "symbolic.h"

namespace pipsgpu {
thread_local std::string g_last_error;
void set_error(const std::string& s) { g_last_error = s; }
}  // namespace pipsgpu

struct pipsgpu_symbolic {
    pipsgpu::LeafSymbolic S;
    int64_t nnzK = 0, nnzB = 0;
};

template <class T, class V>
static void copy_out(T* dst, const V& v) {
    if (dst && !v.empty()) std::memcpy(dst, v.data(), sizeof(T) * v.size());
}

extern "C" {

const char* pipsgpu_last_error(void) { return pipsgpu::g_last_error.c_str(); }
int pipsgpu_version(void) { return 100; }

int pipsgpu_partition(int N, int P, int* lo, int* hi) {
    // mapChildrenToNSubTrees (Readers/Distributed/DistributedTree.C:437-475): floor(N/P) each, the N mod P
    // leftovers spread evenly (by stride) over the ranks, contiguous ascending ranges.
    if (P < 1 || N < 0 || P > (N > 0 ? N : 1) || !lo || !hi) {
        pipsgpu::set_error("pipsgpu_partition: need 1 <= n_ranks <= n_children");
        return PIPSGPU_ERR_INVALID;
    }
    const int everyone_gets = N / P, n_leftovers = N % P;
    std::vector<int> per(P, everyone_gets);
    if (n_leftovers > 0) {
        const int free_in_leftover_row = P - n_leftovers;
        const int free_after_each_leftover = free_in_leftover_row / n_leftovers;
        const int additional_frees = free_in_leftover_row % n_leftovers;
        int assigned = 0;
        for (int i = 0; i < P; i += free_after_each_leftover) {
            ++per[i];
            ++i;
            if (additional_frees != assigned) { ++i; ++assigned; }
        }
    }
    int off = 0;
    for (int r = 0; r < P; ++r) {
        lo[r] = off;
        off += per[r];
        hi[r] = off;
    }
    if (off != N) {
        pipsgpu::set_error("pipsgpu_partition: internal error");
        return PIPSGPU_ERR_INVALID;
    }
    return PIPSGPU_OK;
}

int pipsgpu_symbolic_create(int n, int nx, int ny, int nz, const int* krow, const int* jcol, const double* val, int nB,
                            int64_t nnzB, const int* brow, const int* bcol, int max_width, pipsgpu_symbolic** out) {
    PIPSGPU_TRY
    if (!krow || !jcol || !out) throw std::runtime_error("pipsgpu_symbolic_create: null argument");
    auto* h = new pipsgpu_symbolic();
    pipsgpu::SymbolicOptions opt;
    if (max_width > 0) opt.max_width = max_width;
    try {
        pipsgpu::analyze_leaf(n, nx, ny, nz, krow, jcol, val, nB, nnzB, brow, bcol, opt, h->S);
    } catch (...) {
        delete h;
        throw;
    }
    h->nnzK = krow[n];
    h->nnzB = nnzB;
    *out = h;
    PIPSGPU_CATCH
}

int pipsgpu_symbolic_destroy(pipsgpu_symbolic* sym) {
    delete sym;
    return PIPSGPU_OK;
}

int pipsgpu_symbolic_sizes(pipsgpu_symbolic* sym, int64_t out[13], double* flops) {
    PIPSGPU_TRY
    if (!sym || !out) throw std::runtime_error("pipsgpu_symbolic_sizes: null argument");
    const auto& S = sym->S;
    out[0] = S.n; out[1] = S.nB; out[2] = S.ns; out[3] = S.nlevels; out[4] = (int64_t)S.rowidx.size();
    out[5] = S.Loff.empty() ? 0 : S.Loff.back(); out[6] = S.Coff.empty() ? 0 : S.Coff.back();
    out[7] = sym->nnzK; out[8] = sym->nnzB; out[9] = S.nnzL; out[10] = S.nnzL_border; out[11] = S.n_pairs;
    out[12] = (int64_t)S.child_idx.size();
    if (flops) *flops = S.flops_factor;
    PIPSGPU_CATCH
}

int pipsgpu_symbolic_export(pipsgpu_symbolic* sym, int* perm, int8_t* ptype, int8_t* sgn, int* sn_col,
                            int64_t* sn_rowptr, int* rowidx, int* relidx, int* sn_na, int* sn_nb, int* sn_parent,
                            int* sn_level, int64_t* Loff, int64_t* Coff, int64_t* kmap, int64_t* bmap, int* child_ptr,
                            int* child_idx) {
    PIPSGPU_TRY
    if (!sym) throw std::runtime_error("pipsgpu_symbolic_export: null handle");
    const auto& S = sym->S;
    copy_out(perm, S.perm); copy_out(ptype, S.ptype); copy_out(sgn, S.sgn); copy_out(sn_col, S.sn_col);
    copy_out(sn_rowptr, S.sn_rowptr); copy_out(rowidx, S.rowidx); copy_out(relidx, S.relidx);
    copy_out(sn_na, S.sn_na); copy_out(sn_nb, S.sn_nb); copy_out(sn_parent, S.sn_parent);
    copy_out(sn_level, S.sn_level); copy_out(Loff, S.Loff); copy_out(Coff, S.Coff); copy_out(kmap, S.kmap);
    copy_out(bmap, S.bmap); copy_out(child_ptr, S.child_ptr); copy_out(child_idx, S.child_idx);
    PIPSGPU_CATCH
}

int pipsgpu_symbolic_border_groups(pipsgpu_symbolic* sym, int big_u, double frac, int64_t counts[3], int* root, int* N,
                                   int* mem_ptr, int* mem_front, int64_t* pos_ptr, int* pos) {
    PIPSGPU_TRY
    if (!sym || !counts) throw std::runtime_error("pipsgpu_symbolic_border_groups: null argument");
    pipsgpu::BorderGroups G;
    pipsgpu::build_border_groups(sym->S, big_u, frac, G);
    counts[0] = (int64_t)G.root.size(); counts[1] = (int64_t)G.mem_front.size(); counts[2] = (int64_t)G.pos.size();
    copy_out(root, G.root); copy_out(N, G.N); copy_out(mem_ptr, G.mem_ptr); copy_out(mem_front, G.mem_front);
    copy_out(pos_ptr, G.pos_ptr); copy_out(pos, G.pos);
    PIPSGPU_CATCH
}

}  // extern "C"
