// Symbolic analysis of a leaf KKT matrix: matching -> compressed-graph nested dissection -> postorder ->
// column structures (with border rows) -> relaxed, width-bounded supernodes -> front tree / levels / assembly maps.
// See symbolic.h for the design notes and the reference anchors.
#include "symbolic.h"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <stdexcept>

extern "C" {
// libmetis_static.a shipped with the CUDA toolkit (cuSOLVER's METIS 5, 64-bit idx_t); prototype declared by hand.
int METIS_NodeND(int64_t* nvtxs, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* options, int64_t* perm, int64_t* iperm);
int METIS_SetDefaultOptions(int64_t* options);
}

namespace pipsgpu {

namespace {

// maximum-cardinality matching of the equality rows (zero diagonal) to primal columns, heavy edges first
void match_rows(int n, int nx, int ny, const std::vector<int>& xadj, const std::vector<int>& adj,
                const std::vector<double>& adjw, std::vector<int>& mate) {
    mate.assign(n, -1);
    // greedy pass: rows pick their heaviest still-free column
    for (int y = nx; y < nx + ny; ++y) {
        int best = -1;
        double bw = 0.0;
        for (int p = xadj[y]; p < xadj[y + 1]; ++p) {
            const int x = adj[p];
            if (x >= nx || mate[x] >= 0) continue;
            if (adjw[p] > bw) { bw = adjw[p]; best = x; }
        }
        if (best >= 0) { mate[y] = best; mate[best] = y; }
    }
    // augmenting paths (iterative DFS) for the rows the greedy pass left out
    std::vector<int> stamp(n, -1), stack_y, stack_p, from_x(n, -1);
    for (int y0 = nx; y0 < nx + ny; ++y0) {
        if (mate[y0] >= 0) continue;
        stack_y.assign(1, y0);
        stack_p.assign(1, xadj[y0]);
        stamp[y0] = y0;
        int found_x = -1;
        while (!stack_y.empty() && found_x < 0) {
            const int y = stack_y.back();
            int& p = stack_p.back();
            bool pushed = false;
            for (; p < xadj[y + 1]; ++p) {
                const int x = adj[p];
                if (x >= nx || adjw[p] == 0.0 || stamp[x] == y0) continue;
                stamp[x] = y0;
                from_x[x] = y;
                if (mate[x] < 0) { found_x = x; break; }
                const int y2 = mate[x];
                if (stamp[y2] == y0) continue;
                stamp[y2] = y0;
                ++p;
                stack_y.push_back(y2);
                stack_p.push_back(xadj[y2]);
                pushed = true;
                break;
            }
            if (found_x >= 0) break;
            if (!pushed) { stack_y.pop_back(); stack_p.pop_back(); }
        }
        if (found_x >= 0) {
            int x = found_x;
            while (true) {
                const int y = from_x[x];
                const int prev = mate[y];
                mate[y] = x;
                mate[x] = y;
                if (y == y0) break;
                x = prev;
            }
        }
    }
}

}  // namespace

void analyze_leaf(int n, int nx, int ny, int nz, const int* krow, const int* jcol, const double* val, int nB,
                  int64_t nnzB, const int* brow, const int* bcol, const SymbolicOptions& opt_in, LeafSymbolic& S) {
    SymbolicOptions opt = opt_in;
    if (const char* e = std::getenv("PIPSGPU_SUBTREE_WIDTH")) opt.subtree_width = std::atoi(e);   // tuning knobs
    if (const char* e = std::getenv("PIPSGPU_RELAX_SMALL")) opt.relax_small = std::atof(e);
    if (const char* e = std::getenv("PIPSGPU_RELAX_MID")) opt.relax_mid = std::atof(e);
    if (nx + ny + nz != n) throw std::runtime_error("analyze_leaf: nx+ny+nz != n");
    S = LeafSymbolic();
    S.n = n; S.nx = nx; S.ny = ny; S.nz = nz; S.nB = nB;
    const int64_t nnzK = krow[n];

    // ---- 1. symmetric adjacency of K without the diagonal -------------------------------------------
    std::vector<int> xadj(n + 1, 0);
    for (int i = 0; i < n; ++i)
        for (int p = krow[i]; p < krow[i + 1]; ++p) {
            const int j = jcol[p];
            if (j > i) throw std::runtime_error("analyze_leaf: matrix is not lower triangular CSR");
            if (j != i) { ++xadj[i + 1]; ++xadj[j + 1]; }
        }
    for (int i = 0; i < n; ++i) xadj[i + 1] += xadj[i];
    std::vector<int> adj(xadj[n]);
    std::vector<double> adjw(xadj[n]);
    {
        std::vector<int> fill(xadj.begin(), xadj.end() - 1);
        for (int i = 0; i < n; ++i)
            for (int p = krow[i]; p < krow[i + 1]; ++p) {
                const int j = jcol[p];
                if (j == i) continue;
                const double w = val ? std::fabs(val[p]) : 1.0;
                adj[fill[i]] = j; adjw[fill[i]++] = w;
                adj[fill[j]] = i; adjw[fill[j]++] = w;
            }
    }

    // ---- 2. structural 2x2 pivots: match equality rows to primal columns -------------------------------
    std::vector<int> mate;
    match_rows(n, nx, ny, xadj, adj, adjw, mate);

    // ---- 3. pair-compressed graph ---------------------------------------------------------------------
    std::vector<int> cid(n, -1);
    int nc = 0;
    for (int i = 0; i < n; ++i) {
        if (cid[i] >= 0) continue;
        cid[i] = nc;
        if (mate[i] >= 0) cid[mate[i]] = nc;
        ++nc;
    }
    std::vector<int> cfirst(nc, -1), csecond(nc, -1);   // members: primal first, dual second
    for (int i = 0; i < n; ++i) {
        const int c = cid[i];
        if (mate[i] < 0) cfirst[c] = i;
        else if (i < nx) cfirst[c] = i;
        else csecond[c] = i;
    }
    std::vector<int64_t> cxadj(nc + 1, 0), cadj, cvw(nc);
    {
        std::vector<int> mark(nc, -1);
        cadj.reserve(adj.size());
        for (int c = 0; c < nc; ++c) {
            mark[c] = c;
            const int mem[2] = {cfirst[c], csecond[c]};
            cvw[c] = (csecond[c] >= 0) ? 2 : 1;
            for (int k = 0; k < 2; ++k) {
                const int i = mem[k];
                if (i < 0) continue;
                for (int p = xadj[i]; p < xadj[i + 1]; ++p) {
                    const int d = cid[adj[p]];
                    if (mark[d] != c) { mark[d] = c; cadj.push_back(d); }
                }
            }
            cxadj[c + 1] = (int64_t)cadj.size();
        }
    }

    // ---- 4. fill-reducing order of the compressed graph --------------------------------------------------
    std::vector<int64_t> cperm(nc), ciperm(nc);
    bool metis_ok = false;
    if (opt.use_metis && nc >= 8 && !cadj.empty()) {
        int64_t options[40];
        METIS_SetDefaultOptions(options);
        int64_t nv = nc;
        const int rc = METIS_NodeND(&nv, cxadj.data(), cadj.data(), cvw.data(), options, cperm.data(), ciperm.data());
        metis_ok = (rc == 1);
    }
    if (!metis_ok) std::iota(cperm.begin(), cperm.end(), 0);

    std::vector<int> perm0(n), iperm0(n);
    {
        int pos = 0;
        for (int k = 0; k < nc; ++k) {
            const int c = (int)cperm[k];
            perm0[pos++] = cfirst[c];
            if (csecond[c] >= 0) perm0[pos++] = csecond[c];
        }
        assert(pos == n);
        for (int p = 0; p < n; ++p) iperm0[perm0[p]] = p;
    }

    // ---- 5. elimination tree of the permuted matrix (Liu), then a postorder ---------------------------
    std::vector<int> parent0(n, -1);
    {
        std::vector<int> anc(n, -1);
        for (int j = 0; j < n; ++j) {
            const int oj = perm0[j];
            for (int p = xadj[oj]; p < xadj[oj + 1]; ++p) {
                int i = iperm0[adj[p]];
                while (i >= 0 && i < j) {
                    const int nxt = anc[i];
                    anc[i] = j;
                    if (nxt < 0) { parent0[i] = j; break; }
                    i = nxt;
                }
            }
        }
    }
    std::vector<int> post(n);   // post[k] = position (in order 0) visited k-th
    {
        std::vector<int> sz(n, 1);
        for (int j = 0; j < n; ++j) if (parent0[j] >= 0) sz[parent0[j]] += sz[j];
        // children lists; ordering: pair partner (j-1 matched with j) last, otherwise larger subtrees first
        std::vector<int> cptr(n + 2, 0);
        for (int j = 0; j < n; ++j) ++cptr[(parent0[j] < 0 ? n : parent0[j]) + 1];
        for (int j = 0; j <= n; ++j) cptr[j + 1] += cptr[j];
        std::vector<int> clist(n), cfill(cptr.begin(), cptr.end() - 1);
        for (int j = 0; j < n; ++j) clist[cfill[parent0[j] < 0 ? n : parent0[j]]++] = j;
        auto is_partner = [&](int child, int par) {
            return par < n && mate[perm0[child]] == perm0[par];
        };
        for (int v = 0; v <= n; ++v) {
            std::stable_sort(clist.begin() + cptr[v], clist.begin() + cptr[v + 1], [&](int a, int b) {
                const bool pa = is_partner(a, v), pb = is_partner(b, v);
                if (pa != pb) return pb;          // partner goes last
                return sz[a] > sz[b];             // small subtrees next to the parent: they are the ones that merge
            });
        }
        // iterative DFS from the virtual root n
        int k = 0;
        std::vector<int> st, it;
        st.push_back(n); it.push_back(cptr[n]);
        while (!st.empty()) {
            const int v = st.back();
            int& p = it.back();
            if (p < cptr[v + 1]) {
                const int c = clist[p++];
                st.push_back(c); it.push_back(cptr[c]);
            } else {
                if (v < n) post[k++] = v;
                st.pop_back(); it.pop_back();
            }
        }
        assert(k == n);
    }
    S.perm.resize(n); S.iperm.resize(n);
    for (int k = 0; k < n; ++k) S.perm[k] = perm0[post[k]];
    for (int k = 0; k < n; ++k) S.iperm[S.perm[k]] = k;
    const std::vector<int>& perm = S.perm;
    const std::vector<int>& iperm = S.iperm;

    S.ptype.assign(n, 0); S.sgn.assign(n, 0);
    for (int k = 0; k < n; ++k) {
        const int o = perm[k];
        S.sgn[k] = (o < nx) ? 1 : -1;
        if (mate[o] >= 0) {
            if (o < nx) {
                if (k + 1 >= n || perm[k + 1] != mate[o]) throw std::runtime_error("analyze_leaf: matched pair not adjacent");
                S.ptype[k] = 1; S.ptype[k + 1] = 2; ++S.n_pairs;
            }
        }
    }

    // ---- 6. column structures in the final order, border rows included ---------------------------------
    std::vector<int64_t> bptr(n + 1, 0);
    for (int64_t k = 0; k < nnzB; ++k) ++bptr[brow[k] + 1];
    for (int i = 0; i < n; ++i) bptr[i + 1] += bptr[i];
    std::vector<int> bidx(nnzB);
    {
        std::vector<int64_t> f(bptr.begin(), bptr.end() - 1);
        for (int64_t k = 0; k < nnzB; ++k) bidx[f[brow[k]]++] = bcol[k];
    }
    std::vector<std::vector<int>> cs(n);       // struct(j): rows > j (positions) and border rows n+b, sorted
    std::vector<int> parent(n, -1), nchild(n, 0), cnt(n, 0), cnt_nb(n, 0);
    {
        std::vector<int> mark(n + nB, -1);
        std::vector<int> chead(n, -1), cnext(n, -1);   // children linked lists
        for (int j = 0; j < n; ++j) {
            std::vector<int>& s = cs[j];
            mark[j] = j;
            const int oj = perm[j];
            for (int p = xadj[oj]; p < xadj[oj + 1]; ++p) {
                const int i = iperm[adj[p]];
                if (i > j && mark[i] != j) { mark[i] = j; s.push_back(i); }
            }
            for (int64_t p = bptr[oj]; p < bptr[oj + 1]; ++p) {
                const int r = n + bidx[p];
                if (mark[r] != j) { mark[r] = j; s.push_back(r); }
            }
            for (int c = chead[j]; c >= 0; c = cnext[c]) {
                for (int r : cs[c])
                    if (r != j && mark[r] != j) { mark[r] = j; s.push_back(r); }
            }
            std::sort(s.begin(), s.end());
            cnt[j] = (int)s.size();
            int nb = 0;
            while (nb < cnt[j] && s[nb] < n) ++nb;
            cnt_nb[j] = nb;
            if (nb > 0) {
                parent[j] = s[0];
                ++nchild[s[0]];
                cnext[j] = chead[s[0]];
                chead[s[0]] = j;
            }
        }
    }
    for (int j = 0; j < n; ++j) {
        S.nnzL += 1 + cnt_nb[j];
        S.nnzL_border += cnt[j] - cnt_nb[j];
        S.flops_factor += (double)cnt_nb[j] * (cnt_nb[j] + 2.0);
    }

    // ---- 7. supernodes: fundamental + forced pairs, relaxed amalgamation, width-bounded splitting -------
    std::vector<int> desc(n, 1);   // subtree sizes in the elimination tree of the final order
    for (int j = 0; j < n; ++j) if (parent[j] >= 0) desc[parent[j]] += desc[j];
    struct SN { int a, b; double zeros; };
    std::vector<SN> sn;
    auto merge_ok = [&](const SN& c, const SN& p) {
        const int wc = c.b - c.a, wp = p.b - p.a;
        const int nscol = wc + wp;
        if (nscol > opt.max_width) return false;
        const int par = parent[c.b - 1];
        if (par < p.a || par >= p.b) return false;
        if (!opt.relax) return false;
        const double up = cnt[p.b - 1], uc = cnt[c.b - 1];
        const double z = (double)wc * (wp + up - uc) + c.zeros + p.zeros;
        const double tot = 0.5 * nscol * (nscol + 1.0) + (double)nscol * up;
        const double zf = z / tot;
        if (nscol <= 4) return true;
        if (nscol <= 16) return zf < opt.relax_small;
        if (nscol <= 48) return zf < opt.relax_mid;
        return zf < 0.05;
    };
    for (int j = 0; j < n; ++j) {
        const bool forced = (j > 0 && S.ptype[j] == 2);
        const bool fundamental = (j > 0 && parent[j - 1] == j && cnt[j - 1] == cnt[j] + 1 && nchild[j] == 1);
        if (j > 0 && (forced || fundamental) && !sn.empty() && sn.back().b == j) {
            SN& c = sn.back();
            if (!fundamental) {   // forced pair merge may introduce explicit zeros
                c.zeros += (double)(c.b - c.a) * (1.0 + cnt[j] - cnt[j - 1]);
            }
            c.b = j + 1;
        } else {
            sn.push_back({j, j + 1, 0.0});
        }
        // close the supernode only when the next column cannot join; try relaxed merges when it is complete
        const bool next_joins = (j + 1 < n) && (S.ptype[j + 1] == 2 ||
                                                 (parent[j] == j + 1 && cnt[j] == cnt[j + 1] + 1 && nchild[j + 1] == 1));
        if (!next_joins) {
            while (sn.size() >= 2) {
                const SN& p = sn.back();
                const SN& c = sn[sn.size() - 2];
                if (c.b != p.a) break;
                bool ok = merge_ok(c, p);
                if (!ok && opt.subtree_width > 0) {
                    // subtree amalgamation: a child whose WHOLE subtree already sits in one supernode is absorbed as
                    // long as the merged front stays narrow. Thousands of 4-8 column fronts at the bottom of the
                    // nested-dissection tree become a few hundred 16-32 column fronts: the device pays per front
                    // (dependency latency), not per flop, down there.
                    const int par = parent[c.b - 1];
                    const bool child = par >= p.a && par < p.b;
                    const bool whole = (c.b - c.a) == desc[c.b - 1];
                    ok = child && whole && (p.b - c.a) <= std::min(opt.subtree_width, opt.max_width);
                }
                if (!ok) break;
                const int wc = c.b - c.a, wp = p.b - p.a;
                const double z = (double)wc * (wp + (double)cnt[p.b - 1] - (double)cnt[c.b - 1]) + c.zeros + p.zeros;
                SN m{c.a, p.b, z};
                sn.pop_back(); sn.pop_back();
                sn.push_back(m);
            }
        }
    }
    // split wide supernodes into chains; never between the two columns of a 2x2 pivot
    std::vector<int> fs_a, fs_b, fs_e;   // final supernode [a,b), owner end e (rows = [b,e) + struct(e-1))
    for (const SN& s : sn) {
        int a = s.a;
        while (a < s.b) {
            int b = std::min(a + opt.max_width, s.b);
            if (b < s.b && S.ptype[b] == 2) --b;
            if (b <= a) b = std::min(a + 2, s.b);
            fs_a.push_back(a); fs_b.push_back(b); fs_e.push_back(s.b);
            a = b;
        }
    }
    const int ns = (int)fs_a.size();
    S.ns = ns;
    S.sn_col.resize(ns + 1);
    S.sn_rowptr.assign(ns + 1, 0);
    S.sn_na.resize(ns); S.sn_nb.resize(ns);
    std::vector<int> sn_of(n);
    for (int s = 0; s < ns; ++s) {
        S.sn_col[s] = fs_a[s];
        for (int j = fs_a[s]; j < fs_b[s]; ++j) sn_of[j] = s;
        const int e = fs_e[s];
        S.sn_rowptr[s + 1] = S.sn_rowptr[s] + (e - fs_b[s]) + cnt[e - 1];
    }
    S.sn_col[ns] = n;
    S.rowidx.resize(S.sn_rowptr[ns]);
    for (int s = 0; s < ns; ++s) {
        int64_t q = S.sn_rowptr[s];
        const int e = fs_e[s];
        for (int r = fs_b[s]; r < e; ++r) S.rowidx[q++] = r;
        for (int r : cs[e - 1]) S.rowidx[q++] = r;
        int na = 0;
        for (int64_t t = S.sn_rowptr[s]; t < S.sn_rowptr[s + 1]; ++t) if (S.rowidx[t] < n) ++na;
        S.sn_na[s] = na;
        S.sn_nb[s] = (int)(S.sn_rowptr[s + 1] - S.sn_rowptr[s]) - na;
    }
    { std::vector<std::vector<int>>().swap(cs); }

    // ---- 8. front tree, levels, relative indices, storage offsets ----------------------------------------
    S.sn_parent.assign(ns, -1);
    S.sn_level.assign(ns, 0);
    S.child_ptr.assign(ns + 1, 0);
    for (int s = 0; s < ns; ++s) {
        if (S.sn_na[s] > 0) {
            S.sn_parent[s] = sn_of[S.rowidx[S.sn_rowptr[s]]];
            ++S.child_ptr[S.sn_parent[s] + 1];
        }
    }
    for (int s = 0; s < ns; ++s) S.child_ptr[s + 1] += S.child_ptr[s];
    S.child_idx.resize(S.child_ptr[ns]);
    {
        std::vector<int> f(S.child_ptr.begin(), S.child_ptr.end() - 1);
        for (int s = 0; s < ns; ++s) if (S.sn_parent[s] >= 0) S.child_idx[f[S.sn_parent[s]]++] = s;
    }
    int maxlev = 0;
    for (int s = ns - 1; s >= 0; --s) {
        S.sn_level[s] = (S.sn_parent[s] < 0) ? 0 : S.sn_level[S.sn_parent[s]] + 1;
        maxlev = std::max(maxlev, S.sn_level[s]);
    }
    S.nlevels = ns > 0 ? maxlev + 1 : 0;

    S.relidx.assign(S.rowidx.size(), -1);
    S.Loff.assign(ns + 1, 0);
    S.Coff.assign(ns + 1, 0);
    for (int s = 0; s < ns; ++s) {
        const int w = S.sn_col[s + 1] - S.sn_col[s];
        const int u = S.sn_na[s] + S.sn_nb[s];
        const int m = w + u;
        S.max_w = std::max(S.max_w, w);
        S.max_m = std::max(S.max_m, m);
        S.Loff[s + 1] = S.Loff[s] + (int64_t)m * w;
        S.Coff[s + 1] = S.Coff[s] + (int64_t)u * S.sn_na[s];
        const int p = S.sn_parent[s];
        if (p < 0) continue;
        const int pa = S.sn_col[p], pb = S.sn_col[p + 1], wp = pb - pa;
        int64_t qp = S.sn_rowptr[p];
        const int64_t qpe = S.sn_rowptr[p + 1];
        for (int64_t t = S.sn_rowptr[s]; t < S.sn_rowptr[s + 1]; ++t) {
            const int r = S.rowidx[t];
            if (r < pb) {
                if (r < pa) throw std::runtime_error("analyze_leaf: child row below parent columns");
                S.relidx[t] = r - pa;
            } else {
                while (qp < qpe && S.rowidx[qp] < r) ++qp;
                if (qp >= qpe || S.rowidx[qp] != r) throw std::runtime_error("analyze_leaf: child row missing in parent front");
                S.relidx[t] = wp + (int)(qp - S.sn_rowptr[p]);
            }
        }
    }

    // ---- 9. assembly maps ---------------------------------------------------------------------------------
    auto local_row = [&](int s, int r) -> int {
        const int a = S.sn_col[s], b = S.sn_col[s + 1];
        if (r < b) return r - a;
        const int* beg = S.rowidx.data() + S.sn_rowptr[s];
        const int* end = S.rowidx.data() + S.sn_rowptr[s + 1];
        const int* it = std::lower_bound(beg, end, r);
        if (it == end || *it != r) throw std::runtime_error("analyze_leaf: entry outside the symbolic structure");
        return (b - a) + (int)(it - beg);
    };
    S.kmap.resize(nnzK);
    for (int i = 0; i < n; ++i)
        for (int p = krow[i]; p < krow[i + 1]; ++p) {
            const int pi = iperm[i], pj = iperm[jcol[p]];
            const int c = std::min(pi, pj), r = std::max(pi, pj);
            const int s = sn_of[c];
            const int m = (S.sn_col[s + 1] - S.sn_col[s]) + S.sn_na[s] + S.sn_nb[s];
            S.kmap[p] = S.Loff[s] + (int64_t)(c - S.sn_col[s]) * m + local_row(s, r);
        }
    S.bmap.resize(nnzB);
    for (int64_t k = 0; k < nnzB; ++k) {
        const int c = iperm[brow[k]];
        const int s = sn_of[c];
        const int m = (S.sn_col[s + 1] - S.sn_col[s]) + S.sn_na[s] + S.sn_nb[s];
        S.bmap[k] = S.Loff[s] + (int64_t)(c - S.sn_col[s]) * m + local_row(s, n + bcol[k]);
    }
}

void build_border_groups(const LeafSymbolic& S, int big_u, double frac, BorderGroups& out) {
    const int ns = S.ns;
    out = BorderGroups();
    out.grouped.assign((size_t)ns, 0);
    std::vector<int> grp_of((size_t)ns, -1);
    std::vector<std::vector<int>> pos((size_t)ns);
    std::vector<std::vector<int>> mem;
    // fronts are numbered in postorder: walking backwards visits parents before their children
    for (int f = ns - 1; f >= 0; --f) {
        const int na = S.sn_na[f], nb = S.sn_nb[f];
        if (na + nb < big_u || nb == 0) continue;
        const int p = S.sn_parent[f];
        if (p >= 0 && p <= f) throw std::runtime_error("build_border_groups: front tree is not in postorder");
        std::vector<int>& pf = pos[(size_t)f];
        pf.resize((size_t)nb);
        const int g = p >= 0 ? grp_of[(size_t)p] : -1;
        if (g >= 0 && (double)nb >= frac * out.N[(size_t)g]) {
            const int wp = S.sn_col[p + 1] - S.sn_col[p], nap = S.sn_na[p], nbp = S.sn_nb[p];
            const std::vector<int>& pp = pos[(size_t)p];
            const int* rel = S.relidx.data() + S.sn_rowptr[f] + na;
            for (int t = 0; t < nb; ++t) {
                const int li = rel[t] - wp - nap;
                if (li < 0 || li >= nbp) throw std::runtime_error("build_border_groups: border row of a front is not a border row of its parent");
                pf[(size_t)t] = pp[(size_t)li];
                if (t > 0 && pf[(size_t)t] <= pf[(size_t)t - 1]) throw std::runtime_error("build_border_groups: border rows are not ascending");
            }
            grp_of[(size_t)f] = g;
            mem[(size_t)g].push_back(f);
        } else {
            for (int t = 0; t < nb; ++t) pf[(size_t)t] = t;
            grp_of[(size_t)f] = (int)out.root.size();
            out.root.push_back(f);
            out.N.push_back(nb);
            mem.push_back(std::vector<int>{f});
        }
        out.grouped[(size_t)f] = 1;
    }
    out.mem_ptr.assign(1, 0);
    out.pos_ptr.assign(1, 0);
    for (size_t g = 0; g < mem.size(); ++g) {
        for (int f : mem[g]) {
            out.mem_front.push_back(f);
            out.pos.insert(out.pos.end(), pos[(size_t)f].begin(), pos[(size_t)f].end());
            out.pos_ptr.push_back((int64_t)out.pos.size());
        }
        out.mem_ptr.push_back((int)out.mem_front.size());
    }
}

}  // namespace pipsgpu
