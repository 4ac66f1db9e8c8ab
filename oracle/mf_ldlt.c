/* CPU ORACLE (test infrastructure only - the product never links or calls this file).
 *
 * Serial C restatement of the numeric phase of a multifrontal sparse LDL^T with static 1x1 / 2x2 pivots and
 * PARDISO-style pivot perturbation, including the partial factorisation that leaves the Schur complement
 * -Bord^T K^-1 Bord behind. This is the published algorithm of the third-party solvers the reference calls through
 * SOL/DoubleLinearSolver.h:24-72 (multifrontal LDL^T: Duff & Reid, MA27/MA57, Ma57Solver.C:84-106; incomplete
 * factorisation with Schur output: PARDISO iparm[37], PardisoProjectSchurSolver.C:246-280; pivot perturbation
 * iparm[9]=8, PardisoProjectSolver.C:68-77). None of those libraries is available in /root/reference
 * (ThirdPartyLibs/* holds download scripts only) -> "parity unpinned" at this boundary; the restatement is pinned
 * instead against SuperLU (oracle/linsys.py) and dense numpy in tests/test_oracle_golden.py (test_c_multifrontal_oracle_matches_superlu_oracle).
 *
 * The symbolic structure (permutation, fronts, maps) is an INPUT: tests obtain it from the product's host-side
 * analysis through pipsgpu_symbolic_export (include/pipsgpu.h), which is how that analysis is validated on CPU.
 *
 * Layout (identical to the device layout, DESIGN.md "data layout"):
 *   front s: columns [sn_col[s], sn_col[s+1]), w = width, rows = rowidx[sn_rowptr[s]..) ascending (border rows are
 *   n+b), na non-border + nb border rows, m = w+na+nb. Panel = L + Loff[s], m x w column-major (ld = m).
 *   Contribution block = C + Coff[s], (na+nb) x na column-major (ld = na+nb), lower trapezoid used.
 *   D^-1 is block diagonal: dinv_d[k] diagonal, dinv_e[k] couples k and k+1 when ptype[k]==1.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Optional BLAS-3 front updates (the CPU arm of bench.py must be a fair stand-in for PARDISO / MA57, which use BLAS-3
 * kernels inside their fronts): oracle/mf.py hands over the address of a CBLAS dgemm (scipy's bundled OpenBLAS, loaded at
 * run time; single-threaded per call - the blocks are distributed over host threads). Without it the scalar loops run. */
typedef void (*cblas_dgemm_fn)(int layout, int transa, int transb, int m, int n, int k, double alpha, const double* a, int lda,
                               const double* b, int ldb, double beta, double* c, int ldc);
static cblas_dgemm_fn g_dgemm = 0;
void mf_set_dgemm(void* fn) { g_dgemm = (cblas_dgemm_fn)fn; }
int mf_has_dgemm(void) { return g_dgemm != 0; }
enum { MF_COL_MAJOR = 102, MF_NO_TRANS = 111, MF_TRANS = 112, MF_JB = 192 };

typedef struct {
    int n, nB, ns;
    const int* sn_col;
    const int64_t* sn_rowptr;
    const int* rowidx;
    const int* relidx;
    const int* sn_na;
    const int* sn_nb;
    const int* sn_parent;
    const int64_t* Loff;
    const int64_t* Coff;
    const int8_t* ptype;
    const int8_t* sgn;
} mf_sym;

/* L: pool of size Loff[ns] (zeroed + scattered here); C: pool of size Coff[ns]; D arrays length n;
 * SCb: nB x nB row-major, LOWER triangle in border-index space, receives  -= Bord^T K^-1 Bord  (accumulated). */
int mf_factorize(const mf_sym* S, int64_t nnzK, const int64_t* kmap, const double* kval, int64_t nnzB,
                 const int64_t* bmap, const double* bval, double* L, double* C, double* d_d, double* d_e,
                 double* dinv_d, double* dinv_e, double* SCb, double eps, int* inertia, int* n_perturbed) {
    const int n = S->n, ns = S->ns;
    memset(L, 0, sizeof(double) * (size_t)S->Loff[ns]);
    memset(C, 0, sizeof(double) * (size_t)S->Coff[ns]);
    for (int64_t k = 0; k < nnzK; ++k) L[kmap[k]] += kval[k];
    for (int64_t k = 0; k < nnzB; ++k) L[bmap[k]] += bval[k];
    int pos = 0, neg = 0, zero = 0, npert = 0;
    int maxw = 0, maxm = 0;
    for (int s = 0; s < ns; ++s) {
        int w = S->sn_col[s + 1] - S->sn_col[s];
        int m = w + S->sn_na[s] + S->sn_nb[s];
        if (w > maxw) maxw = w;
        if (m > maxm) maxm = m;
    }
    double* W = (double*)malloc(sizeof(double) * (size_t)maxm * (size_t)(maxw > 0 ? maxw : 1));
    if (!W) return -1;
    double* T = 0;      /* BLAS path: one block column of the update, row-major T[i][j] */
    if (g_dgemm) {
        T = (double*)malloc(sizeof(double) * (size_t)(maxm > 0 ? maxm : 1) * MF_JB);
        if (!T) { free(W); return -1; }
    }

    for (int s = 0; s < ns; ++s) {
        const int c0 = S->sn_col[s];
        const int w = S->sn_col[s + 1] - c0;
        const int na = S->sn_na[s], nb = S->sn_nb[s];
        const int u = na + nb, m = w + u;
        double* F = L + S->Loff[s];
        /* ---- dense partial factorisation of the panel; W keeps the unscaled columns (W = L*D) ---- */
        int k = 0;
        while (k < w) {
            const int g = c0 + k;
            if (S->ptype[g] != 1) {
                double d = F[(size_t)k * m + k];
                const double sg = (double)S->sgn[g];
                /* inertia = sign of the computed pivot, before the static-pivoting perturbation (what PARDISO / MA57
                 * report through get_inertia, LS/LinearSystem.C:296-325 relies on it) */
                if (d >= eps) ++pos; else if (d <= -eps) ++neg; else ++zero;
                if (!(sg * d >= eps)) { d = sg * eps; ++npert; }
                d_d[g] = d; d_e[g] = 0.0; dinv_d[g] = 1.0 / d; dinv_e[g] = 0.0;
                F[(size_t)k * m + k] = 1.0;
                for (int i = k + 1; i < m; ++i) {
                    const double f = F[(size_t)k * m + i];
                    W[(size_t)k * m + i] = f;
                    F[(size_t)k * m + i] = f / d;
                }
                for (int j = k + 1; j < w; ++j) {
                    const double wj = W[(size_t)k * m + j];
                    if (wj == 0.0) continue;
                    for (int i = j; i < m; ++i) F[(size_t)j * m + i] -= F[(size_t)k * m + i] * wj;
                }
                k += 1;
            } else {
                double a = F[(size_t)k * m + k], b = F[(size_t)k * m + k + 1], c = F[(size_t)(k + 1) * m + k + 1];
                double det = a * c - b * b;
                if (det <= -eps * eps) { ++pos; ++neg; }
                else if (det >= eps * eps) { if (a + c > 0) pos += 2; else neg += 2; }
                else { ++zero; if (a + c > 0) ++pos; else ++neg; }
                if (!(det <= -eps * eps)) {   /* expected: one positive, one negative eigenvalue */
                    if (a < eps) a = eps;
                    if (c > -eps) c = -eps;
                    det = a * c - b * b;
                    ++npert;
                }
                const double ia = c / det, ib = -b / det, ic = a / det;
                d_d[g] = a; d_e[g] = b; d_d[g + 1] = c; d_e[g + 1] = 0.0;
                dinv_d[g] = ia; dinv_e[g] = ib; dinv_d[g + 1] = ic; dinv_e[g + 1] = 0.0;
                F[(size_t)k * m + k] = 1.0; F[(size_t)k * m + k + 1] = 0.0; F[(size_t)(k + 1) * m + k + 1] = 1.0;
                for (int i = k + 2; i < m; ++i) {
                    const double f1 = F[(size_t)k * m + i], f2 = F[(size_t)(k + 1) * m + i];
                    W[(size_t)k * m + i] = f1; W[(size_t)(k + 1) * m + i] = f2;
                    F[(size_t)k * m + i] = f1 * ia + f2 * ib;
                    F[(size_t)(k + 1) * m + i] = f1 * ib + f2 * ic;
                }
                for (int j = k + 2; j < w; ++j) {
                    const double w1 = W[(size_t)k * m + j], w2 = W[(size_t)(k + 1) * m + j];
                    if (w1 == 0.0 && w2 == 0.0) continue;
                    for (int i = j; i < m; ++i)
                        F[(size_t)j * m + i] -= F[(size_t)k * m + i] * w1 + F[(size_t)(k + 1) * m + i] * w2;
                }
                k += 2;
            }
        }
        /* ---- contribution block: C(i,j) -= sum_k L(i,k) W(j,k), j < na ; border x border goes to SCb ---- */
        double* Cs = C + S->Coff[s];
        const int* rows = S->rowidx + S->sn_rowptr[s];
        if (g_dgemm && u >= 48) {
            /* U = L21 W21^T by block columns of MF_JB: T^T (jb x ni, column-major) = W21(j-block) * L21(i >= j0)^T, i.e.
             * T[i][j] row-major with leading dimension jb - rows of the Schur complement are then updated contiguously */
            const double* L21 = F + w;       /* u x w, ld m */
            const double* W21 = W + w;       /* u x w, ld m */
            for (int j0 = 0; j0 < u; j0 += MF_JB) {
                const int jb = (u - j0 < MF_JB) ? u - j0 : MF_JB;
                const int ni = u - j0;
                g_dgemm(MF_COL_MAJOR, MF_NO_TRANS, MF_TRANS, jb, ni, w, 1.0, W21 + j0, m, L21 + j0, m, 0.0, T, jb);
                for (int ii = 0; ii < ni; ++ii) {
                    const int i = j0 + ii;
                    const double* Ti = T + (size_t)ii * jb;
                    const int jend = (ii < jb) ? ii + 1 : jb;      /* j <= i */
                    double* scrow = (SCb && i >= na) ? SCb + (size_t)(rows[i] - n) * S->nB : 0;
                    for (int jj = 0; jj < jend; ++jj) {
                        const int j = j0 + jj;
                        if (j < na) Cs[(size_t)j * u + i] -= Ti[jj];
                        else if (scrow) scrow[rows[j] - n] -= Ti[jj];
                    }
                }
            }
        } else
        for (int j = 0; j < u; ++j) {
            for (int kk = 0; kk < w; ++kk) {
                const double wj = W[(size_t)kk * m + w + j];
                if (wj == 0.0) continue;
                const double* Lk = F + (size_t)kk * m + w;
                if (j < na) {
                    double* Cj = Cs + (size_t)j * u;
                    for (int i = j; i < u; ++i) Cj[i] -= Lk[i] * wj;
                } else if (SCb) {
                    const int bj = rows[j] - n;
                    for (int i = j; i < u; ++i) SCb[(size_t)(rows[i] - n) * S->nB + bj] -= Lk[i] * wj;
                }
            }
        }
        /* ---- extend-add into the parent front ---- */
        const int p = S->sn_parent[s];
        if (p >= 0) {
            const int wp = S->sn_col[p + 1] - S->sn_col[p];
            const int up = S->sn_na[p] + S->sn_nb[p], mp = wp + up;
            double* Fp = L + S->Loff[p];
            double* Cp = C + S->Coff[p];
            const int* rel = S->relidx + S->sn_rowptr[s];
            for (int j = 0; j < na; ++j) {
                const int lj = rel[j];
                for (int i = j; i < u; ++i) {
                    const int li = rel[i];
                    const double v = Cs[(size_t)j * u + i];
                    if (lj < wp) Fp[(size_t)lj * mp + li] += v;
                    else Cp[(size_t)(lj - wp) * up + (li - wp)] += v;
                }
            }
        }
    }
    free(W);
    free(T);
    if (inertia) { inertia[0] = pos; inertia[1] = neg; inertia[2] = zero; }
    if (n_perturbed) *n_perturbed = npert;
    return 0;
}

/* x (length n, original ordering) <- K^-1 x using the factors */
int mf_solve(const mf_sym* S, const int* perm, const double* L, const double* dinv_d, const double* dinv_e, double* x) {
    const int n = S->n, ns = S->ns;
    double* y = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    if (!y) return -1;
    for (int p = 0; p < n; ++p) y[p] = x[perm[p]];
    for (int s = 0; s < ns; ++s) {
        const int c0 = S->sn_col[s];
        const int w = S->sn_col[s + 1] - c0;
        const int na = S->sn_na[s], m = w + na + S->sn_nb[s];
        const double* F = L + S->Loff[s];
        const int* rows = S->rowidx + S->sn_rowptr[s];
        for (int k = 0; k < w; ++k) {
            const double yk = y[c0 + k];
            if (yk == 0.0) continue;
            for (int i = k + 1; i < w; ++i) y[c0 + i] -= F[(size_t)k * m + i] * yk;
            for (int i = 0; i < na; ++i) y[rows[i]] -= F[(size_t)k * m + w + i] * yk;
        }
    }
    {
        double prev = 0.0;
        for (int k = 0; k < n; ++k) {
            const double yk = y[k];
            const double nxt = (k + 1 < n) ? y[k + 1] : 0.0;
            double v = dinv_d[k] * yk + dinv_e[k] * nxt;
            if (k > 0) v += dinv_e[k - 1] * prev;
            prev = yk;
            y[k] = v;
        }
    }
    for (int s = ns - 1; s >= 0; --s) {
        const int c0 = S->sn_col[s];
        const int w = S->sn_col[s + 1] - c0;
        const int na = S->sn_na[s], m = w + na + S->sn_nb[s];
        const double* F = L + S->Loff[s];
        const int* rows = S->rowidx + S->sn_rowptr[s];
        for (int k = w - 1; k >= 0; --k) {
            double acc = y[c0 + k];
            for (int i = k + 1; i < w; ++i) acc -= F[(size_t)k * m + i] * y[c0 + i];
            for (int i = 0; i < na; ++i) acc -= F[(size_t)k * m + w + i] * y[rows[i]];
            y[c0 + k] = acc;
        }
    }
    for (int p = 0; p < n; ++p) x[perm[p]] = y[p];
    free(y);
    return 0;
}
