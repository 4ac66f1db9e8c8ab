// Drives the B200 path exactly the way PIPS-IPM++ would: through DoubleLinearSolver* virtual calls on the shim classes
// of include/pipsgpu_shim.hpp (standalone stand-ins for the reference types). Checks size-independent properties:
// residual of the plugin-level solves, and the residual of solveCompressed against the full doubly bordered KKT
// matvec (LS/LinearSystem.C:809-844 system_mult, compressed ordering). Exit code 0 = pass, 3 = library error
// (e.g. no CUDA device: the product path has no CPU fallback), 1 = numerical failure.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define PIPSGPU_SHIM_STANDALONE
#define PIPSGPU_SHIM_ABORT(msg)                                                  \
    do {                                                                         \
        std::fprintf(stderr, "pipsgpu: %s: %s\n", (msg), pipsgpu_last_error()); \
        std::exit(3);                                                            \
    } while (0)
#include "pipsgpu_shim.hpp"

using namespace pipsgpu_shim;

static uint64_t g_state = 0x5EED1234ULL;
static double urand() {   // splitmix64
    uint64_t z = (g_state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return (double)(z >> 11) / 9007199254740992.0;
}

struct Csr {
    int m = 0, n = 0;
    std::vector<int> krow{0}, jcol;
    std::vector<double> val;
    SparseMatrix sparse() const { return SparseMatrix(m, n, krow, jcol, val); }
    void mult(double alpha, const double* x, double* y) const {          // y += alpha A x
        for (int r = 0; r < m; ++r)
            for (int p = krow[r]; p < krow[r + 1]; ++p) y[r] += alpha * val[p] * x[jcol[p]];
    }
    void tmult(double alpha, const double* x, double* y) const {         // y += alpha A^T x
        for (int r = 0; r < m; ++r)
            for (int p = krow[r]; p < krow[r + 1]; ++p) y[jcol[p]] += alpha * val[p] * x[r];
    }
};
static Csr random_rows(int m, int n, int per_row, bool band) {
    Csr a; a.m = m; a.n = n; a.krow.assign(1, 0);
    for (int r = 0; r < m; ++r) {
        std::vector<int> cols;
        if (band) for (int d = -1; d <= 1; ++d) { int c = (int)((long long)r * n / m) + d; if (c >= 0 && c < n) cols.push_back(c); }
        while ((int)cols.size() < per_row && n > 0) {
            int c = (int)(urand() * n); bool dup = false;
            for (int q : cols) dup |= (q == c);
            if (!dup) cols.push_back(c);
        }
        std::sort(cols.begin(), cols.end());
        for (int c : cols) { a.jcol.push_back(c); a.val.push_back(0.5 + urand()); }
        a.krow.push_back((int)a.jcol.size());
    }
    return a;
}

struct Leaf {
    int nx, ny, nz;
    Csr W, D, A, C, F, G;
    std::vector<double> dd, nom;
    std::unique_ptr<SparseSymmetricMatrix> K;
    std::vector<int> diagpos;
};

// create_kkt (LS/DistributedLeafLinearSystem.C:52-72): lower CSR [diag; W 0; D 0 diag], diagonal entry last in each row
static void create_kkt(Leaf& lf) {
    const int n = lf.nx + lf.ny + lf.nz;
    std::vector<int> krow{0}, jcol; std::vector<double> val;
    lf.diagpos.assign(n, 0);
    for (int i = 0; i < n; ++i) {
        const Csr* B = nullptr; int r = 0;
        if (i >= lf.nx && i < lf.nx + lf.ny) { B = &lf.W; r = i - lf.nx; }
        if (i >= lf.nx + lf.ny) { B = &lf.D; r = i - lf.nx - lf.ny; }
        if (B) for (int p = B->krow[r]; p < B->krow[r + 1]; ++p) { jcol.push_back(B->jcol[p]); val.push_back(B->val[p]); }
        lf.diagpos[i] = (int)jcol.size();
        jcol.push_back(i); val.push_back(0.0);
        krow.push_back((int)jcol.size());
    }
    lf.K = std::make_unique<SparseSymmetricMatrix>(n, krow, jcol, val);
}
static void put_diagonals(Leaf& lf) {   // :88-143
    double* M = lf.K->M();
    for (int i = 0; i < lf.nx; ++i) M[lf.diagpos[i]] = lf.dd[i];
    for (int i = 0; i < lf.ny; ++i) M[lf.diagpos[lf.nx + i]] = 0.0;
    for (int i = 0; i < lf.nz; ++i) M[lf.diagpos[lf.nx + lf.ny + i]] = lf.nom[i];
}
static void kkt_mult(const Leaf& lf, const double* x, double* y) {   // y = K x
    const int nx = lf.nx, ny = lf.ny;
    for (int i = 0; i < nx; ++i) y[i] = lf.dd[i] * x[i];
    for (int i = 0; i < ny; ++i) y[nx + i] = 0.0;
    for (int i = 0; i < lf.nz; ++i) y[nx + ny + i] = lf.nom[i] * x[nx + ny + i];
    lf.W.tmult(1.0, x + nx, y); lf.D.tmult(1.0, x + nx + ny, y);
    lf.W.mult(1.0, x, y + nx); lf.D.mult(1.0, x, y + nx + ny);
}
static double norm2(const std::vector<double>& v) { double s = 0; for (double x : v) s += x * x; return std::sqrt(s); }

int main() {
    const int N = 3, n0 = 6, my0 = 2, mz0 = 3, myl = 3, mzl = 2;
    const int nS = n0 + my0 + myl + mzl, nb0 = n0 + my0 + mz0 + myl + mzl;
    std::vector<Leaf> leaves(N);
    for (int i = 0; i < N; ++i) {
        Leaf& lf = leaves[i];
        lf.nx = 60 + 7 * i; lf.ny = 20 + i; lf.nz = 30;
        lf.W = random_rows(lf.ny, lf.nx, 4, true); lf.D = random_rows(lf.nz, lf.nx, 4, true);
        lf.A = random_rows(lf.ny, n0, 1, false); lf.C = random_rows(lf.nz, n0, 1, false);
        lf.F = random_rows(myl, lf.nx, 5, false); lf.G = random_rows(mzl, lf.nx, 5, false);
        lf.dd.resize(lf.nx); lf.nom.resize(lf.nz);
        for (auto& v : lf.dd) v = std::pow(10.0, -2.0 + 4.0 * urand());
        for (auto& v : lf.nom) v = -std::pow(10.0, -2.0 + 4.0 * urand());
        create_kkt(lf); put_diagonals(lf);
    }
    Csr A0 = random_rows(my0, n0, 3, false), C0 = random_rows(mz0, n0, 3, false), F0 = random_rows(myl, n0, 2, false),
        G0 = random_rows(mzl, n0, 2, false);
    std::vector<double> dd0(n0), nom0(mz0), noml(mzl);
    for (auto& v : dd0) v = 0.5 + urand();
    for (auto& v : nom0) v = -(0.5 + urand());
    for (auto& v : noml) v = -(0.5 + urand());

    // ---- system level: sLinsysRootAug-style use ------------------------------------------------------------
    std::vector<std::unique_ptr<CudaSchurLeafSolver>> solvers;
    CudaKktSystem kkt(n0, my0, mz0, myl, mzl, A0.sparse(), C0.sparse(), F0.sparse(), G0.sparse());
    Csr R;   // empty cross Hessian
    for (auto& lf : leaves) {
        solvers.push_back(std::make_unique<CudaSchurLeafSolver>(*lf.K, lf.nx, lf.ny, lf.nz));
        solvers.back()->set_border(R.sparse(), lf.A.sparse(), lf.C.sparse(), lf.F.sparse(), lf.G.sparse(), n0, my0);
        kkt.add_leaf(*solvers.back());
    }
    kkt.setup();
    std::vector<double> ddl, noml_leaves;
    for (auto& lf : leaves) { ddl.insert(ddl.end(), lf.dd.begin(), lf.dd.end()); noml_leaves.insert(noml_leaves.end(), lf.nom.begin(), lf.nom.end()); }
    auto inertia = kkt.factorize(dd0.data(), nom0.data(), noml.data(), ddl.data(), noml_leaves.data());
    if (inertia != std::make_tuple((unsigned)n0, (unsigned)(my0 + myl + mzl), 0u)) { std::fprintf(stderr, "root inertia wrong\n"); return 1; }
    int sumn = 0; for (auto& lf : leaves) sumn += lf.nx + lf.ny + lf.nz;
    std::vector<double> b0(nb0), bl(sumn);
    for (auto& v : b0) v = urand() - 0.5;
    for (auto& v : bl) v = urand() - 0.5;
    std::vector<double> x0 = b0, xl = bl;
    kkt.solveCompressed(x0.data(), xl.data());
    // residual of the full bordered system, b0 layout [x0; y0; z0; yl; zl]
    std::vector<double> r0(nb0, 0.0), rl(sumn, 0.0);
    const double* X0 = x0.data(); const double *Y0 = X0 + n0, *Z0 = Y0 + my0, *YL = Z0 + mz0, *ZL = YL + myl;
    for (int i = 0; i < n0; ++i) r0[i] = dd0[i] * X0[i];
    A0.tmult(1, Y0, r0.data()); C0.tmult(1, Z0, r0.data()); F0.tmult(1, YL, r0.data()); G0.tmult(1, ZL, r0.data());
    A0.mult(1, X0, r0.data() + n0);
    C0.mult(1, X0, r0.data() + n0 + my0); for (int i = 0; i < mz0; ++i) r0[n0 + my0 + i] += nom0[i] * Z0[i];
    F0.mult(1, X0, r0.data() + n0 + my0 + mz0);
    G0.mult(1, X0, r0.data() + n0 + my0 + mz0 + myl); for (int i = 0; i < mzl; ++i) r0[n0 + my0 + mz0 + myl + i] += noml[i] * ZL[i];
    int off = 0;
    for (auto& lf : leaves) {
        const double* xi = xl.data() + off; double* ri = rl.data() + off;
        kkt_mult(lf, xi, ri);
        lf.F.tmult(1, YL, ri); lf.G.tmult(1, ZL, ri);                      // Bord_i u
        lf.A.mult(1, X0, ri + lf.nx); lf.C.mult(1, X0, ri + lf.nx + lf.ny);
        lf.A.tmult(1, xi + lf.nx, r0.data()); lf.C.tmult(1, xi + lf.nx + lf.ny, r0.data());   // Bord_i^T x_i
        lf.F.mult(1, xi, r0.data() + n0 + my0 + mz0); lf.G.mult(1, xi, r0.data() + n0 + my0 + mz0 + myl);
        off += lf.nx + lf.ny + lf.nz;
    }
    double num = 0, den = 0;
    for (int i = 0; i < nb0; ++i) { num += (r0[i] - b0[i]) * (r0[i] - b0[i]); den += b0[i] * b0[i]; }
    for (int i = 0; i < sumn; ++i) { num += (rl[i] - bl[i]) * (rl[i] - bl[i]); den += bl[i] * bl[i]; }
    const double res_kkt = std::sqrt(num / den);
    std::printf("solveCompressed residual %.3e (n_S = %d)\n", res_kkt, nS);
    // LinearSystem::solveCompressedBiCGStab on the device (OUTER_SOLVE = 2): same right-hand side, the reported relative
    // residual is that of the full KKT system (device system_mult) and must agree with the host evaluation above
    std::vector<double> k0 = b0, kl = bl;
    int bicg_iters = -1; double bicg_rel = -1.0;
    const int bicg_status = kkt.solve(k0.data(), kl.data(), 1e-10, 75, &bicg_iters, &bicg_rel);
    double dsol = 0, nsol = 0;
    for (int i = 0; i < nb0; ++i) { dsol += (k0[i] - x0[i]) * (k0[i] - x0[i]); nsol += x0[i] * x0[i]; }
    for (int i = 0; i < sumn; ++i) { dsol += (kl[i] - xl[i]) * (kl[i] - xl[i]); nsol += xl[i] * xl[i]; }
    std::printf("device BiCGStab status %d, %d iterations, relative residual %.3e, distance to solveCompressed %.3e\n", bicg_status,
                bicg_iters, bicg_rel, std::sqrt(dsol / nsol));
    const bool ok_bicg = (bicg_status == 0 || bicg_status == 1) && bicg_rel <= 1e-10 && std::sqrt(dsol / nsol) <= 1e-7;

    // ---- plugin level: DoubleLinearSolver* exactly as DistributedLeafLinearSystem holds it ----------------------
    DoubleLinearSolver* solver = solvers[1].get();
    const Leaf& lf = leaves[1];
    const int n = lf.nx + lf.ny + lf.nz;
    solvers[1]->schur_solve(kkt.schur_complement());                      // addTermToDenseSchurCompl path (refactorises)
    DenseVector<double> rhs(n);
    std::vector<double> b(n), Kx(n);
    for (int i = 0; i < n; ++i) rhs.elements()[i] = b[i] = urand() - 0.5;
    solver->solve(rhs);
    kkt_mult(lf, rhs.elements(), Kx.data());
    double e = 0; for (int i = 0; i < n; ++i) e += (Kx[i] - b[i]) * (Kx[i] - b[i]);
    const double res_leaf = std::sqrt(e) / norm2(b);
    auto in = solver->get_inertia();
    std::printf("leaf solve residual %.3e, inertia (%u,%u,%u)\n", res_leaf, std::get<0>(in), std::get<1>(in), std::get<2>(in));
    std::vector<double> multi(3 * n);
    for (int k = 0; k < 3; ++k) for (int i = 0; i < n; ++i) multi[k * n + i] = (k + 1) * b[i];
    solver->solve(3, multi.data(), nullptr);
    double lin = 0; for (int i = 0; i < n; ++i) lin += std::fabs(multi[2 * n + i] - 3.0 * rhs.elements()[i]);
    // solve(GeneralMatrix&): two right-hand sides handed over as a dense block (PardisoSolver.C:230-251)
    DenseMatrix blockrhs(n, 2);
    for (int i = 0; i < n; ++i) { (&blockrhs[0][0])[i] = b[i]; (&blockrhs[0][0])[n + i] = -2.0 * b[i]; }
    solver->solve(blockrhs);
    for (int i = 0; i < n; ++i)
        lin += std::fabs((&blockrhs[0][0])[i] - rhs.elements()[i]) + std::fabs((&blockrhs[0][0])[n + i] + 2.0 * rhs.elements()[i]);

    // ---- PardisoSchurSolver's exact call sequence: matrixChanged(); schur_solve(R, A, C, F, G, DenseSymmetricMatrix&) ----
    // (sLinsysLeafSchurSlv.C:20-35 unchanged). Sum over the leaves against the batched device path on a fresh SC.
    DenseSymmetricMatrix SCh(nS);
    {
        std::vector<std::unique_ptr<CudaSchurLeafSolverCompat>> cs;
        for (auto& l : leaves) cs.push_back(std::make_unique<CudaSchurLeafSolverCompat>(*l.K, l.nx, l.ny, l.nz));
        for (size_t i = 0; i < leaves.size(); ++i) {
            DoubleLinearSolver* base = cs[i].get();
            base->matrixChanged();
            cs[i]->schur_solve(R.sparse(), leaves[i].A.sparse(), leaves[i].C.sparse(), leaves[i].F.sparse(), leaves[i].G.sparse(), SCh);
        }
    }
    pipsgpu_sc* sc2 = nullptr;
    PIPSGPU_SHIM_CHECK(pipsgpu_sc_create(CudaContext::instance().get(), n0, my0, myl, mzl, &sc2));
    PIPSGPU_SHIM_CHECK(pipsgpu_ctx_factorize_leaves(CudaContext::instance().get(), sc2));
    std::vector<double> sc_dev((size_t)nS * nS);
    PIPSGPU_SHIM_CHECK(pipsgpu_sc_download(sc2, sc_dev.data()));
    pipsgpu_sc_destroy(sc2);
    double sc_num = 0, sc_den = 0;
    for (int i = 0; i < nS; ++i)
        for (int j = 0; j <= i; ++j) {
            const double a = SCh[i][j], b2 = sc_dev[(size_t)i * nS + j];
            sc_num += (a - b2) * (a - b2); sc_den += b2 * b2;
        }
    const double res_compat = std::sqrt(sc_num / std::max(sc_den, 1e-300));
    std::printf("compat schur_solve(R,A,C,F,G,SC) vs batched Schur complement: rel. Frobenius %.3e\n", res_compat);

    // ---- dense root solver through DoubleLinearSolver* ----------------------------------------------------------
    const int nd = 48, nd0 = 20;
    DenseSymmetricMatrix S(nd);
    std::vector<double> full(nd * nd, 0.0);
    for (int i = 0; i < nd; ++i)
        for (int j = 0; j <= i; ++j) {
            double v = (i < nd0 || j >= nd0) ? 0.1 * (urand() - 0.5) : urand() - 0.5;
            if (i == j) v = (i < nd0) ? 3.0 + urand() : -(3.0 + urand());
            S[i][j] = v; full[i * nd + j] = full[j * nd + i] = v;
        }
    CudaDenseLdlSolver dense(S, nd0);
    DoubleLinearSolver* ds = &dense;
    ds->matrixChanged();
    DenseVector<double> dr(nd); std::vector<double> db(nd);
    for (int i = 0; i < nd; ++i) dr.elements()[i] = db[i] = urand() - 0.5;
    ds->solveSynchronized(dr);
    double de = 0;
    for (int i = 0; i < nd; ++i) { double s = -db[i]; for (int j = 0; j < nd; ++j) s += full[i * nd + j] * dr.elements()[j]; de += s * s; }
    const double res_dense = std::sqrt(de) / norm2(db);
    auto din = ds->get_inertia();
    std::printf("dense solve residual %.3e, inertia (%u,%u,%u)\n", res_dense, std::get<0>(din), std::get<1>(din), std::get<2>(din));

    const bool ok = ok_bicg && res_compat <= 1e-10 && sc_den > 0 && res_kkt <= 1e-9 && res_leaf <= 1e-10 && lin <= 1e-9 && res_dense <= 1e-11 &&
                    in == std::make_tuple((unsigned)lf.nx, (unsigned)(lf.ny + lf.nz), 0u) &&
                    din == std::make_tuple((unsigned)nd0, (unsigned)(nd - nd0), 0u);
    std::printf(ok ? "shim_test PASS\n" : "shim_test FAIL\n");
    return ok ? 0 : 1;
}
