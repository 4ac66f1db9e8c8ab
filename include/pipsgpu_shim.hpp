// pipsgpu_shim.hpp - the thin C++17 subclasses a PIPS-IPM++ maintainer compiles into the reference so that its
// linear-system layer calls the B200 path through the C ABI (include/pipsgpu.h). Header-only; no CUDA headers.
//
// Two granularities, both behind the reference's own virtual interfaces (INTEGRATION.md shows the factory edits):
//
//   plugin level     CudaSparseLdlSolver   : DoubleLinearSolver   (what PardisoSolver / Ma27Solver / Ma57Solver are:
//                                                                  SOL/PardisoSolver/PardisoSolver/PardisoSolver.h:20-40)
//                    CudaSchurLeafSolver   : CudaSparseLdlSolver  (adds schur_solve, PardisoSchurSolver.h:44-58)
//                    CudaSchurLeafSolverCompat : DoubleLinearSolver (schur_solve with the reference's exact signature and
//                                                                  call order; private context per leaf, slow path)
//                    CudaDenseLdlSolver    : DoubleLinearSolver   (DeSymIndefSolver.h:19-35, the dense root solver)
//   system level     CudaKktSystem         - owns a pipsgpu_kkt; a subclass of sLinsysRootAug forwards factor2 /
//                                            solveCompressed to it (batched over all scenario blocks of the rank:
//                                            the fast path, LS/DistributedRootLinearSystem.C:206-243,
//                                            LS/DistributedLinearSystem.C:409-420)
//
// Inside the reference tree this header is included AFTER the reference's own headers (DoubleLinearSolver.h,
// SparseSymmetricMatrix.h, DenseSymmetricMatrix.h, DenseVector.hpp). For the repository's own tests it is compiled
// with -DPIPSGPU_SHIM_STANDALONE, which supplies minimal stand-ins of exactly the members the shim touches
// (same names and meaning as the reference classes; nothing else of them is restated).
//
// Error convention: the reference's plugins print and abort on a library error (PardisoSolver.C:201-204); the shim
// does the same through PIPSGPU_SHIM_ABORT (default: message to stderr + std::abort(); define it to MPI_Abort).
#ifndef PIPSGPU_SHIM_HPP
#define PIPSGPU_SHIM_HPP

#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <tuple>
#include <utility>
#include <vector>

#include "pipsgpu.h"

#ifndef PIPSGPU_SHIM_ABORT
#define PIPSGPU_SHIM_ABORT(msg)                                             \
    do {                                                                    \
        std::fprintf(stderr, "pipsgpu: %s: %s\n", (msg), pipsgpu_last_error()); \
        std::abort();                                                       \
    } while (0)
#endif
#define PIPSGPU_SHIM_CHECK(call)                 \
    do {                                         \
        if ((call) != PIPSGPU_OK) PIPSGPU_SHIM_ABORT(#call); \
    } while (0)

#ifdef PIPSGPU_SHIM_STANDALONE
// ---- stand-ins for the reference types (only the members the shim uses) -----------------------------------------
template <typename T>
class Vector {   // LinearAlgebra/Abstract/Vector.hpp
public:
    virtual ~Vector() = default;
    [[nodiscard]] int length() const { return n; }
protected:
    explicit Vector(int n_) : n(n_) {}
    int n;
};
template <typename T>
class DenseVector : public Vector<T> {   // LinearAlgebra/Dense/DenseVector.hpp:20,146
public:
    explicit DenseVector(int n_) : Vector<T>(n_), v(new T[n_ > 0 ? n_ : 1]()) {}
    ~DenseVector() override { delete[] v; }
    T* elements() const { return v; }
private:
    T* v;
};
class AbstractMatrix { public: virtual ~AbstractMatrix() = default; };
class GeneralMatrix : public AbstractMatrix {};
class SparseMatrix : public GeneralMatrix {   // LinearAlgebra/Sparse/SparseMatrix.h:52-54,125-132 (general CSR)
public:
    SparseMatrix(int m, int n, std::vector<int> krow, std::vector<int> jcol, std::vector<double> val)
        : m_(m), n_(n), krow_(std::move(krow)), jcol_(std::move(jcol)), val_(std::move(val)) {}
    [[nodiscard]] long long n_rows() const { return m_; }
    [[nodiscard]] long long n_columns() const { return n_; }
    [[nodiscard]] const int* krowM() const { return krow_.data(); }
    [[nodiscard]] const int* jcolM() const { return jcol_.data(); }
    [[nodiscard]] const double* M() const { return val_.data(); }
private:
    int m_, n_;
    std::vector<int> krow_, jcol_;
    std::vector<double> val_;
};
class SparseSymmetricMatrix : public AbstractMatrix {   // LinearAlgebra/Sparse/SparseSymmetricMatrix.h:18-41,61 (lower CSR)
public:
    SparseSymmetricMatrix(int size, std::vector<int> krow, std::vector<int> jcol, std::vector<double> val)
        : n_(size), krow_(std::move(krow)), jcol_(std::move(jcol)), val_(std::move(val)) {}
    [[nodiscard]] long long size() const { return n_; }
    [[nodiscard]] const int* krowM() const { return krow_.data(); }
    [[nodiscard]] const int* jcolM() const { return jcol_.data(); }
    [[nodiscard]] const double* M() const { return val_.data(); }
    double* M() { return val_.data(); }
private:
    int n_;
    std::vector<int> krow_, jcol_;
    std::vector<double> val_;
};
class DenseMatrix : public GeneralMatrix {   // LinearAlgebra/Dense/DenseMatrix.h (row-major m x n block, operator[] = row pointer)
public:
    DenseMatrix(int m, int n) : m_(m), n_(n), a_((size_t)m * n, 0.0) {}
    [[nodiscard]] std::pair<long long, long long> n_rows_columns() const { return {m_, n_}; }
    double* operator[](int i) { return a_.data() + (size_t)i * n_; }
private:
    int m_, n_;
    std::vector<double> a_;
};
class DenseSymmetricMatrix : public AbstractMatrix {   // LinearAlgebra/Dense/DenseSymmetricMatrix.h:71-79 (row-major, lower used)
public:
    explicit DenseSymmetricMatrix(int size) : n_(size), a_((size_t)size * size, 0.0) {}
    [[nodiscard]] long long size() const { return n_; }
    double* elements() { return a_.data(); }
    [[nodiscard]] const double* elements() const { return a_.data(); }
    double* operator[](int i) { return a_.data() + (size_t)i * n_; }
private:
    int n_;
    std::vector<double> a_;
};
class DoubleLinearSolver {   // LinearSolvers/DoubleLinearSolver.h:24-72 (the virtual set, verbatim names)
public:
    virtual void diagonalChanged(int idiag, int extent) = 0;
    virtual void matrixChanged() = 0;
    virtual void matrixRebuild(const AbstractMatrix&) { assert(0 && "Not implemented"); }
    virtual void solve(Vector<double>& x) = 0;
    [[nodiscard]] virtual bool reports_inertia() const = 0;
    [[nodiscard]] virtual std::tuple<unsigned int, unsigned int, unsigned int> get_inertia() const = 0;
    virtual void solveSynchronized(Vector<double>& x) { solve(x); }
    virtual void solve(GeneralMatrix&) { assert(0 && "Not implemented"); }
    virtual void solve(int, double*, int*) { assert(0 && "Not implemented"); }
    void Dsolve(Vector<double>& x) { solve(x); }
    virtual ~DoubleLinearSolver() = default;
protected:
    DoubleLinearSolver() = default;
};
#endif  // PIPSGPU_SHIM_STANDALONE

namespace pipsgpu_shim {

// One context per process (= per MPI rank = per GPU), shared by every solver of the rank; created on first use.
// device = local rank; the NCCL communicator is attached by the root linear system (CudaKktSystem::attach_comm).
class CudaContext {
public:
    static CudaContext& instance(int device = -1) {
        static CudaContext ctx(device < 0 ? device_from_env() : device);
        return ctx;
    }
    pipsgpu_ctx* get() const { return ctx_; }
    // leaves register at construction; the pools are built when the first factorisation is requested, i.e. after
    // the reference has constructed its whole linear-system tree (LS/DistributedRootLinearSystem.C:711-758)
    void commit_once(int threads = 0) {
        if (!committed_) {
            PIPSGPU_SHIM_CHECK(pipsgpu_ctx_commit(ctx_, threads));
            committed_ = true;
        }
    }
    bool committed() const { return committed_; }
    void mark_committed() { committed_ = true; }
    CudaContext(const CudaContext&) = delete;
    CudaContext& operator=(const CudaContext&) = delete;

private:
    explicit CudaContext(int device) { PIPSGPU_SHIM_CHECK(pipsgpu_ctx_create(device, &ctx_)); }
    ~CudaContext() { pipsgpu_ctx_destroy(ctx_); }
    static int device_from_env() {
        for (const char* k : {"LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID"})
            if (const char* v = std::getenv(k)) return std::atoi(v);
        return 0;
    }
    pipsgpu_ctx* ctx_ = nullptr;
    bool committed_ = false;
};

// ---------------------------------------------------------------------------------------------------------------
// sparse leaf solver: drop-in for PardisoSolver / Ma27Solver / Ma57Solver (LINEAR_LEAF_SOLVER)
// ---------------------------------------------------------------------------------------------------------------
class CudaSparseLdlSolver : public DoubleLinearSolver {
public:
    // borrows the matrix exactly like the reference plugins (SURVEY.md 8(b) "Ownership"); (nx, ny, nz) = (locnx,
    // locmy, locmz) of the leaf: the static pivot order needs to know which rows are primal / equality / inequality
    CudaSparseLdlSolver(const SparseSymmetricMatrix& sgm, int nx, int ny, int nz, CudaContext& ctx = CudaContext::instance())
        : matrix_(sgm), n_((int)sgm.size()), ctx_(ctx) {
        assert(nx + ny + nz == n_);
        PIPSGPU_SHIM_CHECK(pipsgpu_leaf_create(ctx_.get(), n_, nx, ny, nz, sgm.krowM(), sgm.jcolM(), sgm.M(), &leaf_));
    }
    ~CudaSparseLdlSolver() override { pipsgpu_leaf_destroy(leaf_); }

    void diagonalChanged(int /*idiag*/, int /*extent*/) override { matrixChanged(); }   // as PardisoSolver.C:136-139

    void matrixChanged() override {
        ctx_.commit_once();                                                   // firstCall(): ordering + symbolic
        PIPSGPU_SHIM_CHECK(pipsgpu_leaf_set_values(leaf_, matrix_.M()));       // Ma27Solver.C:317-324 copies too
        int in[3];
        PIPSGPU_SHIM_CHECK(pipsgpu_leaf_factorize(leaf_, in));
        inertia_ = {(unsigned)in[0], (unsigned)in[1], (unsigned)in[2]};
    }

    using DoubleLinearSolver::solve;
    void solve(Vector<double>& rhs_in) override {
        auto& rhs = dynamic_cast<DenseVector<double>&>(rhs_in);
        assert(rhs.length() == n_);
        PIPSGPU_SHIM_CHECK(pipsgpu_leaf_solve(leaf_, 1, rhs.elements()));
    }
    // rhs i contiguous at rhss + i*n, in place; colSparsity is an optimisation hint in the reference
    // (PardisoSolver.C:276-352) and may be ignored
    void solve(int nrhss, double* rhss, int* /*colSparsity*/) override {
        PIPSGPU_SHIM_CHECK(pipsgpu_leaf_solve(leaf_, nrhss, rhss));
    }
    // PardisoSolver::solve(GeneralMatrix&) (PardisoSolver.C:230-251): the dense block is handed to the solver as ncols
    // right-hand sides of length n, right-hand side j contiguous at &rhs[0][0] + j*n, solved in place
    void solve(GeneralMatrix& rhs_in) override {
        auto& rhs = dynamic_cast<DenseMatrix&>(rhs_in);
        const auto [nrows, ncols] = rhs.n_rows_columns();
        assert(nrows == n_);
        (void)nrows;
        PIPSGPU_SHIM_CHECK(pipsgpu_leaf_solve(leaf_, (int)ncols, &rhs[0][0]));
    }

    // inertia of the COMPUTED pivots (a wrong-signed / vanishing pivot shows up here even though the factorisation went on
    // with a perturbed one): LinearSystem::factorize_with_correct_inertia (LS/LinearSystem.C:296-325) can react to it
    [[nodiscard]] bool reports_inertia() const override { return true; }
    [[nodiscard]] std::tuple<unsigned int, unsigned int, unsigned int> get_inertia() const override { return inertia_; }
    // this node's own regularisation for the batched system-level factorisation (RegularizationStrategy per node)
    void set_regularization(double regP, double regDy, double regDz) {
        PIPSGPU_SHIM_CHECK(pipsgpu_leaf_set_regularization(leaf_, regP, regDy, regDz));
    }

    pipsgpu_leaf* handle() const { return leaf_; }

protected:
    const SparseSymmetricMatrix& matrix_;
    const int n_;
    CudaContext& ctx_;
    pipsgpu_leaf* leaf_ = nullptr;
    std::tuple<unsigned int, unsigned int, unsigned int> inertia_{0, 0, 0};
};

// ---------------------------------------------------------------------------------------------------------------
// dense root solver: drop-in for DeSymIndefSolver (LINEAR_DENSE_SOLVER = SYM_INDEF)
// ---------------------------------------------------------------------------------------------------------------
class CudaDenseLdlSolver : public DoubleLinearSolver {
public:
    // n0 = number of leading primal indices (expected positive pivots), the rest are dual (negative): the
    // (n0, my0 + myl + mzl, 0) inertia of LS/DistributedRootLinearSystem.C:55
    CudaDenseLdlSolver(const DenseSymmetricMatrix& storage, int n0, CudaContext& ctx = CudaContext::instance())
        : matrix_(storage), n_((int)storage.size()), ctx_(ctx) {
        PIPSGPU_SHIM_CHECK(pipsgpu_sc_create(ctx_.get(), n0, 0, n_ - n0, 0, &sc_));
    }
    ~CudaDenseLdlSolver() override { pipsgpu_sc_destroy(sc_); }

    void diagonalChanged(int, int) override { matrixChanged(); }
    void matrixChanged() override {   // DeSymIndefSolver.C:56-97: copy the matrix, dsytrf
        PIPSGPU_SHIM_CHECK(pipsgpu_sc_upload(sc_, matrix_.elements()));
        int in[3];
        PIPSGPU_SHIM_CHECK(pipsgpu_sc_factorize(sc_, in));
        inertia_ = {(unsigned)in[0], (unsigned)in[1], (unsigned)in[2]};
    }
    using DoubleLinearSolver::solve;
    void solve(Vector<double>& rhs_in) override {
        auto& rhs = dynamic_cast<DenseVector<double>&>(rhs_in);
        PIPSGPU_SHIM_CHECK(pipsgpu_sc_solve(sc_, 1, rhs.elements()));
    }
    void solve(int nrhss, double* rhss, int*) override { PIPSGPU_SHIM_CHECK(pipsgpu_sc_solve(sc_, nrhss, rhss)); }
    [[nodiscard]] bool reports_inertia() const override { return true; }
    [[nodiscard]] std::tuple<unsigned int, unsigned int, unsigned int> get_inertia() const override { return inertia_; }
    pipsgpu_sc* handle() const { return sc_; }

private:
    const DenseSymmetricMatrix& matrix_;
    const int n_;
    CudaContext& ctx_;
    pipsgpu_sc* sc_ = nullptr;
    std::tuple<unsigned int, unsigned int, unsigned int> inertia_{0, 0, 0};
};

// ---------------------------------------------------------------------------------------------------------------
// Schur leaf solver: drop-in for PardisoSchurSolver (sLinsysLeafSchurSlv, SC_COMPUTE_BLOCKWISE = false)
// ---------------------------------------------------------------------------------------------------------------
class CudaSchurLeafSolver : public CudaSparseLdlSolver {
public:
    using CudaSparseLdlSolver::CudaSparseLdlSolver;

    // PardisoSchurSolver::firstSolveCall (PardisoSchurSolver.C:84-389) builds the augmented matrix on the first
    // schur_solve; here the border must be known before the symbolic phase, so the leaf linear system hands it over
    // right after construction (the blocks are available there: LS/DistributedLeafLinearSystem.C:214-252).
    void set_border(const SparseMatrix& R, const SparseMatrix& A, const SparseMatrix& C, const SparseMatrix& F,
                    const SparseMatrix& G, int n0, int my0) {
        assert(R.n_rows() == 0 || R.krowM()[R.n_rows()] == 0);   // cross Hessian must be empty (PardisoSchurSolver.C:118)
        (void)R;
        const int myl = (int)F.n_rows(), mzl = (int)G.n_rows();
        PIPSGPU_SHIM_CHECK(pipsgpu_leaf_set_border(leaf_, n0, my0, myl, mzl, A.krowM(), A.jcolM(), A.M(), C.krowM(),
                                                   C.jcolM(), C.M(), F.krowM(), F.jcolM(), F.M(), G.krowM(), G.jcolM(), G.M()));
    }
    // matrixChanged() only records the new values; the numeric factorisation happens fused with the Schur term,
    // exactly as in the reference (sLinsysLeafSchurSlv.C:20-35 -> PardisoSchurSolver::schur_solve :434-477)
    void matrixChanged() override {
        ctx_.commit_once();
        PIPSGPU_SHIM_CHECK(pipsgpu_leaf_set_values(leaf_, matrix_.M()));
    }
    // SC -= Bord^T K^-1 Bord into the device-resident Schur complement of the root (LS/sLinsysLeafSchurSlv.C:29-34)
    void schur_solve(pipsgpu_sc* SC) {
        PIPSGPU_SHIM_CHECK(pipsgpu_leaf_add_schur(leaf_, SC));
        int in[3]; int npert = 0;
        PIPSGPU_SHIM_CHECK(pipsgpu_leaf_get_inertia(leaf_, in, &npert));
        inertia_ = {(unsigned)in[0], (unsigned)in[1], (unsigned)in[2]};
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Schur leaf solver with the reference's EXACT call sequence and signature (plug-in level only, no system-level wiring):
//     make_leaf_solver(kkt)  ->  matrixChanged() every iteration  ->  schur_solve(R, A, C, F, G, DenseSymmetricMatrix& SC)
// (PardisoSchurSolver.h:52-53, called unchanged from sLinsysLeafSchurSlv::addTermToDenseSchurCompl,
// LS/sLinsysLeafSchurSlv.C:20-35). The border is only known at the first schur_solve - as for
// PardisoSchurSolver::firstSolveCall (PardisoSchurSolver.C:84-389) - so every such solver owns a private context whose
// symbolic phase runs at that first call; the Schur term is formed in a device scratch shared by all solvers of the
// process and added to the caller's host matrix. Compatibility path: one factorisation and one n_S x n_S download per
// leaf and iteration, no batching over leaves - the system-level wiring below is the fast path.
// ---------------------------------------------------------------------------------------------------------------
class CudaSchurLeafSolverCompat : public DoubleLinearSolver {
public:
    CudaSchurLeafSolverCompat(const SparseSymmetricMatrix& sgm, int nx, int ny, int nz, CudaContext& shared = CudaContext::instance())
        : matrix_(sgm), n_((int)sgm.size()), shared_(shared) {
        assert(nx + ny + nz == n_);
        PIPSGPU_SHIM_CHECK(pipsgpu_ctx_create(pipsgpu_ctx_device(shared_.get()), &ctx_));
        PIPSGPU_SHIM_CHECK(pipsgpu_ctx_set_stream(ctx_, pipsgpu_ctx_stream(shared_.get())));   // one stream: scratch ops stay ordered
        PIPSGPU_SHIM_CHECK(pipsgpu_leaf_create(ctx_, n_, nx, ny, nz, sgm.krowM(), sgm.jcolM(), sgm.M(), &leaf_));
    }
    ~CudaSchurLeafSolverCompat() override {
        pipsgpu_leaf_destroy(leaf_);
        pipsgpu_ctx_destroy(ctx_);
    }
    void diagonalChanged(int, int) override { matrixChanged(); }
    void matrixChanged() override { dirty_ = true; }      // the numeric factorisation is fused with the Schur term

    // SC (lower triangle, index order [x0; y0; yl; zl]) -= Bord^T K^-1 Bord with Bord = [R^T A^T C^T | F^T | G^T]
    void schur_solve(const SparseMatrix& R, const SparseMatrix& A, const SparseMatrix& C, const SparseMatrix& F,
                     const SparseMatrix& G, DenseSymmetricMatrix& SC) {
        const int n0 = (int)A.n_columns(), myl = (int)F.n_rows(), mzl = (int)G.n_rows();
        const int nS = (int)SC.size(), my0 = nS - n0 - myl - mzl;
        assert(my0 >= 0);
        if (!committed_) {
            assert(R.n_rows() == 0 || R.krowM()[R.n_rows()] == 0);   // cross Hessian must be empty (PardisoSchurSolver.C:118)
            (void)R;
            PIPSGPU_SHIM_CHECK(pipsgpu_leaf_set_border(leaf_, n0, my0, myl, mzl, A.krowM(), A.jcolM(), A.M(), C.krowM(), C.jcolM(),
                                                       C.M(), F.krowM(), F.jcolM(), F.M(), G.krowM(), G.jcolM(), G.M()));
            PIPSGPU_SHIM_CHECK(pipsgpu_ctx_commit(ctx_, 0));
            committed_ = true;
        }
        pipsgpu_sc* scratch = scratch_for(n0, my0, myl, mzl);
        PIPSGPU_SHIM_CHECK(pipsgpu_sc_zero(scratch));
        PIPSGPU_SHIM_CHECK(pipsgpu_leaf_set_values(leaf_, matrix_.M()));
        PIPSGPU_SHIM_CHECK(pipsgpu_leaf_add_schur(leaf_, scratch));
        dirty_ = false;
        int in[3]; int npert = 0;
        PIPSGPU_SHIM_CHECK(pipsgpu_leaf_get_inertia(leaf_, in, &npert));
        inertia_ = {(unsigned)in[0], (unsigned)in[1], (unsigned)in[2]};
        host_.resize((size_t)nS * nS);
        PIPSGPU_SHIM_CHECK(pipsgpu_sc_download(scratch, host_.data()));
        for (int i = 0; i < nS; ++i) {
            double* row = SC[i];
            const double* add = host_.data() + (size_t)i * nS;
            for (int j = 0; j <= i; ++j) row[j] += add[j];
        }
    }

    using DoubleLinearSolver::solve;
    void solve(Vector<double>& rhs_in) override {
        auto& rhs = dynamic_cast<DenseVector<double>&>(rhs_in);
        solve(1, rhs.elements(), nullptr);
    }
    void solve(int nrhss, double* rhss, int*) override {
        if (!committed_) {      // used without a Schur term: plain sparse solver
            PIPSGPU_SHIM_CHECK(pipsgpu_ctx_commit(ctx_, 0));
            committed_ = true;
        }
        if (dirty_) {
            PIPSGPU_SHIM_CHECK(pipsgpu_leaf_set_values(leaf_, matrix_.M()));
            int in[3];
            PIPSGPU_SHIM_CHECK(pipsgpu_leaf_factorize(leaf_, in));
            inertia_ = {(unsigned)in[0], (unsigned)in[1], (unsigned)in[2]};
            dirty_ = false;
        }
        PIPSGPU_SHIM_CHECK(pipsgpu_leaf_solve(leaf_, nrhss, rhss));
    }
    [[nodiscard]] bool reports_inertia() const override { return true; }
    [[nodiscard]] std::tuple<unsigned int, unsigned int, unsigned int> get_inertia() const override { return inertia_; }

private:
    // one device scratch per Schur-complement shape, shared by every compat solver of the process
    pipsgpu_sc* scratch_for(int n0, int my0, int myl, int mzl) {
        struct Entry { int d[4]; pipsgpu_sc* sc; };
        static std::vector<Entry> table;
        for (auto& e : table)
            if (e.d[0] == n0 && e.d[1] == my0 && e.d[2] == myl && e.d[3] == mzl) return e.sc;
        pipsgpu_sc* sc = nullptr;
        PIPSGPU_SHIM_CHECK(pipsgpu_sc_create(shared_.get(), n0, my0, myl, mzl, &sc));
        table.push_back(Entry{{n0, my0, myl, mzl}, sc});
        return sc;
    }
    const SparseSymmetricMatrix& matrix_;
    const int n_;
    CudaContext& shared_;
    pipsgpu_ctx* ctx_ = nullptr;
    pipsgpu_leaf* leaf_ = nullptr;
    bool committed_ = false, dirty_ = true;
    std::vector<double> host_;
    std::tuple<unsigned int, unsigned int, unsigned int> inertia_{0, 0, 0};
};

// ---------------------------------------------------------------------------------------------------------------
// system level: what a `class sLinsysRootAugCuda : public sLinsysRootAug` holds and forwards to
// ---------------------------------------------------------------------------------------------------------------
class CudaKktSystem {
public:
    CudaKktSystem(int n0, int my0, int mz0, int myl, int mzl, const SparseMatrix& A0, const SparseMatrix& C0,
                  const SparseMatrix& F0, const SparseMatrix& G0, CudaContext& ctx = CudaContext::instance())
        : ctx_(ctx), nb0_(n0 + my0 + mz0 + myl + mzl) {
        PIPSGPU_SHIM_CHECK(pipsgpu_kkt_create(ctx_.get(), n0, my0, mz0, myl, mzl, A0.krowM(), A0.jcolM(), A0.M(), C0.krowM(),
                                              C0.jcolM(), C0.M(), F0.krowM(), F0.jcolM(), F0.M(), G0.krowM(), G0.jcolM(),
                                              G0.M(), &kkt_));
    }
    ~CudaKktSystem() { pipsgpu_kkt_destroy(kkt_); }
    // NCCL bootstrap over the reference's MPI communicator: rank 0 calls pipsgpu_nccl_unique_id, MPI_Bcast(id, 128,
    // MPI_CHAR, 0, mpiComm), then every rank calls attach_comm (replaces mpiComm of DistributedRootLinearSystem)
    void attach_comm(const char id[128], int rank, int nranks) {
        PIPSGPU_SHIM_CHECK(pipsgpu_ctx_init_comm(ctx_.get(), id, rank, nranks));
    }
    // children owned by this rank in ascending order (DistributedDummyLinearSystem children are skipped)
    void add_leaf(CudaSchurLeafSolver& leaf) { PIPSGPU_SHIM_CHECK(pipsgpu_kkt_add_leaf(kkt_, leaf.handle())); }
    void setup(int threads = 0) {
        PIPSGPU_SHIM_CHECK(pipsgpu_kkt_setup(kkt_, threads));
        ctx_.mark_committed();
    }
    // AbstractLinearSystem::factorize for the whole rank: diagonals down, batched leaf LDL^T + Schur terms, NCCL
    // reduceKKT, finalizeKKTdense, dense LDL^T. Vectors are the owned leaves' dd / nomegaInv concatenated.
    std::tuple<unsigned, unsigned, unsigned> factorize(const double* dd0, const double* nomegainv0, const double* nomegainv_link,
                                                       const double* dd_leaves, const double* nomegainv_leaves,
                                                       double regP = 0.0, double regDy = 0.0, double regDz = 0.0) {
        const double reg[3] = {regP, regDy, regDz};
        int in[3];
        PIPSGPU_SHIM_CHECK(pipsgpu_kkt_factorize(kkt_, dd0, nomegainv0, nomegainv_link, dd_leaves, nomegainv_leaves, reg, in));
        return {(unsigned)in[0], (unsigned)in[1], (unsigned)in[2]};
    }
    // DistributedLinearSystem::solveCompressed (Lsolve; Dsolve; Ltsolve) in place
    void solveCompressed(double* b0, double* b_leaves) { PIPSGPU_SHIM_CHECK(pipsgpu_kkt_solve_compressed(kkt_, b0, b_leaves)); }
    // LinearSystem::solveCompressedBiCGStab (LS/LinearSystem.C:550-798) on the device, preconditioned by solveCompressed, with
    // the full-KKT matvec system_mult (:809-844): right-hand side down once, step back once. Returns the reference's
    // IterativeSolverSolutionStatus code (LS/LinearSystem.h:45-47).
    int solve(double* b0, double* b_leaves, double tol, int max_iter, int* n_iter = nullptr, double* rel_residual = nullptr) {
        int status = 0;
        PIPSGPU_SHIM_CHECK(pipsgpu_kkt_solve(kkt_, b0, b_leaves, tol, max_iter, &status, n_iter, rel_residual));
        return status;
    }
    // sparse root KKT (sLinsysRootAug with hasSparseKkt): lower CSR pattern of the Schur complement, declared once
    void set_sparse_pattern(const int* krow, const int* jcol, int threads = 0) {
        PIPSGPU_SHIM_CHECK(pipsgpu_sc_set_sparse_pattern(pipsgpu_kkt_sc(kkt_), krow, jcol, threads));
    }
    pipsgpu_kkt* handle() const { return kkt_; }
    pipsgpu_sc* schur_complement() const { return pipsgpu_kkt_sc(kkt_); }
    int root_rhs_length() const { return nb0_; }

private:
    CudaContext& ctx_;
    pipsgpu_kkt* kkt_ = nullptr;
    int nb0_;
};

}  // namespace pipsgpu_shim

#endif  // PIPSGPU_SHIM_HPP
