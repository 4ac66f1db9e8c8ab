/* pipsgpu.h - C ABI of the B200-native KKT factor/solve path for PIPS-IPM++ (drop-in boundary).
 *
 * Every entry point is what a thin C++ subclass inside PIPS-IPM++ would call; the reference interface each
 * one replaces is cited as  path:line  relative to the reference repository root
 * (LS = PIPS-IPM/Core/KKTFormulation/LinearSystems, SOL = PIPS-IPM/Core/LinearSolvers).
 *
 * Conventions
 *  - plain C types only; all index arrays are 0-based int32 CSR exactly as the reference's SparseStorage
 *    (krowM/jcolM/M, PIPS-IPM/Core/LinearAlgebra/Sparse/SparseStorage.h:46-51); symmetric matrices are LOWER CSR
 *    with a full diagonal (LinearAlgebra/Sparse/SparseSymmetricMatrix.h:18-41).
 *  - every double* argument is a HOST pointer unless its name starts with d_ ; copies are issued on the context's
 *    stream; calls are stream-ordered per context and return after the results they hand back are on the host.
 *  - return value 0 = success, nonzero = error (message via pipsgpu_last_error()); the C++ shim reproduces the
 *    reference's print + MPI_Abort convention (SOL/PardisoSolver/PardisoSolver/PardisoSolver.C:201-204).
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails with PIPSGPU_ERR_NO_DEVICE.
 *    Only the pipsgpu_symbolic_* and pipsgpu_partition entry points are host-only.
 */
#ifndef PIPSGPU_H
#define PIPSGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIPSGPU_OK 0
#define PIPSGPU_ERR_INVALID 1
#define PIPSGPU_ERR_NO_DEVICE 2
#define PIPSGPU_ERR_CUDA 3
#define PIPSGPU_ERR_NCCL 4
#define PIPSGPU_ERR_STATE 5

typedef struct pipsgpu_ctx pipsgpu_ctx;           /* one per process / GPU (one MPI rank of the reference)      */
typedef struct pipsgpu_leaf pipsgpu_leaf;         /* one scenario block: DistributedLeafLinearSystem + its solver */
typedef struct pipsgpu_sc pipsgpu_sc;             /* dense root Schur complement + DeSymIndefSolver               */
typedef struct pipsgpu_kkt pipsgpu_kkt;           /* sLinsysRootAug: root + owned leaves, factorize/solveCompressed */
typedef struct pipsgpu_symbolic pipsgpu_symbolic; /* host-only symbolic analysis result (inspection / tests)      */

const char* pipsgpu_last_error(void);
int pipsgpu_version(void);

/* ---------------------------------------------------------------------------------------------------------
 * context
 * --------------------------------------------------------------------------------------------------------- */
/* one context per GPU; owns the stream, the device pools of all leaves and the (optional) NCCL communicator */
int pipsgpu_ctx_create(int device, pipsgpu_ctx** ctx);
int pipsgpu_ctx_destroy(pipsgpu_ctx* ctx);
int pipsgpu_ctx_synchronize(pipsgpu_ctx* ctx);
/* run on a caller-owned CUDA stream (cudaStream_t passed as void*), e.g. the host application's or torch's current
 * stream, so that the caller's own events bracket the work. The context does not destroy a borrowed stream. */
int pipsgpu_ctx_set_stream(pipsgpu_ctx* ctx, void* cuda_stream);
/* the cudaStream_t every call of this context is enqueued on (its own, or the one given to pipsgpu_ctx_set_stream) */
void* pipsgpu_ctx_stream(pipsgpu_ctx* ctx);
/* CUDA device index of the context (-1 for a null handle) */
int pipsgpu_ctx_device(pipsgpu_ctx* ctx);
/* NCCL bootstrap without MPI: rank 0 creates the id, the host program ships the 128 bytes to the other ranks
 * (torch.distributed / MPI_Bcast / a file), every rank calls init_comm. Replaces the MPI communicator the reference
 * passes around (LS/DistributedRootLinearSystem.C:47-57, mpiComm). */
int pipsgpu_nccl_unique_id(char id[128]);
int pipsgpu_ctx_init_comm(pipsgpu_ctx* ctx, const char id[128], int rank, int nranks);
/* contiguous balanced partition of N children over P ranks, Readers/Distributed/DistributedTree.C:437-475
 * (mapChildrenToNSubTrees): writes lo[rank], hi[rank]. Host-only. */
int pipsgpu_partition(int n_children, int n_ranks, int* lo, int* hi);

/* ---------------------------------------------------------------------------------------------------------
 * leaf = DistributedLeafLinearSystem (LS/DistributedLeafLinearSystem.C) + DoubleLinearSolver plugin
 * (SOL/DoubleLinearSolver.h:24-72; PardisoSolver / PardisoSchurSolver / Ma27Solver / Ma57Solver / MumpsSolverLeaf)
 * --------------------------------------------------------------------------------------------------------- */
/* create_kkt (LS/DistributedLeafLinearSystem.C:52-72) has been done by the caller: lower CSR of
 * [Q+D W^T Dm^T; W 0 0; Dm 0 0] with full diagonal; (nx,ny,nz) = (locnx,locmy,locmz). `val` holds the constant
 * off-diagonal values (and any diagonal, overwritten later). The pattern is fixed for the lifetime of the leaf
 * (matrixRebuild is not supported). Replaces the plugin constructors, e.g. PardisoSolver.C:30-56. */
int pipsgpu_leaf_create(pipsgpu_ctx* ctx, int n, int nx, int ny, int nz, const int* krow, const int* jcol,
                        const double* val, pipsgpu_leaf** leaf);
/* Border Bord_i = [R 0 F^T G^T; A 0 0 0; C 0 0 0] as the reference's BorderBiBlock (LS/RACFG_BLOCK.h:14-53):
 * A (ny x n0) = getLocalA, C (nz x n0) = getLocalC, F (myl x nx) = getLocalF, G (mzl x nx) = getLocalG, all CSR;
 * R (cross Hessian) must be empty, as for the reference's own Schur solver (PardisoSchurSolver.C:118).
 * Schur complement index space: [0,n0) | [n0,n0+my0) untouched | myl | mzl. Any matrix may be NULL/empty. */
int pipsgpu_leaf_set_border(pipsgpu_leaf* leaf, int n0, int my0, int myl, int mzl,
                            const int* A_krow, const int* A_jcol, const double* A_val,
                            const int* C_krow, const int* C_jcol, const double* C_val,
                            const int* F_krow, const int* F_jcol, const double* F_val,
                            const int* G_krow, const int* G_jcol, const double* G_val);
/* ordering + symbolic factorisation on the host (first matrixChanged() -> firstCall() in the reference plugins:
 * PardisoSolver.C:142-145, Ma27Solver.C:44-88, Ma57Solver.C:56-78). Done once, reused by every factorisation. */
int pipsgpu_leaf_analyze(pipsgpu_leaf* leaf);
/* analyze every not-yet-analyzed leaf of the context with `threads` host threads, then build the device pools and
 * the level schedule shared by all leaves (batched launches). Must be called once after the last leaf_create. */
int pipsgpu_ctx_commit(pipsgpu_ctx* ctx, int threads);
/* put_primal_diagonal / clear_dual_equality_diagonal / put_dual_inequalites_diagonal / add_regularization_local_kkt
 * (LS/DistributedLeafLinearSystem.C:88-143): K[0:nx]=dd+regP, K[nx:nx+ny]=-regDy, K[nx+ny:]=nomegainv-regDz.
 * Any pointer may be NULL (= zeros). Marks the leaf as changed. */
int pipsgpu_leaf_set_diagonal(pipsgpu_leaf* leaf, const double* dd, const double* nomegainv, double regP,
                              double regDy, double regDz);
/* this leaf's own regularisation (LS/DistributedLeafLinearSystem.C:121-143: every node of the tree carries its own
 * primal / dual regularisation, chosen by its RegularizationStrategy in LinearSystem::factorize_with_correct_inertia,
 * LS/LinearSystem.C:296-325). The values are ADDED to the regP / regDy / regDz of every later diagonal refresh of this
 * leaf (pipsgpu_leaf_set_diagonal, pipsgpu_kkt_factorize) until changed again. */
int pipsgpu_leaf_set_regularization(pipsgpu_leaf* leaf, double regP, double regDy, double regDz);
/* full value refresh of the lower CSR (Ma27Solver.C:317-324 copies all values on every matrixChanged) */
int pipsgpu_leaf_set_values(pipsgpu_leaf* leaf, const double* val);
/* DoubleLinearSolver::matrixChanged() + get_inertia(): numeric LDL^T of this leaf only (batched entry point below
 * is the fast path). inertia = {positive, negative, zero}; may be NULL. */
int pipsgpu_leaf_factorize(pipsgpu_leaf* leaf, int inertia[3]);
/* DoubleLinearSolver::solve(int nrhss, double* rhss, int* colSparsity) (SOL/DoubleLinearSolver.h:58-59):
 * rhs i contiguous at rhs + i*n, solved in place, original (unpermuted) ordering. */
int pipsgpu_leaf_solve(pipsgpu_leaf* leaf, int nrhs, double* rhs);
/* DoubleLinearSolver::get_inertia of the last factorisation; n_perturbed = number of perturbed pivots
 * (PARDISO iparm[13] analogue) */
int pipsgpu_leaf_get_inertia(pipsgpu_leaf* leaf, int inertia[3], int* n_perturbed);
/* statistics of the symbolic phase: out[0]=nnz(L) (scalar, K part), out[1]=border-row entries, out[2]=stored panel
 * entries, out[3]=#fronts, out[4]=#levels, out[5]=#2x2 pivots, out[6]=max front rows, out[7]=contribution entries;
 * *flops = sum_j eta_j(eta_j+2) (SURVEY.md 8(d) stage-1 count) */
int pipsgpu_leaf_get_stats(pipsgpu_leaf* leaf, int64_t out[8], double* flops);
/* executed flops of this leaf's numeric phase (valid after commit): out[0] = canonical sum_j eta_j(eta_j+2), out[1] =
 * Schur-term flops = lower triangle of L_B D L_B^T over all fronts (sum_f w_f nb_f (nb_f+1)), out[2] = all update flops
 * (ancestor columns + border x border), out[3] = panel flops (W21 = A21 X^T). Measurement only. */
int pipsgpu_leaf_get_flops(pipsgpu_leaf* leaf, double out[4]);
/* factor download for parity checks: L = the leaf's panel pool (out[2] = stored panel entries of get_stats; per front an
 * m x w column-major panel, unit diagonal, layout of pipsgpu_symbolic_export), dinv_d / dinv_e = block-diagonal D^-1
 * (length n each, permuted order; dinv_e[k] couples k and k+1 of a 2x2 pivot). Any pointer may be NULL. */
int pipsgpu_leaf_download_factor(pipsgpu_leaf* leaf, double* L, double* dinv_d, double* dinv_e);
int pipsgpu_leaf_destroy(pipsgpu_leaf* leaf);

/* batched numeric factorisation of all changed leaves of the context AND accumulation of their Schur terms:
 * sc -= sum_i Bord_i^T K_i^-1 Bord_i  (factor2 loop + assembleLocalKKT, LS/DistributedRootLinearSystem.C:206-243,
 * LS/sLinsysRootAug.C:210-227; fused the way PardisoSchurSolver::schur_solve does it, PardisoSchurSolver.C:434-477).
 * sc may be NULL (factorisation only). */
int pipsgpu_ctx_factorize_leaves(pipsgpu_ctx* ctx, pipsgpu_sc* sc);
/* addTermToSchurCompl for a single leaf (LS/DistributedRootLinearSystem.C:1490-1506). The Schur term is a by-product
 * of the elimination, so this REFACTORISES the leaf with the Schur complement attached (the way
 * PardisoSchurSolver::schur_solve does, PardisoSchurSolver.C:434-477) and leaves it factorised. The context must be
 * committed. Per-leaf path, for drop-in completeness; pipsgpu_ctx_factorize_leaves is the batched fast path. */
int pipsgpu_leaf_add_schur(pipsgpu_leaf* leaf, pipsgpu_sc* sc);

/* ---------------------------------------------------------------------------------------------------------
 * sc = dense root Schur complement (LS/DistributedRootLinearSystem.C:161-173 createKKT, DenseSymmetricMatrix) and
 * its solver (SOL/DenseSymmetricIndefinitSolver/DeSymIndefSolver.C). Lower triangle, row-major, device resident.
 * --------------------------------------------------------------------------------------------------------- */
int pipsgpu_sc_create(pipsgpu_ctx* ctx, int n0, int my0, int myl, int mzl, pipsgpu_sc** sc);
int pipsgpu_sc_destroy(pipsgpu_sc* sc);
/* sLinsysRootAug with hasSparseKkt (LS/sLinsysRootAug.C:79-208 finalizeKKTdist, 1629-1767 finalizeKKTsparse: kkt is a
 * SparseSymmetricMatrix whose pattern is precomputed once). Declares the lower-triangular CSR pattern (every diagonal entry
 * stored) of the Schur complement in its index order [x0; y0; yl; zl]. From then on
 *   - reduceKKT all-reduces only the nnz pattern values (reduceKKTsparse / reduceToAllProcs,
 *     LS/DistributedRootLinearSystem.C:885-925) instead of the packed dense triangle,
 *   - factorize / solve run the sparse multifrontal LDL^T that also serves the leaves, in a private context on the same
 *     stream (replaces the sparse root solver, SOL/PardisoSolver/PardisoIndefSolver, Ma57 root solver); symbolic analysis
 *     happens here, once. If a pivot has to be perturbed or the inertia is not (n0, n - n0, 0) the dense solver (which can
 *     pivot) takes over for that factorisation.
 * The assembly target stays the dense device buffer (leaves add their Schur terms there); entries outside the pattern
 * are ignored. Calling it again replaces the pattern and redoes the symbolic analysis (DoubleLinearSolver::matrixRebuild,
 * SOL/DoubleLinearSolver.h:39). */
int pipsgpu_sc_set_sparse_pattern(pipsgpu_sc* sc, const int* krow, const int* jcol, int threads);
/* 1 if the last pipsgpu_sc_factorize / pipsgpu_kkt_factorize used the sparse root solver */
int pipsgpu_sc_uses_sparse_solver(pipsgpu_sc* sc);
/* initializeKKT / myAtPutZeros (LS/DistributedRootLinearSystem.C:840-847, 1467-1488) */
int pipsgpu_sc_zero(pipsgpu_sc* sc);
/* reduceKKT -> reduceKKTdense (LS/DistributedRootLinearSystem.C:849-881): NCCL all-reduce (sum, f64) of the packed
 * pieces  lower(n0) | rect (myl+mzl) x n0 | lower(myl+mzl)  in one buffer. No-op without a communicator. */
int pipsgpu_sc_allreduce(pipsgpu_sc* sc);
/* DenseSymmetricMatrix::add_matrix_at (LinearAlgebra/Dense/DenseSymmetricMatrix.C:378): sc[row0+i][col0+j] += alpha*M_ij
 * for a CSR matrix M (m x n); entries above the diagonal are mirrored into the lower triangle. */
int pipsgpu_sc_add_sparse(pipsgpu_sc* sc, int row0, int col0, int m, const int* krow, const int* jcol,
                          const double* val, double alpha);
/* atAddDiagonal (LS/sLinsysRootAug.C:229-263, 1792-1796) */
int pipsgpu_sc_add_diag(pipsgpu_sc* sc, int offset, int n, const double* d, double alpha);
/* DeSymIndefSolver::matrixChanged (DeSymIndefSolver.C:56-97) + get_inertia (:135-160): dense LDL^T. */
int pipsgpu_sc_factorize(pipsgpu_sc* sc, int inertia[3]);
/* DeSymIndefSolver::solve / solveSynchronized (DeSymIndefSolver.C:99-130): nrhs right-hand sides of length n_S,
 * rhs i contiguous, in place. */
int pipsgpu_sc_solve(pipsgpu_sc* sc, int nrhs, double* rhs);
/* lower triangle, row-major n_S x n_S (upper part zero) - for parity checks; upload for solver-only tests */
int pipsgpu_sc_download(pipsgpu_sc* sc, double* lower_rowmajor);
int pipsgpu_sc_upload(pipsgpu_sc* sc, const double* lower_rowmajor);
int pipsgpu_sc_dim(pipsgpu_sc* sc);

/* ---------------------------------------------------------------------------------------------------------
 * kkt = sLinsysRootAug + DistributedRootLinearSystem for the two-level dense-SC case: the per-iteration entry
 * points AbstractLinearSystem::factorize / DistributedLinearSystem::solveCompressed use.
 * --------------------------------------------------------------------------------------------------------- */
/* root blocks: A0 (my0 x n0; the root's "B", LS/sLinsysRootAug.C:287,315), C0 (mz0 x n0), F0 (myl x n0), G0 (mzl x n0) */
int pipsgpu_kkt_create(pipsgpu_ctx* ctx, int n0, int my0, int mz0, int myl, int mzl,
                       const int* A0_krow, const int* A0_jcol, const double* A0_val,
                       const int* C0_krow, const int* C0_jcol, const double* C0_val,
                       const int* F0_krow, const int* F0_jcol, const double* F0_val,
                       const int* G0_krow, const int* G0_jcol, const double* G0_val, pipsgpu_kkt** kkt);
int pipsgpu_kkt_destroy(pipsgpu_kkt* kkt);
/* children owned by this rank, in ascending child order (createChildren, LS/DistributedRootLinearSystem.C:711-758);
 * non-owned children are simply not added (DistributedDummyLinearSystem). Border must be set on the leaf. */
int pipsgpu_kkt_add_leaf(pipsgpu_kkt* kkt, pipsgpu_leaf* leaf);
/* after the last add_leaf: analyse + commit (calls pipsgpu_ctx_commit) */
int pipsgpu_kkt_setup(pipsgpu_kkt* kkt, int threads);
/* LinearSystem::factorize data movement + DistributedRootLinearSystem::factor2 (LS/LinearSystem.C:171-202,
 * LS/DistributedRootLinearSystem.C:206-243): diagonals down, leaf LDL^T, Schur terms, allreduce, finalizeKKTdense
 * (LS/sLinsysRootAug.C:1769-1813), dense LDL^T. dd_leaves / nomegainv_leaves are the owned leaves' vectors
 * concatenated in add order (lengths sum nx_i / sum nz_i). reg = {regP, regDy, regDz} or NULL.
 * inertia_root may be NULL. */
int pipsgpu_kkt_factorize(pipsgpu_kkt* kkt, const double* dd0, const double* nomegainv0, const double* nomegainv_link,
                          const double* dd_leaves, const double* nomegainv_leaves, const double reg[3],
                          int inertia_root[3]);
/* DistributedLinearSystem::solveCompressed = Lsolve; Dsolve; Ltsolve (LS/DistributedLinearSystem.C:409-420,
 * LS/sLinsysRootAug.C:323-466, LS/DistributedLeafLinearSystem.C:153-212): b0 = [x0;y0;z0;yl;zl]
 * (n0+my0+mz0+myl+mzl), b_leaves = owned leaves' [x_i;y_i;z_i] concatenated; solved in place. The b0 all-reduce of
 * Lsolve (LS/sLinsysRootAug.C:340-341) runs over NCCL when a communicator is attached; on ranks > 0 the incoming b0
 * is ignored exactly as in the reference (:332-333). */
int pipsgpu_kkt_solve_compressed(pipsgpu_kkt* kkt, double* b0, double* b_leaves);
/* LinearSystem::solveXYZS with OUTER_SOLVE = 2 (LS/LinearSystem.C:449-548): solveCompressedBiCGStab (:550-798) on the
 * device - BiCGStab on the (regularised) full KKT system in the node layout of DistributedVector, preconditioned by
 * solveCompressed, full-KKT matvec = system_mult (:809-844). b0 / b_leaves as for pipsgpu_kkt_solve_compressed: the
 * right-hand side goes down once, the solution comes back once (in place); dot products are reduced over NCCL.
 * tol = OUTER_BICG_TOL, max_iter = OUTER_BICG_MAX_ITER (75). *status receives IterativeSolverSolutionStatus
 * (LS/LinearSystem.h:45-47: -1 max iterations, 0 converged, 1 skipped, 3 stagnation, 4 breakdown, 5 diverged), *n_iter the
 * iteration count (0 when skipped), *rel_residual ||b - K x|| / ||b||. Any of the three may be NULL. */
int pipsgpu_kkt_solve(pipsgpu_kkt* kkt, double* b0, double* b_leaves, double tol, int max_iter, int* status, int* n_iter,
                      double* rel_residual);
/* device-resident variants for measuring the kernels without PCIe traffic: reuse the diagonals / the right-hand side
 * of the last pipsgpu_kkt_factorize / pipsgpu_kkt_solve_compressed call, which stay in HBM. solve_compressed_resident
 * is asynchronous (no D2H, no host synchronisation); factorize_resident synchronises once to read the inertia. */
int pipsgpu_kkt_factorize_resident(pipsgpu_kkt* kkt, int inertia_root[3]);
int pipsgpu_kkt_solve_compressed_resident(pipsgpu_kkt* kkt);
/* the root Schur complement object (for download / parity checks) */
pipsgpu_sc* pipsgpu_kkt_sc(pipsgpu_kkt* kkt);
/* device time in ms of the stages of the last factorize: out[0]=H2D+assembly, [1]=leaf LDL^T + Schur terms,
 * [2]=allreduce, [3]=finalize, [4]=root LDL^T ; and of the last solve_compressed: [5]=Lsolve, [6]=Dsolve, [7]=Ltsolve */
int pipsgpu_kkt_stage_times(pipsgpu_kkt* kkt, double out[8]);
/* device time in ms of single kernels, CUDA events on the context's stream: out[0] = the streaming solve launch of the
 * last Lsolve (one K^-1 application over all blocks), [1] = factor -> solve-pool pack of the last factorize,
 * [2] = dense root solve kernel of the last Dsolve, [3] = the Schur-term (border x border) kernels of the last
 * factorize, summed over the factorisation batches (measurement only, no reference counterpart) */
int pipsgpu_kkt_kernel_times(pipsgpu_kkt* kkt, double out[4]);
/* sizes of the context pools after commit: out = {L entries, C entries, solve-pool entries, fronts, KKT rows, tree
 * levels, nnz(K), nnz(Bord)} (measurement only) */
int pipsgpu_ctx_pool_sizes(pipsgpu_ctx* ctx, int64_t out[8]);
/* number of leaf batches of the numeric factorisation (contribution / scratch pools are sized for one batch) */
int pipsgpu_ctx_num_batches(pipsgpu_ctx* ctx);
/* measured FP64 tensor-pipe (DMMA, mma.sync m8n8k4 f64) peak of this GPU in TFLOP/s: register-resident chains on all
 * SMs, CUDA-event timed; the denominator of the "fraction of FP64 tensor peak" figures (no reference counterpart) */
int pipsgpu_bench_dmma_peak(pipsgpu_ctx* ctx, double* tflops);
/* measured rate in TFLOP/s of the library's own FP64 tensor-pipe GEMM (dgemm_nt_kernel: TMA-staged operands, DMMA) on an
 * n x n x n lower-triangular update - the kernel behind the root LDL^T trailing update and the packed Schur-term update
 * (measurement only, no reference counterpart) */
int pipsgpu_bench_dgemm(pipsgpu_ctx* ctx, int n, double* tflops);
/* number of kernels launched by this context since creation (bench.py "gpu_launches") */
int64_t pipsgpu_ctx_launch_count(pipsgpu_ctx* ctx);

/* ---------------------------------------------------------------------------------------------------------
 * host-only symbolic analysis (no GPU needed): used by the CPU tests and by oracle/mf_ldlt.c
 * --------------------------------------------------------------------------------------------------------- */
int pipsgpu_symbolic_create(int n, int nx, int ny, int nz, const int* krow, const int* jcol, const double* val,
                            int nB, int64_t nnzB, const int* brow, const int* bcol, int max_width,
                            pipsgpu_symbolic** sym);
int pipsgpu_symbolic_destroy(pipsgpu_symbolic* sym);
/* sizes: out[0]=n out[1]=nB out[2]=ns out[3]=nlevels out[4]=rowidx length out[5]=Lsize out[6]=Csize out[7]=nnzK
 * out[8]=nnzB out[9]=nnzL out[10]=nnzL_border out[11]=n_pairs out[12]=#children entries */
int pipsgpu_symbolic_sizes(pipsgpu_symbolic* sym, int64_t out[13], double* flops);
/* copy-out of the arrays (caller allocates with the sizes above; any pointer may be NULL) */
int pipsgpu_symbolic_export(pipsgpu_symbolic* sym, int* perm, int8_t* ptype, int8_t* sgn, int* sn_col,
                            int64_t* sn_rowptr, int* rowidx, int* relidx, int* sn_na, int* sn_nb, int* sn_parent,
                            int* sn_level, int64_t* Loff, int64_t* Coff, int64_t* kmap, int64_t* bmap,
                            int* child_ptr, int* child_idx);

/* grouping of the big fronts for the deferred border x border Schur-complement updates (host-only inspection of what
 * pipsgpu_ctx_commit builds; no counterpart in the reference, whose plugins keep their supernode scheduling private -
 * PardisoSchurSolver.C:434-477 just receives the finished Schur complement). Call once with null arrays to get
 * counts = {groups, members, positions}, then with arrays of those sizes (mem_ptr: groups+1, pos_ptr: members+1). */
int pipsgpu_symbolic_border_groups(pipsgpu_symbolic* sym, int big_u, double frac, int64_t counts[3], int* root, int* N,
                                   int* mem_ptr, int* mem_front, int64_t* pos_ptr, int* pos);

#ifdef __cplusplus
}
#endif
#endif /* PIPSGPU_H */
