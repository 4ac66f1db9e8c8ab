// Vector / sparse kernels of the outer solve on the device: KKT matvec of the leaves (symmetric COO SpMV on the resident
// lower-CSR values), border products with a scale factor, deterministic two-stage reductions, axpby. They serve
//   * the iterative refinement of the leaf solves (PARDISO iparm[7]=2 analogue, PardisoProjectSolver.C:68-77,
//     Ma27Solver.C:184-249), and
//   * the device-resident BiCGStab of LinearSystem::solveCompressedBiCGStab with the full-KKT matvec
//     LinearSystem::system_mult (LS/LinearSystem.C:550-798, 809-844), see pipsgpu_kkt_solve in pipsgpu.cu.
// All of them are HBM-bound streaming kernels (12 bytes per nonzero, 8-24 bytes per vector entry).
#include "kernels.cuh"

namespace pipsgpu {

static inline int kr_cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// y[row] += alpha * v * x[col] (+ the mirrored entry): lower-triangular COO of a symmetric matrix, global positions
__global__ void sym_spmv_kernel(int64_t nnz, const int* __restrict__ row, const int* __restrict__ col,
                                const double* __restrict__ val, double alpha, const double* __restrict__ x, double* __restrict__ y) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    const int r = row[k], c = col[k];
    const double v = alpha * val[k];
    atomicAdd(&y[r], v * x[c]);
    if (r != c) atomicAdd(&y[c], v * x[r]);
}
void launch_sym_spmv(int64_t nnz, const int* row, const int* col, const double* val, double alpha, const double* x, double* y,
                     cudaStream_t s) {
    if (nnz <= 0) return;
    sym_spmv_kernel<<<kr_cdiv(nnz, 256), 256, 0, s>>>(nnz, row, col, val, alpha, x, y);
    count_launch("sym_spmv_kernel");
}

// t[brow] += alpha * val * u[bidx]   and   b0[bidx] += alpha * val * z[brow]   (border triplets)
__global__ void border_axpy_kernel(const int* __restrict__ brow, const int* __restrict__ bidx, const double* __restrict__ val,
                                   int64_t nnz, double alpha, const double* __restrict__ u, double* __restrict__ t, bool transpose) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    if (transpose) atomicAdd(&t[bidx[k]], alpha * val[k] * u[brow[k]]);
    else atomicAdd(&t[brow[k]], alpha * val[k] * u[bidx[k]]);
}
void launch_border_axpy(const int* brow, const int* bidx, const double* val, int64_t nnz, double alpha, const double* u, double* t,
                        bool transpose, cudaStream_t s) {
    if (nnz <= 0) return;
    border_axpy_kernel<<<kr_cdiv(nnz, 256), 256, 0, s>>>(brow, bidx, val, nnz, alpha, u, t, transpose);
    count_launch("border_axpy_kernel");
}

// y = a * x + b * y   (b == 0: y is not read)
__global__ void axpby_kernel(int64_t n, double a, const double* __restrict__ x, double b, double* __restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    y[i] = (b == 0.0) ? a * x[i] : fma(a, x[i], b * y[i]);
}
void launch_axpby(int64_t n, double a, const double* x, double b, double* y, cudaStream_t s) {
    if (n <= 0) return;
    axpby_kernel<<<kr_cdiv(n, 256), 256, 0, s>>>(n, a, x, b, y);
    count_launch("axpby_kernel");
}

// y += alpha * d .* x
__global__ void diag_axpy_kernel(int n, double alpha, const double* __restrict__ d, const double* __restrict__ x, double* __restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = fma(alpha * (d ? d[i] : 1.0), x[i], y[i]);
}
void launch_diag_axpy(int n, double alpha, const double* d, const double* x, double* y, cudaStream_t s) {
    if (n <= 0) return;
    diag_axpy_kernel<<<kr_cdiv(n, 128), 128, 0, s>>>(n, alpha, d, x, y);
    count_launch("diag_axpy_kernel");
}

// deterministic dot product: DOT_BLOCKS partial sums in a fixed order, then one block adds them in a fixed order
constexpr int DOT_BLOCKS = 1024;
__global__ void __launch_bounds__(256) dot_partial_kernel(int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                                                          double* __restrict__ partial) {
    __shared__ double sh[256];
    double acc = 0.0;
    const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * chunk, hi = (lo + chunk < n) ? lo + chunk : n;
    for (int64_t i = lo + threadIdx.x; i < hi; i += 256) acc = fma(x[i], y[i], acc);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) dot_final_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ out) {
    __shared__ double sh[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) acc += partial[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}
// out[0] = <x, y> over n entries; `partial` holds DOT_BLOCKS doubles of scratch
void launch_dot(int64_t n, const double* x, const double* y, double* partial, double* out, cudaStream_t s) {
    if (n <= 0) { cudaMemsetAsync(out, 0, sizeof(double), s); return; }
    const int nb = (int)((n + 255) / 256 < DOT_BLOCKS ? (n + 255) / 256 : DOT_BLOCKS);
    dot_partial_kernel<<<nb, 256, 0, s>>>(n, x, y, partial);
    count_launch("dot_partial_kernel");
    dot_final_kernel<<<1, 256, 0, s>>>(partial, nb, out);
    count_launch("dot_final_kernel");
}

// ---- sparse root Schur complement (pipsgpu_sc_set_sparse_pattern): pattern values <-> dense assembly buffer ----
// out[e] = A[idx[e]]  /  A[idx[e]] = in[e]
__global__ void gather_idx_kernel(const double* __restrict__ A, const int64_t* __restrict__ idx, int64_t n, double* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) out[e] = A[idx[e]];
}
__global__ void scatter_idx_kernel(double* __restrict__ A, const int64_t* __restrict__ idx, int64_t n, const double* __restrict__ in) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) A[idx[e]] = in[e];
}
void launch_gather_idx(const double* A, const int64_t* idx, int64_t n, double* out, cudaStream_t s) {
    if (n <= 0) return;
    gather_idx_kernel<<<kr_cdiv(n, 256), 256, 0, s>>>(A, idx, n, out);
    count_launch("gather_idx_kernel");
}
void launch_scatter_idx(double* A, const int64_t* idx, int64_t n, const double* in, cudaStream_t s) {
    if (n <= 0) return;
    scatter_idx_kernel<<<kr_cdiv(n, 256), 256, 0, s>>>(A, idx, n, in);
    count_launch("scatter_idx_kernel");
}
// *eps_out = rel * max(1e-300, max_i |v_i|): static-pivot threshold of the sparse root factorisation, one block
__global__ void __launch_bounds__(1024) absmax_eps_kernel(const double* __restrict__ v, int64_t n, double rel, double* __restrict__ eps_out) {
    __shared__ double sh[1024];
    double m = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) m = fmax(m, fabs(v[i]));
    sh[threadIdx.x] = m;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) *eps_out = rel * fmax(sh[0], 1e-300);
}
void launch_absmax_eps(const double* v, int64_t n, double rel, double* eps_out, cudaStream_t s) {
    absmax_eps_kernel<<<1, 1024, 0, s>>>(v, n, rel, eps_out);
    count_launch("absmax_eps_kernel");
}

// ---- leaf diagonals of one factorisation, assembled on the device (pipsgpu_kkt_factorize) ----
// dvals[p] of leaf l (positions pos_off .. pos_off + n): [ dd_l + regP | -regDy | nomegainv_l - regDz ] with the leaf's own
// regularisation added; dd / nz are the caller's concatenated vectors (null = zeros). One thread per position, the leaf
// is found by bisection of the (ascending) position offsets.
__global__ void build_dvals_kernel(int64_t npos, const LeafDiag* __restrict__ tab, int nl, const double* __restrict__ dd,
                                   const double* __restrict__ nz, double regP, double regDy, double regDz, double* __restrict__ dvals) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npos) return;
    int lo = 0, hi = nl - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tab[mid].pos_off <= p) lo = mid; else hi = mid - 1;
    }
    const LeafDiag t = tab[lo];
    const int i = (int)(p - t.pos_off);
    double v;
    if (i < t.nx) v = (dd ? dd[t.dd_off + i] : 0.0) + regP + t.reg[0];
    else if (i < t.nx + t.ny) v = -(regDy + t.reg[1]);
    else v = (nz ? nz[t.nz_off + (i - t.nx - t.ny)] : 0.0) - (regDz + t.reg[2]);
    dvals[p] = v;
}
void launch_build_dvals(int64_t npos, const LeafDiag* tab, int nl, const double* dd, const double* nz, double regP, double regDy,
                        double regDz, double* dvals, cudaStream_t s) {
    if (npos <= 0 || nl <= 0) return;
    build_dvals_kernel<<<kr_cdiv(npos, 256), 256, 0, s>>>(npos, tab, nl, dd, nz, regP, regDy, regDz, dvals);
    count_launch("build_dvals_kernel");
}

}  // namespace pipsgpu
