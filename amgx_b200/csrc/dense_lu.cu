// dense_lu.cu -- DENSE_LU_SOLVER: direct solve of the coarsest level (the reference's default coarse_solver).
//   DenseLUSolver::solver_setup / solve_iteration   src/solvers/dense_lu_solver.cu:745-900, 905-985
//   csr_to_dense_kernel                             src/solvers/dense_lu_solver.cu:54-130
// The reference copies the coarsest block-CSR matrix into a column-major dense array and calls cuSOLVER getrf / getrs
// (LU with partial pivoting).  cuSOLVER is a library call on a matrix of at most dense_lu_num_rows (default 128)
// rows; here the same factorisation is two small single-CTA kernels:
//   * factor: right-looking LU, pivot = first entry of largest magnitude in the column (idamax), rows swapped across
//     the whole matrix, multipliers scaled by the reciprocal pivot (dgetf2), rank-1 update with one FMA per entry;
//   * solve:  x = rhs, row interchanges, unit-lower forward substitution, upper backward substitution, one column per
//     step, LU staged in shared memory when it fits (n <= 158 in fp64), else read from L2.
// The arithmetic order is sequential per entry (k ascending), so the CPU restatement reproduces it bit for bit;
// against cuSOLVER's blocked getrf the results agree to rounding (backward stable either way).
// Row-partitioned matrices: like the reference's default each rank factors the diagonal block it owns (block Jacobi over the partitions);
// halo values enter only through the right-hand side when the initial guess is not zero.  Limit: n = local rows * block_dim <= 8192.
#include "solvers.h"
#include "dist.h"

namespace amgxb {
namespace {

constexpr int LU_THREADS = 1024;
constexpr int LU_MAX_N = 8192;      // 512 MB of factors in fp64; the reference's own unit test factors 4096 rows (src/tests/dense_lu.cu:232)

template <class MatT, class T>
__global__ void csr_to_dense_kernel(int n_rows, int bdim, const int *__restrict__ rp, const int *__restrict__ ci, const MatT *__restrict__ va, int nnz,
                                    int has_ext_diag, T *dense, int lda)
{
    const int bs = bdim * bdim;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += gridDim.x * blockDim.x) {
        for (int k = rp[i]; k < rp[i + 1]; k++) {
            const int j = ci[k];
            if (j >= n_rows) continue;     // row-partitioned matrix: only the diagonal block this rank owns is factored (dense_lu_solver.cu:893-905)
            for (int r = 0; r < bdim; r++)
                for (int c = 0; c < bdim; c++) dense[(size_t)(i * bdim + r) + (size_t)(j * bdim + c) * lda] = (T)va[(size_t)k * bs + r * bdim + c];
        }
        if (has_ext_diag)
            for (int r = 0; r < bdim; r++)
                for (int c = 0; c < bdim; c++) dense[(size_t)(i * bdim + r) + (size_t)(i * bdim + c) * lda] = (T)va[(size_t)(nnz + i) * bs + r * bdim + c];
    }
}

// row-partitioned matrix, non-zero initial guess: new_rhs = b - A_halo x (distributed_rhs_mod, dense_lu_solver.cu:148-490): only the
// entries whose column lives on another rank contribute
template <class MatT, class T>
__global__ void halo_rhs_kernel(int n_rows, int bdim, const int *__restrict__ rp, const int *__restrict__ ci, const MatT *__restrict__ va,
                                const T *__restrict__ x, const T *__restrict__ b, T *out)
{
    const int bs = bdim * bdim;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_rows * bdim; t += gridDim.x * blockDim.x) {
        const int i = t / bdim, r = t % bdim;
        T acc = b[t];
        for (int k = rp[i]; k < rp[i + 1]; k++) {
            const int j = ci[k];
            if (j < n_rows) continue;
            for (int c = 0; c < bdim; c++) acc -= (T)va[(size_t)k * bs + r * bdim + c] * x[(size_t)j * bdim + c];
        }
        out[t] = acc;
    }
}

// one CTA; a is column-major n x n with leading dimension lda
template <class T> __global__ void __launch_bounds__(LU_THREADS) lu_factor_kernel(int n, T *a, int lda, int *ipiv, int *info)
{
    __shared__ int s_p;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int k = 0; k < n; k++) {
        if (tid == 0) {
            int p = k;
            T best = fabs(a[(size_t)k + (size_t)k * lda]);
            for (int i = k + 1; i < n; i++) {
                const T v = fabs(a[(size_t)i + (size_t)k * lda]);
                if (v > best) { best = v; p = i; }
            }
            ipiv[k] = p;
            s_p = p;
            if (best == (T)0 && *info == 0) *info = k + 1;
        }
        __syncthreads();
        const int p = s_p;
        if (p != k)
            for (int j = tid; j < n; j += nt) {
                const T t = a[(size_t)k + (size_t)j * lda];
                a[(size_t)k + (size_t)j * lda] = a[(size_t)p + (size_t)j * lda];
                a[(size_t)p + (size_t)j * lda] = t;
            }
        __syncthreads();
        const T piv = a[(size_t)k + (size_t)k * lda];
        if (piv != (T)0) {
            const T r = (T)1 / piv;
            for (int i = k + 1 + tid; i < n; i += nt) a[(size_t)i + (size_t)k * lda] *= r;
        }
        __syncthreads();
        const int m = n - k - 1;
        for (int idx = tid; idx < m * m; idx += nt) {
            const int i = k + 1 + idx % m, j = k + 1 + idx / m;
            a[(size_t)i + (size_t)j * lda] = fma(-a[(size_t)i + (size_t)k * lda], a[(size_t)k + (size_t)j * lda], a[(size_t)i + (size_t)j * lda]);
        }
        __syncthreads();
    }
}

// x = A^-1 rhs with the factors of lu_factor_kernel.  IN_SMEM: the factors are first staged into shared memory.
template <class T, bool IN_SMEM>
__global__ void __launch_bounds__(LU_THREADS) lu_solve_kernel(int n, const T *__restrict__ lu_g, int lda, const int *__restrict__ ipiv, const T *__restrict__ rhs, T *out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *x = reinterpret_cast<T *>(smem_raw);               // n entries
    T *lu_s = x + ((n + 1) & ~1);                         // n*n entries when IN_SMEM
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < n; i += nt) x[i] = rhs[i];
    if (IN_SMEM)
        for (int j = 0; j < n; j++)
            for (int i = tid; i < n; i += nt) lu_s[i + j * n] = lu_g[(size_t)i + (size_t)j * lda];
    __syncthreads();
    const T *lu = IN_SMEM ? lu_s : lu_g;
    const int ld = IN_SMEM ? n : lda;
    if (tid == 0)
        for (int k = 0; k < n; k++) {
            const int p = ipiv[k];
            if (p != k) { const T t = x[k]; x[k] = x[p]; x[p] = t; }
        }
    __syncthreads();
    for (int k = 0; k < n - 1; k++) {                     // L y = P b (unit diagonal)
        const T xk = x[k];
        for (int i = k + 1 + tid; i < n; i += nt) x[i] = fma(-lu[(size_t)i + (size_t)k * ld], xk, x[i]);
        __syncthreads();
    }
    for (int k = n - 1; k >= 0; k--) {                    // U x = y
        if (tid == 0) x[k] = x[k] / lu[(size_t)k + (size_t)k * ld];
        __syncthreads();
        const T xk = x[k];
        for (int i = tid; i < k; i += nt) x[i] = fma(-lu[(size_t)i + (size_t)k * ld], xk, x[i]);
        __syncthreads();
    }
    for (int i = tid; i < n; i += nt) out[i] = x[i];
}

}  // namespace

class DenseLUSolver : public Solver {
public:
    DenseLUSolver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc) : Solver(cfg, scope, std::move(rsc))
    {
        set_max_iters(1);   // "Make sure we don't run more than 1 iteration." (dense_lu_solver.cu:660-661)
    }

protected:
    void solver_setup(bool) override
    {
        Matrix &A = *A_;
        // A.dist: the reference factors the diagonal block each rank owns -- block Jacobi over the partitions, halo values only enter
        // through the right-hand side (dense_lu_solver.cu:893-912).  Its all-gathered exact solve (exact_coarse_solve = 1, CLASSICAL only,
        // dense_lu_solver.cu:667-669) is not provided as such; classical AMG on a partitioned matrix keeps its coarse levels whole on
        // every rank (classical.cu: distribute_finest), so its coarsest solve is exact without it.
        if (A.dist && cfg_->get_int("exact_coarse_solve", scope_) != 0 && cfg_->get_string("algorithm", scope_) == "CLASSICAL")
            fatal(AMGX_RC_NOT_IMPLEMENTED, "DENSE_LU_SOLVER with exact_coarse_solve=1 on a distributed matrix");
        if (A.bx != A.by) fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "DENSE_LU_SOLVER needs square blocks");
        n_ = A.n * A.bx;
        if (n_ > LU_MAX_N) fatal(AMGX_RC_NOT_IMPLEMENTED, "DENSE_LU_SOLVER: coarsest level has more than 8192 rows; lower dense_lu_max_rows or use coarse_solver=NOSOLVER");
        cudaStream_t s = stream();
        dense_.resize((size_t)std::max(n_, 1) * std::max(n_, 1), A.vec_prec);
        dense_.zero(s);
        ipiv_.resize((size_t)std::max(n_, 1) + 1);
        ipiv_.zero(s);
        if (n_ == 0) return;
        AMGXB_DISPATCH(A.mat_prec, A.vec_prec, {
            csr_to_dense_kernel<MatT, VecT><<<std::max(1, ceil_div(A.n, 128)), 128, 0, s>>>(A.n, A.bx, A.row_ptr.ptr(), A.col_idx.ptr(), A.values.as<MatT>(), A.nnz,
                                                                                             A.has_ext_diag ? 1 : 0, dense_.as<VecT>(), n_);
            lu_factor_kernel<VecT><<<1, LU_THREADS, 0, s>>>(n_, dense_.as<VecT>(), n_, ipiv_.ptr(), ipiv_.ptr() + n_);
        });
        count_launch(2);
        AMGXB_LAUNCH_CHECK();
        {   // a zero pivot is a setup error, as in the reference (dense_lu_solver.cu:542-560), not inf / NaN at solve time
            int info = 0;
            AMGXB_CUDA_CHECK(cudaMemcpyAsync(&info, ipiv_.ptr() + n_, sizeof(int), cudaMemcpyDeviceToHost, s));
            AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
            if (info != 0) fatal(AMGX_RC_INTERNAL, "Dense LU factorization failed due to a singular matrix");
        }
        const size_t esz = prec_size(A.vec_prec);
        smem_small_ = (size_t)((n_ + 1) & ~1) * esz;
        smem_full_ = smem_small_ + (size_t)n_ * n_ * esz;
        in_smem_ = smem_full_ <= (size_t)200 * 1024;
        AMGXB_DISPATCH_VEC(A.vec_prec, {
            if (in_smem_) AMGXB_CUDA_CHECK(cudaFuncSetAttribute(lu_solve_kernel<VecT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_full_));
            else if (smem_small_ > (size_t)48 * 1024)
                AMGXB_CUDA_CHECK(cudaFuncSetAttribute(lu_solve_kernel<VecT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_small_));
        });
    }

    Status solve_iteration(DevVec &b, DevVec &x, bool xIsZero) override
    {
        if (n_ == 0) return ST_CONVERGED;
        cudaStream_t s = stream();
        const void *rhs = b.ptr();
        if (A_->dist && !xIsZero) {
            Matrix &A = *A_;
            dist_exchange_halo(A, x, s);
            dist_wait_halo(A, s);
            if (rhs_mod_.n != (size_t)n_) rhs_mod_.resize((size_t)n_, A.vec_prec);
            AMGXB_DISPATCH(A.mat_prec, A.vec_prec, {
                halo_rhs_kernel<MatT, VecT><<<std::max(1, ceil_div(n_, 128)), 128, 0, s>>>(A.n, A.bx, A.row_ptr.ptr(), A.col_idx.ptr(), A.values.as<MatT>(),
                                                                                           x.as<VecT>(), b.as<VecT>(), rhs_mod_.as<VecT>());
            });
            count_launch();
            rhs = rhs_mod_.ptr();
        }
        AMGXB_DISPATCH_VEC(A_->vec_prec, {
            if (in_smem_) lu_solve_kernel<VecT, true><<<1, LU_THREADS, smem_full_, s>>>(n_, dense_.as<VecT>(), n_, ipiv_.ptr(), (const VecT *)rhs, x.as<VecT>());
            else lu_solve_kernel<VecT, false><<<1, LU_THREADS, smem_small_, s>>>(n_, dense_.as<VecT>(), n_, ipiv_.ptr(), (const VecT *)rhs, x.as<VecT>());
        });
        count_launch();
        AMGXB_LAUNCH_CHECK();
        return ST_CONVERGED;   // direct solver always converges (dense_lu_solver.cu:984)
    }

    int n_ = 0;
    DevVec dense_;          // n x n column-major LU factors (vector precision)
    DevVec rhs_mod_;        // distributed, non-zero initial guess: b - A_halo x
    DevBuf<int> ipiv_;      // n pivots + 1 info word
    size_t smem_small_ = 0, smem_full_ = 0;
    bool in_smem_ = false;
};

std::unique_ptr<Solver> make_dense_lu_solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc)
{
    return std::unique_ptr<Solver>(new DenseLUSolver(cfg, scope, std::move(rsc)));
}

}  // namespace amgxb
