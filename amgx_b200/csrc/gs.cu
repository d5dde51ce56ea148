// gs.cu -- MULTICOLOR_GS smoother (scalar matrices): weighted Gauss-Seidel swept colour by colour, in place.
//   kernel   multicolorGSSmoothCsrKernel_nPerRow   src/solvers/multicolor_gauss_seidel_solver.cu:496-551
//   host     smooth_1x1 / solve_iteration          :971-1145, 1412-1545   (symmetric_GS: colours ascending, then descending)
// For every row i of the colour:  x_i <- x_i + w * (b_i - sum_j a_ij x_j) / a_ii  (the sum includes j = i and reads the current x:
// rows of earlier colours are already updated, rows of one colour do not couple).  The reference runs N = 4 or 32 lanes per row
// (4 by default; 32 when nnz/rows > 20 or nnz/colours < 500000, :1004-1013): lane l accumulates -a x over entries l, l+N, ... and
// lane 0 of the shuffle-down tree adds b, divides and updates.  The same decomposition (same N, same tree) is kept here so that a
// row's sum associates exactly like the reference's; only the grid is sized for 148 SMs instead of 1024 CTAs of 672 threads.
// The colouring comes from the matrix (coloring.cu, MIN_MAX; or AMGX_matrix_attach_coloring) exactly as for MULTICOLOR_DILU.
#include "solvers.h"
#include "dist.h"

namespace amgxb {

void color_matrix(Matrix &A, const std::string &scheme, double max_uncolored_fraction, cudaStream_t s);   // coloring.cu

namespace {

template <class T> __device__ __forceinline__ T gs_guard(T d);
template <> __device__ __forceinline__ double gs_guard<double>(double d) { return fabs(d) < 1e-12 ? copysign(1e-12, d) : d; }
template <> __device__ __forceinline__ float gs_guard<float>(float d) { return fabs((double)d) < 1e-7 ? copysignf((float)1e-7, d) : d; }

template <class MatT, class VecT, int NPR>
__global__ void __launch_bounds__(256) gs_color_sweep(const int *__restrict__ rp, const int *__restrict__ ci, const int *__restrict__ diag,
                                                      const MatT *__restrict__ va, const VecT *__restrict__ b, VecT *x, VecT weight,
                                                      const int *__restrict__ rows, int nrows)
{
    const int l = threadIdx.x % NPR;
    const int rows_per_grid = gridDim.x * (blockDim.x / NPR);
    for (int it = blockIdx.x * (blockDim.x / NPR) + threadIdx.x / NPR; __any_sync(0xffffffffu, it < nrows); it += rows_per_grid) {
        const bool act = it < nrows;
        const int i = act ? rows[it] : 0;
        VecT acc = 0;
        if (act) {
            const int k1 = rp[i + 1];
            for (int k = rp[i] + l; k < k1; k += NPR) acc -= (VecT)va[k] * x[ci[k]];
        }
#pragma unroll
        for (int m = NPR / 2; m > 0; m >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, m);
        if (act && l == 0) {
            const int d = diag[i];
            const VecT dt = d >= 0 ? (VecT)va[d] : (VecT)0;
            const MatT dia = (MatT)gs_guard<VecT>(dt);
            acc += b[i];
            acc /= dia;
            x[i] = x[i] + weight * acc;
        }
    }
}

class MulticolorGSSolver : public Solver {
public:
    MulticolorGSSolver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc) : Solver(cfg, scope, std::move(rsc))
    {
        weight_ = cfg.get_double("relaxation_factor", scope);
        sym_ = cfg.get_int("symmetric_GS", scope) == 1;
        if (weight_ == 0) {
            weight_ = 1.;
            amgx_printf("Warning, setting weight to 1 instead of estimating largest_eigen_value in Multicolor GaussSeidel smoother\n");
        }
        scheme_ = cfg.get_string("matrix_coloring_scheme", scope);
        if (scheme_ != "MIN_MAX" && scheme_ != "PARALLEL_GREEDY")
            fatal(AMGX_RC_BAD_CONFIGURATION, "matrix_coloring_scheme '" + scheme_ + "' is not supported by this engine (MIN_MAX, PARALLEL_GREEDY, or AMGX_matrix_attach_coloring)");
        if (cfg.get_int("coloring_level", scope) < 1)
            fatal(AMGX_RC_NOT_IMPLEMENTED, "Matrix must be colored to use multicolor gauss-seidel solver. Try setting: coloring_level=1 in the configuration file");
        if (cfg.get_int("coloring_level", scope) != 1) fatal(AMGX_RC_BAD_CONFIGURATION, "MULTICOLOR_GS: coloring_level must be 1");
        // reorder_cols_by_color / insert_diag_while_reordering (src/matrix.cu:749-812): the reference sorts the entries of every row by the colour of
        // their column so that its sweeps can split a row into "earlier colours | later colours" without reading the colour array.  A memory
        // layout, not an algorithm: the sweeps here look up the colour of every column (or work on their own colour-sorted copy), whatever the
        // caller's entry order is, so the two
        // switches are accepted and change nothing (the caller's matrix is not permuted; sums differ from the reference's by their rounding only).
        if (cfg.get_int("use_bsrxmv", scope) != 0) fatal(AMGX_RC_NOT_IMPLEMENTED, "MULTICOLOR_GS with use_bsrxmv=1");
        uncolored_fraction_ = cfg.get_int("determinism_flag", "default") ? 0.0 : cfg.get_double("max_uncolored_percentage", scope);
    }
    bool is_coloring_needed() const override { return true; }
    void smooth(DevVec &b, DevVec &x, bool xIsZero, int sweeps, const SmoothFuse *fuse, bool input_in_alt = false) override
    {
        if (input_in_alt || (fuse && (fuse->agg || fuse->dot_b_x))) fatal(AMGX_RC_INTERNAL, "Gauss-Seidel does not support fused sweeps");
        for (int it = 0; it < sweeps; it++) sweep(b, x, xIsZero && it == 0);
    }

protected:
    void solver_setup(bool) override
    {
        Matrix &A = *A_;
        if (A.bs() != 1) fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "MULTICOLOR_GS: scalar matrices only in this engine");
        if (A.has_ext_diag) fatal(AMGX_RC_NOT_IMPLEMENTED, "MULTICOLOR_GS with an external diagonal");
        if (A.dist) fatal(AMGX_RC_NOT_IMPLEMENTED, "MULTICOLOR_GS on a distributed matrix");
        if (A.num_colors == 0) color_matrix(A, scheme_, uncolored_fraction_, stream());
        // KernelMethod::DEFAULT selection (multicolor_gauss_seidel_solver.cu:1004-1013)
        lanes_ = 4;
        if (A.n > 0 && A.nnz / A.n > 20) lanes_ = 32;
        if (A.num_colors > 0 && A.nnz / A.num_colors < 500000) lanes_ = 32;
    }

    template <class MatT, class VecT, int NPR> void launch_color(DevVec &b, DevVec &x, int off, int cnt)
    {
        Matrix &A = *A_;
        const int grid = std::min(B200_SMS * 8, ceil_div(cnt, 256 / NPR));
        gs_color_sweep<MatT, VecT, NPR><<<grid, 256, 0, stream()>>>(A.row_ptr.ptr(), A.col_idx.ptr(), A.diag_idx.ptr(), A.values.as<MatT>(), b.as<VecT>(),
                                                                   x.as<VecT>(), (VecT)weight_, A.sorted_rows_by_color.ptr() + off, cnt);
        count_launch();
    }

    void sweep(DevVec &b, DevVec &x, bool xIsZero)
    {
        Matrix &A = *A_;
        if (xIsZero) x.zero(stream());
        const int nc = A.num_colors;
        auto color = [&](int c) {
            const int off = A.color_offsets[c], cnt = A.color_offsets[c + 1] - off;
            if (cnt == 0) return;
            AMGXB_DISPATCH(A.mat_prec, A.vec_prec, {
                if (lanes_ == 4) launch_color<MatT, VecT, 4>(b, x, off, cnt);
                else launch_color<MatT, VecT, 32>(b, x, off, cnt);
            });
        };
        for (int c = 0; c < nc; c++) color(c);
        if (sym_)
            for (int c = nc - 1; c >= 0; c--) color(c);
        AMGXB_LAUNCH_CHECK();
    }

    Status solve_iteration(DevVec &b, DevVec &x, bool xIsZero) override
    {
        sweep(b, x, xIsZero);
        return converged(b, x);
    }

    double weight_ = 0.9, uncolored_fraction_ = 0.15;
    std::string scheme_ = "MIN_MAX";
    bool sym_ = false;
    int lanes_ = 4;
};

}  // namespace

std::unique_ptr<Solver> make_gs_solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc)
{
    return std::unique_ptr<Solver>(new MulticolorGSSolver(cfg, scope, std::move(rsc)));
}

}  // namespace amgxb
