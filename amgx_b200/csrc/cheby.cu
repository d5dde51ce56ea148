// cheby.cu -- CHEBYSHEV and CHEBYSHEV_POLY smoothers: polynomial smoothers made of the hot path's SpMV / residual / axpy kernels.
//   Chebyshev_Solver      src/solvers/cheb_solver.cu:64-370      (preconditioned three-term Chebyshev recurrence)
//   ChebyshevPolySolver   src/solvers/chebyshev_poly.cu:60-330   (x += tau_i (b - A x), i = 0 .. order-1, damped Chebyshev roots)
// Spectrum estimates: chebyshev_lambda_estimate_mode 2 (max row sum of |a_ij|, or 0.9 behind a preconditioner) and 3 (user supplied
// cheby_max_lambda / cheby_min_lambda behind a preconditioner); modes 0 and 1 need the reference's Lanczos eigensolver, which is
// outside this engine (AMGX_RC_NOT_IMPLEMENTED).
// Like the reference neither solver clears x when the caller says "x is zero": the residual is taken as b, but x itself is read /
// updated as it is.  Every shipped configuration runs them with presweeps = 0, where the cycle has zero-filled x beforehand.
#include "solvers.h"
#include "dist.h"
#include <cmath>

namespace amgxb {
namespace {

// lambda = max_i sum_j |a_ij|   (getLambdaEstimate + max_element, cheb_solver.cu:15-70)
double max_abs_row_sum(Solver &sv, Matrix &A, const ReduceCtx &red, ScalarBlock &sb, cudaStream_t s)
{
    if (A.bs() != 1) fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "Chebyshev smoothers: the row-sum eigenvalue estimate is implemented for scalar matrices");
    if (A.has_ext_diag) fatal(AMGX_RC_NOT_IMPLEMENTED, "Chebyshev smoothers with an external diagonal");
    DevVec d;
    l1_row_norms(A, d, s);
    vec_nrmmax(d.ptr(), d.prec, (size_t)A.n, red, S_TMP0, 1, s);
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    double v = sb.host[S_TMP0];
    if (A.dist) dist_allreduce_host(A, &v, 1, 2);
    (void)sv;
    return v;
}

class ChebyshevSolver : public Solver {
public:
    ChebyshevSolver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc) : Solver(cfg, scope, std::move(rsc))
    {
        std::string name, ns;
        cfg.get_scoped("preconditioner", scope, name, ns);
        mode_ = cfg.get_int("chebyshev_lambda_estimate_mode", scope);
        order_ = cfg.get_int("chebyshev_polynomial_order", scope);
        if (mode_ == 3) {
            user_max_ = cfg.get_double("cheby_max_lambda", scope);
            user_min_ = cfg.get_double("cheby_min_lambda", scope);
        }
        if (mode_ == 0 || mode_ == 1)
            fatal(AMGX_RC_NOT_IMPLEMENTED, "chebyshev_lambda_estimate_mode 0 and 1 use the Lanczos eigensolver, which this engine does not provide (use 2 or 3)");
        if (mode_ != 2 && mode_ != 3) fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "Not supported chebyshev_lambda_estimate_mode.");
        if (name != "NOSOLVER") precond_ = Solver::allocate(cfg, scope, "preconditioner", rsc_);
    }
    bool is_coloring_needed() const override { return precond_ ? precond_->is_coloring_needed() : false; }
    void print_grid_stats() override { if (precond_) precond_->print_grid_stats(); }
    double lambda_max() const { return lmax_; }
    double lambda_min() const { return lmin_; }

protected:
    bool is_residual_needed() const override { return true; }
    void solver_setup(bool reuse) override
    {
        if (precond_) precond_->setup(*A_, reuse);
        if (!precond_) {                       // modes 2 and 3 coincide without a preconditioner (cheb_solver.cu:186-213)
            lmax_ = max_abs_row_sum(*this, *A_, red_ctx(), sb_, stream());
            lmin_ = lmax_ * 0.125;
        } else if (mode_ == 2) {               // "this preconditioner would be good enough to reduce spectrum to the largest eigen value = 1.0"
            lmax_ = 0.9;
            lmin_ = lmax_ * 0.125;
        } else {
            lmax_ = user_max_;
            lmin_ = user_min_;
        }
        const size_t N = (size_t)A_->n_cols * A_->by;
        p_.resize(N, A_->vec_prec);
        z_.resize(N, A_->vec_prec);
        p_.zero(stream());
        z_.zero(stream());
    }
    void precondition()      // z = M^-1 r (zero initial guess) or a copy
    {
        if (precond_) precond_->solve(r_, z_, true);
        else vec_copy(z_.ptr(), r_.ptr(), r_.prec, vec_len(), stream());
    }
    void solve_init(DevVec &, DevVec &, bool) override
    {
        precondition();
        vec_copy(p_.ptr(), z_.ptr(), z_.prec, vec_len(), stream());
        gamma_ = 0.;
        beta_ = 0.;
        first_iter_ = 0;
    }
    Status solve_iteration(DevVec &b, DevVec &x, bool) override
    {
        cudaStream_t s = stream();
        const size_t n = vec_len();
        const double a = (lmax_ + lmin_) / 2, c = (lmax_ - lmin_) / 2;
        for (int i = 0; i < order_; i++) {
            precondition();
            if (first_iter_ == 0) {
                gamma_ = 1. / a;
                first_iter_ = 1;
            } else {
                beta_ = c * c * gamma_ * gamma_ / 4.;
                if (gamma_ != 0.0 && (a - (beta_ / gamma_)) != 0.0) gamma_ = 1. / (a - beta_ / gamma_);
                vec_axpby(z_.ptr(), p_.ptr(), p_.ptr(), x.prec, n, 1.0, beta_, s);
            }
            vec_axpy(p_.ptr(), x.ptr(), x.prec, n, gamma_, s);
            compute_residual(b, x);
        }
        Status st = ST_NOT_CONVERGED;
        if (monitor_convergence_ && is_done(st = compute_norm_and_converged())) return st;
        return monitor_convergence_ ? ST_NOT_CONVERGED : ST_CONVERGED;
    }

    std::unique_ptr<Solver> precond_;
    int mode_ = 0, order_ = 5, first_iter_ = 0;
    double user_max_ = 1.0, user_min_ = 0.125, lmax_ = 0, lmin_ = 0, gamma_ = 0, beta_ = 0;
    DevVec p_, z_;
};

class ChebyshevPolySolver : public Solver {
public:
    ChebyshevPolySolver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc) : Solver(cfg, scope, std::move(rsc))
    {
        order_ = cfg.get_int("chebyshev_polynomial_order", scope);
        order_ = std::min(10, std::max(order_, 1));
        tau_.resize(order_);
    }

protected:
    void solver_setup(bool) override
    {
        if (A_->bs() != 1) fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "Unsupported block size for BlockJacobi_Solver");
        const double lambda = max_abs_row_sum(*this, *A_, red_ctx(), sb_, stream());
        // magicDampBeta / magicDamp (chebyshev_poly.cu:63-74)
        const double beta = M_PI / (4 * (double)order_ + 2);
        for (int i = 0; i < order_; i++) {
            const double cb = std::cos(beta), c2 = std::cos(beta * (2 * i + 1)), sb2 = std::sin(beta);
            tau_[i] = (cb * cb / (c2 * c2 - sb2 * sb2)) / lambda;
        }
        y_.resize((size_t)A_->n_cols * A_->by, A_->vec_prec);
        y_.zero(stream());
    }
    Status solve_iteration(DevVec &b, DevVec &x, bool) override
    {
        cudaStream_t s = stream();
        for (int i = 0; i < order_; i++) {
            // y = b - A x ; x = x + tau_i y  -- the reference forms A x and x + tau (b - y) in two passes; the subtraction is the same
            // single rounding either way, so the residual epilogue of the SpMV kernel reproduces it
            dist_exchange_halo(*A_, x, s);
            CsrOpArgs g;
            g.x = x.ptr();
            g.b = b.ptr();
            g.y = y_.ptr();
            matrix_apply(*A_, EPI_RESID, g, s);
            vec_axpy(y_.ptr(), x.ptr(), x.prec, vec_len(), tau_[i], s);
        }
        return converged(b, x);
    }
    int order_ = 5;
    std::vector<double> tau_;
    DevVec y_;
};

}  // namespace

std::unique_ptr<Solver> make_chebyshev_solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc)
{
    return std::unique_ptr<Solver>(new ChebyshevSolver(cfg, scope, std::move(rsc)));
}
std::unique_ptr<Solver> make_chebyshev_poly_solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc)
{
    return std::unique_ptr<Solver>(new ChebyshevPolySolver(cfg, scope, std::move(rsc)));
}

}  // namespace amgxb
