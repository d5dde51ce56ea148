// krylov_extra.cu -- CG, PCGF, PBICGSTAB and GMRES: the remaining Krylov drivers of the reference that reuse the kernels of the
// hot path.  Straight restatements of
//   CG_Solver          src/solvers/cg_solver.cu:36-103
//   PCGF_Solver        src/solvers/pcgf_solver.cu:76-175      (flexible PCG, beta = <z, r_new - r_old> / <r, z>)
//   PBiCGStab_Solver   src/solvers/pbicgstab_solver.cu:60-262
//   GMRES_Solver       src/solvers/gmres_solver.cu:23-395     (right-preconditioned GMRES(m), one Z vector)
// with the recurrence scalars on the host, as in the reference (one synchronisation per inner product); the PCG of the headline
// configurations keeps its device-scalar / fused / CUDA-graph form in solvers.cu.  Vector updates use the reference's expression
// order (x*a + y*b [+ z*c]).
#include "solvers.h"
#include "dist.h"
#include <cmath>

namespace amgxb {
namespace {

class KrylovBase : public Solver {
public:
    KrylovBase(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc, bool with_precond = true) : Solver(cfg, scope, std::move(rsc))
    {
        std::string name, ns;
        cfg.get_scoped("preconditioner", scope, name, ns);
        if (with_precond && name != "NOSOLVER") precond_ = Solver::allocate(cfg, scope, "preconditioner", rsc_);
    }
    void print_grid_stats() override { if (precond_) precond_->print_grid_stats(); }

protected:
    bool is_residual_needed() const override { return true; }
    // <x, y> over the owned rows of all ranks, on the host
    double hdot(const DevVec &x, const DevVec &y)
    {
        cudaStream_t s = stream();
        ReduceCtx red = red_ctx();
        vec_dot(x.ptr(), y.ptr(), x.prec, vec_len(), red, FIN_STORE, S_TMP0, 0, s);
        if (A_->dist) dist_allreduce_scalar_fin(*A_, red, S_TMP0, FIN_STORE, s);
        double h = 0;
        AMGXB_CUDA_CHECK(cudaMemcpyAsync(&h, red.scal + S_TMP0, sizeof(double), cudaMemcpyDeviceToHost, s));
        AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
        return h;
    }
    void apply_A(DevVec &x, DevVec &y)
    {
        cudaStream_t s = stream();
        dist_exchange_halo(*A_, x, s);
        CsrOpArgs g;
        g.x = x.ptr();
        g.y = y.ptr();
        matrix_apply(*A_, EPI_SPMV, g, s);
    }
    void precondition(DevVec &in, DevVec &out)   // out = M^-1 in (zero initial guess) or a copy
    {
        if (precond_) precond_->solve(in, out, true);
        else vec_copy(out.ptr(), in.ptr(), in.prec, vec_len(), stream());
    }
    void alloc(DevVec &v) { v.resize((size_t)A_->n_cols * A_->by, A_->vec_prec); v.zero(stream()); }
    Status norm_of_and_check(const DevVec &v)     // compute_norm_and_converged(v, nrm)
    {
        std::vector<double> nv;
        compute_norm_of(v, nv);
        return conv_.update_and_check(nv, nrm_ini_);
    }
    std::unique_ptr<Solver> precond_;
};

class CGSolver : public KrylovBase {
public:
    // CG_Solver never reads the "preconditioner" parameter (cg_solver.cu): CG_DILU.json runs plain CG in the reference too
    CGSolver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc) : KrylovBase(cfg, scope, std::move(rsc), false) {}
protected:
    void solver_setup(bool) override { alloc(p_); alloc(Ap_); }
    void solve_init(DevVec &, DevVec &, bool) override
    {
        vec_copy(p_.ptr(), r_.ptr(), r_.prec, vec_len(), stream());
        rr_ = hdot(r_, r_);
    }
    Status solve_iteration(DevVec &b, DevVec &x, bool) override
    {
        cudaStream_t s = stream();
        const size_t n = vec_len();
        apply_A(p_, Ap_);
        const double alpha = rr_ / hdot(Ap_, p_);
        vec_axpy(p_.ptr(), x.ptr(), x.prec, n, alpha, s);
        vec_axpy(Ap_.ptr(), r_.ptr(), x.prec, n, -alpha, s);
        Status st = ST_NOT_CONVERGED;
        if (monitor_convergence_ && is_done(st = compute_norm_and_converged())) return st;
        if (is_last_iter()) return monitor_convergence_ ? ST_NOT_CONVERGED : ST_CONVERGED;
        const double rr_old = rr_;
        rr_ = hdot(r_, r_);
        vec_axpby(r_.ptr(), p_.ptr(), p_.ptr(), x.prec, n, 1.0, rr_ / rr_old, s);
        return monitor_convergence_ ? ST_NOT_CONVERGED : ST_CONVERGED;
    }
    DevVec p_, Ap_;
    double rr_ = 0;
};

class PCGFSolver : public KrylovBase {
public:
    using KrylovBase::KrylovBase;
protected:
    void solver_setup(bool reuse) override
    {
        alloc(p_); alloc(z_); alloc(Ap_); alloc(d_);
        if (precond_) precond_->setup(*A_, reuse);
    }
    void solve_init(DevVec &, DevVec &, bool) override
    {
        precondition(r_, z_);
        vec_copy(p_.ptr(), z_.ptr(), z_.prec, vec_len(), stream());
    }
    Status solve_iteration(DevVec &b, DevVec &x, bool) override
    {
        cudaStream_t s = stream();
        const size_t n = vec_len();
        apply_A(p_, Ap_);
        const double rz = hdot(r_, z_);
        const double alpha = rz / hdot(Ap_, p_);
        vec_axpy(p_.ptr(), x.ptr(), x.prec, n, alpha, s);
        vec_copy(d_.ptr(), r_.ptr(), x.prec, n, s);
        vec_axpy(Ap_.ptr(), r_.ptr(), x.prec, n, -alpha, s);
        Status st = ST_NOT_CONVERGED;
        if (monitor_convergence_ && is_done(st = compute_norm_and_converged())) return st;
        if (is_last_iter()) return monitor_convergence_ ? ST_NOT_CONVERGED : ST_CONVERGED;
        vec_axpby(r_.ptr(), d_.ptr(), d_.ptr(), x.prec, n, 1.0, -1.0, s);     // d = r_new - r_old
        precondition(r_, z_);
        const double beta = hdot(z_, d_) / rz;
        vec_axpby(z_.ptr(), p_.ptr(), p_.ptr(), x.prec, n, 1.0, beta, s);
        return monitor_convergence_ ? ST_NOT_CONVERGED : ST_CONVERGED;
    }
    DevVec p_, z_, Ap_, d_;
};

class PBiCGStabSolver : public KrylovBase {
public:
    using KrylovBase::KrylovBase;
protected:
    void solver_setup(bool reuse) override
    {
        alloc(p_); alloc(Mp_); alloc(s_); alloc(Ms_); alloc(t_); alloc(v_); alloc(rt_);
        if (precond_) precond_->setup(*A_, reuse);
    }
    void solve_init(DevVec &, DevVec &, bool) override
    {
        vec_copy(rt_.ptr(), r_.ptr(), r_.prec, vec_len(), stream());
        rho_ = hdot(rt_, r_);
        vec_copy(p_.ptr(), r_.ptr(), r_.prec, vec_len(), stream());
    }
    Status solve_iteration(DevVec &b, DevVec &x, bool) override
    {
        cudaStream_t s = stream();
        const size_t n = vec_len();
        const Prec vp = x.prec;
        precondition(p_, Mp_);
        apply_A(Mp_, v_);
        double red = hdot(rt_, v_);
        const double alpha = (red != 0.0) ? rho_ / red : 0.0;
        vec_axpby(r_.ptr(), v_.ptr(), s_.ptr(), vp, n, 1.0, -alpha, s);        // s = r - alpha v
        Status st = ST_NOT_CONVERGED;
        if (monitor_convergence_ && is_done(st = norm_of_and_check(s_))) {     // early exit on ||s||
            vec_axpby(x.ptr(), Mp_.ptr(), x.ptr(), vp, n, 1.0, alpha, s);
            compute_residual(b, x);
            compute_norm();
            return st;
        }
        precondition(s_, Ms_);
        apply_A(Ms_, t_);
        red = hdot(t_, t_);
        double omega = hdot(t_, s_);
        omega = (red == 0.0) ? 0.0 : omega / red;
        vec_axpbypcz(x.ptr(), Mp_.ptr(), Ms_.ptr(), x.ptr(), vp, n, 1.0, alpha, omega, s);     // x += alpha Mp + omega Ms
        vec_axpby(s_.ptr(), t_.ptr(), r_.ptr(), vp, n, 1.0, -omega, s);                        // r = s - omega t
        if (monitor_convergence_ && is_done(st = compute_norm_and_converged())) return st;
        if (is_last_iter()) return monitor_convergence_ ? ST_NOT_CONVERGED : ST_CONVERGED;
        const double rho_new = hdot(rt_, r_);
        double beta = 0.0;
        if (rho_ != 0.0 && omega != 0.0) beta = (rho_new / rho_) * (alpha / omega);
        rho_ = rho_new;
        vec_axpbypcz(r_.ptr(), p_.ptr(), v_.ptr(), p_.ptr(), vp, n, 1.0, beta, -beta * omega, s);   // p = r + beta p - beta omega v
        return monitor_convergence_ ? ST_NOT_CONVERGED : ST_CONVERGED;
    }
    DevVec p_, Mp_, s_, Ms_, t_, v_, rt_;
    double rho_ = 0;
};

class GMRESSolver : public KrylovBase {
public:
    GMRESSolver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc) : KrylovBase(cfg, scope, std::move(rsc))
    {
        R_ = cfg.get_int("gmres_n_restart", scope);
        if (R_ < 1) fatal(AMGX_RC_BAD_CONFIGURATION, "gmres_n_restart must be >= 1");
        K_ = std::min(max_iters_, R_);
        if (norm_type_ != NORM_L2) fatal(AMGX_RC_NOT_SUPPORTED_TARGET, "GMRES only works with L2 norm. Other norms would require extra computations. ");
    }
protected:
    bool is_residual_needed() const override { return false; }
    void solver_setup(bool reuse) override
    {
        if (A_->by != 1 && !use_scalar_norm_)
            fatal(AMGX_RC_NOT_SUPPORTED_TARGET, "GMRES solver only works on block matrix if configuration parameter use_scalar_norm=1");
        if (precond_) precond_->setup(*A_, reuse);
        K_ = std::min(max_iters_, R_);
        V_.resize(K_ + 1);
        for (auto &v : V_) alloc(v);
        alloc(Z_);
        H_.assign((size_t)(K_ + 1) * K_, 0.0);
        s_.assign(K_ + 1, 0.0);
        cs_.assign(K_, 0.0);
        sn_.assign(K_, 0.0);
    }
    double &H(int i, int j) { return H_[(size_t)i + (size_t)j * (K_ + 1)]; }
    double hnrm2(const DevVec &v) { return std::sqrt(hdot(v, v)); }
    void start_cycle(DevVec &b, DevVec &x, double &beta)     // V0 = A x - b ; beta = ||V0||
    {
        apply_A(x, V_[0]);
        vec_axpy(b.ptr(), V_[0].ptr(), x.prec, vec_len(), -1.0, stream());
        beta = hnrm2(V_[0]);
    }
    void rotate(int i)    // PlaneRotation(H, cs, sn, s, i)  (gmres_solver.cu:190-206)
    {
        for (int k = 0; k < i; k++) {
            const double t = cs_[k] * H(k, i) + sn_[k] * H(k + 1, i);
            H(k + 1, i) = cs_[k] * H(k + 1, i) - sn_[k] * H(k, i);
            H(k, i) = t;
        }
        const double dx = H(i, i), dy = H(i + 1, i);
        if (dy < 0.0) { cs_[i] = 1.0; sn_[i] = 0.0; }
        else if (std::fabs(dy) > std::fabs(dx)) { const double t = dx / dy; sn_[i] = 1.0 / std::sqrt(1.0 + t * t); cs_[i] = t * sn_[i]; }
        else { const double t = dy / dx; cs_[i] = 1.0 / std::sqrt(1.0 + t * t); sn_[i] = t * cs_[i]; }
        const double t = cs_[i] * s_[i];
        s_[i + 1] = -sn_[i] * s_[i];
        s_[i] = t;
        H(i, i) = cs_[i] * H(i, i) + sn_[i] * H(i + 1, i);
        H(i + 1, i) = 0.0;
    }
    Status solve_one_iteration(DevVec &b, DevVec &x)        // max_iters == 1 (gmres_solver.cu:215-268)
    {
        cudaStream_t s = stream();
        const size_t n = vec_len();
        const Prec vp = x.prec;
        double beta;
        start_cycle(b, x, beta);
        vec_scal(V_[0].ptr(), vp, n, -1.0 / beta, s);
        std::fill(s_.begin(), s_.end(), 0.0);
        s_[0] = beta;
        precondition(V_[0], Z_);
        apply_A(Z_, V_[1]);
        H(0, 0) = hdot(V_[1], V_[0]);
        vec_axpy(V_[0].ptr(), V_[1].ptr(), vp, n, -H(0, 0), s);
        H(1, 0) = hnrm2(V_[1]);
        rotate(0);
        if (monitor_convergence_) nrm_.assign(1, std::fabs(s_[1]));
        s_[0] = s_[0] / H(0, 0);
        vec_axpy(Z_.ptr(), x.ptr(), vp, n, s_[0], s);
        return monitor_convergence_ ? converged() : ST_CONVERGED;
    }
    Status solve_iteration(DevVec &b, DevVec &x, bool) override
    {
        if (max_iters_ == 1) return solve_one_iteration(b, x);
        cudaStream_t s = stream();
        const size_t n = vec_len();
        const Prec vp = x.prec;
        Status conv_stat = ST_NOT_CONVERGED;
        const int i = curr_iter_ % R_;
        if (i == 0) {
            double beta;
            start_cycle(b, x, beta);
            if (monitor_convergence_) {
                nrm_.assign(1, beta);
                if (is_done(conv_stat = converged())) return conv_stat;
            }
            vec_scal(V_[0].ptr(), vp, n, -1.0 / beta, s);
            std::fill(s_.begin(), s_.end(), 0.0);
            s_[0] = beta;
        }
        precondition(V_[i], Z_);
        apply_A(Z_, V_[i + 1]);
        for (int k = 0; k <= i; k++) {        // modified Gram-Schmidt
            H(k, i) = hdot(V_[i + 1], V_[k]);
            vec_axpy(V_[k].ptr(), V_[i + 1].ptr(), vp, n, -H(k, i), s);
        }
        H(i + 1, i) = hnrm2(V_[i + 1]);
        vec_scal(V_[i + 1].ptr(), vp, n, 1.0 / H(i + 1, i), s);
        rotate(i);
        if (monitor_convergence_) {
            nrm_.assign(1, std::fabs(s_[i + 1]));
            conv_stat = converged();
        }
        if (i == R_ - 1 || is_last_iter() || is_done(conv_stat)) {
            for (int j = i; j >= 0; j--) {
                s_[j] = s_[j] / H(j, j);
                for (int k = j - 1; k >= 0; k--) s_[k] = s_[k] - H(k, j) * s_[j];
            }
            Z_.zero(s);
            for (int j = 0; j <= i; j++) vec_axpy(V_[j].ptr(), Z_.ptr(), vp, n, s_[j], s);
            precondition(Z_, V_[0]);           // M^-1 (sum_j y_j v_j), stored in V0
            vec_axpy(V_[0].ptr(), x.ptr(), vp, n, 1.0, s);
        }
        return monitor_convergence_ ? conv_stat : ST_CONVERGED;
    }
    int R_ = 20, K_ = 20;
    std::vector<DevVec> V_;
    DevVec Z_;
    std::vector<double> H_, s_, cs_, sn_;
};

}  // namespace

std::unique_ptr<Solver> make_cg_solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc) { return std::unique_ptr<Solver>(new CGSolver(cfg, scope, std::move(rsc))); }
std::unique_ptr<Solver> make_pcgf_solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc) { return std::unique_ptr<Solver>(new PCGFSolver(cfg, scope, std::move(rsc))); }
std::unique_ptr<Solver> make_gmres_solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc) { return std::unique_ptr<Solver>(new GMRESSolver(cfg, scope, std::move(rsc))); }
std::unique_ptr<Solver> make_pbicgstab_solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc)
{
    return std::unique_ptr<Solver>(new PBiCGStabSolver(cfg, scope, std::move(rsc)));
}

}  // namespace amgxb
