// fgmres.cu -- flexible GMRES(m) with modified Gram-Schmidt and host Givens rotations.
// Follows FGMRES_Solver (src/solvers/fgmres_solver.cu:215-569): same recurrences, same convergence estimate
// (|s[m+1]| for the scalar L2 norm), same restart / final triangular solve.  Differences by design:
//   * the MGS coefficients h(i,m) = <V_i, V_{m+1}> never travel to the host inside the chain: each dot leaves its
//     value in device memory, the following axpy reads it there; the whole Hessenberg column is mirrored to pinned
//     host memory and read after ONE synchronisation per iteration (the reference synchronises m+2 times);
//   * gmres_krylov_dim < min(max_iters, gmres_n_restart) selects the reference's truncated variant ("DQGMRES", fgmres_solver.cu:17-211,
//     284-298, 448-548): rings of krylov_dim + 2 V- and krylov_dim + 1 Z-vectors, Gram-Schmidt over the last krylov_dim + 1 vectors only,
//     x updated every iteration, the residual VECTOR updated by the recursion of :520-533 and its L2 norm monitored (one more host
//     synchronisation per iteration).  H is allocated once and never cleared, as m_H in the reference: after a restart the rotations
//     meet what the previous cycle left above the band.  (With krylov_dim >= min(max_iters, restart) but max_iters < restart the
//     reference also updates x every iteration; here that case stays on the standard path -- same iterates, one triangular solve.)
#include "solvers.h"
#include "dist.h"
#include <cmath>

namespace amgxb {

FGMRESSolver::FGMRESSolver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc) : Solver(cfg, scope, rsc)
{
    std::string name, ns;
    cfg.get_scoped("preconditioner", scope, name, ns);
    if (name != "NOSOLVER") precond_ = Solver::allocate(cfg, scope, "preconditioner", rsc);
    R_ = cfg.get_int("gmres_n_restart", scope);
    krylov_dim_ = std::min(max_iters_, R_);
    const int kp = cfg.get_int("gmres_krylov_dim", scope);
    if (kp > 0) krylov_dim_ = std::min(krylov_dim_, kp);
    if (R_ < 1) fatal(AMGX_RC_BAD_CONFIGURATION, "gmres_n_restart must be >= 1");
    H_.assign((size_t)(R_ + 2) * (R_ + 1), 0.0);
    s_.assign(R_ + 2, 0.0);
    cs_.assign(R_ + 1, 0.0);
    sn_.assign(R_ + 1, 0.0);
    gamma_.assign(R_ + 2, 0.0);
}

FGMRESSolver::~FGMRESSolver()
{
    if (hs_dev_) cudaFree(hs_dev_);
    if (hs_host_) cudaFreeHost(hs_host_);
}

void FGMRESSolver::solver_setup(bool reuse)
{
    if (precond_) precond_->setup(*A_, reuse);
    const int kmax = std::min(R_, max_iters_);
    trunc_ = krylov_dim_ < kmax;          // implies R_ > 1 and max_iters_ > 1
    use_scalar_L2_ = (A_->by == 1 || use_scalar_norm_) && norm_type_ == NORM_L2;
    if (monitor_convergence_ && !use_scalar_L2_)
        fatal(AMGX_RC_NOT_IMPLEMENTED, "FGMRES convergence monitoring supports the scalar L2 norm only");
    const size_t N = (size_t)A_->n_cols * A_->by;
    // standard: V(0..kmax), Z(0..kmax-1), index == iteration within the restart.  Truncated: rings (KrylovSubspaceBuffer, max_dimension = K + 1)
    V_.resize(trunc_ ? krylov_dim_ + 2 : kmax + 1);
    Z_.resize(trunc_ ? krylov_dim_ + 1 : kmax);
    // Krylov vectors are allocated when an iteration first needs them (gmres_n_restart = 100 in the shipped classical
    // config would otherwise reserve 201 vectors; a solve that converges in 25 iterations touches 51)
    for (auto &v : V_) { v.resize(0, A_->vec_prec); }
    for (auto &z : Z_) { z.resize(0, A_->vec_prec); }
    krylov_len_ = N;
    if (!hs_dev_) {
        AMGXB_CUDA_CHECK(cudaMalloc(&hs_dev_, (R_ + 8) * sizeof(double)));
        AMGXB_CUDA_CHECK(cudaHostAlloc(&hs_host_, (R_ + 8) * sizeof(double), cudaHostAllocMapped));
        AMGXB_CUDA_CHECK(cudaHostGetDevicePointer(&hs_host_dev_, hs_host_, 0));
    }
    update_x_every_iteration_ = (R_ == 1 || max_iters_ == 1) || trunc_;
    update_r_every_iteration_ = trunc_ && monitor_convergence_;
    resid_.resize(0, A_->vec_prec);
}

void FGMRESSolver::solve_init(DevVec &b, DevVec &x, bool xIsZero) {}

// GeneratePlaneRotation / PlaneRotation (fgmres_solver.cu:302-346)
static void generate_plane_rotation(double dx, double dy, double &cs, double &sn)
{
    if (dy < 0.0) { cs = 1.0; sn = 0.0; }
    else if (std::fabs(dy) > std::fabs(dx)) { double t = dx / dy; sn = 1.0 / std::sqrt(1.0 + t * t); cs = t * sn; }
    else { double t = dy / dx; cs = 1.0 / std::sqrt(1.0 + t * t); sn = t * cs; }
}

Status FGMRESSolver::solve_iteration(DevVec &b, DevVec &x, bool xIsZero)
{
    cudaStream_t s = stream();
    const size_t n = vec_len();
    const Prec vp = A_->vec_prec;
    Status conv_stat = ST_CONVERGED;
    const int m = curr_iter_ % R_;
    ReduceCtx red = red_ctx();
    red.scal = hs_dev_;            // Hessenberg column / norms live in our own scalar array
    red.host_mirror = hs_host_dev_;
    const int SLOT_BETA = R_ + 2;
    auto dist_fin = [&](int slot, bool is_norm) {   // distributed: all-reduce the partial, finish, mirror
        if (!A_->dist) return;
        dist_allreduce_norm(*A_, red, slot, is_norm ? 1 : 0, s);
    };
    auto need = [&](DevVec &v) { if (v.n != krylov_len_) { v.resize(krylov_len_, vp); v.zero(s); } };
    DevVec &Vm = Vr(m), &Vm1 = Vr(m + 1), &Zm = Zr(m);
    need(Vm);
    need(Vm1);
    need(Zm);
    const int sm = trunc_ ? std::max(m - krylov_dim_, 0) : 0;      // get_smallest_m(): the oldest vector still kept
    if (m == 0) {
        // r0 = b - A x ; beta = ||r0||
        dist_exchange_halo(*A_, x, s);
        CsrOpArgs g;
        g.x = x.ptr();
        g.b = b.ptr();
        g.y = Vm.ptr();
        matrix_apply(*A_, EPI_RESID, g, s);
        vec_dot(Vm.ptr(), Vm.ptr(), vp, n, red, A_->dist ? FIN_STORE : FIN_SQRT, SLOT_BETA, A_->dist ? 0 : 1, s);
        dist_fin(SLOT_BETA, true);
        AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
        beta_ = hs_host_[SLOT_BETA];
        if (curr_iter_ == 0 && monitor_convergence_) {
            nrm_.assign(1, beta_);
            conv_stat = converged();
            if (is_done(conv_stat)) return conv_stat;
        }
        vec_scal(Vm.ptr(), vp, n, 1.0 / beta_, s);
        std::fill(s_.begin(), s_.end(), 0.0);
        s_[0] = beta_;
    }
    // z_m = M^-1 v_m (zero initial guess)
    if (precond_) precond_->solve(Vm, Zm, true);
    else vec_copy(Zm.ptr(), Vm.ptr(), vp, n, s);
    // v_{m+1} = A z_m
    dist_exchange_halo(*A_, Zm, s);
    {
        CsrOpArgs g;
        g.x = Zm.ptr();
        g.y = Vm1.ptr();
        matrix_apply(*A_, EPI_SPMV, g, s);
    }
    // modified Gram-Schmidt over V(sm..m), coefficients stay on the device
    if (!A_->dist) {
        // single GPU: the update w -= h_i v_i and the next coefficient <v_{i+1}, w> (finally ||w||) share one pass over w
        vec_dot(Vr(sm).ptr(), Vm1.ptr(), vp, n, red, FIN_STORE, sm, 1, s);
        for (int i = sm; i < m; i++)
            vec_axpy_dot_dev(Vr(i).ptr(), Vm1.ptr(), Vr(i + 1).ptr(), vp, n, hs_dev_, i, -1.0, red, FIN_STORE, i + 1, 1, s);
        vec_axpy_dot_dev(Vm.ptr(), Vm1.ptr(), nullptr, vp, n, hs_dev_, m, -1.0, red, FIN_SQRT, m + 1, 1, s);
    } else {
        for (int i = sm; i <= m; i++) {
            vec_dot(Vr(i).ptr(), Vm1.ptr(), vp, n, red, FIN_STORE, i, 0, s);
            dist_fin(i, false);
            vec_axpy_dev(Vr(i).ptr(), Vm1.ptr(), vp, n, hs_dev_, i, -1.0, s);
        }
        vec_dot(Vm1.ptr(), Vm1.ptr(), vp, n, red, FIN_STORE, m + 1, 0, s);
        dist_fin(m + 1, true);
    }
    vec_scal_dev_inv(Vm1.ptr(), vp, n, hs_dev_, m + 1, s);
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));    // the one host sync of the iteration (standard path)
    for (int i = sm; i <= m + 1; i++) H(i, m) = hs_host_[i];      // rows above sm keep what is there (zero, or the previous cycle's: see the header)
    gamma_[m] = s_[m];
    // PlaneRotation(H, cs, sn, s, m)
    for (int k = 0; k < m; k++) {
        const double t = cs_[k] * H(k, m) + sn_[k] * H(k + 1, m);
        H(k + 1, m) = -sn_[k] * H(k, m) + cs_[k] * H(k + 1, m);
        H(k, m) = t;
    }
    generate_plane_rotation(H(m, m), H(m + 1, m), cs_[m], sn_[m]);
    H(m, m) = cs_[m] * H(m, m) + sn_[m] * H(m + 1, m);
    H(m + 1, m) = 0.0;
    {
        const double t = cs_[m] * s_[m];
        s_[m + 1] = -sn_[m] * s_[m];
        s_[m] = t;
    }
    if (update_x_every_iteration_) {
        // p_m = (z_m - sum_{i >= sm} h_im p_i) / h_mm ; x += s_m p_m   (fgmres_solver.cu:504-516)
        for (int i = sm; i < m; i++) vec_axpy(Zr(i).ptr(), Zm.ptr(), vp, n, -H(i, m), s);
        vec_scal(Zm.ptr(), vp, n, 1.0 / H(m, m), s);
        vec_axpy(Zm.ptr(), x.ptr(), vp, n, s_[m], s);
    }
    beta_ = std::fabs(s_[m + 1]);
    if (update_r_every_iteration_) {
        // r_m = (gamma_{m+1} c_m) v_{m+1} + (-gamma_{m+1} s_m / gamma_m) r_{m-1}   (fgmres_solver.cu:518-533); its norm is what is monitored
        need(resid_);
        if (m == 0) vec_axpby(Vr(1).ptr(), Vr(0).ptr(), resid_.ptr(), vp, n, s_[1] * cs_[0], -1.0 * s_[1] * sn_[0], s);
        else vec_axpby(Vm1.ptr(), resid_.ptr(), resid_.ptr(), vp, n, s_[m + 1] * cs_[m], -1.0 * s_[m + 1] * sn_[m] / gamma_[m], s);
        vec_dot(resid_.ptr(), resid_.ptr(), vp, n, red, A_->dist ? FIN_STORE : FIN_SQRT, SLOT_BETA, A_->dist ? 0 : 1, s);
        dist_fin(SLOT_BETA, true);
        AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
        nrm_.assign(1, hs_host_[SLOT_BETA]);
        conv_stat = converged();
    } else if (monitor_convergence_) {
        nrm_.assign(1, beta_);
        conv_stat = converged();
    } else conv_stat = ST_CONVERGED;
    if (!update_x_every_iteration_ && (m == R_ - 1 || is_last_iter() || is_done(conv_stat))) {
        for (int j = m; j >= 0; j--) {
            s_[j] /= H(j, j);
            for (int k = j - 1; k >= 0; k--) s_[k] -= H(k, j) * s_[j];
        }
        for (int j = 0; j <= m; j++) vec_axpy(Zr(j).ptr(), x.ptr(), vp, n, s_[j], s);
    }
    return monitor_convergence_ ? conv_stat : ST_CONVERGED;
}

}  // namespace amgxb
