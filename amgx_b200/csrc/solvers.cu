// solvers.cu -- Solver base class, convergence criteria, Jacobi smoothers, PCG.  See solvers.h.
#include "solvers.h"
#include <atomic>
#include "dist.h"
#include <cmath>
#include <algorithm>
#include <sstream>
#include <iomanip>
#include <map>
#include <mutex>

namespace amgxb {

// ---------------------------------------------------------------------------------------------
// Convergence  (src/convergence/*.cu)
// ---------------------------------------------------------------------------------------------
void Convergence::init(const Config &cfg, const std::string &scope)
{
    tolerance = cfg.get_double("tolerance", scope);
    alt_rel_tolerance = cfg.get_double("alt_rel_tolerance", scope);
    max_nrm.clear();
}

Status Convergence::update_and_check(const std::vector<double> &nrm, const std::vector<double> &nrm_ini)
{
    const double eps_conv = fp32 ? (double)1.0e-6f : 1.0e-12;   // Epsilon_conv<T>, include/convergence/convergence.h:21-46
    switch (kind) {
    case ABSOLUTE: {
        bool ok = true;
        for (size_t i = 0; i < nrm.size(); i++) ok = ok && (nrm[i] < tolerance);
        return ok ? ST_CONVERGED : ST_NOT_CONVERGED;
    }
    case RELATIVE_INI: {
        bool res = true, res_abs = true;
        const double eps = 1e-20;
        for (size_t i = 0; i < nrm.size(); i++) {
            bool conv = (nrm_ini[i] <= eps) ? true : (nrm[i] / nrm_ini[i] <= tolerance);
            res = res && conv;
            bool conv_abs = nrm[i] <= std::max(nrm_ini[i] * eps_conv, eps);
            res_abs = res_abs && conv_abs;
        }
        if (res_abs) return ST_CONVERGED;
        return res ? ST_CONVERGED : ST_NOT_CONVERGED;
    }
    case RELATIVE_MAX: {
        const double eps = fp32 ? 1e-10 : 1e-20;   // AMGX_NUMERICAL_SZERO / DZERO
        if (max_nrm.empty()) max_nrm = nrm;
        else for (size_t i = 0; i < nrm.size(); i++) max_nrm[i] = nrm[i] > max_nrm[i] ? nrm[i] : max_nrm[i];
        bool res = true, res_abs = true;
        for (size_t i = 0; i < nrm.size(); i++) {
            bool conv = (max_nrm[i] <= eps) ? true : ((nrm[i] / max_nrm[i]) <= tolerance);
            res = res && conv;
            bool conv_abs = nrm[i] <= std::max(max_nrm[i] * eps_conv, 1e-20);
            res_abs = res_abs && conv_abs;
        }
        if (res_abs) return ST_CONVERGED;
        return res ? ST_CONVERGED : ST_NOT_CONVERGED;
    }
    case COMBINED_REL_INI_ABS: {
        bool res = true, res_abs = true, res_prec = true;
        const double eps = 1e-20;
        for (size_t i = 0; i < nrm.size(); i++) {
            res_abs = res_abs && (nrm[i] < tolerance);
            res = res && ((nrm_ini[i] <= eps) ? true : (nrm[i] / nrm_ini[i] <= alt_rel_tolerance));
            res_prec = res_prec && (nrm[i] <= std::max(nrm_ini[i] * eps_conv, eps));
        }
        if (res_prec) return ST_CONVERGED;
        return (res || res_abs) ? ST_CONVERGED : ST_NOT_CONVERGED;
    }
    }
    return ST_NOT_CONVERGED;
}

// ---------------------------------------------------------------------------------------------
// scalar block / reduction scratch
// ---------------------------------------------------------------------------------------------
void ScalarBlock::create()
{
    if (scal) return;
    AMGXB_CUDA_CHECK(cudaMalloc(&scal, S_COUNT * sizeof(double)));
    AMGXB_CUDA_CHECK(cudaMemset(scal, 0, S_COUNT * sizeof(double)));
    AMGXB_CUDA_CHECK(cudaHostAlloc(&host, S_COUNT * sizeof(double), cudaHostAllocMapped));
    memset(host, 0, S_COUNT * sizeof(double));
    AMGXB_CUDA_CHECK(cudaHostGetDevicePointer(&host_dev, host, 0));
}
void ScalarBlock::destroy()
{
    if (scal) cudaFree(scal);
    if (host) cudaFreeHost(host);
    scal = host = host_dev = nullptr;
}

void ReduceScratch::ensure(cudaStream_t s)
{
    const size_t need = B200_SMS * 16 + 64;
    if (partials.size() < need) {
        partials.resize(need);
        counter.resize(1);
        counter.zero(s);
    }
}

static std::map<Resources *, std::unique_ptr<ReduceScratch>> g_scratch;
static std::mutex g_scratch_mu;
ReduceScratch &reduce_scratch(Resources *rsc)
{
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    auto &p = g_scratch[rsc];
    if (!p) p.reset(new ReduceScratch);
    p->ensure(rsc->stream);
    return *p;
}
void release_reduce_scratch(Resources *rsc)
{
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    g_scratch.erase(rsc);
}

void GraphSegment::reset()
{
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    exec = nullptr;
    graph = nullptr;
    uses = 0;
    launches = 0;
    key0 = key1 = nullptr;
    failed = false;
}

// distributed PCG: fuse <Ap,p>, <r,z> and ||r|| into the producing kernels (AMGXB_DIST_FUSE=0 falls back to separate dot kernels)
static bool dist_fuse()
{
    static const bool on = getenv("AMGXB_DIST_FUSE") ? atoi(getenv("AMGXB_DIST_FUSE")) != 0 : true;
    return on;
}

thread_local bool g_dry_run = false;
static std::atomic<int> g_graph_inhibit{0};
void GraphInhibit::set() { if (!on) { on = true; g_graph_inhibit++; } }
GraphInhibit::~GraphInhibit() { if (on) g_graph_inhibit--; }

bool phase_timing_on()
{
    static const bool on = getenv("AMGXB_PHASE_TIMING") ? atoi(getenv("AMGXB_PHASE_TIMING")) != 0 : false;
    return on;
}
namespace {
struct PhaseMark { cudaEvent_t ev; std::string label; };
std::vector<PhaseMark> g_marks;
std::vector<cudaEvent_t> g_mark_pool;
}
void phase_mark(const char *label, int level, cudaStream_t s)
{
    if (!phase_timing_on() || g_marks.size() > 200000) return;
    cudaEvent_t e;
    if (!g_mark_pool.empty()) { e = g_mark_pool.back(); g_mark_pool.pop_back(); }
    else if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, s);
    char buf[96];
    if (level >= 0) snprintf(buf, sizeof(buf), "L%02d %s", level, label);
    else snprintf(buf, sizeof(buf), "%s", label);
    g_marks.push_back(PhaseMark{e, buf});
}
void phase_report(cudaStream_t s, int iterations)
{
    if (!phase_timing_on() || g_marks.size() < 2) return;
    cudaStreamSynchronize(s);
    std::map<std::string, std::pair<double, int>> acc;
    double total = 0;
    for (size_t i = 1; i < g_marks.size(); i++) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, g_marks[i - 1].ev, g_marks[i].ev) != cudaSuccess) { cudaGetLastError(); continue; }
        acc[g_marks[i].label].first += ms;
        acc[g_marks[i].label].second++;
        total += ms;
    }
    const int it = std::max(iterations, 1);
    fprintf(stderr, "[amgx_b200 phase timing] %d iterations, %.3f ms marked, %.3f ms / iteration (time since the previous mark, per label)\n", iterations, total, total / it);
    for (auto &kv : acc)
        fprintf(stderr, "  %-28s %9.3f ms total  %8.1f us / iteration  (%d marks)\n", kv.first.c_str(), kv.second.first, kv.second.first / it * 1e3, kv.second.second);
    for (auto &m : g_marks) g_mark_pool.push_back(m.ev);
    g_marks.clear();
}

bool graphs_enabled()
{
    static const bool on = getenv("AMGXB_GRAPHS") ? atoi(getenv("AMGXB_GRAPHS")) != 0 : true;
    return on && g_graph_inhibit.load() == 0 && !phase_timing_on();
}

// ---------------------------------------------------------------------------------------------
// Solver base
// ---------------------------------------------------------------------------------------------
static NormType parse_norm(const std::string &s)
{
    if (s == "L1") return NORM_L1;
    if (s == "L2") return NORM_L2;
    if (s == "LMAX") return NORM_LMAX;
    fatal(AMGX_RC_BAD_CONFIGURATION, "norm '" + s + "' is not supported by this engine (L1, L2, LMAX)");
}

Solver::Solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc) : cfg_(&cfg), scope_(scope), rsc_(std::move(rsc))
{
    verbosity_ = cfg.get_int("verbosity_level", scope);
    monitor_residual_ = cfg.get_int("monitor_residual", scope) != 0;
    store_res_history_ = cfg.get_int("store_res_history", scope) != 0;
    obtain_timings_ = cfg.get_int("obtain_timings", scope) != 0;
    print_solve_stats_ = cfg.get_int("print_solve_stats", scope) != 0;
    print_grid_stats_ = cfg.get_int("print_grid_stats", scope) != 0;
    {
        const std::string sc = cfg.get_string("scaling", scope);
        if (sc == "DIAGONAL_SYMMETRIC") scaling_ = true;
        else if (sc != "NONE") fatal(AMGX_RC_BAD_CONFIGURATION, "scaling '" + sc + "' is not supported by this engine (NONE, DIAGONAL_SYMMETRIC)");
    }
    monitor_convergence_ = monitor_residual_;
    if (scope == "default") {   // backward compatibility rule of the reference (solver.cu:52-60)
        // the reference zeroes the print parameters of the default scope after reading the monitoring flags
        print_solve_stats_ = print_grid_stats_ = false;
    }
    if (print_solve_stats_ && !monitor_residual_)
        fatal(AMGX_RC_BAD_PARAMETERS, "Cannot print solver information if residual is not monitored (i.e. print_solve_stats=1 and monitor_residual=0) ");
    if (store_res_history_ && !monitor_residual_)
        fatal(AMGX_RC_BAD_PARAMETERS, "Cannot store residual information if residual is not monitored (i.e. store_res_history=1 and monitor_residual=0) ");
    max_iters_ = cfg.get_int("max_iters", scope);
    norm_type_ = parse_norm(cfg.get_string("norm", scope));
    use_scalar_norm_ = cfg.get_int("use_scalar_norm", scope) != 0;
    std::string cv = cfg.get_string("convergence", scope);
    if (cv == "ABSOLUTE") conv_.kind = Convergence::ABSOLUTE;
    else if (cv == "RELATIVE_INI" || cv == "RELATIVE_INI_CORE") conv_.kind = Convergence::RELATIVE_INI;
    else if (cv == "RELATIVE_MAX" || cv == "RELATIVE_MAX_CORE") conv_.kind = Convergence::RELATIVE_MAX;
    else if (cv == "COMBINED_REL_INI_ABS") conv_.kind = Convergence::COMBINED_REL_INI_ABS;
    else fatal(AMGX_RC_BAD_CONFIGURATION, "ConvergenceFactory '" + cv + "' has not been registered");
    if (store_res_history_) res_history_.resize(max_iters_ + 1);
    if (g_dry_run) return;
    sb_.create();
    for (auto &e : ev_) AMGXB_CUDA_CHECK(cudaEventCreate(&e));
}

Solver::~Solver()
{
    sb_.destroy();
    for (auto &e : ev_) if (e) cudaEventDestroy(e);
}

void Solver::set_max_iters(int m)
{
    if (store_res_history_ && (size_t)(m + 1) > res_history_.size()) res_history_.resize(m + 1);
    max_iters_ = m;
}

const std::vector<double> &Solver::get_residual(int idx) const
{
    if (!store_res_history_) fatal(AMGX_RC_BAD_PARAMETERS, "Residual history was not recorded");
    if (idx < 0 || (size_t)idx >= res_history_.size()) fatal(AMGX_RC_BAD_PARAMETERS, "Invalid iteration index while retrieving residual");
    return res_history_[idx];
}

ReduceCtx Solver::red_ctx()
{
    ReduceScratch &rs = reduce_scratch(rsc_.get());
    ReduceCtx c;
    c.partials = rs.partials.ptr();
    c.counter = rs.counter.ptr();
    c.scal = sb_.scal;
    c.host_mirror = sb_.host_dev;
    return c;
}

void Solver::setup(Matrix &A, bool reuse)
{
    if (obtain_timings_) AMGXB_CUDA_CHECK(cudaEventRecord(ev_[0], stream()));
    if (!A.initialized) fatal(AMGX_RC_BAD_PARAMETERS, "Trying to setup from the uninitialized matrix");
    if (reuse && A_ != &A) fatal(AMGX_RC_UNKNOWN, "Cannot call resetup with a different matrix");
    A_ = &A;
    conv_.fp32 = (A.vec_prec == Prec::F32);
    if (scaling_) {
        // src/solvers/solver.cu:440-477: scale the matrix in place, set the solver up on the SCALED matrix, undo the scaling.  (The
        // reference calls this "very slow" and unfinished; it is reproduced as it is, including the a*s/s round trip of the values.)
        if (A.dist) fatal(AMGX_RC_NOT_IMPLEMENTED, "Diagonal Symmetric scaling not supported for distributed matrices");
        if (A.bs() != 1 || A.has_ext_diag) fatal(AMGX_RC_NOT_IMPLEMENTED, "DIAGONAL_SYMMETRIC scaling: scalar matrices with the diagonal inside the CSR structure");
        if (diag_sym_scale_setup(A, scale_, stream()))
            fatal(AMGX_RC_NOT_IMPLEMENTED, "Diagonal symmetric scaling only applies to SPD systems with positive diagonal entries");
        diag_sym_scale_matrix(A, scale_, false, stream());
    }
    solver_setup(reuse);
    if (scaling_) diag_sym_scale_matrix(A, scale_, true, stream());
    if (monitor_residual_ || is_residual_needed()) {
        r_.resize((size_t)A.n_cols * A.by, A.vec_prec);
        r_.zero(stream());
        has_r_ = true;
    }
    if (monitor_convergence_) {
        const int bsize = use_scalar_norm_ ? 1 : A.by;
        nrm_.assign(bsize, 0.0);
        nrm_ini_.assign(bsize, 0.0);
    }
    if (obtain_timings_) {
        AMGXB_CUDA_CHECK(cudaEventRecord(ev_[1], stream()));
        AMGXB_CUDA_CHECK(cudaEventSynchronize(ev_[1]));
        float ms = 0;
        cudaEventElapsedTime(&ms, ev_[0], ev_[1]);
        setup_time_ = ms * 1e-3;
    }
    if (verbosity_ > 2 && print_grid_stats_) print_grid_stats();
    is_setup_ = true;
}

void Solver::compute_residual(const DevVec &b, DevVec &x)
{
    dist_exchange_halo(*A_, x, stream());
    CsrOpArgs g;
    g.x = x.ptr();
    g.b = b.ptr();
    g.y = r_.ptr();
    matrix_apply(*A_, EPI_RESID, g, stream());
}

void Solver::enqueue_norm(const DevVec &v)
{
    ReduceCtx red = red_ctx();
    const size_t n = vec_len();
    const bool d = (bool)A_->dist;
    if (norm_type_ == NORM_L2) vec_dot(v.ptr(), v.ptr(), v.prec, n, red, d ? FIN_STORE : FIN_SQRT, S_NRM, d ? 0 : 1, stream());
    else if (norm_type_ == NORM_L1) vec_nrm1(v.ptr(), v.prec, n, red, S_NRM, d ? 0 : 1, stream());
    else vec_nrmmax(v.ptr(), v.prec, n, red, S_NRM, d ? 0 : 1, stream());
    if (d) dist_allreduce_norm(*A_, red, S_NRM, (int)norm_type_, stream());
}

void Solver::read_norm(std::vector<double> &out)
{
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(stream()));
    out.assign(1, sb_.host[S_NRM]);
}

void Solver::compute_norm_of(const DevVec &v, std::vector<double> &out)
{
    const int bsize = use_scalar_norm_ ? 1 : A_->by;
    if (bsize == 1) {
        enqueue_norm(v);
        read_norm(out);
    } else {
        out.resize(bsize);
        ReduceCtx red = red_ctx();
        block_norms(v, A_->n, bsize, (int)norm_type_, red, sb_, out, stream(), A_);
    }
}

void Solver::compute_norm() { compute_norm_of(r_, nrm_); }

Status Solver::converged(const DevVec &b, DevVec &x)
{
    if (monitor_residual_) compute_residual(b, x);
    Status st = ST_NOT_CONVERGED;
    if (monitor_convergence_) {
        compute_norm();
        st = converged();
    }
    return st;
}

Status Solver::compute_norm_and_converged()
{
    compute_norm();
    return converged();
}

static void print_norm(std::stringstream &ss, const std::vector<double> &nrm)
{
    for (double v : nrm) ss << std::scientific << std::setprecision(6) << std::setw(15) << v;
}

// The iteration loop of the reference's Solver::solve (src/solvers/solver.cu:585-970), same order
// of residual / norm / convergence operations so that iteration counts agree.
Status Solver::solve(DevVec &b, DevVec &x, bool xIsZero)
{
    if (!is_setup_) fatal(AMGX_RC_BAD_CONFIGURATION, "Error, setup must be called before calling solve");
    if (obtain_timings_) AMGXB_CUDA_CHECK(cudaEventRecord(ev_[2], stream()));
    if (scaling_) {   // solver.cu:667-675: A <- S A S, b <- S b, x <- S^-1 x, all in place; undone before returning
        diag_sym_scale_matrix(*A_, scale_, false, stream());
        vec_scale_entrywise(b.ptr(), scale_.ptr(), b.prec, vec_len(), false, stream());
        vec_scale_entrywise(x.ptr(), scale_.ptr(), x.prec, vec_len(), true, stream());
    }
    struct Unscale {   // runs on every way out of solve(), exceptions included
        Solver *sv; DevVec *b, *x;
        ~Unscale()
        {
            if (!sv->scaling_) return;
            try {
                vec_scale_entrywise(x->ptr(), sv->scale_.ptr(), x->prec, sv->vec_len(), false, sv->stream());
                vec_scale_entrywise(b->ptr(), sv->scale_.ptr(), b->prec, sv->vec_len(), true, sv->stream());
                diag_sym_scale_matrix(*sv->A_, sv->scale_, true, sv->stream());
            } catch (...) {   // already unwinding from an error: nothing more to report
            }
        }
    } unscale_guard{this, &b, &x};
    if (monitor_residual_ || is_residual_needed()) {
        if (xIsZero) vec_copy(r_.ptr(), b.ptr(), b.prec, vec_len(), stream());
        else compute_residual(b, x);
    }
    std::vector<double> last_nrm;
    if (monitor_convergence_) {
        compute_norm();
        last_nrm = nrm_ini_ = nrm_;
    }
    if (store_res_history_) res_history_[0] = nrm_;
    const bool pstats = verbosity_ > 2 && print_solve_stats_;
    if (pstats) {
        std::stringstream ss;
        ss << std::setw(15) << "iter" << std::setw(20) << " Mem Usage (GB)" << std::setw(15) << "residual";
        for (size_t i = 0; i + 1 < nrm_.size(); i++) ss << std::setw(15) << " ";
        ss << std::setw(15) << "rate";
        for (size_t i = 0; i + 1 < nrm_.size(); i++) ss << std::setw(15) << " ";
        ss << std::endl;
        ss << "         ----------------------------------------------------------------------";
        for (size_t i = 0; i + 1 < nrm_.size(); i++) ss << "-----------------------";
        ss << std::endl;
        ss << std::setw(15) << "Ini";
        ss << std::setw(20) << device_mem_used_gb();
        print_norm(ss, nrm_);
        ss << std::endl;
        amgx_output(ss.str().c_str(), (int)ss.str().length());
    }
    if (monitor_convergence_) {
        conv_.init(*cfg_, scope_);
        if (tol_override_) conv_.tolerance = tol_value_;
        conv_.update_and_check(nrm_, nrm_ini_);
    }
    bool done = monitor_convergence_ ? (converged() == ST_CONVERGED) : false;
    if (max_iters_ == 0) return monitor_convergence_ ? ST_NOT_CONVERGED : ST_CONVERGED;
    if (!done) solve_init(b, x, xIsZero);
    Status conv_stat = done ? ST_CONVERGED : ST_NOT_CONVERGED;
    std::stringstream ss;
    for (curr_iter_ = 0; curr_iter_ < max_iters_ && !done; ++curr_iter_) {
        conv_stat = solve_iteration(b, x, xIsZero);
        xIsZero = false;
        done = monitor_convergence_ && is_done(conv_stat);
        if (pstats) {
            ss.str(std::string());
            ss << std::setw(15) << curr_iter_;
            ss << std::setw(20) << device_mem_used_gb();
            print_norm(ss, nrm_);
            ss << std::setw(15);
            for (size_t i = 0; i < last_nrm.size(); i++) ss << std::fixed << std::setprecision(4) << nrm_[i] / last_nrm[i] << std::setw(8);
            ss << std::endl;
            amgx_output(ss.str().c_str(), (int)ss.str().length());
            last_nrm = nrm_;
        }
        if (store_res_history_) res_history_[curr_iter_ + 1] = nrm_;
    }
    num_iters_ = curr_iter_;
    if (num_iters_ > 0) solve_finalize(b, x);
    if (obtain_timings_) {
        AMGXB_CUDA_CHECK(cudaEventRecord(ev_[3], stream()));
        AMGXB_CUDA_CHECK(cudaEventSynchronize(ev_[3]));
        float ms = 0;
        cudaEventElapsedTime(&ms, ev_[2], ev_[3]);
        solve_time_ = ms * 1e-3;
    }
    if (pstats) {
        ss.str(std::string());
        ss << "         ----------------------------------------------------------------------";
        for (size_t i = 0; i + 1 < nrm_.size(); i++) ss << "-----------------------";
        ss << std::endl;
        ss << "         Total Iterations: " << num_iters_ << std::endl;
        ss << "         Avg Convergence Rate: \t\t";
        const double eps = conv_.fp32 ? 1e-10 : 1e-20;
        for (size_t i = 0; i < last_nrm.size(); i++)
            ss << std::fixed << std::setw(15) << ((nrm_ini_[i] > eps) ? pow(last_nrm[i] / nrm_ini_[i], 1.0 / num_iters_) : nrm_ini_[i]);
        ss << std::endl;
        ss << "         Final Residual: \t\t" << std::setprecision(6);
        for (size_t i = 0; i < last_nrm.size(); i++) ss << std::scientific << std::setw(15) << last_nrm[i] << std::fixed;
        ss << std::endl;
        ss << "         Total Reduction in Residual: \t" << std::setprecision(6);
        for (size_t i = 0; i < last_nrm.size(); i++)
            ss << std::scientific << std::setw(15) << ((nrm_ini_[i] > eps) ? last_nrm[i] / nrm_ini_[i] : nrm_ini_[i]) << std::fixed;
        ss << std::endl;
        ss << "         Maximum Memory Usage: \t\t" << std::setprecision(3) << std::setw(15) << device_mem_used_gb() << " GB" << std::endl;
        ss << "         ----------------------------------------------------------------------";
        for (size_t i = 0; i + 1 < nrm_.size(); i++) ss << "-----------------------";
        ss << std::endl;
        amgx_output(ss.str().c_str(), (int)ss.str().length());
    }
    if (verbosity_ > 2 && obtain_timings_) {
        std::stringstream ts;
        ts << "Total Time: " << setup_time_ + solve_time_ << std::endl;
        ts << "    setup: " << setup_time_ << " s\n";
        ts << "    solve: " << solve_time_ << " s\n";
        ts << "    solve(per iteration): " << ((num_iters_ == 0) ? num_iters_ : solve_time_ / num_iters_) << " s\n";
        amgx_output(ts.str().c_str(), (int)ts.str().length());
    }
    return conv_stat;
}

// generic smoother entry: `sweeps` iterations of solve_iteration without monitoring
void Solver::smooth(DevVec &b, DevVec &x, bool xIsZero, int sweeps, const SmoothFuse *fuse, bool input_in_alt)
{
    if (input_in_alt) fatal(AMGX_RC_INTERNAL, "smoother has no alternate input buffer");
    if (fuse && (fuse->agg || fuse->dot_b_x)) fatal(AMGX_RC_INTERNAL, "smoother does not support fused sweeps");
    // the smoother's own settings come back on every path, a throwing solve included (graceful-failure recovery: CAPIFailure)
    struct Restore {
        Solver &s;
        int max_iters;
        bool mr, mc, tol_override;
        double tol_value, conv_tol;
        ~Restore()
        {
            s.max_iters_ = max_iters; s.monitor_residual_ = mr; s.monitor_convergence_ = mc;
            s.tol_override_ = tol_override; s.tol_value_ = tol_value; s.conv_.tolerance = conv_tol;
        }
    } restore{*this, max_iters_, monitor_residual_, monitor_convergence_, tol_override_, tol_value_, conv_.tolerance};
    set_max_iters(sweeps);
    set_tolerance(0.0);
    solve(b, x, xIsZero);
}

// ---------------------------------------------------------------------------------------------
// NOSOLVER
// ---------------------------------------------------------------------------------------------
Status NoSolver::solve_iteration(DevVec &b, DevVec &x, bool xIsZero)
{
    if (xIsZero) x.zero(stream());
    return converged(b, x);
}

// ---------------------------------------------------------------------------------------------
// BLOCK_JACOBI / JACOBI_L1
// ---------------------------------------------------------------------------------------------
BlockJacobiSolver::BlockJacobiSolver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc) : Solver(cfg, scope, std::move(rsc))
{
    weight_ = cfg.get_double("relaxation_factor", scope);
    if (weight_ == 0) {
        weight_ = 1.;
        amgx_printf("Warning, setting weight to 1 instead of estimating largest_eigen_value in Block Jacobi smoother\n");
    }
}

void BlockJacobiSolver::compute_d() { extract_diagonal(*A_, dinv_, stream()); }

void BlockJacobiSolver::solver_setup(bool)
{
    if (A_->bx != A_->by) fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "Unsupported block size for BlockJacobi_Solver");
    compute_d();
    if (A_->bs() != 1) block_jacobi_setup(*A_, dinv_, stream());   // invert diagonal blocks (k_block.cu)
    tmp_.resize((size_t)A_->n_cols * A_->by, A_->vec_prec);
    tmp_.zero(stream());
}

void BlockJacobiSolver::smooth(DevVec &b, DevVec &x, bool xIsZero, int sweeps, const SmoothFuse *fuse, bool input_in_alt)
{
    // `sweeps` Jacobi iterations ending in x's own buffer, no buffer swaps: the iterates ping-pong between x and
    // tmp_, and the FIRST write goes wherever an even number of hops remains.
    cudaStream_t s = stream();
    const bool scalar = A_->bs() == 1;
    if (sweeps <= 0) return;
    if (!scalar && fuse && (fuse->agg || fuse->dot_b_x)) fatal(AMGX_RC_INTERNAL, "block Jacobi: fused sweeps need block size 1");
    const bool via_agg = fuse && fuse->agg;
    void *xp = x.ptr(), *tp = tmp_.ptr();
    void *cur;
    int m;   // number of SpMV sweeps still to run
    if (xIsZero && !via_agg) {
        m = sweeps - 1;
        cur = (m & 1) ? tp : xp;
        if (scalar) jacobi_zero_guess(b.ptr(), dinv_.ptr(), cur, A_->mat_prec, A_->vec_prec, vec_len(), weight_, s);   // x = w b / d, no SpMV
        else block_jacobi_zero(*A_, dinv_, b, cur, weight_, s);
        if (m == 0 && fuse && fuse->dot_b_x) vec_dot(b.ptr(), cur, x.prec, vec_len(), fuse->red, fuse->fin_op, fuse->fin_slot, 0, s);
    } else {
        m = sweeps;
        if (via_agg) cur = nullptr;   // first sweep reads P xc on the fly
        else if (m & 1) {
            if (!input_in_alt) vec_copy(tp, xp, x.prec, vec_len(), s);
            cur = tp;
        } else {
            if (input_in_alt) fatal(AMGX_RC_INTERNAL, "Jacobi: alternate input with an even sweep count");
            cur = xp;
        }
    }
    for (int it = 0; it < m; it++) {
        const bool last = (it == m - 1);
        const bool want_dot = fuse && fuse->dot_b_x && last;
        void *out;
        if (cur == nullptr) out = (m & 1) ? xp : tp;   // agg-fused first sweep: choose so that the last write is x
        else out = (cur == xp) ? tp : xp;
        if (scalar) {
            CsrOpArgs g;
            g.b = b.ptr();
            g.d = dinv_.ptr();
            g.omega = weight_;
            g.y = out;
            if (cur == nullptr) {
                g.x = fuse->xc;
                g.agg = fuse->agg;
            } else {
                dist_exchange_halo_ptr(*A_, cur, x.prec, s);
                g.x = cur;
            }
            if (want_dot) {
                g.red = fuse->red;
                g.fin_op = fuse->fin_op;
                g.fin_slot = fuse->fin_slot;
                matrix_apply(*A_, EPI_JACOBI_DOT, g, s);
            } else {
                matrix_apply(*A_, EPI_JACOBI, g, s);
            }
        } else {
            dist_exchange_halo_ptr(*A_, cur, x.prec, s);
            block_jacobi_sweep(*A_, dinv_, b, cur, out, weight_, s);
        }
        cur = out;
    }
    if (cur != xp) fatal(AMGX_RC_INTERNAL, "Jacobi ping-pong did not end in x");
}

Status BlockJacobiSolver::solve_iteration(DevVec &b, DevVec &x, bool xIsZero)
{
    smooth(b, x, xIsZero, 1, nullptr);
    return converged(b, x);
}

JacobiL1Solver::JacobiL1Solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc) : BlockJacobiSolver(cfg, scope, std::move(rsc)) {}

void JacobiL1Solver::compute_d()
{
    if (A_->bs() != 1) fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "Unsupported block size for JacobiL1_Solver");
    l1_row_norms(*A_, dinv_, stream());
}

// ---------------------------------------------------------------------------------------------
// PCG   (src/solvers/pcg_solver.cu)
// ---------------------------------------------------------------------------------------------
PCGSolver::PCGSolver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc) : Solver(cfg, scope, rsc)
{
    std::string name, ns;
    cfg.get_scoped("preconditioner", scope, name, ns);
    if (name != "NOSOLVER") precond_ = Solver::allocate(cfg, scope, "preconditioner", rsc);
}

void PCGSolver::solver_setup(bool reuse)
{
    segA_.reset();
    segB_.reset();
    if (precond_) precond_->setup(*A_, reuse);
    const size_t N = (size_t)A_->n_cols * A_->by;
    p_.resize(N, A_->vec_prec);
    z_.resize(N, A_->vec_prec);
    Ap_.resize(N, A_->vec_prec);
    p_.zero(stream());
    z_.zero(stream());
    Ap_.zero(stream());
}

// z = M^-1 r with zero initial guess, then <r,z> finished on the device by fin_op
void PCGSolver::apply_precond_and_rz(int fin_op)
{
    cudaStream_t s = stream();
    ReduceCtx red = red_ctx();
    if (A_->dist) red = dist_wrap_reduce(*A_, red);
    if (!precond_) {
        vec_copy(z_.ptr(), r_.ptr(), r_.prec, vec_len(), s);
        vec_dot(r_.ptr(), z_.ptr(), r_.prec, vec_len(), red, A_->dist ? FIN_STORE : fin_op, S_TMP0, 0, s);
    } else {
        AMGSolver *amg = dynamic_cast<AMGSolver *>(precond_.get());
        bool fused = false;
        if (amg && (!A_->dist || dist_fuse())) fused = amg->solve_fused_dot(r_, z_, red, A_->dist ? FIN_STORE : fin_op, S_TMP0);
        if (!fused) {
            precond_->solve(r_, z_, true);
            vec_dot(r_.ptr(), z_.ptr(), r_.prec, vec_len(), red, A_->dist ? FIN_STORE : fin_op, S_TMP0, 0, s);
        }
    }
    if (A_->dist) dist_allreduce_scalar_fin(*A_, red, S_TMP0, fin_op, s);
}

void PCGSolver::solve_init(DevVec &b, DevVec &x, bool xIsZero)
{
    cudaStream_t s = stream();
    // rz = <r, z>: FIN_PCG_BETA stores the sum in S_RZ (beta is garbage here and unused)
    AMGXB_CUDA_CHECK(cudaMemsetAsync(sb_.scal + S_RZ, 0, sizeof(double), s));
    apply_precond_and_rz(FIN_PCG_BETA);
    vec_copy(p_.ptr(), z_.ptr(), z_.prec, vec_len(), s);
}

// Segment A of an iteration: Ap = A p, alpha, x += alpha p, r -= alpha Ap, ||r|| (everything up to the host's
// convergence check).  Segment B: z = M^-1 r, rz, beta, p = z + beta p.  Each is one CUDA graph after warm-up.
void PCGSolver::enqueue_A(DevVec &x)
{
    cudaStream_t s = stream();
    ReduceCtx red = red_ctx();
    dist_exchange_halo(*A_, p_, s);
    if (A_->bs() == 1 && (!A_->dist || dist_fuse())) {
        CsrOpArgs g;
        g.x = p_.ptr();
        g.y = Ap_.ptr();
        g.red = red;
        g.fin_op = A_->dist ? FIN_STORE : FIN_PCG_ALPHA;
        g.fin_slot = A_->dist ? S_TMP0 : S_DOT;
        matrix_apply(*A_, EPI_SPMV_DOT, g, s);      // Ap = A p fused with <Ap,p>; alpha = rz / <Ap,p> on the device
        if (A_->dist) dist_allreduce_scalar_fin(*A_, red, S_TMP0, FIN_PCG_ALPHA, s);   // partial sums of the ranks, then alpha
    } else {
        CsrOpArgs g;
        g.x = p_.ptr();
        g.y = Ap_.ptr();
        matrix_apply(*A_, EPI_SPMV, g, s);
        vec_dot(Ap_.ptr(), p_.ptr(), p_.prec, vec_len(), red, A_->dist ? FIN_STORE : FIN_PCG_ALPHA, S_TMP0, 0, s);
        if (A_->dist) dist_allreduce_scalar_fin(*A_, red, S_TMP0, FIN_PCG_ALPHA, s);
    }
    const bool scalar_norm = use_scalar_norm_ || A_->by == 1;
    if (monitor_convergence_ && scalar_norm && (!A_->dist || dist_fuse())) {
        const bool d = (bool)A_->dist;
        pcg_update_xr(p_.ptr(), Ap_.ptr(), x.ptr(), r_.ptr(), x.prec, vec_len(), red, (int)norm_type_, S_NRM, d ? 0 : 1, s, d);   // one pass
        if (d) dist_allreduce_norm(*A_, red, S_NRM, (int)norm_type_, s);
    } else {
        vec_axpy_dev(p_.ptr(), x.ptr(), x.prec, vec_len(), sb_.scal, S_ALPHA, 1.0, s);
        vec_axpy_dev(Ap_.ptr(), r_.ptr(), x.prec, vec_len(), sb_.scal, S_NEG_ALPHA, 1.0, s);
        if (monitor_convergence_ && scalar_norm) enqueue_norm(r_);
    }
}

void PCGSolver::enqueue_B()
{
    apply_precond_and_rz(FIN_PCG_BETA);                                                         // z = M^-1 r ; rz ; beta
    vec_axpby_dev(z_.ptr(), p_.ptr(), p_.ptr(), p_.prec, vec_len(), 1.0, sb_.scal, S_BETA, stream());   // p = z*1 + p*beta
}

Status PCGSolver::solve_iteration(DevVec &b, DevVec &x, bool xIsZero)
{
    Status conv_stat = ST_NOT_CONVERGED;
    phase_mark("pcg: begin iteration", -1, stream());
    run_segment(segA_, x.ptr(), r_.ptr(), [&] { enqueue_A(x); });
    phase_mark("pcg: A p, alpha, x/r update, norm", -1, stream());
    if (monitor_convergence_) {
        const bool scalar_norm = use_scalar_norm_ || A_->by == 1;
        if (scalar_norm) read_norm(nrm_);
        else compute_norm();
        conv_stat = converged();
        if (is_done(conv_stat)) return conv_stat;
    }
    if (is_last_iter()) return monitor_convergence_ ? ST_NOT_CONVERGED : ST_CONVERGED;
    run_segment(segB_, z_.ptr(), r_.ptr(), [&] { enqueue_B(); });
    phase_mark("pcg: beta, p update (after V-cycle)", -1, stream());
    return monitor_convergence_ ? ST_NOT_CONVERGED : ST_CONVERGED;
}

// ---------------------------------------------------------------------------------------------
// factory  (SolverFactory::allocate, src/solvers/solver.cu:1098-1134)
// ---------------------------------------------------------------------------------------------
std::unique_ptr<Solver> Solver::allocate(Config &cfg, const std::string &current_scope, const std::string &solver_type,
                                         std::shared_ptr<Resources> rsc)
{
    std::string name, new_scope;
    cfg.get_scoped(solver_type, current_scope, name, new_scope);
    if ((solver_type == "coarse_solver" || solver_type == "smoother" || solver_type == "preconditioner") &&
        (name == "AMG" || name == "FGMRES" || name == "PCGF" || name == "PBICGSTAB" || name == "PCG") && new_scope == "default")
        fatal(AMGX_RC_BAD_PARAMETERS, "Solver " + name + " uses an inner solver (i.e. a preconditioner, smoother or coarse_solver) and therefore cannot be used as an inner solver with the default scope due to the possibility of an infinite number of nested solvers. Please use config_version=2, and specify a new scope name for the inner solver. For example: preconditioner(amg_solver) = AMG. \n");
    std::unique_ptr<Solver> sv;
    if (name == "PCG") sv.reset(new PCGSolver(cfg, new_scope, rsc));
    else if (name == "FGMRES") sv.reset(new FGMRESSolver(cfg, new_scope, rsc));
    else if (name == "AMG") sv.reset(new AMGSolver(cfg, new_scope, rsc));
    else if (name == "BLOCK_JACOBI") sv.reset(new BlockJacobiSolver(cfg, new_scope, rsc));
    else if (name == "JACOBI_L1") sv.reset(new JacobiL1Solver(cfg, new_scope, rsc));
    else if (name == "MULTICOLOR_DILU") sv = make_dilu_solver(cfg, new_scope, rsc);
    else if (name == "MULTICOLOR_GS") sv = make_gs_solver(cfg, new_scope, rsc);
    else if (name == "CHEBYSHEV") sv = make_chebyshev_solver(cfg, new_scope, rsc);
    else if (name == "CHEBYSHEV_POLY") sv = make_chebyshev_poly_solver(cfg, new_scope, rsc);
    else if (name == "CG") sv = make_cg_solver(cfg, new_scope, rsc);
    else if (name == "PCGF") sv = make_pcgf_solver(cfg, new_scope, rsc);
    else if (name == "PBICGSTAB") sv = make_pbicgstab_solver(cfg, new_scope, rsc);
    else if (name == "GMRES") sv = make_gmres_solver(cfg, new_scope, rsc);
    else if (name == "DENSE_LU_SOLVER") sv = make_dense_lu_solver(cfg, new_scope, rsc);
    else if (name == "NOSOLVER") sv.reset(new NoSolver(cfg, new_scope, rsc));
    else
        fatal(AMGX_RC_BAD_CONFIGURATION, "Solver '" + name + "' is outside the scope of the B200 solve-phase engine "
              "(supported: PCG, PCGF, CG, PBICGSTAB, GMRES, FGMRES, AMG, BLOCK_JACOBI, JACOBI_L1, MULTICOLOR_DILU, MULTICOLOR_GS, CHEBYSHEV, CHEBYSHEV_POLY, DENSE_LU_SOLVER, NOSOLVER)");
    sv->set_name(name);
    return sv;
}

}  // namespace amgxb
