// solvers.h -- host-side solver framework of the engine: the template-method Solver base
// (setup / solve loop / residual monitoring / convergence), Krylov solvers, smoothers and the AMG
// preconditioner.  Mirrors the reference's interface for this path:
//   Solver<TConfig>            include/solvers/solver.h:89-104, src/solvers/solver.cu:332-970
//   convergence criteria       src/convergence/*.cu
//   PCG_Solver                 src/solvers/pcg_solver.cu:39-190
//   FGMRES_Solver              src/solvers/fgmres_solver.cu:17-569
//   BlockJacobiSolver          src/solvers/block_jacobi_solver.cu:834-896, 1286-1347
//   JacobiL1Solver             src/solvers/jacobi_l1_solver.cu
//   AlgebraicMultigrid_Solver  src/solvers/algebraic_multigrid_solver.cu, src/amg.cu, src/cycles/fixed_cycle.cu
#pragma once
#include "base.h"
#include "config.h"
#include "matrix.h"
#include "kernels.h"
#include <functional>

namespace amgxb {

enum Status { ST_CONVERGED = 0, ST_NOT_CONVERGED = 1, ST_DIVERGED = 2, ST_FAILED = 3 };
enum NormType { NORM_L1 = 0, NORM_L2 = 1, NORM_LMAX = 2 };

inline bool is_done(Status s) { return s != ST_NOT_CONVERGED; }

// Convergence criteria on host norms; arithmetic copied literally in spirit from
// src/convergence/{absolute,relative_ini,relative_max,combined_rel_ini_abs}.cu since it decides
// iteration counts.
struct Convergence {
    enum Kind { ABSOLUTE, RELATIVE_INI, RELATIVE_MAX, COMBINED_REL_INI_ABS } kind = ABSOLUTE;
    double tolerance = 1e-12, alt_rel_tolerance = 1e-12;
    bool fp32 = false;
    std::vector<double> max_nrm;
    void init(const Config &cfg, const std::string &scope);
    Status update_and_check(const std::vector<double> &nrm, const std::vector<double> &nrm_ini);
};

// Per-solver device scalars + pinned host mirror; reduction scratch shared through Resources.
struct ScalarBlock {
    double *scal = nullptr;      // device, S_COUNT doubles
    double *host = nullptr;      // pinned + mapped host copy written by reduction epilogues
    double *host_dev = nullptr;  // device alias of `host`
    void create();
    void destroy();
};

// AMGXB_PHASE_TIMING=1: cudaEvent marks along the solve (graphs off), summed per label and printed to stderr when the outer solve ends.
// A diagnostic for where an iteration's time goes (per level, per phase); not used in timed runs.
bool phase_timing_on();
void phase_mark(const char *label, int level, cudaStream_t s);
void phase_report(cudaStream_t s, int iterations);

struct ReduceScratch {           // one per Resources (kernels on one stream run in order)
    DevBuf<double> partials;
    DevBuf<unsigned> counter;
    void ensure(cudaStream_t s);
};
ReduceScratch &reduce_scratch(Resources *rsc);

// Extra work a caller may ask a smoother to fuse into its sweeps.
struct SmoothFuse {
    const int *agg = nullptr;    // first sweep reads x := xc[agg[.]] (x itself holds no data yet)
    const void *xc = nullptr;
    bool dot_b_x = false;        // last sweep also reduces <b, x_new> ...
    int fin_op = FIN_STORE;      // ... finished with this scalar op into scal block `red`
    int fin_slot = S_TMP0;
    ReduceCtx red;
};

// A stretch of stream work between two host synchronisation points, replayed as a CUDA graph once its
// pointers have been seen twice (first use runs eagerly: lazy allocations, function attributes; second use is
// captured; later uses replay).  Kernel, memcpy/memset, NCCL and cross-stream event nodes are all captured.
struct GraphSegment {
    cudaGraphExec_t exec = nullptr;
    cudaGraph_t graph = nullptr;
    long long launches = 0;
    int uses = 0;
    const void *key0 = nullptr, *key1 = nullptr;
    bool failed = false;
    void reset();
    ~GraphSegment() { reset(); }
};
bool graphs_enabled();
// Dry run (AMGXB200_config_check): constructors parse and validate their configuration but create no device resources, so the
// whole solver tree of a configuration can be instantiated on a machine without a GPU.
extern thread_local bool g_dry_run;
// Solvers that synchronise with the host inside what an enclosing solver would capture (CG / CGF cycles) or whose launch
// sequence changes from one invocation to the next (reuse_scale) hold one of these for their lifetime.
struct GraphInhibit {
    bool on = false;
    void set();
    ~GraphInhibit();
};

class Solver {
public:
    Solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc);
    virtual ~Solver();

    void setup(Matrix &A, bool reuse_matrix_structure);
    Status solve(DevVec &b, DevVec &x, bool xIsZero);

    void set_max_iters(int m);
    void set_tolerance(double t) { conv_.tolerance = t; tol_override_ = true; tol_value_ = t; }
    int  get_num_iters() const { return num_iters_; }
    const std::vector<double> &get_residual(int idx) const;
    bool stores_history() const { return store_res_history_; }
    void norm_of(const DevVec &v, std::vector<double> &out) { compute_norm_of(v, out); }
    const std::string &name() const { return name_; }
    void set_name(const std::string &n) { name_ = n; }
    Matrix &get_A() { return *A_; }
    double setup_time() const { return setup_time_; }
    double solve_time() const { return solve_time_; }

    // smoother entry used by the cycles: `sweeps` iterations with no residual monitoring
    // (the reference: smoother->setTolerance(0); set_max_iters(n); solve(b, x, xIsZero),
    //  src/cycles/fixed_cycle.cu:97-102)
    // `input_in_alt`: the caller placed the initial x in smooth_input(x, sweeps) instead of x (see below).
    virtual void smooth(DevVec &b, DevVec &x, bool xIsZero, int sweeps, const SmoothFuse *fuse, bool input_in_alt = false);
    // Out-of-place smoothers ping-pong between x and a private buffer.  To end in x without a copy and without
    // swapping buffers (pointer stability is what makes the cycle capturable in a CUDA graph) the producer of
    // the initial x (prolongation) writes it where an odd/even number of sweeps needs it.
    virtual void *smooth_input(DevVec &x, int sweeps) { (void)sweeps; return x.ptr(); }
    virtual bool supports_fusion() const { return false; }
    virtual bool is_coloring_needed() const { return false; }
    virtual void print_grid_stats() {}
    // introspection for parity tests
    virtual const DevVec *smoother_data() const { return nullptr; }

    static std::unique_ptr<Solver> allocate(Config &cfg, const std::string &current_scope, const std::string &solver_type,
                                            std::shared_ptr<Resources> rsc);

protected:
    virtual void solver_setup(bool reuse_matrix_structure) = 0;
    virtual void solve_init(DevVec &b, DevVec &x, bool xIsZero) {}
    virtual Status solve_iteration(DevVec &b, DevVec &x, bool xIsZero) = 0;
    virtual void solve_finalize(DevVec &b, DevVec &x) {}
    virtual bool is_residual_needed() const { return false; }

    // helpers for derived classes
    void compute_residual(const DevVec &b, DevVec &x);            // r_ = b - A x
    void compute_norm();                                          // nrm_ from r_ (host sync)
    void compute_norm_of(const DevVec &v, std::vector<double> &out);
    void enqueue_norm(const DevVec &v);                           // scalar norm of v -> host mirror, no sync
    void read_norm(std::vector<double> &out);                     // sync + read what enqueue_norm produced
    template <class F> void run_segment(GraphSegment &g, const void *k0, const void *k1, F &&body);
    Status converged() { return conv_.update_and_check(nrm_, nrm_ini_); }
    Status converged(const DevVec &b, DevVec &x);                 // residual + norm + check when monitoring
    Status compute_norm_and_converged();
    bool is_last_iter() const { return curr_iter_ == max_iters_ - 1; }
    ReduceCtx red_ctx();
    cudaStream_t stream() const { return rsc_->stream; }
    size_t vec_len() const { return (size_t)A_->n * A_->by; }     // owned scalars

    Config *cfg_;
    std::string scope_, name_ = "SolverNameNotSet";
    std::shared_ptr<Resources> rsc_;
    Matrix *A_ = nullptr;
    DevVec r_;
    bool has_r_ = false;
    int max_iters_ = 100, num_iters_ = 0, curr_iter_ = 0;
    bool monitor_residual_ = false, monitor_convergence_ = false, store_res_history_ = false, obtain_timings_ = false;
    bool print_solve_stats_ = false, print_grid_stats_ = false, use_scalar_norm_ = false;
    int verbosity_ = 3;
    NormType norm_type_ = NORM_L2;
    Convergence conv_;
    bool tol_override_ = false;
    double tol_value_ = 0;
    std::vector<double> nrm_, nrm_ini_;
    std::vector<std::vector<double>> res_history_;
    ScalarBlock sb_;
    bool scaling_ = false;      // scaling = DIAGONAL_SYMMETRIC in this solver's scope
    DevVec scale_;              // s_i = 1 / sqrt(a_ii)
    bool is_setup_ = false;
    double setup_time_ = 0, solve_time_ = 0;
    cudaEvent_t ev_[4] = {nullptr, nullptr, nullptr, nullptr};
    friend class AMGSolver;
    friend class PCGSolver;
    friend class FGMRESSolver;
};

// ------------------------------------------------------------------------------------------
class NoSolver : public Solver {   // Dummy_Solver, src/solvers/dummy_solver.cu
public:
    using Solver::Solver;
protected:
    void solver_setup(bool) override {}
    Status solve_iteration(DevVec &b, DevVec &x, bool xIsZero) override;
};

class BlockJacobiSolver : public Solver {
public:
    BlockJacobiSolver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc);
    void smooth(DevVec &b, DevVec &x, bool xIsZero, int sweeps, const SmoothFuse *fuse, bool input_in_alt = false) override;
    void *smooth_input(DevVec &x, int sweeps) override { return (sweeps & 1) ? tmp_.ptr() : x.ptr(); }
    bool supports_fusion() const override { return A_ && A_->bs() == 1; }
    const DevVec *smoother_data() const override { return &dinv_; }
protected:
    void solver_setup(bool) override;
    Status solve_iteration(DevVec &b, DevVec &x, bool xIsZero) override;
    virtual void compute_d();
    double weight_ = 0.9;
    DevVec dinv_;      // 1x1: the diagonal itself (the reference stores d, not 1/d: block_jacobi_solver.cu:941-957)
    DevVec tmp_;       // ping-pong target of a sweep
};

class JacobiL1Solver : public BlockJacobiSolver {
public:
    JacobiL1Solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc);
protected:
    void compute_d() override;   // d_i = sum_j |a_ij|   (jacobi_l1_solver.cu:60-91)
};

class PCGSolver : public Solver {
public:
    PCGSolver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc);
    void print_grid_stats() override { if (precond_) precond_->print_grid_stats(); }
    Solver *preconditioner() { return precond_.get(); }
protected:
    void solver_setup(bool reuse) override;
    void solve_init(DevVec &b, DevVec &x, bool xIsZero) override;
    Status solve_iteration(DevVec &b, DevVec &x, bool xIsZero) override;
    bool is_residual_needed() const override { return true; }
    void apply_precond_and_rz(int fin_op);   // z = M^-1 r ; <r,z> -> scalars through fin_op
    std::unique_ptr<Solver> precond_;
    DevVec p_, z_, Ap_;
    GraphSegment segA_, segB_;
    void enqueue_A(DevVec &x);
    void enqueue_B();
};

class FGMRESSolver : public Solver {
public:
    FGMRESSolver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc);
    ~FGMRESSolver() override;
    void print_grid_stats() override { if (precond_) precond_->print_grid_stats(); }
    Solver *preconditioner() { return precond_.get(); }
protected:
    void solver_setup(bool reuse) override;
    void solve_init(DevVec &b, DevVec &x, bool xIsZero) override;
    Status solve_iteration(DevVec &b, DevVec &x, bool xIsZero) override;
    bool is_residual_needed() const override { return false; }
    std::unique_ptr<Solver> precond_;
    int R_ = 20, krylov_dim_ = 20;
    bool use_scalar_L2_ = true;
    std::vector<DevVec> V_, Z_;
    std::vector<double> H_, s_, cs_, sn_, gamma_;
    double beta_ = 0;
    size_t krylov_len_ = 0;
    bool update_x_every_iteration_ = false, update_r_every_iteration_ = false;
    bool trunc_ = false;                // gmres_krylov_dim < restart: the truncated variant (rings of krylov_dim + 2 / + 1 vectors)
    DevVec resid_;                      // truncated variant: the recursively updated residual vector whose norm is monitored
    DevVec &Vr(int i) { return V_[(size_t)i % V_.size()]; }
    DevVec &Zr(int i) { return Z_[(size_t)i % Z_.size()]; }
    double &H(int i, int j) { return H_[(size_t)i * (R_ + 1) + j]; }   // (R+2) x (R+1) storage
    double *hs_dev_ = nullptr, *hs_host_ = nullptr, *hs_host_dev_ = nullptr;   // Hessenberg column on device + pinned mirror
};

// ------------------------------------------------------------------------------------------
// AMG hierarchy
// ------------------------------------------------------------------------------------------
struct AMGLevel {
    Matrix *A = nullptr;                  // level 0: the caller's matrix; else owned below
    std::unique_ptr<Matrix> owned_A;
    int index = 0;
    bool coarsest = false;
    std::unique_ptr<Solver> smoother;
    // aggregation
    DevBuf<int> aggregates, R_row_offsets, R_column_indices;
    int n_coarse = 0;
    // classical
    std::unique_ptr<Matrix> P, R;
    DevBuf<int> cf_map;
    const Matrix *cla_reduce_over = nullptr;   // partitioned finest level: R holds this rank's columns only, the restricted residual is summed over the ranks of this matrix
    // cycle work vectors (sized for the NEXT level: bc, xc) and residual of this level
    DevVec bc, xc, r;
    bool init_cycle = false;
    // CG / CGF cycles: work vectors of the CG iterations run ON this level (src/cycles/cg_cycle.cu:25-40)
    DevVec cg_y, cg_z, cg_r, cg_p, cg_d;
    // error_scaling 2 / 3 (aggregation): prolongated correction, A * correction, the scale (device) and its reuse counter
    DevVec ef, Aef;
    DevBuf<double> scale;
    int scale_counter = 0;
    // distributed hierarchy: the NEXT level is replicated on every rank (dist.cu, "replicated coarse tail")
    bool tail_gather = false;
    int tail_off = 0;
    std::vector<int> tail_counts, tail_offs;
};

class AMGSolver : public Solver {
public:
    AMGSolver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc);
    void print_grid_stats() override;
    int num_levels() const { return (int)levels_.size(); }
    AMGLevel &level(int i) { return *levels_[i]; }
    // PCG asks for z = M^-1 r with <r,z> fused into the last finest-level sweep when possible
    bool solve_fused_dot(DevVec &b, DevVec &x, const ReduceCtx &red, int fin_op, int fin_slot);
protected:
    void solver_setup(bool reuse) override;
    void solve_init(DevVec &b, DevVec &x, bool xIsZero) override;
    Status solve_iteration(DevVec &b, DevVec &x, bool xIsZero) override;
    enum CycleType { CYC_V = 0, CYC_W = 1, CYC_F = 2, CYC_CG = 3, CYC_CGF = 4 };
    void cg_cycle_dispatch(int lvl, DevVec &b, DevVec &x, bool flex);               // CG(F)_CycleDispatcher::dispatch
    void scaled_correction(AMGLevel &L, const DevVec &rf, DevVec &x);              // x += lambda * smoothed(P xc)
    double host_dot(const DevVec &x, const DevVec &y, size_t n, const Matrix *over = nullptr);   // over: all-reduce across the ranks of this matrix
    int cycle_iters_ = 2, scaling_smoother_steps_ = 2, reuse_scale_ = 0;
    GraphInhibit inhibit_;
    std::vector<DevBuf<int>> reuse_aggregates_;     // resetup with structure_reuse_levels: aggregates carried over, per level
    std::vector<int> reuse_n_coarse_;
    std::vector<std::unique_ptr<Matrix>> reuse_P_, reuse_R_;   // classical: P and R carried over whole
    std::vector<DevBuf<int>> reuse_cf_;
    void cycle(int lvl, DevVec &b, DevVec &x, const SmoothFuse *top_fuse, int type = -1);   // type -1: the configured cycle
    int cycle_type_ = CYC_V;
    void setup_aggregation();
    void setup_classical();
    void distribute_finest(const std::vector<int> &g_offs);   // classical.cu: classical AMG on a row-partitioned matrix
    void validate_config();       // everything the setup would reject for configuration reasons alone (called by the constructor)
    void replicate_tail(long long tail_rows);   // distributed hierarchy: assemble the small levels on every rank (amg.cu)
    std::unique_ptr<Solver> make_smoother();
    std::vector<std::unique_ptr<AMGLevel>> levels_;
    std::string algorithm_, cycle_name_, selector_, coarse_solver_name_;
    int max_levels_ = 100, min_coarse_rows_ = 2, presweeps_ = 1, postsweeps_ = 1, finest_sweeps_ = -1, coarsest_sweeps_ = 2;
    int intensive_smoothing_ = 0, error_scaling_ = 0, dense_lu_num_rows_ = 0, dense_lu_max_rows_ = 0;
    double coarsen_threshold_ = 1.0;
    std::unique_ptr<Solver> coarse_solver_;
    GraphSegment seg_cycle_, seg_cycle_zero_;   // stand-alone AMG solver: the V-cycle between two convergence checks
    GraphSegment seg_coarse_;                   // AMGXB_GRAPH_COARSE=1: levels >= 1 of a V-cycle used as a preconditioner (fixed bc / xc buffers)
};

template <class F> void Solver::run_segment(GraphSegment &g, const void *k0, const void *k1, F &&body)
{
    if (!graphs_enabled() || g.failed) { body(); return; }
    if (g.exec && (g.key0 != k0 || g.key1 != k1)) g.reset();
    if (g.uses == 0 || g.key0 != k0 || g.key1 != k1) {   // first sight of these pointers: eager
        g.key0 = k0;
        g.key1 = k1;
        g.uses = 1;
        body();
        return;
    }
    cudaStream_t s = stream();
    if (!g.exec) {
        const long long l0 = g_kernel_launches;
        if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); g.failed = true; body(); return; }
        bool ok = true;
        try { body(); } catch (...) { ok = false; }
        cudaGraph_t gr = nullptr;
        cudaError_t e = cudaStreamEndCapture(s, &gr);
        if (!ok || e != cudaSuccess || !gr) {
            cudaGetLastError();
            if (gr) cudaGraphDestroy(gr);
            g.failed = true;
            g_kernel_launches = l0;
            body();   // the captured work never ran: run it eagerly
            return;
        }
        g.launches = g_kernel_launches - l0;
        g_kernel_launches = l0;
        g.graph = gr;
        if (cudaGraphInstantiate(&g.exec, gr, 0) != cudaSuccess) {
            cudaGetLastError();
            g.exec = nullptr;
            g.failed = true;
            body();
            return;
        }
    }
    AMGXB_CUDA_CHECK(cudaGraphLaunch(g.exec, s));
    g_kernel_launches += g.launches;
    g.uses++;
}

// pieces implemented in other translation units
std::unique_ptr<Solver> make_cg_solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc);         // krylov_extra.cu
std::unique_ptr<Solver> make_pcgf_solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc);
std::unique_ptr<Solver> make_pbicgstab_solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc);
std::unique_ptr<Solver> make_chebyshev_solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc);     // cheby.cu
std::unique_ptr<Solver> make_chebyshev_poly_solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc);
std::unique_ptr<Solver> make_gs_solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc);            // gs.cu
std::unique_ptr<Solver> make_gmres_solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc);
std::unique_ptr<Solver> make_dense_lu_solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc);   // dense_lu.cu
void classical_restrict(AMGLevel &L, const DevVec &r, cudaStream_t s);        // classical.cu: bc = R r
void classical_prolong_add(AMGLevel &L, const void *x, void *xout, cudaStream_t s);   // classical.cu: xout = x + P xc (x == nullptr: xout = P xc)
inline void solver_norm_of(Solver &sv, const DevVec &v, std::vector<double> &out) { sv.norm_of(v, out); }

}  // namespace amgxb
