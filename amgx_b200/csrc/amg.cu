// amg.cu -- the AMG preconditioner: hierarchy construction loop and the V-cycle.
//   level loop / stopping rules   src/amg.cu:152-421
//   V-cycle body                  src/cycles/fixed_cycle.cu:25-248
//   aggregation level             src/aggregation/aggregation_amg_level.cu:237-299, 842-907, 1945-2023
// The cycle is restructured around fused kernels (see DESIGN.md "V-cycle dataflow"):
//   * with zero pre-sweeps and a zero initial guess the residual IS b: no SpMV, no copy;
//   * the prolongation of the coarse correction is folded into the first post-smoothing sweep
//     (x is read through aggregates[] from xc and never written unsmoothed);
//   * the last finest-level sweep can carry PCG's <r,z> reduction.
// Each shortcut produces bit-identical vectors to the unfused sequence (0 + e == e, b - A*0 == b).
#include "solvers.h"
#include "dist.h"
#include "p2p.h"
#include <sstream>
#include <iomanip>

namespace amgxb {

AMGSolver::AMGSolver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc) : Solver(cfg, scope, rsc)
{
    algorithm_ = cfg.get_string("algorithm", scope);
    cycle_name_ = cfg.get_string("cycle", scope);
    selector_ = cfg.get_string("selector", scope);
    max_levels_ = cfg.get_int("max_levels", scope);
    min_coarse_rows_ = cfg.get_int("min_coarse_rows", scope);
    coarsen_threshold_ = cfg.get_double("coarsen_threshold", scope);
    presweeps_ = cfg.get_int("presweeps", scope);
    postsweeps_ = cfg.get_int("postsweeps", scope);
    finest_sweeps_ = cfg.get_int("finest_sweeps", scope);
    coarsest_sweeps_ = cfg.get_int("coarsest_sweeps", scope);
    intensive_smoothing_ = cfg.get_int("intensive_smoothing", scope);
    error_scaling_ = cfg.get_int("error_scaling", scope);
    if (cycle_name_ == "V") cycle_type_ = CYC_V;
    else if (cycle_name_ == "W") cycle_type_ = CYC_W;
    else if (cycle_name_ == "F") cycle_type_ = CYC_F;
    else if (cycle_name_ == "CG") cycle_type_ = CYC_CG;
    else if (cycle_name_ == "CGF") cycle_type_ = CYC_CGF;
    else fatal(AMGX_RC_BAD_CONFIGURATION, "CycleFactory '" + cycle_name_ + "' has not been registered");
    cycle_iters_ = cfg.get_int("cycle_iters", scope);
    scaling_smoother_steps_ = cfg.get_int("scaling_smoother_steps", scope);
    reuse_scale_ = cfg.get_int("reuse_scale", scope);
    // allowed values of the reference's parameter registry: 0, 2, 3 (src/core.cu:437-441)
    if (error_scaling_ != 0 && error_scaling_ != 2 && error_scaling_ != 3)
        fatal(AMGX_RC_BAD_CONFIGURATION, "error_scaling must be 0, 2 or 3");
    if ((cycle_type_ == CYC_CG || cycle_type_ == CYC_CGF) && cycle_iters_ < 1) fatal(AMGX_RC_BAD_CONFIGURATION, "cycle_iters must be >= 1");
    // host scalars inside the cycle (CG / CGF) or a launch sequence that differs between cycles (reuse_scale): no graph capture around us
    if (cycle_type_ == CYC_CG || cycle_type_ == CYC_CGF || (error_scaling_ >= 2 && reuse_scale_ > 0)) inhibit_.set();
    if (algorithm_ != "AGGREGATION" && algorithm_ != "CLASSICAL")
        fatal(AMGX_RC_BAD_CONFIGURATION, "algorithm '" + algorithm_ + "' is not supported (AGGREGATION, CLASSICAL)");
    std::string ns;
    cfg.get_scoped("coarse_solver", scope, coarse_solver_name_, ns);
    if (coarse_solver_name_ == "DENSE_LU_SOLVER") {
        dense_lu_num_rows_ = cfg.get_int("dense_lu_num_rows", scope);
        dense_lu_max_rows_ = cfg.get_int("dense_lu_max_rows", scope);
        coarse_solver_ = make_dense_lu_solver(cfg, ns, rsc);
        coarse_solver_->set_name("DENSE_LU_SOLVER");
    } else if (coarse_solver_name_ != "NOSOLVER") {
        coarse_solver_ = Solver::allocate(cfg, scope, "coarse_solver", rsc);
    }
    validate_config();
    if (g_dry_run) make_smoother();     // smoothers are otherwise created per level at setup: instantiate one so that its configuration is checked too
}

// Configuration-only checks of the two hierarchy builders, so that an unsupported option is reported at AMGX_solver_create (and by
// AMGXB200_config_check) rather than at the first setup.
void AMGSolver::validate_config()
{
    if (algorithm_ == "AGGREGATION") {
        if (selector_ != "SIZE_2" && selector_ != "SIZE_4")
            fatal(AMGX_RC_BAD_CONFIGURATION, "aggregation selector '" + selector_ + "' is not supported by this engine (SIZE_2, SIZE_4)");
        if (cfg_->get_int("handshaking_phases", scope_) == 2 && selector_ == "SIZE_2")
            fatal(AMGX_RC_NOT_IMPLEMENTED, "SIZE_2 selector: handshaking_phases=2 is not implemented");
        return;
    }
    const std::string strength = cfg_->get_string("strength", scope_), interp = cfg_->get_string("interpolator", scope_);
    const std::string agg_sel = cfg_->get_string("aggressive_selector", scope_), agg_int = cfg_->get_string("aggressive_interpolator", scope_);
    if (strength != "AHAT") fatal(AMGX_RC_BAD_CONFIGURATION, "strength '" + strength + "' is not supported by this engine (AHAT)");
    if (selector_ != "PMIS" && selector_ != "HMIS")
        fatal(AMGX_RC_BAD_CONFIGURATION, "classical selector '" + selector_ + "' is not supported by this engine (PMIS, HMIS)");
    if (interp != "D1" && interp != "D2" && interp != "MULTIPASS")
        fatal(AMGX_RC_BAD_CONFIGURATION, "interpolator '" + interp + "' is not supported by this engine (D1, D2, MULTIPASS)");
    if (cfg_->get_int("aggressive_levels", scope_) > 0) {
        if (agg_sel != "DEFAULT" && agg_sel != "PMIS" && agg_sel != "HMIS")
            fatal(AMGX_RC_BAD_CONFIGURATION, "aggressive_selector '" + agg_sel + "' is not supported (DEFAULT, PMIS, HMIS)");
        if (agg_int != "MULTIPASS") fatal(AMGX_RC_BAD_CONFIGURATION, "aggressive_interpolator '" + agg_int + "' is not supported (MULTIPASS)");
    }
}

std::unique_ptr<Solver> AMGSolver::make_smoother() { return Solver::allocate(*cfg_, scope_, "smoother", rsc_); }

void AMGSolver::solver_setup(bool reuse)
{
    // AMGX_solver_resetup with structure_reuse_levels = k: the aggregates (hence R and P) of the first k-1 coarsenings are
    // kept and only the Galerkin values, the smoothers and everything below are recomputed (src/amg.cu:229-272: a level is
    // rebuilt when structure_reuse_levels <= its 1-based index; -1 keeps the structure of every level).
    seg_cycle_.reset();          // captured cycles refer to the buffers of the previous hierarchy
    seg_cycle_zero_.reset();
    seg_coarse_.reset();
    reuse_aggregates_.clear();
    reuse_n_coarse_.clear();
    reuse_P_.clear();
    reuse_R_.clear();
    reuse_cf_.clear();
    const int reuse_levels = cfg_->get_int("structure_reuse_levels", scope_);
    if (reuse && reuse_levels != 0 && algorithm_ == "AGGREGATION" && !levels_.empty() && !A_->dist) {
        for (size_t l = 0; l + 1 < levels_.size(); l++) {
            if (reuse_levels != -1 && reuse_levels <= (int)l + 1) break;
            if (levels_[l]->n_coarse <= 0 || (l == 0 && levels_[0]->A->n != A_->n)) break;
            reuse_aggregates_.emplace_back();
            reuse_aggregates_.back().swap(levels_[l]->aggregates);
            reuse_n_coarse_.push_back(levels_[l]->n_coarse);
        }
    } else if (reuse && reuse_levels != 0 && algorithm_ == "CLASSICAL" && !levels_.empty() && !A_->dist) {
        // classical levels keep P and R whole -- pattern AND values -- and only A_c = R A P is recomputed (classical_amg_level.cu:274-291)
        for (size_t l = 0; l + 1 < levels_.size(); l++) {
            if (reuse_levels != -1 && reuse_levels <= (int)l + 1) break;
            if (!levels_[l]->P || !levels_[l]->R || (l == 0 && levels_[0]->A->n != A_->n)) break;
            reuse_P_.push_back(std::move(levels_[l]->P));
            reuse_R_.push_back(std::move(levels_[l]->R));
            reuse_cf_.emplace_back();
            reuse_cf_.back().swap(levels_[l]->cf_map);
            reuse_n_coarse_.push_back(levels_[l]->n_coarse);
        }
    }
    levels_.clear();
    if (dense_lu_num_rows_ > 0) min_coarse_rows_ = dense_lu_num_rows_ / A_->by;   // src/amg.cu:1154-1157
    if (algorithm_ == "AGGREGATION") setup_aggregation();
    else setup_classical();
}

// The level-building loop of AMG_Setup::setup (src/amg.cu:201-418), single-partition form.
void AMGSolver::setup_aggregation()
{
    if (selector_ != "SIZE_2" && selector_ != "SIZE_4")
        fatal(AMGX_RC_BAD_CONFIGURATION, "aggregation selector '" + selector_ + "' is not supported by this engine (SIZE_2, SIZE_4)");
    if (selector_ == "SIZE_4" && A_->dist) fatal(AMGX_RC_NOT_IMPLEMENTED, "SIZE_4 selector on a distributed matrix (use SIZE_2)");
    cudaStream_t s = stream();
    AggSetupParams prm;
    prm.deterministic = cfg_->get_int("determinism_flag", "default");
    prm.max_iterations = cfg_->get_int("max_matching_iterations", scope_);
    prm.max_unassigned = cfg_->get_double("max_unassigned_percentage", scope_);
    prm.two_phase = cfg_->get_int("handshaking_phases", scope_) == 2;
    prm.edge_weight_component = cfg_->get_int("aggregation_edge_weight_component", scope_);
    prm.merge_singletons = cfg_->get_int("merge_singletons", scope_) == 1;
    prm.weight_formula = cfg_->get_int("weight_formula", scope_);

    levels_.emplace_back(new AMGLevel);
    levels_[0]->A = A_;
    levels_[0]->index = 0;
    int num_levels = 1;
    bool coarse_solver_exists = (bool)coarse_solver_;
    while (true) {
        AMGLevel &L = *levels_.back();
        Matrix &A = *L.A;
        A.level = num_levels - 1;
        const int rows = A.n;
        // stopping rules use the minimum / the sum over the partitions (src/amg.cu:186-200, 282-356)
        const long long min_part_rows = dist_allreduce_ll(A, rows, 1);
        if (num_levels >= max_levels_ || min_part_rows <= min_coarse_rows_) {
            if (dense_lu_max_rows_ != 0 && min_part_rows > dense_lu_max_rows_) { coarse_solver_.reset(); coarse_solver_exists = false; }
            L.coarsest = true;
            if (!coarse_solver_exists) { L.smoother = make_smoother(); L.smoother->setup(A, false); }
            break;
        }
        // createCoarseVertices
        int n_agg;
        const size_t li = (size_t)num_levels - 1;
        if (li < reuse_aggregates_.size() && (int)reuse_aggregates_[li].size() >= rows) {
            L.aggregates.swap(reuse_aggregates_[li]);       // structure reuse: keep the previous setup's aggregates
            n_agg = reuse_n_coarse_[li];
        } else {
            n_agg = selector_ == "SIZE_4" ? size4_select(A, prm, L.aggregates, s) : size2_select(A, prm, L.aggregates, s);
        }
        L.n_coarse = n_agg;
        const long long N = dist_allreduce_ll(A, rows, 0) * A.by, nextN = dist_allreduce_ll(A, n_agg, 0) * A.by;
        const long long min_part_next = dist_allreduce_ll(A, n_agg, 1);
        bool built_next = false;
        if ((double)nextN <= coarsen_threshold_ * (double)N && nextN != N && min_part_next >= min_coarse_rows_) {
            std::shared_ptr<DistManager> cdist;
            int n_int_c = 0;
            if (A.dist) cdist = dist_coarsen(A, L.aggregates, n_agg, &n_int_c);   // relabels aggregates, appends halo aggregates
            build_restriction(L.aggregates, rows, n_agg, L.R_row_offsets, L.R_column_indices, s);
            std::unique_ptr<AMGLevel> next(new AMGLevel);
            next->owned_A.reset(new Matrix);
            galerkin_aggregation(A, L.aggregates, n_agg, *next->owned_A, s);
            if (cdist) {
                next->owned_A->dist = cdist;
                next->owned_A->n_cols = n_agg + cdist->n_halo;
                next->owned_A->split_row = n_int_c;
                next->owned_A->rsc = A.rsc;
                next->owned_A->bx = A.bx;
                next->owned_A->by = A.by;
                p2p_manager_setup(*next->owned_A);      // peer-memory receive window of the coarse level (collective)
            }
            next->owned_A->compute_diag_and_plan();
            next->A = next->owned_A.get();
            next->index = num_levels;
            const size_t nc = (size_t)next->A->n_cols * A.by;
            L.bc.resize(nc, A.vec_prec);
            L.xc.resize(nc, A.vec_prec);
            L.bc.zero(s);
            L.xc.zero(s);
            L.r.resize((size_t)A.n_cols * A.by, A.vec_prec);
            L.r.zero(s);
            levels_.push_back(std::move(next));
            built_next = true;
        } else {
            L.aggregates.release();
            L.n_coarse = 0;
            L.coarsest = true;
        }
        AMGLevel &Lcur = *levels_[num_levels - 1];
        if (!Lcur.coarsest || !coarse_solver_exists) { Lcur.smoother = make_smoother(); Lcur.smoother->setup(*Lcur.A, false); }
        if (!built_next) break;
        num_levels++;
    }
    // replicated coarse tail below AMGXB_TAIL_ROWS global rows (0 turns it off)
    static const long long tail_rows = getenv("AMGXB_TAIL_ROWS") ? atoll(getenv("AMGXB_TAIL_ROWS")) : 131072;      // r02, N = 2: +9 % (257 -> 282 global it/s), same 77 iterations
    if (tail_rows > 0 && levels_[0]->A->dist) replicate_tail(tail_rows);
    if (coarse_solver_) coarse_solver_->setup(*levels_.back()->A, false);
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
}

namespace {
__global__ void add_offset_kernel(int n, const int *__restrict__ in, int off, int *out)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = in[i] + off;
}
}  // namespace

// Replicated coarse tail.  The hierarchy has been built distributed (aggregates never cross partitions).  From the first
// level whose GLOBAL size is <= tail_rows on, every rank assembles the whole level -- matrix (dist_gather_matrix) and
// aggregates (local ids shifted by the owner's offset on the next level) -- and the rest of the V-cycle runs
// redundantly on every GPU: the same operators, the same aggregates, the same per-row arithmetic as the distributed
// cycle (rows keep their column order, an aggregate keeps its fine rows in ascending order), but no halo exchange below
// the switch; the level above all-gathers its restricted residual instead.
// History: a first version re-aggregated the gathered level globally (SIZE_2 on the assembled matrix).  That was exact
// as an operator but a worse preconditioner on slab-stacked domains: 96x96x384 needs 85 PCG iterations with one global
// SIZE_2 hierarchy and 51 with four per-slab hierarchies (CPU restatement and 4-GPU runs agree), so the tail now keeps
// the partitioned aggregates.  NOT yet run on a device in this form (no GPU minutes left in round 1): off by default.
void AMGSolver::replicate_tail(long long tail_rows)
{
    cudaStream_t s = stream();
    const int nl = (int)levels_.size();
    int t = -1;
    for (int l = 1; l < nl; l++) {
        Matrix &A = *levels_[l]->A;
        if (!A.dist) return;
        if (dist_allreduce_ll(A, A.n, 0) * A.by <= tail_rows) { t = l; break; }
    }
    if (t < 0) return;
    const int rank = levels_[t]->A->dist->rank;
    std::vector<std::vector<int>> counts(nl), offs(nl);
    std::vector<std::unique_ptr<Matrix>> G(nl);
    for (int l = t; l < nl; l++) G[l] = dist_gather_matrix(*levels_[l]->A, counts[l], offs[l]);
    // global aggregates + restriction patterns (needs the still-distributed matrices for the communicator)
    for (int l = t; l + 1 < nl; l++) {
        AMGLevel &L = *levels_[l];
        const int Nl = offs[l].back(), Nnext = offs[l + 1].back();
        DevBuf<int> ag;
        ag.resize((size_t)std::max(Nl, 1));
        ag.zero(s);
        if (L.A->n) add_offset_kernel<<<std::max(1, std::min(ceil_div(L.A->n, 256), 1024)), 256, 0, s>>>(L.A->n, L.aggregates.ptr(), offs[l + 1][rank], ag.ptr() + offs[l][rank]);
        count_launch();
        dist_allgatherv_int_inplace(*L.A, ag.ptr(), counts[l], offs[l], s);
        L.aggregates.swap(ag);
        L.n_coarse = Nnext;
        build_restriction(L.aggregates, Nl, Nnext, L.R_row_offsets, L.R_column_indices, s);
    }
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    // swap in the assembled operators, resize the cycle vectors, rebuild the smoothers
    for (int l = t; l < nl; l++) {
        AMGLevel &L = *levels_[l];
        const int by = L.A->by;
        const Prec vp = L.A->vec_prec;
        L.owned_A = std::move(G[l]);
        L.A = L.owned_A.get();
        L.A->level = l;
        if (l + 1 < nl) {
            const size_t nc = (size_t)offs[l + 1].back() * by;
            L.bc.resize(nc, vp);
            L.xc.resize(nc, vp);
            L.bc.zero(s);
            L.xc.zero(s);
            L.r.resize((size_t)L.A->n_cols * by, vp);
            L.r.zero(s);
        }
        if (L.smoother) { L.smoother = make_smoother(); L.smoother->setup(*L.A, false); }
    }
    AMGLevel &U = *levels_[t - 1];
    const int by = U.A->by;
    U.tail_gather = true;
    U.tail_counts = counts[t];
    U.tail_offs = offs[t];
    U.tail_off = offs[t][rank];
    U.bc.resize((size_t)offs[t].back() * by, U.A->vec_prec);
    U.xc.resize((size_t)offs[t].back() * by, U.A->vec_prec);
    U.bc.zero(s);
    U.xc.zero(s);
}

void AMGSolver::solve_init(DevVec &b, DevVec &x, bool xIsZero)
{
    if (xIsZero && !levels_.empty()) levels_[0]->init_cycle = true;
}

Status AMGSolver::solve_iteration(DevVec &b, DevVec &x, bool xIsZero)
{
    // As the main solver (monitored, many iterations on the same b / x) the cycle is replayed as a CUDA graph; as a
    // preconditioner the caller owns the graph (PCG) or hands in a different vector pair every iteration (FGMRES).
    if (monitor_residual_ && max_iters_ > 2 && !A_->dist) {
        const bool zero = levels_[0]->init_cycle;
        run_segment(zero ? seg_cycle_zero_ : seg_cycle_, b.ptr(), x.ptr(), [&] { levels_[0]->init_cycle = zero; cycle(0, b, x, nullptr); });
    } else {
        cycle(0, b, x, nullptr);
    }
    levels_[0]->init_cycle = false;
    return converged(b, x);
}

bool AMGSolver::solve_fused_dot(DevVec &b, DevVec &x, const ReduceCtx &red, int fin_op, int fin_slot)
{
    // one V-cycle with zero initial guess (what Solver::solve does with max_iters == 1 and no
    // monitoring) whose last finest-level sweep also reduces <b, x>
    if (max_iters_ != 1 || monitor_residual_ || levels_.empty()) return false;
    AMGLevel &L0 = *levels_[0];
    Solver *sm = L0.smoother.get();
    const bool single = L0.coarsest;
    int last_sweeps;
    if (single) last_sweeps = coarse_solver_ ? 0 : coarsest_sweeps_;
    else last_sweeps = (finest_sweeps_ != -1) ? (postsweeps_ == 0 ? 0 : finest_sweeps_) : postsweeps_;
    if (!sm || !sm->supports_fusion() || last_sweeps <= 0) return false;
    SmoothFuse f;
    f.dot_b_x = true;
    f.red = red;
    f.fin_op = fin_op;
    f.fin_slot = fin_slot;
    L0.init_cycle = true;
    cycle(0, b, x, &f);
    L0.init_cycle = false;
    num_iters_ = 1;
    return true;
}

// FixedCycle::cycle (src/cycles/fixed_cycle.cu:25-248) with the V / W / F dispatchers
void AMGSolver::cycle(int lvl, DevVec &b, DevVec &x, const SmoothFuse *top_fuse, int type)
{
    if (type < 0) type = cycle_type_;
    cudaStream_t s = stream();
    AMGLevel &L = *levels_[lvl];
    Matrix &A = *L.A;
    Solver *sm = L.smoother.get();
    const bool finest = (lvl == 0);
    bool x_is_zero = L.init_cycle;
    L.init_cycle = false;

    // ---- pre-smoothing ----
    int n_pre;
    if (L.coarsest && coarse_solver_) n_pre = 0;
    else if (L.coarsest) n_pre = coarsest_sweeps_;
    else if (finest && finest_sweeps_ != -1) n_pre = presweeps_ == 0 ? 0 : finest_sweeps_;
    else {
        n_pre = presweeps_;
        if (presweeps_ != 0 && intensive_smoothing_) n_pre = std::max(n_pre + lvl - 2, 0);
    }
    if (L.coarsest) {
        if (n_pre > 0) {
            SmoothFuse f;
            const SmoothFuse *pf = nullptr;
            if (top_fuse && finest) { f = *top_fuse; pf = &f; }
            sm->smooth(b, x, x_is_zero, n_pre, pf);
        } else if (x_is_zero) {
            x.zero(s);
        }
        if (coarse_solver_) coarse_solver_->solve(b, x, x_is_zero && n_pre == 0);
        phase_mark("coarsest", lvl, s);
        return;
    }
    bool x_virtual_zero = false;   // x holds no data yet and is mathematically zero
    if (n_pre > 0) sm->smooth(b, x, x_is_zero, n_pre, nullptr);
    else if (x_is_zero) x_virtual_zero = true;

    // ---- residual + restriction ----
    const DevVec *rsrc;
    if (x_virtual_zero) {
        rsrc = &b;   // r = b - A*0 = b exactly
    } else {
        dist_exchange_halo(A, x, s);
        CsrOpArgs g;
        g.x = x.ptr();
        g.b = b.ptr();
        g.y = L.r.ptr();
        matrix_apply(A, EPI_RESID, g, s);
        rsrc = &L.r;
    }
    if (algorithm_ == "AGGREGATION") {
        const size_t tail_bytes = (size_t)L.tail_off * A.by * prec_size(A.vec_prec);
        agg_restrict(L.R_row_offsets.ptr(), L.R_column_indices.ptr(), rsrc->ptr(), (char *)L.bc.ptr() + tail_bytes, A.vec_prec, L.n_coarse, A.by, s);
        if (L.tail_gather) dist_allgatherv_inplace(A, L.bc.ptr(), A.vec_prec, A.by, L.tail_counts, L.tail_offs, s);
    } else {
        classical_restrict(L, *rsrc, s);
    }

    phase_mark("down: pre-smooth, residual, restrict", lvl, s);
    // ---- coarse-grid correction ----
    // V: one cycle on the next level; W: two W cycles; F: a W cycle then a V cycle (src/cycles/{v,w,f}_cycle.cu).  When the
    // next level is the coarsest a single fixed cycle is launched whatever the type (fixed_cycle.cu:169-179).  The second
    // visit continues from the xc the first one left (its init flag has been cleared).
    levels_[lvl + 1]->init_cycle = true;
    // AMGXB_GRAPH_COARSE (default on since r02): when the cycle is a preconditioner whose caller cannot capture it (FGMRES hands in a
    // different vector pair every iteration), everything below the finest level still works on fixed buffers (bc, xc of level 0) and
    // is replayed as one CUDA graph: the launch-latency-bound tail of the hierarchy costs one graph launch.
    static const bool graph_coarse = getenv("AMGXB_GRAPH_COARSE") ? atoi(getenv("AMGXB_GRAPH_COARSE")) != 0 : true;      // r02: parity green (33 tests), +2 % at 256^3, +10 % at 128^3 on config 3
    bool capturing = false;
    if (graph_coarse && finest && type == CYC_V) {
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &st) != cudaSuccess) { cudaGetLastError(); st = cudaStreamCaptureStatusActive; }
        capturing = (st != cudaStreamCaptureStatusNone);
    }
    if (graph_coarse && finest && type == CYC_V && !capturing && !levels_[1]->A->dist) {
        run_segment(seg_coarse_, L.bc.ptr(), L.xc.ptr(), [&] { levels_[1]->init_cycle = true; cycle(1, L.bc, L.xc, nullptr, CYC_V); });
        levels_[1]->init_cycle = false;
    } else if (type == CYC_V || levels_[lvl + 1]->coarsest) {
        cycle(lvl + 1, L.bc, L.xc, nullptr, CYC_V);
    } else if (type == CYC_W) {
        cycle(lvl + 1, L.bc, L.xc, nullptr, CYC_W);
        cycle(lvl + 1, L.bc, L.xc, nullptr, CYC_W);
    } else if (type == CYC_F) {
        cycle(lvl + 1, L.bc, L.xc, nullptr, CYC_W);
        cycle(lvl + 1, L.bc, L.xc, nullptr, CYC_V);
    } else {
        cg_cycle_dispatch(lvl + 1, L.bc, L.xc, type == CYC_CGF);
    }

    // ---- prolongation + post-smoothing ----
    int n_post;
    if (finest && finest_sweeps_ != -1) n_post = postsweeps_ == 0 ? 0 : finest_sweeps_;
    else {
        n_post = postsweeps_;
        if (postsweeps_ != 0 && intensive_smoothing_) n_post = std::max(n_post + lvl - 2, 0);
    }
    SmoothFuse f;
    bool have_fuse = false, in_alt = false;
    if (top_fuse && finest && n_post > 0) { f = *top_fuse; have_fuse = true; }
    if (algorithm_ == "AGGREGATION" && error_scaling_ >= 2) {
        // x += lambda * (smoothed P xc), lambda from the residual the restriction was computed from (b itself when x is still zero)
        if (x_virtual_zero) x.zero(s);
        scaled_correction(L, *rsrc, x);
    } else if (algorithm_ == "AGGREGATION") {
        static const bool fuse_prolong = getenv("AMGXB_FUSE_PROLONG") ? atoi(getenv("AMGXB_FUSE_PROLONG")) != 0 : false;
        // where the smoother wants its initial iterate so that n_post sweeps end in x without a copy
        void *xin = (n_post > 0) ? sm->smooth_input(x, n_post) : x.ptr();
        in_alt = (xin != x.ptr());
        if (x_virtual_zero && n_post > 0 && sm->supports_fusion() && fuse_prolong && !A.dist) {
            // x := P xc is read on the fly by the first sweep
            f.agg = L.aggregates.ptr();
            f.xc = (const char *)L.xc.ptr() + (size_t)L.tail_off * A.by * prec_size(A.vec_prec);
            have_fuse = true;
            in_alt = false;
        } else if (x_virtual_zero) {
            agg_prolong_set(L.aggregates.ptr(), (const char *)L.xc.ptr() + (size_t)L.tail_off * A.by * prec_size(A.vec_prec), xin, A.vec_prec, A.n, A.by, s);               // x = 0 + P xc
        } else {
            agg_prolong_add(L.aggregates.ptr(), (const char *)L.xc.ptr() + (size_t)L.tail_off * A.by * prec_size(A.vec_prec), x.ptr(), xin, A.vec_prec, A.n, A.by, s);      // xin = x + P xc
        }
    } else {
        void *xin = (n_post > 0) ? sm->smooth_input(x, n_post) : x.ptr();
        in_alt = (xin != x.ptr());
        classical_prolong_add(L, x_virtual_zero ? nullptr : x.ptr(), xin, s);   // xin = x + P xc  (0 + P xc == P xc exactly)
    }
    if (n_post > 0) sm->smooth(b, x, false, n_post, have_fuse ? &f : nullptr, in_alt);
    phase_mark("up: prolong, post-smooth", lvl, s);
}

double AMGSolver::host_dot(const DevVec &x, const DevVec &y, size_t n, const Matrix *over)
{
    cudaStream_t s = stream();
    ReduceCtx red = red_ctx();
    vec_dot(x.ptr(), y.ptr(), x.prec, n, red, FIN_STORE, S_TMP0, 0, s);
    double h = 0;
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(&h, red.scal + S_TMP0, sizeof(double), cudaMemcpyDeviceToHost, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    if (over && over->dist) dist_allreduce_host(*over, &h, 1, 0);      // owned rows of every rank
    return h;
}

// CG_CycleDispatcher / CG_Flex_CycleDispatcher::dispatch (src/cycles/cg_cycle.cu:18-101, cg_flex_cycle.cu:18-103): cycle_iters
// iterations of (flexible) PCG on level `lvl`, preconditioned by one CG(F) fixed cycle of that level started from zero.  The
// recurrence scalars live on the host as in the reference (dotc).
void AMGSolver::cg_cycle_dispatch(int lvl, DevVec &b, DevVec &x, bool flex)
{
    cudaStream_t s = stream();
    AMGLevel &L = *levels_[lvl];
    Matrix &A = *L.A;
    const size_t n = (size_t)A.n * A.by, N = (size_t)A.n_cols * A.by;
    const Prec vp = A.vec_prec;
    const int type = flex ? CYC_CGF : CYC_CG;
    for (DevVec *v : {&L.cg_y, &L.cg_z, &L.cg_r, &L.cg_p, &L.cg_d})
        if (v->n != N) { v->resize(N, vp); v->zero(s); }
    DevVec &y = L.cg_y, &z = L.cg_z, &r = L.cg_r, &p = L.cg_p, &d = L.cg_d;
    auto apply = [&](DevVec &in, DevVec &out) {
        dist_exchange_halo(A, in, s);
        CsrOpArgs g;
        g.x = in.ptr();
        g.y = out.ptr();
        matrix_apply(A, EPI_SPMV, g, s);
    };
    if (L.init_cycle) {
        x.zero(s);
        L.init_cycle = false;
    }
    apply(x, y);                                                   // y = A x
    vec_axpby(b.ptr(), y.ptr(), r.ptr(), vp, n, 1.0, -1.0, s);     // r = b - y
    L.init_cycle = true;
    cycle(lvl, r, z, nullptr, type);                               // z = M r
    vec_copy(p.ptr(), z.ptr(), vp, n, s);
    double rz = flex ? 0.0 : host_dot(r, z, n, &A);
    int k = 0;
    while (true) {
        apply(p, y);
        if (flex) rz = host_dot(r, z, n, &A);
        const double alpha = rz / host_dot(y, p, n, &A);
        vec_axpy(p.ptr(), x.ptr(), vp, n, alpha, s);
        if (++k == cycle_iters_) break;
        if (flex) vec_copy(d.ptr(), r.ptr(), vp, n, s);
        vec_axpy(y.ptr(), r.ptr(), vp, n, alpha * -1.0, s);
        if (flex) vec_axpby(r.ptr(), d.ptr(), d.ptr(), vp, n, 1.0, -1.0, s);
        L.init_cycle = true;
        cycle(lvl, r, z, nullptr, type);
        double beta;
        if (flex) beta = host_dot(z, d, n, &A) / rz;
        else {
            const double rz_old = rz;
            rz = host_dot(r, z, n, &A);
            beta = rz / rz_old;
        }
        vec_axpby(z.ptr(), p.ptr(), p.ptr(), vp, n, 1.0, beta, s);
    }
}

// prolongateAndApplyCorrection with error_scaling = 2 (lambda = <r, A e> / <A e, A e>) or 3 (lambda = <r, e> / <e, A e>), e = P xc
// smoothed scaling_smoother_steps times against the residual (aggregation_amg_level.cu:700-824).  The two inner products and the
// clamped quotient stay on the device; the scale is kept for reuse_scale further corrections.
void AMGSolver::scaled_correction(AMGLevel &L, const DevVec &rf, DevVec &x)
{
    cudaStream_t s = stream();
    Matrix &A = *L.A;
    const size_t n = (size_t)A.n * A.by, N = (size_t)A.n_cols * A.by;
    const Prec vp = A.vec_prec;
    if (L.ef.n != N) {
        L.ef.resize(N, vp);
        L.ef.zero(s);
        L.Aef.resize(N, vp);
        L.Aef.zero(s);
        L.scale.resize(1);
        L.scale.zero(s);
        L.scale_counter = 0;
    }
    const char *xc = (const char *)L.xc.ptr() + (size_t)L.tail_off * A.by * prec_size(vp);
    agg_prolong_set(L.aggregates.ptr(), xc, L.ef.ptr(), vp, A.n, A.by, s);      // ef = P xc
    if (L.scale_counter > 0) {
        vec_axpy_dev(L.ef.ptr(), x.ptr(), vp, n, L.scale.ptr(), 0, 1.0, s);
        L.scale_counter--;
        return;
    }
    if (scaling_smoother_steps_ > 0) L.smoother->smooth(const_cast<DevVec &>(rf), L.ef, false, scaling_smoother_steps_, nullptr);
    dist_exchange_halo(A, L.ef, s);
    {
        CsrOpArgs g;
        g.x = L.ef.ptr();
        g.y = L.Aef.ptr();
        matrix_apply(A, EPI_SPMV, g, s);
    }
    ReduceCtx red = red_ctx();
    if (error_scaling_ == 2) {
        vec_dot(rf.ptr(), L.Aef.ptr(), vp, n, red, FIN_STORE, S_TMP0, 0, s);
        vec_dot(L.Aef.ptr(), L.Aef.ptr(), vp, n, red, FIN_STORE, S_TMP1, 0, s);
    } else {
        vec_dot(rf.ptr(), L.ef.ptr(), vp, n, red, FIN_STORE, S_TMP0, 0, s);
        vec_dot(L.ef.ptr(), L.Aef.ptr(), vp, n, red, FIN_STORE, S_TMP1, 0, s);
    }
    if (A.dist) {
        dist_allreduce_scalar_fin(A, red, S_TMP0, FIN_STORE, s);
        dist_allreduce_scalar_fin(A, red, S_TMP1, FIN_STORE, s);
    }
    scalar_error_scale(red.scal, S_TMP0, S_TMP1, L.scale.ptr(), s);
    vec_axpy_dev(L.ef.ptr(), x.ptr(), vp, n, L.scale.ptr(), 0, 1.0, s);        // x += lambda e
    L.scale_counter = reuse_scale_;
}

// print_grid_stats of the reference (src/amg.cu:1231-1350): same table layout
void AMGSolver::print_grid_stats()
{
    std::stringstream ss;
    const int nl = (int)levels_.size();
    long long total_rows = 0, total_nnz = 0;
    for (auto &l : levels_) { total_rows += l->A->n; total_nnz += l->A->nnz + (l->A->has_ext_diag ? l->A->n : 0); }
    ss << "AMG Grid:\n";
    ss << "         Number of Levels: " << nl << "\n";
    ss << "            LVL         ROWS               NNZ  PARTS    SPRSTY       Mem (GB)\n";
    ss << "         ----------------------------------------------------------------------\n";
    for (int i = 0; i < nl; i++) {
        const Matrix &M = *levels_[i]->A;
        const long long nnz = M.nnz + (M.has_ext_diag ? M.n : 0);
        const double sp = M.n ? (double)nnz / ((double)M.n * (double)M.n) : 0.0;
        const double mem = (double)(M.row_ptr.size() * 4 + M.col_idx.size() * 4 + M.values.nbytes()) / (1024.0 * 1024.0 * 1024.0);
        ss << std::setw(12) << i << "(D)" << std::setw(13) << M.n << std::setw(18) << nnz << std::setw(7) << 1
           << std::setw(10) << std::setprecision(3) << std::scientific << sp << std::setw(15) << std::setprecision(3) << std::scientific << mem << "\n";
    }
    ss << "         ----------------------------------------------------------------------\n";
    const Matrix &F = *levels_[0]->A;
    ss << std::fixed << std::setprecision(5);
    ss << "         Grid Complexity: " << (double)total_rows / std::max(1, F.n) << "\n";
    ss << "         Operator Complexity: " << (double)total_nnz / std::max(1ll, (long long)F.nnz + (F.has_ext_diag ? F.n : 0)) << "\n";
    ss << "         Total Memory Usage: " << device_mem_used_gb() << " GB\n";
    ss << "         ----------------------------------------------------------------------\n";
    amgx_output(ss.str().c_str(), (int)ss.str().length());
}

}  // namespace amgxb
