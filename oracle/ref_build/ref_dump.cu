// ref_dump.cu -- TEST INFRASTRUCTURE.  Drives the UNMODIFIED reference (oracle/_ref/libamgx_ref.so, built
// from /root/reference by oracle/ref_build/Makefile) through its own C API on a system we hand it and
// dumps what the parity tests pin: the residual history, the solution, and -- by peeking at the
// reference's internal objects through its own headers -- every level of the AMG hierarchy
// (matrices, aggregates / R pattern or P / R, smoother data, colouring).
//
// usage: ref_dump <system.bin> <config.json> <out.bin> [mode=dDDI] [reps=1]
// system.bin : int32 n, nnz, bx, by, has_diag, has_x0 ; row_ptr[n+1] ; col[nnz] ; val f64[nnz*bx*by] ;
//              (diag f64[n*bx*by]) ; rhs f64[n*by] ; (x0 f64[n*bx])
// out.bin    : records  u32 name_len | name | u8 dtype('i','d','f') | u64 count | raw data
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <map>
#include <memory>
#include <sstream>
#include <iostream>
#include <fstream>
#include <algorithm>
#include <typeinfo>
#include <cuda_runtime.h>
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#define private public
#define protected public
#include <amgx_c.h>
#include <amg_solver.h>
#include <amg.h>
#include <amg_level.h>
#include <solvers/solver.h>
#include <solvers/pcg_solver.h>
#include <solvers/fgmres_solver.h>
#include <solvers/algebraic_multigrid_solver.h>
#include <solvers/block_jacobi_solver.h>
#include <solvers/jacobi_l1_solver.h>
#include <solvers/multicolor_dilu_solver.h>
#include <aggregation/aggregation_amg_level.h>
#include <classical/classical_amg_level.h>
#include <classical/strength/strength.h>
#include <classical/selectors/selector.h>
#include <classical/interpolators/interpolator.h>
#include <truncate.h>
#include <matrix_coloring/matrix_coloring.h>
#include <amgx_c_common.h>
#include <amgx_c_wrappers.inl>
#undef private
#undef protected

using namespace amgx;

static FILE *g_out = nullptr;
static void rec(const std::string &name, char dt, const void *data, uint64_t count, size_t elem)
{
    uint32_t nl = (uint32_t)name.size();
    fwrite(&nl, 4, 1, g_out);
    fwrite(name.data(), 1, nl, g_out);
    fwrite(&dt, 1, 1, g_out);
    fwrite(&count, 8, 1, g_out);
    if (count) fwrite(data, elem, count, g_out);
}
template <class V> static void rec_ivec(const std::string &name, const V &v, size_t count)
{
    std::vector<int> h(count);
    if (count) cudaMemcpy(h.data(), v.raw(), count * sizeof(int), cudaMemcpyDefault);
    rec(name, 'i', h.data(), count, 4);
}
template <class V> static void rec_dvec(const std::string &name, const V &v, size_t count)
{
    typedef typename V::value_type T;
    std::vector<T> h(count);
    if (count) cudaMemcpy(h.data(), v.raw(), count * sizeof(T), cudaMemcpyDefault);
    rec(name, sizeof(T) == 8 ? 'd' : 'f', h.data(), count, sizeof(T));
}

static std::string g_log;
static void print_cb(const char *msg, int length) { g_log.append(msg, length); fwrite(msg, 1, length, stdout); }

template <AMGX_Mode CASE> static void dump_hierarchy(AMGX_solver_handle slv)
{
    typedef typename TemplateMode<CASE>::Type TConfig;
    typedef TConfig TConfig_d;
    AMG_Solver<TConfig> *as = get_mode_object_from<CASE, AMG_Solver, AMGX_solver_handle>(slv);
    Solver<TConfig> *top = as->getSolverObject();
    Solver<TConfig> *cur = top;
    AlgebraicMultigrid_Solver<TConfig> *ams = nullptr;
    for (int depth = 0; depth < 4 && cur; depth++) {
        if ((ams = dynamic_cast<AlgebraicMultigrid_Solver<TConfig> *>(cur))) break;
        if (auto *p = dynamic_cast<PCG_Solver<TConfig> *>(cur)) { cur = p->m_preconditioner; continue; }
        if (auto *f = dynamic_cast<FGMRES_Solver<TConfig> *>(cur)) { cur = f->m_preconditioner; continue; }
        break;
    }
    // top-level smoother-only solvers (no AMG): dump their data as level 0
    auto dump_smoother = [&](Solver<TConfig> *sm, const std::string &pfx) {
        if (!sm) return;
        if (auto *bj = dynamic_cast<block_jacobi_solver::BlockJacobiSolver_Base<TConfig> *>(sm)) rec_dvec(pfx + "Dinv", bj->Dinv, bj->Dinv.size());
        if (auto *l1 = dynamic_cast<JacobiL1Solver_Base<TConfig> *>(sm)) rec_dvec(pfx + "l1_d", l1->m_d, l1->m_d.size());
        if (auto *di = dynamic_cast<multicolor_dilu_solver::MulticolorDILUSolver_Base<TConfig> *>(sm)) rec_dvec(pfx + "Einv", di->Einv, di->Einv.size());
    };
    auto dump_matrix = [&](Matrix<TConfig> &A, const std::string &pfx) {
        int n = A.get_num_rows(), nnz = A.get_num_nz(), bs = A.get_block_size();
        int info[6] = {n, nnz, A.get_block_dimx(), A.get_block_dimy(), A.hasProps(DIAG) ? 1 : 0, (int)A.get_num_cols()};
        rec(pfx + "info", 'i', info, 6, 4);
        rec_ivec(pfx + "row_offsets", A.row_offsets, n + 1);
        rec_ivec(pfx + "col_indices", A.col_indices, nnz);
        rec_dvec(pfx + "values", A.values, (size_t)(nnz + (A.hasProps(DIAG) ? n : 0)) * bs);
        rec_ivec(pfx + "diag", A.diag, std::min<size_t>(A.diag.size(), n));
        if (A.hasProps(COLORING)) {
            const typename Matrix<TConfig>::IVector &rc = A.getMatrixColoring().getRowColors();
            int nc = A.getMatrixColoring().getNumColors();
            rec(pfx + "num_colors", 'i', &nc, 1, 4);
            rec_ivec(pfx + "row_colors", rc, std::min<size_t>(rc.size(), n));
        }
    };
    if (!ams) {
        dump_smoother(cur, "L0.");
        return;
    }
    auto &amg = ams->m_amg;
    int nl = amg.num_levels;
    rec("num_levels", 'i', &nl, 1, 4);
    AMG_Level<TConfig_d> *lvl = amg.getFinestLevel(cusp::device_memory());
    int li = 0;
    while (lvl) {
        std::string pfx = "L" + std::to_string(li) + ".";
        dump_matrix(lvl->getA(), pfx);
        dump_smoother(lvl->getSmoother(), pfx);
        if (auto *ag = dynamic_cast<aggregation::Aggregation_AMG_Level_Base<TConfig_d> *>(lvl)) {
            if (lvl->getNextLevel(cusp::device_memory())) {
                int na = ag->m_num_aggregates;
                rec(pfx + "num_aggregates", 'i', &na, 1, 4);
                rec_ivec(pfx + "aggregates", ag->m_aggregates, lvl->getA().get_num_rows());
                rec_ivec(pfx + "R_row_offsets", ag->m_R_row_offsets, ag->m_R_row_offsets.size());
                rec_ivec(pfx + "R_column_indices", ag->m_R_column_indices, ag->m_R_column_indices.size());
            }
        }
        if (auto *cl = dynamic_cast<classical::Classical_AMG_Level_Base<TConfig_d> *>(lvl)) {
            if (lvl->getNextLevel(cusp::device_memory())) {
                dump_matrix(cl->P, pfx + "P.");
                dump_matrix(cl->R, pfx + "R.");
                rec_ivec(pfx + "cf_map", cl->m_cf_map, cl->m_cf_map.size());
            }
        }
        lvl = lvl->getNextLevel(cusp::device_memory());
        li++;
    }
}


// Classical AMG stage dump (level 0): runs the reference's own strength / selector / interpolator / truncation objects
// on the uploaded matrix, outside the solver, so that the C/F map (which the level discards) and the untruncated P can
// be pinned.  Enabled by REFDUMP_CLASSICAL="strength_threshold,max_row_sum,interp_max_elements".
template <AMGX_Mode CASE> static void dump_classical_stages(AMGX_matrix_handle mtx, double theta, double max_row_sum, int max_elmts)
{
    typedef typename TemplateMode<CASE>::Type TConfig;
    typedef Vector<typename TConfig::template setVecPrec<AMGX_vecInt>::Type> IVector;
    typedef Vector<typename TConfig::template setVecPrec<AMGX_vecBool>::Type> BVector;
    typedef Vector<typename TConfig::template setVecPrec<AMGX_vecFloat>::Type> FVector;
    Matrix<TConfig> &A = *get_mode_object_from<CASE, Matrix, AMGX_matrix_handle>(mtx);
    const int n = A.get_num_rows(), nnz = A.get_num_nz();
    auto dump_csr = [&](Matrix<TConfig> &M, const std::string &pfx) {
        int mn = M.get_num_rows(), mnnz = M.get_num_nz();
        int info[3] = {mn, mnnz, (int)M.get_num_cols()};
        rec(pfx + "info", 'i', info, 3, 4);
        rec_ivec(pfx + "row_offsets", M.row_offsets, mn + 1);
        rec_ivec(pfx + "col_indices", M.col_indices, mnnz);
        rec_dvec(pfx + "values", M.values, mnnz);
    };
    char buf[256];
    snprintf(buf, sizeof(buf), "strength_threshold=%.17g, max_row_sum=%.17g, strength=AHAT", theta, max_row_sum);
    AMG_Config c;
    c.parseParameterString(buf);
    Strength<TConfig> *st = StrengthFactory<TConfig>::allocate(c, "default");
    for (int aggressive = 0; aggressive < 2; aggressive++) {
        const std::string pfx = aggressive ? "stage.aggr." : "stage.pmis.";
        AMG_Config cs;
        cs.parseParameterString(aggressive ? "selector=AGGRESSIVE_PMIS" : "selector=PMIS");
        classical::Selector<TConfig> *sel = classical::SelectorFactory<TConfig>::allocate(cs, "default");
        BVector s_con(nnz);
        FVector w(n);
        IVector cf(n), scratch(n);
        thrust_wrapper::fill<TConfig::memSpace>(s_con.begin(), s_con.end(), false);
        thrust_wrapper::fill<TConfig::memSpace>(w.begin(), w.end(), 0.0f);
        thrust_wrapper::fill<TConfig::memSpace>(cf.begin(), cf.end(), 0);
        thrust_wrapper::fill<TConfig::memSpace>(scratch.begin(), scratch.end(), 0);
        st->computeStrongConnectionsAndWeights(A, s_con, w, max_row_sum);
        if (!aggressive) {
            std::vector<char> hb(nnz);
            cudaMemcpy(hb.data(), s_con.raw(), nnz, cudaMemcpyDefault);
            std::vector<int> hi(hb.begin(), hb.end());
            rec("stage.s_con", 'i', hi.data(), nnz, 4);
            rec_dvec("stage.weights", w, n);
        }
        sel->markCoarseFinePoints(A, w, s_con, cf, scratch);
        rec_ivec(pfx + "cf_map", cf, n);
        int nc = 0;
        sel->renumberAndCountCoarsePoints(cf, nc, n);
        rec(pfx + "num_coarse", 'i', &nc, 1, 4);
        AMG_Config ci;
        ci.parseParameterString(aggressive ? "interpolator=MULTIPASS" : "interpolator=D2");
        Interpolator<TConfig> *ip = InterpolatorFactory<TConfig>::allocate(ci, "default");
        Matrix<TConfig> P;
        ip->generateInterpolationMatrix(A, cf, s_con, scratch, P);
        cudaDeviceSynchronize();
        dump_csr(P, pfx + "P.");
        if (max_elmts > 0) {
            Truncate<TConfig>::truncateByMaxElements(P, max_elmts);
            cudaDeviceSynchronize();
            dump_csr(P, pfx + "Ptrunc.");
        }
        delete ip;
        delete sel;
    }
    delete st;
}

#define CK(x)                                                                   \
    do {                                                                        \
        AMGX_RC _rc = (x);                                                      \
        if (_rc != AMGX_RC_OK) {                                                \
            char msg[1024];                                                     \
            AMGX_get_error_string(_rc, msg, 1024);                              \
            fprintf(stderr, "ref_dump: %s failed: %s\n", #x, msg);              \
            exit(2);                                                            \
        }                                                                       \
    } while (0)

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: ref_dump system.bin config.json out.bin [mode] [reps]\n"); return 1; }
    std::string mode_s = argc > 4 ? argv[4] : "dDDI";
    int reps = argc > 5 ? atoi(argv[5]) : 1;
    AMGX_Mode mode = mode_s == "dDDI" ? AMGX_mode_dDDI : mode_s == "dDFI" ? AMGX_mode_dDFI : mode_s == "dFFI" ? AMGX_mode_dFFI : AMGX_mode_hDDI;
    const bool mat32 = (mode == AMGX_mode_dDFI || mode == AMGX_mode_dFFI), vec32 = (mode == AMGX_mode_dFFI);
    int n = 0, nnz = 0, bx = 1, by = 1, has_diag = 0, has_x0 = 0;
    std::vector<int> rp, ci;
    std::vector<double> va, dg, rhs, x0;
    bool dump_levels = getenv("REFDUMP_NO_LEVELS") == nullptr;   // timing runs on large systems skip the hierarchy dump
    if (!strncmp(argv[1], "poisson:", 8)) {
        // generated 7-point Poisson nx^3 (diagonal first, then -1 for i-1,i+1,j-1,j+1,k-1,k+1), b = 1, x0 = 0; timing runs: no hierarchy dump
        const int nx = atoi(argv[1] + 8);
        dump_levels = nx <= 64;
        n = nx * nx * nx;
        rp.resize((size_t)n + 1);
        ci.reserve((size_t)n * 7);
        va.reserve((size_t)n * 7);
        for (int r = 0; r < n; r++) {
            const int i = r % nx, j = (r / nx) % nx, k = r / (nx * nx);
            rp[r] = (int)ci.size();
            ci.push_back(r); va.push_back(6.0);
            if (i > 0) { ci.push_back(r - 1); va.push_back(-1.0); }
            if (i < nx - 1) { ci.push_back(r + 1); va.push_back(-1.0); }
            if (j > 0) { ci.push_back(r - nx); va.push_back(-1.0); }
            if (j < nx - 1) { ci.push_back(r + nx); va.push_back(-1.0); }
            if (k > 0) { ci.push_back(r - nx * nx); va.push_back(-1.0); }
            if (k < nx - 1) { ci.push_back(r + nx * nx); va.push_back(-1.0); }
        }
        rp[n] = (int)ci.size();
        nnz = (int)ci.size();
        rhs.assign((size_t)n, 1.0);
        x0.assign((size_t)n, 0.0);
    } else {
        FILE *f = fopen(argv[1], "rb");
        if (!f) { perror("system"); return 1; }
        int hdr[6];
        if (fread(hdr, 4, 6, f) != 6) return 1;
        n = hdr[0]; nnz = hdr[1]; bx = hdr[2]; by = hdr[3]; has_diag = hdr[4]; has_x0 = hdr[5];
        rp.resize(n + 1); ci.resize(nnz);
        va.resize((size_t)nnz * bx * by); rhs.resize((size_t)n * by); x0.assign((size_t)n * bx, 0.0);
        fread(rp.data(), 4, n + 1, f);
        fread(ci.data(), 4, nnz, f);
        fread(va.data(), 8, va.size(), f);
        if (has_diag) { dg.resize((size_t)n * bx * by); fread(dg.data(), 8, dg.size(), f); }
        fread(rhs.data(), 8, rhs.size(), f);
        if (has_x0) fread(x0.data(), 8, x0.size(), f);
        fclose(f);
    }
    std::vector<float> vaf, dgf, rhsf, x0f;
    if (mat32) { vaf.assign(va.begin(), va.end()); dgf.assign(dg.begin(), dg.end()); }
    if (vec32) { rhsf.assign(rhs.begin(), rhs.end()); x0f.assign(x0.begin(), x0.end()); }

    g_out = fopen(argv[3], "wb");
    if (!g_out) { perror("out"); return 1; }
    CK(AMGX_initialize());
    CK(AMGX_register_print_callback(&print_cb));
    AMGX_config_handle cfg;
    CK(AMGX_config_create_from_file(&cfg, argv[2]));
    AMGX_resources_handle rsrc;
    CK(AMGX_resources_create_simple(&rsrc, cfg));
    AMGX_matrix_handle A;
    AMGX_vector_handle b, x;
    AMGX_solver_handle solver;
    CK(AMGX_matrix_create(&A, rsrc, mode));
    CK(AMGX_vector_create(&b, rsrc, mode));
    CK(AMGX_vector_create(&x, rsrc, mode));
    CK(AMGX_solver_create(&solver, rsrc, mode, cfg));
    CK(AMGX_matrix_upload_all(A, n, nnz, bx, by, rp.data(), ci.data(), mat32 ? (void *)vaf.data() : (void *)va.data(),
                              has_diag ? (mat32 ? (void *)dgf.data() : (void *)dg.data()) : nullptr));
    CK(AMGX_vector_upload(b, n, by, vec32 ? (void *)rhsf.data() : (void *)rhs.data()));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    CK(AMGX_solver_setup(solver, A));
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float setup_ms = 0;
    cudaEventElapsedTime(&setup_ms, e0, e1);
    double solve_ms_best = 1e30;
    int nit = 0;
    for (int r = 0; r < reps; r++) {
        CK(AMGX_vector_upload(x, n, bx, vec32 ? (void *)x0f.data() : (void *)x0.data()));
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        CK(AMGX_solver_solve(solver, b, x));
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        solve_ms_best = std::min<double>(solve_ms_best, ms);
    }
    AMGX_SOLVE_STATUS st;
    CK(AMGX_solver_get_status(solver, &st));
    CK(AMGX_solver_get_iterations_number(solver, &nit));
    int sti = (int)st;
    rec("status", 'i', &sti, 1, 4);
    rec("iterations", 'i', &nit, 1, 4);
    double tm[2] = {setup_ms * 1e-3, solve_ms_best * 1e-3};
    rec("times", 'd', tm, 2, 8);
    std::vector<double> hist;
    for (int i = 0; i <= nit; i++) {
        for (int c = 0; c < by; c++) {
            double r;
            if (AMGX_solver_get_iteration_residual(solver, i, c, &r) != AMGX_RC_OK) { c = by; break; }
            hist.push_back(r);
        }
    }
    rec("res_history", 'd', hist.data(), hist.size(), 8);
    if (!dump_levels) {
    } else if (vec32) {
        std::vector<float> xs((size_t)n * bx);
        CK(AMGX_vector_download(x, xs.data()));
        rec("solution", 'f', xs.data(), xs.size(), 4);
    } else {
        std::vector<double> xs((size_t)n * bx);
        CK(AMGX_vector_download(x, xs.data()));
        rec("solution", 'd', xs.data(), xs.size(), 8);
    }
    // one SpMV through the public API: y = A * rhs
    if (dump_levels) {
        AMGX_vector_handle y;
        CK(AMGX_vector_create(&y, rsrc, mode));
        CK(AMGX_vector_set_zero(y, n, by));
        CK(AMGX_matrix_vector_multiply(A, b, y));
        if (vec32) { std::vector<float> ys((size_t)n * by); CK(AMGX_vector_download(y, ys.data())); rec("spmv_A_rhs", 'f', ys.data(), ys.size(), 4); }
        else { std::vector<double> ys((size_t)n * by); CK(AMGX_vector_download(y, ys.data())); rec("spmv_A_rhs", 'd', ys.data(), ys.size(), 8); }
        AMGX_vector_destroy(y);
    }
    if (!dump_levels) { /* timing run */ }
    else if (mode == AMGX_mode_dDDI) dump_hierarchy<AMGX_mode_dDDI>(solver);
    else if (mode == AMGX_mode_dDFI) dump_hierarchy<AMGX_mode_dDFI>(solver);
    else if (mode == AMGX_mode_dFFI) dump_hierarchy<AMGX_mode_dFFI>(solver);
    if (dump_levels && mode == AMGX_mode_dDDI && getenv("REFDUMP_CLASSICAL")) {
        double th = 0.25, mrs = 1.1;
        int me = -1;
        sscanf(getenv("REFDUMP_CLASSICAL"), "%lf,%lf,%d", &th, &mrs, &me);
        dump_classical_stages<AMGX_mode_dDDI>(A, th, mrs, me);
    }
    rec("log", 'i', nullptr, 0, 4);
    {
        uint32_t nl = 8;
        fwrite(&nl, 4, 1, g_out);
        fwrite("log_text", 1, 8, g_out);
        char dt = 'c';
        fwrite(&dt, 1, 1, g_out);
        uint64_t cnt = g_log.size();
        fwrite(&cnt, 8, 1, g_out);
        fwrite(g_log.data(), 1, cnt, g_out);
    }
    fclose(g_out);
    AMGX_solver_destroy(solver);
    AMGX_vector_destroy(x);
    AMGX_vector_destroy(b);
    AMGX_matrix_destroy(A);
    AMGX_resources_destroy(rsrc);
    AMGX_config_destroy(cfg);
    AMGX_finalize();
    printf("ref_dump: status %d iterations %d setup %.6f s solve %.6f s\n", sti, nit, tm[0], tm[1]);
    return 0;
}
