// capi2.cu -- remaining C-ABI entry points: system IO, generators, distributed uploads, bench hooks.
#include "capi_internal.h"
#include <fstream>
#include <sstream>
#include <algorithm>
#include <map>

using namespace amgxb;

namespace amgxb {

template <class H> static H *chk(void *p, unsigned magic, const char *what)
{
    H *h = reinterpret_cast<H *>(p);
    if (!h || h->magic != magic) fatal(AMGX_RC_BAD_PARAMETERS, std::string("invalid ") + what + " handle");
    return h;
}

static AMGX_RC on_exception(const char *where)
{
    AMGX_RC rc = AMGX_RC_UNKNOWN;
    std::string msg;
    try { throw; }
    catch (const Error &e) { rc = e.rc; msg = e.msg; }
    catch (const std::bad_alloc &) { rc = AMGX_RC_NO_MEMORY; msg = "Not enough memory"; }
    catch (const std::exception &e) { msg = e.what(); }
    catch (...) { msg = "unknown exception"; }
    std::string full = std::string("AMGX error in ") + where + ": " + msg + "\n";
    amgx_output(full.c_str(), (int)full.size());
    return rc;
}
#define API2_BEGIN try {
#define API2_END } catch (...) { return on_exception(__func__); } return AMGX_RC_OK;

void residual_norm_external(SolverH &h, Matrix &A, Vector &b, Vector &x, std::vector<double> &nrm)
{
    // r = b - A x ; norm per the solver's norm type (Solver::compute_residual_norm_external, solver.cu:225-246)
    cudaStream_t s = A.stream();
    dist_prepare_vector(A, b);
    dist_prepare_vector(A, x);
    DevVec r;
    r.resize((size_t)A.n_cols * A.by, A.vec_prec);
    dist_exchange_halo(A, x.data, s);
    CsrOpArgs g;
    g.x = x.data.ptr();
    g.b = b.data.ptr();
    g.y = r.ptr();
    matrix_apply(A, EPI_RESID, g, s);
    solver_norm_of(*h.solver, r, nrm);
}

// ---------------------------------------------------------------------------------------------
// MatrixMarket (+ the "%%AMGX" extension header: block dims, rhs, solution, diagonal) reader.
// Covers the layouts of src/readers.cu:700-1100 that the examples use: coordinate real
// general|symmetric, optional block size, optional trailing rhs / solution sections.
// ---------------------------------------------------------------------------------------------
struct MMSystem {
    int n = 0, nnz = 0, bx = 1, by = 1;
    std::vector<int> rp, ci;
    std::vector<double> va, diag, rhs, sol;
    bool has_diag = false;
};

static void read_mm(const char *filename, MMSystem &S)
{
    std::ifstream fin(filename);
    if (!fin) fatal(AMGX_RC_IO_ERROR, std::string("Error opening file '") + (filename ? filename : "(null)") + "'");
    std::string line;
    bool symmetric = false, skew = false, pattern = false, has_rhs = false, has_sol = false, base0 = false, sorted_hint = false;
    (void)sorted_hint;
    std::vector<std::string> header;
    while (fin.peek() == '%') {
        std::getline(fin, line);
        std::istringstream is(line);
        std::string tok;
        is >> tok;
        if (tok == "%%MatrixMarket") {
            while (is >> tok) {
                std::transform(tok.begin(), tok.end(), tok.begin(), ::tolower);
                if (tok == "symmetric") symmetric = true;
                if (tok == "skew-symmetric") { symmetric = true; skew = true; }
                if (tok == "pattern") pattern = true;
                if (tok == "complex") fatal(AMGX_RC_IO_ERROR, "Matrix is in complex format, but reading as real AMGX mode");
                if (tok == "array") fatal(AMGX_RC_IO_ERROR, "dense MatrixMarket arrays are not supported");
            }
        } else if (tok == "%%AMGX" || tok == "%%NVAMG") {
            while (is >> tok) {
                if (tok == "diagonal") S.has_diag = true;
                else if (tok == "rhs") has_rhs = true;
                else if (tok == "solution") has_sol = true;
                else if (tok == "base0") base0 = true;
                else if (tok == "sorted") sorted_hint = true;
                else if (isdigit((unsigned char)tok[0])) {
                    int a = atoi(tok.c_str());
                    int b = a;
                    std::string t2;
                    std::streampos pos = is.tellg();
                    if (is >> t2 && isdigit((unsigned char)t2[0])) b = atoi(t2.c_str());
                    else { is.clear(); is.seekg(pos); }
                    S.bx = a;
                    S.by = b;
                }
            }
        }
    }
    long long rows, cols, entries;
    fin >> rows >> cols >> entries;
    if (!fin || rows != cols) fatal(AMGX_RC_IO_ERROR, "MatrixMarket: bad size line or non-square matrix");
    const int bsq = S.bx * S.by;
    if (rows % S.bx) fatal(AMGX_RC_IO_ERROR, "MatrixMarket: matrix size is not a multiple of the block size");
    const int n = (int)(rows / S.bx);
    // scalar entries -> block entries keyed by (block row, block col)
    std::vector<std::map<int, std::vector<double>>> rowsmap(n);
    for (long long e = 0; e < entries; e++) {
        long long i, j;
        double v = 1.0;
        fin >> i >> j;
        if (!pattern) fin >> v;
        if (!fin) fatal(AMGX_RC_IO_ERROR, "MatrixMarket: unexpected end of file in entries");
        if (!base0) { i--; j--; }
        if (i < 0 || j < 0 || i >= rows || j >= cols) fatal(AMGX_RC_IO_ERROR, "Matrix Market format requires 1-based indexing. Use 'base0' AMGX format option to override.");
        auto put = [&](long long r, long long c, double val) {
            auto &blk = rowsmap[(int)(r / S.bx)][(int)(c / S.by)];
            if (blk.empty()) blk.assign(bsq, 0.0);
            blk[(size_t)(r % S.bx) * S.by + (c % S.by)] += val;
        };
        put(i, j, v);
        if (symmetric && i != j) put(j, i, skew ? -v : v);
    }
    S.n = n;
    S.rp.assign(n + 1, 0);
    if (S.has_diag) S.diag.assign((size_t)n * bsq, 0.0);
    for (int i = 0; i < n; i++) {
        for (auto &kv : rowsmap[i]) {
            if (S.has_diag && kv.first == i) { std::copy(kv.second.begin(), kv.second.end(), S.diag.begin() + (size_t)i * bsq); continue; }
            S.ci.push_back(kv.first);
            S.va.insert(S.va.end(), kv.second.begin(), kv.second.end());
        }
        S.rp[i + 1] = (int)S.ci.size();
    }
    S.nnz = (int)S.ci.size();
    auto read_vec = [&](std::vector<double> &v, size_t len) {
        v.resize(len);
        for (size_t k = 0; k < len; k++) {
            fin >> v[k];
            if (!fin) fatal(AMGX_RC_IO_ERROR, "MatrixMarket: unexpected end of file in rhs/solution");
        }
    };
    if (has_rhs) read_vec(S.rhs, (size_t)n * S.by);
    if (has_sol) read_vec(S.sol, (size_t)n * S.bx);
}

template <class T> static std::vector<T> convert(const std::vector<double> &v)
{
    return std::vector<T>(v.begin(), v.end());
}

static void upload_vec(Vector &v, const std::vector<double> &h, int n, int bd)
{
    v.n = n;
    v.block_dim = bd;
    v.data.resize(h.size(), v.prec);
    if (v.prec == Prec::F64) AMGXB_CUDA_CHECK(cudaMemcpy(v.data.ptr(), h.data(), h.size() * 8, cudaMemcpyHostToDevice));
    else {
        auto f = convert<float>(h);
        AMGXB_CUDA_CHECK(cudaMemcpy(v.data.ptr(), f.data(), f.size() * 4, cudaMemcpyHostToDevice));
    }
    v.user_order = true;
}

}  // namespace amgxb

extern "C" {

AMGX_RC AMGX_read_system(AMGX_matrix_handle mtx, AMGX_vector_handle rhs, AMGX_vector_handle sol, const char *filename)
{
    API2_BEGIN
    MMSystem S;
    read_mm(filename, S);
    if (mtx) {
        MatrixH *m = chk<MatrixH>(mtx, MAGIC_MTX, "matrix");
        AMGXB_CUDA_CHECK(cudaSetDevice(m->m->rsc->device));
        if (m->m->mat_prec == Prec::F64)
            upload_matrix(*m->m, S.n, S.nnz, S.bx, S.by, S.rp.data(), S.ci.data(), S.va.data(), S.has_diag ? S.diag.data() : nullptr);
        else {
            auto vf = convert<float>(S.va), df = convert<float>(S.diag);
            upload_matrix(*m->m, S.n, S.nnz, S.bx, S.by, S.rp.data(), S.ci.data(), vf.data(), S.has_diag ? df.data() : nullptr);
        }
    }
    if (rhs) {
        VectorH *b = chk<VectorH>(rhs, MAGIC_VEC, "vector");
        std::vector<double> h = S.rhs;
        if (h.empty()) h.assign((size_t)S.n * S.by, 1.0);   // rhs_from_a = 0: b = [1,...,1]^T
        upload_vec(*b->v, h, S.n, S.by);
    }
    if (sol) {
        VectorH *x = chk<VectorH>(sol, MAGIC_VEC, "vector");
        if (!S.sol.empty()) upload_vec(*x->v, S.sol, S.n, S.bx);
        else { x->v->n = 0; x->v->block_dim = S.bx; x->v->data.resize(0, x->v->prec); }
    }
    API2_END
}

AMGX_RC AMGX_write_system(const AMGX_matrix_handle mtx, const AMGX_vector_handle rhs, const AMGX_vector_handle sol, const char *filename)
{
    API2_BEGIN
    MatrixH *m = chk<MatrixH>(mtx, MAGIC_MTX, "matrix");
    Matrix &A = *m->m;
    AMGXB_CUDA_CHECK(cudaSetDevice(A.rsc->device));
    if (A.dist) fatal(AMGX_RC_NOT_IMPLEMENTED, "write_system of a distributed matrix");
    std::ofstream f(filename);
    if (!f) fatal(AMGX_RC_IO_ERROR, "cannot open output file");
    std::vector<int> rp = A.row_ptr.to_host(A.stream()), ci = A.col_idx.to_host(A.stream());
    const int bsq = A.bs();
    std::vector<double> va((size_t)A.nnz * bsq);
    if (A.mat_prec == Prec::F64) AMGXB_CUDA_CHECK(cudaMemcpy(va.data(), A.values.ptr(), va.size() * 8, cudaMemcpyDeviceToHost));
    else {
        std::vector<float> vf(va.size());
        AMGXB_CUDA_CHECK(cudaMemcpy(vf.data(), A.values.ptr(), vf.size() * 4, cudaMemcpyDeviceToHost));
        std::copy(vf.begin(), vf.end(), va.begin());
    }
    auto get_vec = [&](AMGX_vector_handle vh, std::vector<double> &out) {
        if (!vh) return;
        VectorH *v = chk<VectorH>(vh, MAGIC_VEC, "vector");
        const size_t len = (size_t)v->v->n * v->v->block_dim;
        out.resize(len);
        if (!len) return;
        if (v->v->prec == Prec::F64) AMGXB_CUDA_CHECK(cudaMemcpy(out.data(), v->v->data.ptr(), len * 8, cudaMemcpyDeviceToHost));
        else {
            std::vector<float> t(len);
            AMGXB_CUDA_CHECK(cudaMemcpy(t.data(), v->v->data.ptr(), len * 4, cudaMemcpyDeviceToHost));
            std::copy(t.begin(), t.end(), out.begin());
        }
    };
    std::vector<double> b, x;
    get_vec(rhs, b);
    get_vec(sol, x);
    f << "%%MatrixMarket matrix coordinate real general\n";
    f << "%%AMGX " << A.bx << " " << A.by << " sorted" << (b.empty() ? "" : " rhs") << (x.empty() ? "" : " solution") << "\n";
    f << (long long)A.n * A.bx << " " << (long long)A.n * A.by << " " << (long long)A.nnz * bsq << "\n";
    f.precision(17);
    for (int i = 0; i < A.n; i++)
        for (int k = rp[i]; k < rp[i + 1]; k++)
            for (int r = 0; r < A.bx; r++)
                for (int c = 0; c < A.by; c++)
                    f << (long long)i * A.bx + r + 1 << " " << (long long)ci[k] * A.by + c + 1 << " " << va[(size_t)k * bsq + r * A.by + c] << "\n";
    for (double v : b) f << v << "\n";
    for (double v : x) f << v << "\n";
    API2_END
}

AMGX_RC AMGXB200_get_nccl_unique_id(char *id128)
{
    API2_BEGIN
    if (!id128) fatal(AMGX_RC_BAD_PARAMETERS, "null pointer");
    dist_get_unique_id(id128);
    API2_END
}

// kind: 0 SpMV, 1 fused Jacobi sweep, 2 SpMV+dot
AMGX_RC AMGXB200_bench_kernel(AMGX_matrix_handle mtx, int kind, int warmup, int reps, int flush_l2, double *avg_ms)
{
    API2_BEGIN
    MatrixH *m = chk<MatrixH>(mtx, MAGIC_MTX, "matrix");
    Matrix &A = *m->m;
    AMGXB_CUDA_CHECK(cudaSetDevice(A.rsc->device));
    if (!A.initialized || A.dist) fatal(AMGX_RC_BAD_PARAMETERS, "bench_kernel needs an initialized single-GPU matrix");
    cudaStream_t s = A.stream();
    const size_t N = (size_t)A.n * A.by;
    DevVec x, y, b, d;
    x.resize(N, A.vec_prec);
    y.resize(N, A.vec_prec);
    b.resize(N, A.vec_prec);
    vec_fill(x.ptr(), A.vec_prec, N, 1.0, s);
    vec_fill(b.ptr(), A.vec_prec, N, 1.0, s);
    y.zero(s);
    if (A.bs() == 1) extract_diagonal(A, d, s);
    ScalarBlock sb;
    sb.create();
    ReduceScratch &rs = reduce_scratch(A.rsc.get());
    ReduceCtx red;
    red.partials = rs.partials.ptr();
    red.counter = rs.counter.ptr();
    red.scal = sb.scal;
    red.host_mirror = sb.host_dev;
    DevBuf<char> flush;
    const size_t flush_bytes = (size_t)256 << 20;
    if (flush_l2) flush.resize(flush_bytes);
    CsrOpArgs g;
    g.x = x.ptr();
    g.y = y.ptr();
    g.b = b.ptr();
    g.d = d.ptr();
    g.omega = 0.8;
    g.red = red;
    g.fin_op = FIN_STORE;
    g.fin_slot = S_TMP0;
    const CsrEpi epi = kind == 0 ? EPI_SPMV : kind == 1 ? EPI_JACOBI : EPI_SPMV_DOT;
    for (int i = 0; i < warmup; i++) matrix_apply(A, epi, g, s);
    cudaEvent_t e0, e1;
    AMGXB_CUDA_CHECK(cudaEventCreate(&e0));
    AMGXB_CUDA_CHECK(cudaEventCreate(&e1));
    double total = 0;
    if (flush_l2) {
        for (int i = 0; i < reps; i++) {
            AMGXB_CUDA_CHECK(cudaMemsetAsync(flush.ptr(), i & 0xff, flush_bytes, s));
            AMGXB_CUDA_CHECK(cudaEventRecord(e0, s));
            matrix_apply(A, epi, g, s);
            AMGXB_CUDA_CHECK(cudaEventRecord(e1, s));
            AMGXB_CUDA_CHECK(cudaEventSynchronize(e1));
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            total += ms;
        }
    } else {
        AMGXB_CUDA_CHECK(cudaEventRecord(e0, s));
        for (int i = 0; i < reps; i++) matrix_apply(A, epi, g, s);
        AMGXB_CUDA_CHECK(cudaEventRecord(e1, s));
        AMGXB_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        total = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    sb.destroy();
    if (avg_ms) *avg_ms = total / std::max(1, reps);
    API2_END
}

}  // extern "C"
