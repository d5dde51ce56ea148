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
#define API2_END_NORETURN } catch (...) { return on_exception(__func__); }

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

// The reference's binary system file (writer src/matrix_io.cu:267-405, reader src/readers.cu:1676-1960): "%%NVAMGBinary\n", nine
// uint32 {is_mtx, is_rhs, is_soln, format (0 = CSR), diag, block_dimx, block_dimy, rows, nnz}, int32 row offsets, int32 column
// indices, float64 values (off-diagonal blocks, then the diagonal blocks when diag), float64 rhs, float64 solution.
static const char NVAMG_BINARY_HEADER[] = "%%NVAMGBinary\n";
static bool read_binary(const char *filename, MMSystem &S)
{
    FILE *f = fopen(filename, "rb");
    if (!f) return false;
    char head[sizeof(NVAMG_BINARY_HEADER)] = {0};
    const size_t hl = strlen(NVAMG_BINARY_HEADER);
    if (fread(head, 1, hl, f) != hl || memcmp(head, NVAMG_BINARY_HEADER, hl) != 0) { fclose(f); return false; }
    auto need = [&](bool ok, const char *what) { if (!ok) { fclose(f); fatal(AMGX_RC_IO_ERROR, std::string("fread failed reading ") + what + ", exiting"); } };
    uint32_t fl[9];
    need(fread(fl, sizeof(uint32_t), 9, f) == 9, "the system header");
    if ((fl[3] & 0xffu) != 0) { fclose(f); fatal(AMGX_RC_IO_ERROR, "binary system file: only the CSR real format is supported"); }
    S.has_diag = fl[4] != 0;
    S.bx = (int)fl[5];
    S.by = (int)fl[6];
    S.n = (int)fl[7];
    S.nnz = (int)fl[8];
    const size_t bsq = (size_t)S.bx * S.by;
    S.rp.resize((size_t)S.n + 1);
    S.ci.resize((size_t)S.nnz);
    S.va.resize((size_t)S.nnz * bsq);
    need(fread(S.rp.data(), sizeof(int), S.rp.size(), f) == S.rp.size(), "row_offsets");
    need(S.nnz == 0 || fread(S.ci.data(), sizeof(int), S.ci.size(), f) == S.ci.size(), "column_indices");
    need(S.va.empty() || fread(S.va.data(), sizeof(double), S.va.size(), f) == S.va.size(), "off-diagonal values");
    if (S.has_diag) {
        S.diag.resize((size_t)S.n * bsq);
        need(fread(S.diag.data(), sizeof(double), S.diag.size(), f) == S.diag.size(), "diagonal values");
    }
    if (fl[1]) { S.rhs.resize((size_t)S.n * S.by); need(fread(S.rhs.data(), sizeof(double), S.rhs.size(), f) == S.rhs.size(), "rhs"); }
    if (fl[2]) { S.sol.resize((size_t)S.n * S.bx); need(fread(S.sol.data(), sizeof(double), S.sol.size(), f) == S.sol.size(), "solution"); }
    fclose(f);
    if (S.rp[0] != 0 || S.rp[S.n] != S.nnz) fatal(AMGX_RC_IO_ERROR, "binary system file: inconsistent row offsets");
    return true;
}

static void write_binary(const char *filename, int n, int nnz, int bx, int by, const std::vector<int> &rp, const std::vector<int> &ci, const std::vector<double> &va,
                         bool has_diag, const std::vector<double> &b, const std::vector<double> &x)
{
    FILE *f = fopen(filename, "wb");
    if (!f) fatal(AMGX_RC_BAD_PARAMETERS, "Cannot open output file!11");
    const uint32_t fl[9] = {1u, (uint32_t)!b.empty(), (uint32_t)!x.empty(), 0u, (uint32_t)has_diag, (uint32_t)bx, (uint32_t)by, (uint32_t)n, (uint32_t)nnz};
    bool ok = fwrite(NVAMG_BINARY_HEADER, 1, strlen(NVAMG_BINARY_HEADER), f) == strlen(NVAMG_BINARY_HEADER) && fwrite(fl, sizeof(uint32_t), 9, f) == 9;
    ok = ok && fwrite(rp.data(), sizeof(int), (size_t)n + 1, f) == (size_t)n + 1;
    ok = ok && (nnz == 0 || fwrite(ci.data(), sizeof(int), (size_t)nnz, f) == (size_t)nnz);
    ok = ok && (va.empty() || fwrite(va.data(), sizeof(double), va.size(), f) == va.size());
    ok = ok && (b.empty() || fwrite(b.data(), sizeof(double), b.size(), f) == b.size());
    ok = ok && (x.empty() || fwrite(x.data(), sizeof(double), x.size(), f) == x.size());
    fclose(f);
    if (!ok) fatal(AMGX_RC_IO_ERROR, "error while writing the binary system file");
}

static void read_mm(const char *filename, MMSystem &S)
{
    if (!filename) fatal(AMGX_RC_IO_ERROR, "Error opening file '(null)'");
    if (read_binary(filename, S)) return;          // "%%NVAMGBinary" files are detected by their header
    std::ifstream fin(filename);
    if (!fin) fatal(AMGX_RC_IO_ERROR, std::string("Error opening file '") + filename + "'");
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
    // What follows the entries in the reference's format (writer src/matrix_io.cu:222-258, reader src/readers.cu:1290-1406): with
    // "diagonal" one line of block_size values per row (the diagonal blocks, which are then NOT among the entries); with "rhs" /
    // "solution" a line holding the vector length, then the values.  Files without the length lines, and files that carry the diagonal
    // inline as (i, i) entries, are accepted too: the layout is told apart by the number of values left in the file.
    std::vector<double> tail;
    for (double v; fin >> v;) tail.push_back(v);
    const size_t need_b = has_rhs ? (size_t)n * S.by : 0, need_x = has_sol ? (size_t)n * S.bx : 0, L = (size_t)has_rhs + (size_t)has_sol;
    const size_t D = S.has_diag ? (size_t)n * bsq : 0;
    // Candidate layouts in the reference's order of preference (its own writer first); a candidate is accepted when the value count fits
    // AND, where it claims length lines, those tokens really hold the vector lengths -- the count alone is ambiguous for 1- and 2-row
    // systems (n * block size can equal the number of length lines).
    bool diag_section = false, lengths = false, found = false;
    auto fits = [&](bool ds, bool ln) -> bool {
        const size_t d = ds ? D : 0, l = ln ? L : 0;
        if (tail.size() != d + need_b + need_x + l) return false;
        if (ln) {
            size_t pos = d;
            if (has_rhs) { if ((size_t)tail[pos] != need_b || tail[pos] != (double)need_b) return false; pos += 1 + need_b; }
            if (has_sol) { if ((size_t)tail[pos] != need_x || tail[pos] != (double)need_x) return false; }
        }
        return true;
    };
    const bool cand[4][2] = {{true, true}, {true, false}, {false, true}, {false, false}};
    for (int c = 0; c < 4 && !found; c++) {
        const bool ds = cand[c][0], ln = cand[c][1] && L > 0;
        if (ds && !S.has_diag) continue;
        if (cand[c][1] && L == 0) continue;
        if (fits(ds, ln)) { diag_section = ds; lengths = ln; found = true; }
    }
    if (!found) fatal(AMGX_RC_IO_ERROR, "MatrixMarket: unexpected number of values after the matrix entries (diagonal / rhs / solution sections)");
    S.n = n;
    S.rp.assign(n + 1, 0);
    if (S.has_diag) S.diag.assign((size_t)n * bsq, 0.0);
    for (int i = 0; i < n; i++) {
        for (auto &kv : rowsmap[i]) {
            if (S.has_diag && !diag_section && kv.first == i) { std::copy(kv.second.begin(), kv.second.end(), S.diag.begin() + (size_t)i * bsq); continue; }
            S.ci.push_back(kv.first);
            S.va.insert(S.va.end(), kv.second.begin(), kv.second.end());
        }
        S.rp[i + 1] = (int)S.ci.size();
    }
    S.nnz = (int)S.ci.size();
    size_t pos = 0;
    if (diag_section) { std::copy(tail.begin(), tail.begin() + (long)D, S.diag.begin()); pos = D; }
    auto take = [&](std::vector<double> &v, size_t len) {
        if (lengths) {
            if ((size_t)tail[pos] != len) fatal(AMGX_RC_IO_ERROR, "MatrixMarket: rhs / solution length line does not match the matrix size");
            pos++;
        }
        v.assign(tail.begin() + (long)pos, tail.begin() + (long)(pos + len));
        pos += len;
    };
    if (has_rhs) take(S.rhs, need_b);
    if (has_sol) take(S.sol, need_x);
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

// ---------------------------------------------------------------------------------------------
// Partitioned reads: AMGX_read_system_distributed / _global / _maps_one_ring (src/amgx_c.cu:1497-1700, 4069-4430).
// Every rank reads the whole file and keeps its rows.  Who owns a row: the partition vector if given (num_partitions > ranks:
// consecutive groups of partitions share a rank, amgx_c.cu:1578-1590), else partition_sizes (contiguous blocks), else equal contiguous
// blocks (p * n / ranks).  A rank's rows are kept in increasing global id; in the engine's contiguous numbering (the reference's
// ipartition_map) rank r owns [offsets[r], offsets[r+1]).
// ---------------------------------------------------------------------------------------------
struct LocalPart {
    int n = 0, nnz = 0;
    std::vector<int64_t> offsets;            // [world+1], contiguous numbering
    std::vector<int> rows;                   // original global ids of my rows, ascending
    std::vector<int> rp;
    std::vector<int64_t> cols_contig, cols_orig;
    std::vector<double> va, diag, rhs, sol;
};

static void partition_system(const MMSystem &S, int rank, int world, int num_partitions, const int *partition_sizes, int pv_size, const int *pv_in, LocalPart &L)
{
    const int ng = S.n, bsq = S.bx * S.by;
    std::vector<int> pv(ng);
    if (pv_in) {
        if (pv_size != ng) fatal(AMGX_RC_BAD_PARAMETERS, "partition_vector_size does not match the global vector size");
        int maxp = 0;
        for (int i = 0; i < ng; i++) { if (pv_in[i] < 0) fatal(AMGX_RC_BAD_PARAMETERS, "negative partition id"); maxp = std::max(maxp, pv_in[i]); }
        int nparts = num_partitions > 0 ? num_partitions : maxp + 1;
        if (nparts < maxp + 1) fatal(AMGX_RC_BAD_PARAMETERS, "partition vector names more partitions than num_partitions");
        if (nparts % world) fatal(AMGX_RC_BAD_PARAMETERS, "the number of partitions must be a multiple of the number of ranks");
        const int per_rank = nparts / world;
        for (int i = 0; i < ng; i++) pv[i] = pv_in[i] / per_rank;
    } else if (partition_sizes) {
        const int nparts = num_partitions > 0 ? num_partitions : world;
        if (nparts % world) fatal(AMGX_RC_BAD_PARAMETERS, "the number of partitions must be a multiple of the number of ranks");
        const int per_rank = nparts / world;
        long long g = 0;
        for (int p = 0; p < nparts; p++)
            for (int k = 0; k < partition_sizes[p]; k++, g++) {
                if (g >= ng) fatal(AMGX_RC_BAD_PARAMETERS, "partition_sizes add up to more rows than the matrix has");
                pv[g] = p / per_rank;
            }
        if (g != ng) fatal(AMGX_RC_BAD_PARAMETERS, "partition_sizes do not add up to the number of rows");
    } else {
        int p = 0;
        for (int i = 0; i < ng; i++) {
            while (p + 1 < world && i >= (long long)(p + 1) * ng / world) p++;
            pv[i] = p;
        }
    }
    L.offsets.assign((size_t)world + 1, 0);
    std::vector<int64_t> new_global((size_t)std::max(ng, 1));
    if (!partition_vector_to_contiguous(ng, world, pv.data(), L.offsets.data(), new_global.data()))
        fatal(AMGX_RC_BAD_PARAMETERS, "partition vector names a rank outside [0, number of ranks)");
    L.rows.clear();
    for (int g = 0; g < ng; g++) if (pv[g] == rank) L.rows.push_back(g);
    L.n = (int)L.rows.size();
    L.rp.assign((size_t)L.n + 1, 0);
    for (int i = 0; i < L.n; i++) L.rp[i + 1] = L.rp[i] + (S.rp[L.rows[i] + 1] - S.rp[L.rows[i]]);
    L.nnz = L.rp[L.n];
    L.cols_contig.resize((size_t)std::max(L.nnz, 1));
    L.cols_orig.resize((size_t)std::max(L.nnz, 1));
    L.va.resize((size_t)L.nnz * bsq);
    if (S.has_diag) L.diag.resize((size_t)L.n * bsq);
    if (!S.rhs.empty()) L.rhs.resize((size_t)L.n * S.by);
    if (!S.sol.empty()) L.sol.resize((size_t)L.n * S.bx);
    for (int i = 0; i < L.n; i++) {
        const int g = L.rows[i];
        for (int k = S.rp[g], o = L.rp[i]; k < S.rp[g + 1]; k++, o++) {
            L.cols_orig[o] = S.ci[k];
            L.cols_contig[o] = new_global[S.ci[k]];
            std::copy(S.va.begin() + (size_t)k * bsq, S.va.begin() + (size_t)(k + 1) * bsq, L.va.begin() + (size_t)o * bsq);
        }
        if (S.has_diag) std::copy(S.diag.begin() + (size_t)g * bsq, S.diag.begin() + (size_t)(g + 1) * bsq, L.diag.begin() + (size_t)i * bsq);
        if (!S.rhs.empty()) std::copy(S.rhs.begin() + (size_t)g * S.by, S.rhs.begin() + (size_t)(g + 1) * S.by, L.rhs.begin() + (size_t)i * S.by);
        if (!S.sol.empty()) std::copy(S.sol.begin() + (size_t)g * S.bx, S.sol.begin() + (size_t)(g + 1) * S.bx, L.sol.begin() + (size_t)i * S.bx);
    }
}

template <class T> static T *c_dup(const std::vector<double> &v)
{
    T *p = (T *)malloc(sizeof(T) * std::max<size_t>(v.size(), 1));
    if (!p) fatal(AMGX_RC_NO_MEMORY, "out of host memory");
    for (size_t i = 0; i < v.size(); i++) p[i] = (T)v[i];
    return p;
}
static void *c_dup_prec(const std::vector<double> &v, bool f64) { return f64 ? (void *)c_dup<double>(v) : (void *)c_dup<float>(v); }
template <class T, class U> static T *c_dup_i(const std::vector<U> &v)
{
    T *p = (T *)malloc(sizeof(T) * std::max<size_t>(v.size(), 1));
    if (!p) fatal(AMGX_RC_NO_MEMORY, "out of host memory");
    for (size_t i = 0; i < v.size(); i++) p[i] = (T)v[i];
    return p;
}

// What AMGX_read_system_maps_one_ring hands back (examples/amgx_mpi_capi_agg.c:386-420): rows in their original order, owned columns
// numbered 0..n-1, halo columns from n on grouped by neighbour (ascending rank) and by ascending global id inside a group; for every
// neighbour the local rows it needs (send, ascending) and the halo columns its values land in (recv).  Exactly the conventions of
// the partition planner, whose row renumbering [interior | boundary] is undone here.
static void maps_one_ring(const MMSystem &S, const LocalPart &L, int rank, int world, std::vector<int> &local_cols, std::vector<int> &neighbors,
                          std::vector<std::vector<int>> &send, std::vector<std::vector<int>> &recv)
{
    AMGXB200_partition_plan pl;
    memset(&pl, 0, sizeof(pl));
    partition_plan_create(&pl, rank, world, L.offsets.data(), L.n, L.nnz, L.rp.data(), L.cols_contig.data());
    std::vector<int> inv((size_t)std::max(L.n, 1));
    for (int i = 0; i < L.n; i++) inv[pl.perm_old_to_new[i]] = i;
    local_cols.resize((size_t)std::max(L.nnz, 1));
    for (int k = 0; k < L.nnz; k++) local_cols[k] = pl.local_cols[k] < L.n ? inv[pl.local_cols[k]] : pl.local_cols[k];
    neighbors.assign(pl.neighbors, pl.neighbors + pl.num_neighbors);
    send.assign(pl.num_neighbors, {});
    recv.assign(pl.num_neighbors, {});
    for (int q = 0; q < pl.num_neighbors; q++) {
        for (int k = pl.send_offsets[q]; k < pl.send_offsets[q + 1]; k++) send[q].push_back(inv[pl.send_maps[k]]);
        for (int k = pl.halo_offsets[q]; k < pl.halo_offsets[q + 1]; k++) recv[q].push_back(L.n + k);
    }
    AMGXB200_partition_plan_free(&pl);
    (void)S;
}

static AMGX_RC read_maps_impl(int rank, int world, int mode, const char *filename, int num_partitions, const int *partition_sizes, int pv_size, const int *pv,
                              int *n, int *nnz, int *bx, int *by, int **row_ptrs, int **col_local, int64_t **col_global, void **data, void **diag_data,
                              void **rhs, void **sol, int *num_neighbors, int **neighbors, int **send_sizes, int ***send_maps, int **recv_sizes,
                              int ***recv_maps)
{
    MMSystem S;
    read_mm(filename, S);
    LocalPart L;
    partition_system(S, rank, world, num_partitions, partition_sizes, pv_size, pv, L);
    // mode = mem + 16 * vec + 256 * mat + 4096 * ind, 0 = double (include/amgx_config.h:81-124)
    const bool mat64 = ((mode >> 8) & 15) == 0, vec64 = ((mode >> 4) & 15) == 0;
    *n = L.n;
    *nnz = L.nnz;
    *bx = S.bx;
    *by = S.by;
    *row_ptrs = c_dup_i<int>(L.rp);
    *data = c_dup_prec(L.va, mat64);
    *diag_data = S.has_diag ? c_dup_prec(L.diag, mat64) : nullptr;
    std::vector<double> b = L.rhs, x = L.sol;
    if (b.empty()) b.assign((size_t)L.n * S.by, 1.0);             // no rhs in the file: b = [1,...,1]^T (rhs_from_a = 0)
    if (x.empty()) x.assign((size_t)L.n * S.bx, 0.0);             // "Initializing solution vector with zeroes..."
    *rhs = c_dup_prec(b, vec64);
    *sol = c_dup_prec(x, vec64);
    if (col_global) *col_global = c_dup_i<int64_t>(L.cols_orig);  // ORIGINAL numbering, to go with the partition vector (amgx_c.cu:4412-4426)
    if (col_local) {
        std::vector<int> lc, nb;
        std::vector<std::vector<int>> sm, rm;
        maps_one_ring(S, L, rank, world, lc, nb, sm, rm);
        lc.resize((size_t)L.nnz);
        *col_local = c_dup_i<int>(lc);
        const int nn = (int)nb.size();
        *num_neighbors = nn;
        *neighbors = c_dup_i<int>(nb);
        *send_sizes = (int *)malloc(sizeof(int) * std::max(nn, 1));
        *recv_sizes = (int *)malloc(sizeof(int) * std::max(nn, 1));
        *send_maps = (int **)malloc(sizeof(int *) * std::max(nn, 1));
        *recv_maps = (int **)malloc(sizeof(int *) * std::max(nn, 1));
        for (int q = 0; q < nn; q++) {
            (*send_sizes)[q] = (int)sm[q].size();
            (*recv_sizes)[q] = (int)rm[q].size();
            (*send_maps)[q] = c_dup_i<int>(sm[q]);
            (*recv_maps)[q] = c_dup_i<int>(rm[q]);
        }
    }
    return AMGX_RC_OK;
}

}  // namespace amgxb

extern "C" {

AMGX_RC AMGX_read_system_maps_one_ring(int *n, int *nnz, int *block_dimx, int *block_dimy, int **row_ptrs, int **col_indices, void **data, void **diag_data,
                                       void **rhs, void **sol, int *num_neighbors, int **neighbors, int **send_sizes, int ***send_maps, int **recv_sizes,
                                       int ***recv_maps, AMGX_resources_handle rsc, AMGX_Mode mode, const char *filename, int allocated_halo_depth,
                                       int num_partitions, const int *partition_sizes, int partition_vector_size, const int *partition_vector)
{
    API2_BEGIN
    (void)allocated_halo_depth;
    ResourcesH *r = chk<ResourcesH>(rsc, MAGIC_RSC, "resources");
    if (!n || !nnz || !block_dimx || !block_dimy || !row_ptrs || !col_indices || !data || !diag_data || !rhs || !sol || !num_neighbors || !neighbors ||
        !send_sizes || !send_maps || !recv_sizes || !recv_maps)
        fatal(AMGX_RC_BAD_PARAMETERS, "AMGX_read_system_maps_one_ring: null output pointer");
    read_maps_impl(r->rsc->rank, r->rsc->world, (int)mode, filename, num_partitions, partition_sizes, partition_vector_size, partition_vector, n, nnz,
                   block_dimx, block_dimy, row_ptrs, col_indices, nullptr, data, diag_data, rhs, sol, num_neighbors, neighbors, send_sizes, send_maps,
                   recv_sizes, recv_maps);
    API2_END
}

AMGX_RC AMGX_read_system_global(int *n, int *nnz, int *block_dimx, int *block_dimy, int **row_ptrs, void **col_indices_global, void **data, void **diag_data,
                                void **rhs, void **sol, AMGX_resources_handle rsc, AMGX_Mode mode, const char *filename, int allocated_halo_depth,
                                int num_partitions, const int *partition_sizes, int partition_vector_size, const int *partition_vector)
{
    API2_BEGIN
    (void)allocated_halo_depth;
    ResourcesH *r = chk<ResourcesH>(rsc, MAGIC_RSC, "resources");
    if (!n || !nnz || !block_dimx || !block_dimy || !row_ptrs || !col_indices_global || !data || !diag_data || !rhs || !sol)
        fatal(AMGX_RC_BAD_PARAMETERS, "AMGX_read_system_global: null output pointer");
    int64_t *cg = nullptr;
    read_maps_impl(r->rsc->rank, r->rsc->world, (int)mode, filename, num_partitions, partition_sizes, partition_vector_size, partition_vector, n, nnz,
                   block_dimx, block_dimy, row_ptrs, nullptr, &cg, data, diag_data, rhs, sol, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    *col_indices_global = cg;
    API2_END
}

AMGX_RC AMGX_free_system_maps_one_ring(int *row_ptrs, int *col_indices, void *data, void *diag_data, void *rhs, void *sol, int num_neighbors, int *neighbors,
                                       int *send_sizes, int **send_maps, int *recv_sizes, int **recv_maps)
{
    free(row_ptrs);
    free(col_indices);
    free(data);
    free(diag_data);
    free(rhs);
    free(sol);
    free(neighbors);
    free(send_sizes);
    free(recv_sizes);
    if (send_maps) { for (int q = 0; q < num_neighbors; q++) free(send_maps[q]); free(send_maps); }
    if (recv_maps) { for (int q = 0; q < num_neighbors; q++) free(recv_maps[q]); free(recv_maps); }
    return AMGX_RC_OK;
}

/* Test / tooling hook: AMGX_read_system_maps_one_ring + _global for an explicit (rank, world) pair, no resources handle, no GPU. */
AMGX_RC AMGXB200_read_system_partition(int rank, int world_size, AMGX_Mode mode, const char *filename, int num_partitions, const int *partition_sizes,
                                       int partition_vector_size, const int *partition_vector, int *n, int *nnz, int *block_dimx, int *block_dimy,
                                       int **row_ptrs, int **col_indices_local, int64_t **col_indices_global, void **data, void **diag_data, void **rhs,
                                       void **sol, int *num_neighbors, int **neighbors, int **send_sizes, int ***send_maps, int **recv_sizes, int ***recv_maps)
{
    API2_BEGIN
    if (rank < 0 || world_size < 1 || rank >= world_size) fatal(AMGX_RC_BAD_PARAMETERS, "bad rank / world size");
    read_maps_impl(rank, world_size, (int)mode, filename, num_partitions, partition_sizes, partition_vector_size, partition_vector, n, nnz, block_dimx, block_dimy,
                   row_ptrs, col_indices_local, col_indices_global, data, diag_data, rhs, sol, num_neighbors, neighbors, send_sizes, send_maps, recv_sizes,
                   recv_maps);
    API2_END
}

AMGX_RC AMGX_read_system_distributed(AMGX_matrix_handle mtx, AMGX_vector_handle rhs, AMGX_vector_handle sol, const char *filename, int allocated_halo_depth,
                                     int num_partitions, const int *partition_sizes, int partition_vector_size, const int *partition_vector)
{
    API2_BEGIN
    (void)allocated_halo_depth;
    MatrixH *m = mtx ? chk<MatrixH>(mtx, MAGIC_MTX, "matrix") : nullptr;
    VectorH *b = rhs ? chk<VectorH>(rhs, MAGIC_VEC, "vector") : nullptr;
    VectorH *x = sol ? chk<VectorH>(sol, MAGIC_VEC, "vector") : nullptr;
    Resources *rs = m ? m->m->rsc.get() : b ? b->v->rsc.get() : x ? x->v->rsc.get() : nullptr;
    if (!rs) fatal(AMGX_RC_BAD_PARAMETERS, "AMGX_read_system_distributed: no matrix or vector handle");
    AMGXB_CUDA_CHECK(cudaSetDevice(rs->device));
    MMSystem S;
    read_mm(filename, S);
    LocalPart L;
    partition_system(S, rs->rank, rs->world, num_partitions, partition_sizes, partition_vector_size, partition_vector, L);
    if (m) {
        Matrix &A = *m->m;
        const void *vals, *dg = nullptr;
        std::vector<float> vf, df;
        if (A.mat_prec == Prec::F64) { vals = L.va.data(); if (S.has_diag) dg = L.diag.data(); }
        else { vf = convert<float>(L.va); df = convert<float>(L.diag); vals = vf.data(); if (S.has_diag) dg = df.data(); }
        if (rs->world == 1) {
            std::vector<int> c32(L.cols_contig.begin(), L.cols_contig.end());
            upload_matrix(A, L.n, L.nnz, S.bx, S.by, L.rp.data(), c32.data(), vals, dg);
        } else {
            dist_build_matrix(A, L.offsets.data(), L.n, L.nnz, S.bx, S.by, L.rp.data(), L.cols_contig.data(), vals, dg);
        }
    }
    auto put = [&](VectorH *vh, const std::vector<double> &h, int bd, bool allow_empty) {
        if (!vh) return;
        Vector &v = *vh->v;
        if (h.empty() && allow_empty) { v.n = 0; v.block_dim = bd; v.data.resize(0, v.prec); return; }
        upload_vec(v, h, L.n, bd);
        if (m && m->m->dist) v.dist = m->m->dist;       // bound: the solve permutes from the caller's row order
    };
    std::vector<double> hb = L.rhs;
    if (hb.empty()) hb.assign((size_t)L.n * S.by, 1.0);
    put(b, hb, S.by, false);
    std::vector<double> hx = L.sol;
    if (hx.empty()) hx.assign((size_t)L.n * S.bx, 0.0);         // no solution section: zeros, as in AMGX_read_system
    put(x, hx, S.bx, false);
    API2_END
}

// matrix_writer = matrixmarket (default) | binary, read from the configuration the resources were created with (matrix_io.cu:505-530).
// va holds nnz blocks, followed by n diagonal blocks when ext_diag (written by the binary format only).
static void write_host_system(const char *filename, const std::string &writer, int n, int nnz, int bx, int by, const std::vector<int> &rp, const std::vector<int> &ci,
                              const std::vector<double> &va, bool ext_diag, const std::vector<double> &b, const std::vector<double> &x)
{
    const int bsq = bx * by;
    if (writer == "binary") {
        write_binary(filename, n, nnz, bx, by, rp, ci, va, ext_diag, b, x);
        return;
    }
    if (writer != "matrixmarket") fatal(AMGX_RC_BAD_CONFIGURATION, "matrix_writer '" + writer + "' is not supported (matrixmarket, binary)");
    std::ofstream f(filename);
    if (!f) fatal(AMGX_RC_IO_ERROR, "cannot open output file");
    // the reference's writer, src/matrix_io.cu:120-258: "%%NVAMG bx by [diagonal] [rhs] [solution]" (no "sorted": the columns need not be),
    // entries, the diagonal blocks one per line when they are stored outside the CSR structure, then length + values of rhs and solution;
    // 17 significant digits instead of the reference's 16 so that a round trip is exact
    f << "%%MatrixMarket matrix coordinate real general\n";
    f << "%%NVAMG " << bx << " " << by << (ext_diag ? " diagonal" : "") << (b.empty() ? "" : " rhs") << (x.empty() ? "" : " solution") << "\n";
    f << (long long)n * bx << " " << (long long)n * by << " " << (long long)nnz * bsq << "\n";
    f.precision(17);
    f << std::scientific;
    for (int i = 0; i < n; i++)
        for (int k = rp[i]; k < rp[i + 1]; k++)
            for (int r = 0; r < bx; r++)
                for (int c = 0; c < by; c++)
                    f << (long long)i * bx + r + 1 << " " << (long long)ci[k] * by + c + 1 << " " << va[(size_t)k * bsq + r * by + c] << "\n";
    if (ext_diag)
        for (int i = 0; i < n; i++) {
            for (int k = 0; k < bsq; k++) f << va[((size_t)nnz + i) * bsq + k] << " ";
            f << "\n";
        }
    if (!b.empty()) { f << b.size() << "\n"; for (double v : b) f << v << "\n"; }
    if (!x.empty()) { f << x.size() << "\n"; for (double v : x) f << v << "\n"; }
}

AMGX_RC AMGXB200_write_system_host(const char *filename, const char *writer, int n, int nnz, int block_dimx, int block_dimy, const int *row_ptrs,
                                   const int *col_indices, const double *values, int ext_diag, const double *rhs, const double *sol)
{
    API2_BEGIN
    if (!filename || !writer || n < 0 || nnz < 0 || block_dimx < 1 || block_dimy < 1 || !row_ptrs || (nnz > 0 && (!col_indices || !values)))
        fatal(AMGX_RC_BAD_PARAMETERS, "AMGXB200_write_system_host: bad arguments");
    const size_t bsq = (size_t)block_dimx * block_dimy;
    std::vector<int> rp(row_ptrs, row_ptrs + n + 1), ci(col_indices, col_indices + nnz);
    std::vector<double> va(values, values + ((size_t)nnz + (ext_diag ? (size_t)n : 0)) * bsq), b, x;
    if (rhs) b.assign(rhs, rhs + (size_t)n * block_dimy);
    if (sol) x.assign(sol, sol + (size_t)n * block_dimx);
    write_host_system(filename, writer, n, nnz, block_dimx, block_dimy, rp, ci, va, ext_diag != 0, b, x);
    API2_END
}

/* AMGX_write_system_distributed (src/amgx_c.cu:1406-1491, 3557-3600): the partitions are gathered and rank 0 writes ONE global system.
 * Every rank assembles the global matrix in the callers' row order (dist_gather_matrix) and the vectors with one all-gather; when the
 * matrix was uploaded through a partition vector, passing the same vector here restores the original global numbering (the reference's
 * construct_global_matrix); without it the rows appear in the contiguous per-rank numbering.  Collective: every rank must call it. */
AMGX_RC AMGX_write_system_distributed(const AMGX_matrix_handle mtx, const AMGX_vector_handle rhs, const AMGX_vector_handle sol, const char *filename,
                                      int allocated_halo_depth, int num_partitions, const int *partition_sizes, int partition_vector_size,
                                      const int *partition_vector)
{
    (void)allocated_halo_depth; (void)num_partitions; (void)partition_sizes;
    bool is_dist = false;
    {
        API2_BEGIN
        MatrixH *m = chk<MatrixH>(mtx, MAGIC_MTX, "matrix");
        is_dist = (bool)m->m->dist;
        API2_END_NORETURN
    }
    if (!is_dist) return AMGX_write_system(mtx, rhs, sol, filename);
    API2_BEGIN
    if (!filename) fatal(AMGX_RC_BAD_PARAMETERS, "null file name");
    MatrixH *m = chk<MatrixH>(mtx, MAGIC_MTX, "matrix");
    Matrix &A = *m->m;
    AMGXB_CUDA_CHECK(cudaSetDevice(A.rsc->device));
    if (A.has_ext_diag) fatal(AMGX_RC_NOT_IMPLEMENTED, "AMGX_write_system_distributed with an external diagonal");
    cudaStream_t s = A.stream();
    const int rank = A.rsc->rank, world = A.rsc->world;
    std::vector<int> counts, offs;
    std::unique_ptr<Matrix> G = dist_gather_matrix(A, counts, offs, true);
    const int N = G->n, NNZ = G->nnz, bsq = A.bs();
    auto gather_vec = [&](AMGX_vector_handle vh, std::vector<double> &out) {
        if (!vh) return;
        VectorH *v = chk<VectorH>(vh, MAGIC_VEC, "vector");
        const int bd = v->v->block_dim;
        if (v->v->n != A.n) fatal(AMGX_RC_BAD_PARAMETERS, "AMGX_write_system_distributed: vector and matrix sizes differ");
        const size_t len = (size_t)A.n * bd;
        std::vector<double> mine(len);
        if (len) {
            if (v->v->prec == Prec::F64) {
                if (v->v->dist && !v->v->user_order) dist_download_vector(*v->v, mine.data());
                else AMGXB_CUDA_CHECK(cudaMemcpy(mine.data(), v->v->data.ptr(), len * 8, cudaMemcpyDeviceToHost));
            } else {
                std::vector<float> t(len);
                if (v->v->dist && !v->v->user_order) dist_download_vector(*v->v, t.data());
                else AMGXB_CUDA_CHECK(cudaMemcpy(t.data(), v->v->data.ptr(), len * 4, cudaMemcpyDeviceToHost));
                std::copy(t.begin(), t.end(), mine.begin());
            }
        }
        DevVec g;
        g.resize((size_t)N * bd, Prec::F64);
        g.zero(s);
        if (len) AMGXB_CUDA_CHECK(cudaMemcpyAsync((double *)g.ptr() + (size_t)offs[rank] * bd, mine.data(), len * 8, cudaMemcpyHostToDevice, s));
        dist_allgatherv_inplace(A, g.ptr(), Prec::F64, bd, counts, offs, s);
        out.resize((size_t)N * bd);
        if (!out.empty()) AMGXB_CUDA_CHECK(cudaMemcpyAsync(out.data(), g.ptr(), out.size() * 8, cudaMemcpyDeviceToHost, s));
        AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    };
    std::vector<double> b, x;
    gather_vec(rhs, b);
    gather_vec(sol, x);
    if (rank != 0) return AMGX_RC_OK;
    std::vector<int> rp = G->row_ptr.to_host(s), ci = G->col_idx.to_host(s);
    std::vector<double> va((size_t)NNZ * bsq);
    if (NNZ) {
        if (A.mat_prec == Prec::F64) AMGXB_CUDA_CHECK(cudaMemcpy(va.data(), G->values.ptr(), va.size() * 8, cudaMemcpyDeviceToHost));
        else {
            std::vector<float> vf(va.size());
            AMGXB_CUDA_CHECK(cudaMemcpy(vf.data(), G->values.ptr(), vf.size() * 4, cudaMemcpyDeviceToHost));
            std::copy(vf.begin(), vf.end(), va.begin());
        }
    }
    if (partition_vector) {
        // contiguous id (what the ranks hold) -> original global id
        if (partition_vector_size != N) fatal(AMGX_RC_BAD_PARAMETERS, "AMGX_write_system_distributed: partition vector size != global number of rows");
        std::vector<int64_t> off((size_t)world + 1), newg((size_t)std::max(N, 1));
        if (!partition_vector_to_contiguous(N, world, partition_vector, off.data(), newg.data()))
            fatal(AMGX_RC_BAD_PARAMETERS, "partition vector names a rank outside [0, number of ranks)");
        std::vector<int> orig((size_t)std::max(N, 1));
        for (int g = 0; g < N; g++) orig[(size_t)newg[g]] = g;
        std::vector<int> rp2((size_t)N + 1, 0), ci2(ci.size());
        std::vector<double> va2(va.size());
        for (int i = 0; i < N; i++) rp2[(size_t)orig[i] + 1] = rp[i + 1] - rp[i];
        for (int i = 0; i < N; i++) rp2[i + 1] += rp2[i];
        for (int i = 0; i < N; i++) {
            const int d = rp2[orig[i]];
            for (int k = rp[i]; k < rp[i + 1]; k++) {
                ci2[(size_t)d + (k - rp[i])] = orig[ci[k]];
                std::copy(va.begin() + (size_t)k * bsq, va.begin() + (size_t)(k + 1) * bsq, va2.begin() + (size_t)(d + (k - rp[i])) * bsq);
            }
        }
        rp.swap(rp2); ci.swap(ci2); va.swap(va2);
        auto unpermute = [&](std::vector<double> &v) {
            if (v.empty()) return;
            const size_t bd = v.size() / (size_t)N;
            std::vector<double> t(v.size());
            for (int i = 0; i < N; i++) std::copy(v.begin() + (size_t)i * bd, v.begin() + (size_t)(i + 1) * bd, t.begin() + (size_t)orig[i] * bd);
            v.swap(t);
        };
        unpermute(b);
        unpermute(x);
    }
    write_host_system(filename, A.rsc->cfg ? A.rsc->cfg->get_string("matrix_writer", "default") : std::string("matrixmarket"), N, NNZ, A.bx, A.by, rp, ci, va, false, b, x);
    API2_END
}

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
        if (h.empty()) {
            // no rhs in the file (src/readers.cu:1364-1388): b = [1,...,1]^T, or b = A e with e = [1,...,1]^T when rhs_from_a = 1
            const bool from_a = b->v->rsc && b->v->rsc->cfg && b->v->rsc->cfg->get_int("rhs_from_a", "default") == 1;
            if (!from_a) h.assign((size_t)S.n * S.by, 1.0);
            else {
                h.assign((size_t)S.n * S.by, 0.0);
                const int bsq = S.bx * S.by;
                for (int i = 0; i < S.n; i++) {
                    for (int k = S.rp[i]; k < S.rp[i + 1]; k++)
                        for (int r = 0; r < S.bx; r++)
                            for (int c = 0; c < S.by; c++) h[(size_t)i * S.by + r] += S.va[(size_t)k * bsq + r * S.by + c];
                    if (S.has_diag)
                        for (int r = 0; r < S.bx; r++)
                            for (int c = 0; c < S.by; c++) h[(size_t)i * S.by + r] += S.diag[(size_t)i * bsq + r * S.by + c];
                }
            }
        }
        upload_vec(*b->v, h, S.n, S.by);
    }
    if (sol) {
        VectorH *x = chk<VectorH>(sol, MAGIC_VEC, "vector");
        // no solution section: x = [0, ..., 0]^T of the matrix size, as the reference's reader leaves it (src/readers.cu:1392-1412)
        std::vector<double> h = S.sol;
        if (h.empty()) h.assign((size_t)S.n * S.bx, 0.0);
        upload_vec(*x->v, h, S.n, S.bx);
    }
    API2_END
}

AMGX_RC AMGX_write_system(const AMGX_matrix_handle mtx, const AMGX_vector_handle rhs, const AMGX_vector_handle sol, const char *filename)
{
    API2_BEGIN
    if (!filename) fatal(AMGX_RC_BAD_PARAMETERS, "null file name");
    MatrixH *m = chk<MatrixH>(mtx, MAGIC_MTX, "matrix");
    Matrix &A = *m->m;
    AMGXB_CUDA_CHECK(cudaSetDevice(A.rsc->device));
    if (A.dist) fatal(AMGX_RC_NOT_IMPLEMENTED, "write_system of a distributed matrix: use AMGX_write_system_distributed");
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
    std::vector<double> vall;
    if (A.has_ext_diag) {
        vall.resize((size_t)(A.nnz + A.n) * bsq);
        if (A.mat_prec == Prec::F64) AMGXB_CUDA_CHECK(cudaMemcpy(vall.data(), A.values.ptr(), vall.size() * 8, cudaMemcpyDeviceToHost));
        else {
            std::vector<float> vf(vall.size());
            AMGXB_CUDA_CHECK(cudaMemcpy(vf.data(), A.values.ptr(), vf.size() * 4, cudaMemcpyDeviceToHost));
            std::copy(vf.begin(), vf.end(), vall.begin());
        }
    }
    f.close();
    write_host_system(filename, A.rsc->cfg ? A.rsc->cfg->get_string("matrix_writer", "default") : std::string("matrixmarket"), A.n, A.nnz, A.bx, A.by, rp, ci,
                      A.has_ext_diag ? vall : va, A.has_ext_diag, b, x);
    API2_END
}

AMGX_RC AMGXB200_get_nccl_unique_id(char *id128)
{
    API2_BEGIN
    if (!id128) fatal(AMGX_RC_BAD_PARAMETERS, "null pointer");
    dist_get_unique_id(id128);
    API2_END
}

// which scalar CSR kernel family the matrix was planned for (tests assert that the path they mean to cover is the one that runs)
AMGX_RC AMGXB200_matrix_get_kernel_plan(AMGX_matrix_handle mtx, int *tile_rows, int *coded_tiles, int *pair_tiles, int *row_pattern_tiles, int *window)
{
    API2_BEGIN
    MatrixH *m = chk<MatrixH>(mtx, MAGIC_MTX, "matrix");
    const Matrix &A = *m->m;
    if (!A.initialized) fatal(AMGX_RC_BAD_PARAMETERS, "matrix not uploaded");
    if (tile_rows) *tile_rows = A.plan.use_tiles ? A.plan.tile_rows : 0;
    if (coded_tiles) *coded_tiles = A.colenc.on ? A.colenc.tiles_dict8 + A.colenc.tiles_off16 : 0;
    if (pair_tiles) *pair_tiles = A.colenc.on ? A.colenc.tiles_pair : 0;
    if (row_pattern_tiles) *row_pattern_tiles = A.colenc.on ? A.colenc.tiles_rowpat : 0;
    if (window) *window = A.win.on ? A.win.ring : 0;
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
