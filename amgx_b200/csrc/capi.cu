// capi.cu -- the C-ABI of libamgx_b200.so: the reference's AMGX_* entry points for the solve path
// (declared in include/amgx_b200.h with the reference lines they replace) plus AMGXB200_*
// extensions.  C++ exceptions never cross this boundary: they are mapped to AMGX_RC here, the way
// the reference's AMGX_TRIES / AMGX_CATCHES / getCAPIerror_x do (include/amgx_c_common.h:26-47).
#include "solvers.h"
#include "dist.h"
#include "capi_internal.h"
#include <fstream>
#include <sstream>
#include <algorithm>
#include <random>

using namespace amgxb;

namespace amgxb {

static bool g_initialized = false;

static AMGX_RC handle_exception(const char *where, Resources *rsc)
{
    AMGX_RC rc = AMGX_RC_UNKNOWN;
    std::string msg;
    try { throw; }
    catch (const Error &e) { rc = e.rc; msg = e.msg; }
    catch (const std::bad_alloc &) { rc = AMGX_RC_NO_MEMORY; msg = "Not enough memory"; }
    catch (const std::exception &e) { rc = AMGX_RC_UNKNOWN; msg = e.what(); }
    catch (...) { rc = AMGX_RC_UNKNOWN; msg = "unknown exception"; }
    std::string full = std::string("AMGX error in ") + where + ": " + msg + "\n";
    amgx_output(full.c_str(), (int)full.size());
    if (rsc && rsc->cfg && rsc->cfg->get_int("exception_handling", "default") == 1) {
        // internal error handling requested: print and terminate (amgx_error_exit)
        fprintf(stderr, "%s", full.c_str());
        exit(1);
    }
    return rc;
}

#define API_BEGIN try {
#define API_END(rsc_ptr)                                                      \
    }                                                                         \
    catch (...) { return handle_exception(__func__, (rsc_ptr)); }             \
    return AMGX_RC_OK;

template <class H> static H *check(void *p, unsigned magic, const char *what)
{
    H *h = reinterpret_cast<H *>(p);
    if (!h || h->magic != magic) fatal(AMGX_RC_BAD_PARAMETERS, std::string("invalid ") + what + " handle");
    return h;
}
static ConfigH *cfgH(AMGX_config_handle h) { return check<ConfigH>(h, MAGIC_CFG, "config"); }
static ResourcesH *rscH(AMGX_resources_handle h) { return check<ResourcesH>(h, MAGIC_RSC, "resources"); }
static MatrixH *mtxH(AMGX_matrix_handle h) { return check<MatrixH>(h, MAGIC_MTX, "matrix"); }
static VectorH *vecH(AMGX_vector_handle h) { return check<VectorH>(h, MAGIC_VEC, "vector"); }
static SolverH *slvH(AMGX_solver_handle h) { return check<SolverH>(h, MAGIC_SLV, "solver"); }

static void use_device(const std::shared_ptr<Resources> &r) { AMGXB_CUDA_CHECK(cudaSetDevice(r->device)); }

// ---- scalar matrices with an external diagonal are merged into plain CSR (diagonal first in each
// row, the layout the reference's own Poisson generator uses) so that one kernel family serves them.
void upload_matrix(Matrix &A, int n, int nnz, int bx, int by, const int *row_ptrs, const int *col_indices, const void *data,
                   const void *diag_data)
{
    if (n < 1 || nnz < 0 || bx < 1 || by < 1) fatal(AMGX_RC_BAD_PARAMETERS, "Error: Failure in matrix_upload_all().");
    if (bx != by) fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "rectangular blocks are not supported");
    cudaStream_t s = A.stream();
    const size_t bs = (size_t)bx * by, msz = prec_size(A.mat_prec);
    A.initialized = false;
    A.n = n;
    A.n_cols = n;
    A.bx = bx;
    A.by = by;
    A.dist.reset();
    if (diag_data && bs == 1) {
        // merge on the host (setup-time path)
        std::vector<int> rp(n + 1), ci(nnz);
        std::vector<char> va((size_t)nnz * msz), dg((size_t)n * msz);
        AMGXB_CUDA_CHECK(cudaMemcpy(rp.data(), row_ptrs, sizeof(int) * (n + 1), cudaMemcpyDefault));
        if (nnz) AMGXB_CUDA_CHECK(cudaMemcpy(ci.data(), col_indices, sizeof(int) * nnz, cudaMemcpyDefault));
        if (nnz) AMGXB_CUDA_CHECK(cudaMemcpy(va.data(), data, msz * nnz, cudaMemcpyDefault));
        AMGXB_CUDA_CHECK(cudaMemcpy(dg.data(), diag_data, msz * n, cudaMemcpyDefault));
        std::vector<int> rp2(n + 1), ci2((size_t)nnz + n);
        std::vector<char> va2(((size_t)nnz + n) * msz);
        size_t o = 0;
        for (int i = 0; i < n; i++) {
            rp2[i] = (int)o;
            ci2[o] = i;
            memcpy(&va2[o * msz], &dg[(size_t)i * msz], msz);
            o++;
            for (int k = rp[i]; k < rp[i + 1]; k++) {
                ci2[o] = ci[k];
                memcpy(&va2[o * msz], &va[(size_t)k * msz], msz);
                o++;
            }
        }
        rp2[n] = (int)o;
        A.nnz = nnz + n;
        A.has_ext_diag = false;
        A.merged_ext_diag = true;
        A.row_ptr.from_any(rp2.data(), n + 1, s);
        A.col_idx.from_any(ci2.data(), A.nnz, s);
        A.values.resize((size_t)A.nnz, A.mat_prec);
        AMGXB_CUDA_CHECK(cudaMemcpyAsync(A.values.ptr(), va2.data(), (size_t)A.nnz * msz, cudaMemcpyHostToDevice, s));
        AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    } else {
        A.nnz = nnz;
        A.has_ext_diag = diag_data != nullptr;
        A.merged_ext_diag = false;
        A.row_ptr.from_any(row_ptrs, n + 1, s);
        A.col_idx.from_any(col_indices, nnz, s);
        const size_t nblocks = (size_t)nnz + (diag_data ? n : 0);
        A.values.resize(nblocks * bs, A.mat_prec);
        if (nnz) AMGXB_CUDA_CHECK(cudaMemcpyAsync(A.values.ptr(), data, (size_t)nnz * bs * msz, cudaMemcpyDefault, s));
        if (diag_data)
            AMGXB_CUDA_CHECK(cudaMemcpyAsync((char *)A.values.ptr() + (size_t)nnz * bs * msz, diag_data, (size_t)n * bs * msz, cudaMemcpyDefault, s));
        AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));   // uploads copy synchronously (src/amgx_c.cu:894-907)
    }
    A.compute_diag_and_plan();
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
}

}  // namespace amgxb

extern "C" {

// ---------------------------------------------------------------------------------------------
// build / init / system
// ---------------------------------------------------------------------------------------------
AMGX_RC AMGX_get_api_version(int *major, int *minor)
{
    if (!major || !minor) return AMGX_RC_BAD_PARAMETERS;
    *major = 1;   // src/api_version.cu
    *minor = 0;
    return AMGX_RC_OK;
}

AMGX_RC AMGX_get_build_info_strings(char **version, char **date, char **time)
{
    static char v[] = "2.5.0-b200", d[] = __DATE__, t[] = __TIME__;
    if (version) *version = v;
    if (date) *date = d;
    if (time) *time = t;
    return AMGX_RC_OK;
}

AMGX_RC AMGX_get_error_string(AMGX_RC err, char *buf, int buf_len)
{
    static const char *msgs[] = {"No error.", "Incorrect parameters for amgx call.", "Unknown error.", "Unsupported device/host algorithm.",
                                 "Unsupported block size for the algorithm.", "CUDA kernel launch error.", "Thrust failure.",
                                 "Insufficient memory.", "I/O error.", "Incorrect C API mode.", "Error initializing amgx core.",
                                 "Error initializing plugins.", "Incorrect amgx configuration provided.",
                                 "Configuration feature is not implemented.", "Valid license is not found.", "Internal error."};
    if (!buf || buf_len < 1) return AMGX_RC_BAD_PARAMETERS;
    const char *m = ((int)err >= 0 && (int)err <= 15) ? msgs[(int)err] : msgs[2];
    strncpy(buf, m, buf_len);
    buf[buf_len - 1] = 0;
    return AMGX_RC_OK;
}

AMGX_RC AMGX_initialize(void)
{
    g_initialized = true;
    return AMGX_RC_OK;
}
AMGX_RC AMGX_initialize_plugins(void) { return AMGX_RC_OK; }
AMGX_RC AMGX_finalize(void)
{
    g_initialized = false;
    return AMGX_RC_OK;
}
AMGX_RC AMGX_finalize_plugins(void) { return AMGX_RC_OK; }
void AMGX_abort(AMGX_resources_handle, int err) { exit(err); }
AMGX_RC AMGX_pin_memory(void *ptr, unsigned int bytes)
{
    if (bytes > 0) {
        cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterMapped);
        if (e != cudaSuccess) { cudaGetLastError(); return AMGX_RC_CUDA_FAILURE; }
    }
    return AMGX_RC_OK;
}
AMGX_RC AMGX_unpin_memory(void *ptr)
{
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return AMGX_RC_CUDA_FAILURE; }
    return AMGX_RC_OK;
}
AMGX_RC AMGX_install_signal_handler(void) { return AMGX_RC_OK; }
AMGX_RC AMGX_reset_signal_handler(void) { return AMGX_RC_OK; }
AMGX_RC AMGX_register_print_callback(AMGX_print_callback func)
{
    set_print_callback(func);
    return AMGX_RC_OK;
}
AMGX_RC AMGX_solver_register_print_callback(AMGX_print_callback func) { return AMGX_register_print_callback(func); }

// ---------------------------------------------------------------------------------------------
// config
// ---------------------------------------------------------------------------------------------
AMGX_RC AMGX_config_create(AMGX_config_handle *cfg, const char *options)
{
    API_BEGIN
    if (!cfg) fatal(AMGX_RC_BAD_PARAMETERS, "null handle pointer");
    std::unique_ptr<ConfigH> h(new ConfigH);
    h->cfg = std::make_shared<Config>();
    h->cfg->parse_string(options);
    *cfg = reinterpret_cast<AMGX_config_handle>(h.release());
    API_END(nullptr)
}

AMGX_RC AMGX_config_add_parameters(AMGX_config_handle *cfg, const char *options)
{
    API_BEGIN
    if (!cfg) fatal(AMGX_RC_BAD_PARAMETERS, "null handle pointer");
    ConfigH *h = cfgH(*cfg);
    h->cfg->allow_mod = true;
    try { h->cfg->parse_string(options); }
    catch (...) { h->cfg->allow_mod = false; throw; }
    h->cfg->allow_mod = false;
    API_END(nullptr)
}

AMGX_RC AMGX_config_create_from_file(AMGX_config_handle *cfg, const char *param_file)
{
    API_BEGIN
    if (!cfg) fatal(AMGX_RC_BAD_PARAMETERS, "null handle pointer");
    std::unique_ptr<ConfigH> h(new ConfigH);
    h->cfg = std::make_shared<Config>();
    h->cfg->parse_file(param_file);
    *cfg = reinterpret_cast<AMGX_config_handle>(h.release());
    API_END(nullptr)
}

AMGX_RC AMGX_config_create_from_file_and_string(AMGX_config_handle *cfg, const char *param_file, const char *options)
{
    API_BEGIN
    if (!cfg) fatal(AMGX_RC_BAD_PARAMETERS, "null handle pointer");
    std::unique_ptr<ConfigH> h(new ConfigH);
    h->cfg = std::make_shared<Config>();
    h->cfg->parse_file(param_file);
    h->cfg->allow_mod = true;
    h->cfg->parse_string(options);
    h->cfg->allow_mod = false;
    *cfg = reinterpret_cast<AMGX_config_handle>(h.release());
    API_END(nullptr)
}

AMGX_RC AMGX_config_get_default_number_of_rings(AMGX_config_handle cfg, int *num_import_rings)
{
    API_BEGIN
    ConfigH *h = cfgH(cfg);
    if (!num_import_rings) fatal(AMGX_RC_BAD_PARAMETERS, "null pointer");
    // 2 rings for CLASSICAL AMG as solver or preconditioner, else 1 (src/amgx_c.cu:2529-2590)
    std::string sv, ss, pv, ps;
    h->cfg->get_scoped("solver", "default", sv, ss);
    std::string alg_s = h->cfg->get_string("algorithm", ss);
    h->cfg->get_scoped("preconditioner", ss, pv, ps);
    std::string alg_p = h->cfg->get_string("algorithm", ps);
    if (sv == "AMG") *num_import_rings = (alg_s == "CLASSICAL") ? 2 : 1;
    else if (pv == "AMG") *num_import_rings = (alg_p == "CLASSICAL") ? 2 : 1;
    else *num_import_rings = 1;
    API_END(nullptr)
}

AMGX_RC AMGX_config_destroy(AMGX_config_handle cfg)
{
    API_BEGIN
    ConfigH *h = cfgH(cfg);
    h->magic = 0;
    delete h;
    API_END(nullptr)
}

// ---------------------------------------------------------------------------------------------
// resources
// ---------------------------------------------------------------------------------------------
static std::shared_ptr<Resources> make_resources(const std::shared_ptr<Config> &cfg, int device, const AMGXB200_comm *comm)
{
    auto r = std::make_shared<Resources>();
    r->cfg = cfg;
    r->device = device;
    AMGXB_CUDA_CHECK(cudaSetDevice(device));
    AMGXB_CUDA_CHECK(cudaFree(0));
    cudaDeviceProp prop;
    AMGXB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    r->num_sms = prop.multiProcessorCount;
    AMGXB_CUDA_CHECK(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    AMGXB_CUDA_CHECK(cudaStreamCreateWithFlags(&r->side_stream, cudaStreamNonBlocking));
    if (comm && comm->world_size > 1) dist_init_comm(r.get(), comm);
    return r;
}

AMGX_RC AMGX_resources_create(AMGX_resources_handle *rsc, AMGX_config_handle cfg, void *comm, int device_num, const int *devices)
{
    API_BEGIN
    if (!rsc) fatal(AMGX_RC_BAD_PARAMETERS, "null handle pointer");
    ConfigH *c = cfgH(cfg);
    if (device_num != 1 || !devices) fatal(AMGX_RC_BAD_PARAMETERS, "exactly one device per process is supported");
    std::unique_ptr<ResourcesH> h(new ResourcesH);
    h->rsc = make_resources(c->cfg, devices[0], reinterpret_cast<const AMGXB200_comm *>(comm));
    *rsc = reinterpret_cast<AMGX_resources_handle>(h.release());
    API_END(nullptr)
}

AMGX_RC AMGX_resources_create_simple(AMGX_resources_handle *rsc, AMGX_config_handle cfg)
{
    API_BEGIN
    if (!rsc) fatal(AMGX_RC_BAD_PARAMETERS, "null handle pointer");
    ConfigH *c = cfgH(cfg);
    std::unique_ptr<ResourcesH> h(new ResourcesH);
    h->rsc = make_resources(c->cfg, 0, nullptr);
    *rsc = reinterpret_cast<AMGX_resources_handle>(h.release());
    API_END(nullptr)
}

AMGX_RC AMGX_resources_destroy(AMGX_resources_handle rsc)
{
    API_BEGIN
    ResourcesH *h = rscH(rsc);
    h->magic = 0;
    delete h;
    API_END(nullptr)
}

// ---------------------------------------------------------------------------------------------
// matrix
// ---------------------------------------------------------------------------------------------
AMGX_RC AMGX_matrix_create(AMGX_matrix_handle *mtx, AMGX_resources_handle rsc, AMGX_Mode mode)
{
    Resources *rp = nullptr;
    API_BEGIN
    if (!mtx) fatal(AMGX_RC_BAD_PARAMETERS, "null handle pointer");
    ResourcesH *r = rscH(rsc);
    rp = r->rsc.get();
    ModeInfo mi = decode_mode((int)mode);
    std::unique_ptr<MatrixH> h(new MatrixH);
    h->m = std::make_shared<Matrix>();
    h->m->rsc = r->rsc;
    h->m->mode = (int)mode;
    h->m->mat_prec = mi.mat;
    h->m->vec_prec = mi.vec;
    *mtx = reinterpret_cast<AMGX_matrix_handle>(h.release());
    API_END(rp)
}

AMGX_RC AMGX_matrix_destroy(AMGX_matrix_handle mtx)
{
    API_BEGIN
    MatrixH *h = mtxH(mtx);
    h->magic = 0;
    delete h;
    API_END(nullptr)
}

AMGX_RC AMGX_matrix_upload_all(AMGX_matrix_handle mtx, int n, int nnz, int block_dimx, int block_dimy, const int *row_ptrs,
                               const int *col_indices, const void *data, const void *diag_data)
{
    Resources *rp = nullptr;
    API_BEGIN
    MatrixH *h = mtxH(mtx);
    rp = h->m->rsc.get();
    use_device(h->m->rsc);
    if (h->m->dist_pending) dist_upload_local(*h->m, n, nnz, block_dimx, block_dimy, row_ptrs, col_indices, data, diag_data);
    else upload_matrix(*h->m, n, nnz, block_dimx, block_dimy, row_ptrs, col_indices, data, diag_data);
    API_END(rp)
}

AMGX_RC AMGX_matrix_replace_coefficients(AMGX_matrix_handle mtx, int n, int nnz, const void *data, const void *diag_data)
{
    Resources *rp = nullptr;
    API_BEGIN
    MatrixH *h = mtxH(mtx);
    Matrix &A = *h->m;
    rp = A.rsc.get();
    use_device(A.rsc);
    if (!A.initialized) fatal(AMGX_RC_BAD_PARAMETERS, "matrix not initialized");
    if (A.merged_ext_diag) fatal(AMGX_RC_NOT_IMPLEMENTED, "replace_coefficients on a matrix uploaded with a separate scalar diagonal");
    const int user_nnz = A.nnz;
    if (n != A.n || nnz != user_nnz) fatal(AMGX_RC_BAD_PARAMETERS, "replace_coefficients: size mismatch");
    if (A.dist) {
        if (diag_data) fatal(AMGX_RC_NOT_IMPLEMENTED, "replace_coefficients with an external diagonal on a distributed matrix");
        if (data) {
            dist_replace_values(A, nnz, data);
            csr_values_changed(A, A.stream());      // the value codes of the tile kernels follow the values
            AMGXB_CUDA_CHECK(cudaStreamSynchronize(A.stream()));
        }
        return AMGX_RC_OK;
    }
    const size_t bs = A.bs(), msz = prec_size(A.mat_prec);
    if (data) AMGXB_CUDA_CHECK(cudaMemcpyAsync(A.values.ptr(), data, (size_t)nnz * bs * msz, cudaMemcpyDefault, A.stream()));
    if (diag_data && A.has_ext_diag)
        AMGXB_CUDA_CHECK(cudaMemcpyAsync((char *)A.values.ptr() + (size_t)nnz * bs * msz, diag_data, (size_t)n * bs * msz, cudaMemcpyDefault, A.stream()));
    if (data) csr_values_changed(A, A.stream());
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(A.stream()));
    API_END(rp)
}

AMGX_RC AMGX_matrix_get_size(const AMGX_matrix_handle mtx, int *n, int *block_dimx, int *block_dimy)
{
    API_BEGIN
    MatrixH *h = mtxH(mtx);
    if (n) *n = h->m->n;
    if (block_dimx) *block_dimx = h->m->bx;
    if (block_dimy) *block_dimy = h->m->by;
    API_END(nullptr)
}

AMGX_RC AMGX_matrix_get_nnz(const AMGX_matrix_handle mtx, int *nnz)
{
    API_BEGIN
    MatrixH *h = mtxH(mtx);
    if (nnz) *nnz = h->m->merged_ext_diag ? h->m->nnz - h->m->n : h->m->nnz;
    API_END(nullptr)
}

AMGX_RC AMGX_matrix_download_all(const AMGX_matrix_handle mtx, int *row_ptrs, int *col_indices, void *data, void **diag_data)
{
    Resources *rp = nullptr;
    API_BEGIN
    MatrixH *h = mtxH(mtx);
    Matrix &A = *h->m;
    rp = A.rsc.get();
    use_device(A.rsc);
    if (A.dist) fatal(AMGX_RC_NOT_IMPLEMENTED, "download of a distributed matrix");
    const size_t bs = A.bs(), msz = prec_size(A.mat_prec);
    if (A.merged_ext_diag) {
        // undo the merge: the first entry of each row is the former external diagonal
        std::vector<int> rp2 = A.row_ptr.to_host(A.stream()), ci2 = A.col_idx.to_host(A.stream());
        std::vector<char> va2((size_t)A.nnz * msz);
        AMGXB_CUDA_CHECK(cudaMemcpy(va2.data(), A.values.ptr(), va2.size(), cudaMemcpyDeviceToHost));
        char *dg = (char *)malloc((size_t)A.n * msz);
        size_t o = 0;
        for (int i = 0; i < A.n; i++) {
            row_ptrs[i] = (int)o;
            memcpy(dg + (size_t)i * msz, &va2[(size_t)rp2[i] * msz], msz);
            for (int k = rp2[i] + 1; k < rp2[i + 1]; k++) {
                col_indices[o] = ci2[k];
                memcpy((char *)data + o * msz, &va2[(size_t)k * msz], msz);
                o++;
            }
        }
        row_ptrs[A.n] = (int)o;
        if (diag_data) *diag_data = dg; else free(dg);
    } else {
        AMGXB_CUDA_CHECK(cudaMemcpy(row_ptrs, A.row_ptr.ptr(), sizeof(int) * (A.n + 1), cudaMemcpyDefault));
        if (A.nnz) AMGXB_CUDA_CHECK(cudaMemcpy(col_indices, A.col_idx.ptr(), sizeof(int) * A.nnz, cudaMemcpyDefault));
        if (A.nnz) AMGXB_CUDA_CHECK(cudaMemcpy(data, A.values.ptr(), (size_t)A.nnz * bs * msz, cudaMemcpyDefault));
        if (diag_data) {
            *diag_data = nullptr;
            if (A.has_ext_diag) {
                *diag_data = malloc((size_t)A.n * bs * msz);
                AMGXB_CUDA_CHECK(cudaMemcpy(*diag_data, (char *)A.values.ptr() + (size_t)A.nnz * bs * msz, (size_t)A.n * bs * msz, cudaMemcpyDefault));
            }
        }
    }
    API_END(rp)
}

AMGX_RC AMGX_matrix_vector_multiply(AMGX_matrix_handle mtx, AMGX_vector_handle x, AMGX_vector_handle y)
{
    Resources *rp = nullptr;
    API_BEGIN
    MatrixH *hm = mtxH(mtx);
    VectorH *hx = vecH(x), *hy = vecH(y);
    Matrix &A = *hm->m;
    rp = A.rsc.get();
    use_device(A.rsc);
    if (hx->v->mode != A.mode || hy->v->mode != A.mode) fatal(AMGX_RC_BAD_PARAMETERS, "Error: mismatch between Matrix mode and Vector Mode.");
    if (!A.initialized) fatal(AMGX_RC_BAD_PARAMETERS, "matrix not initialized");
    dist_prepare_vector(A, *hx->v);
    const size_t need = (size_t)A.n_cols * A.by;
    if (hy->v->data.n < need || hy->v->n != A.n) {
        hy->v->data.resize(need, A.vec_prec);
        hy->v->n = A.n;
        hy->v->block_dim = A.by;
        hy->v->prec = A.vec_prec;
    }
    if (A.dist) { hy->v->dist = A.dist; hy->v->user_order = false; }
    if ((size_t)hx->v->n * hx->v->block_dim != (size_t)A.n * A.bx) fatal(AMGX_RC_BAD_PARAMETERS, "x size does not match the matrix");
    cudaStream_t s = A.stream();
    dist_exchange_halo(A, hx->v->data, s);
    CsrOpArgs g;
    g.x = hx->v->data.ptr();
    g.y = hy->v->data.ptr();
    matrix_apply(A, EPI_SPMV, g, s);
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    API_END(rp)
}

AMGX_RC AMGX_matrix_set_boundary_separation(AMGX_matrix_handle mtx, int) { (void)mtx; return AMGX_RC_OK; }

AMGX_RC AMGX_matrix_attach_coloring(AMGX_matrix_handle mtx, int *row_coloring, int num_rows, int num_colors)
{
    Resources *rp = nullptr;
    API_BEGIN
    MatrixH *h = mtxH(mtx);
    rp = h->m->rsc.get();
    use_device(h->m->rsc);
    attach_user_coloring(*h->m, row_coloring, num_rows, num_colors);
    API_END(rp)
}

AMGX_RC AMGX_matrix_attach_geometry(AMGX_matrix_handle, double *, double *, double *, int) { return AMGX_RC_OK; }

AMGX_RC AMGX_matrix_check_symmetry(AMGX_matrix_handle mtx, int *structurally_symmetric, int *symmetric)
{
    Resources *rp = nullptr;
    API_BEGIN
    MatrixH *h = mtxH(mtx);
    Matrix &A = *h->m;
    rp = A.rsc.get();
    use_device(A.rsc);
    if (A.bs() != 1 || A.dist) fatal(AMGX_RC_NOT_IMPLEMENTED, "check_symmetry: scalar single-GPU matrices only");
    std::vector<int> rp_h = A.row_ptr.to_host(A.stream()), ci = A.col_idx.to_host(A.stream());
    std::vector<double> va(A.nnz);
    if (A.mat_prec == Prec::F64) AMGXB_CUDA_CHECK(cudaMemcpy(va.data(), A.values.ptr(), sizeof(double) * A.nnz, cudaMemcpyDeviceToHost));
    else {
        std::vector<float> vf(A.nnz);
        AMGXB_CUDA_CHECK(cudaMemcpy(vf.data(), A.values.ptr(), sizeof(float) * A.nnz, cudaMemcpyDeviceToHost));
        for (int k = 0; k < A.nnz; k++) va[k] = vf[k];
    }
    bool ss = true, sy = true;
    for (int i = 0; i < A.n && ss; i++)
        for (int k = rp_h[i]; k < rp_h[i + 1]; k++) {
            const int j = ci[k];
            bool found = false;
            for (int kk = rp_h[j]; kk < rp_h[j + 1]; kk++)
                if (ci[kk] == i) { found = true; if (va[kk] != va[k]) sy = false; break; }
            if (!found) { ss = false; sy = false; break; }
        }
    if (structurally_symmetric) *structurally_symmetric = ss;
    if (symmetric) *symmetric = sy;
    API_END(rp)
}

AMGX_RC AMGX_matrix_check_diag_dominant(const AMGX_matrix_handle mtx, int *diag_dominant)
{
    Resources *rp = nullptr;
    API_BEGIN
    MatrixH *h = mtxH(mtx);
    Matrix &A = *h->m;
    rp = A.rsc.get();
    use_device(A.rsc);
    if (A.bs() != 1 || A.dist || A.mat_prec != Prec::F64) fatal(AMGX_RC_NOT_IMPLEMENTED, "check_diag_dominant: scalar fp64 single-GPU matrices only");
    std::vector<int> rp_h = A.row_ptr.to_host(A.stream()), ci = A.col_idx.to_host(A.stream());
    std::vector<double> va(A.nnz);
    AMGXB_CUDA_CHECK(cudaMemcpy(va.data(), A.values.ptr(), sizeof(double) * A.nnz, cudaMemcpyDeviceToHost));
    bool dd = true;
    for (int i = 0; i < A.n; i++) {
        double d = 0, o = 0;
        for (int k = rp_h[i]; k < rp_h[i + 1]; k++) { if (ci[k] == i) d = fabs(va[k]); else o += fabs(va[k]); }
        if (d < o) { dd = false; break; }
    }
    if (diag_dominant) *diag_dominant = dd;
    API_END(rp)
}

// ---------------------------------------------------------------------------------------------
// vector
// ---------------------------------------------------------------------------------------------
AMGX_RC AMGX_vector_create(AMGX_vector_handle *vec, AMGX_resources_handle rsc, AMGX_Mode mode)
{
    Resources *rp = nullptr;
    API_BEGIN
    if (!vec) fatal(AMGX_RC_BAD_PARAMETERS, "null handle pointer");
    ResourcesH *r = rscH(rsc);
    rp = r->rsc.get();
    ModeInfo mi = decode_mode((int)mode);
    std::unique_ptr<VectorH> h(new VectorH);
    h->v = std::make_shared<Vector>();
    h->v->rsc = r->rsc;
    h->v->mode = (int)mode;
    h->v->prec = mi.vec;
    *vec = reinterpret_cast<AMGX_vector_handle>(h.release());
    API_END(rp)
}

AMGX_RC AMGX_vector_destroy(AMGX_vector_handle vec)
{
    API_BEGIN
    VectorH *h = vecH(vec);
    h->magic = 0;
    delete h;
    API_END(nullptr)
}

AMGX_RC AMGX_vector_upload(AMGX_vector_handle vec, int n, int block_dim, const void *data)
{
    Resources *rp = nullptr;
    API_BEGIN
    VectorH *h = vecH(vec);
    Vector &v = *h->v;
    rp = v.rsc.get();
    use_device(v.rsc);
    if (n < 0 || block_dim < 1) fatal(AMGX_RC_BAD_PARAMETERS, "vector_upload: bad sizes");
    v.n = n;
    v.block_dim = block_dim;
    const size_t len = (size_t)n * block_dim;
    const size_t alloc = v.dist ? std::max(len, (size_t)(v.dist->n_owned + v.dist->n_halo) * block_dim) : len;
    v.data.resize(alloc, v.prec);
    if (alloc > len) v.data.zero(v.rsc->stream);
    if (len) AMGXB_CUDA_CHECK(cudaMemcpyAsync(v.data.ptr(), data, len * prec_size(v.prec), cudaMemcpyDefault, v.rsc->stream));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(v.rsc->stream));
    v.user_order = true;
    API_END(rp)
}

AMGX_RC AMGX_vector_set_zero(AMGX_vector_handle vec, int n, int block_dim)
{
    Resources *rp = nullptr;
    API_BEGIN
    VectorH *h = vecH(vec);
    Vector &v = *h->v;
    rp = v.rsc.get();
    use_device(v.rsc);
    if (n < 0 || block_dim < 1) fatal(AMGX_RC_BAD_PARAMETERS, "vector_set_zero: bad sizes");
    v.n = n;
    v.block_dim = block_dim;
    const size_t len = (size_t)n * block_dim;
    const size_t alloc = v.dist ? std::max(len, (size_t)(v.dist->n_owned + v.dist->n_halo) * block_dim) : len;
    v.data.resize(alloc, v.prec);
    v.data.zero(v.rsc->stream);
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(v.rsc->stream));
    v.user_order = !v.dist;   // all zeros: any order
    API_END(rp)
}

AMGX_RC AMGX_vector_set_random(AMGX_vector_handle vec, int n)
{
    Resources *rp = nullptr;
    API_BEGIN
    VectorH *h = vecH(vec);
    Vector &v = *h->v;
    rp = v.rsc.get();
    use_device(v.rsc);
    if (n < 0) fatal(AMGX_RC_BAD_PARAMETERS, "vector_set_random: bad size");
    const int bd = std::max(1, v.block_dim);
    std::mt19937 gen(12345u);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    const size_t len = (size_t)n * bd;
    v.n = n;
    v.block_dim = bd;
    v.data.resize(len, v.prec);
    if (v.prec == Prec::F64) {
        std::vector<double> hbuf(len);
        for (auto &x : hbuf) x = U(gen);
        AMGXB_CUDA_CHECK(cudaMemcpy(v.data.ptr(), hbuf.data(), len * 8, cudaMemcpyHostToDevice));
    } else {
        std::vector<float> hbuf(len);
        for (auto &x : hbuf) x = (float)U(gen);
        AMGXB_CUDA_CHECK(cudaMemcpy(v.data.ptr(), hbuf.data(), len * 4, cudaMemcpyHostToDevice));
    }
    v.user_order = true;
    API_END(rp)
}

AMGX_RC AMGX_vector_download(const AMGX_vector_handle vec, void *data)
{
    Resources *rp = nullptr;
    API_BEGIN
    VectorH *h = vecH(vec);
    Vector &v = *h->v;
    rp = v.rsc.get();
    use_device(v.rsc);
    const size_t len = (size_t)v.n * v.block_dim;
    if (v.dist && !v.user_order) dist_download_vector(v, data);
    else if (len) AMGXB_CUDA_CHECK(cudaMemcpy(data, v.data.ptr(), len * prec_size(v.prec), cudaMemcpyDefault));
    API_END(rp)
}

AMGX_RC AMGX_vector_get_size(const AMGX_vector_handle vec, int *n, int *block_dim)
{
    API_BEGIN
    VectorH *h = vecH(vec);
    if (n) *n = h->v->n;
    if (block_dim) *block_dim = h->v->block_dim;
    API_END(nullptr)
}

AMGX_RC AMGX_vector_bind(AMGX_vector_handle vec, const AMGX_matrix_handle mtx)
{
    API_BEGIN
    VectorH *h = vecH(vec);
    MatrixH *m = mtxH(mtx);
    h->v->dist = m->m->dist;
    API_END(nullptr)
}

// ---------------------------------------------------------------------------------------------
// solver
// ---------------------------------------------------------------------------------------------
AMGX_RC AMGX_solver_create(AMGX_solver_handle *slv, AMGX_resources_handle rsc, AMGX_Mode mode, const AMGX_config_handle cfg_solver)
{
    Resources *rp = nullptr;
    API_BEGIN
    if (!slv) fatal(AMGX_RC_BAD_PARAMETERS, "null handle pointer");
    ResourcesH *r = rscH(rsc);
    rp = r->rsc.get();
    ConfigH *c = cfgH(cfg_solver);
    use_device(r->rsc);
    decode_mode((int)mode);
    std::unique_ptr<SolverH> h(new SolverH);
    h->rsc = r->rsc;
    h->mode = (int)mode;
    h->cfg = std::make_shared<Config>(*c->cfg);   // the solver keeps its own copy (AMG_Solver(Resources*, AMG_Configuration&))
    h->solver = Solver::allocate(*h->cfg, "default", "solver", r->rsc);
    *slv = reinterpret_cast<AMGX_solver_handle>(h.release());
    API_END(rp)
}

/* Does every component this configuration names exist in the engine?  Instantiates the whole solver tree in dry-run mode (the
 * constructors parse and validate, no device resource is created): AMGX_RC_OK, or the return code AMGX_solver_create would give, with
 * the message in msg.  Needs no GPU.  What only a matrix can decide (block size, distribution, precision mode) is not covered. */
AMGX_RC AMGXB200_config_check(const AMGX_config_handle cfg, AMGX_Mode mode, char *msg, int msg_len)
{
    if (msg && msg_len > 0) msg[0] = 0;
    AMGX_RC rc = AMGX_RC_OK;
    std::string text;
    g_dry_run = true;
    try {
        ConfigH *c = cfgH(cfg);
        decode_mode((int)mode);
        auto rsc = std::make_shared<Resources>();
        rsc->cfg = c->cfg;
        auto own = std::make_shared<Config>(*c->cfg);
        std::unique_ptr<Solver> sv = Solver::allocate(*own, "default", "solver", rsc);
    } catch (const Error &e) { rc = e.rc; text = e.msg; }
    catch (const std::exception &e) { rc = AMGX_RC_UNKNOWN; text = e.what(); }
    catch (...) { rc = AMGX_RC_UNKNOWN; text = "unknown exception"; }
    g_dry_run = false;
    if (msg && msg_len > 0) { strncpy(msg, text.c_str(), (size_t)msg_len - 1); msg[msg_len - 1] = 0; }
    return rc;
}

AMGX_RC AMGX_solver_destroy(AMGX_solver_handle slv)
{
    API_BEGIN
    SolverH *h = slvH(slv);
    cudaSetDevice(h->rsc->device);
    h->magic = 0;
    delete h;
    API_END(nullptr)
}

static AMGX_RC solver_setup_impl(AMGX_solver_handle slv, AMGX_matrix_handle mtx, bool reuse)
{
    Resources *rp = nullptr;
    API_BEGIN
    SolverH *h = slvH(slv);
    MatrixH *m = mtxH(mtx);
    rp = h->rsc.get();
    use_device(h->rsc);
    if (m->m->mode != h->mode) fatal(AMGX_RC_BAD_PARAMETERS, "Error: mismatch between Matrix mode and Solver Mode.");
    if (m->m->rsc.get() != h->rsc.get()) fatal(AMGX_RC_BAD_PARAMETERS, "Error: Inconsistency between solver and matrix resources object, exiting");
    h->A = m->m;   // the solver shares ownership of the matrix after setup (src/amg_solver.cu:257-261)
    h->solver->setup(*h->A, reuse && h->was_setup);
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(h->rsc->stream));
    h->was_setup = true;
    API_END(rp)
}

AMGX_RC AMGX_solver_setup(AMGX_solver_handle slv, AMGX_matrix_handle mtx) { return solver_setup_impl(slv, mtx, false); }
AMGX_RC AMGX_solver_resetup(AMGX_solver_handle slv, AMGX_matrix_handle mtx) { return solver_setup_impl(slv, mtx, true); }

static AMGX_RC solver_solve_impl(AMGX_solver_handle slv, AMGX_vector_handle rhs, AMGX_vector_handle sol, bool xIsZero)
{
    Resources *rp = nullptr;
    API_BEGIN
    SolverH *h = slvH(slv);
    VectorH *b = vecH(rhs), *x = vecH(sol);
    rp = h->rsc.get();
    use_device(h->rsc);
    if (b->v->mode != h->mode) fatal(AMGX_RC_BAD_PARAMETERS, "Error: mismatch between RHS mode and Solver Mode.\n");
    if (b->v->mode != x->v->mode) fatal(AMGX_RC_BAD_PARAMETERS, "Error: mismatch between RHS mode and Sol Mode.\n");
    if (b->v->rsc.get() != h->rsc.get() || x->v->rsc.get() != h->rsc.get())
        fatal(AMGX_RC_BAD_PARAMETERS, "Error: Inconsistency between solver and rhs/sol resources object, exiting");
    if (!h->A) fatal(AMGX_RC_BAD_CONFIGURATION, "Error, setup must be called before calling solve");
    Matrix &A = *h->A;
    if (b->v->block_dim != A.by) fatal(AMGX_RC_BAD_PARAMETERS, "Block sizes do not match");
    if (b->v->n != A.n) fatal(AMGX_RC_BAD_PARAMETERS, "rhs size does not match the matrix");
    const size_t need = (size_t)A.n_cols * A.by;
    if (!xIsZero) {     // a non-zero initial guess is read: it must have the matrix's shape (a block_dim-1 upload for a 4x4 system must not pass)
        if (x->v->n != A.n) fatal(AMGX_RC_BAD_PARAMETERS, "solution size does not match the matrix");
        if (x->v->block_dim != A.bx || x->v->data.n < (size_t)A.n * A.bx) fatal(AMGX_RC_BAD_PARAMETERS, "Block sizes do not match");
    }
    dist_prepare_vector(A, *b->v);
    if (xIsZero && (x->v->n != A.n || x->v->data.n < need)) {
        x->v->n = A.n;
        x->v->block_dim = A.by;
        x->v->data.resize(need, A.vec_prec);
        x->v->data.zero(h->rsc->stream);
        x->v->dist = A.dist;
        x->v->user_order = !A.dist;
    } else {
        dist_prepare_vector(A, *x->v);
    }
    cudaEvent_t e0, e1;
    AMGXB_CUDA_CHECK(cudaEventCreate(&e0));
    AMGXB_CUDA_CHECK(cudaEventCreate(&e1));
    const long long launches0 = g_kernel_launches;
    AMGXB_CUDA_CHECK(cudaEventRecord(e0, h->rsc->stream));
    Status st;
    try { st = h->solver->solve(b->v->data, x->v->data, xIsZero); }
    catch (...) { cudaEventDestroy(e0); cudaEventDestroy(e1); h->last_status = ST_FAILED; throw; }
    AMGXB_CUDA_CHECK(cudaEventRecord(e1, h->rsc->stream));
    AMGXB_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    h->last_solve_seconds = ms * 1e-3;
    phase_report(h->rsc->stream, h->solver->get_num_iters());      // AMGXB_PHASE_TIMING=1 only
    h->last_solve_launches = g_kernel_launches - launches0;
    h->last_status = st;
    API_END(rp)
}

AMGX_RC AMGX_solver_solve(AMGX_solver_handle slv, AMGX_vector_handle rhs, AMGX_vector_handle sol) { return solver_solve_impl(slv, rhs, sol, false); }
AMGX_RC AMGX_solver_solve_with_0_initial_guess(AMGX_solver_handle slv, AMGX_vector_handle rhs, AMGX_vector_handle sol)
{
    return solver_solve_impl(slv, rhs, sol, true);
}

AMGX_RC AMGX_solver_get_iterations_number(AMGX_solver_handle slv, int *n)
{
    API_BEGIN
    SolverH *h = slvH(slv);
    if (n) *n = h->solver->get_num_iters();
    API_END(nullptr)
}

AMGX_RC AMGX_solver_get_iteration_residual(AMGX_solver_handle slv, int it, int idx, double *res)
{
    Resources *rp = nullptr;
    API_BEGIN
    SolverH *h = slvH(slv);
    rp = h->rsc.get();
    if (!res) fatal(AMGX_RC_BAD_PARAMETERS, "null pointer");
    *res = -1.;
    const std::vector<double> &r = h->solver->get_residual(it);
    if (idx < 0 || idx >= (int)r.size()) {
        amgx_printf("Incorrect block index");
        return AMGX_RC_BAD_PARAMETERS;
    }
    *res = r[idx];
    API_END(rp)
}

AMGX_RC AMGX_solver_get_status(AMGX_solver_handle slv, AMGX_SOLVE_STATUS *st)
{
    API_BEGIN
    SolverH *h = slvH(slv);
    if (!st) fatal(AMGX_RC_BAD_PARAMETERS, "null pointer");
    switch (h->last_status) {
    case ST_CONVERGED: *st = AMGX_SOLVE_SUCCESS; break;
    case ST_DIVERGED: *st = AMGX_SOLVE_DIVERGED; break;
    case ST_NOT_CONVERGED: *st = AMGX_SOLVE_NOT_CONVERGED; break;
    default: *st = AMGX_SOLVE_FAILED;
    }
    API_END(nullptr)
}

AMGX_RC AMGX_solver_calculate_residual_norm(AMGX_solver_handle solver, AMGX_matrix_handle mtx, AMGX_vector_handle rhs, AMGX_vector_handle x,
                                            void *norm_vector)
{
    Resources *rp = nullptr;
    API_BEGIN
    SolverH *h = slvH(solver);
    MatrixH *m = mtxH(mtx);
    VectorH *b = vecH(rhs), *xv = vecH(x);
    rp = h->rsc.get();
    use_device(h->rsc);
    std::vector<double> nrm;
    residual_norm_external(*h, *m->m, *b->v, *xv->v, nrm);
    if (m->m->vec_prec == Prec::F64) for (size_t i = 0; i < nrm.size(); i++) ((double *)norm_vector)[i] = nrm[i];
    else for (size_t i = 0; i < nrm.size(); i++) ((float *)norm_vector)[i] = (float)nrm[i];
    API_END(rp)
}

AMGX_RC AMGX_write_parameters_description(char *filename, AMGX_GET_PARAMS_DESC_FLAG mode)
{
    API_BEGIN
    if (mode != AMGX_GET_PARAMS_DESC_JSON_TO_FILE) fatal(AMGX_RC_NOT_IMPLEMENTED, "only AMGX_GET_PARAMS_DESC_JSON_TO_FILE is implemented");
    std::ofstream f(filename);
    if (!f) fatal(AMGX_RC_IO_ERROR, "cannot open output file");
    size_t n;
    const ParamDesc *r = Config::registry(&n);
    f << "{\n";
    for (size_t i = 0; i < n; i++) {
        const char *t = r[i].type == PType::INT ? "int" : r[i].type == PType::SIZE ? "size_t" : r[i].type == PType::DOUBLE ? "double" : "string";
        f << "  \"" << r[i].name << "\": {\"parameter_type\": \"" << t << "\", \"default_value\": ";
        if (r[i].type == PType::STRING) f << "\"" << r[i].def << "\"";
        else f << r[i].def;
        f << "}" << (i + 1 < n ? "," : "") << "\n";
    }
    f << "}\n";
    API_END(nullptr)
}

// ---------------------------------------------------------------------------------------------
// extensions: introspection
// ---------------------------------------------------------------------------------------------
static AMGSolver *find_amg(Solver *s)
{
    if (!s) return nullptr;
    if (auto *a = dynamic_cast<AMGSolver *>(s)) return a;
    if (auto *p = dynamic_cast<PCGSolver *>(s)) return find_amg(p->preconditioner());
    if (auto *f = dynamic_cast<FGMRESSolver *>(s)) return find_amg(f->preconditioner());
    return nullptr;
}

AMGX_RC AMGXB200_solver_get_num_levels(AMGX_solver_handle slv, int *num_levels)
{
    API_BEGIN
    SolverH *h = slvH(slv);
    AMGSolver *a = find_amg(h->solver.get());
    if (!a) fatal(AMGX_RC_BAD_PARAMETERS, "solver has no AMG hierarchy");
    *num_levels = a->num_levels();
    API_END(nullptr)
}

static AMGLevel &get_level(AMGX_solver_handle slv, int level, SolverH **hh = nullptr)
{
    SolverH *h = slvH(slv);
    if (hh) *hh = h;
    AMGSolver *a = find_amg(h->solver.get());
    if (!a) fatal(AMGX_RC_BAD_PARAMETERS, "solver has no AMG hierarchy");
    if (level < 0 || level >= a->num_levels()) fatal(AMGX_RC_BAD_PARAMETERS, "level out of range");
    AMGXB_CUDA_CHECK(cudaSetDevice(h->rsc->device));
    return a->level(level);
}

AMGX_RC AMGXB200_solver_get_level_info(AMGX_solver_handle slv, int level, int *n, int *nnz, int *block_dim, int *n_coarse)
{
    API_BEGIN
    AMGLevel &L = get_level(slv, level);
    if (n) *n = L.A->n;
    if (nnz) *nnz = L.A->nnz;
    if (block_dim) *block_dim = L.A->bx;
    if (n_coarse) *n_coarse = L.n_coarse;
    API_END(nullptr)
}

AMGX_RC AMGXB200_solver_get_level_matrix(AMGX_solver_handle slv, int level, int *row_ptrs, int *col_indices, void *values)
{
    API_BEGIN
    AMGLevel &L = get_level(slv, level);
    const Matrix &A = *L.A;
    if (row_ptrs) AMGXB_CUDA_CHECK(cudaMemcpy(row_ptrs, A.row_ptr.ptr(), sizeof(int) * (A.n + 1), cudaMemcpyDeviceToHost));
    if (col_indices && A.nnz) AMGXB_CUDA_CHECK(cudaMemcpy(col_indices, A.col_idx.ptr(), sizeof(int) * A.nnz, cudaMemcpyDeviceToHost));
    if (values && A.nnz) AMGXB_CUDA_CHECK(cudaMemcpy(values, A.values.ptr(), (size_t)A.nnz * A.bs() * prec_size(A.mat_prec), cudaMemcpyDeviceToHost));
    API_END(nullptr)
}

AMGX_RC AMGXB200_solver_get_level_aggregates(AMGX_solver_handle slv, int level, int *aggregates, int *R_row_offsets, int *R_column_indices)
{
    API_BEGIN
    AMGLevel &L = get_level(slv, level);
    if (L.aggregates.size() == 0) fatal(AMGX_RC_BAD_PARAMETERS, "level has no aggregates");
    if (aggregates) AMGXB_CUDA_CHECK(cudaMemcpy(aggregates, L.aggregates.ptr(), sizeof(int) * L.A->n, cudaMemcpyDeviceToHost));
    if (R_row_offsets) AMGXB_CUDA_CHECK(cudaMemcpy(R_row_offsets, L.R_row_offsets.ptr(), sizeof(int) * (L.n_coarse + 1), cudaMemcpyDeviceToHost));
    if (R_column_indices) AMGXB_CUDA_CHECK(cudaMemcpy(R_column_indices, L.R_column_indices.ptr(), sizeof(int) * L.A->n, cudaMemcpyDeviceToHost));
    API_END(nullptr)
}

static void copy_out_matrix(const Matrix *M, int *nnz, int *row_ptrs, int *col_indices, void *values)
{
    if (!M) fatal(AMGX_RC_BAD_PARAMETERS, "level has no such operator");
    if (nnz) *nnz = M->nnz;
    if (row_ptrs) AMGXB_CUDA_CHECK(cudaMemcpy(row_ptrs, M->row_ptr.ptr(), sizeof(int) * (M->n + 1), cudaMemcpyDeviceToHost));
    if (col_indices && M->nnz) AMGXB_CUDA_CHECK(cudaMemcpy(col_indices, M->col_idx.ptr(), sizeof(int) * M->nnz, cudaMemcpyDeviceToHost));
    if (values && M->nnz) AMGXB_CUDA_CHECK(cudaMemcpy(values, M->values.ptr(), (size_t)M->nnz * prec_size(M->mat_prec), cudaMemcpyDeviceToHost));
}

AMGX_RC AMGXB200_solver_get_level_P(AMGX_solver_handle slv, int level, int *nnz, int *row_ptrs, int *col_indices, void *values)
{
    API_BEGIN
    AMGLevel &L = get_level(slv, level);
    copy_out_matrix(L.P.get(), nnz, row_ptrs, col_indices, values);
    API_END(nullptr)
}

AMGX_RC AMGXB200_solver_get_level_R(AMGX_solver_handle slv, int level, int *nnz, int *row_ptrs, int *col_indices, void *values)
{
    API_BEGIN
    AMGLevel &L = get_level(slv, level);
    copy_out_matrix(L.R.get(), nnz, row_ptrs, col_indices, values);
    API_END(nullptr)
}

AMGX_RC AMGXB200_solver_get_level_cf_map(AMGX_solver_handle slv, int level, int *cf_map)
{
    API_BEGIN
    AMGLevel &L = get_level(slv, level);
    if (L.cf_map.size() == 0) fatal(AMGX_RC_BAD_PARAMETERS, "level has no C/F map");
    AMGXB_CUDA_CHECK(cudaMemcpy(cf_map, L.cf_map.ptr(), sizeof(int) * L.A->n, cudaMemcpyDeviceToHost));
    API_END(nullptr)
}

AMGX_RC AMGXB200_solver_get_level_smoother_data(AMGX_solver_handle slv, int level, void *data)
{
    API_BEGIN
    AMGLevel &L = get_level(slv, level);
    if (!L.smoother || !L.smoother->smoother_data()) fatal(AMGX_RC_BAD_PARAMETERS, "level has no smoother data");
    const DevVec *d = L.smoother->smoother_data();
    AMGXB_CUDA_CHECK(cudaMemcpy(data, d->ptr(), d->nbytes(), cudaMemcpyDeviceToHost));
    API_END(nullptr)
}

AMGX_RC AMGXB200_solver_get_level_coloring(AMGX_solver_handle slv, int level, int *num_colors, int *row_colors)
{
    API_BEGIN
    AMGLevel &L = get_level(slv, level);
    if (num_colors) *num_colors = L.A->num_colors;
    if (row_colors && L.A->row_colors.size()) AMGXB_CUDA_CHECK(cudaMemcpy(row_colors, L.A->row_colors.ptr(), sizeof(int) * L.A->n, cudaMemcpyDeviceToHost));
    API_END(nullptr)
}

AMGX_RC AMGXB200_solver_get_last_solve_stats(AMGX_solver_handle slv, double *solve_seconds, long long *kernel_launches)
{
    API_BEGIN
    SolverH *h = slvH(slv);
    if (solve_seconds) *solve_seconds = h->last_solve_seconds;
    if (kernel_launches) *kernel_launches = h->last_solve_launches;
    API_END(nullptr)
}

}  // extern "C"
