// capi_dist.cu -- distributed entry points of the C-ABI and the 7-point Poisson generator.
#include "capi_internal.h"
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
#define API3_BEGIN try {
#define API3_END } catch (...) { return on_exception(__func__); } return AMGX_RC_OK;

// 7-point Poisson on an nx*ny*nz box, natural ordering row = i + nx*j + nx*ny*k; per row the
// diagonal (6) first, then -1 for i-1, i+1, j-1, j+1, k-1, k+1 when inside -- the entry order of
// poisson7pt_set_col_values (src/distributed/distributed_manager.cu:86-260).
__global__ void poisson7_count_kernel(int nx, int ny, int nz, int *row_ptr)
{
    const long long n = (long long)nx * ny * nz;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(r % nx), j = (int)((r / nx) % ny), k = (int)(r / ((long long)nx * ny));
        row_ptr[r] = 1 + (i > 0) + (i < nx - 1) + (j > 0) + (j < ny - 1) + (k > 0) + (k < nz - 1);
    }
}
template <class MatT> __global__ void poisson7_fill_kernel(int nx, int ny, int nz, const int *row_ptr, int *col, MatT *val)
{
    const long long n = (long long)nx * ny * nz;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(r % nx), j = (int)((r / nx) % ny), k = (int)(r / ((long long)nx * ny));
        int p = row_ptr[r];
        col[p] = (int)r; val[p++] = (MatT)6;
        if (i > 0) { col[p] = (int)(r - 1); val[p++] = (MatT)-1; }
        if (i < nx - 1) { col[p] = (int)(r + 1); val[p++] = (MatT)-1; }
        if (j > 0) { col[p] = (int)(r - nx); val[p++] = (MatT)-1; }
        if (j < ny - 1) { col[p] = (int)(r + nx); val[p++] = (MatT)-1; }
        if (k > 0) { col[p] = (int)(r - (long long)nx * ny); val[p++] = (MatT)-1; }
        if (k < nz - 1) { col[p] = (int)(r + (long long)nx * ny); val[p++] = (MatT)-1; }
    }
}
}  // namespace amgxb

#include <cub/cub.cuh>

extern "C" {

AMGX_RC AMGX_generate_distributed_poisson_7pt(AMGX_matrix_handle mtx, AMGX_vector_handle rhs, AMGX_vector_handle sol, int allocated_halo_depth,
                                              int num_import_rings, int nx, int ny, int nz, int px, int py, int pz)
{
    API3_BEGIN
    (void)allocated_halo_depth; (void)num_import_rings;
    MatrixH *m = chk<MatrixH>(mtx, MAGIC_MTX, "matrix");
    Matrix &A = *m->m;
    AMGXB_CUDA_CHECK(cudaSetDevice(A.rsc->device));
    if (nx < 1 || ny < 1 || nz < 1 || px < 1 || py < 1 || pz < 1) fatal(AMGX_RC_BAD_PARAMETERS, "bad grid sizes");
    if (px * py * pz != A.rsc->world) fatal(AMGX_RC_BAD_PARAMETERS, "px*py*pz must equal the number of ranks");
    if (A.rsc->world > 1) {
        dist_generate_poisson7(A, nx, ny, nz, px, py, pz);
    } else {
        const long long n = (long long)nx * ny * nz;
        if (n >= (1ll << 31) / 7) fatal(AMGX_RC_BAD_PARAMETERS, "grid too large for 32-bit indices");
        cudaStream_t s = A.stream();
        A.initialized = false;
        A.n = A.n_cols = (int)n;
        A.bx = A.by = 1;
        A.has_ext_diag = A.merged_ext_diag = false;
        A.dist.reset();
        A.row_ptr.resize(n + 1);
        A.row_ptr.zero(s);
        const int grid = (int)std::min<long long>((n + 255) / 256, B200_SMS * 16);
        poisson7_count_kernel<<<grid, 256, 0, s>>>(nx, ny, nz, A.row_ptr.ptr());
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, A.row_ptr.ptr(), A.row_ptr.ptr(), (int)n + 1, s);
        DevBytes tmp;
        tmp.resize(tb);
        cub::DeviceScan::ExclusiveSum(tmp.p, tb, A.row_ptr.ptr(), A.row_ptr.ptr(), (int)n + 1, s);
        int nnz = 0;
        AMGXB_CUDA_CHECK(cudaMemcpyAsync(&nnz, A.row_ptr.ptr() + n, sizeof(int), cudaMemcpyDeviceToHost, s));
        AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
        A.nnz = nnz;
        A.col_idx.resize(nnz);
        A.values.resize(nnz, A.mat_prec);
        if (A.mat_prec == Prec::F64) poisson7_fill_kernel<double><<<grid, 256, 0, s>>>(nx, ny, nz, A.row_ptr.ptr(), A.col_idx.ptr(), A.values.as<double>());
        else poisson7_fill_kernel<float><<<grid, 256, 0, s>>>(nx, ny, nz, A.row_ptr.ptr(), A.col_idx.ptr(), A.values.as<float>());
        AMGXB_LAUNCH_CHECK();
        A.compute_diag_and_plan();
        AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    }
    // rhs and sol are filled with ones (src/amgx_c.cu:1731-1733)
    auto ones = [&](AMGX_vector_handle vh) {
        if (!vh) return;
        VectorH *v = chk<VectorH>(vh, MAGIC_VEC, "vector");
        v->v->n = A.n;
        v->v->block_dim = 1;
        v->v->dist = A.dist;
        v->v->data.resize((size_t)A.n_cols, v->v->prec);
        v->v->data.zero(A.stream());
        vec_fill(v->v->data.ptr(), v->v->prec, (size_t)A.n, 1.0, A.stream());
        v->v->user_order = !A.dist;
    };
    ones(rhs);
    ones(sol);
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(A.stream()));
    API3_END
}

AMGX_RC AMGX_distribution_create(AMGX_distribution_handle *dist, AMGX_config_handle)
{
    API3_BEGIN
    if (!dist) fatal(AMGX_RC_BAD_PARAMETERS, "null handle pointer");
    *dist = reinterpret_cast<AMGX_distribution_handle>(new DistributionH);
    API3_END
}
AMGX_RC AMGX_distribution_destroy(AMGX_distribution_handle dist)
{
    API3_BEGIN
    DistributionH *h = chk<DistributionH>(dist, MAGIC_DST, "distribution");
    h->magic = 0;
    delete h;
    API3_END
}
AMGX_RC AMGX_distribution_set_partition_data(AMGX_distribution_handle dist, AMGX_DIST_PARTITION_INFO info, const void *partition_data)
{
    API3_BEGIN
    DistributionH *h = chk<DistributionH>(dist, MAGIC_DST, "distribution");
    h->info = (int)info;
    h->partition_data = partition_data;
    API3_END
}
AMGX_RC AMGX_distribution_set_32bit_colindices(AMGX_distribution_handle dist, int use32bit)
{
    API3_BEGIN
    DistributionH *h = chk<DistributionH>(dist, MAGIC_DST, "distribution");
    h->use32bit = use32bit;
    API3_END
}

AMGX_RC AMGX_matrix_upload_distributed(AMGX_matrix_handle mtx, int n_global, int n, int nnz, int block_dimx, int block_dimy, const int *row_ptrs,
                                       const void *col_indices_global, const void *data, const void *diag_data, AMGX_distribution_handle distribution)
{
    API3_BEGIN
    MatrixH *m = chk<MatrixH>(mtx, MAGIC_MTX, "matrix");
    DistributionH *d = chk<DistributionH>(distribution, MAGIC_DST, "distribution");
    AMGXB_CUDA_CHECK(cudaSetDevice(m->m->rsc->device));
    dist_upload_global(*m->m, n_global, n, nnz, block_dimx, block_dimy, row_ptrs, col_indices_global, d->use32bit != 0, data, diag_data,
                       d->info, d->partition_data);
    API3_END
}

AMGX_RC AMGX_matrix_upload_all_global(AMGX_matrix_handle mtx, int n_global, int n, int nnz, int block_dimx, int block_dimy, const int *row_ptrs,
                                      const void *col_indices_global, const void *data, const void *diag_data, int, int, const int *partition_vector)
{
    API3_BEGIN
    MatrixH *m = chk<MatrixH>(mtx, MAGIC_MTX, "matrix");
    AMGXB_CUDA_CHECK(cudaSetDevice(m->m->rsc->device));
    dist_upload_global(*m->m, n_global, n, nnz, block_dimx, block_dimy, row_ptrs, col_indices_global, false, data, diag_data,
                       AMGX_DIST_PARTITION_VECTOR, partition_vector);
    API3_END
}

AMGX_RC AMGX_matrix_upload_all_global_32(AMGX_matrix_handle mtx, int n_global, int n, int nnz, int block_dimx, int block_dimy, const int *row_ptrs,
                                         const void *col_indices_global, const void *data, const void *diag_data, int, int, const int *partition_vector)
{
    API3_BEGIN
    MatrixH *m = chk<MatrixH>(mtx, MAGIC_MTX, "matrix");
    AMGXB_CUDA_CHECK(cudaSetDevice(m->m->rsc->device));
    dist_upload_global(*m->m, n_global, n, nnz, block_dimx, block_dimy, row_ptrs, col_indices_global, true, data, diag_data,
                       AMGX_DIST_PARTITION_VECTOR, partition_vector);
    API3_END
}

// include/amgx_c.h:310-323: the multi-ring form, neighbour q's maps are send_maps[send_ptrs[q] .. send_ptrs[q+1]) etc.  The engine
// exchanges one ring (what aggregation AMG needs); more import rings are rejected, not ignored.
AMGX_RC AMGX_matrix_comm_from_maps(AMGX_matrix_handle mtx, int allocated_halo_depth, int num_import_rings, int max_num_neighbors, const int *neighbors,
                                   const int *send_ptrs, const int *send_maps, const int *recv_ptrs, const int *recv_maps)
{
    API3_BEGIN
    MatrixH *m = chk<MatrixH>(mtx, MAGIC_MTX, "matrix");
    // the reference's own limits (src/amgx_c.cu:1871-1887)
    if (allocated_halo_depth > 1) fatal(AMGX_RC_BAD_PARAMETERS, "Allocated_halo_depth > 1 currently not supported");
    if (num_import_rings > 1) fatal(AMGX_RC_BAD_PARAMETERS, "num_import_rings > 1 currently not supported");
    if (allocated_halo_depth != num_import_rings) fatal(AMGX_RC_BAD_PARAMETERS, "num_import_rings != allocated_halo_depth currently not supported");
    if (max_num_neighbors < 0 || (max_num_neighbors > 0 && (!neighbors || !send_ptrs || !send_maps || !recv_ptrs || !recv_maps)))
        fatal(AMGX_RC_BAD_PARAMETERS, "AMGX_matrix_comm_from_maps: null map arrays");
    std::vector<int> ssz(max_num_neighbors), rsz(max_num_neighbors);
    std::vector<const int *> sm(max_num_neighbors), rm(max_num_neighbors);
    for (int q = 0; q < max_num_neighbors; q++) {
        ssz[q] = send_ptrs[q + 1] - send_ptrs[q];
        rsz[q] = recv_ptrs[q + 1] - recv_ptrs[q];
        sm[q] = send_maps + send_ptrs[q];
        rm[q] = recv_maps + recv_ptrs[q];
    }
    dist_comm_from_maps_one_ring(*m->m, max_num_neighbors, neighbors, ssz.data(), sm.data(), rsz.data(), rm.data());
    API3_END
}

AMGX_RC AMGX_matrix_comm_from_maps_one_ring(AMGX_matrix_handle mtx, int allocated_halo_depth, int num_neighbors, const int *neighbors,
                                            const int *send_sizes, const int **send_maps, const int *recv_sizes, const int **recv_maps)
{
    API3_BEGIN
    MatrixH *m = chk<MatrixH>(mtx, MAGIC_MTX, "matrix");
    (void)allocated_halo_depth;
    dist_comm_from_maps_one_ring(*m->m, num_neighbors, neighbors, send_sizes, send_maps, recv_sizes, recv_maps);
    API3_END
}

AMGX_RC AMGXB200_partition_plan_create(AMGXB200_partition_plan *plan, int rank, int world_size, const int64_t *offsets, int n, int nnz,
                                       const int *row_ptrs, const int64_t *col_indices_global)
{
    API3_BEGIN
    if (!plan) fatal(AMGX_RC_BAD_PARAMETERS, "null plan");
    partition_plan_create(plan, rank, world_size, offsets, n, nnz, row_ptrs, col_indices_global);
    API3_END
}

AMGX_RC AMGXB200_comm_maps_to_global_cols(int n, int nnz, const int *local_cols, int64_t my_offset, int num_neighbors, const int *recv_sizes,
                                          const int *const *recv_maps, const int64_t *const *recv_global, int64_t *cols_out)
{
    API3_BEGIN
    if (n < 0 || nnz < 0 || (nnz > 0 && (!local_cols || !cols_out)) || num_neighbors < 0) fatal(AMGX_RC_BAD_PARAMETERS, "comm_maps_to_global_cols: bad arguments");
    const std::string err = comm_maps_to_global_cols(n, nnz, local_cols, my_offset, num_neighbors, recv_sizes, recv_maps, recv_global, cols_out);
    if (!err.empty()) fatal(AMGX_RC_BAD_PARAMETERS, err);
    API3_END
}

AMGX_RC AMGXB200_partition_vector_to_contiguous(int n_global, int world_size, const int *partition_vector, int64_t *offsets, int64_t *new_global)
{
    API3_BEGIN
    if (!partition_vector || !offsets || n_global < 0 || world_size < 1) fatal(AMGX_RC_BAD_PARAMETERS, "partition_vector_to_contiguous: bad arguments");
    if (!partition_vector_to_contiguous(n_global, world_size, partition_vector, offsets, new_global))
        fatal(AMGX_RC_BAD_PARAMETERS, "partition vector names a rank outside [0, world_size)");
    API3_END
}

void AMGXB200_partition_plan_free(AMGXB200_partition_plan *plan)
{
    if (!plan) return;
    free(plan->neighbors); free(plan->send_offsets); free(plan->send_maps); free(plan->halo_offsets);
    free(plan->halo_global); free(plan->perm_old_to_new); free(plan->local_cols);
    memset(plan, 0, sizeof(*plan));
}

}  // extern "C"
