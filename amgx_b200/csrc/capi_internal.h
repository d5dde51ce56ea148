// capi_internal.h -- handle objects behind the opaque AMGX_*_handle pointers.
#pragma once
#include "solvers.h"
#include "dist.h"

namespace amgxb {

enum : unsigned { MAGIC_CFG = 0xC0F16001u, MAGIC_RSC = 0xC0F16002u, MAGIC_MTX = 0xC0F16003u, MAGIC_VEC = 0xC0F16004u,
                  MAGIC_SLV = 0xC0F16005u, MAGIC_DST = 0xC0F16006u };

struct ConfigH { unsigned magic = MAGIC_CFG; std::shared_ptr<Config> cfg; };
struct ResourcesH { unsigned magic = MAGIC_RSC; std::shared_ptr<Resources> rsc; };
struct MatrixH { unsigned magic = MAGIC_MTX; std::shared_ptr<Matrix> m; };
struct VectorH { unsigned magic = MAGIC_VEC; std::shared_ptr<Vector> v; };
struct SolverH {
    unsigned magic = MAGIC_SLV;
    std::shared_ptr<Resources> rsc;
    int mode = 0;
    std::shared_ptr<Config> cfg;
    std::unique_ptr<Solver> solver;
    std::shared_ptr<Matrix> A;
    Status last_status = ST_FAILED;
    bool was_setup = false;
    double last_solve_seconds = 0;
    long long last_solve_launches = 0;
};
struct DistributionH {
    unsigned magic = MAGIC_DST;
    int info = AMGX_DIST_PARTITION_OFFSETS;
    const void *partition_data = nullptr;
    int use32bit = 0;
};

void upload_matrix(Matrix &A, int n, int nnz, int bx, int by, const int *row_ptrs, const int *col_indices, const void *data, const void *diag_data);
void attach_user_coloring(Matrix &A, const int *row_coloring, int num_rows, int num_colors);   // coloring.cu
void residual_norm_external(SolverH &h, Matrix &A, Vector &b, Vector &x, std::vector<double> &nrm);   // capi2.cu

// distributed plumbing (dist.cu)
void dist_get_unique_id(char *id128);
void dist_init_comm(Resources *rsc, const AMGXB200_comm *comm);
void dist_upload_local(Matrix &A, int n, int nnz, int bx, int by, const int *row_ptrs, const int *col_indices, const void *data, const void *diag_data);
void dist_prepare_vector(const Matrix &A, Vector &v);     // caller order -> local order (+ halo tail) when A is distributed
void dist_download_vector(const Vector &v, void *data);   // local order -> caller order
void dist_generate_poisson7(Matrix &A, int nx, int ny, int nz, int px, int py, int pz);
void dist_upload_global(Matrix &A, int n_global, int n, int nnz, int bx, int by, const int *row_ptrs, const void *cols_global, bool cols32,
                        const void *data, const void *diag_data, int partition_info, const void *partition_data);
void dist_build_matrix(Matrix &A, const int64_t *offsets, int n, int nnz, int bx, int by, const int *rp, const int64_t *cols, const void *vals,
                       const void *diag);
void dist_comm_from_maps_one_ring(Matrix &A, int num_neighbors, const int *neighbors, const int *send_sizes, const int **send_maps,
                                  const int *recv_sizes, const int **recv_maps);
// pure host partition planner (partition.cpp)
std::string comm_maps_to_global_cols(int n, int nnz, const int *local_cols, int64_t my_offset, int num_neighbors, const int *recv_sizes,
                                     const int *const *recv_maps, const int64_t *const *recv_global, int64_t *cols_out);   // partition.cpp
bool partition_vector_to_contiguous(int n_global, int world, const int *pv, int64_t *offsets, int64_t *new_global);   // partition.cpp
void partition_plan_create(AMGXB200_partition_plan *plan, int rank, int world, const int64_t *offsets, int n, int nnz, const int *row_ptrs,
                           const int64_t *cols_global);

}  // namespace amgxb
