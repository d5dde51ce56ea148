// dist.cu -- multi-GPU layer.  (single-GPU behaviour: every exchange is a no-op)
#include "solvers.h"
#include "dist.h"
#include "capi_internal.h"
namespace amgxb {
DistManager::~DistManager()
{
    if (ev_pack) cudaEventDestroy(ev_pack);
    if (ev_done) cudaEventDestroy(ev_done);
    if (allreduce_buf) cudaFree(allreduce_buf);
}
void dist_destroy_comm(Resources *) {}
void dist_exchange_halo(const Matrix &A, DevVec &, cudaStream_t) { if (A.dist) fatal(AMGX_RC_NOT_IMPLEMENTED, "distributed halo exchange"); }
void dist_exchange_halo_coarse(const Matrix &A, const void *, cudaStream_t) { if (A.dist) fatal(AMGX_RC_NOT_IMPLEMENTED, "distributed halo exchange"); }
double dist_reduce_norm(const Matrix &, double local, int) { return local; }
ReduceCtx dist_wrap_reduce(const Matrix &, const ReduceCtx &red) { return red; }
void dist_allreduce_scalar_fin(const Matrix &, const ReduceCtx &, int, int, cudaStream_t) {}
}
namespace amgxb {
void dist_get_unique_id(char *) { fatal(AMGX_RC_NOT_IMPLEMENTED, "NCCL bootstrap"); }
void dist_init_comm(Resources *, const AMGXB200_comm *) { fatal(AMGX_RC_NOT_IMPLEMENTED, "multi-GPU resources"); }
void dist_upload_local(Matrix &, int, int, int, int, const int *, const int *, const void *, const void *) { fatal(AMGX_RC_NOT_IMPLEMENTED, "distributed upload"); }
void dist_prepare_vector(const Matrix &A, Vector &v) { if (A.dist) fatal(AMGX_RC_NOT_IMPLEMENTED, "distributed vectors"); (void)v; }
void dist_download_vector(const Vector &, void *) { fatal(AMGX_RC_NOT_IMPLEMENTED, "distributed vectors"); }
void dist_generate_poisson7(Matrix &, int, int, int, int, int, int) { fatal(AMGX_RC_NOT_IMPLEMENTED, "distributed generator"); }
void dist_upload_global(Matrix &, int, int, int, int, int, const int *, const void *, bool, const void *, const void *, int, const void *) { fatal(AMGX_RC_NOT_IMPLEMENTED, "distributed upload"); }
void dist_comm_from_maps_one_ring(Matrix &, int, const int *, const int *, const int **, const int *, const int **) { fatal(AMGX_RC_NOT_IMPLEMENTED, "comm_from_maps"); }
void partition_plan_create(AMGXB200_partition_plan *, int, int, const int64_t *, int, int, const int *, const int64_t *) { fatal(AMGX_RC_NOT_IMPLEMENTED, "partition planner"); }
void attach_user_coloring(Matrix &, const int *, int, int) { fatal(AMGX_RC_NOT_IMPLEMENTED, "attach_coloring"); }
}
