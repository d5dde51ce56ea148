// dist.h -- row-partitioned multi-GPU layer (one process per GPU): local renumbering
// [interior | boundary | halo], per-neighbour send maps (B2L), NCCL halo exchange overlapped with
// interior work, scalar all-reduces.  Replaces the reference's DistributedManager exchange_halo /
// global_reduce_sum over MPI (include/distributed/distributed_manager.h:955-1170,
// src/distributed/comms_mpi_hostbuffer_stream.cu:598-700, 1226-1268) on the solve path.
// Also hosts the matrix_apply dispatcher and a few block-size helpers used by the solvers.
#pragma once
#include "base.h"
#include "matrix.h"
#include "kernels.h"

namespace amgxb {

struct ScalarBlock;
struct P2PLink;

struct DistManager {
    int rank = 0, world = 1;
    int n_owned = 0, n_interior = 0, n_halo = 0;
    std::vector<int> neighbors;          // ranks
    std::vector<int> send_offsets;       // [nn+1]
    std::vector<int> halo_offsets;       // [nn+1], relative to n_owned
    DevBuf<int> send_maps;               // B2L maps, concatenated
    DevVec send_buf;                     // packed boundary values (vec precision, * block dim)
    DevBuf<int> perm_old_to_new;         // caller (partition) order -> local order, owned rows
    DevBuf<int> caller_row_ptr;          // row pointers in the caller's row order (AMGX_matrix_replace_coefficients on a distributed matrix)
    std::vector<int64_t> halo_global;    // global id of each halo column
    int64_t global_offset = 0;           // first global row owned by this rank
    int64_t n_global = 0;
    cudaEvent_t ev_pack = nullptr, ev_done = nullptr;
    double *allreduce_buf = nullptr;     // device scratch for setup-time collectives
    bool exchange_pending = false;       // a halo exchange is in flight on the side stream
    DevBuf<int> halo_int;                // scratch for integer halo exchanges
    std::shared_ptr<P2PLink> p2p;        // peer-memory link of this manager (p2p.cu); null -> NCCL send/recv
    ~DistManager();
};

// ---- halo exchange (no-ops on a single GPU) ----
void dist_replace_values(Matrix &A, int nnz, const void *data);                     // values in the caller's row order -> local row order
void dist_exchange_halo(const Matrix &A, DevVec &x, cudaStream_t s);
void dist_exchange_halo_ptr(const Matrix &A, void *x, Prec prec, cudaStream_t s);
void dist_exchange_halo_coarse(const Matrix &A, const void *xc, cudaStream_t s);   // vector living on the NEXT level
void dist_wait_halo(const Matrix &A, cudaStream_t s);                              // make `s` wait for the exchange in flight
void dist_exchange_int(const Matrix &A, int *x, cudaStream_t s);                   // blocking (stream-ordered) int exchange, x has n_cols entries
long long dist_allreduce_ll(const Matrix &A, long long v, int op);                 // host value, op: 0 sum, 1 min, 2 max
// replicated coarse tail
std::unique_ptr<Matrix> dist_gather_matrix(const Matrix &A, std::vector<int> &counts, std::vector<int> &offs, bool caller_order = false);
void dist_allreduce_vec(const Matrix &A, void *v, Prec prec, size_t n, cudaStream_t s);   // in-place sum over the ranks, on stream s
void dist_allgatherv_int_inplace(const Matrix &A, int *v, const std::vector<int> &counts, const std::vector<int> &offs, cudaStream_t s);
void dist_allgatherv_inplace(const Matrix &A, void *v, Prec prec, int bsize, const std::vector<int> &counts, const std::vector<int> &offs, cudaStream_t s);
std::shared_ptr<DistManager> dist_coarsen(const Matrix &A, DevBuf<int> &aggregates, int n_agg, int *n_interior_c);
void dist_allreduce_norm(const Matrix &A, const ReduceCtx &red, int slot, int norm_type, cudaStream_t s);
double dist_reduce_norm(const Matrix &A, double local, int norm_type);
void dist_allreduce_host(const Matrix &A, double *vals, int count, int op);      // op 0 sum, 2 max             // host value in, global value out
ReduceCtx dist_wrap_reduce(const Matrix &A, const ReduceCtx &red);
void dist_allreduce_scalar_fin(const Matrix &A, const ReduceCtx &red, int slot, int fin_op, cudaStream_t s);

// y = op(A, x): scalar matrices go to the CSR tile kernels, 4x4 blocks to the block kernels.
void matrix_apply(const Matrix &A, CsrEpi epi, const CsrOpArgs &args, cudaStream_t s);

// ---- block-size > 1 helpers (k_block.cu) ----
void block_norms(const DevVec &v, int n, int bsize, int norm_type, const ReduceCtx &red, ScalarBlock &sb, std::vector<double> &out, cudaStream_t s,
                 const Matrix *dist_of = nullptr);   // dist_of: reduce over the ranks of this matrix
void block_jacobi_setup(const Matrix &A, DevVec &dinv, cudaStream_t s);   // dinv <- inverse of diagonal blocks (in place)
void block_jacobi_zero(const Matrix &A, const DevVec &dinv, const DevVec &b, void *x, double omega, cudaStream_t s);
void block_jacobi_sweep(const Matrix &A, const DevVec &dinv, const DevVec &b, const void *x, void *xout, double omega, cudaStream_t s);
void l1_row_norms(const Matrix &A, DevVec &d, cudaStream_t s);

double device_mem_used_gb();
void block_build_diag(Matrix &A, cudaStream_t s);
void block_apply(const Matrix &A, CsrEpi epi, const CsrOpArgs &args, cudaStream_t s);
void dist_destroy_comm(Resources *rsc);
void set_print_callback(AMGX_print_callback cb);

class Solver;
class Config;
std::unique_ptr<Solver> make_dilu_solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc);

void release_reduce_scratch(Resources *rsc);

}  // namespace amgxb
