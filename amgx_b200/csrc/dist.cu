// dist.cu -- multi-GPU layer: NCCL bootstrap, distributed matrix construction (from the host partition
// plan), halo exchange overlapped with interior rows, scalar all-reduces, distributed vectors.
// See dist.h for what it replaces in the reference.  Single-GPU matrices (A.dist == null) make every
// entry point here a no-op.
#include "solvers.h"
#include "dist.h"
#include <cmath>
#include "capi_internal.h"
#include "p2p.h"
#include <nccl.h>
#include <algorithm>
#include <numeric>

namespace amgxb {

#define AMGXB_NCCL_CHECK(expr)                                                                    \
    do {                                                                                          \
        ncclResult_t _r = (expr);                                                                 \
        if (_r != ncclSuccess) {                                                                  \
            char _b[512];                                                                         \
            snprintf(_b, sizeof(_b), "NCCL error %s at %s:%d", ncclGetErrorString(_r), __FILE__, __LINE__); \
            throw ::amgxb::Error(AMGX_RC_CORE, _b);                                               \
        }                                                                                         \
    } while (0)

DistManager::~DistManager()
{
    if (ev_pack) cudaEventDestroy(ev_pack);
    if (ev_done) cudaEventDestroy(ev_done);
    if (allreduce_buf) cudaFree(allreduce_buf);
}

void dist_get_unique_id(char *id128)
{
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    AMGXB_NCCL_CHECK(ncclGetUniqueId(&id));
    memcpy(id128, &id, 128);
}

void dist_init_comm(Resources *rsc, const AMGXB200_comm *comm)
{
    ncclUniqueId id;
    memcpy(&id, comm->nccl_unique_id, 128);
    ncclComm_t c;
    AMGXB_NCCL_CHECK(ncclCommInitRank(&c, comm->world_size, id, comm->rank));
    rsc->nccl_comm = c;
    rsc->rank = comm->rank;
    rsc->world = comm->world_size;
    p2p_init(rsc);      // NVLink peer-memory windows (CUDA IPC); falls back to NCCL send/recv when unavailable
}

void dist_destroy_comm(Resources *rsc)
{
    p2p_shutdown(rsc);
    if (rsc->nccl_comm) {
        ncclCommDestroy((ncclComm_t)rsc->nccl_comm);
        rsc->nccl_comm = nullptr;
    }
}

static ncclComm_t comm_of(const Matrix &A) { return (ncclComm_t)A.rsc->nccl_comm; }

// ---------------------------------------------------------------------------------------------
// small host-visible collectives used at setup time
// ---------------------------------------------------------------------------------------------
static void ensure_scratch(DistManager &m)
{
    if (!m.allreduce_buf) {
        AMGXB_CUDA_CHECK(cudaMalloc(&m.allreduce_buf, 64 * sizeof(double)));
        AMGXB_CUDA_CHECK(cudaEventCreateWithFlags(&m.ev_pack, cudaEventDisableTiming));
        AMGXB_CUDA_CHECK(cudaEventCreateWithFlags(&m.ev_done, cudaEventDisableTiming));
    }
}

long long dist_allreduce_ll(const Matrix &A, long long v, int op /*0 sum, 1 min, 2 max*/)
{
    if (!A.dist) return v;
    DistManager &m = *A.dist;
    ensure_scratch(m);
    cudaStream_t s = A.stream();
    long long *buf = reinterpret_cast<long long *>(m.allreduce_buf);
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(buf, &v, sizeof(v), cudaMemcpyHostToDevice, s));
    AMGXB_NCCL_CHECK(ncclAllReduce(buf, buf, 1, ncclInt64, op == 0 ? ncclSum : op == 1 ? ncclMin : ncclMax, comm_of(A), s));
    long long out;
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(&out, buf, sizeof(out), cudaMemcpyDeviceToHost, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    return out;
}

// ---------------------------------------------------------------------------------------------
// halo exchange
// ---------------------------------------------------------------------------------------------
namespace {
template <class T> __global__ void pack_kernel(const int *__restrict__ map, int count, int bsize, const T *__restrict__ x, T *__restrict__ buf)
{
    const long long total = (long long)count * bsize;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t / bsize), c = (int)(t % bsize);
        buf[t] = x[(size_t)map[k] * bsize + c];
    }
}
__global__ void sqrt_mirror_kernel(double *scal, int slot, int do_sqrt, double *host_mirror)
{
    double v = scal[slot];
    if (do_sqrt) v = sqrt(v);
    scal[slot] = v;
    if (host_mirror) host_mirror[slot] = v;
    __threadfence_system();
}
__global__ void fin_kernel(double *scal, int slot, int fin_op)
{
    const double sum = scal[slot];
    if (fin_op == FIN_PCG_ALPHA) {
        scal[S_DOT] = sum;
        const double a = (sum != 0.0) ? scal[S_RZ] / sum : 0.0;
        scal[S_ALPHA] = a;
        scal[S_NEG_ALPHA] = -a;
    } else if (fin_op == FIN_PCG_BETA) {
        const double old = scal[S_RZ];
        scal[S_RZ_OLD] = old;
        scal[S_RZ] = sum;
        scal[S_BETA] = (old != 0.0) ? sum / old : 0.0;
    } else if (fin_op == FIN_SQRT) {
        scal[slot] = sqrt(sum);
    }
}
}  // namespace

// Start the exchange of x's halo: pack boundary values (compute stream), then send/recv on the side
// stream.  matrix_apply waits for it only before the rows that read halo columns.
template <class T> static void exchange_typed(const Matrix &A, T *x, int bsize, cudaStream_t s, ncclDataType_t dt)
{
    DistManager &m = *A.dist;
    ensure_scratch(m);
    const int nn = (int)m.neighbors.size();
    if (nn == 0) return;
    const int total_send = m.send_offsets[nn];
    const size_t need = (size_t)total_send * bsize * sizeof(T);
    if (m.send_buf.b.bytes < need) m.send_buf.b.resize(need);
    T *buf = (T *)m.send_buf.b.p;
    if (total_send) {
        const int grid = std::min(ceil_div((long long)total_send * bsize, 256), B200_SMS * 8);
        pack_kernel<T><<<grid, 256, 0, s>>>(m.send_maps.ptr(), total_send, bsize, x, buf);
        count_launch();
        AMGXB_LAUNCH_CHECK();
    }
    cudaStream_t side = A.rsc->side_stream;
    AMGXB_CUDA_CHECK(cudaEventRecord(m.ev_pack, s));
    AMGXB_CUDA_CHECK(cudaStreamWaitEvent(side, m.ev_pack, 0));
    AMGXB_NCCL_CHECK(ncclGroupStart());
    for (int q = 0; q < nn; q++) {
        const int sc = m.send_offsets[q + 1] - m.send_offsets[q], rc = m.halo_offsets[q + 1] - m.halo_offsets[q];
        if (sc) AMGXB_NCCL_CHECK(ncclSend(buf + (size_t)m.send_offsets[q] * bsize, (size_t)sc * bsize, dt, m.neighbors[q], comm_of(A), side));
        if (rc) AMGXB_NCCL_CHECK(ncclRecv(x + ((size_t)m.n_owned + m.halo_offsets[q]) * bsize, (size_t)rc * bsize, dt, m.neighbors[q], comm_of(A), side));
    }
    AMGXB_NCCL_CHECK(ncclGroupEnd());
    AMGXB_CUDA_CHECK(cudaEventRecord(m.ev_done, side));
    m.exchange_pending = true;
}

void dist_exchange_halo_ptr(const Matrix &A, void *x, Prec prec, cudaStream_t s)
{
    if (!A.dist) return;
    if (A.dist->exchange_pending) dist_wait_halo(A, s);
    const int bsize = A.bx;
    // Peer-memory path: the whole exchange is ONE kernel (p2p.cu).  Matrices applied as a whole (small levels, block matrices) run it on
    // the compute stream; split matrices run it on the side stream while the interior rows are processed, and the boundary rows wait
    // for its event -- the structure of the NCCL path without the pack kernel and the NCCL launch.
    if (A.dist->p2p) {
        DistManager &m = *A.dist;
        if (A.plan.split == 0) {
            if (p2p_exchange_blocking(A, x, prec, bsize, s)) { m.exchange_pending = false; return; }
        } else if (!m.neighbors.empty()) {
            ensure_scratch(m);
            cudaStream_t side = A.rsc->side_stream;
            AMGXB_CUDA_CHECK(cudaEventRecord(m.ev_pack, s));
            AMGXB_CUDA_CHECK(cudaStreamWaitEvent(side, m.ev_pack, 0));
            if (p2p_exchange_blocking(A, x, prec, bsize, side)) {
                AMGXB_CUDA_CHECK(cudaEventRecord(m.ev_done, side));
                m.exchange_pending = true;
                return;
            }
        } else return;
    }
    if (prec == Prec::F64) exchange_typed<double>(A, (double *)x, bsize, s, ncclDouble);
    else exchange_typed<float>(A, (float *)x, bsize, s, ncclFloat);
}

void dist_exchange_halo(const Matrix &A, DevVec &x, cudaStream_t s) { dist_exchange_halo_ptr(A, x.ptr(), x.prec, s); }

void dist_wait_halo(const Matrix &A, cudaStream_t s)
{
    if (!A.dist || !A.dist->exchange_pending) return;
    AMGXB_CUDA_CHECK(cudaStreamWaitEvent(s, A.dist->ev_done, 0));
    A.dist->exchange_pending = false;
}

void dist_exchange_int(const Matrix &A, int *x, cudaStream_t s)
{
    if (!A.dist) return;
    exchange_typed<int>(A, x, 1, s, ncclInt32);
    dist_wait_halo(A, s);
}

void dist_exchange_halo_coarse(const Matrix &, const void *, cudaStream_t) {}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
ReduceCtx dist_wrap_reduce(const Matrix &, const ReduceCtx &red) { return red; }

// scal[slot] holds this rank's partial sum: all-reduce it in place on the compute stream, then apply the
// scalar epilogue on the device.
void dist_allreduce_scalar_fin(const Matrix &A, const ReduceCtx &red, int slot, int fin_op, cudaStream_t s)
{
    if (!A.dist) return;
    if (p2p_allreduce_scalar(A, red, slot, 0, 0, fin_op, 0, false, s)) return;
    AMGXB_NCCL_CHECK(ncclAllReduce(red.scal + slot, red.scal + slot, 1, ncclDouble, ncclSum, comm_of(A), s));
    if (fin_op != FIN_STORE) {
        fin_kernel<<<1, 1, 0, s>>>(red.scal, slot, fin_op);
        count_launch();
        AMGXB_LAUNCH_CHECK();
    }
}

// norm of a distributed vector: scal[slot] holds the local sum (L1), sum of squares (L2) or max (LMAX)
void dist_allreduce_norm(const Matrix &A, const ReduceCtx &red, int slot, int norm_type, cudaStream_t s)
{
    if (!A.dist) return;
    if (p2p_allreduce_scalar(A, red, slot, norm_type == 2 ? 2 : 0, 1, FIN_STORE, norm_type == 1, true, s)) return;
    AMGXB_NCCL_CHECK(ncclAllReduce(red.scal + slot, red.scal + slot, 1, ncclDouble, norm_type == 2 ? ncclMax : ncclSum, comm_of(A), s));
    sqrt_mirror_kernel<<<1, 1, 0, s>>>(red.scal, slot, norm_type == 1, red.host_mirror);
    count_launch();
    AMGXB_LAUNCH_CHECK();
}

double dist_reduce_norm(const Matrix &, double local, int) { return local; }

// host values in, globally reduced values out (op 0: sum, 2: max); used by the per-component block norms
void dist_allreduce_host(const Matrix &A, double *vals, int count, int op)
{
    if (!A.dist || count <= 0) return;
    DistManager &m = *A.dist;
    ensure_scratch(m);
    if (count > 32) fatal(AMGX_RC_INTERNAL, "dist_allreduce_host: at most 32 values");
    cudaStream_t s = A.stream();
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(m.allreduce_buf, vals, sizeof(double) * count, cudaMemcpyHostToDevice, s));
    AMGXB_NCCL_CHECK(ncclAllReduce(m.allreduce_buf, m.allreduce_buf, count, ncclDouble, op == 2 ? ncclMax : ncclSum, comm_of(A), s));
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(vals, m.allreduce_buf, sizeof(double) * count, cudaMemcpyDeviceToHost, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
}

// ---------------------------------------------------------------------------------------------
// y = op(A, x) with the halo exchange overlapped: rows [0, split) first, wait, rows [split, n)
// ---------------------------------------------------------------------------------------------
void matrix_apply(const Matrix &A, CsrEpi epi, const CsrOpArgs &args, cudaStream_t s)
{
    if (A.bs() != 1) {
        if (A.dist) dist_wait_halo(A, s);
        block_apply(A, epi, args, s);
        return;
    }
    if (A.has_ext_diag) fatal(AMGX_RC_INTERNAL, "scalar matrix with external diagonal must be merged at upload");
    if (!A.dist) { csr_op(A, epi, args, s, 0); return; }
    if (A.plan.split > 0 && A.dist->exchange_pending) {
        csr_op(A, epi, args, s, 1);
        dist_wait_halo(A, s);
        if (epi == EPI_SPMV_DOT || epi == EPI_JACOBI_DOT || epi == EPI_RESID_NRM2) {
            // the two row segments each finish a partial sum: the second one adds to the first (stream order)
            if (args.fin_op != FIN_STORE) fatal(AMGX_RC_INTERNAL, "distributed fused reduction must use FIN_STORE + all-reduce");
            CsrOpArgs a2 = args;
            a2.fin_op = FIN_ADD;
            csr_op(A, epi, a2, s, 2);
        } else {
            csr_op(A, epi, args, s, 2);
        }
    } else {
        dist_wait_halo(A, s);
        csr_op(A, epi, args, s, 0);
    }
}

// ---------------------------------------------------------------------------------------------
// building a distributed matrix from this rank's rows (global column ids)
// ---------------------------------------------------------------------------------------------
static std::shared_ptr<DistManager> manager_from_plan(Matrix &A, const AMGXB200_partition_plan &pl, const int64_t *offsets)
{
    auto m = std::make_shared<DistManager>();
    m->rank = A.rsc->rank;
    m->world = A.rsc->world;
    m->n_owned = pl.n_owned;
    m->n_interior = pl.n_interior;
    m->n_halo = pl.n_halo;
    m->neighbors.assign(pl.neighbors, pl.neighbors + pl.num_neighbors);
    m->send_offsets.assign(pl.send_offsets, pl.send_offsets + pl.num_neighbors + 1);
    m->halo_offsets.assign(pl.halo_offsets, pl.halo_offsets + pl.num_neighbors + 1);
    m->send_maps.from_any(pl.send_maps, (size_t)m->send_offsets.back(), A.stream());
    m->perm_old_to_new.from_any(pl.perm_old_to_new, (size_t)pl.n_owned, A.stream());
    m->halo_global.assign(pl.halo_global, pl.halo_global + pl.n_halo);
    m->global_offset = offsets[m->rank];
    m->n_global = offsets[m->world];
    return m;
}

// The planner (partition.cpp) derives the send side locally, which is only right when the pattern is structurally symmetric across the
// cut (row i references a column of q <=> q references row i).  The reference accepts any pattern: its B2L maps are built from what the
// RECEIVERS ask for (createOneRingB2Lmaps, src/distributed/distributed_manager.cu).  Do the same: every rank learns how many of its rows
// each other rank needs (all-gather of the need counts), the receivers send the global ids of their halo columns to the owners, and the
// send maps are rebuilt from what arrives.  Neighbour lists become symmetric (a neighbour one only sends to, or only receives from,
// has an empty range on the other side), so every grouped send/recv and every peer-memory flag has its partner.  On a structurally
// symmetric pattern this reproduces the planner's maps exactly (same rows, ascending global id).
static void reconcile_plan(Matrix &A)
{
    DistManager &m = *A.dist;
    const int world = m.world, rank = m.rank;
    cudaStream_t s = A.stream();
    ensure_scratch(m);
    const int nn0 = (int)m.neighbors.size();
    // need[r * world + q] = number of halo columns rank r reads from rank q
    std::vector<int> mine(world, 0);
    for (int q = 0; q < nn0; q++) mine[m.neighbors[q]] = m.halo_offsets[q + 1] - m.halo_offsets[q];
    DevBuf<int> d_mine, d_all;
    d_mine.from_any(mine.data(), (size_t)world, s);
    d_all.resize((size_t)world * world);
    AMGXB_NCCL_CHECK(ncclAllGather(d_mine.ptr(), d_all.ptr(), (size_t)world, ncclInt32, comm_of(A), s));
    const std::vector<int> need = d_all.to_host(s);
    // symmetric neighbour list, ascending rank; halo ranges keep the planner's order (ascending owner)
    std::vector<int> nbrs, halo_off{0}, send_cnt;
    std::vector<int> old_index(world, -1);
    for (int q = 0; q < nn0; q++) old_index[m.neighbors[q]] = q;
    for (int q = 0; q < world; q++) {
        if (q == rank) continue;
        const int i_need = need[(size_t)rank * world + q], q_needs = need[(size_t)q * world + rank];
        if (i_need == 0 && q_needs == 0) continue;
        nbrs.push_back(q);
        halo_off.push_back(halo_off.back() + i_need);
        send_cnt.push_back(q_needs);
    }
    const int nn = (int)nbrs.size();
    if (halo_off.back() != m.n_halo) fatal(AMGX_RC_INTERNAL, "reconcile_plan: halo ranges do not add up");
    // ids I need -> owners; ids the others need from me <- requesters
    std::vector<long long> ask((size_t)std::max(m.n_halo, 1));
    for (int k = 0; k < m.n_halo; k++) ask[k] = (long long)m.halo_global[k];
    std::vector<int> send_off(nn + 1, 0);
    for (int q = 0; q < nn; q++) send_off[q + 1] = send_off[q] + send_cnt[q];
    DevBuf<long long> d_ask, d_req;
    d_ask.from_any(ask.data(), ask.size(), s);
    d_req.resize((size_t)std::max(send_off[nn], 1));
    AMGXB_NCCL_CHECK(ncclGroupStart());
    for (int q = 0; q < nn; q++) {
        const int hc = halo_off[q + 1] - halo_off[q];
        if (hc) AMGXB_NCCL_CHECK(ncclSend(d_ask.ptr() + halo_off[q], (size_t)hc, ncclInt64, nbrs[q], comm_of(A), s));
        if (send_cnt[q]) AMGXB_NCCL_CHECK(ncclRecv(d_req.ptr() + send_off[q], (size_t)send_cnt[q], ncclInt64, nbrs[q], comm_of(A), s));
    }
    AMGXB_NCCL_CHECK(ncclGroupEnd());
    const std::vector<long long> req = d_req.to_host(s);
    const std::vector<int> perm = m.perm_old_to_new.to_host(s);
    std::vector<int> smap((size_t)std::max(send_off[nn], 1));
    for (int k = 0; k < send_off[nn]; k++) {
        const long long g = req[k] - (long long)m.global_offset;
        if (g < 0 || g >= m.n_owned) fatal(AMGX_RC_BAD_PARAMETERS, "distributed matrix: a neighbour asks for a row this rank does not own (inconsistent partition offsets)");
        smap[k] = perm[(size_t)g];
    }
    m.neighbors = nbrs;
    m.halo_offsets = halo_off;
    m.send_offsets = send_off;
    m.send_maps.from_any(smap.data(), (size_t)send_off[nn], s);
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
}

// AMGX_matrix_replace_coefficients on a row-partitioned matrix: the caller's values come in the caller's row order, the engine keeps
// rows as [interior | boundary]; entries keep their order inside a row, so row i's segment moves as a block to row perm[i].
namespace {
template <class T> __global__ void permute_row_values_kernel(int n, const int *__restrict__ old_rp, const int *__restrict__ perm, const int *__restrict__ new_rp,
                                                             int bs, const T *__restrict__ src, T *__restrict__ dst)
{
    const int lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const size_t s0 = (size_t)old_rp[i] * bs, len = (size_t)(old_rp[i + 1] - old_rp[i]) * bs, d0 = (size_t)new_rp[perm[i]] * bs;
        for (size_t k = lane; k < len; k += 32) dst[d0 + k] = src[s0 + k];
    }
}
}  // namespace

void dist_replace_values(Matrix &A, int nnz, const void *data)
{
    DistManager &m = *A.dist;
    if (m.caller_row_ptr.size() != (size_t)A.n + 1) fatal(AMGX_RC_INTERNAL, "distributed matrix without the caller's row pointers");
    cudaStream_t s = A.stream();
    const size_t bs = A.bs(), msz = prec_size(A.mat_prec);
    DevBytes stage;
    stage.resize(std::max<size_t>((size_t)nnz * bs * msz, 1));
    if (nnz) AMGXB_CUDA_CHECK(cudaMemcpyAsync(stage.p, data, (size_t)nnz * bs * msz, cudaMemcpyDefault, s));
    const int grid = std::max(1, std::min(ceil_div(A.n, 8), B200_SMS * 8));
    if (A.mat_prec == Prec::F64)
        permute_row_values_kernel<double><<<grid, 256, 0, s>>>(A.n, m.caller_row_ptr.ptr(), m.perm_old_to_new.ptr(), A.row_ptr.ptr(), (int)bs, (const double *)stage.p,
                                                                A.values.as<double>());
    else
        permute_row_values_kernel<float><<<grid, 256, 0, s>>>(A.n, m.caller_row_ptr.ptr(), m.perm_old_to_new.ptr(), A.row_ptr.ptr(), (int)bs, (const float *)stage.p,
                                                               A.values.as<float>());
    count_launch();
    AMGXB_LAUNCH_CHECK();
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
}

// host arrays: rp[n+1], global cols[nnz], values (mat precision, nnz*bs), optional external diagonal
void dist_build_matrix(Matrix &A, const int64_t *offsets, int n, int nnz, int bx, int by, const int *rp, const int64_t *cols, const void *vals,
                       const void *diag)
{
    if (bx != by) fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "rectangular blocks are not supported");
    if (diag) fatal(AMGX_RC_NOT_IMPLEMENTED, "distributed upload with an external diagonal");
    AMGXB200_partition_plan pl;
    memset(&pl, 0, sizeof(pl));
    partition_plan_create(&pl, A.rsc->rank, A.rsc->world, offsets, n, nnz, rp, cols);
    // permute rows into the local order
    std::vector<int> inv(n);
    for (int i = 0; i < n; i++) inv[pl.perm_old_to_new[i]] = i;
    const size_t bs = (size_t)bx * by, msz = prec_size(A.mat_prec);
    std::vector<int> rp2(n + 1), ci2(std::max(nnz, 1));
    std::vector<char> va2(std::max<size_t>((size_t)nnz * bs * msz, 1));
    size_t o = 0;
    for (int p = 0; p < n; p++) {
        const int i = inv[p];
        rp2[p] = (int)o;
        const int len = rp[i + 1] - rp[i];
        memcpy(&ci2[o], pl.local_cols + rp[i], sizeof(int) * len);
        memcpy(&va2[o * bs * msz], (const char *)vals + (size_t)rp[i] * bs * msz, (size_t)len * bs * msz);
        o += len;
    }
    rp2[n] = (int)o;
    cudaStream_t s = A.stream();
    A.initialized = false;
    A.n = n;
    A.n_cols = n + pl.n_halo;
    A.split_row = pl.n_interior;
    A.nnz = nnz;
    A.bx = bx;
    A.by = by;
    A.has_ext_diag = A.merged_ext_diag = false;
    A.row_ptr.from_any(rp2.data(), n + 1, s);
    A.col_idx.from_any(ci2.data(), nnz, s);
    A.values.resize((size_t)nnz * bs, A.mat_prec);
    if (nnz) AMGXB_CUDA_CHECK(cudaMemcpyAsync(A.values.ptr(), va2.data(), (size_t)nnz * bs * msz, cudaMemcpyHostToDevice, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    A.dist = manager_from_plan(A, pl, offsets);
    A.dist->caller_row_ptr.from_any(rp, (size_t)n + 1, s);
    AMGXB200_partition_plan_free(&pl);
    reconcile_plan(A);      // send side from what the receivers ask for (any pattern, not only structurally symmetric ones)
    p2p_manager_setup(A);
    A.compute_diag_and_plan();
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
}

void dist_upload_global(Matrix &A, int n_global, int n, int nnz, int bx, int by, const int *row_ptrs, const void *cols_global, bool cols32,
                        const void *data, const void *diag_data, int partition_info, const void *partition_data)
{
    const int world = A.rsc->world, rank = A.rsc->rank;
    std::vector<int64_t> offsets(world + 1, 0), new_global;   // new_global: non-contiguous partition vector -> contiguous ids
    if (partition_info == AMGX_DIST_PARTITION_OFFSETS) {
        if (!partition_data) fatal(AMGX_RC_BAD_PARAMETERS, "partition offsets missing");
        for (int r = 0; r <= world; r++) offsets[r] = cols32 ? (int64_t)((const int *)partition_data)[r] : ((const int64_t *)partition_data)[r];
    } else {
        // partition vector: only contiguous, rank-ordered partitions are supported (the layout OFFSETS describes)
        if (!partition_data) {
            // default of the reference: equal contiguous blocks
            for (int r = 0; r <= world; r++) offsets[r] = (int64_t)n_global * r / world;
        } else {
            // any partition vector: rank r's rows are renumbered to [offsets[r], offsets[r+1]) in increasing global id, the
            // reference's ipartition_map (loadDistributedMatrixPartitionVec); callers pass their rows in that order
            const int *pv = (const int *)partition_data;
            new_global.resize((size_t)std::max(n_global, 1));
            if (!partition_vector_to_contiguous(n_global, world, pv, offsets.data(), new_global.data()))
                fatal(AMGX_RC_BAD_PARAMETERS, "partition vector names a rank outside [0, number of ranks)");
            bool identity = true;
            for (int g = 0; g < n_global && identity; g++) identity = (new_global[g] == g);
            if (identity) new_global.clear();
        }
    }
    if (offsets[world] != n_global || offsets[rank + 1] - offsets[rank] != n) fatal(AMGX_RC_BAD_PARAMETERS, "partition does not match n / n_global");
    // bring everything to the host
    const size_t bs = (size_t)bx * by, msz = prec_size(A.mat_prec);
    std::vector<int> rp(n + 1);
    AMGXB_CUDA_CHECK(cudaMemcpy(rp.data(), row_ptrs, sizeof(int) * (n + 1), cudaMemcpyDefault));
    std::vector<int64_t> cols(std::max(nnz, 1));
    if (cols32) {
        std::vector<int> c32(std::max(nnz, 1));
        if (nnz) AMGXB_CUDA_CHECK(cudaMemcpy(c32.data(), cols_global, sizeof(int) * nnz, cudaMemcpyDefault));
        for (int k = 0; k < nnz; k++) cols[k] = c32[k];
    } else if (nnz) AMGXB_CUDA_CHECK(cudaMemcpy(cols.data(), cols_global, sizeof(int64_t) * nnz, cudaMemcpyDefault));
    if (!new_global.empty())
        for (int k = 0; k < nnz; k++) {
            if (cols[k] < 0 || cols[k] >= n_global) fatal(AMGX_RC_BAD_PARAMETERS, "global column index out of range");
            cols[k] = new_global[cols[k]];
        }
    std::vector<char> vals(std::max<size_t>((size_t)nnz * bs * msz, 1));
    if (nnz) AMGXB_CUDA_CHECK(cudaMemcpy(vals.data(), data, (size_t)nnz * bs * msz, cudaMemcpyDefault));
    if (world == 1) {
        std::vector<int> c32(std::max(nnz, 1));
        for (int k = 0; k < nnz; k++) c32[k] = (int)cols[k];
        upload_matrix(A, n, nnz, bx, by, rp.data(), c32.data(), vals.data(), diag_data);
        return;
    }
    dist_build_matrix(A, offsets.data(), n, nnz, bx, by, rp.data(), cols.data(), vals.data(), diag_data);
}

// AMGX_matrix_comm_from_maps_one_ring + AMGX_matrix_upload_all (examples/amgx_mpi_capi_agg.c:480-485): the caller numbers owned
// columns 0..n-1 and halo columns from n on, and says which of its rows each neighbour needs (send_maps) and which halo column each
// received value lands in (recv_maps).  The maps are kept until the upload; the upload turns the local numbering into global column
// ids (ranks own contiguous blocks in rank order; the halo ids come from one exchange of the senders' global row ids) and goes
// through the same planner as the global uploads, which re-derives interior / boundary rows and the send order it needs.
void dist_comm_from_maps_one_ring(Matrix &A, int nn, const int *neighbors, const int *send_sizes, const int **send_maps, const int *recv_sizes,
                                  const int **recv_maps)
{
    if (nn < 0 || (nn > 0 && (!neighbors || !send_sizes || !send_maps || !recv_sizes || !recv_maps)))
        fatal(AMGX_RC_BAD_PARAMETERS, "AMGX_matrix_comm_from_maps_one_ring: null map arrays");
    auto cm = std::make_shared<Matrix::CommMaps>();
    cm->neighbors.assign(neighbors, neighbors + nn);
    cm->send.resize(nn);
    cm->recv.resize(nn);
    for (int q = 0; q < nn; q++) {
        if (neighbors[q] < 0 || neighbors[q] >= A.rsc->world || neighbors[q] == A.rsc->rank)
            fatal(AMGX_RC_BAD_PARAMETERS, "AMGX_matrix_comm_from_maps_one_ring: bad neighbour rank");
        if (send_sizes[q] < 0 || recv_sizes[q] < 0) fatal(AMGX_RC_BAD_PARAMETERS, "AMGX_matrix_comm_from_maps_one_ring: negative map size");
        cm->send[q].assign(send_maps[q], send_maps[q] + send_sizes[q]);
        cm->recv[q].assign(recv_maps[q], recv_maps[q] + recv_sizes[q]);
    }
    A.comm_maps = cm;
    A.dist_pending = true;
}

void dist_upload_local(Matrix &A, int n, int nnz, int bx, int by, const int *row_ptrs, const int *col_indices, const void *data, const void *diag_data)
{
    if (!A.comm_maps) fatal(AMGX_RC_BAD_PARAMETERS, "local distributed upload without communication maps");
    const Matrix::CommMaps &cm = *A.comm_maps;
    const int world = A.rsc->world, rank = A.rsc->rank, nn = (int)cm.neighbors.size();
    if (n < 0 || nnz < 0) fatal(AMGX_RC_BAD_PARAMETERS, "Error: Failure in matrix_upload_all().");
    cudaStream_t s = A.stream();
    const size_t bs = (size_t)bx * by, msz = prec_size(A.mat_prec);
    std::vector<int> rp(n + 1), ci(std::max(nnz, 1));
    AMGXB_CUDA_CHECK(cudaMemcpy(rp.data(), row_ptrs, sizeof(int) * (n + 1), cudaMemcpyDefault));
    if (nnz) AMGXB_CUDA_CHECK(cudaMemcpy(ci.data(), col_indices, sizeof(int) * nnz, cudaMemcpyDefault));
    std::vector<char> vals(std::max<size_t>((size_t)nnz * bs * msz, 1));
    if (nnz) AMGXB_CUDA_CHECK(cudaMemcpy(vals.data(), data, (size_t)nnz * bs * msz, cudaMemcpyDefault));
    if (world == 1) {
        upload_matrix(A, n, nnz, bx, by, rp.data(), ci.data(), vals.data(), diag_data);
        A.dist_pending = false;
        return;
    }
    // 1. rows per rank -> contiguous global offsets
    DevBuf<long long> cnt;
    cnt.resize((size_t)world + 1);
    const long long mine = n;
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(cnt.ptr() + world, &mine, sizeof(mine), cudaMemcpyHostToDevice, s));
    AMGXB_NCCL_CHECK(ncclAllGather(cnt.ptr() + world, cnt.ptr(), 1, ncclInt64, (ncclComm_t)A.rsc->nccl_comm, s));
    std::vector<long long> hc = cnt.to_host(s);
    std::vector<int64_t> offsets((size_t)world + 1, 0);
    for (int r = 0; r < world; r++) offsets[r + 1] = offsets[r] + hc[r];
    // 2. global ids of the rows I send, exchanged for the global ids of my halo columns
    std::vector<int> soff(nn + 1, 0), roff(nn + 1, 0);
    for (int q = 0; q < nn; q++) {
        soff[q + 1] = soff[q] + (int)cm.send[q].size();
        roff[q + 1] = roff[q] + (int)cm.recv[q].size();
    }
    std::vector<long long> sg((size_t)std::max(soff[nn], 1)), rg((size_t)std::max(roff[nn], 1));
    for (int q = 0; q < nn; q++)
        for (size_t k = 0; k < cm.send[q].size(); k++) {
            if (cm.send[q][k] < 0 || cm.send[q][k] >= n) fatal(AMGX_RC_BAD_PARAMETERS, "send_maps holds a row index outside [0, n)");
            sg[(size_t)soff[q] + k] = offsets[rank] + cm.send[q][k];
        }
    DevBuf<long long> dsg, drg;
    dsg.from_any(sg.data(), sg.size(), s);
    drg.resize(rg.size());
    AMGXB_NCCL_CHECK(ncclGroupStart());
    for (int q = 0; q < nn; q++) {
        if (soff[q + 1] > soff[q]) AMGXB_NCCL_CHECK(ncclSend(dsg.ptr() + soff[q], (size_t)(soff[q + 1] - soff[q]), ncclInt64, cm.neighbors[q], (ncclComm_t)A.rsc->nccl_comm, s));
        if (roff[q + 1] > roff[q]) AMGXB_NCCL_CHECK(ncclRecv(drg.ptr() + roff[q], (size_t)(roff[q + 1] - roff[q]), ncclInt64, cm.neighbors[q], (ncclComm_t)A.rsc->nccl_comm, s));
    }
    AMGXB_NCCL_CHECK(ncclGroupEnd());
    rg = drg.to_host(s);
    // 3. local -> global columns, 4. the usual planner
    std::vector<int> rsz(nn);
    std::vector<const int *> rmaps(nn);
    std::vector<std::vector<int64_t>> rglob(nn);
    std::vector<const int64_t *> rgp(nn);
    for (int q = 0; q < nn; q++) {
        rsz[q] = (int)cm.recv[q].size();
        rmaps[q] = cm.recv[q].data();
        rglob[q].assign(rg.begin() + roff[q], rg.begin() + roff[q + 1]);
        rgp[q] = rglob[q].data();
    }
    std::vector<int64_t> cols((size_t)std::max(nnz, 1));
    const std::string err = comm_maps_to_global_cols(n, nnz, ci.data(), offsets[rank], nn, rsz.data(), rmaps.data(), rgp.data(), cols.data());
    if (!err.empty()) fatal(AMGX_RC_BAD_PARAMETERS, "AMGX_matrix_comm_from_maps_one_ring / upload_all: " + err);
    dist_build_matrix(A, offsets.data(), n, nnz, bx, by, rp.data(), cols.data(), vals.data(), diag_data);
    A.dist_pending = false;
}

// 7-point Poisson on a px*py*pz process grid, each rank an nx*ny*nz box (the reference's generator
// semantics: nx,ny,nz are PER RANK, src/amgx_c.cu:1690-1735).  Global numbering is rank-major: rank r owns
// global rows [r*nloc, (r+1)*nloc), local lexicographic order inside the box.
void dist_generate_poisson7(Matrix &A, int nx, int ny, int nz, int px, int py, int pz)
{
    const int world = A.rsc->world, rank = A.rsc->rank;
    const long long nloc = (long long)nx * ny * nz;
    if (nloc * 7 >= (1ll << 31)) fatal(AMGX_RC_BAD_PARAMETERS, "local grid too large for 32-bit indices");
    const int rx = rank % px, ry = (rank / px) % py, rz = rank / (px * py);
    std::vector<int64_t> offsets(world + 1);
    for (int r = 0; r <= world; r++) offsets[r] = nloc * r;
    std::vector<int> rp((size_t)nloc + 1);
    std::vector<int64_t> cols;
    std::vector<double> vals;
    cols.reserve((size_t)nloc * 7);
    vals.reserve((size_t)nloc * 7);
    auto gid = [&](int qx, int qy, int qz, int i, int j, int k) -> int64_t {
        const int q = qx + px * (qy + py * qz);
        return nloc * q + i + (long long)nx * (j + (long long)ny * k);
    };
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) {
                const size_t r = i + (size_t)nx * (j + (size_t)ny * k);
                rp[r] = (int)cols.size();
                cols.push_back(gid(rx, ry, rz, i, j, k)); vals.push_back(6.0);
                if (i > 0) { cols.push_back(gid(rx, ry, rz, i - 1, j, k)); vals.push_back(-1.0); }
                else if (rx > 0) { cols.push_back(gid(rx - 1, ry, rz, nx - 1, j, k)); vals.push_back(-1.0); }
                if (i < nx - 1) { cols.push_back(gid(rx, ry, rz, i + 1, j, k)); vals.push_back(-1.0); }
                else if (rx < px - 1) { cols.push_back(gid(rx + 1, ry, rz, 0, j, k)); vals.push_back(-1.0); }
                if (j > 0) { cols.push_back(gid(rx, ry, rz, i, j - 1, k)); vals.push_back(-1.0); }
                else if (ry > 0) { cols.push_back(gid(rx, ry - 1, rz, i, ny - 1, k)); vals.push_back(-1.0); }
                if (j < ny - 1) { cols.push_back(gid(rx, ry, rz, i, j + 1, k)); vals.push_back(-1.0); }
                else if (ry < py - 1) { cols.push_back(gid(rx, ry + 1, rz, i, 0, k)); vals.push_back(-1.0); }
                if (k > 0) { cols.push_back(gid(rx, ry, rz, i, j, k - 1)); vals.push_back(-1.0); }
                else if (rz > 0) { cols.push_back(gid(rx, ry, rz - 1, i, j, nz - 1)); vals.push_back(-1.0); }
                if (k < nz - 1) { cols.push_back(gid(rx, ry, rz, i, j, k + 1)); vals.push_back(-1.0); }
                else if (rz < pz - 1) { cols.push_back(gid(rx, ry, rz + 1, i, j, 0)); vals.push_back(-1.0); }
            }
    rp[nloc] = (int)cols.size();
    const int nnz = (int)cols.size();
    if (A.mat_prec == Prec::F64) dist_build_matrix(A, offsets.data(), (int)nloc, nnz, 1, 1, rp.data(), cols.data(), vals.data(), nullptr);
    else {
        std::vector<float> vf(vals.begin(), vals.end());
        dist_build_matrix(A, offsets.data(), (int)nloc, nnz, 1, 1, rp.data(), cols.data(), vf.data(), nullptr);
    }
}

// ---------------------------------------------------------------------------------------------
// distributed vectors: the caller's order (partition order) <-> local order [interior|boundary|halo]
// ---------------------------------------------------------------------------------------------
namespace {
template <class T> __global__ void permute_kernel(const int *__restrict__ perm, int n, int bsize, const T *__restrict__ in, T *__restrict__ out, int forward)
{
    const long long total = (long long)n * bsize;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / bsize), c = (int)(t % bsize);
        if (forward) out[(size_t)perm[i] * bsize + c] = in[t];      // out[new] = in[old]
        else out[t] = in[(size_t)perm[i] * bsize + c];             // out[old] = in[new]
    }
}
}  // namespace

void dist_prepare_vector(const Matrix &A, Vector &v)
{
    if (!A.dist) return;
    DistManager &m = *A.dist;
    const int bd = std::max(1, v.block_dim);
    const size_t need = (size_t)(m.n_owned + m.n_halo) * bd;
    cudaStream_t s = A.stream();
    if (v.dist.get() == &m && !v.user_order && v.data.n >= need) return;
    if (v.n != m.n_owned) fatal(AMGX_RC_BAD_PARAMETERS, "distributed vector size does not match the matrix partition");
    if (v.user_order) {
        DevVec nv;
        nv.resize(need, v.prec);
        nv.zero(s);
        const int grid = std::min(ceil_div((long long)m.n_owned * bd, 256), B200_SMS * 8);
        if (v.prec == Prec::F64) permute_kernel<double><<<grid, 256, 0, s>>>(m.perm_old_to_new.ptr(), m.n_owned, bd, v.data.as<double>(), nv.as<double>(), 1);
        else permute_kernel<float><<<grid, 256, 0, s>>>(m.perm_old_to_new.ptr(), m.n_owned, bd, v.data.as<float>(), nv.as<float>(), 1);
        count_launch();
        AMGXB_LAUNCH_CHECK();
        AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
        v.data.swap(nv);
    } else if (v.data.n < need) {
        DevVec nv;
        nv.resize(need, v.prec);
        nv.zero(s);
        AMGXB_CUDA_CHECK(cudaMemcpyAsync(nv.ptr(), v.data.ptr(), (size_t)m.n_owned * bd * prec_size(v.prec), cudaMemcpyDeviceToDevice, s));
        AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
        v.data.swap(nv);
    }
    v.dist = A.dist;
    v.user_order = false;
}

void dist_download_vector(const Vector &v, void *data)
{
    DistManager &m = *v.dist;
    const int bd = std::max(1, v.block_dim);
    cudaStream_t s = v.rsc->stream;
    DevVec tmp;
    tmp.resize((size_t)m.n_owned * bd, v.prec);
    const int grid = std::min(ceil_div((long long)m.n_owned * bd, 256), B200_SMS * 8);
    if (v.prec == Prec::F64) permute_kernel<double><<<grid, 256, 0, s>>>(m.perm_old_to_new.ptr(), m.n_owned, bd, v.data.as<double>(), tmp.as<double>(), 0);
    else permute_kernel<float><<<grid, 256, 0, s>>>(m.perm_old_to_new.ptr(), m.n_owned, bd, v.data.as<float>(), tmp.as<float>(), 0);
    count_launch();
    AMGXB_LAUNCH_CHECK();
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(data, tmp.ptr(), tmp.nbytes(), cudaMemcpyDefault, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
}

// ---------------------------------------------------------------------------------------------
// coarse-level communication pattern of an aggregation level (setup time, host bookkeeping).
// In: aggregates[n_owned] with ids 0..n_agg-1.  Out: aggregates[n_cols] -- owned part relabelled so
// that aggregates containing a boundary row come last (coarse rows [0, n_interior_c) then reference no
// halo column), halo part = local ids (n_agg + k) of the neighbours' aggregates -- and the coarse manager.
// Replaces the reference's setNeighborAggregates / createRenumbering bookkeeping
// (src/aggregation/aggregation_amg_level.cu:1720-1870).
// ---------------------------------------------------------------------------------------------
std::shared_ptr<DistManager> dist_coarsen(const Matrix &A, DevBuf<int> &aggregates, int n_agg, int *n_interior_c)
{
    DistManager &m = *A.dist;
    cudaStream_t s = A.stream();
    const int n = m.n_owned, nn = (int)m.neighbors.size();
    std::vector<int> h(n);
    if (n) AMGXB_CUDA_CHECK(cudaMemcpyAsync(h.data(), aggregates.ptr(), sizeof(int) * n, cudaMemcpyDeviceToHost, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    std::vector<char> bnd(std::max(n_agg, 1), 0);
    for (int i = m.n_interior; i < n; i++) bnd[h[i]] = 1;
    std::vector<int> newid(std::max(n_agg, 1));
    int c = 0;
    for (int I = 0; I < n_agg; I++) if (!bnd[I]) newid[I] = c++;
    *n_interior_c = c;
    for (int I = 0; I < n_agg; I++) if (bnd[I]) newid[I] = c++;
    for (int i = 0; i < n; i++) h[i] = newid[h[i]];
    DevBuf<int> ext;
    ext.resize((size_t)n + m.n_halo);
    ext.zero(s);
    if (n) AMGXB_CUDA_CHECK(cudaMemcpyAsync(ext.ptr(), h.data(), sizeof(int) * n, cudaMemcpyHostToDevice, s));
    dist_exchange_int(A, ext.ptr(), s);
    std::vector<int> hh(std::max(m.n_halo, 1));
    if (m.n_halo) AMGXB_CUDA_CHECK(cudaMemcpyAsync(hh.data(), ext.ptr() + n, sizeof(int) * m.n_halo, cudaMemcpyDeviceToHost, s));
    std::vector<int> smap = m.send_maps.to_host(s);
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    auto cm = std::make_shared<DistManager>();
    cm->rank = m.rank;
    cm->world = m.world;
    cm->n_owned = n_agg;
    cm->n_interior = *n_interior_c;
    cm->neighbors = m.neighbors;
    cm->halo_offsets.assign(1, 0);
    cm->send_offsets.assign(1, 0);
    std::vector<int> csend;
    for (int q = 0; q < nn; q++) {
        // coarse halo from q: unique remote coarse ids, ascending
        std::vector<int> u(hh.begin() + m.halo_offsets[q], hh.begin() + m.halo_offsets[q + 1]);
        std::sort(u.begin(), u.end());
        u.erase(std::unique(u.begin(), u.end()), u.end());
        const int base = n_agg + cm->halo_offsets.back();
        for (int k = m.halo_offsets[q]; k < m.halo_offsets[q + 1]; k++)
            hh[k] = base + (int)(std::lower_bound(u.begin(), u.end(), hh[k]) - u.begin());
        cm->halo_offsets.push_back(cm->halo_offsets.back() + (int)u.size());
        // coarse send map to q: unique aggregates of the fine rows q needs, ascending
        std::vector<int> sq;
        for (int k = m.send_offsets[q]; k < m.send_offsets[q + 1]; k++) sq.push_back(h[smap[k]]);
        std::sort(sq.begin(), sq.end());
        sq.erase(std::unique(sq.begin(), sq.end()), sq.end());
        csend.insert(csend.end(), sq.begin(), sq.end());
        cm->send_offsets.push_back((int)csend.size());
    }
    cm->n_halo = cm->halo_offsets.back();
    cm->send_maps.from_any(csend.data(), csend.size(), s);
    if (m.n_halo) AMGXB_CUDA_CHECK(cudaMemcpyAsync(ext.ptr() + n, hh.data(), sizeof(int) * m.n_halo, cudaMemcpyHostToDevice, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    aggregates.swap(ext);
    return cm;
}


// ---------------------------------------------------------------------------------------------
// Replicated coarse tail: once a level is small, every rank assembles the WHOLE level (rows of rank r occupy the
// global range [offs[r], offs[r] + counts[r]) in r's local order) and the rest of the hierarchy is built and cycled
// redundantly on every GPU -- no halo exchange below this level.  The reference's analogue is consolidation
// (amg_consolidation_flag, src/amg.cu:225-270) onto fewer ranks; with NVSwitch an all-gather of a small vector is
// cheaper than keeping ~15 latency-bound levels distributed.
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void iota_off_kernel(int n, int off, int *v) { for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = off + i; }
}

// caller_order = false: rank r's rows appear at offs[r] + (local row index), the order the distributed vectors have (replicated coarse tail);
// caller_order = true: at offs[r] + (row index in the caller's upload order), i.e. the assembled matrix IS the caller's global matrix, entry
// order included, when the ranks own contiguous blocks (classical AMG on a partitioned matrix builds its hierarchy from it).
std::unique_ptr<Matrix> dist_gather_matrix(const Matrix &A, std::vector<int> &counts, std::vector<int> &offs, bool caller_order)
{
    DistManager &m = *A.dist;
    cudaStream_t s = A.stream();
    const int world = m.world, rank = m.rank;
    const size_t bs = (size_t)A.bs(), msz = prec_size(A.mat_prec);
    // sizes of every rank
    DevBuf<int> sz_send, sz_recv;
    sz_send.resize(2);
    sz_recv.resize((size_t)2 * world);
    const int mine[2] = {A.n, A.nnz};
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(sz_send.ptr(), mine, sizeof(mine), cudaMemcpyHostToDevice, s));
    AMGXB_NCCL_CHECK(ncclAllGather(sz_send.ptr(), sz_recv.ptr(), 2, ncclInt32, comm_of(A), s));
    std::vector<int> sz = sz_recv.to_host(s);
    counts.assign(world, 0);
    offs.assign(world + 1, 0);
    std::vector<long long> nnz_off(world + 1, 0);
    int n_max = 0, nnz_max = 0;
    for (int r = 0; r < world; r++) {
        counts[r] = sz[2 * r];
        offs[r + 1] = offs[r] + counts[r];
        nnz_off[r + 1] = nnz_off[r] + sz[2 * r + 1];
        n_max = std::max(n_max, sz[2 * r]);
        nnz_max = std::max(nnz_max, sz[2 * r + 1]);
    }
    const int N = offs[world];
    const long long NNZ = nnz_off[world];
    if (NNZ > 0x7fffffffll) fatal(AMGX_RC_INTERNAL, "gathered level too large");
    // global ids of the local columns (owned: offset + local index; halo: asked from the owner)
    DevBuf<int> gid;
    gid.resize((size_t)std::max(A.n_cols, 1));
    std::vector<int> h_perm;        // caller_order: position in the caller's order -> local row
    if (caller_order) {
        h_perm = m.perm_old_to_new.to_host(s);
        std::vector<int> g0((size_t)std::max(A.n_cols, 1), 0);
        for (int old = 0; old < A.n; old++) g0[h_perm[old]] = offs[rank] + old;
        gid.from_any(g0.data(), g0.size(), s);
    } else {
        iota_off_kernel<<<std::max(1, std::min(ceil_div(A.n_cols, 256), 1024)), 256, 0, s>>>(A.n_cols, offs[rank], gid.ptr());
        count_launch();
    }
    dist_exchange_int(A, gid.ptr(), s);
    std::vector<int> h_gid = gid.to_host(s), h_rp = A.row_ptr.to_host(s), h_ci = A.col_idx.to_host(s);
    // padded all-gather of row lengths, global columns and values
    std::vector<int> len_pad((size_t)std::max(n_max, 1), 0), col_pad((size_t)std::max(nnz_max, 1), 0);
    DevBuf<int> d_len, d_col, r_len, r_col;
    DevBytes d_val, r_val;
    const size_t val_pad = (size_t)std::max(nnz_max, 1) * bs * msz;
    d_val.resize(val_pad);
    r_val.resize(val_pad * world);
    AMGXB_CUDA_CHECK(cudaMemsetAsync(d_val.p, 0, val_pad, s));
    if (!caller_order) {
        for (int i = 0; i < A.n; i++) len_pad[i] = h_rp[i + 1] - h_rp[i];
        for (int k = 0; k < A.nnz; k++) col_pad[k] = h_gid[h_ci[k]];
        if (A.nnz) AMGXB_CUDA_CHECK(cudaMemcpyAsync(d_val.p, A.values.ptr(), (size_t)A.nnz * bs * msz, cudaMemcpyDeviceToDevice, s));
    } else {
        std::vector<char> v_loc((size_t)std::max(A.nnz, 1) * bs * msz), v_out((size_t)std::max(A.nnz, 1) * bs * msz);
        if (A.nnz) AMGXB_CUDA_CHECK(cudaMemcpyAsync(v_loc.data(), A.values.ptr(), (size_t)A.nnz * bs * msz, cudaMemcpyDeviceToHost, s));
        AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
        size_t k = 0;
        for (int old = 0; old < A.n; old++) {
            const int i = h_perm[old], len = h_rp[i + 1] - h_rp[i];
            len_pad[old] = len;
            for (int kk = h_rp[i]; kk < h_rp[i + 1]; kk++) col_pad[k + (size_t)(kk - h_rp[i])] = h_gid[h_ci[kk]];
            if (len) memcpy(&v_out[k * bs * msz], &v_loc[(size_t)h_rp[i] * bs * msz], (size_t)len * bs * msz);
            k += (size_t)len;
        }
        if (A.nnz) AMGXB_CUDA_CHECK(cudaMemcpyAsync(d_val.p, v_out.data(), (size_t)A.nnz * bs * msz, cudaMemcpyHostToDevice, s));
        AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));      // v_out goes out of scope below
    }
    d_len.from_any(len_pad.data(), len_pad.size(), s);
    d_col.from_any(col_pad.data(), col_pad.size(), s);
    r_len.resize(len_pad.size() * world);
    r_col.resize(col_pad.size() * world);
    AMGXB_NCCL_CHECK(ncclGroupStart());
    AMGXB_NCCL_CHECK(ncclAllGather(d_len.ptr(), r_len.ptr(), len_pad.size(), ncclInt32, comm_of(A), s));
    AMGXB_NCCL_CHECK(ncclAllGather(d_col.ptr(), r_col.ptr(), col_pad.size(), ncclInt32, comm_of(A), s));
    AMGXB_NCCL_CHECK(ncclAllGather(d_val.p, r_val.p, val_pad, ncclChar, comm_of(A), s));
    AMGXB_NCCL_CHECK(ncclGroupEnd());
    std::vector<int> all_len = r_len.to_host(s), all_col = r_col.to_host(s);
    std::vector<char> all_val(val_pad * world);
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(all_val.data(), r_val.p, all_val.size(), cudaMemcpyDeviceToHost, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    std::vector<int> g_rp((size_t)N + 1, 0), g_ci((size_t)std::max<long long>(NNZ, 1));
    std::vector<char> g_va((size_t)std::max<long long>(NNZ, 1) * bs * msz);
    for (int r = 0; r < world; r++) {
        for (int i = 0; i < counts[r]; i++) g_rp[(size_t)offs[r] + i + 1] = all_len[(size_t)r * len_pad.size() + i];
        const int nz = sz[2 * r + 1];
        if (nz) {
            memcpy(&g_ci[(size_t)nnz_off[r]], &all_col[(size_t)r * col_pad.size()], sizeof(int) * (size_t)nz);
            memcpy(&g_va[(size_t)nnz_off[r] * bs * msz], &all_val[(size_t)r * val_pad], (size_t)nz * bs * msz);
        }
    }
    for (int i = 0; i < N; i++) g_rp[i + 1] += g_rp[i];
    std::unique_ptr<Matrix> G(new Matrix);
    G->rsc = A.rsc;
    G->mode = A.mode;
    G->mat_prec = A.mat_prec;
    G->vec_prec = A.vec_prec;
    upload_matrix(*G, N, (int)NNZ, A.bx, A.by, g_rp.data(), g_ci.data(), g_va.data(), nullptr);
    if (getenv("AMGXB_TAIL_CHECK") && !caller_order && bs == 1 && A.vec_prec == Prec::F64) {
        // self-check: y = A x through the distributed operator must equal the owned slice of G x for x(g) = sin(1 + 0.37 g);
        // and an all-gather of the owned global ids must reproduce 0..N-1
        std::vector<double> xg((size_t)N), xl((size_t)A.n_cols);
        for (int g = 0; g < N; g++) xg[g] = std::sin(1.0 + 0.37 * g);
        for (int c = 0; c < A.n_cols; c++) xl[c] = xg[h_gid[c]];
        DevVec dx, dy, gx, gy;
        dx.resize((size_t)A.n_cols, Prec::F64); dy.resize((size_t)A.n_cols, Prec::F64);
        gx.resize((size_t)N, Prec::F64); gy.resize((size_t)N, Prec::F64);
        AMGXB_CUDA_CHECK(cudaMemcpyAsync(dx.ptr(), xl.data(), sizeof(double) * xl.size(), cudaMemcpyHostToDevice, s));
        AMGXB_CUDA_CHECK(cudaMemcpyAsync(gx.ptr(), xg.data(), sizeof(double) * xg.size(), cudaMemcpyHostToDevice, s));
        CsrOpArgs a1, a2;
        a1.x = dx.ptr(); a1.y = dy.ptr();
        csr_op(A, EPI_SPMV, a1, s, 0);
        a2.x = gx.ptr(); a2.y = gy.ptr();
        csr_op(*G, EPI_SPMV, a2, s, 0);
        std::vector<double> y1((size_t)A.n_cols), y2((size_t)N);
        AMGXB_CUDA_CHECK(cudaMemcpyAsync(y1.data(), dy.ptr(), sizeof(double) * y1.size(), cudaMemcpyDeviceToHost, s));
        AMGXB_CUDA_CHECK(cudaMemcpyAsync(y2.data(), gy.ptr(), sizeof(double) * y2.size(), cudaMemcpyDeviceToHost, s));
        AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
        double md = 0, mx = 0;
        for (int i = 0; i < A.n; i++) { md = std::max(md, std::fabs(y1[i] - y2[(size_t)offs[rank] + i])); mx = std::max(mx, std::fabs(y1[i])); }
        // symmetry of G and row sums
        double asym = 0;
        {
            std::vector<double> va((size_t)NNZ);
            memcpy(va.data(), g_va.data(), sizeof(double) * (size_t)NNZ);
            for (int i = 0; i < N && i < 4000; i++)
                for (int k = g_rp[i]; k < g_rp[i + 1]; k++) {
                    const int j = g_ci[k];
                    double aji = 0; bool f = false;
                    for (int kk = g_rp[j]; kk < g_rp[j + 1]; kk++) if (g_ci[kk] == i) { aji = va[kk]; f = true; break; }
                    asym = std::max(asym, f ? std::fabs(aji - va[k]) : 1e30);
                }
        }
        // all-gather check
        DevVec tv;
        tv.resize((size_t)N, Prec::F64);
        tv.zero(s);
        std::vector<double> mine_v((size_t)A.n);
        for (int i = 0; i < A.n; i++) mine_v[i] = offs[rank] + i;
        AMGXB_CUDA_CHECK(cudaMemcpyAsync((double *)tv.ptr() + offs[rank], mine_v.data(), sizeof(double) * mine_v.size(), cudaMemcpyHostToDevice, s));
        dist_allgatherv_inplace(A, tv.ptr(), Prec::F64, 1, counts, offs, s);
        std::vector<double> tvh((size_t)N);
        AMGXB_CUDA_CHECK(cudaMemcpyAsync(tvh.data(), tv.ptr(), sizeof(double) * tvh.size(), cudaMemcpyDeviceToHost, s));
        AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
        int bad = 0;
        for (int g = 0; g < N; g++) bad += (tvh[g] != (double)g);
        unsigned long long cks = 1469598103934665603ull;
        for (long long k = 0; k < NNZ; k++) cks = (cks ^ (unsigned)g_ci[k]) * 1099511628211ull;
        fprintf(stderr, "[tail-check rank %d/%d] N=%d NNZ=%lld n_owned=%d n_halo=%d nbrs=%d | max|A x - (G x)_slice| = %.3e (max|y| %.3e) | asym(G) = %.3e | allgatherv mismatches = %d | cks %llx\n",
                rank, world, N, NNZ, A.n, m.n_halo, (int)m.neighbors.size(), md, mx, asym, bad, cks);
    }
    return G;
}

void dist_allreduce_vec(const Matrix &A, void *v, Prec prec, size_t n, cudaStream_t s)
{
    if (!A.dist || n == 0) return;
    AMGXB_NCCL_CHECK(ncclAllReduce(v, v, n, prec == Prec::F64 ? ncclDouble : ncclFloat, ncclSum, comm_of(A), s));
}

void dist_allgatherv_int_inplace(const Matrix &A, int *v, const std::vector<int> &counts, const std::vector<int> &offs, cudaStream_t s)
{
    AMGXB_NCCL_CHECK(ncclGroupStart());
    for (int r = 0; r < (int)counts.size(); r++) {
        if (counts[r] == 0) continue;
        AMGXB_NCCL_CHECK(ncclBroadcast(v + offs[r], v + offs[r], (size_t)counts[r], ncclInt32, r, comm_of(A), s));
    }
    AMGXB_NCCL_CHECK(ncclGroupEnd());
}

// v (global length offs[world], block dim bsize): every rank contributes its slice [offs[rank], +counts[rank]) in place
void dist_allgatherv_inplace(const Matrix &A, void *v, Prec prec, int bsize, const std::vector<int> &counts, const std::vector<int> &offs, cudaStream_t s)
{
    const size_t esz = prec_size(prec);
    const ncclDataType_t dt = prec == Prec::F64 ? ncclDouble : ncclFloat;
    AMGXB_NCCL_CHECK(ncclGroupStart());
    for (int r = 0; r < (int)counts.size(); r++) {
        if (counts[r] == 0) continue;
        char *p = (char *)v + (size_t)offs[r] * bsize * esz;
        AMGXB_NCCL_CHECK(ncclBroadcast(p, p, (size_t)counts[r] * bsize, dt, r, comm_of(A), s));
    }
    AMGXB_NCCL_CHECK(ncclGroupEnd());
}

}  // namespace amgxb
