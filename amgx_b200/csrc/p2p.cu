// p2p.cu -- halo exchange and scalar all-reduce over NVLink PEER MEMORY (one process per GPU, CUDA IPC), replacing the
// pack-kernel -> event -> grouped ncclSend/ncclRecv -> event sequence and the 8-byte ncclAllReduce + 1-thread epilogue of
// dist.cu on the solve path.  What it replaces in the reference: DistributedManager::exchange_halo / global_reduce_sum over
// host-staged MPI (include/distributed/distributed_manager.h:955-1170, src/distributed/comms_mpi_hostbuffer_stream.cu:598-700).
//
// Why: round 1's 1 -> 2 GPU cliff was latency, not bandwidth: ~22 levels x 3 sweeps of (pack kernel, event, NCCL group launch on
// a side stream, event) plus three serial NCCL all-reduces of one double per PCG iteration.  On NVSwitch every GPU can store into
// every peer's memory directly, so
//   * the boundary values are written by ONE kernel straight into the neighbour's receive window (the "pack" IS the send),
//     followed by a release-store of an epoch flag in the neighbour's memory;
//   * the receiver runs a small kernel in front of the rows that read halo columns: it acquires the flags of its neighbours and
//     copies the window into the halo tail of the vector;
//   * a scalar all-reduce is one 32-thread kernel: every rank stores its partial into every peer's slot array, flags it, waits for
//     the world_size flags of its own array and sums the slots in RANK ORDER (bit-identical on every rank), then applies the same
//     scalar epilogue (alpha / beta / sqrt) the single-GPU reduction kernels apply.
// Everything is stream-ordered kernels, hence capturable in the PCG CUDA-graph segments; epochs live in device memory so a graph
// replay advances them by itself.  Double buffering by epoch parity is sufficient because exchanges are pairwise symmetric and
// every push is followed by its wait before the next push of the same manager (proof sketch in DESIGN.md, multi-GPU section).
// If CUDA IPC is unavailable the ranks agree (all-reduce of a flag) to keep the NCCL path of dist.cu.
#include "solvers.h"
#include "dist.h"
#include "p2p.h"
#include "capi_internal.h"
#include <nccl.h>

namespace amgxb {

#define P2P_NCCL_CHECK(expr)                                                                      \
    do {                                                                                          \
        ncclResult_t _r = (expr);                                                                 \
        if (_r != ncclSuccess) {                                                                  \
            char _b[512];                                                                         \
            snprintf(_b, sizeof(_b), "NCCL error %s at %s:%d", ncclGetErrorString(_r), __FILE__, __LINE__); \
            throw ::amgxb::Error(AMGX_RC_CORE, _b);                                               \
        }                                                                                         \
    } while (0)

namespace {

typedef unsigned long long u64;

__device__ __forceinline__ void st_release_sys(u64 *p, u64 v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld_acquire_sys(const u64 *p)
{
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// spin until *flag >= e; a peer that never arrives (crashed rank) must not hang the GPU for ever: trap after ~30 s
__device__ __forceinline__ void wait_flag(const u64 *flag, u64 e)
{
    const long long t0 = clock64();
    int spins = 0;
    while (ld_acquire_sys(flag) < e) {
        if (++spins > 4096) __nanosleep(128);      // the common case completes within a few microseconds: poll hot first
        if ((spins & 1023) == 0 && clock64() - t0 > 60000000000ll) { printf("[amgx_b200] peer-memory wait timed out (flag %p, epoch %llu)\n", (const void *)flag, e); __trap(); }
    }
}

// per-manager device state
struct LinkDev {
    int nn;
    int send_begin[P2P_MAX_NEIGHBORS], send_end[P2P_MAX_NEIGHBORS];
    char *peer_data[P2P_MAX_NEIGHBORS];        // where MY values land in neighbour q's window (parity 0)
    u64 peer_parity_stride[P2P_MAX_NEIGHBORS]; // bytes between neighbour q's two parity buffers
    u64 *peer_flag[P2P_MAX_NEIGHBORS];         // my flag in neighbour q's window
    char *stage;                               // my receive window: 2 x stage_stride bytes
    u64 stage_stride;
    u64 *flags;                                // my flags, one per neighbour (written by the neighbours)
    u64 send_epoch, recv_epoch;
    unsigned counter_push, counter_wait;
};

// The exchange: push my boundary values into the neighbours' windows, release the epoch flags (last CTA to finish its stores), acquire
// the neighbours' flags, copy my window into the halo tail of x -- ONE kernel.  No intra-grid barrier is needed: what a CTA waits for
// comes from the neighbours' grids.  The grid is small (<= 32 CTAs) so that it is always fully resident (two grids spinning on each
// other's flags must both be able to finish their pushes).  Ordering: a CTA's peer stores are made visible at system scope by ONE
// fence of its thread 0 after the CTA barrier (cumulativity), before its ticket; whoever draws the last ticket therefore knows every
// CTA's data has been performed at system scope and only has to keep its own release stores of the flags behind that observation.
template <class T> __global__ void __launch_bounds__(256) p2p_exchange_kernel(LinkDev *d, const int *__restrict__ map, T *x, int bsize, long long halo_count, long long halo_first)
{
    const u64 e = *(volatile u64 *)&d->send_epoch + 1;
    const int nn = d->nn;
    for (int q = 0; q < nn; q++) {
        T *dst = reinterpret_cast<T *>(d->peer_data[q] + (e & 1) * d->peer_parity_stride[q]);
        const int b0 = d->send_begin[q];
        const long long total = (long long)(d->send_end[q] - b0) * bsize;
        for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
            const int k = (int)(t / bsize), c = (int)(t % bsize);
            dst[t] = x[(size_t)map[b0 + k] * bsize + c];
        }
    }
    __syncthreads();
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        __threadfence_system();
        is_last = (gridDim.x == 1) || (atomicAdd(&d->counter_push, 1u) == gridDim.x - 1);
        __threadfence();
    }
    __syncthreads();
    if (is_last && (int)threadIdx.x < nn) st_release_sys(d->peer_flag[threadIdx.x], e);
    if ((int)threadIdx.x < nn) wait_flag(d->flags + threadIdx.x, e);
    __syncthreads();
    const T *src = reinterpret_cast<const T *>(d->stage + (e & 1) * d->stage_stride);
    T *x_halo = x + halo_first;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < halo_count; t += (long long)gridDim.x * blockDim.x) x_halo[t] = __ldcg(src + t);
    __syncthreads();
    if (threadIdx.x == 0) {
        const bool last2 = (gridDim.x == 1) || (atomicAdd(&d->counter_wait, 1u) == gridDim.x - 1);
        if (last2) { d->send_epoch = e; d->recv_epoch = e; d->counter_push = 0; d->counter_wait = 0; }
    }
}

// per-resources device state of the scalar all-reduce
struct RedDev {
    int world, rank;
    double *peer_slots[P2P_MAX_WORLD];   // peer r's slot array [2][world]
    u64 *peer_flags[P2P_MAX_WORLD];      // peer r's flag array [2][world]
    double *slots;                       // mine
    u64 *flags;
    u64 epoch;
};

// op: 0 sum, 2 max.  post: 0 = FinOp epilogue `fin_op` on scal[slot]; 1 = norm epilogue (sqrt when do_sqrt, host mirror)
__global__ void __launch_bounds__(32) p2p_allreduce_kernel(RedDev *d, double *scal, int slot, int op, int post, int fin_op, int do_sqrt, double *host_mirror)
{
    const int lane = threadIdx.x, world = d->world;
    const u64 e = *(volatile u64 *)&d->epoch + 1;
    const int par = (int)(e & 1);
    const double mine = scal[slot];
    if (lane < world) *(volatile double *)(d->peer_slots[lane] + par * world + d->rank) = mine;
    __threadfence_system();
    if (lane < world) st_release_sys(d->peer_flags[lane] + par * world + d->rank, e);
    if (lane < world) wait_flag(d->flags + par * world + lane, e);
    __syncwarp();
    if (lane == 0) {
        double acc = __ldcg(d->slots + par * world);
        for (int r = 1; r < world; r++) {
            const double v = __ldcg(d->slots + par * world + r);
            acc = (op == 2) ? fmax(acc, v) : acc + v;
        }
        if (post == 1) {
            if (do_sqrt) acc = sqrt(acc);
            scal[slot] = acc;
            if (host_mirror) { host_mirror[slot] = acc; __threadfence_system(); }
        } else {
            if (fin_op == FIN_PCG_ALPHA) {
                scal[S_DOT] = acc;
                const double a = (acc != 0.0) ? scal[S_RZ] / acc : 0.0;
                scal[S_ALPHA] = a;
                scal[S_NEG_ALPHA] = -a;
                scal[slot] = acc;
            } else if (fin_op == FIN_PCG_BETA) {
                const double old = scal[S_RZ];
                scal[S_RZ_OLD] = old;
                scal[S_RZ] = acc;
                scal[S_BETA] = (old != 0.0) ? acc / old : 0.0;
                scal[slot] = acc;
            } else if (fin_op == FIN_SQRT) {
                scal[slot] = sqrt(acc);
            } else {
                scal[slot] = acc;
            }
        }
        d->epoch = e;
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct PeerWindow {
    int rank = 0, world = 1;
    char *base = nullptr;                 // my window
    size_t bytes = 0, used = 0;
    std::vector<char *> peer_base;        // opened windows of the peers (own entry = base)
    RedDev *red = nullptr;                // device
    bool ok = false;
};

struct P2PLink {
    LinkDev *dev = nullptr;
    size_t stage_stride = 0;
    int bsize = 1;
    ~P2PLink() { if (dev) cudaFree(dev); }
};

static PeerWindow *window_of(const Resources *rsc) { return reinterpret_cast<PeerWindow *>(rsc->p2p); }

static bool p2p_enabled_env()
{
    const char *e = getenv("AMGXB_P2P");
    return !e || atoi(e) != 0;
}

// bump allocation inside my window (256-byte aligned); returns the offset or (size_t)-1
static size_t window_alloc(PeerWindow &w, size_t bytes)
{
    const size_t off = (w.used + 255) & ~(size_t)255;
    if (off + bytes > w.bytes) return (size_t)-1;
    w.used = off + bytes;
    return off;
}

// all ranks: gather `mine` (count ints64) from everybody through the device
static std::vector<long long> allgather_ll(Resources *rsc, const std::vector<long long> &mine)
{
    const int world = rsc->world;
    DevBuf<long long> buf;
    buf.resize(mine.size() * (size_t)(world + 1));
    cudaStream_t s = rsc->stream;
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(buf.ptr() + mine.size() * world, mine.data(), mine.size() * sizeof(long long), cudaMemcpyHostToDevice, s));
    P2P_NCCL_CHECK(ncclAllGather(buf.ptr() + mine.size() * world, buf.ptr(), mine.size(), ncclInt64, (ncclComm_t)rsc->nccl_comm, s));
    std::vector<long long> all = buf.to_host(s);
    all.resize(mine.size() * (size_t)world);
    return all;
}

void p2p_init(Resources *rsc)
{
    rsc->p2p = nullptr;
    if (rsc->world <= 1 || !rsc->nccl_comm) return;
    const int world = rsc->world, rank = rsc->rank;
    std::unique_ptr<PeerWindow> w(new PeerWindow);
    w->rank = rank;
    w->world = world;
    long long ok = (p2p_enabled_env() && world <= P2P_MAX_WORLD) ? 1 : 0;
    size_t mb = 64;
    if (const char *e = getenv("AMGXB_P2P_WINDOW_MB")) mb = (size_t)std::max(1, atoi(e));
    cudaIpcMemHandle_t handle;
    memset(&handle, 0, sizeof(handle));
    if (ok) {
        w->bytes = mb << 20;
        if (cudaMalloc((void **)&w->base, w->bytes) != cudaSuccess) { cudaGetLastError(); ok = 0; w->base = nullptr; }
    }
    if (ok) {
        if (cudaMemset(w->base, 0, w->bytes) != cudaSuccess || cudaIpcGetMemHandle(&handle, w->base) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    }
    // exchange (ok, handle) as 1 + 8 int64 words per rank
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    std::vector<long long> mine(9, 0);
    mine[0] = ok;
    memcpy(&mine[1], &handle, 64);
    std::vector<long long> all = allgather_ll(rsc, mine);
    bool all_ok = true;
    for (int r = 0; r < world; r++) all_ok = all_ok && all[(size_t)r * 9] == 1;
    long long opened = all_ok ? 1 : 0;
    if (all_ok) {
        w->peer_base.assign(world, nullptr);
        w->peer_base[rank] = w->base;
        for (int r = 0; r < world && opened; r++) {
            if (r == rank) continue;
            cudaIpcMemHandle_t h;
            memcpy(&h, &all[(size_t)r * 9 + 1], 64);
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); opened = 0; break; }
            w->peer_base[r] = (char *)p;
        }
    }
    // second agreement round: every rank opened every window
    std::vector<long long> all2 = allgather_ll(rsc, std::vector<long long>(1, opened));
    for (int r = 0; r < world; r++) all_ok = all_ok && all2[r] == 1;
    if (!all_ok) {
        for (int r = 0; r < (int)w->peer_base.size(); r++)
            if (r != rank && w->peer_base[r]) cudaIpcCloseMemHandle(w->peer_base[r]);
        if (w->base) cudaFree(w->base);
        if (getenv("AMGXB_P2P_VERBOSE")) fprintf(stderr, "[amgx_b200] rank %d: peer-memory path unavailable, using NCCL send/recv\n", rank);
        return;
    }
    // scalar all-reduce area: slots [2][world] doubles + flags [2][world] u64 at the start of every window (same offset on every rank)
    const size_t slots_off = window_alloc(*w, sizeof(double) * 2 * world), flags_off = window_alloc(*w, sizeof(u64) * 2 * world);
    RedDev h;
    memset(&h, 0, sizeof(h));
    h.world = world;
    h.rank = rank;
    for (int r = 0; r < world; r++) {
        h.peer_slots[r] = reinterpret_cast<double *>(w->peer_base[r] + slots_off);
        h.peer_flags[r] = reinterpret_cast<u64 *>(w->peer_base[r] + flags_off);
    }
    h.slots = reinterpret_cast<double *>(w->base + slots_off);
    h.flags = reinterpret_cast<u64 *>(w->base + flags_off);
    h.epoch = 0;
    AMGXB_CUDA_CHECK(cudaMalloc((void **)&w->red, sizeof(RedDev)));
    AMGXB_CUDA_CHECK(cudaMemcpy(w->red, &h, sizeof(h), cudaMemcpyHostToDevice));
    w->ok = true;
    if (getenv("AMGXB_P2P_VERBOSE")) fprintf(stderr, "[amgx_b200] rank %d/%d: peer-memory window %zu MB mapped on all ranks\n", rank, world, mb);
    rsc->p2p = w.release();
}

void p2p_shutdown(Resources *rsc)
{
    PeerWindow *w = window_of(rsc);
    if (!w) return;
    // nobody may still be storing into my window: meet the peers first (best effort -- a failing collective must not throw from a destructor)
    if (rsc->nccl_comm) {
        double *tmp = nullptr;
        if (cudaMalloc((void **)&tmp, sizeof(double)) == cudaSuccess) {
            cudaMemset(tmp, 0, sizeof(double));
            ncclAllReduce(tmp, tmp, 1, ncclDouble, ncclSum, (ncclComm_t)rsc->nccl_comm, rsc->stream);
            cudaStreamSynchronize(rsc->stream);
            cudaFree(tmp);
        }
    }
    cudaDeviceSynchronize();
    for (int r = 0; r < (int)w->peer_base.size(); r++)
        if (r != w->rank && w->peer_base[r]) cudaIpcCloseMemHandle(w->peer_base[r]);
    if (w->red) cudaFree(w->red);
    if (w->base) cudaFree(w->base);
    delete w;
    rsc->p2p = nullptr;
}

bool p2p_available(const Resources *rsc) { return window_of(rsc) && window_of(rsc)->ok; }

// Collective over the ranks of A (every rank calls it for the same manager in the same order): carve the receive window of this
// manager out of my peer window and tell every neighbour where its values go.
void p2p_manager_setup(const Matrix &A)
{
    if (!A.dist) return;
    DistManager &m = *A.dist;
    m.p2p.reset();
    PeerWindow *w = window_of(A.rsc.get());
    if (!w || !w->ok) return;
    cudaStream_t s = A.stream();
    const int nn = (int)m.neighbors.size();
    const int bsize = std::max(1, A.bx);
    const size_t stride = (((size_t)m.n_halo * bsize * sizeof(double)) + 255) & ~(size_t)255;
    const size_t saved_used = w->used;
    long long ok = nn <= P2P_MAX_NEIGHBORS ? 1 : 0;
    size_t stage_off = 0, flags_off = 0;
    if (ok) {
        stage_off = window_alloc(*w, 2 * std::max<size_t>(stride, 256));
        flags_off = window_alloc(*w, sizeof(u64) * std::max(nn, 1));
        if (stage_off == (size_t)-1 || flags_off == (size_t)-1) ok = 0;
    }
    // all ranks must take the same path for this manager
    {
        DevBuf<long long> f;
        f.resize(1);
        AMGXB_CUDA_CHECK(cudaMemcpyAsync(f.ptr(), &ok, sizeof(ok), cudaMemcpyHostToDevice, s));
        P2P_NCCL_CHECK(ncclAllReduce(f.ptr(), f.ptr(), 1, ncclInt64, ncclMin, (ncclComm_t)A.rsc->nccl_comm, s));
        ok = f.to_host(s)[0];
    }
    if (!ok) { w->used = saved_used; return; }
    // tell neighbour q: (offset of its slice in my window, my parity stride, offset of its flag in my window)
    std::vector<long long> tell((size_t)std::max(nn, 1) * 3, 0), told((size_t)std::max(nn, 1) * 3, 0);
    for (int q = 0; q < nn; q++) {
        tell[(size_t)q * 3 + 0] = (long long)(stage_off + (size_t)m.halo_offsets[q] * bsize * sizeof(double));
        tell[(size_t)q * 3 + 1] = (long long)stride;
        tell[(size_t)q * 3 + 2] = (long long)(flags_off + sizeof(u64) * q);
    }
    if (nn) {
        DevBuf<long long> dt, dr;
        dt.from_any(tell.data(), tell.size(), s);
        dr.resize(told.size());
        P2P_NCCL_CHECK(ncclGroupStart());
        for (int q = 0; q < nn; q++) {
            P2P_NCCL_CHECK(ncclSend(dt.ptr() + (size_t)q * 3, 3, ncclInt64, m.neighbors[q], (ncclComm_t)A.rsc->nccl_comm, s));
            P2P_NCCL_CHECK(ncclRecv(dr.ptr() + (size_t)q * 3, 3, ncclInt64, m.neighbors[q], (ncclComm_t)A.rsc->nccl_comm, s));
        }
        P2P_NCCL_CHECK(ncclGroupEnd());
        told = dr.to_host(s);
    }
    LinkDev h;
    memset(&h, 0, sizeof(h));
    h.nn = nn;
    for (int q = 0; q < nn; q++) {
        h.send_begin[q] = m.send_offsets[q];
        h.send_end[q] = m.send_offsets[q + 1];
        char *pb = w->peer_base[m.neighbors[q]];
        h.peer_data[q] = pb + told[(size_t)q * 3 + 0];
        h.peer_parity_stride[q] = (u64)told[(size_t)q * 3 + 1];
        h.peer_flag[q] = reinterpret_cast<u64 *>(pb + told[(size_t)q * 3 + 2]);
    }
    h.stage = w->base + stage_off;
    h.stage_stride = stride;
    h.flags = reinterpret_cast<u64 *>(w->base + flags_off);
    auto link = std::make_shared<P2PLink>();
    AMGXB_CUDA_CHECK(cudaMalloc((void **)&link->dev, sizeof(LinkDev)));
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(link->dev, &h, sizeof(h), cudaMemcpyHostToDevice, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    link->stage_stride = stride;
    link->bsize = bsize;
    m.p2p = link;
}

// the whole exchange as one kernel on stream s (see p2p_exchange_kernel): in stream order, the halo tail of x is filled afterwards.
// Returns false when this manager has no peer-memory link (the caller then takes the NCCL path).
bool p2p_exchange_blocking(const Matrix &A, void *x, Prec prec, int bsize, cudaStream_t s)
{
    DistManager &m = *A.dist;
    if (!m.p2p || bsize != m.p2p->bsize) return false;
    if (m.neighbors.empty()) return true;
    const int nn = (int)m.neighbors.size();
    const long long work = std::max<long long>((long long)m.send_offsets[nn], (long long)m.n_halo) * bsize;
    static const int max_ctas = getenv("AMGXB_P2P_CTAS") ? std::max(1, std::min(32, atoi(getenv("AMGXB_P2P_CTAS")))) : 32;     // r02, N = 2: 2 CTAs 223, 8 CTAs 272, 32 CTAs 283 global it/s
    const int grid = (int)std::max<long long>(1, std::min<long long>((work + 2047) / 2048, max_ctas));
    const long long halo_count = (long long)m.n_halo * bsize, halo_first = (long long)m.n_owned * bsize;
    if (prec == Prec::F64) p2p_exchange_kernel<double><<<grid, 256, 0, s>>>(m.p2p->dev, m.send_maps.ptr(), (double *)x, bsize, halo_count, halo_first);
    else p2p_exchange_kernel<float><<<grid, 256, 0, s>>>(m.p2p->dev, m.send_maps.ptr(), (float *)x, bsize, halo_count, halo_first);
    count_launch();
    AMGXB_LAUNCH_CHECK();
    return true;
}

// scal[slot] <- op over the ranks of scal[slot], then the scalar epilogue; one 32-thread kernel.  Returns false without a window.
bool p2p_allreduce_scalar(const Matrix &A, const ReduceCtx &red, int slot, int op, int post, int fin_op, int do_sqrt, bool mirror, cudaStream_t s)
{
    PeerWindow *w = window_of(A.rsc.get());
    if (!w || !w->ok) return false;
    p2p_allreduce_kernel<<<1, 32, 0, s>>>(w->red, red.scal, slot, op, post, fin_op, do_sqrt, mirror ? red.host_mirror : nullptr);
    count_launch();
    AMGXB_LAUNCH_CHECK();
    return true;
}

}  // namespace amgxb
