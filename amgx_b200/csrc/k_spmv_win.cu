// k_spmv_win.cu -- the CSR tile kernel for BANDED irregular matrices (SuiteSparse-shaped: random columns within a few thousand rows of
// the diagonal): a sliding window of x lives in shared memory.
//
// Why: on such a matrix every entry gathers 8 bytes of x from a different 32-byte sector.  The matrix streams through TMA at 10-12 B per
// entry, but the gathers pull 32 B per entry through L2 -> L1 (r02: 0.27 ms per SpMV on the 4 M-row matrix = 0.45 of the copy peak by the
// 12 B/entry count, with DRAM far from saturated).  The columns of a row tile, however, fall into [r0 - W, r0 + T + W) for a W of a few
// sigma -- and that window moves by only T rows from one tile to the next.  So:
//   * a CTA owns a CONTIGUOUS range of tiles (one CTA per SM) and keeps x[r0 - W, r0 + S*T + W) in a ring buffer in shared memory
//     (index = column & (R - 1)); advancing one tile costs ONE 2 KB bulk copy of the T new entries -- x is read from L2 / DRAM once per CTA
//     range instead of once per entry;
//   * the columns travel as 16-bit (column - row) offsets (2 B per entry; an offset that does not fit, or a column outside the window,
//     falls back to the 32-bit column / a global gather of x -- the window is a cache, never a restriction);
//   * values and offsets are stored a second time in SLICED-ELL order (per 256-row tile the rows sorted by length, 32-row slices stored
//     entry-major, SELL-32-256): a thread still owns one row and sums it left to right with FMA -- the same bits as every other CSR kernel
//     here -- but the 32 lanes of a warp now read 32 consecutive shared-memory words per step.  First version (CSR order in shared
//     memory, lanes one row length apart): 27 shared-memory wavefronts per warp step, l1tex LSU pipe 70 % busy, 0.271 ms (ncu r02);
//   * the consumers read values, offsets and x from shared memory only.
// Ring safety: the producer runs at most S tiles ahead, tile p's chunk x[r_p + W, r_p + T + W) overwrites the slots of
// x[r_p + W - R, ...), which the oldest tile still being consumed (p - S + 1) no longer addresses iff S*T + 2W <= R.
// Chosen per level at plan time from the matrix's own offset statistics (csr_build_window); AMGXB_WINDOW=0 disables it.
#include "kernels.h"
#include <map>

namespace amgxb {
namespace {

#include "tile_common.cuh"

constexpr int WIN_T = 256;
constexpr int WIN_BIAS = 32768;          // offsets are stored biased (unsigned 16 bits: one LDS.U16, no sign extension)
constexpr unsigned short WIN_FAR = 0;    // "read the 32-bit column": column - row does not fit (lands below the window by construction)

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

struct WinArgs {
    const void *sv;             // values in sliced-ELL order (MatT)
    const unsigned short *so;            // column - row in the same order (WIN_FAR: read the 32-bit column of the CSR copy)
    const unsigned char *perm;  // sorted position -> local row of every tile
    const long long *tbase;     // first entry of every tile in sv / so (num_tiles + 1)
    const int *sbase;           // first entry of every 32-row slice, relative to its tile
    int ring;                   // R: entries of the ring buffer, a power of two and a multiple of WIN_T
    int w;                      // W: half-width of the window, a multiple of WIN_T
    int x_len;                  // entries of x that may be staged (whole 16-byte units of the owned rows)
    int tiles_per_cta;
    const unsigned short *smeta; // streaming kernel: local row | length << 8 of every sorted position
    const int *slens;            // streaming kernel: entries per row of every slice
    const int *row_ptr;
};

// plan 1: how many entries two candidate windows would hold
__global__ void win_stats_kernel(const int *__restrict__ rp, const int *__restrict__ ci, int n, int w1, int w2, int lmax, unsigned long long *stats)
{
    unsigned long long in1 = 0, in2 = 0, reg = 0, nlong = 0;
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < n; row += gridDim.x * blockDim.x) {
        const int k0 = rp[row], k1 = rp[row + 1];
        if (k1 - k0 > lmax) { nlong++; continue; }          // long rows go to the warp-per-row side kernel: not part of the window's business
        reg += (unsigned long long)(k1 - k0);
        for (int k = k0; k < k1; k++) {
            const int d = ci[k] - row;
            const int ad = d < 0 ? -d : d;
            in1 += ad <= w1;
            in2 += ad <= w2;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        in1 += __shfl_xor_sync(0xffffffffu, in1, o);
        in2 += __shfl_xor_sync(0xffffffffu, in2, o);
        reg += __shfl_xor_sync(0xffffffffu, reg, o);
        nlong += __shfl_xor_sync(0xffffffffu, nlong, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(stats + 0, in1);
        atomicAdd(stats + 1, in2);
        atomicAdd(stats + 2, reg);
        atomicAdd(stats + 3, nlong);
    }
}

// plan 2: per tile, the rows sorted by length (descending, ties by index) -> sorted position p is handled by thread p; the 32 rows of a
// slice (= a consumer warp) then have nearly equal lengths, and a slice is stored entry-major: entry j of lane l at slice + 32 j + l.
// A slice is as long as its first row; the few positions past a shorter row's end are padding that is never read.
__global__ void __launch_bounds__(WIN_T) sell_plan_kernel(const int *__restrict__ rp, int n, int num_tiles, int lmax, unsigned char *perm, int *sbase, int *tlen, unsigned short *smeta, int *slens,
                                                          int *long_rows, int *long_count)
{
    __shared__ int len[WIN_T];
    __shared__ int slen[WIN_T / 32];
    const int tid = threadIdx.x;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int row = t * WIN_T + tid;
        const int true_len = row < n ? rp[row + 1] - rp[row] : -1;
        const bool is_long = true_len > lmax;                          // summed by the side kernel (csr_long_rows_kernel); stored with length 0 here
        const int mine = is_long ? 0 : true_len;                      // rows past the end (-1) sort last
        if (is_long) long_rows[atomicAdd(long_count, 1)] = row;
        len[tid] = mine;
        __syncthreads();
        int rank = 0;
        for (int j = 0; j < WIN_T; j++) rank += (len[j] > mine) || (len[j] == mine && j < tid);
        perm[(size_t)t * WIN_T + rank] = (unsigned char)tid;
        smeta[(size_t)t * WIN_T + rank] = (unsigned short)(tid | ((is_long ? 255 : max(mine, 0)) << 8));      // local row and length of sorted position `rank`; 255 = long row
        if ((rank & 31) == 0) slen[rank >> 5] = max(mine, 0);
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int sl = 0; sl < WIN_T / 32; sl++) { sbase[(size_t)t * (WIN_T / 32) + sl] = run; slens[(size_t)t * (WIN_T / 32) + sl] = slen[sl]; run += 32 * slen[sl]; }
            tlen[t] = run;
        }
        __syncthreads();
    }
}

// plan 3 (and after every in-place change of the values, so == nullptr): CSR -> sliced-ELL copy
template <class MatT>
__global__ void __launch_bounds__(WIN_T) sell_fill_kernel(const int *__restrict__ rp, const int *__restrict__ ci, const MatT *__restrict__ val, int n, int num_tiles,
                                                          const unsigned char *__restrict__ perm, const unsigned short *__restrict__ smeta, const long long *__restrict__ tbase, const int *__restrict__ sbase, MatT *sv,
                                                          unsigned short *so)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int row = t * WIN_T + (int)perm[(size_t)t * WIN_T + tid];
        const int k0 = row < n ? rp[row] : 0;
        const int lenb = (int)(smeta[(size_t)t * WIN_T + tid] >> 8);
        const int len = (row < n && lenb != 255) ? lenb : 0;            // long rows are not stored
        const int L = __shfl_sync(0xffffffffu, len, 0);
        const long long base = tbase[t] + sbase[(size_t)t * (WIN_T / 32) + warp] + lane;
        for (int j = 0; j < L; j++) {
            const long long dst = base + (long long)j * 32;
            if (j < len) {
                sv[dst] = val[k0 + j];
                if (so) {
                    const int d = ci[k0 + j] - row;
                    so[dst] = (d > -WIN_BIAS && d < WIN_BIAS) ? (unsigned short)(d + WIN_BIAS) : WIN_FAR;
                }
            } else {
                sv[dst] = (MatT)0;
                if (so) so[dst] = 0;
            }
        }
    }
}

// up to N entries of one row, `rem` of them present: offsets -> columns, x from the ring, FMAs in storage order.  vals / offs point at the
// lane's first entry of the step, consecutive entries of a row are 32 apart.  The common case (every column inside the window) is
// straight-line code: N offset / value loads, N ring loads, N FMAs; a thread with a column outside the window (or a 16-bit overflow:
// WIN_FAR lands below the window by construction) takes the per-entry path with the global gather.
template <class MatT, class VecT, int N, bool FULL>
__device__ __forceinline__ VecT row_sell_step(const MatT *__restrict__ vals, const unsigned short *__restrict__ offs, const int *__restrict__ col, const int *__restrict__ rpp, const int j0, const int rem,
                                              const VecT *__restrict__ ring, const unsigned mask, const int lo, const unsigned span, const int row, const VecT *__restrict__ x, VecT sum)
{
    int c[N];
    MatT v[N];
    VecT xv[N];
    bool inside = true;
#pragma unroll
    for (int u = 0; u < N; u++)
        if (FULL || u < rem) {
            c[u] = row + (int)offs[u * 32];      // row arrives with the bias already subtracted
            v[u] = vals[u * 32];
            inside = inside && ((unsigned)(c[u] - lo) < span);
        }
    if (inside) {
#pragma unroll
        for (int u = 0; u < N; u++)
            if (FULL || u < rem) xv[u] = ring[(unsigned)c[u] & mask];
    } else {
#pragma unroll
        for (int u = 0; u < N; u++)
            if (FULL || u < rem) {
                if ((unsigned)(c[u] - lo) < span) xv[u] = ring[(unsigned)c[u] & mask];
                else {
                    if (offs[u * 32] == WIN_FAR) c[u] = __ldg(col + (__ldg(rpp) + j0 + u));      // rare: the CSR copy's 32-bit column
                    xv[u] = __ldg(x + c[u]);
                }
            }
    }
#pragma unroll
    for (int u = 0; u < N; u++)
        if (FULL || u < rem) sum = fma((VecT)v[u], xv[u], sum);
    return sum;
}

// Rows longer than the plan's lmax (hub rows of aggregated levels: 2 % of the rows, a third of the entries on level 1 of the banded
// hierarchy) would force whole slices to their length.  They are left out of the sliced-ELL copy and summed here, a warp per row over the CSR
// arrays, in exactly the order of csr_vector_kernel (lane-strided FMAs, xor-shuffle tree) -- the kernel those levels ran before, so these
// rows keep their bits.  The dot product is parked in y[row]; the window kernel picks it up and applies the epilogue / reduction as for
// any other row.  (The caller guarantees y aliases neither b nor x.)
template <class MatT, class VecT>
__global__ void __launch_bounds__(256) csr_long_rows_kernel(const int *__restrict__ row_ptr, const int *__restrict__ col, const MatT *__restrict__ val, const VecT *__restrict__ x,
                                                            VecT *__restrict__ y, const int *__restrict__ long_rows, int num_long)
{
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < num_long; i += gridDim.x * wpb) {
        const int row = __ldg(long_rows + i);
        const int k0 = __ldg(row_ptr + row), k1 = __ldg(row_ptr + row + 1);
        VecT sum = 0;
        for (int k = k0 + lane; k < k1; k += 32) sum = fma((VecT)__ldg(val + k), __ldg(x + __ldg(col + k)), sum);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) y[row] = sum;
    }
}

// blockDim.x = WIN_T + 32 (last warp = producer), one CTA per SM.
// smem: barriers + reduction scratch (512 B) | ring[R] | stages x { values cap | offsets cap | row map T x 2 B | slice offsets | header 16 B }
template <class MatT, class VecT, int EPI>
__global__ void __launch_bounds__(WIN_T + PRODUCER_THREADS, 1) csr_window_kernel(const TileArgs<MatT, VecT> a, const WinArgs w)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + MAX_STAGES;
    double *smem_red = reinterpret_cast<double *>(smem_raw + 2 * MAX_STAGES * sizeof(uint64_t));
    VecT *ring = reinterpret_cast<VecT *>(smem_raw + 512);
    unsigned char *stage_base = smem_raw + 512 + (size_t)w.ring * sizeof(VecT);
    const size_t vals_bytes = align16((size_t)a.cap * sizeof(MatT));
    const size_t offs_bytes = align16((size_t)a.cap * sizeof(short));
    constexpr int CONSUMER_WARPS = WIN_T / 32;
    const size_t meta_off = vals_bytes + offs_bytes;                       // the tile's row map (local row | length << 8 per sorted position) and slice offsets
    const size_t sbase_off = meta_off + WIN_T * sizeof(unsigned short);   // travel with it: nothing the consumers need before their first FMA comes from global memory
    const size_t hdr_off = sbase_off + CONSUMER_WARPS * sizeof(int);
    const size_t stage_bytes = hdr_off + 16;
    constexpr bool HAS_RED = (EPI == EPI_SPMV_DOT || EPI == EPI_JACOBI_DOT || EPI == EPI_RESID_NRM2);
    constexpr bool NEED_B = (EPI == EPI_RESID || EPI == EPI_JACOBI || EPI == EPI_JACOBI_DOT || EPI == EPI_RESID_NRM2 || EPI == EPI_JACOBI_L1 || EPI == EPI_ADD);
    constexpr bool NEED_D = (EPI == EPI_JACOBI || EPI == EPI_JACOBI_DOT || EPI == EPI_JACOBI_L1);

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < a.stages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    double acc = 0.0;
    const int t_begin = (int)blockIdx.x * w.tiles_per_cta;
    const int my_tiles = max(0, min(a.num_tiles, t_begin + w.tiles_per_cta) - t_begin);
    const unsigned mask = (unsigned)w.ring - 1u;

    if (tid >= WIN_T) {
        // ------------------------------- producer warp -------------------------------
        if (tid == WIN_T) {
            int s = 0;
            unsigned ph = 0u;
            for (int it = 0; it < my_tiles; it++, s = (s + 1 == a.stages) ? 0 : s + 1, ph ^= (s == 0) ? 1u : 0u) {
                const int tile = t_begin + it;
                if (it >= a.stages) mbar_wait(&empty[s], ph ^ 1u);
                const int r0 = tile * WIN_T;
                const int r1 = min(r0 + WIN_T, a.n);
                const long long tb0 = __ldg(w.tbase + tile);
                const unsigned cnt = (unsigned)(__ldg(w.tbase + tile + 1) - tb0);       // a multiple of 32 entries
                unsigned char *st = stage_base + (size_t)s * stage_bytes;
                const unsigned val_copy = cnt * (unsigned)sizeof(MatT), off_copy = cnt * (unsigned)sizeof(short);
                // the part of x this tile adds to the ring: everything for the CTA's first tile, the T entries at the far end afterwards
                const int win_lo = max(0, r0 - w.w), win_hi = min(w.x_len, r0 + WIN_T + w.w);
                const int x0 = (it == 0) ? win_lo : min(w.x_len, r0 + w.w);
                const unsigned x_copy = win_hi > x0 ? (unsigned)(win_hi - x0) * (unsigned)sizeof(VecT) : 0u;
                *reinterpret_cast<int4 *>(st + hdr_off) = make_int4(win_lo, win_hi > win_lo ? win_hi - win_lo : 0, 0, 0);   // released to the consumers by the arrive below
                mbar_expect_tx(&full[s], val_copy + off_copy + x_copy + (unsigned)(WIN_T * sizeof(unsigned short)) + (unsigned)(CONSUMER_WARPS * sizeof(int)));
                tma_bulk_g2s(st + meta_off, w.smeta + (size_t)tile * WIN_T, (unsigned)(WIN_T * sizeof(unsigned short)), &full[s]);
                tma_bulk_g2s(st + sbase_off, w.sbase + (size_t)tile * CONSUMER_WARPS, (unsigned)(CONSUMER_WARPS * sizeof(int)), &full[s]);
                if (cnt) {
                    tma_bulk_g2s(st, reinterpret_cast<const MatT *>(w.sv) + tb0, val_copy, &full[s]);
                    tma_bulk_g2s(st + vals_bytes, w.so + tb0, off_copy, &full[s]);
                }
                for (int p = x0; p < win_hi;) {                           // pieces that do not cross the end of the ring (one, or two for the first tile)
                    const int e = min(win_hi, (int)(((unsigned)p | mask) + 1u));
                    tma_bulk_g2s(ring + ((unsigned)p & mask), a.x + p, (unsigned)(e - p) * (unsigned)sizeof(VecT), &full[s]);
                    p = e;
                }
                if (a.l2pf) {
                    if (NEED_B) l2_prefetch_span(a.b + r0, r1 - r0);
                    if (NEED_D) l2_prefetch_span(a.d + r0, r1 - r0);
                }
            }
        }
    } else {
        // ------------------------------- consumers: one row per thread, a warp = one slice -------------------------------
        const int lane = tid & 31, warp = tid >> 5;
        int s = 0;
        unsigned ph = 0u;
        for (int it = 0; it < my_tiles; it++, s = (s + 1 == a.stages) ? 0 : s + 1, ph ^= (s == 0) ? 1u : 0u) {
            const int tile = t_begin + it;
            const unsigned char *st = stage_base + (size_t)s * stage_bytes;
            mbar_wait(&full[s], ph);
            const int4 hdr = *reinterpret_cast<const int4 *>(st + hdr_off);
            const unsigned meta = reinterpret_cast<const unsigned short *>(st + meta_off)[tid];
            const int lrow = (int)(meta & 255u);
            const bool is_long = (meta >> 8) == 255u;                    // its dot product was left in y[row] by csr_long_rows_kernel
            const int sb = reinterpret_cast<const int *>(st + sbase_off)[warp];
            const int row = tile * WIN_T + lrow;
            const bool active = row < a.n;
            // the row's vector operands: needed by the epilogue only, their latency (L2: the producer asked for these lines when it issued
            // the tile) hides behind the row's dot product
            VecT bi = 0, xi = 0;
            MatT di = 1;
            if (active) {
                if (NEED_B) bi = __ldg(a.b + row);
                if (NEED_D) di = __ldg(a.d + row);
                if (NEED_D || EPI == EPI_SPMV_DOT) xi = ((unsigned)(row - hdr.x) < (unsigned)hdr.y) ? ring[(unsigned)row & mask] : __ldg(a.x + row);      // the row's own x sits in the window
            }
            const MatT *vals = reinterpret_cast<const MatT *>(st) + sb + lane;
            const unsigned short *offs = reinterpret_cast<const unsigned short *>(st + vals_bytes) + sb + lane;
            const int len = (active && !is_long) ? (int)(meta >> 8) : 0;
            const int L = __shfl_sync(0xffffffffu, len, 0);              // the slice's first row is its longest
            VecT sum = (active && is_long) ? a.y[row] : (VecT)0;
            const int rowb = row - WIN_BIAS;
            auto step8 = [&](const int j) {
                const int rem = len - j;
                if (rem >= 8) sum = row_sell_step<MatT, VecT, 8, true>(vals + j * 32, offs + j * 32, a.col, a.row_ptr + row, j, rem, ring, mask, hdr.x, (unsigned)hdr.y, rowb, a.x, sum);
                else if (rem > 0) sum = row_sell_step<MatT, VecT, 8, false>(vals + j * 32, offs + j * 32, a.col, a.row_ptr + row, j, rem, ring, mask, hdr.x, (unsigned)hdr.y, rowb, a.x, sum);
            };
            int j = 0;
            for (; j < L; j += 8) step8(j);
            if (active) acc += tile_epilogue<MatT, VecT, EPI>(a, row, sum, bi, di, xi);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
    }
    if (HAS_RED) block_reduce_finish(acc, smem_red, a.red, a.fin_op, a.fin_slot, a.mirror);
}

// ---------------------------------------------------------------------------------------------
// Streaming form (AMGXB_WINDOW_STREAM=1, opt-in: measured SLOWER than the whole-tile form, kept as the record of the experiment): the same
// ring, the same sliced-ELL data, but no whole-tile stages.  Motivation: with whole tiles the 8 consumer warps meet at every tile -- the
// warp holding the tile's longest slice (28 entries per row against 17 on average) sets the pace, and the next 48 KB load starts only
// when the slowest warp is done (ncu r02: consumers waiting for data 35-40 % of the time at 56 % DRAM throughput).  Here every consumer
// warp has its own queue of SQ small buffers; a slice travels as chunks of <= 8 entry-rows (2.5 KB), and slice s of tile t goes to warp
// (s + t) mod 8, so long and short slices rotate over the warps.  A warp refills its own queue: after consuming a chunk its lane 0
// issues the bulk copy of the chunk SQ ahead into the same buffer (first attempt: 8 lanes of a producer warp, one per queue -- UBLKCP
// from divergent lanes is serialised by the uniform datapath and the producer became the bottleneck: 0.35 ms).  A separate thread feeds
// the ring; the only coupling left is the ring: x of tile t may be overwritten once all 8 warps are done with tile t - SK + 1 (rempty,
// count 8), i.e. SK * T + 2 W <= R.  Result on the 4 M-row banded matrix: SpMV 0.208 ms against 0.167 ms for the whole-tile form -- the
// data waits shrink to 13 % of the samples, but the per-chunk bookkeeping (76 M warp instructions against 45 M) costs more than they
// did with only two warps per scheduler to hide it.
// ---------------------------------------------------------------------------------------------
constexpr int SQ = 4;        // chunk buffers per consumer warp
constexpr int SCH = 8;       // entry-rows per chunk
constexpr int SK = 2;        // tiles whose part of x may be in use at the same time

template <class MatT, class VecT, int EPI>
__global__ void __launch_bounds__(WIN_T + PRODUCER_THREADS, 1) csr_stream_kernel(const TileArgs<MatT, VecT> a, const WinArgs w)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int CONSUMER_WARPS = WIN_T / 32;
    constexpr unsigned CHUNK_VALS = SCH * 32 * sizeof(MatT), CHUNK_BYTES = CHUNK_VALS + SCH * 32 * sizeof(unsigned short);
    uint64_t *cfull = reinterpret_cast<uint64_t *>(smem_raw);                 // [warp][slot]
    uint64_t *rfull = cfull + CONSUMER_WARPS * SQ;                             // [SK]
    uint64_t *rempty = rfull + SK;
    double *smem_red = reinterpret_cast<double *>(smem_raw + 640);
    VecT *ring = reinterpret_cast<VecT *>(smem_raw + 1024);
    unsigned char *chunk_base = smem_raw + 1024 + (size_t)w.ring * sizeof(VecT);
    constexpr bool HAS_RED = (EPI == EPI_SPMV_DOT || EPI == EPI_JACOBI_DOT || EPI == EPI_RESID_NRM2);
    constexpr bool NEED_B = (EPI == EPI_RESID || EPI == EPI_JACOBI || EPI == EPI_JACOBI_DOT || EPI == EPI_RESID_NRM2 || EPI == EPI_JACOBI_L1 || EPI == EPI_ADD);
    constexpr bool NEED_D = (EPI == EPI_JACOBI || EPI == EPI_JACOBI_DOT || EPI == EPI_JACOBI_L1);

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < CONSUMER_WARPS * SQ; i++) {
            mbar_init(&cfull[i], 1);
        }
        for (int i = 0; i < SK; i++) {
            mbar_init(&rfull[i], 1);
            mbar_init(&rempty[i], CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    double acc = 0.0;
    const int t_begin = (int)blockIdx.x * w.tiles_per_cta;
    const int my_tiles = max(0, min(a.num_tiles, t_begin + w.tiles_per_cta) - t_begin);
    const unsigned mask = (unsigned)w.ring - 1u;

    if (tid >= WIN_T) {
        // ------------------------------- ring warp: one thread keeps x[r0 - W, r0 + SK * T + W) of the CTA's tile range in the ring -------------------------------
        if (tid == WIN_T) {
            if (my_tiles > 0) {            // the CTA's slice descriptors and row maps: small, read by every warp a little ahead of use -- pull them into L2 now
                l2_prefetch_span(w.tbase + t_begin, my_tiles + 1);
                l2_prefetch_span(w.sbase + (size_t)t_begin * CONSUMER_WARPS, my_tiles * CONSUMER_WARPS);
                l2_prefetch_span(w.slens + (size_t)t_begin * CONSUMER_WARPS, my_tiles * CONSUMER_WARPS);
                l2_prefetch_span(w.smeta + (size_t)t_begin * WIN_T, my_tiles * WIN_T);
            }
            for (int it = 0; it < my_tiles; it++) {
                const int slot = it & (SK - 1);
                if (it >= SK) mbar_wait(&rempty[slot], ((unsigned)(it / SK) - 1u) & 1u);
                const int tile = t_begin + it;
                const int r0 = tile * WIN_T, r1 = min(r0 + WIN_T, a.n);
                const int win_lo = max(0, r0 - w.w), win_hi = min(w.x_len, r0 + WIN_T + w.w);
                const int x0 = (it == 0) ? win_lo : min(w.x_len, r0 + w.w);
                const unsigned x_copy = win_hi > x0 ? (unsigned)(win_hi - x0) * (unsigned)sizeof(VecT) : 0u;
                mbar_expect_tx(&rfull[slot], x_copy);
                for (int p = x0; p < win_hi;) {
                    const int e = min(win_hi, (int)(((unsigned)p | mask) + 1u));
                    tma_bulk_g2s(ring + ((unsigned)p & mask), a.x + p, (unsigned)(e - p) * (unsigned)sizeof(VecT), &rfull[slot]);
                    p = e;
                }
                if (a.l2pf) {
                    if (NEED_B) l2_prefetch_span(a.b + r0, r1 - r0);
                    if (NEED_D) l2_prefetch_span(a.d + r0, r1 - r0);
                }
            }
        }
    } else {
        // ------------------------------- consumers: a warp = one slice at a time; it refills its own queue -------------------------------
        const int lane = tid & 31, wq = tid >> 5;
        // load cursor (warp-uniform): the chunk that goes into the buffer the warp frees next.  Chunks of a slice are contiguous, so the cursor
        // is two running pointers; the descriptors of the next two slices are already in registers (their lines were pulled into L2 by the
        // ring thread when the kernel started).
        const MatT *p_v = nullptr;
        const unsigned short *p_o = nullptr;
        int p_it = -1, p_rem = 0;          // tile of the cursor, entry-rows left in its slice
        int n1_L = 0, n2_L = 0;
        long long n1_base = 0, n2_base = 0;
        unsigned gi = 0;                   // chunks issued
        auto load_desc = [&](const int itn, long long &base, int &L) {
            base = 0;
            L = 0;
            if (itn < my_tiles) {
                const int tn = t_begin + itn, sn = (wq - itn) & 7;
                base = __ldg(w.tbase + tn) + __ldg(w.sbase + (size_t)tn * CONSUMER_WARPS + sn);
                L = __ldg(w.slens + (size_t)tn * CONSUMER_WARPS + sn);
            }
        };
        load_desc(0, n1_base, n1_L);
        load_desc(1, n2_base, n2_L);
        auto issue = [&]() {
            while (p_rem <= 0 && p_it < my_tiles) {        // slice exhausted (or empty): on to the next tile's slice
                p_it++;
                p_v = reinterpret_cast<const MatT *>(w.sv) + n1_base;
                p_o = w.so + n1_base;
                p_rem = n1_L;
                n1_base = n2_base;
                n1_L = n2_L;
                load_desc(p_it + 2, n2_base, n2_L);
            }
            if (p_it >= my_tiles) return;
            if (lane == 0) {
                const unsigned q = gi & (SQ - 1);
                const unsigned ne = (unsigned)min(SCH, p_rem);
                unsigned char *buf = chunk_base + (size_t)(wq * SQ + q) * CHUNK_BYTES;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the warp's reads of this buffer precede the bulk copy that refills it
                mbar_expect_tx(&cfull[wq * SQ + q], ne * 32u * (unsigned)(sizeof(MatT) + sizeof(unsigned short)));
                tma_bulk_g2s(buf, p_v, ne * 32u * (unsigned)sizeof(MatT), &cfull[wq * SQ + q]);
                tma_bulk_g2s(buf + CHUNK_VALS, p_o, ne * 32u * (unsigned)sizeof(unsigned short), &cfull[wq * SQ + q]);
            }
            p_v += SCH * 32;
            p_o += SCH * 32;
            p_rem -= SCH;
            gi++;
        };
        for (int i = 0; i < SQ; i++) issue();
        unsigned g = 0;
        unsigned meta_next = my_tiles > 0 ? (unsigned)__ldg(w.smeta + (size_t)t_begin * WIN_T + (wq & 7) * 32 + lane) : 0u;
        for (int it = 0; it < my_tiles; it++) {
            const int tile = t_begin + it;
            const unsigned meta = meta_next;
            if (it + 1 < my_tiles) meta_next = (unsigned)__ldg(w.smeta + (size_t)(tile + 1) * WIN_T + ((wq - it - 1) & 7) * 32 + lane);
            const int lrow = (int)(meta & 255u);
            const bool is_long = (meta >> 8) == 255u;                    // summed by csr_long_rows_kernel, parked in y[row]
            const int len = is_long ? 0 : (int)(meta >> 8);
            const int row = tile * WIN_T + lrow;
            const bool active = row < a.n;
            VecT bi = 0, xi = 0;
            MatT di = 1;
            if (active) {
                if (NEED_B) bi = __ldg(a.b + row);
                if (NEED_D) di = __ldg(a.d + row);
            }
            const int L = __shfl_sync(0xffffffffu, len, 0);
            const int r0 = tile * WIN_T;
            const int win_lo = max(0, r0 - w.w), win_hi = min(w.x_len, r0 + WIN_T + w.w);
            const unsigned span = win_hi > win_lo ? (unsigned)(win_hi - win_lo) : 0u;
            const int slot = it & (SK - 1);
            mbar_wait(&rfull[slot], (unsigned)(it / SK) & 1u);
            if (active && (NEED_D || EPI == EPI_SPMV_DOT)) xi = ((unsigned)(row - win_lo) < span) ? ring[(unsigned)row & mask] : __ldg(a.x + row);
            const int rowb = row - WIN_BIAS;
            VecT sum = (active && is_long) ? a.y[row] : (VecT)0;
            for (int c0 = 0; c0 < L; c0 += SCH, g++) {
                const unsigned q = g & (SQ - 1);
                mbar_wait(&cfull[wq * SQ + q], (g / SQ) & 1u);
                const unsigned char *buf = chunk_base + (size_t)(wq * SQ + q) * CHUNK_BYTES;
                const MatT *vals = reinterpret_cast<const MatT *>(buf) + lane;
                const unsigned short *offs = reinterpret_cast<const unsigned short *>(buf + CHUNK_VALS) + lane;
                const int rem = len - c0;
                if (rem >= SCH) sum = row_sell_step<MatT, VecT, SCH, true>(vals, offs, a.col, w.row_ptr + row, c0, rem, ring, mask, win_lo, span, rowb, a.x, sum);
                else if (rem > 0) sum = row_sell_step<MatT, VecT, SCH, false>(vals, offs, a.col, w.row_ptr + row, c0, rem, ring, mask, win_lo, span, rowb, a.x, sum);
                __syncwarp();
                issue();                   // refill the buffer just consumed with the chunk SQ ahead
            }
            if (active) acc += tile_epilogue<MatT, VecT, EPI>(a, row, sum, bi, di, xi);
            __syncwarp();
            if (lane == 0) mbar_arrive(&rempty[slot]);
        }
    }
    if (HAS_RED) block_reduce_finish(acc, smem_red, a.red, a.fin_op, a.fin_slot, a.mirror);
}

template <class MatT, class VecT, int EPI> void launch_win(const Matrix &A, const TileArgs<MatT, VecT> &ta, const WinArgs &wa, cudaStream_t s)
{
    if (A.win.num_long > 0) {
        const int sms = A.rsc ? A.rsc->num_sms : B200_SMS;
        csr_long_rows_kernel<MatT, VecT><<<std::max(1, std::min(ceil_div(A.win.num_long, 8), sms * 8)), 256, 0, s>>>(ta.row_ptr, ta.col, ta.val, ta.x, ta.y, A.win.long_rows.ptr(), A.win.num_long);
        count_launch();
    }
    if (A.win.stream) {
        const size_t smem = A.win.stream_smem_bytes;
        auto k = csr_stream_kernel<MatT, VecT, EPI>;
        smem_opt_in(reinterpret_cast<const void *>(k), smem);      // exactly what this kernel needs, once per size and device
        k<<<A.win.grid, WIN_T + PRODUCER_THREADS, smem, s>>>(ta, wa);
    } else {
        const size_t smem = A.win.smem_bytes;
        auto k = csr_window_kernel<MatT, VecT, EPI>;
        smem_opt_in(reinterpret_cast<const void *>(k), smem);      // exactly what this kernel needs, once per size and device
        k<<<A.win.grid, WIN_T + PRODUCER_THREADS, smem, s>>>(ta, wa);
    }
    count_launch();
    AMGXB_LAUNCH_CHECK();
}

}  // namespace

static void sell_fill(Matrix &A, bool with_offsets, cudaStream_t s)
{
    WinPlan &P = A.win;
    const int sms = A.rsc ? A.rsc->num_sms : B200_SMS;
    const int grid = std::max(1, std::min(A.plan.num_tiles, sms * 8));
    unsigned short *so = with_offsets ? P.so.ptr() : nullptr;
    if (A.mat_prec == Prec::F64)
        sell_fill_kernel<double><<<grid, WIN_T, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), A.values.as<double>(), A.n, A.plan.num_tiles, P.perm.ptr(), P.smeta.ptr(), P.tbase.ptr(),
                                                        P.sbase.ptr(), (double *)P.sv.ptr(), so);
    else
        sell_fill_kernel<float><<<grid, WIN_T, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), A.values.as<float>(), A.n, A.plan.num_tiles, P.perm.ptr(), P.smeta.ptr(), P.tbase.ptr(),
                                                       P.sbase.ptr(), (float *)P.sv.ptr(), so);
    count_launch();
    AMGXB_LAUNCH_CHECK();
}

// called at the end of csr_build_plan (after the coded streams: a level they pair-code is a stencil level and stays with them)
void csr_build_window(Matrix &A, cudaStream_t s)
{
    WinPlan &P = A.win;
    P.on = false;
    static const int env_on = getenv("AMGXB_WINDOW") ? atoi(getenv("AMGXB_WINDOW")) : 1;
    static const double env_min = getenv("AMGXB_WINDOW_MIN_INSIDE") ? atof(getenv("AMGXB_WINDOW_MIN_INSIDE")) : 0.97;     // a column outside the window costs a global gather inside the FMA chain: r02, 7 % outside = 2.8x slower
    static const int env_ring = getenv("AMGXB_WINDOW_RING") ? atoi(getenv("AMGXB_WINDOW_RING")) : 0;      // experiments: force 16384 (2 stages) or 8192 (3 stages)
    const int sms = A.rsc ? A.rsc->num_sms : B200_SMS;
    // (a level whose plain tile plan failed because a few hub rows blow up a tile -- plan.use_tiles false -- is exactly what the long-row split is for)
    if (!env_on || A.bs() != 1 || A.plan.tile_rows != WIN_T || A.plan.split != 0 || A.plan.use_perm) return;
    if (A.n_cols > A.n) return;                                    // halo columns live behind the owned rows: the plain / coded kernels
    const int nt = A.plan.num_tiles;
    if (nt < sms * 8) return;                                      // a CTA's first tile loads the whole window: needs a range of tiles to pay for it
    if ((double)A.nnz < 8.0 * (double)A.n) return;                 // short rows: the vectors dominate, nothing to gain
    if (A.colenc.on && 2 * A.colenc.tiles_pair > A.colenc.num_tiles) return;
    const size_t msz = prec_size(A.mat_prec), vsz = prec_size(A.vec_prec);
    struct Cand { int ring, stages, w; } cand[2] = {{16384, 2, 0}, {8192, 3, 0}};
    for (Cand &c : cand) c.w = ((c.ring - c.stages * WIN_T) / 2 / WIN_T) * WIN_T;
    // rows longer than lmax leave the sliced-ELL copy (one of them would stretch its whole slice): 3 x the mean, at least 48 and at most what
    // 8 bits hold -- no row of the 4 M-row banded matrix (mean 16, longest 34), the hub rows of its first aggregated level (mean 23; 2.4 % of
    // the rows hold 330-2600 entries each, a third of all entries)
    static const int env_lmax = getenv("AMGXB_WINDOW_LMAX") ? atoi(getenv("AMGXB_WINDOW_LMAX")) : 0;
    const int lmax = env_lmax > 0 ? std::min(env_lmax, 254) : std::max(48, std::min(254, (int)(3.0 * (double)A.nnz / (double)std::max(A.n, 1))));
    DevBuf<unsigned long long> stats;
    stats.resize(4);
    stats.zero(s);
    win_stats_kernel<<<std::min(ceil_div(A.n, 256), sms * 16), 256, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), A.n, cand[0].w, cand[1].w, lmax, stats.ptr());
    count_launch();
    AMGXB_LAUNCH_CHECK();
    const std::vector<unsigned long long> h = stats.to_host(s);
    const long long nnz_reg = (long long)h[2];
    const int num_long = (int)h[3];
    if (nnz_reg < 4 * (long long)A.n) return;                     // what is left after the long rows is too thin to be worth a second copy
    const double inside[2] = {(double)h[0] / (double)std::max(nnz_reg, 1LL), (double)h[1] / (double)std::max(nnz_reg, 1LL)};
    const bool verbose = getenv("AMGXB_WINDOW_VERBOSE") != nullptr;
    if (std::max(inside[0], inside[1]) < env_min) {
        if (verbose) fprintf(stderr, "[amgx_b200] window level %d: %d rows, inside +-%d: %.3f, +-%d: %.3f -> off\n", A.level, A.n, cand[0].w, inside[0], cand[1].w, inside[1]);
        return;
    }
    // sliced-ELL plan: sorted rows, slice and tile offsets
    P.perm.resize((size_t)nt * WIN_T);
    P.sbase.resize((size_t)nt * (WIN_T / 32));
    DevBuf<int> tlen;
    tlen.resize((size_t)nt);
    P.smeta.resize((size_t)nt * WIN_T);
    P.slens.resize((size_t)nt * (WIN_T / 32));
    P.long_rows.resize((size_t)std::max(num_long, 1));
    DevBuf<int> long_count;
    long_count.resize(1);
    long_count.zero(s);
    sell_plan_kernel<<<std::max(1, std::min(nt, sms * 8)), WIN_T, 0, s>>>(A.row_ptr.ptr(), A.n, nt, lmax, P.perm.ptr(), P.sbase.ptr(), tlen.ptr(), P.smeta.ptr(), P.slens.ptr(), P.long_rows.ptr(),
                                                                          long_count.ptr());
    count_launch();
    AMGXB_LAUNCH_CHECK();
    const std::vector<int> hl = tlen.to_host(s);
    std::vector<long long> hb((size_t)nt + 1, 0);
    int cap = 32;
    for (int t = 0; t < nt; t++) { hb[t + 1] = hb[t] + hl[t]; cap = std::max(cap, hl[t]); }
    const size_t stage = align16((size_t)cap * msz) + align16((size_t)cap * sizeof(short)) + WIN_T * sizeof(unsigned short) + (WIN_T / 32) * sizeof(int) + 16;
    int best = -1;
    for (int i = 0; i < 2; i++) {
        const size_t smem = 512 + (size_t)cand[i].ring * vsz + (size_t)cand[i].stages * stage;
        if (smem > (size_t)226 * 1024 || inside[i] < env_min) continue;
        if (env_ring > 0 && cand[i].ring != env_ring) continue;
        if (best < 0 || inside[i] > inside[best] + 0.02) best = i;       // the wider window when it fits; the narrower one must hold clearly more to win
    }
    if (verbose)
        fprintf(stderr, "[amgx_b200] window level %d: %d rows, %.1f entries per row; %d rows longer than %d (%.1f %% of the entries) go to the warp-per-row kernel; sliced-ELL padding %.1f %%, "
                        "longest tile %d; inside +-%d: %.3f, +-%d: %.3f -> %s\n", A.level, A.n, (double)A.nnz / A.n, num_long, lmax, 100.0 * (1.0 - (double)nnz_reg / std::max(A.nnz, 1)),
                100.0 * ((double)hb[nt] / (double)std::max(nnz_reg, 1LL) - 1.0), cap, cand[0].w, inside[0], cand[1].w, inside[1],
                best < 0 ? "off (shared memory)" : (best == 0 ? "ring 16384 x 2 stages" : "ring 8192 x 3 stages"));
    if (best < 0 || (double)hb[nt] > 1.25 * (double)nnz_reg) { P.perm.release(); P.sbase.release(); P.smeta.release(); P.slens.release(); P.long_rows.release(); return; }
    P.num_long = num_long;
    P.lmax = lmax;
    P.tbase.from_any(hb.data(), hb.size(), s);
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));                     // hb is a local
    P.sv.resize((size_t)hb[nt] * msz + 64);
    P.so.resize((size_t)hb[nt] + 64);
    P.cap = cap;
    P.ring = cand[best].ring;
    P.w = cand[best].w;
    P.stages = cand[best].stages;
    P.smem_bytes = 512 + (size_t)P.ring * vsz + (size_t)P.stages * stage;
    P.inside = inside[best];
    static const int env_stream = getenv("AMGXB_WINDOW_STREAM") ? atoi(getenv("AMGXB_WINDOW_STREAM")) : 0;      // opt-in, see the streaming form's header comment
    P.stream_smem_bytes = 1024 + (size_t)P.ring * vsz + (size_t)(WIN_T / 32) * SQ * SCH * 32 * (msz + sizeof(unsigned short));
    // streaming form: lengths travel in 8 bits, SK tiles of x in use at once need SK * T + 2 W <= R (the plan's W satisfies it for stages >= SK)
    P.stream = env_stream != 0 && P.stages >= SK && P.stream_smem_bytes <= (size_t)226 * 1024;
    sell_fill(A, true, s);
    const int ctas = std::min(sms, nt);
    P.tiles_per_cta = ceil_div(nt, ctas);
    P.grid = ceil_div(nt, P.tiles_per_cta);
    P.on = true;
}

// the values of A changed in place: the sliced-ELL copy follows them (called from csr_values_changed)
void csr_window_values_changed(Matrix &A, cudaStream_t s)
{
    if (A.win.on) sell_fill(A, false, s);
}

// csr_op entry of the window path; false: the caller goes on to the coded / plain kernels
bool csr_op_win(const Matrix &A, CsrEpi epi, const CsrOpArgs &g, cudaStream_t s, int segment)
{
    if (!A.win.on || g.agg || segment != 0) return false;
    if ((reinterpret_cast<uintptr_t>(g.x) & 15u) != 0) return false;          // the ring is filled by 16-byte bulk copies of x
    if (A.win.num_long > 0 && (g.y == g.b || g.y == g.x)) return false;         // the long rows' dot products are parked in y before b and x are read
    WinArgs wa;
    wa.sv = A.win.sv.ptr();
    wa.so = A.win.so.ptr();
    wa.perm = A.win.perm.ptr();
    wa.tbase = A.win.tbase.ptr();
    wa.sbase = A.win.sbase.ptr();
    wa.ring = A.win.ring;
    wa.w = A.win.w;
    wa.tiles_per_cta = A.win.tiles_per_cta;
    wa.smeta = A.win.smeta.ptr();
    wa.slens = A.win.slens.ptr();
    wa.row_ptr = A.row_ptr.ptr();
    AMGXB_DISPATCH(A.mat_prec, A.vec_prec, {
        wa.x_len = A.n & ~(int)(16 / sizeof(VecT) - 1);
        TileArgs<MatT, VecT> ta;
        ta.row_ptr = A.row_ptr.ptr();
        ta.col = A.col_idx.ptr();
        ta.val = A.values.as<MatT>();
        ta.n = A.n;
        ta.row0 = 0;
        ta.num_tiles = A.plan.num_tiles;
        ta.cap = A.win.cap;
        ta.stages = A.win.stages;
        ta.unroll = 8;
        ta.perm = nullptr;
        ta.tile_base = 0;
        ta.l2pf = (l2_prefetch_flags() & 1) != 0;
        ta.x = (const VecT *)g.x;
        ta.agg = nullptr;
        ta.b = (const VecT *)g.b;
        ta.d = (const MatT *)g.d;
        ta.y = (VecT *)g.y;
        ta.omega = g.omega;
        ta.red = g.red;
        ta.fin_op = g.fin_op;
        ta.fin_slot = g.fin_slot;
        ta.mirror = g.mirror;
        switch (epi) {
        case EPI_SPMV: launch_win<MatT, VecT, EPI_SPMV>(A, ta, wa, s); break;
        case EPI_RESID: launch_win<MatT, VecT, EPI_RESID>(A, ta, wa, s); break;
        case EPI_ADD: launch_win<MatT, VecT, EPI_ADD>(A, ta, wa, s); break;
        case EPI_JACOBI:
        case EPI_JACOBI_L1: launch_win<MatT, VecT, EPI_JACOBI>(A, ta, wa, s); break;
        case EPI_SPMV_DOT: launch_win<MatT, VecT, EPI_SPMV_DOT>(A, ta, wa, s); break;
        case EPI_JACOBI_DOT: launch_win<MatT, VecT, EPI_JACOBI_DOT>(A, ta, wa, s); break;
        case EPI_RESID_NRM2: launch_win<MatT, VecT, EPI_RESID_NRM2>(A, ta, wa, s); break;
        }
    });
    return true;
}

}  // namespace amgxb
