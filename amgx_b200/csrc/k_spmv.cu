// k_spmv.cu -- the scalar-CSR kernel family of the solve phase for sm_100a:
//   SpMV, residual, fused (L1-)Jacobi sweep, each optionally fused with a dot / norm reduction and
//   optionally reading x through the aggregation prolongation.
//
// Replaces, on the hot path: Cusparse::bsrmv_internal / csrmv / cusparseSpMV
// (src/amgx_cusparse.cu:534-600, 983-1103), axmb (src/blas.cu:601-623), BlockJacobiSolver::smooth_1x1
// (src/solvers/block_jacobi_solver.cu:1286-1319), JacobiL1Solver::smooth_1x1
// (src/solvers/jacobi_l1_solver.cu:503-519) and the dot that follows the SpMV in PCG
// (src/solvers/pcg_solver.cu:118-137).
//
// Design (see DESIGN.md "CSR tile kernel"):
//  * persistent CTAs; a tile = TILE_ROWS consecutive rows.  A dedicated producer warp stages the
//    tile's row_ptr slice, col_idx range and value range into shared memory with TMA 1-D bulk
//    copies (cp.async.bulk ... mbarrier::complete_tx), STAGES deep, so HBM sees only fully
//    coalesced 16-byte-aligned bulk reads of the matrix.
//  * consumers: one thread per row walks its entries in shared memory left to right with an FMA
//    chain -- the exact per-row order of the reference's csrmv (y = a*x + y) -- gathering x
//    through L1/L2 (for stencil-like matrices neighbouring rows gather neighbouring x: coalesced).
//  * rows longer than the stage capacity fall back to a warp-per-row kernel with a shuffle tree.
//  * reductions: per-thread partial -> warp shuffle -> per-CTA partial -> the last CTA to finish
//    sums the partials in index order (deterministic) and applies the scalar epilogue (FinOp).
#include "kernels.h"
#include <cooperative_groups.h>

namespace amgxb {

long long g_kernel_launches = 0;

namespace {

#include "tile_common.cuh"

// One row's dot product from the staged tile: U column loads, then U gathers of x in flight, then the FMA chain in storage order
// (the per-row order of the reference's csrmv: y = a*x + y, amgx_cusparse.cu:1004-1012).  Slots past the end of the row re-read
// the row's first column (an L1 hit) and are not accumulated.
template <class MatT, class VecT, bool AGG, int U>
__device__ __forceinline__ VecT row_dot(const MatT *__restrict__ vals, const int *__restrict__ cols, int k, const int kend, const VecT *__restrict__ x,
                                        const int *__restrict__ agg)
{
    VecT sum = 0;
    for (; k + U <= kend; k += U) {
        int c[U];
        VecT xv[U];
#pragma unroll
        for (int j = 0; j < U; j++) c[j] = cols[k + j];
#pragma unroll
        for (int j = 0; j < U; j++) xv[j] = gather<VecT, AGG>(x, agg, c[j]);
#pragma unroll
        for (int j = 0; j < U; j++) sum = fma((VecT)vals[k + j], xv[j], sum);
    }
    if (k < kend) {
        int c[U - 1];
        VecT xv[U - 1];
#pragma unroll
        for (int j = 0; j < U - 1; j++) c[j] = cols[(k + j < kend) ? k + j : k];
#pragma unroll
        for (int j = 0; j < U - 1; j++) xv[j] = gather<VecT, AGG>(x, agg, c[j]);
#pragma unroll
        for (int j = 0; j < U - 1; j++)
            if (k + j < kend) sum = fma((VecT)vals[k + j], xv[j], sum);
    }
    return sum;
}

// ---------------------------------------------------------------------------------------------
// The tile kernel.  blockDim.x = TILE_ROWS + 32 (last warp = producer).
// shared memory layout: [stages x full mbarrier][stages x empty mbarrier][red scratch 40 doubles]
//                       then per stage: vals[cap] | cols[cap] | rp[TILE_ROWS+4]
// ---------------------------------------------------------------------------------------------
template <class MatT, class VecT, int TILE_ROWS, int EPI, bool AGG>
__global__ void __launch_bounds__(TILE_ROWS + PRODUCER_THREADS) csr_tile_kernel(const TileArgs<MatT, VecT> a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + MAX_STAGES;
    double *smem_red = reinterpret_cast<double *>(smem_raw + 2 * MAX_STAGES * sizeof(uint64_t));
    unsigned char *stage_base = smem_raw + 512;
    const size_t vals_bytes = (size_t)a.cap * sizeof(MatT);
    const size_t cols_bytes = (size_t)a.cap * sizeof(int);
    const size_t rp_bytes = (size_t)(TILE_ROWS + 4) * sizeof(int);
    const size_t stage_bytes = vals_bytes + cols_bytes + rp_bytes;
    constexpr int CONSUMER_WARPS = TILE_ROWS / 32;
    constexpr bool HAS_RED = (EPI == EPI_SPMV_DOT || EPI == EPI_JACOBI_DOT || EPI == EPI_RESID_NRM2);

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
    const int my_tiles = (a.num_tiles > (int)blockIdx.x) ? (a.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid >= TILE_ROWS) {
        // ------------------------------- producer warp -------------------------------
        if (tid == TILE_ROWS) {
            int s = 0;             // stage and mbarrier phase advance by counting: no runtime division per tile
            unsigned ph = 0u;
            for (int it = 0; it < my_tiles; it++, s = (s + 1 == a.stages) ? 0 : s + 1, ph ^= (s == 0) ? 1u : 0u) {
                const int tile = blockIdx.x + it * gridDim.x;
                if (it >= a.stages) mbar_wait(&empty[s], ph ^ 1u);
                const int r0 = a.row0 + tile * TILE_ROWS;
                const int r1 = min(r0 + TILE_ROWS, a.n);
                const int nz0 = __ldg(a.row_ptr + r0), nz1 = __ldg(a.row_ptr + r1);
                const int sa = nz0 & ~3, ea = (nz1 + 3) & ~3;
                unsigned char *st = stage_base + (size_t)s * stage_bytes;
                const unsigned rp_copy = (unsigned)(((r1 - r0 + 1 + 3) & ~3) * sizeof(int));
                const unsigned cnt = (unsigned)(ea - sa);
                mbar_expect_tx(&full[s], rp_copy + cnt * (unsigned)(sizeof(MatT) + sizeof(int)));
                tma_bulk_g2s(st + vals_bytes + cols_bytes, a.row_ptr + r0, rp_copy, &full[s]);
                if (cnt) {
                    tma_bulk_g2s(st, a.val + sa, cnt * (unsigned)sizeof(MatT), &full[s]);
                    tma_bulk_g2s(st + vals_bytes, a.col + sa, cnt * (unsigned)sizeof(int), &full[s]);
                }
                if (a.l2pf) {      // the consumers' per-row vector loads of this tile, `stages` tiles ahead of them (tile_common.cuh)
                    if (EPI == EPI_RESID || EPI == EPI_JACOBI || EPI == EPI_JACOBI_DOT || EPI == EPI_RESID_NRM2 || EPI == EPI_JACOBI_L1 || EPI == EPI_ADD)
                        l2_prefetch_span(a.b + r0, r1 - r0);
                    if (EPI == EPI_JACOBI || EPI == EPI_JACOBI_DOT || EPI == EPI_JACOBI_L1) l2_prefetch_span(a.d + r0, r1 - r0);
                }
            }
        }
    } else {
        // ------------------------------- consumers: one row per thread -------------------------------
        int s = 0;
        unsigned ph = 0u;
        for (int it = 0; it < my_tiles; it++, s = (s + 1 == a.stages) ? 0 : s + 1, ph ^= (s == 0) ? 1u : 0u) {
            const int tile = blockIdx.x + it * gridDim.x;
            // which row of the tile this thread takes: its own, or (irregular matrices) the tid-th longest, so that the rows of a warp
            // have similar lengths and the warp is not paced by its longest row.  A row is still summed left to right by ONE thread.
            const int lrow = a.perm ? (int)__ldg(a.perm + (size_t)(a.tile_base + tile) * TILE_ROWS + tid) : tid;
            const int row = a.row0 + tile * TILE_ROWS + lrow;
            const bool active = row < a.n;
            // operands that do not depend on the staged tile: issue their loads before waiting
            VecT bi = 0, xi = 0;
            MatT di = 1;
            if (active) {
                if (EPI == EPI_RESID || EPI == EPI_JACOBI || EPI == EPI_JACOBI_DOT || EPI == EPI_RESID_NRM2 || EPI == EPI_JACOBI_L1 || EPI == EPI_ADD)
                    bi = __ldg(a.b + row);
                if (EPI == EPI_JACOBI || EPI == EPI_JACOBI_DOT || EPI == EPI_JACOBI_L1) {
                    di = __ldg(a.d + row);
                    xi = gather<VecT, AGG>(a.x, a.agg, row);
                }
                if (EPI == EPI_SPMV_DOT) xi = gather<VecT, AGG>(a.x, a.agg, row);
            }
            const unsigned char *st = stage_base + (size_t)s * stage_bytes;
            const MatT *vals = reinterpret_cast<const MatT *>(st);
            const int *cols = reinterpret_cast<const int *>(st + vals_bytes);
            const int *rp = reinterpret_cast<const int *>(st + vals_bytes + cols_bytes);
            mbar_wait(&full[s], ph);
            if (active) {
                const int sa = rp[0] & ~3;
                int k = rp[lrow] - sa;
                const int kend = rp[lrow + 1] - sa;
                VecT sum = 0;
                // U gathers in flight per step (U = 8 when the plan says so: rows of a 7-point stencil then take ONE dependent
                // LDS -> gather -> FMA round instead of two); the FMA chain runs strictly left to right either way
                if (a.unroll == 8) sum = row_dot<MatT, VecT, AGG, 8>(vals, cols, k, kend, a.x, a.agg);
                else sum = row_dot<MatT, VecT, AGG, 4>(vals, cols, k, kend, a.x, a.agg);
                // ---- epilogue ----
                if (EPI == EPI_SPMV) {
                    a.y[row] = sum;
                } else if (EPI == EPI_SPMV_DOT) {
                    a.y[row] = sum;
                    acc += (double)sum * (double)xi;
                } else if (EPI == EPI_RESID) {
                    a.y[row] = bi - sum;
                } else if (EPI == EPI_ADD) {
                    a.y[row] = bi + sum;
                } else if (EPI == EPI_RESID_NRM2) {
                    const VecT r = bi - sum;
                    a.y[row] = r;
                    acc += (double)r * (double)r;
                } else {
                    // x + ((b - Ax) * w) * (1/d): d = 1/d; b -= y; b *= w; b*d + x  (one FMA)
                    MatT dinv = (MatT)1 / guard_diag<MatT>(di);
                    VecT t = bi - sum;
                    t = (VecT)(t * a.omega);
                    const VecT out = fma(t, (VecT)dinv, xi);
                    a.y[row] = out;
                    if (EPI == EPI_JACOBI_DOT) acc += (double)bi * (double)out;
                }
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty[s]);
        }
    }
    if (HAS_RED) block_reduce_finish(acc, smem_red, a.red, a.fin_op, a.fin_slot, a.mirror);
}

// ---------------------------------------------------------------------------------------------
// Fallback: warp per row, lanes stride the row, shuffle-tree reduction (long / irregular rows).
// ---------------------------------------------------------------------------------------------
template <class MatT, class VecT, int EPI, bool AGG>
__global__ void __launch_bounds__(256) csr_vector_kernel(const TileArgs<MatT, VecT> a)
{
    __shared__ double smem_red[40];
    constexpr bool HAS_RED = (EPI == EPI_SPMV_DOT || EPI == EPI_JACOBI_DOT || EPI == EPI_RESID_NRM2);
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    double acc = 0.0;
    for (int row = a.row0 + blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < a.n; row += gridDim.x * warps_per_block) {
        const int k0 = __ldg(a.row_ptr + row), k1 = __ldg(a.row_ptr + row + 1);
        VecT sum = 0;
        for (int k = k0 + lane; k < k1; k += 32) sum = fma((VecT)__ldg(a.val + k), gather<VecT, AGG>(a.x, a.agg, __ldg(a.col + k)), sum);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) {
            if (EPI == EPI_SPMV) {
                a.y[row] = sum;
            } else if (EPI == EPI_SPMV_DOT) {
                a.y[row] = sum;
                acc += (double)sum * (double)gather<VecT, AGG>(a.x, a.agg, row);
            } else if (EPI == EPI_RESID) {
                a.y[row] = a.b[row] - sum;
            } else if (EPI == EPI_ADD) {
                a.y[row] = a.b[row] + sum;
            } else if (EPI == EPI_RESID_NRM2) {
                const VecT r = a.b[row] - sum;
                a.y[row] = r;
                acc += (double)r * (double)r;
            } else {
                const VecT bi = a.b[row];
                MatT dinv = (MatT)1 / guard_diag<MatT>(a.d[row]);
                VecT t = bi - sum;
                t = (VecT)(t * a.omega);
                const VecT out = fma(t, (VecT)dinv, gather<VecT, AGG>(a.x, a.agg, row));
                a.y[row] = out;
                if (EPI == EPI_JACOBI_DOT) acc += (double)bi * (double)out;
            }
        }
    }
    if (HAS_RED) block_reduce_finish(acc, smem_red, a.red, a.fin_op, a.fin_slot, a.mirror);
}

// ---------------------------------------------------------------------------------------------
// plan construction
// ---------------------------------------------------------------------------------------------
__global__ void find_diag_kernel(const int *row_ptr, const int *col, int n, int *diag_idx)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int d = -1;
        for (int k = row_ptr[i]; k < row_ptr[i + 1]; k++)
            if (col[k] == i) { d = k; break; }
        diag_idx[i] = d;
    }
}

__global__ void tile_stats_kernel(const int *row_ptr, int row0, int n, int tile_rows, int num_tiles, int *max_tile_nnz, int *max_row_nnz)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < num_tiles; t += gridDim.x * blockDim.x) {
        int r0 = row0 + t * tile_rows, r1 = min(r0 + tile_rows, n);
        int sa = row_ptr[r0] & ~3, ea = (row_ptr[r1] + 3) & ~3;
        atomicMax(max_tile_nnz, ea - sa);
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicMax(max_row_nnz, row_ptr[i + 1] - row_ptr[i]);
}

// thread -> row map of every tile: rows sorted by (length descending, index ascending); rows past the end of the segment come last.
// One CTA per tile, rank sort in shared memory (setup time).
template <int TILE_ROWS>
__global__ void __launch_bounds__(TILE_ROWS) tile_perm_kernel(const int *__restrict__ rp, int row0, int n, int num_tiles, int tile_base, unsigned char *perm)
{
    __shared__ int len[TILE_ROWS];
    const int tid = threadIdx.x;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int row = row0 + t * TILE_ROWS + tid;
        const int mine = row < n ? rp[row + 1] - rp[row] : -1;
        len[tid] = mine;
        __syncthreads();
        int rank = 0;
        for (int j = 0; j < TILE_ROWS; j++) {
            const int o = len[j];
            rank += (o > mine) || (o == mine && j < tid);
        }
        perm[(size_t)(tile_base + t) * TILE_ROWS + rank] = (unsigned char)tid;
        __syncthreads();
    }
}

size_t tile_smem_bytes(int cap, int stages, int tile_rows, size_t mat_size)
{
    return 512 + (size_t)stages * ((size_t)cap * (mat_size + 4) + (size_t)(tile_rows + 4) * 4);
}

template <class MatT, class VecT, int TILE_ROWS, int EPI> void launch_tile(const Matrix &A, const TileArgs<MatT, VecT> &ta, int grid, cudaStream_t s)
{
    const size_t smem = A.plan.smem_bytes;
    if (ta.agg) {
        auto k = csr_tile_kernel<MatT, VecT, TILE_ROWS, EPI, true>;
        smem_opt_in(reinterpret_cast<const void *>(k), smem);      // exactly what this kernel needs, once per size and device
        k<<<grid, TILE_ROWS + PRODUCER_THREADS, smem, s>>>(ta);
    } else {
        auto k = csr_tile_kernel<MatT, VecT, TILE_ROWS, EPI, false>;
        smem_opt_in(reinterpret_cast<const void *>(k), smem);      // exactly what this kernel needs, once per size and device
        k<<<grid, TILE_ROWS + PRODUCER_THREADS, smem, s>>>(ta);
    }
}

template <class MatT, class VecT, int EPI> void launch_epi(const Matrix &A, const TileArgs<MatT, VecT> &ta, cudaStream_t s)
{
    if (A.plan.use_tiles) {
        const int grid = std::max(1, std::min(csr_max_grid(A), ta.num_tiles));
        if (A.plan.tile_rows == 256) launch_tile<MatT, VecT, 256, EPI>(A, ta, grid, s);
        else launch_tile<MatT, VecT, 128, EPI>(A, ta, grid, s);
    } else {
        const int grid = std::max(1, std::min(csr_max_grid(A), ceil_div(ta.n - ta.row0, 8)));
        if (ta.agg) csr_vector_kernel<MatT, VecT, EPI, true><<<grid, 256, 0, s>>>(ta);
        else csr_vector_kernel<MatT, VecT, EPI, false><<<grid, 256, 0, s>>>(ta);
    }
    count_launch();
    AMGXB_LAUNCH_CHECK();
}

}  // namespace

int csr_max_grid(const Matrix &A)
{
    const int sms = A.rsc ? A.rsc->num_sms : 148;
    if (A.plan.use_tiles) {
        const int per_sm = std::max(1, A.plan.ctas_per_sm);
        return std::max(1, std::min(A.plan.num_tiles, sms * per_sm));
    }
    return std::max(1, std::min(ceil_div(A.n, 8), sms * 8));
}

void csr_build_plan(Matrix &A, cudaStream_t s)
{
    if (A.bs() != 1) { A.plan = TilePlan(); return; }   // block matrices use the block kernels (k_block.cu)
    A.diag_idx.resize(A.n);
    if (A.n == 0) { A.plan = TilePlan(); return; }
    find_diag_kernel<<<std::min(ceil_div(A.n, 256), 4096), 256, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), A.n, A.diag_idx.ptr());
    count_launch();
    AMGXB_LAUNCH_CHECK();
    const int sms = A.rsc ? A.rsc->num_sms : 148;
    TilePlan p;
    // 256-row tiles once there are enough of them to fill the machine, else 128-row tiles
    p.tile_rows = (ceil_div(A.n, 256) >= 2 * sms) ? 256 : 128;
    p.num_tiles = ceil_div(A.n, p.tile_rows);
    DevBuf<int> stats;
    stats.resize(2);
    stats.zero(s);
    tile_stats_kernel<<<std::min(ceil_div(A.n, 256), 1024), 256, 0, s>>>(A.row_ptr.ptr(), 0, A.n, p.tile_rows, p.num_tiles, stats.ptr(), stats.ptr() + 1);
    count_launch();
    p.split = (A.split_row / 4) * 4;   // distributed: rows [0, split) never touch halo columns
    // The split buys overlap of the halo exchange with the interior rows; below AMGXB_SPLIT_ROWS rows an interior kernel is shorter than
    // the two extra launches the split costs, so small levels exchange first (one kernel on the peer-memory path) and run as a whole.
    // With the peer-memory exchange (one ~10 us kernel, p2p.cu) even the fine level gains nothing from the overlap that the two extra
    // launches and the cross-stream edges cost (r02, N = 2: 283 it/s split at 2^20 rows, 288 never split), so the split is for the NCCL path.
    static const int env_split = getenv("AMGXB_SPLIT_ROWS") ? atoi(getenv("AMGXB_SPLIT_ROWS")) : -1;
    const int split_rows = env_split >= 0 ? env_split : ((A.rsc && A.rsc->p2p) ? (1 << 30) : (1 << 20));
    if (A.n < split_rows) p.split = 0;
    if (p.split > 0 && p.split < A.n) {
        tile_stats_kernel<<<std::min(ceil_div(A.n, 256), 1024), 256, 0, s>>>(A.row_ptr.ptr(), 0, p.split, p.tile_rows, ceil_div(p.split, p.tile_rows), stats.ptr(), stats.ptr() + 1);
        tile_stats_kernel<<<std::min(ceil_div(A.n, 256), 1024), 256, 0, s>>>(A.row_ptr.ptr(), p.split, A.n, p.tile_rows, ceil_div(A.n - p.split, p.tile_rows), stats.ptr(), stats.ptr() + 1);
        count_launch(2);
    } else p.split = 0;
    AMGXB_LAUNCH_CHECK();
    std::vector<int> h = stats.to_host(s);
    p.max_tile_nnz = std::max(4, h[0]);
    const size_t msz = prec_size(A.mat_prec);
    p.use_tiles = false;
    p.max_row_nnz = h[1];
    // tuning knobs (defaults chosen from the r02 sweeps in profiles/): pipeline depth, CTAs per SM, gathers in flight per step
    static const int env_stages = getenv("AMGXB_TILE_STAGES") ? atoi(getenv("AMGXB_TILE_STAGES")) : 0;
    static const int env_unroll = getenv("AMGXB_TILE_UNROLL") ? atoi(getenv("AMGXB_TILE_UNROLL")) : 0;
    p.unroll = (env_unroll == 4 || env_unroll == 8) ? env_unroll : 4;
    // The consumers are latency-bound (one row per thread, dependent LDS -> gather -> FMA rounds), so resident consumer warps are what
    // buys bandwidth: pick the pipeline depth that lets the most CTAs share an SM (ties: the deeper pipeline).  r02 sweep on 256^3 / 512^3
    // 7-point Poisson (profiles/r02_tile_sweep.md): 4 stages x 2 CTAs/SM 0.300 ms, 2 stages x 3 CTAs 0.291 ms, 2 stages x 4 CTAs 0.250 ms.
    static const int env_ctas = getenv("AMGXB_TILE_CTAS") ? atoi(getenv("AMGXB_TILE_CTAS")) : 0;
    const int by_threads = std::min(2048 / (p.tile_rows + PRODUCER_THREADS), 65536 / ((p.tile_rows + PRODUCER_THREADS) * 40));   // threads, registers (40 / thread)
    int best_ctas = 0;
    for (int st = (env_stages >= 2 && env_stages <= MAX_STAGES) ? env_stages : MAX_STAGES; st >= 2; st--) {
        const size_t need = tile_smem_bytes(p.max_tile_nnz, st, p.tile_rows, msz);
        if (need > (size_t)216 * 1024) continue;
        const int ctas = std::max(1, std::min(by_threads, (int)((size_t)227 * 1024 / (need + 1024))));
        if (ctas > best_ctas) { best_ctas = ctas; p.stages = st; p.smem_bytes = need; p.use_tiles = true; }
        if (env_stages) break;
    }
    p.ctas_per_sm = env_ctas > 0 ? std::min(env_ctas, std::max(best_ctas, 1)) : std::max(best_ctas, 1);
    // irregular rows (opt-in, AMGXB_TILE_PERM=1): hand the rows of a tile to the threads sorted by length
    {
        static const int env_perm = getenv("AMGXB_TILE_PERM") ? atoi(getenv("AMGXB_TILE_PERM")) : -1;
        const double mean = (double)A.nnz / std::max(A.n, 1);
        // r02 A/B on B200: OFF by default.  On the Poisson hierarchy (coarse levels: mean 8-15, max ~2x) the sorted assignment costs 8 % of the
        // solve (311 vs 336 it/s): the vectors of a tile (b, d, x, y) are then touched in a scattered order by each warp; on the 4 M-row
        // banded matrix it is neutral (SpMV 0.45 vs 0.43 of peak): that kernel is bound by the L2 sectors of the random gathers, not by imbalance.
        p.use_perm = p.use_tiles && env_perm > 0;
        (void)mean;
        if (p.use_perm) {
            const int T = p.tile_rows;
            int nseg = 1, r0[2] = {0, 0}, r1[2] = {A.n, 0}, nt[2] = {ceil_div(A.n, T), 0};
            if (p.split > 0 && p.split < A.n) { nseg = 2; r1[0] = p.split; nt[0] = ceil_div(p.split, T); r0[1] = p.split; r1[1] = A.n; nt[1] = ceil_div(A.n - p.split, T); }
            A.tile_perm.resize((size_t)(nt[0] + nt[1]) * T);
            for (int g = 0, base = 0; g < nseg; base += nt[g], g++) {
                const int grid = std::max(1, std::min(nt[g], sms * 8));
                if (T == 256) tile_perm_kernel<256><<<grid, 256, 0, s>>>(A.row_ptr.ptr(), r0[g], r1[g], nt[g], base, A.tile_perm.ptr());
                else tile_perm_kernel<128><<<grid, 128, 0, s>>>(A.row_ptr.ptr(), r0[g], r1[g], nt[g], base, A.tile_perm.ptr());
                count_launch();
            }
            AMGXB_LAUNCH_CHECK();
        }
    }
    A.plan = p;
    csr_build_colenc(A, s);     // no-op unless AMGXB_COLENC=1
    csr_build_window(A, s);     // banded irregular levels only
}

void csr_op(const Matrix &A, CsrEpi epi, const CsrOpArgs &g, cudaStream_t s, int segment)
{
    if (A.bs() != 1) fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "csr_op: scalar kernel called on a block matrix");
    if (A.n == 0) return;
    // segment 0: all rows; 1: rows [0, split) (no halo columns); 2: rows [split, n)
    const int row0 = (segment == 2) ? A.plan.split : 0;
    const int row1 = (segment == 1) ? A.plan.split : A.n;
    if (row1 <= row0) return;
    if (A.win.on && csr_op_win(A, epi, g, s, segment)) return;
    if (A.colenc.on && csr_op_enc(A, epi, g, s, segment)) return;
    AMGXB_DISPATCH(A.mat_prec, A.vec_prec, {
        TileArgs<MatT, VecT> ta;
        ta.row_ptr = A.row_ptr.ptr();
        ta.col = A.col_idx.ptr();
        ta.val = A.values.as<MatT>();
        ta.n = row1;
        ta.row0 = row0;
        ta.num_tiles = ceil_div(row1 - row0, std::max(1, A.plan.tile_rows));
        ta.cap = A.plan.max_tile_nnz;
        ta.stages = A.plan.stages;
        ta.unroll = A.plan.unroll;
        // (a split matrix applied as a whole has another tiling than its two segments: no map then)
        ta.perm = (A.plan.use_perm && !(A.plan.split > 0 && segment == 0)) ? A.tile_perm.ptr() : nullptr;
        ta.tile_base = (segment == 2) ? ceil_div(A.plan.split, std::max(1, A.plan.tile_rows)) : 0;
        ta.l2pf = (l2_prefetch_flags() & 1) != 0;
        ta.x = (const VecT *)g.x;
        ta.agg = g.agg;
        ta.b = (const VecT *)g.b;
        ta.d = (const MatT *)g.d;
        ta.y = (VecT *)g.y;
        ta.omega = g.omega;
        ta.red = g.red;
        ta.fin_op = g.fin_op;
        ta.fin_slot = g.fin_slot;
        ta.mirror = g.mirror;
        switch (epi) {
        case EPI_SPMV: launch_epi<MatT, VecT, EPI_SPMV>(A, ta, s); break;
        case EPI_RESID: launch_epi<MatT, VecT, EPI_RESID>(A, ta, s); break;
        case EPI_ADD: launch_epi<MatT, VecT, EPI_ADD>(A, ta, s); break;
        case EPI_JACOBI:
        case EPI_JACOBI_L1: launch_epi<MatT, VecT, EPI_JACOBI>(A, ta, s); break;
        case EPI_SPMV_DOT: launch_epi<MatT, VecT, EPI_SPMV_DOT>(A, ta, s); break;
        case EPI_JACOBI_DOT: launch_epi<MatT, VecT, EPI_JACOBI_DOT>(A, ta, s); break;
        case EPI_RESID_NRM2: launch_epi<MatT, VecT, EPI_RESID_NRM2>(A, ta, s); break;
        }
    });
}

}  // namespace amgxb
