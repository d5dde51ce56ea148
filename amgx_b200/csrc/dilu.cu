// dilu.cu -- MULTICOLOR_DILU smoother: per-colour setup of Einv, forward and backward sweeps, block sizes 1 and 4.
//   setup    DILU_setup_1x1_kernel / DILU_setup_NxN_kernel   src/solvers/multicolor_dilu_solver.cu:643-812, 362-640
//   forward  DILU_forward_1x1_kernel / _4x4_                 :1763-1901, 1585-1759
//   backward DILU_backward_1x1_kernel / _4x4_ / _skip        :2772-2915, 2604-2768
//   host     solve_iteration / smooth_NxN / computeEinv_NxN  :3773-3851, 4021-4242, 3890-4019
// Formulas (boundary_coloring = SYNC_COLORS; j ranges over OWNED columns in the colour-restricted sums):
//   E_i     = A_ii - sum_{colour(j) < colour(i), j != i} A_ij Einv_j A_ji ;  Einv_i = E_i^{-1} (0 stays 0 for 1x1)
//   forward (colour c ascending):  delta_i = Einv_i ( b_i - sum_j A_ij (x_j + [c != 0 and colour(j) < c] delta_j) )
//   backward (colour c descending): Delta_i = delta_i - Einv_i sum_{[c != 0 and colour(j) > c]} A_ij Delta_j ;  x_i += w Delta_i
//   (the "c != 0" guards are the reference's; the last colour uses the cheap Delta = delta shortcut).
// The 1x1 kernels keep the reference's work decomposition (8 lanes per row, per-lane FMA accumulation, xor-butterfly
// reduction; setup: 32 lanes per row) so that their sums associate exactly like the reference's.
#include "solvers.h"
#include "dist.h"
#include <cub/cub.cuh>

namespace amgxb {

void color_matrix(Matrix &A, const std::string &scheme, double max_uncolored_fraction, cudaStream_t s);   // coloring.cu

namespace {

#include "tile_common.cuh"

constexpr int NTPR = 8;   // lanes per row in the 1x1 sweeps

template <class MatT, class VecT>
__global__ void dilu_setup_1x1(const int *__restrict__ rp, const int *__restrict__ ci, const int *__restrict__ diag, const MatT *__restrict__ va,
                               MatT *Einv, const int *__restrict__ rows, const int *__restrict__ colors, int nrows, int color, int n_owned)
{
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = gridDim.x * (blockDim.x >> 5);
    for (int it = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); it < nrows; it += warps_per_grid) {
        const int i = rows[it];
        VecT e_out = 0;
        if (color != 0) {
            const int k1 = rp[i + 1];
            for (int k0 = rp[i]; k0 < k1; k0 += 32) {
                const int k = k0 + lane;
                if (k < k1) {
                    const int j = ci[k];
                    if (j != i && j < n_owned && colors[j] < color) {
                        MatT a_ji = 0;
                        for (int kk = rp[j]; kk < rp[j + 1]; kk++)
                            if (ci[kk] == i) { a_ji = va[kk]; break; }   // first match, as the reference's search order yields
                        e_out += (VecT)(a_ji * Einv[j]) * (VecT)va[k];
                    }
                }
            }
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) e_out += __shfl_xor_sync(0xffffffffu, e_out, m);
        if (lane == 0) {
            const int d = diag[i];
            MatT res = (d >= 0 ? va[d] : (MatT)0) - (MatT)e_out;
            if (res != (MatT)0) res = (MatT)1 / res;
            Einv[i] = res;
        }
    }
}

// Loads of the vectors a sweep updates (x, delta, Delta).  The per-colour kernels see them through L1 (kernel boundaries
// invalidate it); the fused level kernel below runs all colours in ONE launch across the CTAs of a cluster, where another SM's
// L1 may hold a stale line: there the loads go to L2 (ld.global.cg).
template <bool CG, class T> __device__ __forceinline__ T ldv(const T *p) { return CG ? __ldcg(p) : *p; }

// one row of the forward sweep, 8 lanes per row (lane l of the row group); identical for every caller, so the sums associate alike
template <class MatT, class VecT, bool CG>
__device__ __forceinline__ void dilu_fwd_row_1x1(const bool act, const int i, const int l, const int *__restrict__ rp, const int *__restrict__ ci,
                                                 const MatT *__restrict__ va, const VecT *x, const VecT *__restrict__ b, VecT *delta, int color,
                                                 const int *__restrict__ colors, const MatT *__restrict__ Einv, int n_owned)
{
    VecT acc = 0;
    if (act && l == 0) acc = b[i];
    if (act) {
        const int k1 = rp[i + 1];
        for (int k = rp[i] + l; k < k1; k += NTPR) {
            const int j = ci[k];
            VecT xx = ldv<CG>(x + j);
            if (color != 0 && j < n_owned && colors[j] < color) xx += ldv<CG>(delta + j);
            acc -= (VecT)va[k] * xx;
        }
    }
#pragma unroll
    for (int m = NTPR / 2; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if (act && l == 0) delta[i] = (VecT)Einv[i] * acc;
}

template <class MatT, class VecT, bool CG>
__device__ __forceinline__ void dilu_bwd_row_1x1(const bool act, const int i, const int l, const int *__restrict__ rp, const int *__restrict__ ci,
                                                 const MatT *__restrict__ va, VecT *x, double weight, const int *__restrict__ colors,
                                                 const MatT *__restrict__ Einv, const VecT *delta, VecT *Delta, int color, int n_owned)
{
    VecT acc = 0;
    if (act) {
        const int k1 = rp[i + 1];
        for (int k = rp[i] + l; k < k1; k += NTPR) {
            const int j = ci[k];
            const bool valid = color != 0 && j < n_owned && colors[j] > color;
            if (valid) acc += (VecT)va[k] * ldv<CG>(Delta + j);
        }
    }
#pragma unroll
    for (int m = NTPR / 2; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if (act && l == 0) {
        const VecT v = ldv<CG>(delta + i) - (VecT)Einv[i] * acc;
        x[i] = ldv<CG>(x + i) + (VecT)(weight * v);
        Delta[i] = v;
    }
}

template <class MatT, class VecT>
__global__ void dilu_forward_1x1(const int *__restrict__ rp, const int *__restrict__ ci, const MatT *__restrict__ va, const VecT *x,
                                 const VecT *__restrict__ b, VecT *delta, const int *__restrict__ rows, int nrows, int color,
                                 const int *__restrict__ colors, const MatT *__restrict__ Einv, int n_owned)
{
    const int l = threadIdx.x % NTPR;
    const int rows_per_grid = gridDim.x * (blockDim.x / NTPR);
    for (int it = blockIdx.x * (blockDim.x / NTPR) + threadIdx.x / NTPR; __any_sync(0xffffffffu, it < nrows); it += rows_per_grid) {
        const bool act = it < nrows;
        dilu_fwd_row_1x1<MatT, VecT, false>(act, act ? rows[it] : 0, l, rp, ci, va, x, b, delta, color, colors, Einv, n_owned);
    }
}

template <class MatT, class VecT>
__global__ void dilu_backward_1x1(const int *__restrict__ rp, const int *__restrict__ ci, const MatT *__restrict__ va, VecT *x, double weight,
                                  const int *__restrict__ rows, const int *__restrict__ colors, const MatT *__restrict__ Einv,
                                  const VecT *__restrict__ delta, VecT *Delta, int nrows, int color, int n_owned)
{
    const int l = threadIdx.x % NTPR;
    const int rows_per_grid = gridDim.x * (blockDim.x / NTPR);
    for (int it = blockIdx.x * (blockDim.x / NTPR) + threadIdx.x / NTPR; __any_sync(0xffffffffu, it < nrows); it += rows_per_grid) {
        const bool act = it < nrows;
        dilu_bwd_row_1x1<MatT, VecT, false>(act, act ? rows[it] : 0, l, rp, ci, va, x, weight, colors, Einv, delta, Delta, color, n_owned);
    }
}

template <class VecT> __global__ void dilu_backward_skip(VecT *x, double weight, const int *__restrict__ rows, const VecT *__restrict__ delta, VecT *Delta, int nrows, int bs)
{
    const long long total = (long long)nrows * bs;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = rows[(int)(t / bs)], c = (int)(t % bs);
        const size_t idx = (size_t)i * bs + c;
        const VecT v = delta[idx];
        x[idx] = x[idx] + (VecT)(weight * v);
        Delta[idx] = v;
    }
}

// ------------------------------------------ 4x4 blocks ------------------------------------------
// In-place Gauss-Jordan without pivoting, the operation order of the reference's lock-step 16-thread version
// (multicolor_dilu_solver.cu:586-633; include/solvers/block_common_solver.h:106-137).
template <class T> __device__ __forceinline__ T guard0(T d);
template <> __device__ __forceinline__ double guard0<double>(double d) { return fabs(d) < 1e-12 ? copysign(1e-12, d) : d; }
template <> __device__ __forceinline__ float guard0<float>(float d) { return fabs((double)d) < 1e-7 ? copysignf((float)1e-7, d) : d; }

template <class T> __device__ void invert4x4(T *A)
{
    for (int row = 0; row < 4; row++) {
        const T diag = (T)1 / guard0<T>(A[row * 4 + row]);
        for (int j = 0; j < 4; j++) if (j != row) A[row * 4 + j] *= diag;
        for (int i = 0; i < 4; i++) if (i != row)
            for (int j = 0; j < 4; j++) if (j != row) A[i * 4 + j] -= A[i * 4 + row] * A[row * 4 + j];
        for (int j = 0; j < 4; j++) A[j * 4 + row] = (j == row) ? diag : -A[j * 4 + row] * diag;
    }
}

template <class MatT, class VecT>
__global__ void dilu_setup_4x4(const int *__restrict__ rp, const int *__restrict__ ci, const int *__restrict__ diag, const MatT *__restrict__ va,
                               MatT *Einv, const int *__restrict__ rows, const int *__restrict__ colors, int nrows, int color, int n_owned)
{
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < nrows; it += gridDim.x * blockDim.x) {
        const int i = rows[it];
        VecT E[16];
        const int d = diag[i];
        for (int m = 0; m < 16; m++) E[m] = d >= 0 ? (VecT)va[(size_t)d * 16 + m] : (VecT)0;
        if (color != 0) {
            for (int k = rp[i]; k < rp[i + 1]; k++) {
                const int j = ci[k];
                if (j == i || j >= n_owned || colors[j] >= color) continue;
                int kji = -1;
                for (int kk = rp[j]; kk < rp[j + 1]; kk++)
                    if (ci[kk] == i) { kji = kk; break; }
                VecT T[16];
                for (int r = 0; r < 4; r++)
                    for (int c = 0; c < 4; c++) {
                        VecT t = 0;
                        for (int m = 0; m < 4; m++) t += (VecT)va[(size_t)k * 16 + r * 4 + m] * (VecT)Einv[(size_t)j * 16 + m * 4 + c];
                        T[r * 4 + c] = t;
                    }
                if (kji >= 0)
                    for (int r = 0; r < 4; r++)
                        for (int c = 0; c < 4; c++)
                            for (int m = 0; m < 4; m++) E[r * 4 + c] -= T[r * 4 + m] * (VecT)va[(size_t)kji * 16 + m * 4 + c];
            }
        }
        invert4x4<VecT>(E);
        for (int m = 0; m < 16; m++) Einv[(size_t)i * 16 + m] = (MatT)E[m];
    }
}

// One warp per block row: quad q (lanes 4q..4q+3) takes the blocks q, q+8, ... of the row, thread r of a quad owns
// component r; the eight partial 4-vectors are combined with an xor butterfly over the quads.  (A single quad walking
// a 30..60-block coarse row serially is a pure latency chain: ~0.5 us per block.)
// 4 consecutive scalars with 16-byte loads (a block row of the matrix, a block of a vector); CG: through L2 (see ldv)
template <bool CG> __device__ __forceinline__ void ld4v(const double *p, double (&o)[4])
{
    const double2 a = CG ? __ldcg(reinterpret_cast<const double2 *>(p)) : *reinterpret_cast<const double2 *>(p);
    const double2 b = CG ? __ldcg(reinterpret_cast<const double2 *>(p) + 1) : *(reinterpret_cast<const double2 *>(p) + 1);
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
template <bool CG> __device__ __forceinline__ void ld4v(const float *p, float (&o)[4])
{
    const float4 a = CG ? __ldcg(reinterpret_cast<const float4 *>(p)) : *reinterpret_cast<const float4 *>(p);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
}

template <class MatT, class VecT, bool BACKWARD, bool CG, class Mid>
__device__ __forceinline__ void dilu_row_4x4(const int i, const int k0, const int k1, const int lane, const int *__restrict__ ci, const MatT *__restrict__ va, VecT *x,
                                             const VecT *__restrict__ b, VecT *delta, VecT *Delta, double weight, int color, const int *__restrict__ colors,
                                             const MatT *__restrict__ Einv, int n_owned, Mid mid)
{
    const int r = lane & 3, q = lane >> 2;
    VecT acc = 0;
    if (!BACKWARD && q == 0) acc = b[(size_t)i * 4 + r];
    for (int k = k0 + q; k < k1; k += 8) {
        const int j = ci[k];
        MatT a[4];
        ld4v<false>(va + (size_t)k * 16 + r * 4, a);
        if (!BACKWARD) {
            const bool valid = color != 0 && j < n_owned && colors[j] < color;
            VecT xv[4], dv[4] = {0, 0, 0, 0};
            ld4v<CG>(x + (size_t)j * 4, xv);
            if (valid) ld4v<CG>(delta + (size_t)j * 4, dv);
#pragma unroll
            for (int m = 0; m < 4; m++) {
                VecT xx = xv[m];
                if (valid) xx += dv[m];
                acc -= (VecT)a[m] * xx;
            }
        } else {
            const bool valid = color != 0 && j < n_owned && colors[j] > color;
            if (valid) {
                VecT Dv[4];
                ld4v<CG>(Delta + (size_t)j * 4, Dv);
#pragma unroll
                for (int m = 0; m < 4; m++) acc += (VecT)a[m] * Dv[m];
            }
        }
    }
    mid();      // the caller's prefetch of the NEXT row's extent goes here: behind this row's gathers, in front of its reduction
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    // y = Einv_i * acc (4x4 mat-vec inside quad 0; every quad holds the full acc)
    VecT y = 0;
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const VecT am = __shfl_sync(0xffffffffu, acc, m);
        y += (VecT)Einv[(size_t)i * 16 + r * 4 + m] * am;
    }
    if (q == 0) {
        const size_t idx = (size_t)i * 4 + r;
        if (!BACKWARD) delta[idx] = y;
        else {
            const VecT v = ldv<CG>(delta + idx) - y;
            x[idx] = ldv<CG>(x + idx) + (VecT)(weight * v);
            Delta[idx] = v;
        }
    }
}

// One colour of a sweep.  A warp walks its rows with the NEXT row's index and extent already in flight (a row is a chain of dependent
// loads: sorted_rows -> row_ptr -> col -> x; prefetching the first two halves it: r02 ncu, 25 % of the DRAM peak and latency-bound).
template <class MatT, class VecT, bool BACKWARD>
__global__ void __launch_bounds__(128) dilu_sweep_4x4(const int *__restrict__ rp, const int *__restrict__ ci, const MatT *__restrict__ va, VecT *x,
                                                      const VecT *__restrict__ b, VecT *delta, VecT *Delta, double weight, const int *__restrict__ rows, int nrows,
                                                      int color, const int *__restrict__ colors, const MatT *__restrict__ Einv, int n_owned)
{
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = gridDim.x * (blockDim.x >> 5);
    int it = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (it >= nrows) return;
    int i = rows[it];
    int k0 = rp[i], k1 = rp[i + 1];
    while (true) {
        const int itn = it + warps_per_grid;
        const int in = itn < nrows ? __ldg(rows + itn) : -1;
        int k0n = 0, k1n = 0;
        dilu_row_4x4<MatT, VecT, BACKWARD, false>(i, k0, k1, lane, ci, va, x, b, delta, Delta, weight, color, colors, Einv, n_owned,
                                                  [&] { if (in >= 0) { k0n = __ldg(rp + in); k1n = __ldg(rp + in + 1); } });
        if (in < 0) break;
        it = itn; i = in; k0 = k0n; k1 = k1n;
    }
}

// ---------------------------------------------------------------------------------------------
// Fused level kernel: ALL colours of a forward + backward sweep (and all `sweeps` sweeps of a smooth call) in one launch by one
// thread-block CLUSTER; colours are separated by cluster barriers (barrier.cluster, hardware-supported on sm_90+/sm_100) instead of
// kernel boundaries.  For the levels of the hierarchy whose colours hold fewer rows than the machine has warps -- on a 20-level
// block hierarchy that is 12 levels x 2 x (9..40) colours x 2 sweeps of 15 us latency-bound launches (profiles/r02_block.md).
// Same per-row functions as the per-colour kernels => the same bits.  Mutable vectors are read through L2 (see ldv).
// ---------------------------------------------------------------------------------------------
struct DiluLevelArgs {
    const int *rp, *ci, *colors, *sorted_rows, *color_offsets;   // color_offsets: device copy, [num_colors + 1]
    const void *va, *Einv, *b;
    void *x, *delta, *Delta;
    double weight;
    int n, n_owned, num_colors, sweeps, zero_first, bs;
};

__device__ __forceinline__ void cluster_barrier()
{
    __threadfence();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_nctarank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}

template <class MatT, class VecT, int BS>
__global__ void __launch_bounds__(1024) dilu_level_kernel(const DiluLevelArgs a)
{
    const int *__restrict__ rp = a.rp, *__restrict__ ci = a.ci, *__restrict__ colors = a.colors, *__restrict__ rows = a.sorted_rows;
    const MatT *__restrict__ va = (const MatT *)a.va, *__restrict__ Einv = (const MatT *)a.Einv;
    const VecT *__restrict__ b = (const VecT *)a.b;
    VecT *x = (VecT *)a.x, *delta = (VecT *)a.delta, *Delta = (VecT *)a.Delta;
    const int lane = threadIdx.x & 31;
    const int nthreads = (int)(cluster_nctarank() * blockDim.x), tid = (int)(cluster_ctarank() * blockDim.x + threadIdx.x);
    const int nc = a.num_colors;
    for (int sw = 0; sw < a.sweeps; sw++) {
        if (sw == 0 && a.zero_first) {
            for (long long t = tid; t < (long long)a.n * BS; t += nthreads) x[t] = 0;
            cluster_barrier();
        }
        for (int c = 0; c < nc; c++) {          // forward, colours ascending
            const int off = a.color_offsets[c], cnt = a.color_offsets[c + 1] - off;
            if (cnt == 0) continue;
            if (BS == 1) {
                const int l = threadIdx.x % NTPR;
                for (int it = tid / NTPR; __any_sync(0xffffffffu, it < cnt); it += nthreads / NTPR) {
                    const bool act = it < cnt;
                    dilu_fwd_row_1x1<MatT, VecT, true>(act, act ? rows[off + it] : 0, l, rp, ci, va, x, b, delta, c, colors, Einv, a.n_owned);
                }
            } else {
                for (int it = tid >> 5; it < cnt; it += nthreads >> 5)
                    { const int i = rows[off + it]; dilu_row_4x4<MatT, VecT, false, true>(i, rp[i], rp[i + 1], lane, ci, va, x, b, delta, Delta, a.weight, c, colors, Einv, a.n_owned, [] {}); }
            }
            cluster_barrier();
        }
        for (int c = nc - 1; c >= 0; c--) {     // backward, colours descending
            const int off = a.color_offsets[c], cnt = a.color_offsets[c + 1] - off;
            if (cnt == 0) continue;
            if (c == nc - 1) {                  // last colour: Delta = delta (dilu_backward_skip)
                for (long long t = tid; t < (long long)cnt * BS; t += nthreads) {
                    const size_t idx = (size_t)rows[off + (int)(t / BS)] * BS + (size_t)(t % BS);
                    const VecT v = __ldcg(delta + idx);
                    x[idx] = __ldcg(x + idx) + (VecT)(a.weight * v);
                    Delta[idx] = v;
                }
            } else if (BS == 1) {
                const int l = threadIdx.x % NTPR;
                for (int it = tid / NTPR; __any_sync(0xffffffffu, it < cnt); it += nthreads / NTPR) {
                    const bool act = it < cnt;
                    dilu_bwd_row_1x1<MatT, VecT, true>(act, act ? rows[off + it] : 0, l, rp, ci, va, x, a.weight, colors, Einv, delta, Delta, c, a.n_owned);
                }
            } else {
                for (int it = tid >> 5; it < cnt; it += nthreads >> 5)
                    { const int i = rows[off + it]; dilu_row_4x4<MatT, VecT, true, true>(i, rp[i], rp[i + 1], lane, ci, va, x, b, delta, Delta, a.weight, c, colors, Einv, a.n_owned, [] {}); }
            }
            cluster_barrier();
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Colour-sorted DILU for 4x4 blocks on the LARGE levels (r02).  The per-colour kernels above walk rows in colour order through the
// sorted_rows indirection: every row is a chain of dependent loads and the matrix arrives in scattered 64-byte requests -- 25 % of the
// DRAM peak (ncu, profiles/r02_ncu_kernels.md).  Here the smoother keeps its own copy of the matrix with the ROWS SORTED BY COLOUR
// (the reference's reorder idea, src/matrix.cu:749-812, applied to the rows): a colour is then a contiguous row range, staged tile by
// tile with TMA bulk copies exactly like the block SpMV (k_block.cu: block4_tile_kernel), a quad per block row consuming from shared
// memory.  The two colour predicates of the sweeps are decided once at setup and travel in the top bits of the column index:
//   bit 30: the column has a LOWER colour than the row (forward: add delta_j to x_j)         [colors[j] < colors[i], j owned]
//   bit 31: the column has a HIGHER colour than the row and the row's colour is not 0        [the reference's `c != 0` guard]
// Summation order inside a row: storage order, one quad (the 8-quad butterfly of dilu_row_4x4 associates differently: results agree
// to rounding, as they do with the reference's own half-warp kernel).
// ---------------------------------------------------------------------------------------------
constexpr int DT_ROWS = 32, DT_CONSUMERS = DT_ROWS * 4, DT_STAGES = 2;
constexpr unsigned CS_LOWER = 0x40000000u, CS_HIGHER = 0x80000000u, CS_MASK = 0x3fffffffu;

struct DiluTileArgs {
    const int *rp, *ci, *rows;        // colour-sorted row pointers / flagged columns; original row id of every sorted row
    const void *va, *Einv, *b;
    void *x, *delta, *Delta;
    double weight;
    int row0, row1, num_tiles, cap;   // sorted-row range of the colour
};

__global__ void cs_lengths_kernel(const int *__restrict__ rp, const int *__restrict__ rows, int n, int *len)
{
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) { const int i = rows[p]; len[p] = rp[i + 1] - rp[i]; }
    if (blockIdx.x == 0 && threadIdx.x == 0) len[n] = 0;
}

template <class MatT>
__global__ void cs_permute_kernel(const int *__restrict__ rp, const int *__restrict__ ci, const MatT *__restrict__ va, const int *__restrict__ rows,
                                  const int *__restrict__ colors, const int *__restrict__ cs_rp, int n, int n_owned, int *cs_ci, MatT *cs_va)
{
    const int lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
    for (int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < n; p += warps) {
        const int i = rows[p], k0 = rp[i], len = rp[i + 1] - k0, d0 = cs_rp[p], ci_color = colors[i];
        for (int k = lane; k < len; k += 32) {
            const int j = ci[k0 + k];
            unsigned c = (unsigned)j;
            if (j < n_owned && j != i) {
                const int cj = colors[j];
                if (cj < ci_color) c |= CS_LOWER;
                else if (cj > ci_color && ci_color != 0) c |= CS_HIGHER;
            }
            cs_ci[d0 + k] = (int)c;
        }
        const size_t s0 = (size_t)k0 * 16, t0 = (size_t)d0 * 16, total = (size_t)len * 16;
        for (size_t t = lane; t < total; t += 32) cs_va[t0 + t] = va[s0 + t];
    }
}

__global__ void cs_tile_stats_kernel(const int *rp, int row0, int row1, int num_tiles, int *max_tile_nnz)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < num_tiles; t += gridDim.x * blockDim.x) {
        const int r0 = row0 + t * DT_ROWS, r1 = min(r0 + DT_ROWS, row1);
        atomicMax(max_tile_nnz, ((rp[r1] + 3) & ~3) - (rp[r0] & ~3));
    }
}

template <class MatT, class VecT, bool BACKWARD>
__global__ void __launch_bounds__(DT_CONSUMERS + PRODUCER_THREADS) dilu_tile_kernel(const DiluTileArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + MAX_STAGES;
    unsigned char *stage_base = smem_raw + 128;
    const size_t vals_bytes = (size_t)a.cap * 16 * sizeof(MatT);
    const size_t cols_bytes = (size_t)a.cap * sizeof(int);
    const size_t rp_bytes = (size_t)(DT_ROWS + 4) * sizeof(int);
    const size_t stage_bytes = vals_bytes + cols_bytes + rp_bytes;
    const MatT *__restrict__ va = (const MatT *)a.va;
    const MatT *__restrict__ Einv = (const MatT *)a.Einv;
    VecT *x = (VecT *)a.x, *delta = (VecT *)a.delta, *Delta = (VecT *)a.Delta;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < DT_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], DT_CONSUMERS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int my_tiles = (a.num_tiles > (int)blockIdx.x) ? (a.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (tid >= DT_CONSUMERS) {
        if (tid == DT_CONSUMERS) {
            for (int it = 0; it < my_tiles; it++) {
                const int tile = blockIdx.x + it * gridDim.x;
                const int s = it % DT_STAGES;
                const unsigned ph = (unsigned)(it / DT_STAGES) & 1u;
                if (it >= DT_STAGES) mbar_wait(&empty[s], ph ^ 1u);
                const int r0 = a.row0 + tile * DT_ROWS, r1 = min(r0 + DT_ROWS, a.row1);
                const int nz0 = __ldg(a.rp + r0), nz1 = __ldg(a.rp + r1);
                const int sa = nz0 & ~3, ea = (nz1 + 3) & ~3;
                unsigned char *st = stage_base + (size_t)s * stage_bytes;
                // the row_ptr slice starts at an arbitrary sorted row: copy from the 16-byte aligned address below it
                const int ra = r0 & ~3;
                const unsigned rp_copy = (unsigned)(((r1 - ra + 1 + 3) & ~3) * sizeof(int));
                const unsigned cnt = (unsigned)(ea - sa);
                mbar_expect_tx(&full[s], rp_copy + cnt * (unsigned)(16 * sizeof(MatT) + sizeof(int)));
                tma_bulk_g2s(st + vals_bytes + cols_bytes, a.rp + ra, rp_copy, &full[s]);
                if (cnt) {
                    tma_bulk_g2s(st, va + (size_t)sa * 16, cnt * (unsigned)(16 * sizeof(MatT)), &full[s]);
                    tma_bulk_g2s(st + vals_bytes, a.ci + sa, cnt * (unsigned)sizeof(int), &full[s]);
                }
            }
        }
    } else {
        const int r = tid & 3, lane = tid & 31, qbase = lane & ~3, q = tid >> 2;
        for (int it = 0; it < my_tiles; it++) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int s = it % DT_STAGES;
            const unsigned ph = (unsigned)(it / DT_STAGES) & 1u;
            const int r0 = a.row0 + tile * DT_ROWS;
            const int p = r0 + q;
            const bool act = p < a.row1;
            const int i = act ? __ldg(a.rows + p) : 0;
            VecT acc = 0;
            if (act && !BACKWARD) acc = ((const VecT *)a.b)[(size_t)i * 4 + r];
            const unsigned char *st = stage_base + (size_t)s * stage_bytes;
            const MatT *vals = reinterpret_cast<const MatT *>(st);
            const unsigned *cols = reinterpret_cast<const unsigned *>(st + vals_bytes);
            const int *rp = reinterpret_cast<const int *>(st + vals_bytes + cols_bytes) + (r0 & 3);
            mbar_wait(&full[s], ph);
            if (act) {
                const int sa = rp[0] & ~3;
                const int k0 = rp[q] - sa, kend = rp[q + 1] - sa;
                // four blocks per step: their gathers of x (and delta / Delta) are in flight together -- a quad walking its row one block
                // at a time is a chain of L2 round trips (r02: 2.0 ms per sweep on 4.1 M rows that way)
                if (!BACKWARD) {
                    for (int k = k0; k < kend; k += 4) {
                        unsigned c[4];
                        VecT xv[4][4], dv[4][4];
#pragma unroll
                        for (int u = 0; u < 4; u++) c[u] = cols[(k + u < kend) ? k + u : k0];
#pragma unroll
                        for (int u = 0; u < 4; u++) ld4v<false>(x + (size_t)(c[u] & CS_MASK) * 4, xv[u]);
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (k + u < kend && (c[u] & CS_LOWER)) ld4v<false>(delta + (size_t)(c[u] & CS_MASK) * 4, dv[u]);
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            if (k + u < kend) {
                                MatT av[4];
                                ld4v<false>(vals + (size_t)(k + u) * 16 + r * 4, av);
                                if (c[u] & CS_LOWER) {
#pragma unroll
                                    for (int m = 0; m < 4; m++) xv[u][m] += dv[u][m];
                                }
#pragma unroll
                                for (int m = 0; m < 4; m++) acc -= (VecT)av[m] * xv[u][m];
                            }
                        }
                    }
                } else {
                    for (int k = k0; k < kend; k += 4) {
                        unsigned c[4];
                        VecT Dv[4][4];
#pragma unroll
                        for (int u = 0; u < 4; u++) c[u] = (k + u < kend) ? cols[k + u] : 0u;
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (c[u] & CS_HIGHER) ld4v<false>(Delta + (size_t)(c[u] & CS_MASK) * 4, Dv[u]);
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            if (c[u] & CS_HIGHER) {
                                MatT av[4];
                                ld4v<false>(vals + (size_t)(k + u) * 16 + r * 4, av);
#pragma unroll
                                for (int m = 0; m < 4; m++) acc += (VecT)av[m] * Dv[u][m];
                            }
                        }
                    }
                }
            }
            // y = Einv_i * acc inside the quad
            VecT y = 0;
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const VecT am = __shfl_sync(0xffffffffu, acc, qbase + m);
                if (act) y += (VecT)Einv[(size_t)i * 16 + r * 4 + m] * am;
            }
            if (act) {
                const size_t idx = (size_t)i * 4 + r;
                if (!BACKWARD) delta[idx] = y;
                else {
                    const VecT v = delta[idx] - y;
                    x[idx] = x[idx] + (VecT)(a.weight * v);
                    Delta[idx] = v;
                }
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty[s]);
        }
    }
}

template <class MatT, class VecT, bool BACKWARD> void launch_dilu_tile(const DiluTileArgs &ta, int grid, size_t smem, cudaStream_t s)
{
    auto k = dilu_tile_kernel<MatT, VecT, BACKWARD>;
    smem_opt_in(reinterpret_cast<const void *>(k), smem);      // exactly what this kernel needs, once per size and device
    k<<<grid, DT_CONSUMERS + PRODUCER_THREADS, smem, s>>>(ta);
}

}  // namespace

class MulticolorDILUSolver : public Solver {
public:
    MulticolorDILUSolver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc) : Solver(cfg, scope, std::move(rsc))
    {
        weight_ = cfg.get_double("relaxation_factor", scope);
        if (weight_ == 0) {
            weight_ = 1.;
            amgx_printf("Warning, setting weight to 1 instead of estimating largest_eigen_value in Multicolor DILU smoother\n");
        }
        scheme_ = cfg.get_string("matrix_coloring_scheme", scope);
        if (scheme_ != "MIN_MAX" && scheme_ != "PARALLEL_GREEDY")
            fatal(AMGX_RC_BAD_CONFIGURATION, "matrix_coloring_scheme '" + scheme_ + "' is not supported by this engine (MIN_MAX, PARALLEL_GREEDY, or AMGX_matrix_attach_coloring)");
        if (cfg.get_int("coloring_level", scope) != 1) fatal(AMGX_RC_BAD_CONFIGURATION, "MULTICOLOR_DILU: coloring_level must be 1");
        // reorder_cols_by_color / insert_diag_while_reordering (src/matrix.cu:749-812): the reference sorts the entries of every row by the colour of
        // their column so that its sweeps can split a row into "earlier colours | later colours" without reading the colour array.  A memory
        // layout, not an algorithm: the sweeps here look up the colour of every column (or work on their own colour-sorted copy), whatever the
        // caller's entry order is, so the two
        // switches are accepted and change nothing (the caller's matrix is not permuted; sums differ from the reference's by their rounding only).
        uncolored_fraction_ = cfg.get_int("determinism_flag", "default") ? 0.0 : cfg.get_double("max_uncolored_percentage", scope);
    }
    bool is_coloring_needed() const override { return true; }
    const DevVec *smoother_data() const override { return &Einv_; }
    void smooth(DevVec &b, DevVec &x, bool xIsZero, int sweeps, const SmoothFuse *fuse, bool input_in_alt = false) override
    {
        if (input_in_alt || (fuse && (fuse->agg || fuse->dot_b_x))) fatal(AMGX_RC_INTERNAL, "DILU does not support fused sweeps");
        if (fused_level_ && !A_->dist && sweeps > 0) { level_sweeps(b, x, xIsZero, sweeps); return; }     // all colours, all sweeps: one launch
        for (int it = 0; it < sweeps; it++) sweep(b, x, xIsZero && it == 0);
    }

protected:
    void solver_setup(bool) override
    {
        Matrix &A = *A_;
        if (A.bx != A.by) fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "DILU implemented only for squared blocks");
        if (A.bs() != 1 && A.bx != 4) fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "Multicolor-DILU: block sizes 1 and 4 are enabled in this engine");
        if (A.has_ext_diag) fatal(AMGX_RC_NOT_IMPLEMENTED, "Multicolor-DILU with an external diagonal");
        cudaStream_t s = stream();
        if (A.num_colors == 0) color_matrix(A, scheme_, uncolored_fraction_, s);
        const size_t bs = A.bs();
        Einv_.resize((size_t)A.n_cols * bs, A.mat_prec);
        Einv_.zero(s);
        delta_.resize((size_t)A.n_cols * A.by, A.vec_prec);
        Delta_.resize((size_t)A.n_cols * A.by, A.vec_prec);
        delta_.zero(s);
        Delta_.zero(s);
        for (int c = 0; c < A.num_colors; c++) {
            const int off = A.color_offsets[c], cnt = A.color_offsets[c + 1] - off;
            if (cnt == 0) continue;
            AMGXB_DISPATCH(A.mat_prec, A.vec_prec, {
                if (bs == 1) {
                    const int grid = std::min(4096, ceil_div(cnt, 4));
                    dilu_setup_1x1<MatT, VecT><<<grid, 128, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), A.diag_idx.ptr(), A.values.as<MatT>(), Einv_.as<MatT>(),
                                                                    A.sorted_rows_by_color.ptr() + off, A.row_colors.ptr(), cnt, c, A.n);
                } else {
                    const int grid = std::min(4096, ceil_div(cnt, 128));
                    dilu_setup_4x4<MatT, VecT><<<grid, 128, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), A.diag_idx.ptr(), A.values.as<MatT>(), Einv_.as<MatT>(),
                                                                    A.sorted_rows_by_color.ptr() + off, A.row_colors.ptr(), cnt, c, A.n);
                }
            });
            count_launch();
            AMGXB_LAUNCH_CHECK();
        }
        // Levels whose LARGEST colour fits the warps of one cluster in a single pass take the fused kernel (dilu_level_kernel): there a
        // colour costs one barrier instead of one launch.  Larger colours need the whole machine: a cluster has 8 x 32 warps, a per-colour
        // launch up to 148 x 64 (r02: fusing levels of up to 32768 rows made config 5 slower, 53 -> 46 it/s).
        static const int fused_on = getenv("AMGXB_DILU_FUSED") ? atoi(getenv("AMGXB_DILU_FUSED")) : 1;
        int max_color = 0;
        for (int c = 0; c < A.num_colors; c++) max_color = std::max(max_color, A.color_offsets[c + 1] - A.color_offsets[c]);
        const int rows_per_warp = (bs == 1) ? 32 / NTPR : 1;
        fused_level_ = fused_on && A.n > 0 && A.num_colors > 0 && max_color <= 8 * 32 * rows_per_warp;
        if (fused_level_) {
            d_color_offsets_.from_any(A.color_offsets.data(), A.color_offsets.size(), s);
            cluster_size_ = max_color <= 32 * rows_per_warp ? 1 : 8;     // one CTA, or a cluster of 8 (the portable maximum)
        }
        build_color_sorted();
    }

    // colour-sorted copy of a 4x4 block matrix for the tile kernels (large levels only: the fused level kernel takes the small ones)
    void build_color_sorted()
    {
        Matrix &A = *A_;
        cudaStream_t s = stream();
        cs_ready_ = false;
        static const int tiles_on = getenv("AMGXB_DILU_TILES") ? atoi(getenv("AMGXB_DILU_TILES")) : 1;
        // short rows only: a quad walks its row alone here; the 30-60-block rows of the coarse levels keep the 8-quads-per-row kernel
        // ... and fp32 blocks only: with 128-byte blocks three CTAs fit an SM and the tile form loses to form 1 (r02: 45.0 vs 47.6 it/s in dDDI,
        // 63.6 vs 60.2 in dDFI on the 160^3 problem)
        if (!tiles_on || fused_level_ || A.bs() != 16 || A.n == 0 || (long long)A.n_cols >= (1ll << 30) || (double)A.nnz > 12.0 * A.n || A.mat_prec != Prec::F32) return;
        const int n = A.n;
        DevBuf<int> len;
        len.resize((size_t)n + 1);
        cs_lengths_kernel<<<std::max(1, std::min(ceil_div(n, 256), 2048)), 256, 0, s>>>(A.row_ptr.ptr(), A.sorted_rows_by_color.ptr(), n, len.ptr());
        cs_rp_.resize((size_t)n + 1);
        size_t tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, len.ptr(), cs_rp_.ptr(), n + 1, s);
        DevBytes tmp;
        tmp.resize(tmp_bytes);
        cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, len.ptr(), cs_rp_.ptr(), n + 1, s);
        cs_ci_.resize((size_t)std::max(A.nnz, 1) + 8);
        cs_va_.resize(((size_t)std::max(A.nnz, 1) + 8) * 16, A.mat_prec);
        const int grid = std::max(1, std::min(ceil_div(n, 8), B200_SMS * 16));
        if (A.mat_prec == Prec::F64)
            cs_permute_kernel<double><<<grid, 256, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), A.values.as<double>(), A.sorted_rows_by_color.ptr(), A.row_colors.ptr(),
                                                           cs_rp_.ptr(), n, A.n, cs_ci_.ptr(), cs_va_.as<double>());
        else
            cs_permute_kernel<float><<<grid, 256, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), A.values.as<float>(), A.sorted_rows_by_color.ptr(), A.row_colors.ptr(),
                                                          cs_rp_.ptr(), n, A.n, cs_ci_.ptr(), cs_va_.as<float>());
        count_launch(4);
        // stage capacity: the largest tile of any colour
        DevBuf<int> st;
        st.resize(1);
        st.zero(s);
        for (int c = 0; c < A.num_colors; c++) {
            const int off = A.color_offsets[c], cnt = A.color_offsets[c + 1] - off;
            if (cnt == 0) continue;
            const int nt = ceil_div(cnt, DT_ROWS);
            cs_tile_stats_kernel<<<std::max(1, std::min(ceil_div(nt, 256), 1024)), 256, 0, s>>>(cs_rp_.ptr(), off, off + cnt, nt, st.ptr());
            count_launch();
        }
        AMGXB_LAUNCH_CHECK();
        cs_cap_ = std::max(4, st.to_host(s)[0]);
        cs_smem_ = 128 + (size_t)DT_STAGES * ((size_t)cs_cap_ * (16 * prec_size(A.mat_prec) + 4) + (size_t)(DT_ROWS + 4) * 4);
        if (cs_smem_ > (size_t)200 * 1024) { cs_rp_.release(); cs_ci_.release(); return; }
        const int regs = (A.mat_prec == Prec::F64) ? 96 : 64;      // cuobjdump -res-usage, forward kernel (four blocks of gathers in flight)
        const int by_threads = std::min(2048 / (DT_CONSUMERS + PRODUCER_THREADS), 65536 / ((DT_CONSUMERS + PRODUCER_THREADS) * regs));
        cs_ctas_ = std::max(1, std::min(by_threads, (int)((size_t)227 * 1024 / (cs_smem_ + 1024))));
        cs_ready_ = true;
    }

    // one colour of a sweep through the colour-sorted copy
    void tile_color(DevVec &b, DevVec &x, int c, bool backward)
    {
        Matrix &A = *A_;
        cudaStream_t s = stream();
        const int off = A.color_offsets[c], cnt = A.color_offsets[c + 1] - off;
        DiluTileArgs ta;
        ta.rp = cs_rp_.ptr();
        ta.ci = cs_ci_.ptr();
        ta.rows = A.sorted_rows_by_color.ptr();
        ta.va = cs_va_.ptr();
        ta.Einv = Einv_.ptr();
        ta.b = b.ptr();
        ta.x = x.ptr();
        ta.delta = delta_.ptr();
        ta.Delta = Delta_.ptr();
        ta.weight = weight_;
        ta.row0 = off;
        ta.row1 = off + cnt;
        ta.num_tiles = ceil_div(cnt, DT_ROWS);
        ta.cap = cs_cap_;
        const int grid = std::max(1, std::min(ta.num_tiles, (A.rsc ? A.rsc->num_sms : 148) * cs_ctas_));
        AMGXB_DISPATCH(A.mat_prec, A.vec_prec, {
            if (backward) launch_dilu_tile<MatT, VecT, true>(ta, grid, cs_smem_, s);
            else launch_dilu_tile<MatT, VecT, false>(ta, grid, cs_smem_, s);
        });
        count_launch();
    }

    // forward + backward over all colours, `sweeps` times, in ONE launch (one thread-block cluster)
    void level_sweeps(DevVec &b, DevVec &x, bool xIsZero, int sweeps)
    {
        Matrix &A = *A_;
        cudaStream_t s = stream();
        DiluLevelArgs a;
        a.rp = A.row_ptr.ptr();
        a.ci = A.col_idx.ptr();
        a.colors = A.row_colors.ptr();
        a.sorted_rows = A.sorted_rows_by_color.ptr();
        a.color_offsets = d_color_offsets_.ptr();
        a.va = A.values.ptr();
        a.Einv = Einv_.ptr();
        a.b = b.ptr();
        a.x = x.ptr();
        a.delta = delta_.ptr();
        a.Delta = Delta_.ptr();
        a.weight = weight_;
        a.n = A.n;
        a.n_owned = A.n;
        a.num_colors = A.num_colors;
        a.sweeps = sweeps;
        a.zero_first = xIsZero ? 1 : 0;
        a.bs = A.bs();
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)cluster_size_, 1, 1);
        cfg.blockDim = dim3(1024, 1, 1);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cluster_size_;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        AMGXB_DISPATCH(A.mat_prec, A.vec_prec, {
            if (A.bs() == 1) AMGXB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, dilu_level_kernel<MatT, VecT, 1>, a));
            else AMGXB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, dilu_level_kernel<MatT, VecT, 4>, a));
        });
        count_launch();
        AMGXB_LAUNCH_CHECK();
    }

    void sweep(DevVec &b, DevVec &x, bool xIsZero)
    {
        Matrix &A = *A_;
        cudaStream_t s = stream();
        if (xIsZero) x.zero(s);
        else if (A.dist) {
            // row-partitioned: halo values of x are frozen for the sweep (the forward pass reads them in b - A x); the
            // coloured corrections couple owned rows only -- the reference's SYNC_COLORS forward predicate
            // (a_col_id < boundary_index, multicolor_dilu_solver.cu:1848-1858).  Halo Delta is never formed: the sweep
            // is x += w M^-1 (b - A x) with M the DILU factorisation of this rank's diagonal block.
            dist_exchange_halo(A, x, s);
            dist_wait_halo(A, s);
        }
        const int nc = A.num_colors;
        const size_t bs = A.bs();
        AMGXB_DISPATCH(A.mat_prec, A.vec_prec, {
            for (int c = 0; c < nc; c++) {
                const int off = A.color_offsets[c], cnt = A.color_offsets[c + 1] - off;
                if (cnt == 0) continue;
                if (bs == 1) {
                    const int grid = std::min(4096, ceil_div(cnt, 128 / NTPR));
                    dilu_forward_1x1<MatT, VecT><<<grid, 128, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), A.values.as<MatT>(), x.as<VecT>(), b.as<VecT>(), delta_.as<VecT>(),
                                                                      A.sorted_rows_by_color.ptr() + off, cnt, c, A.row_colors.ptr(), Einv_.as<MatT>(), A.n);
                } else if (cs_ready_) {
                    tile_color(b, x, c, false);
                    continue;
                } else {
                    const int grid = std::min(B200_SMS * 16, ceil_div(cnt, 4));
                    dilu_sweep_4x4<MatT, VecT, false><<<grid, 128, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), A.values.as<MatT>(), x.as<VecT>(), b.as<VecT>(), delta_.as<VecT>(),
                                                                           Delta_.as<VecT>(), weight_, A.sorted_rows_by_color.ptr() + off, cnt, c, A.row_colors.ptr(),
                                                                           Einv_.as<MatT>(), A.n);
                }
                count_launch();
            }
            for (int c = nc - 1; c >= 0; c--) {
                const int off = A.color_offsets[c], cnt = A.color_offsets[c + 1] - off;
                if (cnt == 0) continue;
                if (c == nc - 1) {
                    const int grid = std::min(4096, ceil_div((long long)cnt * bs, 128));
                    dilu_backward_skip<VecT><<<grid, 128, 0, s>>>(x.as<VecT>(), weight_, A.sorted_rows_by_color.ptr() + off, delta_.as<VecT>(), Delta_.as<VecT>(), cnt, A.by);
                } else if (bs == 1) {
                    const int grid = std::min(4096, ceil_div(cnt, 128 / NTPR));
                    dilu_backward_1x1<MatT, VecT><<<grid, 128, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), A.values.as<MatT>(), x.as<VecT>(), weight_,
                                                                       A.sorted_rows_by_color.ptr() + off, A.row_colors.ptr(), Einv_.as<MatT>(), delta_.as<VecT>(),
                                                                       Delta_.as<VecT>(), cnt, c, A.n);
                } else if (cs_ready_) {
                    tile_color(b, x, c, true);
                    continue;
                } else {
                    const int grid = std::min(B200_SMS * 16, ceil_div(cnt, 4));
                    dilu_sweep_4x4<MatT, VecT, true><<<grid, 128, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), A.values.as<MatT>(), x.as<VecT>(), b.as<VecT>(), delta_.as<VecT>(),
                                                                          Delta_.as<VecT>(), weight_, A.sorted_rows_by_color.ptr() + off, cnt, c, A.row_colors.ptr(),
                                                                          Einv_.as<MatT>(), A.n);
                }
                count_launch();
            }
        });
        AMGXB_LAUNCH_CHECK();
    }

    Status solve_iteration(DevVec &b, DevVec &x, bool xIsZero) override
    {
        sweep(b, x, xIsZero);
        return converged(b, x);
    }

    double weight_ = 0.9, uncolored_fraction_ = 0.15;
    std::string scheme_ = "MIN_MAX";
    DevVec Einv_, delta_, Delta_;
    bool fused_level_ = false;
    int cluster_size_ = 1;
    DevBuf<int> d_color_offsets_;
    // colour-sorted copy for the tile kernels
    bool cs_ready_ = false;
    DevBuf<int> cs_rp_, cs_ci_;
    DevVec cs_va_;
    int cs_cap_ = 0, cs_ctas_ = 1;
    size_t cs_smem_ = 0;
};

std::unique_ptr<Solver> make_dilu_solver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc)
{
    return std::unique_ptr<Solver>(new MulticolorDILUSolver(cfg, scope, std::move(rsc)));
}

}  // namespace amgxb
