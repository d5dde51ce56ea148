// k_block.cu -- 4x4 block-CSR kernels: SpMV / residual / fused block-Jacobi sweep, diagonal-block inverses, per-component
// norms.  Mixed precision (matrix fp32, vectors fp64: mode dDFI) is supported by every kernel.
//   blockDiaCsrMultiplyKernel(_DiaProps)_4x4      src/multiply.cu:335-547
//   jacobiSmooth4by4BlockDiaCsrKernel_..._Dinv2  src/solvers/block_jacobi_solver.cu:665-773  (same operation order here)
//   setupBlockJacobiSmooth4by4BlockDiaCsrKernel   src/solvers/block_jacobi_solver.cu:637-662 + block_common_solver.h:106-137
//   strided_reduction (block norms)               include/strided_reduction.h, src/norm.cu:308-405
// Layout: a quad of threads owns one block row; thread r of the quad owns row r of every 4x4 block of that block row,
// so each block is read as four contiguous 16- or 32-byte vectors by one quad (full sectors, no shared-memory staging
// needed), x_j is read as one 16/32-byte vector per thread, and the per-thread FMA chain runs left to right.
#include "solvers.h"
#include "dist.h"
#include <cmath>

namespace amgxb {
namespace {

#include "tile_common.cuh"

template <class T> __device__ __forceinline__ T guardz(T d);
template <> __device__ __forceinline__ double guardz<double>(double d) { return fabs(d) < 1e-12 ? copysign(1e-12, d) : d; }
template <> __device__ __forceinline__ float guardz<float>(float d) { return fabs((double)d) < 1e-7 ? copysignf((float)1e-7, d) : d; }

template <class T> struct Vec4 { T v[4]; };
__device__ __forceinline__ Vec4<double> ld4(const double *p)
{
    const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
    return Vec4<double>{{a.x, a.y, b.x, b.y}};
}
__device__ __forceinline__ Vec4<float> ld4(const float *p)
{
    const float4 a = *reinterpret_cast<const float4 *>(p);
    return Vec4<float>{{a.x, a.y, a.z, a.w}};
}

__global__ void find_block_diag_kernel(const int *rp, const int *ci, int n, int nnz, int has_ext, int *diag)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int d = -1;
        if (has_ext) d = nnz + i;
        else for (int k = rp[i]; k < rp[i + 1]; k++) if (ci[k] == i) { d = k; break; }
        diag[i] = d;
    }
}

enum { B_SPMV = 0, B_RESID = 1, B_JACOBI = 2 };

template <class MatT, class VecT, int MODE, bool EXT_DIAG>
__global__ void __launch_bounds__(256) block4_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const int *__restrict__ diag,
                                                     const MatT *__restrict__ va, const VecT *__restrict__ x, const VecT *__restrict__ b,
                                                     const MatT *__restrict__ dinv, VecT *__restrict__ y, double omega)
{
    const int r = threadIdx.x & 3, lane = threadIdx.x & 31, qbase = lane & ~3;
    const int quads_per_grid = gridDim.x * (blockDim.x >> 2);
    for (int i = blockIdx.x * (blockDim.x >> 2) + (threadIdx.x >> 2); __any_sync(0xffffffffu, i < n); i += quads_per_grid) {
        const bool act = i < n;
        VecT acc = 0, xin = 0;
        if (act) {
            const int k0 = rp[i], k1 = rp[i + 1];
            if (MODE == B_JACOBI) {
                // bmAx = b ; diagonal block first ; then the other blocks in row order (block_jacobi_solver.cu:686-722)
                acc = b[(size_t)i * 4 + r];
                xin = x[(size_t)i * 4 + r];
                const int d = diag[i];
                if (d >= 0) {
                    const Vec4<MatT> a = ld4(va + (size_t)d * 16 + r * 4);
                    const Vec4<VecT> xv = ld4(x + (size_t)i * 4);
#pragma unroll
                    for (int m = 0; m < 4; m++) acc = fma(-(VecT)a.v[m], xv.v[m], acc);
                }
                for (int k = k0; k < k1; k++) {
                    const int j = ci[k];
                    if (j == i) continue;
                    const Vec4<MatT> a = ld4(va + (size_t)k * 16 + r * 4);
                    const Vec4<VecT> xv = ld4(x + (size_t)j * 4);
#pragma unroll
                    for (int m = 0; m < 4; m++) acc = fma(-(VecT)a.v[m], xv.v[m], acc);
                }
            } else {
                for (int k = k0; k < k1; k++) {
                    const int j = ci[k];
                    const Vec4<MatT> a = ld4(va + (size_t)k * 16 + r * 4);
                    const Vec4<VecT> xv = ld4(x + (size_t)j * 4);
#pragma unroll
                    for (int m = 0; m < 4; m++) acc = fma((VecT)a.v[m], xv.v[m], acc);
                }
                if (EXT_DIAG) {   // external diagonal: second pass with the identity column map (amgx_cusparse.cu:574-599)
                    const Vec4<MatT> a = ld4(va + (size_t)diag[i] * 16 + r * 4);
                    const Vec4<VecT> xv = ld4(x + (size_t)i * 4);
#pragma unroll
                    for (int m = 0; m < 4; m++) acc = fma((VecT)a.v[m], xv.v[m], acc);
                }
            }
        }
        if (MODE == B_JACOBI) {
            VecT t = 0;
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const VecT am = __shfl_sync(0xffffffffu, acc, qbase + m);
                if (act) t = fma((VecT)dinv[(size_t)i * 16 + r * 4 + m], am, t);
            }
            if (act) y[(size_t)i * 4 + r] = fma(t, (VecT)omega, xin);     // xin + bmAx * weight
        } else if (act) {
            if (MODE == B_SPMV) y[(size_t)i * 4 + r] = acc;
            else y[(size_t)i * 4 + r] = b[(size_t)i * 4 + r] - acc;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// TMA-staged 4x4 block tile kernel (r02).  Same arithmetic as block4_kernel above (a quad per block row, thread r owns row r of every
// block, blocks in storage order, 4 FMAs per block in component order => the same bits), but the matrix reaches the SM the way the
// scalar tile kernel's does: a producer warp bulk-copies the tile's row_ptr slice, its column indices and its 64- / 128-byte blocks
// into shared memory (cp.async.bulk + mbarrier, 2 stages), so HBM sees long contiguous reads instead of one 64-byte request per
// quad and step, and the consumers' only global loads are the gathers of x.  Tile = 32 block rows (128 consumer threads).
// Replaces blockDiaCsrMultiplyKernel_4x4 (src/multiply.cu:400-547) and the fused 4x4 Jacobi (block_jacobi_solver.cu:665-739).
// ---------------------------------------------------------------------------------------------
constexpr int BT_ROWS = 32;                         // block rows per tile
constexpr int BT_CONSUMERS = BT_ROWS * 4;           // a quad per block row
constexpr int BT_STAGES = 2;

struct BlockTileArgs {
    const int *rp, *ci, *diag;
    const void *va, *x, *b, *dinv;
    void *y;
    double omega;
    int n, num_tiles, cap;                          // cap: blocks a stage holds (multiple of 4)
};

template <class MatT, class VecT, int MODE>
__global__ void __launch_bounds__(BT_CONSUMERS + PRODUCER_THREADS) block4_tile_kernel(const BlockTileArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + MAX_STAGES;
    unsigned char *stage_base = smem_raw + 128;
    const size_t vals_bytes = (size_t)a.cap * 16 * sizeof(MatT);
    const size_t cols_bytes = (size_t)a.cap * sizeof(int);
    const size_t rp_bytes = (size_t)(BT_ROWS + 4) * sizeof(int);
    const size_t stage_bytes = vals_bytes + cols_bytes + rp_bytes;
    const MatT *__restrict__ va = (const MatT *)a.va;
    const VecT *__restrict__ x = (const VecT *)a.x;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < BT_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], BT_CONSUMERS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int my_tiles = (a.num_tiles > (int)blockIdx.x) ? (a.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (tid >= BT_CONSUMERS) {
        if (tid == BT_CONSUMERS) {
            for (int it = 0; it < my_tiles; it++) {
                const int tile = blockIdx.x + it * gridDim.x;
                const int s = it % BT_STAGES;
                const unsigned ph = (unsigned)(it / BT_STAGES) & 1u;
                if (it >= BT_STAGES) mbar_wait(&empty[s], ph ^ 1u);
                const int r0 = tile * BT_ROWS, r1 = min(r0 + BT_ROWS, a.n);
                const int nz0 = __ldg(a.rp + r0), nz1 = __ldg(a.rp + r1);
                const int sa = nz0 & ~3, ea = (nz1 + 3) & ~3;
                unsigned char *st = stage_base + (size_t)s * stage_bytes;
                const unsigned rp_copy = (unsigned)(((r1 - r0 + 1 + 3) & ~3) * sizeof(int));
                const unsigned cnt = (unsigned)(ea - sa);
                mbar_expect_tx(&full[s], rp_copy + cnt * (unsigned)(16 * sizeof(MatT) + sizeof(int)));
                tma_bulk_g2s(st + vals_bytes + cols_bytes, a.rp + r0, rp_copy, &full[s]);
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
            const int s = it % BT_STAGES;
            const unsigned ph = (unsigned)(it / BT_STAGES) & 1u;
            const int i = tile * BT_ROWS + q;
            const bool act = i < a.n;
            VecT acc = 0, xin = 0, bi = 0;
            int d = -1;
            if (act) {
                if (MODE != B_SPMV) bi = ((const VecT *)a.b)[(size_t)i * 4 + r];
                if (MODE == B_JACOBI) { xin = x[(size_t)i * 4 + r]; d = __ldg(a.diag + i); }
            }
            const unsigned char *st = stage_base + (size_t)s * stage_bytes;
            const MatT *vals = reinterpret_cast<const MatT *>(st);
            const int *cols = reinterpret_cast<const int *>(st + vals_bytes);
            const int *rp = reinterpret_cast<const int *>(st + vals_bytes + cols_bytes);
            mbar_wait(&full[s], ph);
            if (act) {
                const int sa = rp[0] & ~3;
                int k = rp[q] - sa;
                const int kend = rp[q + 1] - sa;
                if (MODE == B_JACOBI) {
                    // bmAx = b ; diagonal block first ; then the other blocks in row order (block_jacobi_solver.cu:686-722)
                    acc = bi;
                    if (d >= 0) {
                        const Vec4<MatT> av = ld4(vals + (size_t)(d - sa) * 16 + r * 4);
                        const Vec4<VecT> xv = ld4(x + (size_t)i * 4);
#pragma unroll
                        for (int m = 0; m < 4; m++) acc = fma(-(VecT)av.v[m], xv.v[m], acc);
                    }
                }
                // two blocks per step: their gathers of x are in flight together
                for (; k + 2 <= kend; k += 2) {
                    const int j0 = cols[k], j1 = cols[k + 1];
                    const Vec4<VecT> x0 = ld4(x + (size_t)j0 * 4), x1 = ld4(x + (size_t)j1 * 4);
                    const Vec4<MatT> a0 = ld4(vals + (size_t)k * 16 + r * 4), a1 = ld4(vals + (size_t)(k + 1) * 16 + r * 4);
                    if (MODE == B_JACOBI) {
                        if (j0 != i) {
#pragma unroll
                            for (int m = 0; m < 4; m++) acc = fma(-(VecT)a0.v[m], x0.v[m], acc);
                        }
                        if (j1 != i) {
#pragma unroll
                            for (int m = 0; m < 4; m++) acc = fma(-(VecT)a1.v[m], x1.v[m], acc);
                        }
                    } else {
#pragma unroll
                        for (int m = 0; m < 4; m++) acc = fma((VecT)a0.v[m], x0.v[m], acc);
#pragma unroll
                        for (int m = 0; m < 4; m++) acc = fma((VecT)a1.v[m], x1.v[m], acc);
                    }
                }
                if (k < kend) {
                    const int j0 = cols[k];
                    const Vec4<VecT> x0 = ld4(x + (size_t)j0 * 4);
                    const Vec4<MatT> a0 = ld4(vals + (size_t)k * 16 + r * 4);
                    if (MODE == B_JACOBI) {
                        if (j0 != i) {
#pragma unroll
                            for (int m = 0; m < 4; m++) acc = fma(-(VecT)a0.v[m], x0.v[m], acc);
                        }
                    } else {
#pragma unroll
                        for (int m = 0; m < 4; m++) acc = fma((VecT)a0.v[m], x0.v[m], acc);
                    }
                }
            }
            if (MODE == B_JACOBI) {
                VecT t = 0;
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const VecT am = __shfl_sync(0xffffffffu, acc, qbase + m);
                    if (act) t = fma((VecT)((const MatT *)a.dinv)[(size_t)i * 16 + r * 4 + m], am, t);
                }
                if (act) ((VecT *)a.y)[(size_t)i * 4 + r] = fma(t, (VecT)a.omega, xin);     // xin + bmAx * weight
            } else if (act) {
                if (MODE == B_SPMV) ((VecT *)a.y)[(size_t)i * 4 + r] = acc;
                else ((VecT *)a.y)[(size_t)i * 4 + r] = bi - acc;
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty[s]);
        }
    }
}

template <class MatT, class VecT, int MODE> void launch_block_tile(const BlockTileArgs &ta, int grid, size_t smem, cudaStream_t s)
{
    auto k = block4_tile_kernel<MatT, VecT, MODE>;
    smem_opt_in(reinterpret_cast<const void *>(k), smem);      // exactly what this kernel needs, once per size and device
    k<<<grid, BT_CONSUMERS + PRODUCER_THREADS, smem, s>>>(ta);
}

__global__ void block_tile_stats_kernel(const int *rp, int n, int num_tiles, int *max_tile_nnz)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < num_tiles; t += gridDim.x * blockDim.x) {
        const int r0 = t * BT_ROWS, r1 = min(r0 + BT_ROWS, n);
        atomicMax(max_tile_nnz, ((rp[r1] + 3) & ~3) - (rp[r0] & ~3));
    }
}

template <class MatT, class VecT> __global__ void block4_jacobi_zero(int n, const MatT *__restrict__ dinv, const VecT *__restrict__ b, VecT *__restrict__ x, double omega)
{
    const int r = threadIdx.x & 3;
    for (int i = blockIdx.x * (blockDim.x >> 2) + (threadIdx.x >> 2); i < n; i += gridDim.x * (blockDim.x >> 2)) {
        const Vec4<VecT> bv = ld4(b + (size_t)i * 4);
        VecT t = 0;
#pragma unroll
        for (int m = 0; m < 4; m++) t = fma((VecT)dinv[(size_t)i * 16 + r * 4 + m], bv.v[m], t);
        x[(size_t)i * 4 + r] = t * (VecT)omega;     // bmAx * weight
    }
}

// Dinv = inverse of each diagonal block: Gauss-Jordan without pivoting in the reference's operation order
template <class MatT> __global__ void block4_invert_diag(int n, MatT *d)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        MatT A[16];
        for (int m = 0; m < 16; m++) A[m] = d[(size_t)i * 16 + m];
        for (int row = 0; row < 4; row++) {
            const MatT diag = (MatT)1 / guardz<MatT>(A[row * 4 + row]);
            for (int j = 0; j < 4; j++) if (j != row) A[row * 4 + j] = A[row * 4 + j] * diag;
            for (int ii = 0; ii < 4; ii++) if (ii != row)
                for (int j = 0; j < 4; j++) if (j != row) A[ii * 4 + j] = fma(-A[ii * 4 + row], A[row * 4 + j], A[ii * 4 + j]);
            for (int j = 0; j < 4; j++) A[j * 4 + row] = (j == row) ? diag : -(A[j * 4 + row] * diag);
        }
        for (int m = 0; m < 16; m++) d[(size_t)i * 16 + m] = A[m];
    }
}

// per-component norms: one launch, bsize accumulators
template <class VecT, int NORM> __global__ void block_norm_kernel(const VecT *__restrict__ v, int n, int bsize, double *out)
{
    __shared__ double sm[8][32];
    double acc[8];
    for (int c = 0; c < bsize; c++) acc[c] = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        for (int c = 0; c < bsize; c++) {
            const double t = (double)v[(size_t)i * bsize + c];
            if (NORM == 1) acc[c] += t * t;
            else if (NORM == 0) acc[c] += fabs(t);
            else acc[c] = fmax(acc[c], fabs(t));
        }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = 0; c < bsize; c++) {
        double t = acc[c];
        for (int o = 16; o > 0; o >>= 1) { const double u = __shfl_xor_sync(0xffffffffu, t, o); t = (NORM == 2) ? fmax(t, u) : t + u; }
        if (lane == 0) sm[c][warp] = t;
    }
    __syncthreads();
    if (warp == 0)
        for (int c = 0; c < bsize; c++) {
            double t = (lane < (int)(blockDim.x >> 5)) ? sm[c][lane] : 0.0;
            for (int o = 16; o > 0; o >>= 1) { const double u = __shfl_xor_sync(0xffffffffu, t, o); t = (NORM == 2) ? fmax(t, u) : t + u; }
            if (lane == 0) out[(size_t)blockIdx.x * 8 + c] = t;
        }
}

}  // namespace

void block_build_diag(Matrix &A, cudaStream_t s)
{
    if (A.bx != A.by) fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "rectangular blocks are not supported");
    A.plan = TilePlan();
    A.diag_idx.resize(std::max(A.n, 1));
    if (A.n == 0) return;
    find_block_diag_kernel<<<std::min(ceil_div(A.n, 256), 4096), 256, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), A.n, A.nnz, A.has_ext_diag ? 1 : 0, A.diag_idx.ptr());
    count_launch();
    AMGXB_LAUNCH_CHECK();
    // plan of the TMA-staged tile kernel (4x4 blocks, diagonal inside the CSR structure)
    static const int tiles_on = getenv("AMGXB_BLOCK_TILES") ? atoi(getenv("AMGXB_BLOCK_TILES")) : 1;
    if (tiles_on && A.bx == 4 && A.by == 4 && !A.has_ext_diag) {
        TilePlan p;
        p.tile_rows = BT_ROWS;
        p.num_tiles = ceil_div(A.n, BT_ROWS);
        DevBuf<int> st;
        st.resize(1);
        st.zero(s);
        block_tile_stats_kernel<<<std::max(1, std::min(ceil_div(p.num_tiles, 256), 1024)), 256, 0, s>>>(A.row_ptr.ptr(), A.n, p.num_tiles, st.ptr());
        count_launch();
        p.max_tile_nnz = std::max(4, st.to_host(s)[0]);
        p.stages = BT_STAGES;
        p.smem_bytes = 128 + (size_t)BT_STAGES * ((size_t)p.max_tile_nnz * (16 * prec_size(A.mat_prec) + 4) + (size_t)(BT_ROWS + 4) * 4);
        const int by_threads = std::min(2048 / (BT_CONSUMERS + PRODUCER_THREADS), 65536 / ((BT_CONSUMERS + PRODUCER_THREADS) * 40));     // threads, registers (<= 40 / thread, cuobjdump -res-usage)
        p.ctas_per_sm = std::max(1, std::min(by_threads, (int)((size_t)227 * 1024 / (p.smem_bytes + 1024))));
        p.use_tiles = p.smem_bytes <= (size_t)200 * 1024;
        A.plan = p;
    }
}

void block_apply(const Matrix &A, CsrEpi epi, const CsrOpArgs &g, cudaStream_t s)
{
    if (A.bx != 4 || A.by != 4) fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "block kernels of this engine support 4x4 blocks (and scalar matrices)");
    if (g.agg) fatal(AMGX_RC_INTERNAL, "aggregated gather is a scalar-kernel feature");
    if (A.n == 0) return;
    if (A.plan.use_tiles && (epi == EPI_SPMV || epi == EPI_RESID || epi == EPI_JACOBI)) {
        BlockTileArgs ta;
        ta.rp = A.row_ptr.ptr();
        ta.ci = A.col_idx.ptr();
        ta.diag = A.diag_idx.ptr();
        ta.va = A.values.ptr();
        ta.x = g.x;
        ta.b = g.b;
        ta.dinv = g.d;
        ta.y = g.y;
        ta.omega = g.omega;
        ta.n = A.n;
        ta.num_tiles = A.plan.num_tiles;
        ta.cap = A.plan.max_tile_nnz;
        const int tgrid = std::max(1, std::min(A.plan.num_tiles, (A.rsc ? A.rsc->num_sms : 148) * A.plan.ctas_per_sm));
        const size_t smem = A.plan.smem_bytes;
        AMGXB_DISPATCH(A.mat_prec, A.vec_prec, {
            if (epi == EPI_SPMV) launch_block_tile<MatT, VecT, B_SPMV>(ta, tgrid, smem, s);
            else if (epi == EPI_RESID) launch_block_tile<MatT, VecT, B_RESID>(ta, tgrid, smem, s);
            else launch_block_tile<MatT, VecT, B_JACOBI>(ta, tgrid, smem, s);
        });
        count_launch();
        AMGXB_LAUNCH_CHECK();
        return;
    }
    const int grid = std::max(1, std::min(ceil_div(A.n, 64), (A.rsc ? A.rsc->num_sms : 148) * 8));
    AMGXB_DISPATCH(A.mat_prec, A.vec_prec, {
        const MatT *va = A.values.as<MatT>();
        const VecT *x = (const VecT *)g.x;
        const VecT *b = (const VecT *)g.b;
        VecT *y = (VecT *)g.y;
        const MatT *d = (const MatT *)g.d;
        switch (epi) {
        case EPI_SPMV:
            if (A.has_ext_diag) block4_kernel<MatT, VecT, B_SPMV, true><<<grid, 256, 0, s>>>(A.n, A.row_ptr.ptr(), A.col_idx.ptr(), A.diag_idx.ptr(), va, x, b, d, y, g.omega);
            else block4_kernel<MatT, VecT, B_SPMV, false><<<grid, 256, 0, s>>>(A.n, A.row_ptr.ptr(), A.col_idx.ptr(), A.diag_idx.ptr(), va, x, b, d, y, g.omega);
            break;
        case EPI_RESID:
            if (A.has_ext_diag) block4_kernel<MatT, VecT, B_RESID, true><<<grid, 256, 0, s>>>(A.n, A.row_ptr.ptr(), A.col_idx.ptr(), A.diag_idx.ptr(), va, x, b, d, y, g.omega);
            else block4_kernel<MatT, VecT, B_RESID, false><<<grid, 256, 0, s>>>(A.n, A.row_ptr.ptr(), A.col_idx.ptr(), A.diag_idx.ptr(), va, x, b, d, y, g.omega);
            break;
        case EPI_JACOBI:
            block4_kernel<MatT, VecT, B_JACOBI, false><<<grid, 256, 0, s>>>(A.n, A.row_ptr.ptr(), A.col_idx.ptr(), A.diag_idx.ptr(), va, x, b, d, y, g.omega);
            break;
        default: fatal(AMGX_RC_INTERNAL, "block_apply: fused reductions are scalar-kernel features");
        }
    });
    count_launch();
    AMGXB_LAUNCH_CHECK();
}

void block_jacobi_setup(const Matrix &A, DevVec &dinv, cudaStream_t s)
{
    if (A.bx != 4) fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "Unsupported block size for BlockJacobi_Solver (1 and 4 are enabled)");
    if (A.n == 0) return;
    const int grid = std::min(ceil_div(A.n, 128), 4096);
    if (A.mat_prec == Prec::F64) block4_invert_diag<double><<<grid, 128, 0, s>>>(A.n, dinv.as<double>());
    else block4_invert_diag<float><<<grid, 128, 0, s>>>(A.n, dinv.as<float>());
    count_launch();
    AMGXB_LAUNCH_CHECK();
}

void block_jacobi_zero(const Matrix &A, const DevVec &dinv, const DevVec &b, void *x, double omega, cudaStream_t s)
{
    if (A.n == 0) return;
    const int grid = std::max(1, std::min(ceil_div(A.n, 64), B200_SMS * 8));
    AMGXB_DISPATCH(A.mat_prec, A.vec_prec, { block4_jacobi_zero<MatT, VecT><<<grid, 256, 0, s>>>(A.n, dinv.as<MatT>(), b.as<VecT>(), (VecT *)x, omega); });
    count_launch();
    AMGXB_LAUNCH_CHECK();
}

void block_jacobi_sweep(const Matrix &A, const DevVec &dinv, const DevVec &b, const void *x, void *xout, double omega, cudaStream_t s)
{
    CsrOpArgs g;
    g.x = x;
    g.b = b.ptr();
    g.d = dinv.ptr();
    g.y = xout;
    g.omega = omega;
    block_apply(A, EPI_JACOBI, g, s);
}

void block_norms(const DevVec &v, int n, int bsize, int norm_type, const ReduceCtx &red, ScalarBlock &sb, std::vector<double> &out, cudaStream_t s,
                 const Matrix *dist_of)
{
    if (bsize > 8) fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "block norms support block sizes up to 8");
    const int grid = std::max(1, std::min(ceil_div(n, 256), 256));
    static DevBuf<double> part;          // one process drives one device: a single cached partials buffer (no cudaMalloc per norm)
    static double *hpart = nullptr;      // pinned landing zone
    if (part.size() < (size_t)256 * 8) {
        part.resize((size_t)256 * 8);
        AMGXB_CUDA_CHECK(cudaHostAlloc(&hpart, sizeof(double) * 256 * 8, cudaHostAllocDefault));
    }
    AMGXB_DISPATCH_VEC(v.prec, {
        if (norm_type == 1) block_norm_kernel<VecT, 1><<<grid, 256, 0, s>>>(v.as<VecT>(), n, bsize, part.ptr());
        else if (norm_type == 0) block_norm_kernel<VecT, 0><<<grid, 256, 0, s>>>(v.as<VecT>(), n, bsize, part.ptr());
        else block_norm_kernel<VecT, 2><<<grid, 256, 0, s>>>(v.as<VecT>(), n, bsize, part.ptr());
    });
    count_launch();
    AMGXB_LAUNCH_CHECK();
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(hpart, part.ptr(), sizeof(double) * (size_t)grid * 8, cudaMemcpyDeviceToHost, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    const double *h = hpart;
    out.assign(bsize, 0.0);
    for (int b = 0; b < grid; b++)
        for (int c = 0; c < bsize; c++) {
            if (norm_type == 2) out[c] = std::max(out[c], h[(size_t)b * 8 + c]);
            else out[c] += h[(size_t)b * 8 + c];
        }
    if (dist_of && dist_of->dist) dist_allreduce_host(*dist_of, out.data(), bsize, norm_type == 2 ? 2 : 0);
    if (norm_type == 1) for (auto &o : out) o = std::sqrt(o);
    (void)red; (void)sb;
}

}  // namespace amgxb
