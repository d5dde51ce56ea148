// tile_common.cuh -- device helpers shared by the CSR tile kernels (k_spmv.cu, k_spmv_enc.cu): mbarrier / TMA bulk-copy wrappers,
// the deterministic "last CTA finalises" reduction with its scalar epilogues, the kernel argument block.  Included INSIDE
// `namespace amgxb { namespace { ... } }` of each translation unit (internal linkage, as when this code lived in k_spmv.cu).
#pragma once

constexpr int PRODUCER_THREADS = 32;
constexpr int MAX_STAGES = 4;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
// non-blocking probe of a phase (polling loops that serve several barriers)
__device__ __forceinline__ bool mbar_test(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <class T> __device__ __forceinline__ T guard_diag(T d);
template <> __device__ __forceinline__ double guard_diag<double>(double d) { return fabs(d) < 1e-12 ? copysign(1e-12, d) : d; }
template <> __device__ __forceinline__ float guard_diag<float>(float d) { return fabs((double)d) < 1e-7 ? copysignf((float)1e-7, d) : d; }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// scalar epilogue executed by one thread of the last CTA
__device__ void apply_fin(double sum, double *scal, int fin_op, int slot, double *host_mirror, int mirror)
{
    double out = sum;
    switch (fin_op) {
    case FIN_STORE:
    case FIN_ABS: scal[slot] = sum; break;
    case FIN_SQRT: out = sqrt(sum); scal[slot] = out; break;
    case FIN_ADD: out = scal[slot] + sum; scal[slot] = out; break;
    case FIN_PCG_ALPHA: {
        scal[S_DOT] = sum;
        double a = (sum != 0.0) ? scal[S_RZ] / sum : 0.0;
        scal[S_ALPHA] = a;
        scal[S_NEG_ALPHA] = -a;
        out = a;
        break;
    }
    case FIN_PCG_BETA: {
        double old = scal[S_RZ];
        scal[S_RZ_OLD] = old;
        scal[S_RZ] = sum;
        double bta = (old != 0.0) ? sum / old : 0.0;
        scal[S_BETA] = bta;
        out = sum;
        break;
    }
    }
    if (mirror && host_mirror) { host_mirror[slot] = out; }
}

// Block-level deterministic reduction + "last block finalises" pattern.
// All threads of the CTA must call it; `nthreads` = blockDim.x; smem_red has >= 33 doubles.
__device__ void block_reduce_finish(double v, double *smem_red, const ReduceCtx &red, int fin_op, int slot, int mirror)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    if (lane == 0) smem_red[warp] = v;
    __syncthreads();
    __shared__ bool is_last;
    if (warp == 0) {
        double t = (lane < nwarps) ? smem_red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) {
            red.partials[blockIdx.x] = t;
            __threadfence();
            unsigned ticket = atomicAdd(red.counter, 1u);
            is_last = (ticket == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        // fixed order: thread t sums partials t, t+blockDim, ... then a fixed tree
        double t = 0.0;
        for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) t += ((volatile double *)red.partials)[i];
        t = warp_sum(t);
        if (lane == 0) smem_red[warp] = t;
        __syncthreads();
        if (warp == 0) {
            double u = (lane < nwarps) ? smem_red[lane] : 0.0;
            u = warp_sum(u);
            if (lane == 0) {
                apply_fin(u, red.scal, fin_op, slot, red.host_mirror, mirror);
                *red.counter = 0u;
                __threadfence_system();
            }
        }
    }
}

// Pull `count` elements at `ptr` towards L2 ahead of their consumers (UBLKPF.L2): the producer warp runs `stages` tiles ahead, so the
// per-row vector loads and the furthest-ahead gathers of a tile find their lines in L2 instead of paying a DRAM round trip on the
// consumers' critical path (one row per thread = one dependent chain per tile).  Rounded inwards to 16 bytes: never reads outside.
template <class T> __device__ __forceinline__ void l2_prefetch_span(const T *ptr, int count)
{
    if (count <= 0) return;
    const unsigned long long a0 = (reinterpret_cast<unsigned long long>(ptr) + 15ull) & ~15ull;
    const unsigned long long a1 = (reinterpret_cast<unsigned long long>(ptr + count)) & ~15ull;
    if (a1 > a0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"((unsigned)(a1 - a0)) : "memory");
}

// AMGXB_L2_PREFETCH: bit 0 = the plain tile kernel (b / d slices), bit 1 = the coded kernel (b / d and the x rows ahead).  r02 A/B at 256^3:
// plain fused Jacobi 0.297 -> 0.289 ms (solve 260 -> 265 it/s), coded kernels 0.159 -> 0.170 ms (their consumers are bound by the SM's
// load/store path, not by DRAM latency, and the extra requests only add to it): default 1.
inline int l2_prefetch_flags()
{
    static const int on = getenv("AMGXB_L2_PREFETCH") ? atoi(getenv("AMGXB_L2_PREFETCH")) : 1;
    return on;
}

template <class MatT, class VecT> struct TileArgs {
    const int *row_ptr;
    const int *col;
    const MatT *val;
    int n, num_tiles, cap, stages;   // n = end row of the segment
    int row0;                        // first row of the segment (multiple of 4)
    int unroll;                      // gathers in flight per consumer step: 4 or 8
    const unsigned char *perm;       // length-sorted thread -> row map of every tile (null: thread t takes row t), see tile_perm_kernel
    int tile_base;                   // index of the segment's first tile in `perm` (tiles are numbered over the row segments)
    int l2pf;                        // producer prefetches the tile's b / d slices (and, coded stencil tiles, the x rows its furthest neighbour reads) into L2
    const VecT *x;
    const int *agg;
    const VecT *b;
    const MatT *d;
    VecT *y;
    double omega;
    ReduceCtx red;
    int fin_op, fin_slot, mirror;
};

template <class VecT, bool AGG> __device__ __forceinline__ VecT gather(const VecT *x, const int *agg, int c)
{
    if (AGG) return __ldg(x + __ldg(agg + c));
    return __ldg(x + c);
}

// what a consumer does with its row's dot product (identical to csr_tile_kernel); returns the row's contribution to the fused reduction
template <class MatT, class VecT, int EPI>
__device__ __forceinline__ double tile_epilogue(const TileArgs<MatT, VecT> &a, const int row, const VecT sum, const VecT bi, const MatT di, const VecT xi)
{
    if (EPI == EPI_SPMV) {
        a.y[row] = sum;
        return 0.0;
    } else if (EPI == EPI_SPMV_DOT) {
        a.y[row] = sum;
        return (double)sum * (double)xi;
    } else if (EPI == EPI_RESID) {
        a.y[row] = bi - sum;
        return 0.0;
    } else if (EPI == EPI_ADD) {
        a.y[row] = bi + sum;
        return 0.0;
    } else if (EPI == EPI_RESID_NRM2) {
        const VecT r = bi - sum;
        a.y[row] = r;
        return (double)r * (double)r;
    } else {
        // x + ((b - Ax) * w) * (1/d): d = 1/d; b -= y; b *= w; b*d + x  (one FMA)
        MatT dinv = (MatT)1 / guard_diag<MatT>(di);
        VecT t = bi - sum;
        t = (VecT)(t * a.omega);
        const VecT out = fma(t, (VecT)dinv, xi);
        a.y[row] = out;
        return (EPI == EPI_JACOBI_DOT) ? (double)bi * (double)out : 0.0;
    }
}
