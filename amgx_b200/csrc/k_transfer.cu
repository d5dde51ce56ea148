// k_transfer.cu -- aggregation-AMG level transfer operators.
// Replaces restrictResidualKernel / prolongateAndApplyCorrectionKernel and their block variants
// (src/aggregation/aggregation_amg_level.cu:91-181).  Summation order inside an aggregate is the
// order of R_column_indices (ascending fine row), exactly as in the reference.
#include "kernels.h"

namespace amgxb {
namespace {

template <class VecT> __global__ void restrict_kernel(const int *__restrict__ Rp, const int *__restrict__ Rc, const VecT *__restrict__ r,
                                                      VecT *__restrict__ rc, int n_agg)
{
    for (int I = blockIdx.x * blockDim.x + threadIdx.x; I < n_agg; I += gridDim.x * blockDim.x) {
        VecT t = 0;
        const int j1 = Rp[I + 1];
        for (int j = Rp[I]; j < j1; j++) t = t + __ldg(r + Rc[j]);
        rc[I] = t;
    }
}

template <class VecT> __global__ void restrict_block_kernel(const int *__restrict__ Rp, const int *__restrict__ Rc, const VecT *__restrict__ r,
                                                            VecT *__restrict__ rc, int n_agg, int bsize)
{
    const long long total = (long long)n_agg * bsize;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int I = (int)(t / bsize), m = (int)(t % bsize);
        VecT acc = 0;
        const int j1 = Rp[I + 1];
        for (int j = Rp[I]; j < j1; j++) acc = acc + __ldg(r + (size_t)Rc[j] * bsize + m);
        rc[t] = acc;
    }
}

template <class VecT> __global__ void prolong_kernel(const int *__restrict__ agg, const VecT *__restrict__ e, const VecT *x, VecT *xout, int n, int bsize)
{
    const long long total = (long long)n * bsize;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / bsize), m = (int)(t % bsize);
        xout[t] = x[t] + __ldg(e + (size_t)agg[i] * bsize + m);   // alpha == 1
    }
}

template <class VecT> __global__ void prolong_set_kernel(const int *__restrict__ agg, const VecT *__restrict__ e, VecT *__restrict__ x, int n, int bsize)
{
    const long long total = (long long)n * bsize;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / bsize), m = (int)(t % bsize);
        x[t] = __ldg(e + (size_t)agg[i] * bsize + m);   // 0 + e == e: identical to zero-fill followed by prolongateAndApplyCorrection
    }
}

}  // namespace

void agg_prolong_set(const int *aggregates, const void *e, void *x, Prec p, int n, int bsize, cudaStream_t s)
{
    if (n == 0) return;
    AMGXB_DISPATCH_VEC(p, {
        int grid = std::min(ceil_div((long long)n * bsize, 256), B200_SMS * 16);
        prolong_set_kernel<VecT><<<grid, 256, 0, s>>>(aggregates, (const VecT *)e, (VecT *)x, n, bsize);
    });
    count_launch();
    AMGXB_LAUNCH_CHECK();
}

void agg_restrict(const int *Rp, const int *Rc, const void *r, void *rc, Prec p, int n_agg, int bsize, cudaStream_t s)
{
    if (n_agg == 0) return;
    AMGXB_DISPATCH_VEC(p, {
        if (bsize == 1) {
            int grid = std::min(ceil_div(n_agg, 256), B200_SMS * 16);
            restrict_kernel<VecT><<<grid, 256, 0, s>>>(Rp, Rc, (const VecT *)r, (VecT *)rc, n_agg);
        } else {
            int grid = std::min(ceil_div((long long)n_agg * bsize, 256), B200_SMS * 16);
            restrict_block_kernel<VecT><<<grid, 256, 0, s>>>(Rp, Rc, (const VecT *)r, (VecT *)rc, n_agg, bsize);
        }
    });
    count_launch();
    AMGXB_LAUNCH_CHECK();
}

void agg_prolong_add(const int *aggregates, const void *e, const void *x, void *xout, Prec p, int n, int bsize, cudaStream_t s)
{
    if (n == 0) return;
    AMGXB_DISPATCH_VEC(p, {
        int grid = std::min(ceil_div((long long)n * bsize, 256), B200_SMS * 16);
        prolong_kernel<VecT><<<grid, 256, 0, s>>>(aggregates, (const VecT *)e, (const VecT *)x, (VecT *)xout, n, bsize);
    });
    count_launch();
    AMGXB_LAUNCH_CHECK();
}

}  // namespace amgxb
