// k_blas.cu -- level-1 kernels of the Krylov loop (PCG / FGMRES) and the Jacobi zero-guess sweep.
// Replaces the cuBLAS / Thrust calls of the reference's blas layer (src/blas.cu:148-321, 501-917,
// src/amgx_cublas.cu:438-500, src/norm.cu:34-90).  Differences by design:
//   * scalars (alpha, beta, norms) stay in device memory (ReduceCtx::scal) -- no host sync per dot;
//   * every reduction is ONE kernel: warp shuffle -> CTA partial -> last CTA sums the partials in a
//     fixed order (bit-reproducible run to run) and applies the scalar epilogue;
//   * PCG's "x += a p; r -= a Ap; ||r||" is one pass instead of three.
#include "kernels.h"

namespace amgxb {
namespace {

constexpr int BLK = 256;
constexpr int UNROLL = 4;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ void fin_apply(double sum, double *scal, int fin_op, int slot, double *host_mirror, int mirror)
{
    double out = sum;
    switch (fin_op) {
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
        scal[S_BETA] = (old != 0.0) ? sum / old : 0.0;
        break;
    }
    default: scal[slot] = sum; break;
    }
    if (mirror && host_mirror) host_mirror[slot] = out;
}

// IS_MAX: reduction is max instead of sum
template <bool IS_MAX> __device__ void block_finish(double v, const ReduceCtx &red, int fin_op, int slot, int mirror)
{
    __shared__ double sred[32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    v = IS_MAX ? warp_max(v) : warp_sum(v);
    if (lane == 0) sred[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double t = (lane < nwarps) ? sred[lane] : 0.0;
        t = IS_MAX ? warp_max(t) : warp_sum(t);
        if (lane == 0) {
            red.partials[blockIdx.x] = t;
            __threadfence();
            is_last = (atomicAdd(red.counter, 1u) == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double t = 0.0;
        for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
            double p = ((volatile double *)red.partials)[i];
            t = IS_MAX ? fmax(t, p) : t + p;
        }
        t = IS_MAX ? warp_max(t) : warp_sum(t);
        if (lane == 0) sred[warp] = t;
        __syncthreads();
        if (warp == 0) {
            double u = (lane < nwarps) ? sred[lane] : 0.0;
            u = IS_MAX ? warp_max(u) : warp_sum(u);
            if (lane == 0) {
                fin_apply(u, red.scal, fin_op, slot, red.host_mirror, mirror);
                *red.counter = 0u;
                __threadfence_system();
            }
        }
    }
}

template <class T, class F> __global__ void __launch_bounds__(BLK) map_kernel(size_t n, F f)
{
    const size_t stride = (size_t)gridDim.x * BLK;
    for (size_t i = (size_t)blockIdx.x * BLK + threadIdx.x; i < n; i += stride) f(i);
}

template <class F> void launch_map(size_t n, F f, cudaStream_t s)
{
    if (n == 0) return;
    int grid = (int)std::min<size_t>((n + BLK - 1) / BLK, (size_t)blas_max_grid());
    map_kernel<double, F><<<grid, BLK, 0, s>>>(n, f);
    count_launch();
    AMGXB_LAUNCH_CHECK();
}

// generic reduction: v = f(i) summed (or maxed)
template <bool IS_MAX, class F> __global__ void __launch_bounds__(BLK) reduce_kernel(size_t n, F f, ReduceCtx red, int fin_op, int slot, int mirror)
{
    const size_t stride = (size_t)gridDim.x * BLK;
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * BLK + threadIdx.x; i < n; i += stride) {
        double v = f(i);
        acc = IS_MAX ? fmax(acc, v) : acc + v;
    }
    block_finish<IS_MAX>(acc, red, fin_op, slot, mirror);
}

template <bool IS_MAX, class F> void launch_reduce(size_t n, F f, const ReduceCtx &red, int fin_op, int slot, int mirror, cudaStream_t s)
{
    int grid = (int)std::max<size_t>(1, std::min<size_t>((n + (size_t)BLK * UNROLL - 1) / ((size_t)BLK * UNROLL), (size_t)blas_max_grid()));
    reduce_kernel<IS_MAX, F><<<grid, BLK, 0, s>>>(n, f, red, fin_op, slot, mirror);
    count_launch();
    AMGXB_LAUNCH_CHECK();
}

}  // namespace

int blas_max_grid() { return B200_SMS * 8; }

void vec_fill(void *x, Prec p, size_t n, double v, cudaStream_t s)
{
    AMGXB_DISPATCH_VEC(p, { VecT *X = (VecT *)x; VecT val = (VecT)v; launch_map(n, [=] __device__(size_t i) { X[i] = val; }, s); });
}

void vec_copy(void *dst, const void *src, Prec p, size_t n, cudaStream_t s)
{
    if (n == 0 || dst == src) return;
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(dst, src, n * prec_size(p), cudaMemcpyDeviceToDevice, s));
}

// out = x*a + y*b   (AXPBY functor, src/blas.cu:88-104)
void vec_axpby(const void *x, const void *y, void *out, Prec p, size_t n, double a, double b, cudaStream_t s)
{
    AMGXB_DISPATCH_VEC(p, {
        const VecT *X = (const VecT *)x; const VecT *Y = (const VecT *)y; VecT *O = (VecT *)out;
        VecT aa = (VecT)a, bb = (VecT)b;
        launch_map(n, [=] __device__(size_t i) { O[i] = X[i] * aa + Y[i] * bb; }, s);
    });
}

// y = a*x + y   (cublas?axpy)
void vec_axpy(const void *x, void *y, Prec p, size_t n, double a, cudaStream_t s)
{
    AMGXB_DISPATCH_VEC(p, {
        const VecT *X = (const VecT *)x; VecT *Y = (VecT *)y; VecT aa = (VecT)a;
        launch_map(n, [=] __device__(size_t i) { Y[i] = fma(aa, X[i], Y[i]); }, s);
    });
}

void vec_scal(void *x, Prec p, size_t n, double a, cudaStream_t s)
{
    AMGXB_DISPATCH_VEC(p, { VecT *X = (VecT *)x; VecT aa = (VecT)a; launch_map(n, [=] __device__(size_t i) { X[i] = X[i] * aa; }, s); });
}

// out = x*a + y*b + z*c, the expression of the reference's AXPBYPCZ functor (src/blas.cu:107-124)
void vec_axpbypcz(const void *x, const void *y, const void *z, void *out, Prec p, size_t n, double a, double b, double c, cudaStream_t s)
{
    AMGXB_DISPATCH_VEC(p, {
        const VecT *X = (const VecT *)x; const VecT *Y = (const VecT *)y; const VecT *Z = (const VecT *)z; VecT *O = (VecT *)out;
        const VecT aa = (VecT)a, bb = (VecT)b, cc = (VecT)c;
        launch_map(n, [=] __device__(size_t i) { O[i] = X[i] * aa + Y[i] * bb + Z[i] * cc; }, s);
    });
}

// ---- scaling = DIAGONAL_SYMMETRIC (src/scalers/diagonal_symmetric.cu): s_i = 1/sqrt(a_ii); A <- S A S in place (values[jj] *= s_i*s_j),
// undone by the division; vectors are multiplied / divided entry-wise.  negative[0] is set when a diagonal entry is negative.
namespace {
template <class MatT, class VecT> __global__ void diag_inv_sqrt_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const MatT *__restrict__ va,
                                                                        VecT *s, int *negative)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        VecT d = 0;
        for (int k = rp[i]; k < rp[i + 1]; k++) if (ci[k] == i) d = (VecT)va[k];      // grabDiagonalVector: the last match wins
        if (d < (VecT)0) negative[0] = 1;
        s[i] = 1. / sqrt(d);
    }
}
template <class MatT, class VecT> __global__ void scale_matrix_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, MatT *va, const VecT *__restrict__ s,
                                                                       int unscale)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const VecT si = s[i];
        for (int k = rp[i]; k < rp[i + 1]; k++) {
            const VecT sj = s[ci[k]];
            if (unscale) va[k] /= si * sj;
            else va[k] *= si * sj;
        }
    }
}
}  // namespace

int diag_sym_scale_setup(const Matrix &A, DevVec &scale, cudaStream_t s)
{
    scale.resize((size_t)A.n, A.vec_prec);
    DevBuf<int> neg;
    neg.resize(1);
    neg.zero(s);
    if (A.n) {
        const int grid = std::min(ceil_div(A.n, 256), blas_max_grid());
        AMGXB_DISPATCH(A.mat_prec, A.vec_prec, {
            diag_inv_sqrt_kernel<MatT, VecT><<<grid, 256, 0, s>>>(A.n, A.row_ptr.ptr(), A.col_idx.ptr(), A.values.as<MatT>(), scale.as<VecT>(), neg.ptr());
        });
        count_launch();
        AMGXB_LAUNCH_CHECK();
    }
    return neg.to_host(s)[0];
}

void diag_sym_scale_matrix(Matrix &A, const DevVec &scale, bool unscale, cudaStream_t s)
{
    if (!A.n) return;
    const int grid = std::min(ceil_div(A.n, 256), blas_max_grid());
    AMGXB_DISPATCH(A.mat_prec, A.vec_prec, {
        scale_matrix_kernel<MatT, VecT><<<grid, 256, 0, s>>>(A.n, A.row_ptr.ptr(), A.col_idx.ptr(), A.values.as<MatT>(), scale.as<VecT>(), unscale ? 1 : 0);
    });
    count_launch();
    AMGXB_LAUNCH_CHECK();
    csr_values_changed(A, s);
}

void vec_scale_entrywise(void *v, const void *d, Prec p, size_t n, bool divide, cudaStream_t s)
{
    AMGXB_DISPATCH_VEC(p, {
        VecT *V = (VecT *)v; const VecT *D = (const VecT *)d;
        if (divide) launch_map(n, [=] __device__(size_t i) { V[i] /= D[i]; }, s);
        else launch_map(n, [=] __device__(size_t i) { V[i] *= D[i]; }, s);
    });
}

// scale of the coarse-grid correction, error_scaling = 2, 3: out[0] = clamp(scal[nom] / scal[den]) exactly as
// aggregation_amg_level.cu:797-817 words it (|den| == 0 -> 1; |alpha| < .3 -> sign * .3; |alpha| > 10 -> sign * 10)
void scalar_error_scale(const double *scal, int slot_nom, int slot_den, double *out, cudaStream_t s)
{
    launch_map(1, [=] __device__(size_t) {
        double nom = scal[slot_nom], den = scal[slot_den];
        if (fabs(den) == 0.0) nom = den = 1.0;
        double alpha = nom / den;
        if (fabs(alpha) < .3) alpha = (alpha / fabs(alpha)) * .3;
        if (fabs(alpha) > 10) alpha = (alpha / fabs(alpha)) * 10.;
        out[0] = alpha;
    }, s);
}

void vec_axpy_dev(const void *x, void *y, Prec p, size_t n, const double *scal, int slot, double sign, cudaStream_t s)
{
    AMGXB_DISPATCH_VEC(p, {
        const VecT *X = (const VecT *)x; VecT *Y = (VecT *)y;
        launch_map(n, [=] __device__(size_t i) { VecT aa = (VecT)(sign * scal[slot]); Y[i] = fma(aa, X[i], Y[i]); }, s);
    });
}

// y += sign*scal[slot]*x and, in the same pass, <z, y_new> (z == nullptr: <y_new, y_new>).  Element arithmetic and the
// reduction tree are those of vec_axpy_dev followed by vec_dot, so the fused chain reproduces the unfused one bit for bit.
void vec_axpy_dot_dev(const void *x, void *y, const void *z, Prec p, size_t n, const double *scal, int slot, double sign, const ReduceCtx &red, int fin_op,
                      int fin_slot, int mirror, cudaStream_t s)
{
    AMGXB_DISPATCH_VEC(p, {
        const VecT *X = (const VecT *)x; VecT *Y = (VecT *)y; const VecT *Z = (const VecT *)z;
        launch_reduce<false>(n, [=] __device__(size_t i) {
            const VecT aa = (VecT)(sign * scal[slot]);
            const VecT yn = fma(aa, X[i], Y[i]);
            Y[i] = yn;
            const VecT zz = Z ? Z[i] : yn;
            return (double)zz * (double)yn;
        }, red, fin_op, fin_slot, mirror, s);
    });
}

void vec_axpby_dev(const void *x, const void *y, void *out, Prec p, size_t n, double a, const double *scal, int slot_b, cudaStream_t s)
{
    AMGXB_DISPATCH_VEC(p, {
        const VecT *X = (const VecT *)x; const VecT *Y = (const VecT *)y; VecT *O = (VecT *)out; VecT aa = (VecT)a;
        launch_map(n, [=] __device__(size_t i) { VecT bb = (VecT)scal[slot_b]; O[i] = X[i] * aa + Y[i] * bb; }, s);
    });
}

void vec_scal_dev_inv(void *x, Prec p, size_t n, const double *scal, int slot, cudaStream_t s)
{
    AMGXB_DISPATCH_VEC(p, {
        VecT *X = (VecT *)x;
        launch_map(n, [=] __device__(size_t i) { VecT aa = (VecT)(1.0 / scal[slot]); X[i] = X[i] * aa; }, s);
    });
}

void vec_dot(const void *x, const void *y, Prec p, size_t n, const ReduceCtx &red, int fin_op, int fin_slot, int mirror, cudaStream_t s)
{
    AMGXB_DISPATCH_VEC(p, {
        const VecT *X = (const VecT *)x; const VecT *Y = (const VecT *)y;
        launch_reduce<false>(n, [=] __device__(size_t i) { return (double)X[i] * (double)Y[i]; }, red, fin_op, fin_slot, mirror, s);
    });
}

void vec_nrm1(const void *x, Prec p, size_t n, const ReduceCtx &red, int fin_slot, int mirror, cudaStream_t s)
{
    AMGXB_DISPATCH_VEC(p, {
        const VecT *X = (const VecT *)x;
        launch_reduce<false>(n, [=] __device__(size_t i) { return fabs((double)X[i]); }, red, FIN_STORE, fin_slot, mirror, s);
    });
}

void vec_nrmmax(const void *x, Prec p, size_t n, const ReduceCtx &red, int fin_slot, int mirror, cudaStream_t s)
{
    AMGXB_DISPATCH_VEC(p, {
        const VecT *X = (const VecT *)x;
        launch_reduce<true>(n, [=] __device__(size_t i) { return fabs((double)X[i]); }, red, FIN_STORE, fin_slot, mirror, s);
    });
}

// norm_type: 0 = L1, 1 = L2, 2 = LMAX
void pcg_update_xr(const void *p, const void *Ap, void *x, void *r, Prec pr, size_t n, const ReduceCtx &red, int norm_type,
                   int fin_slot, int mirror, cudaStream_t s, bool partial)
{
    const double *scal = red.scal;
    AMGXB_DISPATCH_VEC(pr, {
        const VecT *P = (const VecT *)p; const VecT *AP = (const VecT *)Ap; VecT *X = (VecT *)x; VecT *R = (VecT *)r;
        if (norm_type == 1) {
            launch_reduce<false>(n, [=] __device__(size_t i) {
                const VecT a = (VecT)scal[S_ALPHA], na = (VecT)scal[S_NEG_ALPHA];
                X[i] = fma(a, P[i], X[i]);
                const VecT rn = fma(na, AP[i], R[i]);
                R[i] = rn;
                return (double)rn * (double)rn; }, red, partial ? FIN_STORE : FIN_SQRT, fin_slot, mirror, s);
        } else if (norm_type == 0) {
            launch_reduce<false>(n, [=] __device__(size_t i) {
                const VecT a = (VecT)scal[S_ALPHA], na = (VecT)scal[S_NEG_ALPHA];
                X[i] = fma(a, P[i], X[i]);
                const VecT rn = fma(na, AP[i], R[i]);
                R[i] = rn;
                return fabs((double)rn); }, red, FIN_STORE, fin_slot, mirror, s);
        } else {
            launch_reduce<true>(n, [=] __device__(size_t i) {
                const VecT a = (VecT)scal[S_ALPHA], na = (VecT)scal[S_NEG_ALPHA];
                X[i] = fma(a, P[i], X[i]);
                const VecT rn = fma(na, AP[i], R[i]);
                R[i] = rn;
                return fabs((double)rn); }, red, FIN_STORE, fin_slot, mirror, s);
        }
    });
}

namespace {
template <class T> __device__ __forceinline__ T guard_d(T d);
template <> __device__ __forceinline__ double guard_d<double>(double d) { return fabs(d) < 1e-12 ? copysign(1e-12, d) : d; }
template <> __device__ __forceinline__ float guard_d<float>(float d) { return fabs((double)d) < 1e-7 ? copysignf((float)1e-7, d) : d; }
}  // namespace

// x = b * w / d    (jacobi_presmooth_functor, src/solvers/block_jacobi_solver.cu:24-30)
void jacobi_zero_guess(const void *b, const void *d, void *x, Prec matp, Prec vecp, size_t n, double omega, cudaStream_t s)
{
    AMGXB_DISPATCH(matp, vecp, {
        const VecT *B = (const VecT *)b; const MatT *D = (const MatT *)d; VecT *X = (VecT *)x;
        launch_map(n, [=] __device__(size_t i) { X[i] = (VecT)(B[i] * omega / guard_d<MatT>(D[i])); }, s);
    });
}

}  // namespace amgxb
