// misc.cu -- output callback, memory accounting, matrix_apply dispatcher, L1 row norms.
#include "solvers.h"
#include "dist.h"
#include <cstdarg>
#include <map>
#include <utility>

namespace amgxb {

static AMGX_print_callback g_print_cb = nullptr;
void set_print_callback(AMGX_print_callback cb) { g_print_cb = cb; }

void amgx_output(const char *msg, int len)
{
    if (g_print_cb) g_print_cb(msg, len);
    else { fwrite(msg, 1, (size_t)len, stdout); fflush(stdout); }
}

void amgx_printf(const char *fmt, ...)
{
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (n > 0) amgx_output(buf, std::min(n, (int)sizeof(buf) - 1));
}

void smem_opt_in(const void *kernel, size_t smem)
{
    static std::map<std::pair<const void *, int>, size_t> granted;      // (kernel, device) -> bytes; host threads do not share handles (as in the reference)
    int dev = 0;
    AMGXB_CUDA_CHECK(cudaGetDevice(&dev));
    size_t &have = granted[std::make_pair(kernel, dev)];
    if (smem > have) {
        AMGXB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        have = smem;
    }
}

double device_mem_used_gb()
{
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) return 0.0;
    return (double)(tot - fr) / (1024.0 * 1024.0 * 1024.0);
}

Resources::~Resources()
{
    release_reduce_scratch(this);
    dist_destroy_comm(this);
    if (stream) cudaStreamDestroy(stream);
    if (side_stream) cudaStreamDestroy(side_stream);
}

void Matrix::compute_diag_and_plan()
{
    if (bs() == 1) csr_build_plan(*this, stream());
    else block_build_diag(*this, stream());
    initialized = true;
}

namespace {
template <class MatT, class VecT> __global__ void l1_kernel(int n, const int *__restrict__ rp, const int *__restrict__ col, const MatT *__restrict__ val, MatT *__restrict__ d)
{
    // compute_d_kernel, src/solvers/jacobi_l1_solver.cu:60-91: accumulate |a_ij| in the VECTOR precision
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        VecT acc = 0;
        bool npd = false;
        for (int k = rp[i]; k < rp[i + 1]; k++) {
            VecT a = (VecT)val[k];
            if (col[k] == i && a < 0.) npd = true;
            a = fabs(a);
            acc += a;
        }
        d[i] = (MatT)(npd ? -acc : acc);
    }
}
}  // namespace

void l1_row_norms(const Matrix &A, DevVec &d, cudaStream_t s)
{
    d.resize((size_t)A.n, A.mat_prec);
    if (A.n == 0) return;
    const int grid = std::min(ceil_div(A.n, 256), B200_SMS * 16);
    AMGXB_DISPATCH(A.mat_prec, A.vec_prec, {
        l1_kernel<MatT, VecT><<<grid, 256, 0, s>>>(A.n, A.row_ptr.ptr(), A.col_idx.ptr(), A.values.as<MatT>(), d.as<MatT>());
    });
    count_launch();
    AMGXB_LAUNCH_CHECK();
}

}  // namespace amgxb
