// coloring.cu -- MIN_MAX graph colouring (one ring) and the per-colour row lists the multicolour smoothers use.
//   colorRowsKernel + host loop      src/matrix_coloring/min_max.cu:103-160, 380-420
//   createColorArrays                src/matrix_coloring/matrix_coloring.cu:230-330 (stable sort of rows by colour)
// Also AMGX_matrix_attach_coloring (user supplied colours, include/amgx_c.h:512-516).
#include "solvers.h"
#include "dist.h"
#include "capi_internal.h"
#include <cub/cub.cuh>

namespace amgxb {
namespace {

__host__ __device__ inline unsigned hash_function(unsigned a, unsigned seed)   // min_max.cu:27-37
{
    a ^= seed;
    a = (a + 0x7ed55d16u) + (a << 12);
    a = (a ^ 0xc761c23cu) + (a >> 19);
    a = (a + 0x165667b1u) + (a << 5);
    a = (a ^ 0xd3a2646cu) + (a << 9);
    a = (a + 0xfd7046c5u) + (a << 3);
    a = (a ^ 0xb55a4f09u) + (a >> 16);
    return a;
}

// One pass: local maxima of the (signed) hash among uncolored neighbours get `current_color`, local minima
// `current_color + 1`.  Reads of row_colors race with writes of the same pass exactly as in the reference; the
// (0 | current | current+1) tests make the outcome independent of the interleaving.
__global__ void color_rows_kernel(const int *__restrict__ rp, const int *__restrict__ ci, int *row_colors, int current_color, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (row_colors[i] != 0) continue;
        const int hash_i = (int)hash_function((unsigned)i, 0);
        bool max_i = true, min_i = true;
        for (int r = rp[i]; r < rp[i + 1]; r++) {
            const int j = ci[r];
            if (j >= n) continue;
            const int hash_j = (int)hash_function((unsigned)j, 0);
            const int cj = ((volatile int *)row_colors)[j];
            if (hash_j > hash_i && (cj == 0 || cj == current_color)) max_i = false;
            if (hash_j < hash_i && (cj == 0 || cj == current_color + 1)) min_i = false;
        }
        int c = 0;
        if (max_i) c = current_color;
        else if (min_i) c = current_color + 1;
        if (c != 0) row_colors[i] = c;
    }
}

__global__ void count_zero_kernel(int n, const int *__restrict__ v, int *count, int *maxv)
{
    int c = 0, m = 0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) { c += (v[t] == 0); m = max(m, v[t]); }
    for (int o = 16; o > 0; o >>= 1) { c += __shfl_xor_sync(0xffffffffu, c, o); m = max(m, __shfl_xor_sync(0xffffffffu, m, o)); }
    if ((threadIdx.x & 31) == 0) { if (c) atomicAdd(count, c); atomicMax(maxv, m); }
}
__global__ void iota_k(int n, int *v) { for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) v[t] = t; }
__global__ void offsets_k(int n, const int *__restrict__ keys, int n_keys, int *__restrict__ offsets)
{
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p <= n; p += gridDim.x * blockDim.x) {
        const int prev = (p == 0) ? -1 : keys[p - 1];
        const int cur = (p == n) ? n_keys : keys[p];
        for (int I = prev + 1; I <= cur; I++) offsets[I] = p;
    }
}

inline int grid_for(long long n) { return std::max(1, std::min(ceil_div(n, 256), 148 * 16)); }

void build_color_arrays(Matrix &A, cudaStream_t s)
{
    const int n = A.n;
    A.sorted_rows_by_color.resize(n);
    DevBuf<int> keys_out, vals_in, offs;
    keys_out.resize(n);
    vals_in.resize(n);
    iota_k<<<grid_for(n), 256, 0, s>>>(n, vals_in.ptr());
    int bits = 1;
    while ((1ll << bits) < (long long)A.num_colors + 1) bits++;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, A.row_colors.ptr(), keys_out.ptr(), vals_in.ptr(), A.sorted_rows_by_color.ptr(), n, 0, bits, s);
    DevBytes tmp;
    tmp.resize(tb);
    cub::DeviceRadixSort::SortPairs(tmp.p, tb, A.row_colors.ptr(), keys_out.ptr(), vals_in.ptr(), A.sorted_rows_by_color.ptr(), n, 0, bits, s);
    offs.resize(A.num_colors + 1);
    offsets_k<<<grid_for(n + 1), 256, 0, s>>>(n, keys_out.ptr(), A.num_colors, offs.ptr());
    count_launch(2);
    AMGXB_LAUNCH_CHECK();
    A.color_offsets = offs.to_host(s);
}

}  // namespace

// colorMatrixOneRing: loop until the uncoloured fraction is small enough (0 when determinism_flag is set)
void color_matrix_min_max(Matrix &A, double max_uncolored_fraction, cudaStream_t s)
{
    if (A.user_coloring) return;
    const int n = A.n;
    A.row_colors.resize(std::max(A.n_cols, n));
    A.row_colors.zero(s);
    const int max_uncolored = (int)(max_uncolored_fraction * (double)n);
    DevBuf<int> cnt;
    cnt.resize(2);
    int num_colors = 1;
    for (int num_uncolored = n; num_uncolored > max_uncolored;) {
        color_rows_kernel<<<grid_for(n), 256, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), A.row_colors.ptr(), num_colors, n);
        num_colors += 2;
        cnt.zero(s);
        count_zero_kernel<<<grid_for(n), 256, 0, s>>>(n, A.row_colors.ptr(), cnt.ptr(), cnt.ptr() + 1);
        count_launch(2);
        AMGXB_LAUNCH_CHECK();
        num_uncolored = cnt.to_host(s)[0];
    }
    cnt.zero(s);
    count_zero_kernel<<<grid_for(n), 256, 0, s>>>(n, A.row_colors.ptr(), cnt.ptr(), cnt.ptr() + 1);
    count_launch();
    A.num_colors = cnt.to_host(s)[1] + 1;
    build_color_arrays(A, s);
}

void attach_user_coloring(Matrix &A, const int *row_coloring, int num_rows, int num_colors)
{
    if (num_rows != A.n || num_colors < 1 || !row_coloring) fatal(AMGX_RC_BAD_PARAMETERS, "attach_coloring: sizes do not match the matrix");
    cudaStream_t s = A.stream();
    A.row_colors.resize(std::max(A.n_cols, A.n));
    A.row_colors.zero(s);
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(A.row_colors.ptr(), row_coloring, sizeof(int) * num_rows, cudaMemcpyDefault, s));
    A.num_colors = num_colors;
    A.user_coloring = true;
    build_color_arrays(A, s);
}

}  // namespace amgxb
