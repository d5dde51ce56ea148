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

// PARALLEL_GREEDY, coloring_level 1 (src/matrix_coloring/parallel_greedy.cu:148-215): an uncoloured row whose (signed) hash beats
// every uncoloured neighbour takes the smallest colour >= 1 none of its neighbours holds.  The reference updates the colours in
// place, so whether a row sees a neighbour coloured earlier in the SAME launch depends on scheduling and its colouring is not
// reproducible run to run; this is the synchronous form of the same rule (every row reads the colours of the previous launch):
// always a proper colouring, reproducible, and what the reference produces whenever no such same-launch propagation happens.
__global__ void pg_color_kernel(const int *__restrict__ rp, const int *__restrict__ ci, const int *__restrict__ colors_in, int *colors_out, int n, int *max_color)
{
    int mc = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int c = colors_in[i];
        if (c == 0) {
            const int hash_i = (int)hash_function((unsigned)i, 0);
            unsigned long long used = 0ull;
            bool max_row = true;
            for (int r = rp[i]; r < rp[i + 1]; r++) {
                const int j = ci[r];
                if (j >= n || j == i) continue;
                const int cj = colors_in[j];
                if (cj > 0 && cj <= 64) used |= 1ull << (64 - cj);
                max_row &= (hash_i > (int)hash_function((unsigned)j, 0) || cj != 0);
            }
            if (max_row) {
                const unsigned long long free_mask = ~used;
                if (free_mask != 0ull) c = 64 - (63 - __clzll((long long)free_mask));      // 64 - bfind(~used)
            }
        }
        colors_out[i] = c;
        mc = max(mc, c);
    }
    for (int o = 16; o > 0; o >>= 1) mc = max(mc, __shfl_xor_sync(0xffffffffu, mc, o));
    if ((threadIdx.x & 31) == 0 && mc) atomicMax(max_color, mc);
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

inline int grid_for(long long n) { return std::max(1, std::min(ceil_div(n, 256), B200_SMS * 16)); }

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

// Parallel_Greedy_Matrix_Coloring::colorMatrix (parallel_greedy.cu:665-790), coloring_level 1: launches until the number of
// uncoloured rows counted BEFORE a launch is <= max_uncolored (that launch still colours), stops progressing, or 64 colours are in use
void color_matrix_parallel_greedy(Matrix &A, double max_uncolored_fraction, cudaStream_t s)
{
    if (A.user_coloring) return;
    const int n = A.n;
    A.row_colors.resize(std::max(A.n_cols, n));
    A.row_colors.zero(s);
    DevBuf<int> next, cnt;
    next.resize(std::max(A.n_cols, n));
    next.zero(s);
    cnt.resize(2);
    const int max_uncolored = (int)(max_uncolored_fraction * (double)n);
    int prev_uncolored = 0, max_color = 0;
    for (int iter = 0; n > 0; iter++) {
        cnt.zero(s);
        count_zero_kernel<<<grid_for(n), 256, 0, s>>>(n, A.row_colors.ptr(), cnt.ptr(), cnt.ptr() + 1);      // uncoloured before this launch
        pg_color_kernel<<<grid_for(n), 256, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), A.row_colors.ptr(), next.ptr(), n, cnt.ptr() + 1);
        count_launch(2);
        AMGXB_LAUNCH_CHECK();
        A.row_colors.swap(next);
        const std::vector<int> h = cnt.to_host(s);
        const int num_uncolored = h[0];
        max_color = h[1];
        if (max_color + 1 >= 64 || prev_uncolored == num_uncolored || num_uncolored <= max_uncolored) break;
        prev_uncolored = num_uncolored;
        if (iter > 100000) fatal(AMGX_RC_INTERNAL, "PARALLEL_GREEDY colouring did not terminate");
    }
    if (max_color + 1 >= 64) {
        cnt.zero(s);
        count_zero_kernel<<<grid_for(n), 256, 0, s>>>(n, A.row_colors.ptr(), cnt.ptr(), cnt.ptr() + 1);
        count_launch();
        if (cnt.to_host(s)[0] > max_uncolored)
            fatal(AMGX_RC_NOT_IMPLEMENTED, "PARALLEL_GREEDY colouring needs more than 63 colours (the reference's max-colour fallback is not implemented); use MIN_MAX");
    }
    A.num_colors = max_color + 1;
    build_color_arrays(A, s);
}

// matrix_coloring_scheme dispatcher used by the multicolour smoothers
void color_matrix(Matrix &A, const std::string &scheme, double max_uncolored_fraction, cudaStream_t s)
{
    if (scheme == "MIN_MAX") color_matrix_min_max(A, max_uncolored_fraction, s);
    else if (scheme == "PARALLEL_GREEDY") color_matrix_parallel_greedy(A, max_uncolored_fraction, s);
    else fatal(AMGX_RC_BAD_CONFIGURATION, "matrix_coloring_scheme '" + scheme + "' is not supported by this engine (MIN_MAX, PARALLEL_GREEDY, or AMGX_matrix_attach_coloring)");
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
