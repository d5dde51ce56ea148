// spgemm.cu -- deterministic CSR * CSR product used by the classical-AMG Galerkin operator A_c = R (A P).
//   reference: CSR_Multiply::csr_galerkin_product (src/csr_multiply.cu, src/csr_multiply_detail.cu; called from
//   src/classical/classical_amg_level.cu:581-583) -- a hash SpGEMM whose per-entry sums arrive through atomics in an
//   unspecified order and whose output columns are left in hash order (spmm_no_sort, src/core.cu:507).
// This engine fixes the order instead (so that a sequential CPU restatement reproduces every bit):
//   * one warp per output row; A's row is walked left to right; for each a_ik the lanes take the entries of row k of
//     B (their columns are distinct), so every column receives its contributions in the storage order of A's row;
//   * a product is rounded before it is added (this translation unit is compiled with -fmad=false);
//   * output columns are written in ascending order (rank sort inside the warp).
// Two passes (symbolic, numeric) with open-addressing tables in shared memory; table size escalates per row.
#include "kernels.h"
#include <cub/cub.cuh>

namespace amgxb {
namespace {

__device__ __forceinline__ unsigned hash_col(int c) { return (unsigned)c * 2654435761u; }

// ---- symbolic: number of distinct columns of each row of A*B ----------------------------------------------------
// rows == nullptr: all rows [0, m); else the listed rows.  A row whose distinct count exceeds LIMIT is appended to
// `fail_list` (and its row_nnz left untouched).
template <int TABLE>
__global__ void __launch_bounds__(256) spgemm_symbolic_kernel(int m, const int *__restrict__ rows, const int *__restrict__ arp, const int *__restrict__ aci,
                                                              const int *__restrict__ brp, const int *__restrict__ bci, int *__restrict__ row_nnz,
                                                              int *fail_list, int *fail_count)
{
    extern __shared__ int smem_i[];
    constexpr int LIMIT = TABLE / 4 * 3;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    int *keys = smem_i + (size_t)wib * TABLE;
    for (int w = blockIdx.x * wpb + wib; w < m; w += gridDim.x * wpb) {
        const int row = rows ? rows[w] : w;
        for (int t = lane; t < TABLE; t += 32) keys[t] = -1;
        __syncwarp();
        int count = 0;
        bool failed = false;
        for (int j = arp[row]; j < arp[row + 1] && !failed; j++) {
            const int k = aci[j];
            const int q0 = brp[k], q1 = brp[k + 1];
            for (int q = q0; q < q1; q += 32) {
                int added = 0;
                if (q + lane < q1) {
                    const int c = bci[q + lane];
                    unsigned h = hash_col(c) & (TABLE - 1);
                    while (true) {
                        const int old = atomicCAS(&keys[h], -1, c);
                        if (old == -1) { added = 1; break; }
                        if (old == c) break;
                        h = (h + 1) & (TABLE - 1);
                    }
                }
                count += __reduce_add_sync(0xffffffffu, added);
                if (count > LIMIT) { failed = true; break; }
            }
        }
        __syncwarp();
        if (lane == 0) {
            if (failed) fail_list[atomicAdd(fail_count, 1)] = row;
            else row_nnz[row] = count;
        }
    }
}

// ---- numeric ------------------------------------------------------------------------------------------------------
// Processes the rows whose nnz lies in (lo, hi]; hi <= TABLE*3/4.
template <class T, int TABLE>
__global__ void __launch_bounds__(256) spgemm_numeric_kernel(int m, const int *__restrict__ arp, const int *__restrict__ aci, const T *__restrict__ ava,
                                                             const int *__restrict__ brp, const int *__restrict__ bci, const T *__restrict__ bva,
                                                             const int *__restrict__ crp, int *__restrict__ cci, T *__restrict__ cva, int lo, int hi)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int LIMIT = TABLE / 4 * 3;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    // per warp: vals[TABLE] | lvals[LIMIT] | keys[TABLE] | lkeys[LIMIT]
    unsigned char *base = smem_raw + (size_t)wib * ((size_t)(TABLE + LIMIT) * (sizeof(T) + sizeof(int)));
    T *vals = (T *)base;
    T *lvals = vals + TABLE;
    int *keys = (int *)(lvals + LIMIT);
    int *lkeys = keys + TABLE;
    for (int row = blockIdx.x * wpb + wib; row < m; row += gridDim.x * wpb) {
        const int c0 = crp[row], nnz = crp[row + 1] - c0;
        if (nnz <= lo || nnz > hi) continue;
        for (int t = lane; t < TABLE; t += 32) { keys[t] = -1; vals[t] = (T)0; }
        __syncwarp();
        for (int j = arp[row]; j < arp[row + 1]; j++) {
            const int k = aci[j];
            const T a = ava[j];
            const int q0 = brp[k], q1 = brp[k + 1];
            for (int q = q0 + lane; q < q1; q += 32) {
                const int c = bci[q];
                const T t = a * bva[q];
                unsigned h = hash_col(c) & (TABLE - 1);
                while (true) {
                    const int old = atomicCAS(&keys[h], -1, c);
                    if (old == -1 || old == c) break;
                    h = (h + 1) & (TABLE - 1);
                }
                vals[h] = vals[h] + t;     // this lane is the only one holding column c in this step
            }
            __syncwarp();
        }
        // compact the table into a list
        int cnt = 0;
        for (int t0 = 0; t0 < TABLE; t0 += 32) {
            const int key = keys[t0 + lane];
            const unsigned bal = __ballot_sync(0xffffffffu, key != -1);
            if (key != -1) {
                const int p = cnt + __popc(bal & ((1u << lane) - 1));
                lkeys[p] = key;
                lvals[p] = vals[t0 + lane];
            }
            cnt += __popc(bal);
        }
        __syncwarp();
        // rank sort (columns are distinct)
        for (int e = lane; e < cnt; e += 32) {
            const int key = lkeys[e];
            int rank = 0;
            for (int f = 0; f < cnt; f++) rank += (lkeys[f] < key);
            cci[c0 + rank] = key;
            cva[c0 + rank] = lvals[e];
        }
        __syncwarp();
    }
}

inline int warps_grid(long long rows, int wpb, int cap) { return (int)std::max(1ll, std::min((rows + wpb - 1) / wpb, (long long)cap)); }

template <int TABLE> void run_symbolic(int m, const int *rows, const int *arp, const int *aci, const int *brp, const int *bci, int *row_nnz, int *fail_list,
                                       int *fail_count, cudaStream_t s)
{
    const size_t per_warp = (size_t)TABLE * sizeof(int);
    int wpb = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)200 * 1024 / per_warp));
    const size_t smem = per_warp * wpb;
    auto k = spgemm_symbolic_kernel<TABLE>;
    AMGXB_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<warps_grid(m, wpb, B200_SMS * 16), wpb * 32, smem, s>>>(m, rows, arp, aci, brp, bci, row_nnz, fail_list, fail_count);
    count_launch();
    AMGXB_LAUNCH_CHECK();
}

template <class T, int TABLE> void run_numeric(int m, const int *arp, const int *aci, const T *ava, const int *brp, const int *bci, const T *bva, const int *crp,
                                               int *cci, T *cva, int lo, cudaStream_t s)
{
    constexpr int LIMIT = TABLE / 4 * 3;
    const size_t per_warp = (size_t)(TABLE + LIMIT) * (sizeof(T) + sizeof(int));
    int wpb = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)200 * 1024 / per_warp));
    const size_t smem = per_warp * wpb;
    auto k = spgemm_numeric_kernel<T, TABLE>;
    AMGXB_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<warps_grid(m, wpb, B200_SMS * 16), wpb * 32, smem, s>>>(m, arp, aci, ava, brp, bci, bva, crp, cci, cva, lo, LIMIT);
    count_launch();
    AMGXB_LAUNCH_CHECK();
}

__global__ void max_int_kernel(int n, const int *__restrict__ v, int *out)
{
    int m = 0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) m = max(m, v[t]);
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

}  // namespace

// C = A (m x k) * B (k x ncols).  B's rows must hold distinct columns.  C.row_ptr/col_idx/values are (re)allocated.
void spgemm_csr(int m, const DevBuf<int> &arp, const DevBuf<int> &aci, const DevVec &ava, const DevBuf<int> &brp, const DevBuf<int> &bci, const DevVec &bva,
                DevBuf<int> &crp, DevBuf<int> &cci, DevVec &cva, int *c_nnz, cudaStream_t s)
{
    crp.resize((size_t)m + 1);
    if (m == 0) { crp.zero(s); cci.resize(0); cva.resize(0, ava.prec); *c_nnz = 0; return; }
    DevBuf<int> row_nnz, fail_a, fail_b, counters;
    row_nnz.resize((size_t)m + 1);
    row_nnz.zero(s);
    fail_a.resize(m);
    counters.resize(4);
    counters.zero(s);
    run_symbolic<512>(m, nullptr, arp.ptr(), aci.ptr(), brp.ptr(), bci.ptr(), row_nnz.ptr(), fail_a.ptr(), counters.ptr(), s);
    int nfail = counters.to_host(s)[0];
    if (nfail > 0) {
        fail_b.resize(nfail);
        run_symbolic<4096>(nfail, fail_a.ptr(), arp.ptr(), aci.ptr(), brp.ptr(), bci.ptr(), row_nnz.ptr(), fail_b.ptr(), counters.ptr() + 1, s);
        const int nfail2 = counters.to_host(s)[1];
        if (nfail2 > 0) {
            run_symbolic<32768>(nfail2, fail_b.ptr(), arp.ptr(), aci.ptr(), brp.ptr(), bci.ptr(), row_nnz.ptr(), fail_a.ptr(), counters.ptr() + 2, s);
            if (counters.to_host(s)[2] > 0) fatal(AMGX_RC_NOT_IMPLEMENTED, "Galerkin product: a coarse row has more than 24576 entries");
        }
    }
    max_int_kernel<<<std::min(ceil_div(m, 256), 1024), 256, 0, s>>>(m, row_nnz.ptr(), counters.ptr() + 3);
    count_launch();
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, row_nnz.ptr(), crp.ptr(), m + 1, s);
    DevBytes tmp;
    tmp.resize(tb);
    cub::DeviceScan::ExclusiveSum(tmp.p, tb, row_nnz.ptr(), crp.ptr(), m + 1, s);
    count_launch();
    int nnz = 0;
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(&nnz, crp.ptr() + m, sizeof(int), cudaMemcpyDeviceToHost, s));
    const int max_row = counters.to_host(s)[3];
    *c_nnz = nnz;
    cci.resize((size_t)std::max(nnz, 1));
    cva.resize((size_t)std::max(nnz, 1), ava.prec);
    if (max_row > 6144) fatal(AMGX_RC_NOT_IMPLEMENTED, "Galerkin product: a coarse row has more than 6144 entries");
    AMGXB_DISPATCH_VEC(ava.prec, {
        const VecT *av = ava.as<VecT>(), *bv = bva.as<VecT>();
        VecT *cv = cva.as<VecT>();
        run_numeric<VecT, 128>(m, arp.ptr(), aci.ptr(), av, brp.ptr(), bci.ptr(), bv, crp.ptr(), cci.ptr(), cv, 0, s);
        if (max_row > 96) run_numeric<VecT, 512>(m, arp.ptr(), aci.ptr(), av, brp.ptr(), bci.ptr(), bv, crp.ptr(), cci.ptr(), cv, 96, s);
        if (max_row > 384) run_numeric<VecT, 2048>(m, arp.ptr(), aci.ptr(), av, brp.ptr(), bci.ptr(), bv, crp.ptr(), cci.ptr(), cv, 384, s);
        if (max_row > 1536) run_numeric<VecT, 8192>(m, arp.ptr(), aci.ptr(), av, brp.ptr(), bci.ptr(), bv, crp.ptr(), cci.ptr(), cv, 1536, s);
    });
}

}  // namespace amgxb
