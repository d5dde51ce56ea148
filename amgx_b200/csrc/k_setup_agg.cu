// k_setup_agg.cu -- setup producers of the aggregation hierarchy (parity-critical, not
// roofline-critical): SIZE_2 handshake selector, restriction pattern, Galerkin product.
//
// Follows, stage by stage:
//   edge weights        include/aggregation/selectors/common_selector.h:19-29, 63-137 (float, hash-perturbed)
//   handshake matching  src/aggregation/selectors/size2_selector.cu:224-323, 787-847
//   leftovers           src/aggregation/selectors/size2_selector.cu:508-566, 424-440, 853-887 (deterministic path)
//   renumbering         src/aggregation/selectors/agg_selector.cu:18-43
//   R pattern           src/aggregation/aggregation_amg_level.cu:237-299 (stable sort by aggregate)
//   Galerkin Ac=R A P   src/aggregation/coarseAgenerators/low_deg_coarse_A_generator.cu:1135-1320
// Deliberate differences (documented in DESIGN.md):
//   * the reference's findStrongestNeighbour kernel merges singletons by writing aggregates[] that
//     other threads of the same launch read (a benign race, outcome timing dependent); here every
//     thread sees the aggregates of the previous step ("snapshot"), which is one of the outcomes
//     the reference can produce and is reproducible;
//   * coarse rows are emitted with ascending column order (the reference emits hash-table order);
//     entries and their summation order (fine row ascending, in-row order) are the same.
#include "kernels.h"
#include <cub/cub.cuh>

namespace amgxb {
namespace {

__host__ __device__ inline unsigned hash_val(unsigned a, unsigned seed)
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

template <class MatT>
__global__ void edge_weights_kernel(const int *__restrict__ row_ptr, const int *__restrict__ col, const int *__restrict__ diag,
                                    const MatT *__restrict__ val, int n, int bsq, int entry, int weight_formula,
                                    float *__restrict__ w)
{
    // one thread per row walks its entries (the reference uses one thread per non-zero with a
    // row_indices array; the per-entry arithmetic is identical)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int di = diag[i];
        const MatT aii = di >= 0 ? val[(size_t)di * bsq + entry] : (MatT)0;
        for (int k = row_ptr[i]; k < row_ptr[i + 1]; k++) {
            const int j = col[k];
            if (i == j || j >= n) { w[k] = -1.0f; continue; }
            const int dj = diag[j];
            const MatT ajj = dj >= 0 ? val[(size_t)dj * bsq + entry] : (MatT)0;
            const MatT mx = fabs(aii) > fabs(ajj) ? fabs(aii) : fabs(ajj);   // max(|a_ii|,|a_jj|) in MatT
            const float den = (float)mx;
            MatT kval = 0;
            bool found = false;
            for (int kk = row_ptr[j]; kk < row_ptr[j + 1]; kk++)
                if (col[kk] == i) { kval = val[(size_t)kk * bsq + entry]; found = true; break; }
            float ew = 0.0f;
            if (found) {
                if (weight_formula == 0) {
                    const MatT ssum = fabs(val[(size_t)k * bsq + entry]) + fabs(kval);
                    const double t = 0.5 * (double)ssum;
                    ew = (float)(t / (double)den);
                } else {
                    const MatT rz = val[(size_t)k * bsq + entry] / aii + kval / ajj;
                    ew = (float)(-0.5 * (double)(float)rz);
                }
            }
            const unsigned h = hash_val((unsigned)min(i, j), (unsigned)max(i, j));
            const float small_fraction = __fdiv_rn(__fmul_rn(1e-5f, __uint2float_rn(h)), 4294967296.0f);
            ew = __fmaf_rn(small_fraction, ew, ew);
            w[k] = ew;
        }
    }
}

// phase-1 handshake: strongest unaggregated neighbour, or merge target when every neighbour is aggregated
__global__ void find_strongest_kernel(const int *__restrict__ row_ptr, const int *__restrict__ col, const float *__restrict__ w, int n,
                                      const int *__restrict__ agg, int *__restrict__ strongest, int *__restrict__ merge_to,
                                      int merge_singletons)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        merge_to[t] = -1;
        if (agg[t] != -1) continue;
        int s_un = -1, s_ag = -1;
        float m_un = 0.f, m_ag = 0.f;
        for (int k = row_ptr[t]; k < row_ptr[t + 1]; k++) {
            const int j = col[k];
            const float wt = w[k];
            if (j == t || j >= n) continue;
            const int aj = agg[j];
            if (aj == -1 && (wt > m_un || (wt == m_un && j > s_un))) { m_un = wt; s_un = j; }
            else if (aj != -1 && (wt > m_ag || (wt == m_ag && j > s_ag))) { m_ag = wt; s_ag = j; }
        }
        if (s_un == -1 && s_ag != -1) {
            merge_to[t] = merge_singletons ? agg[s_ag] : t;
        } else if (s_un != -1) {
            strongest[t] = s_un;
        } else {
            strongest[t] = t;
        }
    }
}

__global__ void match_edges_kernel(int n, int *__restrict__ agg, const int *__restrict__ strongest, const int *__restrict__ merge_to)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        if (agg[t] != -1) continue;
        if (merge_to[t] != -1) { agg[t] = merge_to[t]; continue; }
        const int pm = strongest[t];
        if (pm < 0) continue;
        // strongest[pm] is fresh iff pm was unaggregated and did not merge in this step
        if (merge_to[pm] == -1 && strongest[pm] == t) agg[t] = (pm > t) ? t : pm;
    }
}

__global__ void count_unaggregated_kernel(int n, const int *__restrict__ agg, int *count)
{
    int c = 0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) c += (agg[t] == -1);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

__global__ void merge_candidates_kernel(const int *__restrict__ row_ptr, const int *__restrict__ col, const float *__restrict__ w, int n,
                                        const int *__restrict__ agg, int *__restrict__ cand)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        if (agg[t] != -1) continue;
        float m_ag = 0.f;
        int s_ag = -1;
        for (int k = row_ptr[t]; k < row_ptr[t + 1]; k++) {
            const float wt = w[k];
            const int j = col[k];
            if (j == t || j >= n) continue;
            if (agg[j] != -1 && (wt > m_ag || (wt == m_ag && j > s_ag))) { m_ag = wt; s_ag = j; }
        }
        cand[t] = (s_ag != -1) ? agg[s_ag] : t;
    }
}

__global__ void join_kernel(int n, int *__restrict__ agg, const int *__restrict__ cand)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
        if (agg[t] == -1 && cand[t] != -1) agg[t] = cand[t];
}

__global__ void singletons_kernel(int n, int *__restrict__ agg)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
        if (agg[t] == -1) agg[t] = t;
}

__global__ void mark_kernel(int n, const int *__restrict__ agg, int *__restrict__ scratch)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) scratch[agg[t]] = 1;
}
__global__ void relabel_kernel(int n, int *__restrict__ agg, const int *__restrict__ scratch)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) agg[t] = scratch[agg[t]];
}
__global__ void iota_kernel(int n, int *__restrict__ v)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) v[t] = t;
}
__global__ void fill_int_kernel(int n, int *__restrict__ v, int x)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) v[t] = x;
}
// sorted keys -> CSR offsets (every key in [0,n_keys) may be empty)
__global__ void offsets_from_sorted_kernel(int n, const int *__restrict__ keys, int n_keys, int *__restrict__ offsets)
{
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p <= n; p += gridDim.x * blockDim.x) {
        const int prev = (p == 0) ? -1 : keys[p - 1];
        const int cur = (p == n) ? n_keys : keys[p];
        for (int I = prev + 1; I <= cur; I++) offsets[I] = p;
    }
}

inline int grid_for(long long n) { return std::max(1, std::min(ceil_div(n, 256), B200_SMS * 16)); }

template <class T> T read_scalar(const T *dptr, cudaStream_t s)
{
    T h;
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(&h, dptr, sizeof(T), cudaMemcpyDeviceToHost, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    return h;
}

// ------------------------------------ Galerkin ------------------------------------
__global__ void galerkin_keys_kernel(const int *__restrict__ row_ptr, const int *__restrict__ col, const int *__restrict__ agg, int n,
                                     unsigned long long *__restrict__ keys, int *__restrict__ idx)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long I = (unsigned long long)(unsigned)agg[i] << 32;
        for (int k = row_ptr[i]; k < row_ptr[i + 1]; k++) {
            keys[k] = I | (unsigned)agg[col[k]];
            idx[k] = k;
        }
    }
}
__global__ void head_flags_kernel(int nnz, const unsigned long long *__restrict__ keys, int *__restrict__ flags)
{
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += gridDim.x * blockDim.x)
        flags[p] = (p == 0 || keys[p] != keys[p - 1]) ? 1 : 0;
}
// pos = exclusive scan of flags.  At heads: column, segment start, row count.
__global__ void coarse_structure_kernel(int nnz, const unsigned long long *__restrict__ keys, const int *__restrict__ flags,
                                        const int *__restrict__ pos, int *__restrict__ col_c, int *__restrict__ seg_start,
                                        int *__restrict__ row_count)
{
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += gridDim.x * blockDim.x) {
        if (flags[p]) {
            const int e = pos[p];
            col_c[e] = (int)(unsigned)(keys[p] & 0xffffffffull);
            seg_start[e] = p;
            atomicAdd(&row_count[(int)(keys[p] >> 32)], 1);
        }
    }
}
template <class MatT>
__global__ void coarse_values_kernel(int nnz_c, int nnz, int bsq, const int *__restrict__ seg_start, const int *__restrict__ idx,
                                     const MatT *__restrict__ val, MatT *__restrict__ val_c)
{
    const long long total = (long long)nnz_c * bsq;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(t / bsq), m = (int)(t % bsq);
        const int p0 = seg_start[e], p1 = (e + 1 < nnz_c) ? seg_start[e + 1] : nnz;
        MatT acc = val[(size_t)idx[p0] * bsq + m];
        for (int p = p0 + 1; p < p1; p++) acc = acc + val[(size_t)idx[p] * bsq + m];
        val_c[t] = acc;
    }
}

template <class MatT> __global__ void extract_diag_kernel(int n, int bsq, const int *__restrict__ diag, const MatT *__restrict__ val, MatT *__restrict__ d)
{
    const long long total = (long long)n * bsq;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / bsq), m = (int)(t % bsq);
        const int k = diag[i];
        d[t] = k >= 0 ? val[(size_t)k * bsq + m] : (MatT)0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// SIZE_4 selector (src/aggregation/selectors/size4_selector.cu:96-224; kernels include/aggregation/selectors/common_selector.h:178-525):
// handshake pairs, then handshake pairs of pairs, then the deterministic merge of the leftovers.  In the deterministic flow none
// of these kernels reads what another thread of the same launch writes, so a launch is a pure function of its inputs.
// ---------------------------------------------------------------------------------------------------------------
__global__ void s4_find_strongest_nomerge(const int *__restrict__ rp, const int *__restrict__ ci, const float *__restrict__ w, int n,
                                          const int *__restrict__ partner, int *strongest)
{
    for (int tid = blockIdx.x * blockDim.x + threadIdx.x; tid < n; tid += gridDim.x * blockDim.x) {
        if (partner[tid] != -1) continue;
        float max_w = 0.f;
        int best = -1;
        for (int j = rp[tid]; j < rp[tid + 1]; j++) {
            const int jc = ci[j];
            if (tid == jc || jc >= n) continue;
            const float wt = w[j];
            if (partner[jc] == -1 && (wt > max_w || (wt == max_w && jc > best))) { max_w = wt; best = jc; }
        }
        if (best != -1) strongest[tid] = best;      // nothing found: the previous proposal stays (deterministic flow)
    }
}
__global__ void s4_match_edges(int n, int *partner, int *aggregates, const int *__restrict__ strongest)
{
    for (int tid = blockIdx.x * blockDim.x + threadIdx.x; tid < n; tid += gridDim.x * blockDim.x) {
        if (partner[tid] != -1) continue;
        const int pm = strongest[tid];
        if (pm != -1 && strongest[pm] == tid) {
            partner[tid] = pm;
            aggregates[tid] = pm > tid ? tid : pm;
        }
    }
}
__global__ void s4_count_minus_one(int n, const int *__restrict__ v, int *count)
{
    int c = 0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) c += (v[t] == -1);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}
__global__ void s4_assign_unassigned(int n, int *partner)
{
    for (int tid = blockIdx.x * blockDim.x + threadIdx.x; tid < n; tid += gridDim.x * blockDim.x)
        if (partner[tid] == -1) partner[tid] = tid;
}
__global__ void s4_store_weight(const int *__restrict__ rp, const int *__restrict__ ci, const float *__restrict__ w, int n, const int *__restrict__ aggregated,
                                const int *__restrict__ aggregates, int *strongest, const int *__restrict__ partner, float *wsn)
{
    for (int tid = blockIdx.x * blockDim.x + threadIdx.x; tid < n; tid += gridDim.x * blockDim.x) {
        if (aggregated[tid] != -1) continue;
        const int p = partner[tid];
        float max_w = 0.f;
        int best = -1;
        for (int j = rp[tid]; j < rp[tid + 1]; j++) {
            const int jc = ci[j];
            if (tid == jc || jc >= n) continue;
            const float wt = w[j];
            if (aggregated[jc] == -1 && jc != p && (wt > max_w || (wt == max_w && jc > best))) { max_w = wt; best = jc; }
        }
        if (best != -1) {
            wsn[tid] = max_w;
            strongest[tid] = aggregates[best];
        }
    }
}
__global__ void s4_agree(int n, int *aggregated, int *strongest, const float *__restrict__ wsn, const int *__restrict__ partner)
{
    for (int tid = blockIdx.x * blockDim.x + threadIdx.x; tid < n; tid += gridDim.x * blockDim.x) {
        if (aggregated[tid] != -1) continue;
        const int p = partner[tid];
        const float mine = wsn[tid];
        float theirs = -1;
        if (p != -1) theirs = wsn[p];
        if (mine < 0.f && theirs < 0.f) {             // every neighbour is aggregated: the pair stays as it is
            aggregated[tid] = 1;
            strongest[tid] = -1;
        } else if (mine < theirs) {
            strongest[tid] = strongest[p];            // the weaker half adopts its partner's proposal (the partner does not write: its weight is larger)
        }
    }
}
__global__ void s4_match_aggregates(int n, int *aggregates, int *aggregated, const int *__restrict__ strongest)
{
    for (int tid = blockIdx.x * blockDim.x + threadIdx.x; tid < n; tid += gridDim.x * blockDim.x) {
        if (aggregated[tid] != -1) continue;
        const int pm = strongest[tid];
        if (pm == -1) continue;
        const int mine = aggregates[tid];
        if (strongest[pm] == mine) {
            aggregated[tid] = 1;
            aggregates[tid] = pm > mine ? mine : pm;
        }
    }
}
__global__ void s4_merge_candidates(const int *__restrict__ rp, const int *__restrict__ ci, const float *__restrict__ w, int n, const int *__restrict__ aggregates,
                                    const int *__restrict__ aggregated, int *cand)
{
    for (int tid = blockIdx.x * blockDim.x + threadIdx.x; tid < n; tid += gridDim.x * blockDim.x) {
        if (aggregated[tid] != -1) continue;
        float max_w = 0.f;
        int best = -1;
        for (int j = rp[tid]; j < rp[tid + 1]; j++) {
            const int jc = ci[j];
            if (tid == jc || jc >= n) continue;
            if (aggregated[jc] != -1) {
                const float wt = w[j];
                if (wt > max_w || (wt == max_w && jc > best)) { max_w = wt; best = jc; }
            }
        }
        cand[tid] = best != -1 ? aggregates[best] : tid;
    }
}
__global__ void s4_join(int n, int *aggregates, int *aggregated, const int *__restrict__ cand)
{
    for (int tid = blockIdx.x * blockDim.x + threadIdx.x; tid < n; tid += gridDim.x * blockDim.x)
        if (aggregated[tid] == -1 && cand[tid] != -1) { aggregates[tid] = cand[tid]; aggregated[tid] = 1; }
}
__global__ void fill_float_kernel(int n, float *v, float x)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) v[t] = x;
}

// labels (minimum member row ids) -> 0..n_agg-1 in label order (renumberAndCountAggregates, agg_selector.cu:18-43)
int renumber_aggregates(int n, DevBuf<int> &aggregates, cudaStream_t s)
{
    const int g = grid_for(n);
    DevBuf<int> scratch;
    scratch.resize(n + 1);
    scratch.zero(s);
    mark_kernel<<<g, 256, 0, s>>>(n, aggregates.ptr(), scratch.ptr());
    count_launch();
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, scratch.ptr(), scratch.ptr(), n + 1, s);
    DevBytes tmp;
    tmp.resize(tmp_bytes);
    cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, scratch.ptr(), scratch.ptr(), n + 1, s);
    relabel_kernel<<<g, 256, 0, s>>>(n, aggregates.ptr(), scratch.ptr());
    count_launch();
    AMGXB_LAUNCH_CHECK();
    return read_scalar(scratch.ptr() + n, s);
}

}  // namespace

int size4_select(const Matrix &A, const AggSetupParams &prm, DevBuf<int> &aggregates, cudaStream_t s)
{
    const int n = A.n;
    aggregates.resize(n);
    if (n == 0) return 0;
    const int g = grid_for(n);
    DevBuf<float> w, wsn;
    DevBuf<int> strongest, partner, aggregated, counter, cand;
    w.resize(std::max(A.nnz, 1));
    wsn.resize(n);
    strongest.resize(n);
    partner.resize(n);
    aggregated.resize(n);
    cand.resize(n);
    counter.resize(1);
    iota_kernel<<<g, 256, 0, s>>>(n, aggregates.ptr());
    fill_int_kernel<<<g, 256, 0, s>>>(n, strongest.ptr(), -1);
    fill_int_kernel<<<g, 256, 0, s>>>(n, partner.ptr(), -1);
    count_launch(3);
    const int bsq = A.bs();
    const int entry = prm.edge_weight_component * A.bx + prm.edge_weight_component;
    if (A.mat_prec == Prec::F64)
        edge_weights_kernel<double><<<g, 256, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), A.diag_idx.ptr(), A.values.as<double>(), n, bsq, entry, prm.weight_formula, w.ptr());
    else
        edge_weights_kernel<float><<<g, 256, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), A.diag_idx.ptr(), A.values.as<float>(), n, bsq, entry, prm.weight_formula, w.ptr());
    count_launch();
    auto count_m1 = [&](const int *v) {
        counter.zero(s);
        s4_count_minus_one<<<g, 256, 0, s>>>(n, v, counter.ptr());
        count_launch();
        AMGXB_LAUNCH_CHECK();
        return read_scalar(counter.ptr(), s);
    };
    // ---- pairs.  The reference counts -1 over its 3n-entry partner_index array, whose upper 2n entries are still -1 in this phase
    // (size4_selector.cu:126, 158): the count never reaches 0 nor the tolerance, the loop ends on "no progress" or the iteration cap.
    int num_unassigned = n, prev = n, icount = 0;
    do {
        s4_find_strongest_nomerge<<<g, 256, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), w.ptr(), n, partner.ptr(), strongest.ptr());
        s4_match_edges<<<g, 256, 0, s>>>(n, partner.ptr(), aggregates.ptr(), strongest.ptr());
        count_launch(2);
        prev = num_unassigned;
        num_unassigned = count_m1(partner.ptr()) + 2 * n;
        icount++;
    } while (!(num_unassigned == 0 || icount > prm.max_iterations || 1.0 * num_unassigned / n < prm.max_unassigned || prev == num_unassigned));
    s4_assign_unassigned<<<g, 256, 0, s>>>(n, partner.ptr());
    // ---- pairs of pairs
    fill_float_kernel<<<g, 256, 0, s>>>(n, wsn.ptr(), -1.f);
    fill_int_kernel<<<g, 256, 0, s>>>(n, aggregated.ptr(), -1);
    count_launch(3);
    icount = 0;
    num_unassigned = prev = n;
    do {
        s4_store_weight<<<g, 256, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), w.ptr(), n, aggregated.ptr(), aggregates.ptr(), strongest.ptr(), partner.ptr(), wsn.ptr());
        s4_agree<<<g, 256, 0, s>>>(n, aggregated.ptr(), strongest.ptr(), wsn.ptr(), partner.ptr());
        s4_match_aggregates<<<g, 256, 0, s>>>(n, aggregates.ptr(), aggregated.ptr(), strongest.ptr());
        count_launch(3);
        prev = num_unassigned;
        num_unassigned = count_m1(aggregated.ptr());
        icount++;
    } while (!(num_unassigned == 0 || icount > prm.max_iterations || 1.0 * num_unassigned / n < prm.max_unassigned || prev == num_unassigned));
    // ---- leftovers join the aggregate of their strongest aggregated neighbour (deterministic: candidates first, then join)
    fill_int_kernel<<<g, 256, 0, s>>>(n, cand.ptr(), -1);
    count_launch();
    while (num_unassigned != 0) {
        s4_merge_candidates<<<g, 256, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), w.ptr(), n, aggregates.ptr(), aggregated.ptr(), cand.ptr());
        s4_join<<<g, 256, 0, s>>>(n, aggregates.ptr(), aggregated.ptr(), cand.ptr());
        count_launch(2);
        num_unassigned = count_m1(aggregated.ptr());
    }
    return renumber_aggregates(n, aggregates, s);
}

int size2_select(const Matrix &A, const AggSetupParams &prm, DevBuf<int> &aggregates, cudaStream_t s)
{
    if (prm.two_phase) fatal(AMGX_RC_NOT_IMPLEMENTED, "SIZE_2 selector: handshaking_phases=2 is not implemented");
    const int n = A.n;
    aggregates.resize(n);
    if (n == 0) return 0;
    DevBuf<float> w;
    w.resize(std::max(A.nnz, 1));
    DevBuf<int> strongest, merge_to, counter, cand;
    strongest.resize(n);
    merge_to.resize(n);
    counter.resize(1);
    const int g = grid_for(n);
    fill_int_kernel<<<g, 256, 0, s>>>(n, aggregates.ptr(), -1);
    fill_int_kernel<<<g, 256, 0, s>>>(n, strongest.ptr(), -1);
    count_launch(2);
    const int bsq = A.bs();
    const int entry = prm.edge_weight_component * A.bx + prm.edge_weight_component;
    if (A.mat_prec == Prec::F64)
        edge_weights_kernel<double><<<g, 256, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), A.diag_idx.ptr(), A.values.as<double>(), n, bsq, entry,
                                                      prm.weight_formula, w.ptr());
    else
        edge_weights_kernel<float><<<g, 256, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), A.diag_idx.ptr(), A.values.as<float>(), n, bsq, entry,
                                                     prm.weight_formula, w.ptr());
    count_launch();
    AMGXB_LAUNCH_CHECK();

    // The reference builds this loop with EXPERIMENTAL_ITERATIVE_MATCHING (size2_selector.cu:22, 808-847): the count of
    // unaggregated rows is taken after even iterations only and consumed one iteration later, so the loop runs an even
    // number of handshake steps and the exit test (and the merge phase's first test) see a one-step-stale count.
    int num_unassigned = n, prev = n, icount = 0, sflag = 1, pending = n;
    do {
        find_strongest_kernel<<<g, 256, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), w.ptr(), n, aggregates.ptr(), strongest.ptr(), merge_to.ptr(),
                                                prm.merge_singletons);
        match_edges_kernel<<<g, 256, 0, s>>>(n, aggregates.ptr(), strongest.ptr(), merge_to.ptr());
        count_launch(2);
        sflag = (icount & 1);
        if (sflag == 0) {
            counter.zero(s);
            count_unaggregated_kernel<<<g, 256, 0, s>>>(n, aggregates.ptr(), counter.ptr());
            count_launch();
            AMGXB_LAUNCH_CHECK();
            pending = read_scalar(counter.ptr(), s);
        } else {
            prev = num_unassigned;
            num_unassigned = pending;
        }
        icount++;
    } while (sflag == 0 || !(num_unassigned == 0 || icount > prm.max_iterations || 1.0 * num_unassigned / n < prm.max_unassigned || num_unassigned == prev));

    if (prm.merge_singletons) {
        cand.resize(n);
        fill_int_kernel<<<g, 256, 0, s>>>(n, cand.ptr(), -1);
        count_launch();
        while (num_unassigned != 0) {
            merge_candidates_kernel<<<g, 256, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), w.ptr(), n, aggregates.ptr(), cand.ptr());
            join_kernel<<<g, 256, 0, s>>>(n, aggregates.ptr(), cand.ptr());
            counter.zero(s);
            count_unaggregated_kernel<<<g, 256, 0, s>>>(n, aggregates.ptr(), counter.ptr());
            count_launch(3);
            AMGXB_LAUNCH_CHECK();
            num_unassigned = read_scalar(counter.ptr(), s);
        }
    } else {
        singletons_kernel<<<g, 256, 0, s>>>(n, aggregates.ptr());
        count_launch();
    }

    // renumber: labels (minimum member row ids) -> 0..n_agg-1 in label order
    DevBuf<int> scratch;
    scratch.resize(n + 1);
    scratch.zero(s);
    mark_kernel<<<g, 256, 0, s>>>(n, aggregates.ptr(), scratch.ptr());
    count_launch();
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, scratch.ptr(), scratch.ptr(), n + 1, s);
    DevBytes tmp;
    tmp.resize(tmp_bytes);
    cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, scratch.ptr(), scratch.ptr(), n + 1, s);
    relabel_kernel<<<g, 256, 0, s>>>(n, aggregates.ptr(), scratch.ptr());
    count_launch();
    AMGXB_LAUNCH_CHECK();
    return read_scalar(scratch.ptr() + n, s);
}

void build_restriction(const DevBuf<int> &aggregates, int n, int n_agg, DevBuf<int> &Rp, DevBuf<int> &Rc, cudaStream_t s)
{
    Rp.resize(n_agg + 1);
    Rc.resize(n);
    if (n == 0) { Rp.zero(s); return; }
    DevBuf<int> keys_out, vals_in;
    keys_out.resize(n);
    vals_in.resize(n);
    iota_kernel<<<grid_for(n), 256, 0, s>>>(n, vals_in.ptr());
    count_launch();
    int bits = 1;
    while ((1ll << bits) < (long long)n_agg + 1) bits++;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, aggregates.ptr(), keys_out.ptr(), vals_in.ptr(), Rc.ptr(), n, 0, bits, s);
    DevBytes tmp;
    tmp.resize(tmp_bytes);
    cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, aggregates.ptr(), keys_out.ptr(), vals_in.ptr(), Rc.ptr(), n, 0, bits, s);
    offsets_from_sorted_kernel<<<grid_for(n + 1), 256, 0, s>>>(n, keys_out.ptr(), n_agg, Rp.ptr());
    count_launch();
    AMGXB_LAUNCH_CHECK();
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));   // temporaries die here
}

void galerkin_aggregation(const Matrix &A, const DevBuf<int> &aggregates, int n_agg, Matrix &Ac, cudaStream_t s)
{
    const int n = A.n, nnz = A.nnz, bsq = A.bs();
    Ac.rsc = A.rsc;
    Ac.mode = A.mode;
    Ac.mat_prec = A.mat_prec;
    Ac.vec_prec = A.vec_prec;
    Ac.bx = A.bx;
    Ac.by = A.by;
    Ac.n = n_agg;
    Ac.n_cols = n_agg;
    Ac.has_ext_diag = false;
    Ac.level = A.level + 1;
    Ac.row_ptr.resize(n_agg + 1);
    if (nnz == 0 || n_agg == 0) {
        Ac.nnz = 0;
        Ac.row_ptr.zero(s);
        Ac.col_idx.resize(0);
        Ac.values.resize(0, A.mat_prec);
        return;
    }
    DevBuf<unsigned long long> keys, keys_sorted;
    DevBuf<int> idx, idx_sorted, flags, pos;
    keys.resize(nnz);
    keys_sorted.resize(nnz);
    idx.resize(nnz);
    idx_sorted.resize(nnz);
    galerkin_keys_kernel<<<grid_for(n), 256, 0, s>>>(A.row_ptr.ptr(), A.col_idx.ptr(), aggregates.ptr(), n, keys.ptr(), idx.ptr());
    count_launch();
    int bits = 1;
    while ((1ll << bits) < (long long)n_agg) bits++;   // row part of the key; the column part (incl. halo aggregates) uses the low 32 bits
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.ptr(), keys_sorted.ptr(), idx.ptr(), idx_sorted.ptr(), nnz, 0, 32 + bits, s);
    DevBytes tmp;
    tmp.resize(tmp_bytes);
    cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.ptr(), keys_sorted.ptr(), idx.ptr(), idx_sorted.ptr(), nnz, 0, 32 + bits, s);
    keys.release();
    idx.release();
    flags.resize(nnz + 1);
    pos.resize(nnz + 1);
    flags.zero(s);
    head_flags_kernel<<<grid_for(nnz), 256, 0, s>>>(nnz, keys_sorted.ptr(), flags.ptr());
    count_launch();
    size_t tmp2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp2, flags.ptr(), pos.ptr(), nnz + 1, s);
    if (tmp2 > tmp.cap) tmp.resize(tmp2);
    cub::DeviceScan::ExclusiveSum(tmp.p, tmp2, flags.ptr(), pos.ptr(), nnz + 1, s);
    const int nnz_c = read_scalar(pos.ptr() + nnz, s);
    Ac.nnz = nnz_c;
    Ac.col_idx.resize(nnz_c);
    Ac.values.resize((size_t)nnz_c * bsq, A.mat_prec);
    DevBuf<int> seg_start, row_count;
    seg_start.resize(nnz_c);
    row_count.resize(n_agg + 1);
    row_count.zero(s);
    coarse_structure_kernel<<<grid_for(nnz), 256, 0, s>>>(nnz, keys_sorted.ptr(), flags.ptr(), pos.ptr(), Ac.col_idx.ptr(), seg_start.ptr(),
                                                          row_count.ptr());
    count_launch();
    size_t tmp3 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp3, row_count.ptr(), Ac.row_ptr.ptr(), n_agg + 1, s);
    if (tmp3 > tmp.cap) tmp.resize(tmp3);
    cub::DeviceScan::ExclusiveSum(tmp.p, tmp3, row_count.ptr(), Ac.row_ptr.ptr(), n_agg + 1, s);
    if (A.mat_prec == Prec::F64)
        coarse_values_kernel<double><<<grid_for((long long)nnz_c * bsq), 256, 0, s>>>(nnz_c, nnz, bsq, seg_start.ptr(), idx_sorted.ptr(),
                                                                                      A.values.as<double>(), Ac.values.as<double>());
    else
        coarse_values_kernel<float><<<grid_for((long long)nnz_c * bsq), 256, 0, s>>>(nnz_c, nnz, bsq, seg_start.ptr(), idx_sorted.ptr(),
                                                                                     A.values.as<float>(), Ac.values.as<float>());
    count_launch();
    AMGXB_LAUNCH_CHECK();
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    Ac.initialized = true;
}

void extract_diagonal(const Matrix &A, DevVec &d, cudaStream_t s)
{
    const int bsq = A.bs();
    d.resize((size_t)A.n * bsq, A.mat_prec);
    if (A.n == 0) return;
    if (A.has_ext_diag) {
        AMGXB_CUDA_CHECK(cudaMemcpyAsync(d.ptr(), (const char *)A.values.ptr() + (size_t)A.nnz * bsq * prec_size(A.mat_prec),
                                         (size_t)A.n * bsq * prec_size(A.mat_prec), cudaMemcpyDeviceToDevice, s));
        return;
    }
    if (A.mat_prec == Prec::F64)
        extract_diag_kernel<double><<<grid_for((long long)A.n * bsq), 256, 0, s>>>(A.n, bsq, A.diag_idx.ptr(), A.values.as<double>(), d.as<double>());
    else
        extract_diag_kernel<float><<<grid_for((long long)A.n * bsq), 256, 0, s>>>(A.n, bsq, A.diag_idx.ptr(), A.values.as<float>(), d.as<float>());
    count_launch();
    AMGXB_LAUNCH_CHECK();
}

}  // namespace amgxb
