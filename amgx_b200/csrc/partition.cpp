// partition.cpp -- pure host partition planner (no CUDA calls): local renumbering of a row-partitioned
// matrix into [interior | boundary | halo], per-neighbour send maps (the reference's B2L maps) and halo
// layout.  Replaces, for contiguous row partitions, what DistributedManager / DistributedArranger derive
// in the reference (include/distributed/distributed_manager.h:940-952, include/vector.h:10-28,
// src/distributed/distributed_arranger.cu).  Exposed through AMGXB200_partition_plan_create so that the
// CPU (gloo) tests can exercise it without a GPU.
//
// Conventions:
//  * rank r owns global rows [offsets[r], offsets[r+1]).
//  * halo columns are grouped by owning rank (ascending), inside a group sorted by global id; the k-th
//    halo column gets local id n_owned + k.
//  * boundary rows = owned rows with at least one halo column; they are renumbered after the interior
//    rows (both groups keep their relative order), so rows [0, n_interior) never touch remote data and can
//    be processed while the halo exchange is in flight.
//  * send map for neighbour q = my rows (renumbered ids) that q needs, in ascending GLOBAL id order.  The
//    matrix pattern is assumed structurally symmetric across partitions (row i references a column of q
//    <=> q references row i), which makes the send side derivable locally; the engine verifies the
//    neighbours' halo sizes against these maps when the communicator is built.
#include "capi_internal.h"
#include <algorithm>
#include <numeric>

namespace amgxb {

static int owner_of(int64_t g, const int64_t *offsets, int world)
{
    const int64_t *p = std::upper_bound(offsets, offsets + world + 1, g);
    return (int)(p - offsets) - 1;
}

template <class T> static T *dup(const std::vector<T> &v)
{
    T *p = (T *)malloc(sizeof(T) * std::max<size_t>(v.size(), 1));
    if (!p) fatal(AMGX_RC_NO_MEMORY, "partition plan: out of host memory");
    if (!v.empty()) memcpy(p, v.data(), sizeof(T) * v.size());
    return p;
}

void partition_plan_create(AMGXB200_partition_plan *plan, int rank, int world, const int64_t *offsets, int n, int nnz, const int *row_ptrs,
                           const int64_t *cols)
{
    if (world < 1 || rank < 0 || rank >= world || !offsets || n < 0 || !row_ptrs) fatal(AMGX_RC_BAD_PARAMETERS, "partition plan: bad arguments");
    const int64_t lo = offsets[rank], hi = offsets[rank + 1];
    if (hi - lo != n) fatal(AMGX_RC_BAD_PARAMETERS, "partition plan: n does not match the partition offsets");
    if (row_ptrs[n] != nnz) fatal(AMGX_RC_BAD_PARAMETERS, "partition plan: nnz does not match row_ptrs");
    const int64_t n_global = offsets[world];
    // ---- halo columns: unique (owner, global id) ----
    std::vector<int64_t> halo;
    std::vector<char> is_boundary(n, 0);
    for (int i = 0; i < n; i++)
        for (int k = row_ptrs[i]; k < row_ptrs[i + 1]; k++) {
            const int64_t g = cols[k];
            if (g < 0 || g >= n_global) fatal(AMGX_RC_BAD_PARAMETERS, "partition plan: column index out of range");
            if (g < lo || g >= hi) { halo.push_back(g); is_boundary[i] = 1; }
        }
    std::sort(halo.begin(), halo.end());     // global order == (owner, id) order for contiguous partitions
    halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
    std::vector<int> neighbors, halo_offsets{0};
    for (size_t k = 0; k < halo.size(); k++) {
        const int o = owner_of(halo[k], offsets, world);
        if (neighbors.empty() || neighbors.back() != o) {
            if (!neighbors.empty()) halo_offsets.push_back((int)k);
            neighbors.push_back(o);
        }
    }
    halo_offsets.push_back((int)halo.size());
    if (neighbors.empty()) halo_offsets = {0};
    // ---- renumber owned rows: interior first, then boundary ----
    std::vector<int> perm(n);
    int n_interior = 0;
    for (int i = 0; i < n; i++) if (!is_boundary[i]) perm[i] = n_interior++;
    int nb = n_interior;
    for (int i = 0; i < n; i++) if (is_boundary[i]) perm[i] = nb++;
    // ---- send maps: rows that reference a column of neighbour q, ascending global id ----
    const int nn = (int)neighbors.size();
    std::vector<int> nb_index(world, -1);
    for (int q = 0; q < nn; q++) nb_index[neighbors[q]] = q;
    std::vector<std::vector<int>> sends(nn);
    std::vector<int> last_mark(nn, -1);
    for (int i = 0; i < n; i++) {
        if (!is_boundary[i]) continue;
        for (int k = row_ptrs[i]; k < row_ptrs[i + 1]; k++) {
            const int64_t g = cols[k];
            if (g >= lo && g < hi) continue;
            const int q = nb_index[owner_of(g, offsets, world)];
            if (last_mark[q] != i) { last_mark[q] = i; sends[q].push_back(perm[i]); }
        }
    }
    std::vector<int> send_offsets{0}, send_maps;
    for (int q = 0; q < nn; q++) {
        send_maps.insert(send_maps.end(), sends[q].begin(), sends[q].end());
        send_offsets.push_back((int)send_maps.size());
    }
    // ---- local column ids ----
    std::vector<int> local_cols(std::max(nnz, 1));
    for (int k = 0; k < nnz; k++) {
        const int64_t g = cols[k];
        if (g >= lo && g < hi) local_cols[k] = perm[(int)(g - lo)];
        else local_cols[k] = n + (int)(std::lower_bound(halo.begin(), halo.end(), g) - halo.begin());
    }
    plan->n_owned = n;
    plan->n_interior = n_interior;
    plan->n_halo = (int)halo.size();
    plan->num_neighbors = nn;
    plan->neighbors = dup(neighbors);
    plan->send_offsets = dup(send_offsets);
    plan->send_maps = dup(send_maps);
    plan->halo_offsets = dup(halo_offsets);
    plan->halo_global = dup(halo);
    plan->perm_old_to_new = dup(perm);
    plan->local_cols = dup(local_cols);
}

// Arbitrary partition vector -> contiguous partition (DistributedManager::loadDistributedMatrixPartitionVec,
// src/distributed/distributed_manager.cu:1133-1203): rank r's rows become global ids [offsets[r], offsets[r+1]) in increasing
// order of their original global id; new_global[g] = offsets[pv[g]] + #{g' < g : pv[g'] == pv[g]} (the reference's
// ipartition_map).  offsets has world+1 entries, new_global n_global.  Returns false on an out-of-range rank id.
bool partition_vector_to_contiguous(int n_global, int world, const int *pv, int64_t *offsets, int64_t *new_global)
{
    std::vector<int64_t> cnt((size_t)world + 1, 0);
    for (int g = 0; g < n_global; g++) {
        if (pv[g] < 0 || pv[g] >= world) return false;
        cnt[(size_t)pv[g] + 1]++;
    }
    for (int r = 0; r < world; r++) cnt[r + 1] += cnt[r];
    for (int r = 0; r <= world; r++) offsets[r] = cnt[r];
    std::vector<int64_t> next(cnt.begin(), cnt.end() - 1);
    if (new_global)
        for (int g = 0; g < n_global; g++) new_global[g] = next[pv[g]]++;
    return true;
}

// Caller-supplied communication maps (AMGX_matrix_comm_from_maps_one_ring, include/amgx_c.h:325-333) -> global column ids, so
// that a matrix uploaded in LOCAL numbering (owned columns < n, halo columns >= n) can go through the same planner as the
// global uploads.  recv_global[q][k] is the global id of the k-th value neighbour q sends (= global id of its send_maps row k);
// it lands in local halo column recv_maps[q][k].  Returns an empty string, or what is wrong with the maps.
std::string comm_maps_to_global_cols(int n, int nnz, const int *local_cols, int64_t my_offset, int num_neighbors, const int *recv_sizes,
                                     const int *const *recv_maps, const int64_t *const *recv_global, int64_t *cols_out)
{
    int n_halo = 0;
    for (int k = 0; k < nnz; k++) {
        if (local_cols[k] < 0) return "negative column index";
        if (local_cols[k] >= n) n_halo = std::max(n_halo, local_cols[k] - n + 1);
    }
    std::vector<int64_t> halo((size_t)n_halo, -1);
    for (int q = 0; q < num_neighbors; q++)
        for (int k = 0; k < recv_sizes[q]; k++) {
            const int c = recv_maps[q][k];
            if (c < n) return "recv_maps holds an owned index (< n)";
            if (c - n >= n_halo) continue;                     // a halo slot no row references: harmless
            if (halo[c - n] >= 0 && halo[c - n] != recv_global[q][k]) return "a halo column is received from two different rows";
            halo[c - n] = recv_global[q][k];
        }
    for (int k = 0; k < nnz; k++) {
        const int c = local_cols[k];
        if (c < n) cols_out[k] = my_offset + c;
        else {
            if (halo[c - n] < 0) return "a halo column is not covered by recv_maps";
            cols_out[k] = halo[c - n];
        }
    }
    return std::string();
}

}  // namespace amgxb
