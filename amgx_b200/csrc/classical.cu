// classical.cu -- classical (Ruge-Stueben) AMG: setup producers and transfer operators of a classical level.
//   strength AHAT + PMIS weights       src/classical/strength/strength_base.cu:185-330
//   PMIS                               src/classical/selectors/pmis.cu:221-266, 370-466, 468-622
//   aggressive PMIS (S2 second pass)   src/classical/selectors/aggressive_pmis.cu:22-150, selector.cu:116-230, 427-580, 942-1004, 1057-1070
//   D2 "extended+i" interpolation      src/classical/interpolators/distance2.cu:600-716, 1178-1362, 1562-1796
//   MULTIPASS interpolation            src/classical/interpolators/multipass.cu:94-147, 244-287, 1057-1206, 1538-1720
//   truncation (interp_max_elements)   src/truncate.cu:352-456, 78-92, 783-862
//   R = P^T, A_c = R A P               src/classical/classical_amg_level.cu:440-468, 501-586
//   restrict / prolongate              src/classical/classical_amg_level.cu:590-644, 851-913
// Order conventions (see DESIGN.md "classical setup"): rows of P are emitted in the reference's own order (its
// hash-table slot order, emulated: it decides which of several equal weights the max-elements truncation keeps); sums
// run left to right in storage order and a product is rounded before it is added (this file is compiled with
// -fmad=false).  Selection arrays, the pattern and the row order of P are bit-comparable with the reference on the
// finest level; weights agree to rounding.
#include "solvers.h"
#include <queue>
#include "dist.h"
#include <cub/cub.cuh>
#include <climits>

namespace amgxb {

void spgemm_csr(int m, const DevBuf<int> &arp, const DevBuf<int> &aci, const DevVec &ava, const DevBuf<int> &brp, const DevBuf<int> &bci, const DevVec &bva,
                DevBuf<int> &crp, DevBuf<int> &cci, DevVec &cva, int *c_nnz, cudaStream_t s);

namespace {

constexpr int COARSE = -1, FINE = -2, STRONG_FINE = -3, UNASSIGNED = -4;
typedef unsigned char u8;
typedef long long i64;

inline int grid_for(i64 n) { return (int)std::max<i64>(1, std::min<i64>((n + 255) / 256, B200_SMS * 16)); }
#define ROW_LOOP(i, n) for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += gridDim.x * blockDim.x)

__host__ __device__ inline float cla_hash(int i)   // ourHash, strength_base.cu:41-55
{
    unsigned a = (unsigned)i;
    a = (a + 0x7ed55d16u) + (a << 12);
    a = (a ^ 0xc761c23cu) + (a >> 19);
    a = (a + 0x165667b1u) + (a << 5);
    a = (a ^ 0xd3a2646cu) + (a << 9);
    a = (a + 0xfd7046c5u) + (a << 3);
    a = (a ^ 0xb55a4f09u) + (a >> 16);
    return (float)(a ^ 0x4a51e590u) / (float)UINT_MAX;
}

// ---------------------------------------------------------------------------------------------------------------
// strength of connection
// ---------------------------------------------------------------------------------------------------------------
__global__ void strength_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const double *__restrict__ va, double alpha, double max_row_sum,
                                int compute_row_sum, u8 *__restrict__ s_con, int *cnt)
{
    ROW_LOOP(row, n) {
        double diag = 0, minv = 0, maxv = 0, sum = 0, dsum = 0;
        const int r0 = rp[row], r1 = rp[row + 1];
        for (int j = r0; j < r1; j++) {
            const double v = va[j];
            if (ci[j] == row) diag = v;
            else { minv = fmin(minv, v); maxv = fmax(maxv, v); }
            sum += v;
            if (ci[j] == row && v != 0) dsum = v;
        }
        const double row_sum = compute_row_sum ? fabs(sum / dsum) : -1.0;
        const double thr = ((diag < 0) ? maxv : minv) * alpha;
        for (int j = r0; j < r1; j++) {
            bool strong = false;
            if (!(compute_row_sum && row_sum > max_row_sum)) strong = ci[j] != row && ((diag < 0) ? va[j] > thr : va[j] < thr);
            s_con[j] = strong;
            if (strong && ci[j] < n) atomicAdd(&cnt[ci[j]], 1);
        }
    }
}
__global__ void pattern_count_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, int *cnt)   // computeWeightsKernel
{
    ROW_LOOP(i, n)
        for (int j = rp[i]; j < rp[i + 1]; j++)
            if (ci[j] != i) atomicAdd(&cnt[ci[j]], 1);
}
__global__ void weights_kernel(int n, const int *__restrict__ cnt, float *w) { ROW_LOOP(i, n) w[i] = (float)cnt[i] + cla_hash(i); }

// ---------------------------------------------------------------------------------------------------------------
// PMIS
// ---------------------------------------------------------------------------------------------------------------
__global__ void pmis_init_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const u8 *__restrict__ s_con, float *w, int *cf, int init)
{
    ROW_LOOP(i, n) {
        const int r0 = rp[i], numj = rp[i + 1] - r0;
        int c;
        if (numj == 0) c = FINE;
        else if (numj == 1 && ci[r0] == i) c = FINE;
        else if (w[i] < 1) c = FINE;
        else c = UNASSIGNED;
        bool isolated = true;
        for (int j = r0; j < r0 + numj; j++)
            if (!s_con || s_con[j]) { isolated = false; break; }
        if (isolated) { c = (init == 3) ? COARSE : STRONG_FINE; w[i] = 0.f; }
        cf[i] = c;
    }
}
// initialMarkingCfInitKernel (pmis.cu:316-360): cf already holds a C/F splitting (HMIS: the Ruge-Stueben first pass); its F points
// are reconsidered, its C points and strong-F points stand
__global__ void pmis_init_from_cf_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, float *w, int *cf, int *mark)
{
    ROW_LOOP(i, n) {
        const int r0 = rp[i], numj = rp[i + 1] - r0, in = cf[i];
        if (numj == 0) cf[i] = FINE;
        else if (numj == 1 && ci[r0] == i) cf[i] = FINE;
        else if (w[i] < 1) cf[i] = FINE;
        else if (in == STRONG_FINE) w[i] = 0.f;
        else if (in == FINE) { cf[i] = UNASSIGNED; mark[i] = 1; }
    }
}
__global__ void pmis_mark_coarse_kernel(int n, const float *__restrict__ w, const int *__restrict__ cf_in, int *cf_out, int *mark)
{
    ROW_LOOP(i, n) {
        const int in = cf_in[i], un = (in == UNASSIGNED);
        mark[i] = un;
        cf_out[i] = (w[i] > 1.f) ? (un ? COARSE : in) : in;
    }
}
__global__ void pmis_unmark_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const u8 *__restrict__ s_con, const float *__restrict__ w,
                                   int *cf_out, const int *__restrict__ mark)
{
    ROW_LOOP(i, n) {
        if (mark[i] <= 0) continue;
        const float wr = w[i];
        for (int j = rp[i]; j < rp[i + 1]; j++) {
            if (s_con && !s_con[j]) continue;
            const int jc = ci[j];
            if (jc >= n) continue;
            const float wc = w[jc];
            if (mark[jc] && wc > 1.0f) {
                if (wr > wc) cf_out[jc] = UNASSIGNED;        // every write stores UNASSIGNED: the interleaving does not matter
                else if (wc > wr) cf_out[i] = UNASSIGNED;
            }
        }
    }
}
__global__ void pmis_mark_fine_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const u8 *__restrict__ s_con, const int *__restrict__ cf_in,
                                      int *cf_out, int *num_unassigned)
{
    int un = 0;
    ROW_LOOP(i, n) {
        const int in = cf_in[i];
        bool fine = false;
        if (in == UNASSIGNED)
            for (int j = rp[i]; !fine && j < rp[i + 1]; j++) {
                if (s_con && !s_con[j]) continue;
                if (ci[j] < n) fine = cf_in[ci[j]] == COARSE;
            }
        const int out = fine ? FINE : in;
        cf_out[i] = out;
        un += (out == UNASSIGNED);
    }
    un = __reduce_add_sync(0xffffffffu, un);
    if ((threadIdx.x & 31) == 0 && un) atomicAdd(num_unassigned, un);
}

void pmis(int n, const int *rp, const int *ci, const u8 *s_con, float *w, int *cf, int init, cudaStream_t s)
{
    if (n == 0) return;
    DevBuf<int> scratch, mark, cnt;
    scratch.resize(n);
    mark.resize(n);
    mark.zero(s);
    cnt.resize(1);
    const int g = grid_for(n);
    if (init == 1) pmis_init_from_cf_kernel<<<g, 256, 0, s>>>(n, rp, ci, w, cf, mark.ptr());
    else pmis_init_kernel<<<g, 256, 0, s>>>(n, rp, ci, s_con, w, cf, init);
    count_launch();
    int iter = 0, num_unassigned;
    do {
        if (iter || !init) {
            pmis_mark_coarse_kernel<<<g, 256, 0, s>>>(n, w, cf, scratch.ptr(), mark.ptr());
            pmis_unmark_kernel<<<g, 256, 0, s>>>(n, rp, ci, s_con, w, scratch.ptr(), mark.ptr());
            count_launch(2);
        } else {
            AMGXB_CUDA_CHECK(cudaMemcpyAsync(scratch.ptr(), cf, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, s));
        }
        cnt.zero(s);
        pmis_mark_fine_kernel<<<g, 256, 0, s>>>(n, rp, ci, s_con, scratch.ptr(), cf, cnt.ptr());
        count_launch();
        AMGXB_LAUNCH_CHECK();
        num_unassigned = cnt.to_host(s)[0];
        iter++;
        if (iter > 10000) fatal(AMGX_RC_INTERNAL, "PMIS did not terminate");
    } while (num_unassigned != 0);
}

// ---------------------------------------------------------------------------------------------------------------
// HMIS = first pass of Ruge-Stueben coarsening, then PMIS on what it left fine.  The reference runs the first pass on the HOST
// (RS_Selector<host>, src/classical/selectors/rs.cu:36-262: "it's a sequential algorithm", :275) after copying matrix, strength
// flags and maps back (hmis.cu:58-88); so does this engine.  The reference keeps (measure, row) pairs in a std::set and always takes
// the largest measure, smallest row among equals; here a priority queue with lazy deletion gives the same sequence.
// ---------------------------------------------------------------------------------------------------------------
static void rs_first_pass_host(int n, const std::vector<int> &rp, const std::vector<int> &ci, const std::vector<u8> *s_con, std::vector<int> &cf)
{
    auto strong = [&](int k) { return (!s_con || (*s_con)[k]) && ci[k] < n; };
    std::vector<int> stp((size_t)n + 1, 0), stc((size_t)std::max(rp[n], 1));
    for (int k = 0; k < rp[n]; k++) if (strong(k)) stp[ci[k] + 1]++;
    for (int i = 0; i < n; i++) stp[i + 1] += stp[i];
    {
        std::vector<int> fill(stp.begin(), stp.end() - 1);
        for (int i = 0; i < n; i++) for (int k = rp[i]; k < rp[i + 1]; k++) if (strong(k)) stc[fill[ci[k]]++] = i;
    }
    std::vector<int> iw(n);
    std::vector<char> in_set(n, 0);
    struct Ent { int w, i; };
    struct Later { bool operator()(const Ent &a, const Ent &b) const { return a.w < b.w || (a.w == b.w && a.i > b.i); } };   // top(): max w, min i
    std::priority_queue<Ent, std::vector<Ent>, Later> pq;
    auto erase = [&](int i) { in_set[i] = 0; };
    auto insert = [&](int i) { in_set[i] = 1; pq.push(Ent{iw[i], i}); };
    for (int i = 0; i < n; i++) iw[i] = stp[i + 1] - stp[i];
    cf.assign((size_t)std::max(n, 1), 0);
    int num_left = 0;
    for (int j = 0; j < n; j++) {
        bool isolated = true;
        for (int k = rp[j]; k < rp[j + 1] && isolated; k++) isolated = !strong(k);
        if (isolated) { cf[j] = STRONG_FINE; iw[j] = 0; }
        else { cf[j] = UNASSIGNED; num_left++; }
    }
    for (int j = 0; j < n; j++) {
        if (cf[j] == STRONG_FINE) continue;
        if (iw[j] > 0) { insert(j); continue; }
        cf[j] = FINE;
        for (int k = rp[j]; k < rp[j + 1]; k++) {
            if (!strong(k)) continue;
            const int nb = ci[k];
            if (cf[nb] == STRONG_FINE) continue;
            if (nb < j) { if (iw[nb] > 0) erase(nb); ++iw[nb]; insert(nb); }
            else ++iw[nb];
        }
        --num_left;
    }
    auto bump_unassigned_neighbours = [&](int row) {
        for (int k = rp[row]; k < rp[row + 1]; k++) {
            if (!strong(k)) continue;
            const int d2 = ci[k];
            if (cf[d2] == UNASSIGNED) { erase(d2); ++iw[d2]; insert(d2); }
        }
    };
    while (num_left > 0) {
        int index = -1;
        while (!pq.empty()) {
            const Ent t = pq.top();
            pq.pop();
            if (in_set[t.i] && iw[t.i] == t.w) { index = t.i; break; }
        }
        if (index < 0) break;
        cf[index] = COARSE;
        iw[index] = 0;
        --num_left;
        erase(index);
        for (int j = stp[index]; j < stp[index + 1]; j++) {
            const int nb = stc[j];
            if (cf[nb] != UNASSIGNED) continue;
            cf[nb] = FINE;
            erase(nb);
            --num_left;
            bump_unassigned_neighbours(nb);
        }
        for (int j = rp[index]; j < rp[index + 1]; j++) {
            if (!strong(j)) continue;
            const int nb = ci[j];
            if (cf[nb] != UNASSIGNED) continue;
            erase(nb);
            const int wgt = --iw[nb];
            if (wgt > 0) { insert(nb); continue; }
            cf[nb] = FINE;
            --num_left;
            bump_unassigned_neighbours(nb);
        }
    }
}

void hmis(int n, const int *rp, const int *ci, int nnz, const u8 *s_con, float *w, int *cf, cudaStream_t s)
{
    if (n == 0) return;
    std::vector<int> hrp((size_t)n + 1), hci((size_t)std::max(nnz, 1)), hcf;
    std::vector<u8> hs;
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(hrp.data(), rp, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost, s));
    if (nnz) AMGXB_CUDA_CHECK(cudaMemcpyAsync(hci.data(), ci, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost, s));
    if (s_con) {
        hs.resize((size_t)std::max(nnz, 1));
        if (nnz) AMGXB_CUDA_CHECK(cudaMemcpyAsync(hs.data(), s_con, (size_t)nnz, cudaMemcpyDeviceToHost, s));
    }
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    rs_first_pass_host(n, hrp, hci, s_con ? &hs : nullptr, hcf);
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(cf, hcf.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    pmis(n, rp, ci, s_con, w, cf, 1, s);
}

// ---------------------------------------------------------------------------------------------------------------
// scans / renumbering
// ---------------------------------------------------------------------------------------------------------------
template <class T> void exclusive_scan(const T *in, T *out, size_t count, cudaStream_t s)
{
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, count, s);
    DevBytes tmp;
    tmp.resize(tb);
    cub::DeviceScan::ExclusiveSum(tmp.p, tb, in, out, count, s);
    count_launch();
}
__global__ void flag_coarse_kernel(int n, const int *__restrict__ cf, int *flag) { ROW_LOOP(i, n) flag[i] = (cf[i] == COARSE); if (blockIdx.x == 0 && threadIdx.x == 0) flag[n] = 0; }
__global__ void assign_coarse_kernel(int n, int *cf, const int *__restrict__ scan) { ROW_LOOP(i, n) if (cf[i] == COARSE) cf[i] = scan[i]; }

int renumber_coarse(int n, int *cf, cudaStream_t s)   // renumberAndCountCoarsePoints
{
    if (n == 0) return 0;
    DevBuf<int> flag, scan;
    flag.resize((size_t)n + 1);
    scan.resize((size_t)n + 1);
    flag_coarse_kernel<<<grid_for(n), 256, 0, s>>>(n, cf, flag.ptr());
    exclusive_scan(flag.ptr(), scan.ptr(), (size_t)n + 1, s);
    assign_coarse_kernel<<<grid_for(n), 256, 0, s>>>(n, cf, scan.ptr());
    count_launch(2);
    int nc = 0;
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(&nc, scan.ptr() + n, sizeof(int), cudaMemcpyDeviceToHost, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    return nc;
}

// ---------------------------------------------------------------------------------------------------------------
// sorted coarse sets (thread per row; segments live in global memory)
// ---------------------------------------------------------------------------------------------------------------
__device__ inline int insert_sorted(int *a, int m, int key)
{
    int lo = 0, hi = m;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    if (lo < m && a[lo] == key) return m;
    for (int k = m; k > lo; k--) a[k] = a[k - 1];
    a[lo] = key;
    return m + 1;
}
__device__ inline int find_sorted(const int *a, int m, int key)
{
    int lo = 0, hi = m;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return (lo < m && a[lo] == key) ? lo : -1;
}
// which rows own a distance-two set: mode 0 = coarse rows (S2), mode 1 = FINE rows (D2; coarse rows get 1 slot)
__device__ inline int chat_role(int cfv, int mode)
{
    if (mode == 0) return cfv >= 0 ? 2 : 0;
    if (cfv >= 0) return 1;
    return cfv == STRONG_FINE ? 0 : 2;
}
__global__ void chat_upper_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const u8 *__restrict__ s_con, const int *__restrict__ cf,
                                  int mode, i64 *ub)
{
    ROW_LOOP(i, n) {
        const int role = chat_role(cf[i], mode);
        i64 count = 0;
        if (role == 1) count = 1;
        else if (role == 2)
            for (int j = rp[i]; j < rp[i + 1]; j++) {
                const int c = ci[j];
                if (c == i || !s_con[j]) continue;
                const int cc = cf[c];
                if (cc == FINE) {
                    for (int jj = rp[c]; jj < rp[c + 1]; jj++) {
                        const int c2 = ci[jj];
                        if (c2 != c && s_con[jj]) { const int c3 = cf[c2]; if (c3 != FINE && c3 != STRONG_FINE) count++; }
                    }
                } else if (cc != STRONG_FINE) count++;
            }
        ub[i] = count;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ub[n] = 0;
}
__global__ void chat_fill_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const u8 *__restrict__ s_con, const int *__restrict__ cf,
                                 int mode, const i64 *__restrict__ off, int *chat, int *len)
{
    ROW_LOOP(i, n) {
        const int role = chat_role(cf[i], mode);
        int *out = chat + off[i];
        int m = 0;
        if (role == 1) { out[0] = i; m = 1; }
        else if (role == 2)
            for (int j = rp[i]; j < rp[i + 1]; j++) {
                const int c = ci[j];
                if (c == i || !s_con[j]) continue;
                const int cc = cf[c];
                if (cc == FINE) {
                    for (int jj = rp[c]; jj < rp[c + 1]; jj++) {
                        const int c2 = ci[jj];
                        if (c2 != c && s_con[jj]) { const int c3 = cf[c2]; if (c3 != FINE && c3 != STRONG_FINE) m = insert_sorted(out, m, c2); }
                    }
                } else if (cc != STRONG_FINE) m = insert_sorted(out, m, c);
            }
        len[i] = m;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) len[n] = 0;
}

// ---------------------------------------------------------------------------------------------------------------
// The reference's row order.  A row of P leaves the reference in the slot order of its Hash_set / Hash_map
// (include/hash_containers_detail.inl): 128 shared-memory slots, slot = ((key ^ c[f]) + c[4+f]) & 127 for the first of
// four hash functions whose slot is free or already holds the key; a warp inserts up to 32 keys at once and when
// several lanes race for one empty slot the lowest lane wins (this rule reproduces every row of the reference dumps);
// keys that lose all four rounds go to a global-memory table (same functions, mask gmem_size-1) stored after the
// shared-memory part.  The max-elements truncation keeps the FIRST of several equal weights, so this order decides
// which coarse points survive on symmetric stencils.  Sequential emulation, one thread per row: a "step" is one
// warp-wide insert, lanes visited in ascending order inside every hash round.
// ---------------------------------------------------------------------------------------------------------------
__constant__ unsigned c_cla_hash_keys[8] = {3499211612u, 581869302u, 3890346734u, 3586334585u, 545404204u, 4161255391u, 3922919429u, 949333985u};
constexpr int SS_SLOTS = 128, SS_OVF = 64;
struct SlotSet {
    int tab[SS_SLOTS];
    int ovf_slot[SS_OVF], ovf_key[SS_OVF];
    int n_ovf, gmem_mask;
    __device__ void clear(int gmem_size) { for (int s = 0; s < SS_SLOTS; s++) tab[s] = -1; n_ovf = 0; gmem_mask = gmem_size - 1; }
    __device__ static unsigned hash(int key, int f) { return ((unsigned)key ^ c_cla_hash_keys[f]) + c_cla_hash_keys[4 + f]; }
    __device__ void insert_step(int *keys)   // keys[0..31]: key of each lane or -1; destroyed
    {
        for (int f = 0; f < 4; f++) {
            bool any = false;
            for (int l = 0; l < 32; l++) {
                const int k = keys[l];
                if (k == -1) continue;
                const int s = (int)(hash(k, f) & (SS_SLOTS - 1));
                const int t = tab[s];
                if (t == -1) { tab[s] = k; keys[l] = -1; }
                else if (t == k) keys[l] = -1;
                else any = true;
            }
            if (!any) return;
        }
        for (int f = 0; f < 4; f++) {
            bool any = false;
            for (int l = 0; l < 32; l++) {
                const int k = keys[l];
                if (k == -1) continue;
                const int s = (int)(hash(k, f) & (unsigned)gmem_mask);
                int q = -1;
                for (int t = 0; t < n_ovf; t++) if (ovf_slot[t] == s) { q = t; break; }
                if (q < 0) {
                    if (n_ovf < SS_OVF) { ovf_slot[n_ovf] = s; ovf_key[n_ovf] = k; n_ovf++; }
                    keys[l] = -1;
                } else if (ovf_key[q] == k) keys[l] = -1;
                else any = true;
            }
            if (!any) return;
        }
    }
    __device__ int store(int *out) const   // shared-memory slots ascending, then global-memory slots ascending
    {
        int m = 0;
        for (int s = 0; s < SS_SLOTS; s++) if (tab[s] != -1) out[m++] = tab[s];
        int last = -1;
        for (int t = 0; t < n_ovf; t++) {       // selection by ascending slot (slots are distinct)
            int best = -1;
            for (int u = 0; u < n_ovf; u++) if (ovf_slot[u] > last && (best < 0 || ovf_slot[u] < ovf_slot[best])) best = u;
            out[m++] = ovf_key[best];
            last = ovf_slot[best];
        }
        return m;
    }
};
__device__ inline int find_linear(const int *a, int m, int key) { for (int k = 0; k < m; k++) if (a[k] == key) return k; return -1; }

// distance2::compute_c_hat_kernel (distance2.cu:848-1170): coarse set of the FINE rows in the reference's order.
// wide == 0: the 8-lanes-per-row variant (average nnz per row < 16), four rows of B in flight; wide == 1: 32 lanes per row of B.
__global__ void chat_fill_ref_order_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const u8 *__restrict__ s_con, const int *__restrict__ cf,
                                           int wide, const i64 *__restrict__ off, int *chat, int *len)
{
    ROW_LOOP(i, n) {
        const int cfi = cf[i];
        int *out = chat + off[i];
        if (cfi >= 0) { out[0] = i; len[i] = 1; continue; }
        if (cfi == STRONG_FINE) { len[i] = 0; continue; }
        SlotSet h;
        h.clear(512);                       // gmem_size of the reference on sm >= 7 (distance2.cu:1912-1914)
        int keys[32], fines[32];
        const int r1 = rp[i + 1];
        for (int c0 = rp[i]; c0 < r1; c0 += 32) {
            int nf = 0;
            for (int l = 0; l < 32; l++) {
                keys[l] = -1;
                const int k = c0 + l;
                if (k >= r1) continue;
                const int c = ci[k];
                if (c == i || !s_con[k]) continue;
                const int cc = cf[c];
                if (cc == FINE) fines[nf++] = c;
                else if (cc != STRONG_FINE) keys[l] = c;
            }
            h.insert_step(keys);
            if (!wide) {
                for (int g0 = 0; g0 < nf; g0 += 4)
                    for (int t = 0;; t++) {
                        bool any = false;
                        for (int l = 0; l < 32; l++) {
                            keys[l] = -1;
                            const int gi = l >> 3, m = l & 7;
                            if (g0 + gi >= nf) continue;
                            const int b = fines[g0 + gi], k = rp[b] + m + 8 * t;
                            if (k >= rp[b + 1]) continue;
                            any = true;
                            const int c = ci[k];
                            if (c != b && s_con[k]) { const int cc = cf[c]; if (cc != FINE && cc != STRONG_FINE) keys[l] = c; }
                        }
                        if (!any) break;
                        h.insert_step(keys);
                    }
            } else {
                for (int g = 0; g < nf; g++) {
                    const int b = fines[g], b1 = rp[b + 1];
                    for (int k0 = rp[b]; k0 < b1; k0 += 32) {
                        for (int l = 0; l < 32; l++) {
                            keys[l] = -1;
                            const int k = k0 + l;
                            if (k >= b1) continue;
                            const int c = ci[k];
                            if (c != b && s_con[k]) { const int cc = cf[c]; if (cc != FINE && cc != STRONG_FINE) keys[l] = c; }
                        }
                        h.insert_step(keys);
                    }
                }
            }
        }
        len[i] = h.store(out);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) len[n] = 0;
}

// S2: row of coarse point cf[i] <- its set mapped to coarse ids
__global__ void s2_compact_kernel(int n, const int *__restrict__ cf, const i64 *__restrict__ off, const int *__restrict__ chat, const int *__restrict__ len,
                                  const int *__restrict__ s2_rp, int *s2_ci)
{
    ROW_LOOP(i, n) {
        const int c = cf[i];
        if (c < 0) continue;
        const int *src = chat + off[i];
        int *dst = s2_ci + s2_rp[c];
        for (int k = 0; k < len[i]; k++) dst[k] = cf[src[k]];
    }
}
__global__ void s2_len_kernel(int n, const int *__restrict__ cf, const int *__restrict__ len, int *s2_len, int nc)
{
    ROW_LOOP(i, n) if (cf[i] >= 0) s2_len[cf[i]] = len[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) s2_len[nc] = 0;
}
__global__ void correct_cf_kernel(int n, int *cf, const int *__restrict__ scanned, const int *__restrict__ cf2)
{
    ROW_LOOP(i, n) if (cf[i] == COARSE) { const int c2 = cf2[scanned[i]]; cf[i] = (c2 == STRONG_FINE) ? COARSE : c2; }
}
__global__ void copy_segments_int_kernel(int n, const i64 *__restrict__ off, const int *__restrict__ len, const int *__restrict__ dst_rp, const int *__restrict__ src, int *dst)
{
    ROW_LOOP(i, n) { const int *a = src + off[i]; int *b = dst + dst_rp[i]; for (int k = 0; k < len[i]; k++) b[k] = a[k]; }
}
__global__ void copy_segments_val_kernel(int n, const i64 *__restrict__ off, const int *__restrict__ len, const int *__restrict__ dst_rp, const double *__restrict__ src, double *dst)
{
    ROW_LOOP(i, n) { const double *a = src + off[i]; double *b = dst + dst_rp[i]; for (int k = 0; k < len[i]; k++) b[k] = a[k]; }
}

// Aggressive_PMIS / Aggressive_HMIS selectors: the same two passes, with the plain selector of the same name on A and on S2
// (aggressive_pmis.cu:40,132; aggressive_hmis.cu:41,131 -- HMIS ignores the cf_map_init = 3 it is handed, hmis.cu:58-88)
void aggressive_pmis(int n, const int *rp, const int *ci, int nnz, const u8 *s_con, float *w, int *cf, bool use_hmis, cudaStream_t s)
{
    if (use_hmis) hmis(n, rp, ci, nnz, s_con, w, cf, s);
    else pmis(n, rp, ci, s_con, w, cf, 0, s);
    DevBuf<int> scanned;
    scanned.resize(n);
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(scanned.ptr(), cf, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    const int nc = renumber_coarse(n, scanned.ptr(), s);
    if (nc == 0) return;
    // S2 = distance-two strength graph among the coarse points (createS2)
    DevBuf<i64> ub, off;
    ub.resize((size_t)n + 1);
    off.resize((size_t)n + 1);
    const int g = grid_for(n);
    chat_upper_kernel<<<g, 256, 0, s>>>(n, rp, ci, s_con, scanned.ptr(), 0, ub.ptr());
    exclusive_scan(ub.ptr(), off.ptr(), (size_t)n + 1, s);
    i64 total = 0;
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(&total, off.ptr() + n, sizeof(i64), cudaMemcpyDeviceToHost, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    ub.release();
    DevBuf<int> chat, len, s2_len, s2_rp, s2_ci;
    chat.resize((size_t)std::max<i64>(total, 1));
    len.resize((size_t)n + 1);
    chat_fill_kernel<<<g, 256, 0, s>>>(n, rp, ci, s_con, scanned.ptr(), 0, off.ptr(), chat.ptr(), len.ptr());
    s2_len.resize((size_t)nc + 1);
    s2_rp.resize((size_t)nc + 1);
    s2_len_kernel<<<g, 256, 0, s>>>(n, scanned.ptr(), len.ptr(), s2_len.ptr(), nc);
    exclusive_scan(s2_len.ptr(), s2_rp.ptr(), (size_t)nc + 1, s);
    int s2_nnz = 0;
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(&s2_nnz, s2_rp.ptr() + nc, sizeof(int), cudaMemcpyDeviceToHost, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    s2_ci.resize((size_t)std::max(s2_nnz, 1));
    s2_compact_kernel<<<g, 256, 0, s>>>(n, scanned.ptr(), off.ptr(), chat.ptr(), len.ptr(), s2_rp.ptr(), s2_ci.ptr());
    count_launch(4);
    chat.release();
    off.release();
    // weights on S2 (Strength_All::computeWeights), PMIS with cf_map_init = 3, correctCfMap
    DevBuf<int> cnt, cf2;
    DevBuf<float> w2;
    cnt.resize(nc);
    cnt.zero(s);
    w2.resize(nc);
    cf2.resize(nc);
    pattern_count_kernel<<<grid_for(nc), 256, 0, s>>>(nc, s2_rp.ptr(), s2_ci.ptr(), cnt.ptr());
    weights_kernel<<<grid_for(nc), 256, 0, s>>>(nc, cnt.ptr(), w2.ptr());
    count_launch(2);
    if (use_hmis) hmis(nc, s2_rp.ptr(), s2_ci.ptr(), s2_nnz, nullptr, w2.ptr(), cf2.ptr(), s);
    else pmis(nc, s2_rp.ptr(), s2_ci.ptr(), nullptr, w2.ptr(), cf2.ptr(), 3, s);
    correct_cf_kernel<<<g, 256, 0, s>>>(n, cf, scanned.ptr(), cf2.ptr());
    count_launch();
    AMGXB_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------------------
// D2 interpolation
// ---------------------------------------------------------------------------------------------------------------
__device__ inline bool cla_sign(double x) { return x >= 0.0; }

__global__ void diag_value_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const double *__restrict__ va, double *d)
{
    ROW_LOOP(i, n) { double v = 0; for (int j = rp[i]; j < rp[i + 1]; j++) if (ci[j] == i) { v = va[j]; break; } d[i] = v; }
}

// P's structure already holds, for each row, its coarse set (reference order) as FINE-GRID ids in p_ci; the kernel computes the
// weights in place and finally rewrites the ids as coarse ids.
__global__ void d2_weights_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const double *__restrict__ va, const int *__restrict__ cf,
                                  const u8 *__restrict__ s_con, const double *__restrict__ diag, const int *__restrict__ p_rp, int *p_ci, double *p_va)
{
    ROW_LOOP(i, n) {
        const int p0 = p_rp[i], m = p_rp[i + 1] - p0;
        int *ch = p_ci + p0;
        double *val = p_va + p0;
        const int cfi = cf[i];
        if (cfi >= 0) { ch[0] = cfi; val[0] = 1.0; continue; }
        if (cfi == STRONG_FINE) continue;
        for (int k = 0; k < m; k++) val[k] = 0.0;
        const bool sign_i = cla_sign(diag[i]);
        double weak = 0.0;
        for (int j = rp[i]; j < rp[i + 1]; j++) {
            const int c = ci[j];
            const double a = va[j];
            const bool offd = (c != i);
            const bool strong = offd && s_con[j];
            const int p = find_linear(ch, m, c);
            if (p >= 0) val[p] += a;
            const int cfc = cf[c];
            if (offd && !strong && p < 0 && cfc != STRONG_FINE) weak += a;
            if (strong && cfc == FINE) {
                double bottom = 0.0;
                for (int jj = rp[c]; jj < rp[c + 1]; jj++) {
                    const int l = ci[jj];
                    const bool needed = (l == i) || find_linear(ch, m, l) >= 0;
                    const double b = needed ? va[jj] : 0.0;
                    if (sign_i != cla_sign(b)) bottom += b;
                }
                const double inner = (bottom != 0.0) ? a / bottom : a;
                const double dk = diag[c];
                double aki = 0.0;
                for (int jj = rp[c]; jj < rp[c + 1]; jj++) {
                    const int l = ci[jj];
                    double b = va[jj];
                    if (cla_sign(dk) == cla_sign(b)) b = 0.0;
                    if (l == i) aki = b;
                    const int q = find_linear(ch, m, l);
                    if (q >= 0) { const double t = b * inner; val[q] += t; }
                }
                const double t = aki * inner;
                weak += t;
            }
        }
        weak += diag[i];
        const double scale = -1.0 / weak;
        for (int k = 0; k < m; k++) { val[k] = scale * val[k]; ch[k] = cf[ch[k]]; }
    }
}

struct Csr {   // plain device CSR (fp64 values)
    int n = 0, nc = 0, nnz = 0;
    DevBuf<int> rp, ci;
    DevVec va;
};

void interp_d2(const Matrix &A, const int *cf, const u8 *s_con, int nc, Csr &P, cudaStream_t s)
{
    const int n = A.n;
    const int *rp = A.row_ptr.ptr(), *ci = A.col_idx.ptr();
    const double *va = A.values.as<double>();
    const int g = grid_for(n);
    DevBuf<i64> ub, off;
    ub.resize((size_t)n + 1);
    off.resize((size_t)n + 1);
    chat_upper_kernel<<<g, 256, 0, s>>>(n, rp, ci, s_con, cf, 1, ub.ptr());
    exclusive_scan(ub.ptr(), off.ptr(), (size_t)n + 1, s);
    i64 total = 0;
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(&total, off.ptr() + n, sizeof(i64), cudaMemcpyDeviceToHost, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    ub.release();
    DevBuf<int> chat, len;
    chat.resize((size_t)std::max<i64>(total, 1));
    len.resize((size_t)n + 1);
    const int wide = !(n > 0 && A.nnz / n < 16);   // kernel variant the reference picks (distance2.cu:1921-1945)
    chat_fill_ref_order_kernel<<<g, 128, 0, s>>>(n, rp, ci, s_con, cf, wide, off.ptr(), chat.ptr(), len.ptr());
    P.n = n;
    P.nc = nc;
    P.rp.resize((size_t)n + 1);
    exclusive_scan(len.ptr(), P.rp.ptr(), (size_t)n + 1, s);
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(&P.nnz, P.rp.ptr() + n, sizeof(int), cudaMemcpyDeviceToHost, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    P.ci.resize((size_t)std::max(P.nnz, 1));
    P.va.resize((size_t)std::max(P.nnz, 1), Prec::F64);
    copy_segments_int_kernel<<<g, 256, 0, s>>>(n, off.ptr(), len.ptr(), P.rp.ptr(), chat.ptr(), P.ci.ptr());
    chat.release();
    DevVec diag;
    diag.resize((size_t)std::max(n, 1), Prec::F64);
    diag_value_kernel<<<g, 256, 0, s>>>(n, rp, ci, va, diag.as<double>());
    d2_weights_kernel<<<g, 256, 0, s>>>(n, rp, ci, va, cf, s_con, diag.as<double>(), P.rp.ptr(), P.ci.ptr(), P.va.as<double>());
    count_launch(5);
    AMGXB_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------------------
// D1 interpolation -- the reference's DEFAULT interpolator (src/classical/interpolators/distance1.cu:353-867, device flow).  One thread
// per row, everything sequential inside a row, so the restatement is exact.  For a fine row i with strong coarse set C_i, strong fine
// set F_i (with or without a common coarse neighbour) and weak fine set W_i:
//     w_ij = -(a_ij + sum_{k in F_i} a_ik abar_kj / sum_{m in C_i} abar_km) / (a_ii + sum_{W_i} a_ik + sum_{k in F_i, empty denominator} a_ik)
// where abar keeps only entries whose sign is opposite to a_kk.  Two reference quirks are kept on purpose: a STRONG_FINE row gets one
// explicit (column 0, value 0) entry, and calculateBKernel's "return" inside its grid-stride loop (4096 x 64 launch) skips row i when
// some row i - m * 262144 is coarse (B and the D increment of that row stay 0).
// ---------------------------------------------------------------------------------------------------------------
constexpr int D1_STRONG_COARSE = 1, D1_WEAK_COARSE = 2, D1_STRONG_FINE = 4, D1_STRONG_FINE_NO_COMMON = 8, D1_WEAK_FINE = 16;
constexpr int D1_REF_THREADS = 262144;

__global__ void d1_count_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const int *__restrict__ cf, const u8 *__restrict__ s_con, int *nz)
{
    ROW_LOOP(i, n) {
        int cnt = 0;
        if (cf[i] == FINE) { for (int j = rp[i]; j < rp[i + 1]; j++) if (s_con[j] && cf[ci[j]] >= 0) cnt++; }
        else cnt = 1;
        nz[i] = cnt;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) nz[n] = 0;
}
__global__ void d1_mark_coarse_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const int *__restrict__ cf, u8 *mark)
{
    ROW_LOOP(i, n) for (int j = rp[i]; j < rp[i + 1]; j++) mark[j] = (ci[j] != i && cf[ci[j]] >= 0) ? 1 : 0;
}
__device__ inline bool d1_intersect(const int *__restrict__ ci, const u8 *__restrict__ mark, int b1, int e1, int b2, int e2)
{
    int i1 = b1, i2 = b2;
    if (b1 >= e1 || b2 >= e2) return false;
    for (;;) {
        const int c1 = ci[i1], c2 = ci[i2];
        if (c1 == c2) {
            if (mark[i1] && mark[i2]) return true;
            i1++; i2++;
            if (i1 >= e1 || i2 >= e2) return false;
        } else if (c1 > c2) { if (++i2 >= e2) return false; }
        else { if (++i1 >= e1) return false; }
    }
}
__global__ void d1_categorise_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const int *__restrict__ cf, const u8 *__restrict__ s_con,
                                     const u8 *__restrict__ mark, int *set)
{
    ROW_LOOP(i, n) {
        const bool coarse_row = cf[i] >= 0;
        for (int j = rp[i]; j < rp[i + 1]; j++) {
            int v = 0;
            const int jc = ci[j];
            if (!coarse_row && jc != i) {
                if (cf[jc] >= 0) v = s_con[j] ? D1_STRONG_COARSE : D1_WEAK_COARSE;
                else if (!s_con[j]) v = D1_WEAK_FINE;
                else v = d1_intersect(ci, mark, rp[i], rp[i + 1], rp[jc], rp[jc + 1]) ? D1_STRONG_FINE : D1_STRONG_FINE_NO_COMMON;
            }
            set[j] = v;
        }
    }
}
__global__ void d1_B_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const double *__restrict__ va, const double *__restrict__ diag,
                            const int *__restrict__ set, const int *__restrict__ cf, const int *__restrict__ prp, double *B, double *D)
{
    ROW_LOOP(i, n) {
        bool skipped = false;       // the reference thread owning row i left its loop at an earlier coarse row
        for (int e = i - D1_REF_THREADS; e >= 0 && !skipped; e -= D1_REF_THREADS) skipped = cf[e] >= 0;
        if (skipped) continue;
        if (cf[i] >= 0) { B[prp[i]] = 1; continue; }
        double dinc = 0;
        int flag = 0, first_j_loop = 0, local = 0;
        const double tol = 1e-10;
        for (int j = rp[i]; j < rp[i + 1]; j++) {
            if (!(set[j] & D1_STRONG_COARSE)) continue;
            const int jcol = ci[j];
            if (flag == 0) { first_j_loop = 1; flag = 1; } else first_j_loop = 0;
            double sum = 0.0;
            for (int k = rp[i]; k < rp[i + 1]; k++) {
                if (!((set[k] & D1_STRONG_FINE) || (set[k] & D1_STRONG_FINE_NO_COMMON))) continue;
                const int kcol = ci[k];
                const double a_ik = va[k];
                const int sgn = diag[kcol] < 0.0 ? -1 : 1;
                double top = 0.0, bottom = 0.0;
                for (int q = rp[kcol]; q < rp[kcol + 1]; q++)
                    if (ci[q] == jcol && sgn * va[q] < 0) top = a_ik * va[q];
                for (int m = rp[i]; m < rp[i + 1]; m++) {
                    if (!(set[m] & D1_STRONG_COARSE)) continue;
                    const int mcol = ci[m];
                    for (int q = rp[kcol]; q < rp[kcol + 1]; q++)
                        if (ci[q] == mcol && sgn * va[q] < 0) bottom += va[q];
                }
                if (fabs(bottom) < tol) { if (first_j_loop == 1) dinc += va[k]; }
                else sum += top / bottom;
            }
            B[prp[i] + local] = sum;
            local++;
        }
        D[i] = dinc;
    }
}
__global__ void d1_W_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const double *__restrict__ va, const double *__restrict__ diag,
                            const int *__restrict__ set, const int *__restrict__ cf, const int *__restrict__ prp, const double *__restrict__ B,
                            const double *__restrict__ D, int *pc, double *pv)
{
    ROW_LOOP(i, n) {
        if (cf[i] >= 0) { pv[prp[i]] = 1.0; pc[prp[i]] = cf[i]; continue; }
        double wf = 0;                                   // calculateDKernel: weak-fine entries join the diagonal
        for (int k = rp[i]; k < rp[i + 1]; k++) if (set[k] & D1_WEAK_FINE) wf += va[k];
        const double Di = D[i] + wf;
        int local = 0;
        for (int j = rp[i]; j < rp[i + 1]; j++) {
            if (!(set[j] & D1_STRONG_COARSE)) continue;
            const double bottom = (fabs(diag[i] + Di) < 1e-10) ? 1. : diag[i] + Di;
            pc[prp[i] + local] = cf[ci[j]];
            pv[prp[i] + local] = -1.0 / bottom * (va[j] + B[prp[i] + local]);
            local++;
        }
    }
}

void interp_d1(const Matrix &A, const int *cf, const u8 *s_con, int nc, Csr &P, cudaStream_t s)
{
    const int n = A.n;
    const int *rp = A.row_ptr.ptr(), *ci = A.col_idx.ptr();
    const double *va = A.values.as<double>();
    const int g = grid_for(n);
    DevBuf<int> nz, set;
    DevBuf<u8> mark;
    nz.resize((size_t)n + 1);
    P.n = n;
    P.nc = nc;
    P.rp.resize((size_t)n + 1);
    d1_count_kernel<<<g, 256, 0, s>>>(n, rp, ci, cf, s_con, nz.ptr());
    exclusive_scan(nz.ptr(), P.rp.ptr(), (size_t)n + 1, s);
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(&P.nnz, P.rp.ptr() + n, sizeof(int), cudaMemcpyDeviceToHost, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    P.ci.resize((size_t)std::max(P.nnz, 1));
    P.ci.zero(s);                                        // STRONG_FINE rows keep (column 0, value 0)
    P.va.resize((size_t)std::max(P.nnz, 1), Prec::F64);
    P.va.zero(s);
    mark.resize((size_t)std::max(A.nnz, 1));
    set.resize((size_t)std::max(A.nnz, 1));
    DevVec diag, B, D;
    diag.resize((size_t)std::max(n, 1), Prec::F64);
    B.resize((size_t)std::max(P.nnz, 1), Prec::F64);
    D.resize((size_t)std::max(n, 1), Prec::F64);
    B.zero(s);
    D.zero(s);
    diag_value_kernel<<<g, 256, 0, s>>>(n, rp, ci, va, diag.as<double>());
    d1_mark_coarse_kernel<<<g, 256, 0, s>>>(n, rp, ci, cf, mark.ptr());
    d1_categorise_kernel<<<g, 256, 0, s>>>(n, rp, ci, cf, s_con, mark.ptr(), set.ptr());
    d1_B_kernel<<<g, 256, 0, s>>>(n, rp, ci, va, diag.as<double>(), set.ptr(), cf, P.rp.ptr(), B.as<double>(), D.as<double>());
    d1_W_kernel<<<g, 256, 0, s>>>(n, rp, ci, va, diag.as<double>(), set.ptr(), cf, P.rp.ptr(), B.as<double>(), D.as<double>(), P.ci.ptr(), P.va.as<double>());
    count_launch(6);
    AMGXB_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------------------
// MULTIPASS interpolation
// ---------------------------------------------------------------------------------------------------------------
__global__ void mp_init_assigned_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const u8 *__restrict__ s_con, const int *__restrict__ cf,
                                        int *assigned, i64 *ub, int *counters)
{
    int un = 0, sf = 0;
    ROW_LOOP(i, n) {
        int a = -1;
        i64 u = 0;
        const int cfi = cf[i];
        if (cfi >= 0) { a = 0; u = 1; }
        else if (cfi == FINE) {
            int cc = 0;
            for (int j = rp[i]; j < rp[i + 1]; j++) if (ci[j] != i && s_con[j] && cf[ci[j]] >= 0) cc++;
            if (cc) { a = 1; u = cc; }
        }
        assigned[i] = a;
        ub[i] = u;
        un += (a < 0);
        sf += (cfi == STRONG_FINE);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ub[n] = 0;
    un = __reduce_add_sync(0xffffffffu, un);
    sf = __reduce_add_sync(0xffffffffu, sf);
    if ((threadIdx.x & 31) == 0) { if (un) atomicAdd(&counters[0], un); if (sf) atomicAdd(&counters[1], sf); }
}
// assigned[i] = pass if a strong neighbour was assigned in pass-1.  Concurrent writes store `pass`, which no reader
// of this launch tests for (readers compare with pass-1): the result is independent of the interleaving.
__global__ void mp_fill_assigned_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const u8 *__restrict__ s_con, int *assigned, int pass,
                                        int *num_unassigned)
{
    int un = 0;
    ROW_LOOP(i, n) {
        int a = ((volatile int *)assigned)[i];
        if (a == -1) {
            for (int j = rp[i]; j < rp[i + 1]; j++)
                if (ci[j] != i && s_con[j] && ((volatile int *)assigned)[ci[j]] == pass - 1) { a = pass; break; }
            if (a == pass) assigned[i] = pass;
        }
        un += (a < 0);
    }
    un = __reduce_add_sync(0xffffffffu, un);
    if ((threadIdx.x & 31) == 0 && un) atomicAdd(num_unassigned, un);
}
__global__ void mp_upper_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const u8 *__restrict__ s_con, const int *__restrict__ assigned,
                                i64 *ub, int pass)
{
    ROW_LOOP(i, n) {
        if (assigned[i] != pass) continue;
        i64 c = 0;
        for (int j = rp[i]; j < rp[i + 1]; j++) if (ci[j] != i && s_con[j] && assigned[ci[j]] == pass - 1) c += ub[ci[j]];
        ub[i] = c;
    }
}
__global__ void mp_first_pass_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const double *__restrict__ va, const int *__restrict__ cf,
                                     const u8 *__restrict__ s_con, const int *__restrict__ assigned, const double *__restrict__ diag,
                                     const i64 *__restrict__ off, int *cols, double *vals, int *len)
{
    ROW_LOOP(i, n) {
        int *pc = cols + off[i];
        double *pv = vals + off[i];
        const int a = assigned[i];
        int m = 0;
        if (a == 0) { pc[0] = cf[i]; pv[0] = 1.0; m = 1; }
        else if (a == 1) {
            double sum_N = 0.0, sum_C = 0.0;
            for (int j = rp[i]; j < rp[i + 1]; j++) {
                const int c = ci[j];
                if (c == i) continue;
                const double v = va[j];
                if (cf[c] != STRONG_FINE) sum_N += v;
                if (s_con[j] && assigned[c] == 0) { sum_C += v; pc[m] = cf[c]; pv[m] = v; m++; }
            }
            const double sd = sum_C * diag[i];
            const double div = (fabs(sd) == 0.0) ? 1.0 : sd;
            const double alfa = -sum_N / div;
            for (int k = 0; k < m; k++) pv[k] *= alfa;
        }
        len[i] = m;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) len[n] = 0;
}
__global__ void mp_pass_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const double *__restrict__ va, const int *__restrict__ cf,
                               const u8 *__restrict__ s_con, const int *__restrict__ assigned, const double *__restrict__ diag, const i64 *__restrict__ off,
                               int *cols, double *vals, int *len, int pass)
{
    ROW_LOOP(i, n) {
        if (assigned[i] != pass) continue;
        int *pc = cols + off[i];
        double *pv = vals + off[i];
        // multipass::compute_c_hat_kernel<8,...> (multipass.cu:762-905): union of the coarse sets of the strong neighbours
        // assigned in the previous pass, four neighbours in flight, 8 lanes each; keys = coarse ids; reference order
        SlotSet h;
        h.clear(2048);
        int keys[32], nb[32];
        const int r1 = rp[i + 1];
        for (int c0 = rp[i]; c0 < r1; c0 += 32) {
            int nn = 0;
            const int c1 = min(c0 + 32, r1);
            for (int j = c0; j < c1; j++) if (ci[j] != i && s_con[j] && assigned[ci[j]] == pass - 1) nb[nn++] = ci[j];
            for (int g0 = 0; g0 < nn; g0 += 4)
                for (int t = 0;; t++) {
                    bool any = false;
                    for (int l = 0; l < 32; l++) {
                        keys[l] = -1;
                        const int gi = l >> 3, mm = l & 7;
                        if (g0 + gi >= nn) continue;
                        const int b = nb[g0 + gi], idx = mm + 8 * t;
                        if (idx >= len[b]) continue;
                        any = true;
                        keys[l] = cols[off[b] + idx];
                    }
                    if (!any) break;
                    h.insert_step(keys);
                }
        }
        const int m = h.store(pc);
        for (int q = 0; q < m; q++) pv[q] = 0.0;
        double sum_N = 0.0, sum_C = 0.0;
        for (int j = rp[i]; j < rp[i + 1]; j++) {
            const int k = ci[j];
            if (k == i) continue;
            const double a = va[j];
            const bool sa = s_con[j] && assigned[k] == pass - 1;
            if (!sa) { if (cf[k] != STRONG_FINE) sum_N += a; continue; }
            const int *kc = cols + off[k];
            const double *kv = vals + off[k];
            for (int q = 0; q < len[k]; q++) {
                const double tmp = kv[q] * a;
                sum_C += tmp;
                sum_N += tmp;
                pv[find_linear(pc, m, kc[q])] += tmp;
            }
        }
        const double sd = sum_C * diag[i];
        const double div = (fabs(sd) == 0.0) ? 1.0 : sd;
        const double alfa = -sum_N / div;
        for (int q = 0; q < m; q++) pv[q] = alfa * pv[q];
        len[i] = m;
    }
}

void interp_multipass(const Matrix &A, const int *cf, const u8 *s_con, int nc, Csr &P, cudaStream_t s)
{
    const int n = A.n;
    const int *rp = A.row_ptr.ptr(), *ci = A.col_idx.ptr();
    const double *va = A.values.as<double>();
    const int g = grid_for(n);
    DevBuf<int> assigned, counters, len;
    DevBuf<i64> ub, off;
    assigned.resize(std::max(n, 1));
    ub.resize((size_t)n + 1);
    off.resize((size_t)n + 1);
    counters.resize(2);
    counters.zero(s);
    mp_init_assigned_kernel<<<g, 256, 0, s>>>(n, rp, ci, s_con, cf, assigned.ptr(), ub.ptr(), counters.ptr());
    count_launch();
    std::vector<int> h = counters.to_host(s);
    const int num_sf = h[1];
    int remaining = h[0] - num_sf;
    int pass = 2;
    while (remaining && pass < 10) {
        counters.zero(s);
        mp_fill_assigned_kernel<<<g, 256, 0, s>>>(n, rp, ci, s_con, assigned.ptr(), pass, counters.ptr());
        count_launch();
        remaining = counters.to_host(s)[0] - num_sf;
        pass++;
    }
    const int num_passes = pass;
    for (int p = 2; p < num_passes; p++) { mp_upper_kernel<<<g, 256, 0, s>>>(n, rp, ci, s_con, assigned.ptr(), ub.ptr(), p); count_launch(); }
    exclusive_scan(ub.ptr(), off.ptr(), (size_t)n + 1, s);
    i64 total = 0;
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(&total, off.ptr() + n, sizeof(i64), cudaMemcpyDeviceToHost, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    ub.release();
    DevBuf<int> cols;
    DevVec vals, diag;
    cols.resize((size_t)std::max<i64>(total, 1));
    vals.resize((size_t)std::max<i64>(total, 1), Prec::F64);
    len.resize((size_t)n + 1);
    diag.resize((size_t)std::max(n, 1), Prec::F64);
    diag_value_kernel<<<g, 256, 0, s>>>(n, rp, ci, va, diag.as<double>());
    mp_first_pass_kernel<<<g, 256, 0, s>>>(n, rp, ci, va, cf, s_con, assigned.ptr(), diag.as<double>(), off.ptr(), cols.ptr(), vals.as<double>(), len.ptr());
    count_launch(2);
    for (int p = 2; p < num_passes; p++) {
        mp_pass_kernel<<<g, 256, 0, s>>>(n, rp, ci, va, cf, s_con, assigned.ptr(), diag.as<double>(), off.ptr(), cols.ptr(), vals.as<double>(), len.ptr(), p);
        count_launch();
    }
    P.n = n;
    P.nc = nc;
    P.rp.resize((size_t)n + 1);
    exclusive_scan(len.ptr(), P.rp.ptr(), (size_t)n + 1, s);
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(&P.nnz, P.rp.ptr() + n, sizeof(int), cudaMemcpyDeviceToHost, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    P.ci.resize((size_t)std::max(P.nnz, 1));
    P.va.resize((size_t)std::max(P.nnz, 1), Prec::F64);
    copy_segments_int_kernel<<<g, 256, 0, s>>>(n, off.ptr(), len.ptr(), P.rp.ptr(), cols.ptr(), P.ci.ptr());
    copy_segments_val_kernel<<<g, 256, 0, s>>>(n, off.ptr(), len.ptr(), P.rp.ptr(), vals.as<double>(), P.va.as<double>());
    count_launch(2);
    AMGXB_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------------------
// truncation to the max_elmts largest |weights| per row + rescaling to the original row sum
// ---------------------------------------------------------------------------------------------------------------
__global__ void trunc_len_kernel(int n, const int *__restrict__ rp, int max_elmts, int *len)
{
    ROW_LOOP(i, n) len[i] = min(rp[i + 1] - rp[i], max_elmts);
    if (blockIdx.x == 0 && threadIdx.x == 0) len[n] = 0;
}
__global__ void truncate_kernel(int n, const int *__restrict__ rp, const int *__restrict__ ci, const double *__restrict__ va, const int *__restrict__ nrp,
                                int *nci, double *nva, int max_elmts)
{
    ROW_LOOP(i, n) {
        const int s = rp[i], e = rp[i + 1], rl = e - s;
        int oc[32];
        double ov[32];
        double orig = 0.0;
        for (int j = s; j < e; j++) orig += va[j];
        int m;
        if (rl <= max_elmts) {
            m = rl;
            for (int j = 0; j < m; j++) { oc[j] = ci[s + j]; ov[j] = va[s + j]; }
        } else {
            m = max_elmts;
            for (int j = 0; j < m; j++) { oc[j] = ci[s + j]; ov[j] = va[s + j]; }
            int nn = m;
            do {
                int newn = 0;
                for (int q = 1; q < nn; q++)
                    if (fabs(ov[q - 1]) < fabs(ov[q])) {
                        const double tv = ov[q - 1]; const int ti = oc[q - 1];
                        ov[q - 1] = ov[q]; oc[q - 1] = oc[q]; ov[q] = tv; oc[q] = ti;
                        newn = q;
                    }
                nn = newn;
            } while (nn > 0);
            for (int j = s + m; j < e; j++) {
                const double v = va[j];
                for (int q = 0; q < m; q++)
                    if (fabs(v) > fabs(ov[q])) {
                        for (int k = m - 1; k > q; k--) { ov[k] = ov[k - 1]; oc[k] = oc[k - 1]; }
                        ov[q] = v; oc[q] = ci[j];
                        break;
                    }
            }
        }
        double nsum = 0.0;
        for (int j = 0; j < m; j++) nsum += ov[j];
        const double mult = (fabs(nsum) == 0.0) ? 1.0 : orig / nsum;
        int *dc = nci + nrp[i];
        double *dv = nva + nrp[i];
        for (int j = 0; j < m; j++) { dc[j] = oc[j]; dv[j] = ov[j] * mult; }
    }
}
void truncate_max_elements(Csr &P, int max_elmts, cudaStream_t s)
{
    if (max_elmts > 32) fatal(AMGX_RC_BAD_PARAMETERS, "Matrix truncation to > 32 elements not supported");   // truncate.cu:786-789
    const int n = P.n;
    if (n == 0) return;
    DevBuf<int> len, nrp, nci;
    DevVec nva;
    len.resize((size_t)n + 1);
    nrp.resize((size_t)n + 1);
    trunc_len_kernel<<<grid_for(n), 256, 0, s>>>(n, P.rp.ptr(), max_elmts, len.ptr());
    exclusive_scan(len.ptr(), nrp.ptr(), (size_t)n + 1, s);
    int nnz = 0;
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(&nnz, nrp.ptr() + n, sizeof(int), cudaMemcpyDeviceToHost, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    nci.resize((size_t)std::max(nnz, 1));
    nva.resize((size_t)std::max(nnz, 1), Prec::F64);
    truncate_kernel<<<grid_for(n), 256, 0, s>>>(n, P.rp.ptr(), P.ci.ptr(), P.va.as<double>(), nrp.ptr(), nci.ptr(), nva.as<double>(), max_elmts);
    count_launch(2);
    AMGXB_LAUNCH_CHECK();
    P.rp.swap(nrp);
    P.ci.swap(nci);
    P.va.swap(nva);
    P.nnz = nnz;
}

// ---------------------------------------------------------------------------------------------------------------
// transpose (stable: rows of R list the fine rows in ascending order)
// ---------------------------------------------------------------------------------------------------------------
__global__ void expand_rows_kernel(int n, const int *__restrict__ rp, int *row_of) { ROW_LOOP(i, n) for (int k = rp[i]; k < rp[i + 1]; k++) row_of[k] = i; }
__global__ void iota_kernel(int n, int *v) { ROW_LOOP(i, n) v[i] = i; }
__global__ void gather_transpose_kernel(int nnz, const int *__restrict__ perm, const int *__restrict__ row_of, const double *__restrict__ va, int *rci, double *rva)
{
    ROW_LOOP(q, nnz) { const int e = perm[q]; rci[q] = row_of[e]; rva[q] = va[e]; }
}
__global__ void offsets_from_sorted_kernel(int nnz, const int *__restrict__ keys, int n_keys, int *offsets)
{
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p <= nnz; p += gridDim.x * blockDim.x) {
        const int prev = (p == 0) ? -1 : keys[p - 1];
        const int cur = (p == nnz) ? n_keys : keys[p];
        for (int I = prev + 1; I <= cur; I++) offsets[I] = p;
    }
}
void transpose_csr(const Csr &P, Csr &R, cudaStream_t s)
{
    R.n = P.nc;
    R.nc = P.n;
    R.nnz = P.nnz;
    R.rp.resize((size_t)R.n + 1);
    R.ci.resize((size_t)std::max(P.nnz, 1));
    R.va.resize((size_t)std::max(P.nnz, 1), Prec::F64);
    if (P.nnz == 0) { R.rp.zero(s); return; }
    DevBuf<int> row_of, idx, keys_out, perm;
    row_of.resize(P.nnz);
    idx.resize(P.nnz);
    keys_out.resize(P.nnz);
    perm.resize(P.nnz);
    expand_rows_kernel<<<grid_for(P.n), 256, 0, s>>>(P.n, P.rp.ptr(), row_of.ptr());
    iota_kernel<<<grid_for(P.nnz), 256, 0, s>>>(P.nnz, idx.ptr());
    int bits = 1;
    while ((1ll << bits) < (i64)P.nc + 1) bits++;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, P.ci.ptr(), keys_out.ptr(), idx.ptr(), perm.ptr(), P.nnz, 0, bits, s);
    DevBytes tmp;
    tmp.resize(tb);
    cub::DeviceRadixSort::SortPairs(tmp.p, tb, P.ci.ptr(), keys_out.ptr(), idx.ptr(), perm.ptr(), P.nnz, 0, bits, s);
    gather_transpose_kernel<<<grid_for(P.nnz), 256, 0, s>>>(P.nnz, perm.ptr(), row_of.ptr(), P.va.as<double>(), R.ci.ptr(), R.va.as<double>());
    offsets_from_sorted_kernel<<<grid_for((i64)P.nnz + 1), 256, 0, s>>>(P.nnz, keys_out.ptr(), R.n, R.rp.ptr());
    count_launch(5);
    AMGXB_LAUNCH_CHECK();
}

std::unique_ptr<Matrix> to_matrix(Csr &C, const Matrix &like, cudaStream_t s)
{
    std::unique_ptr<Matrix> M(new Matrix);
    M->rsc = like.rsc;
    M->mode = like.mode;
    M->mat_prec = like.mat_prec;
    M->vec_prec = like.vec_prec;
    M->n = C.n;
    M->n_cols = C.nc;
    M->nnz = C.nnz;
    M->row_ptr.swap(C.rp);
    M->col_idx.swap(C.ci);
    M->values.swap(C.va);
    M->values.n = (size_t)C.nnz;
    (void)s;
    return M;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// level construction
// ---------------------------------------------------------------------------------------------------------------
struct ClassicalParams {
    double strength_threshold, max_row_sum;
    int max_elmts, aggressive_levels;
    bool d2, d1 = false, aggressive_multipass, hmis = false, aggressive_hmis = false;
};

// createCoarseVertices: strength + C/F splitting; returns the number of coarse points, cf_map renumbered
static int classical_select(const Matrix &A, const ClassicalParams &prm, int level, DevBuf<u8> &s_con, DevBuf<int> &cf, cudaStream_t s)
{
    const int n = A.n;
    s_con.resize((size_t)std::max(A.nnz, 1));
    cf.resize((size_t)std::max(n, 1));
    DevBuf<int> cnt;
    DevBuf<float> w;
    cnt.resize(std::max(n, 1));
    cnt.zero(s);
    w.resize(std::max(n, 1));
    const int compute_row_sum = (prm.max_row_sum < 1.0 && n > 0) ? 1 : 0;
    strength_kernel<<<grid_for(n), 256, 0, s>>>(n, A.row_ptr.ptr(), A.col_idx.ptr(), A.values.as<double>(), prm.strength_threshold, prm.max_row_sum,
                                                  compute_row_sum, s_con.ptr(), cnt.ptr());
    weights_kernel<<<grid_for(n), 256, 0, s>>>(n, cnt.ptr(), w.ptr());
    count_launch(2);
    AMGXB_LAUNCH_CHECK();
    if (level < prm.aggressive_levels) aggressive_pmis(n, A.row_ptr.ptr(), A.col_idx.ptr(), A.nnz, s_con.ptr(), w.ptr(), cf.ptr(), prm.aggressive_hmis, s);
    else if (prm.hmis) hmis(n, A.row_ptr.ptr(), A.col_idx.ptr(), A.nnz, s_con.ptr(), w.ptr(), cf.ptr(), s);
    else pmis(n, A.row_ptr.ptr(), A.col_idx.ptr(), s_con.ptr(), w.ptr(), cf.ptr(), 0, s);
    return renumber_coarse(n, cf.ptr(), s);
}

void AMGSolver::setup_classical()
{
    cudaStream_t s = stream();
    // Row-partitioned matrix: the hierarchy is built from the assembled global matrix, redundantly on every rank (identical to the
    // single-GPU hierarchy of the caller's global matrix), then the finest level alone is distributed -- see distribute_finest() below.
    std::unique_ptr<Matrix> assembled;
    std::vector<int> g_counts, g_offs;
    if (A_->bs() != 1) fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "Unsupported block size for strong connections");   // strength_base.cu:672-681
    if (A_->mat_prec != Prec::F64 || A_->vec_prec != Prec::F64) fatal(AMGX_RC_BAD_MODE, "classical AMG setup needs mode dDDI");
    if (A_->dist && cfg_->get_string("interpolator", scope_) == "D1")      // classical_amg_level.cu:268-273
        fatal(AMGX_RC_NOT_IMPLEMENTED, "D1 interpolation is not supported in distributed settings");
    if (A_->dist) {
        if (A_->has_ext_diag) fatal(AMGX_RC_NOT_IMPLEMENTED, "classical AMG on a partitioned matrix with an external diagonal");
        assembled = dist_gather_matrix(*A_, g_counts, g_offs, true);
        assembled->compute_diag_and_plan();
    }
    Matrix *top = assembled ? assembled.get() : A_;
    ClassicalParams prm;
    prm.strength_threshold = cfg_->get_double("strength_threshold", scope_);
    prm.max_row_sum = cfg_->get_double("max_row_sum", scope_);
    prm.max_elmts = cfg_->get_int("interp_max_elements", scope_);
    prm.aggressive_levels = cfg_->get_int("aggressive_levels", scope_);
    const std::string strength = cfg_->get_string("strength", scope_), interp = cfg_->get_string("interpolator", scope_);
    const std::string agg_sel = cfg_->get_string("aggressive_selector", scope_), agg_int = cfg_->get_string("aggressive_interpolator", scope_);
    if (strength != "AHAT") fatal(AMGX_RC_BAD_CONFIGURATION, "strength '" + strength + "' is not supported by this engine (AHAT)");
    if (selector_ != "PMIS" && selector_ != "HMIS")
        fatal(AMGX_RC_BAD_CONFIGURATION, "classical selector '" + selector_ + "' is not supported by this engine (PMIS, HMIS)");
    prm.hmis = (selector_ == "HMIS");
    if (interp != "D1" && interp != "D2" && interp != "MULTIPASS")
        fatal(AMGX_RC_BAD_CONFIGURATION, "interpolator '" + interp + "' is not supported by this engine (D1, D2, MULTIPASS)");
    if (prm.aggressive_levels > 0) {
        if (agg_sel != "DEFAULT" && agg_sel != "PMIS" && agg_sel != "HMIS")
            fatal(AMGX_RC_BAD_CONFIGURATION, "aggressive_selector '" + agg_sel + "' is not supported (DEFAULT, PMIS, HMIS)");
        prm.aggressive_hmis = (agg_sel == "HMIS") || (agg_sel == "DEFAULT" && prm.hmis);   // classical_amg_level.cu:132-147
        if (agg_int != "MULTIPASS") fatal(AMGX_RC_BAD_CONFIGURATION, "aggressive_interpolator '" + agg_int + "' is not supported (MULTIPASS)");
    }
    prm.d2 = (interp == "D2");
    prm.d1 = (interp == "D1");
    prm.aggressive_multipass = true;

    levels_.emplace_back(new AMGLevel);
    levels_[0]->A = top;
    levels_[0]->index = 0;
    int num_levels = 1;
    bool coarse_solver_exists = (bool)coarse_solver_;
    while (true) {     // AMG_Setup::setup level loop (src/amg.cu:201-418)
        AMGLevel &L = *levels_.back();
        Matrix &A = *L.A;
        A.level = num_levels - 1;
        const int rows = A.n;
        if (num_levels >= max_levels_ || rows <= min_coarse_rows_) {
            if (dense_lu_max_rows_ != 0 && rows > dense_lu_max_rows_) { coarse_solver_.reset(); coarse_solver_exists = false; }
            L.coarsest = true;
            if (!coarse_solver_exists) { L.smoother = make_smoother(); L.smoother->setup(num_levels == 1 ? *A_ : A, false); }
            break;
        }
        DevBuf<u8> s_con;
        const int lvl = num_levels - 1;
        // AMGX_solver_resetup with structure_reuse_levels: P and R of the previous setup stand, only the Galerkin product follows the new values
        const bool reused = (size_t)lvl < reuse_P_.size() && reuse_P_[lvl] && reuse_P_[lvl]->n == rows;
        if (!reused) { reuse_P_.resize(std::min(reuse_P_.size(), (size_t)lvl)); reuse_R_.resize(reuse_P_.size()); }   // the chain ends at the first rebuilt level
        int nc;
        if (reused) {
            L.cf_map.swap(reuse_cf_[lvl]);
            nc = reuse_n_coarse_[lvl];
        } else {
            nc = classical_select(A, prm, lvl, s_con, L.cf_map, s);
        }
        L.n_coarse = nc;
        bool built_next = false;
        if ((double)nc <= coarsen_threshold_ * (double)rows && nc != rows && nc >= min_coarse_rows_) {
            Csr P, R;
            if (reused) {
                L.P = std::move(reuse_P_[lvl]);
                L.R = std::move(reuse_R_[lvl]);
            } else {
                if (lvl < prm.aggressive_levels || (!prm.d2 && !prm.d1)) interp_multipass(A, L.cf_map.ptr(), s_con.ptr(), nc, P, s);
                else if (prm.d1) interp_d1(A, L.cf_map.ptr(), s_con.ptr(), nc, P, s);
                else interp_d2(A, L.cf_map.ptr(), s_con.ptr(), nc, P, s);
                s_con.release();
                if (prm.max_elmts > 0 && P.n > 0) truncate_max_elements(P, prm.max_elmts, s);
                transpose_csr(P, R, s);
            }
            // the operands of the Galerkin product: the carried-over matrices, or the fresh P / R (turned into matrices further down)
            const DevBuf<int> &p_rp = reused ? L.P->row_ptr : P.rp, &p_ci = reused ? L.P->col_idx : P.ci;
            const DevVec &p_va = reused ? L.P->values : P.va;
            const DevBuf<int> &r_rp = reused ? L.R->row_ptr : R.rp, &r_ci = reused ? L.R->col_idx : R.ci;
            const DevVec &r_va = reused ? L.R->values : R.va;
            // A_c = R (A P)
            Csr AP;
            AP.n = rows;
            AP.nc = nc;
            spgemm_csr(rows, A.row_ptr, A.col_idx, A.values, p_rp, p_ci, p_va, AP.rp, AP.ci, AP.va, &AP.nnz, s);
            std::unique_ptr<AMGLevel> next(new AMGLevel);
            next->owned_A.reset(new Matrix);
            Matrix &Ac = *next->owned_A;
            Ac.rsc = A.rsc;
            Ac.mode = A.mode;
            Ac.mat_prec = A.mat_prec;
            Ac.vec_prec = A.vec_prec;
            Ac.n = nc;
            Ac.n_cols = nc;
            spgemm_csr(nc, r_rp, r_ci, r_va, AP.rp, AP.ci, AP.va, Ac.row_ptr, Ac.col_idx, Ac.values, &Ac.nnz, s);
            Ac.values.n = (size_t)Ac.nnz;
            AP.rp.release(); AP.ci.release(); AP.va.b.release();
            Ac.compute_diag_and_plan();
            if (!reused) {
                L.P = to_matrix(P, A, s);
                L.R = to_matrix(R, A, s);
                csr_build_plan(*L.P, s);
                csr_build_plan(*L.R, s);
            }
            next->A = next->owned_A.get();
            next->index = num_levels;
            L.bc.resize((size_t)nc, A.vec_prec);
            L.xc.resize((size_t)nc, A.vec_prec);
            L.bc.zero(s);
            L.xc.zero(s);
            L.r.resize((size_t)A.n_cols, A.vec_prec);
            L.r.zero(s);
            levels_.push_back(std::move(next));
            built_next = true;
        } else {
            L.cf_map.release();
            L.n_coarse = 0;
            L.coarsest = true;
        }
        AMGLevel &Lcur = *levels_[num_levels - 1];
        // the smoother of the finest level works on the caller's (possibly partitioned) matrix, not on the assembled copy
        if (!Lcur.coarsest || !coarse_solver_exists) { Lcur.smoother = make_smoother(); Lcur.smoother->setup(num_levels == 1 ? *A_ : *Lcur.A, false); }
        if (!built_next) break;
        num_levels++;
    }
    if (assembled) distribute_finest(g_offs);
    if (coarse_solver_) coarse_solver_->setup(*levels_.back()->A, false);
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
}

namespace {
__global__ void invert_perm_offset_kernel(int n, const int *__restrict__ perm_old_to_new, int offset, int *global_row_of_local)
{
    ROW_LOOP(old, n) global_row_of_local[perm_old_to_new[old]] = offset + old;
}
__global__ void gather_rows_count_kernel(int n, const int *__restrict__ rows, const int *__restrict__ rp, int *len)
{
    ROW_LOOP(i, n) len[i] = rp[rows[i] + 1] - rp[rows[i]];
    if (blockIdx.x == 0 && threadIdx.x == 0) len[n] = 0;
}
__global__ void gather_rows_fill_kernel(int n, const int *__restrict__ rows, const int *__restrict__ rp, const int *__restrict__ ci, const double *__restrict__ va,
                                        const int *__restrict__ orp, int *oci, double *ova)
{
    ROW_LOOP(i, n) {
        const int src = rp[rows[i]], len = rp[rows[i] + 1] - src, dst = orp[i];
        for (int k = 0; k < len; k++) { oci[dst + k] = ci[src + k]; ova[dst + k] = va[src + k]; }
    }
}
}  // namespace

// Classical AMG on a row-partitioned matrix.  Levels >= 1 stay what the loop above built from the assembled matrix: whole, replicated,
// cycled redundantly by every rank (aggressive coarsening leaves them 6 % of the rows at 512^3; they are latency-, not bandwidth-bound).
// Level 0 -- where the bytes are -- is distributed: smoothing and residual run on the caller's partitioned matrix with the usual halo
// exchange; prolongation uses this rank's rows of P (in local row order; columns are global coarse ids, the coarse vector is whole on
// every rank, so no exchange); restriction applies the transpose of those rows -- this rank's contribution to every coarse residual --
// and one all-reduce over the coarse vector sums the contributions (the one collective the path adds; it replaces the reference's
// reverse halo exchange of R's halo rows, classical_amg_level.cu:590-644).  Same operators as the single-GPU hierarchy of the global
// matrix; the restricted residual differs from it only by the association of that sum.
void AMGSolver::distribute_finest(const std::vector<int> &g_offs)
{
    cudaStream_t s = stream();
    AMGLevel &L = *levels_[0];
    const Matrix &A = *A_;
    DistManager &m = *A.dist;
    L.A = A_;
    A_->level = 0;
    L.cf_map.release();
    if (L.coarsest) return;                   // a single level: nothing to transfer
    const int n = A.n, nc = L.P->n_cols;
    DevBuf<int> rows, len;
    rows.resize((size_t)std::max(n, 1));
    len.resize((size_t)n + 1);
    const int g = grid_for(n);
    invert_perm_offset_kernel<<<g, 256, 0, s>>>(n, m.perm_old_to_new.ptr(), g_offs[m.rank], rows.ptr());
    Csr Pl, Rl;
    Pl.n = n;
    Pl.nc = nc;
    Pl.rp.resize((size_t)n + 1);
    gather_rows_count_kernel<<<g, 256, 0, s>>>(n, rows.ptr(), L.P->row_ptr.ptr(), len.ptr());
    exclusive_scan(len.ptr(), Pl.rp.ptr(), (size_t)n + 1, s);
    AMGXB_CUDA_CHECK(cudaMemcpyAsync(&Pl.nnz, Pl.rp.ptr() + n, sizeof(int), cudaMemcpyDeviceToHost, s));
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
    Pl.ci.resize((size_t)std::max(Pl.nnz, 1));
    Pl.va.resize((size_t)std::max(Pl.nnz, 1), Prec::F64);
    gather_rows_fill_kernel<<<g, 256, 0, s>>>(n, rows.ptr(), L.P->row_ptr.ptr(), L.P->col_idx.ptr(), L.P->values.as<double>(), Pl.rp.ptr(), Pl.ci.ptr(),
                                              Pl.va.as<double>());
    count_launch(3);
    AMGXB_LAUNCH_CHECK();
    transpose_csr(Pl, Rl, s);
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));      // the assembled P and R are released next
    L.P = to_matrix(Pl, A, s);
    L.R = to_matrix(Rl, A, s);
    csr_build_plan(*L.P, s);
    csr_build_plan(*L.R, s);
    L.cla_reduce_over = A_;
    L.r.resize((size_t)A.n_cols, A.vec_prec);
    L.r.zero(s);
    AMGXB_CUDA_CHECK(cudaStreamSynchronize(s));
}

// rr = R r  (Classical_AMG_Level_Base::restrictResidual, classical_amg_level.cu:590-644)
void classical_restrict(AMGLevel &L, const DevVec &r, cudaStream_t s)
{
    CsrOpArgs g;
    g.x = r.ptr();
    g.y = L.bc.ptr();
    csr_op(*L.R, EPI_SPMV, g, s, 0);
    if (L.cla_reduce_over) dist_allreduce_vec(*L.cla_reduce_over, L.bc.ptr(), L.bc.prec, (size_t)L.R->n, s);
}
// xout = x + P e  (prolongateAndApplyCorrection: multiply(P, e, tmp); axpby(x, tmp, x, 1, 1); classical_amg_level.cu:851-913)
void classical_prolong_add(AMGLevel &L, const void *x, void *xout, cudaStream_t s)
{
    CsrOpArgs g;
    g.x = L.xc.ptr();
    g.b = x;
    g.y = xout;
    csr_op(*L.P, x ? EPI_ADD : EPI_SPMV, g, s, 0);
}

}  // namespace amgxb
