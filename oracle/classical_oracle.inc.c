/*
 * classical_oracle.inc.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (included by amg_oracle.c).
 *
 * CPU restatement of the classical (Ruge-Stueben) AMG setup producers and transfer operators of the
 * reference, the pieces behind BASELINE config 3 (FGMRES_CLASSICAL_AGGRESSIVE_PMIS.json):
 *   strength of connection AHAT + PMIS weights   src/classical/strength/strength_base.cu:185-330
 *   PMIS C/F splitting                           src/classical/selectors/pmis.cu:221-266, 370-466, 468-622
 *   aggressive PMIS (second pass on S2)          src/classical/selectors/aggressive_pmis.cu:22-150,
 *                                                src/classical/selectors/selector.cu:116-230, 427-580, 942-1004, 1057-1070
 *   distance-2 "extended+i" interpolation        src/classical/interpolators/distance2.cu:600-716, 1178-1362, 1562-1796
 *   multipass interpolation                      src/classical/interpolators/multipass.cu:94-147, 244-287, 1057-1206, 1538-1720
 *   truncation to interp_max_elements            src/truncate.cu:352-456, 78-92, 783-862
 *   R = P^T, A_c = R A P                         src/classical/classical_amg_level.cu:440-468, 501-586
 *   restriction / prolongation                   src/classical/classical_amg_level.cu:590-644, 851-913
 *
 * Row order of P: the reference's (hash-table slot order, emulated below -- it decides which of several EQUAL
 * weights the max-elements truncation keeps).  Where the reference leaves the order of floating-point
 * contributions to atomics and lane-partial sums, this restatement fixes ONE order -- the same one the CUDA
 * engine uses: sums run left to right in storage order, products are rounded before they are added (no FMA).
 * Consequences, stated in DESIGN.md: selection arrays (strong connections, C/F maps), the pattern AND the row
 * order of P on the finest level are comparable bit for bit with the reference; interpolation weights agree to
 * rounding.  On coarser levels the reference's A has hash-ordered columns (its SpGEMM), ours ascending ones, so
 * the emulated insertion sequence -- and with it the choice among equal weights -- can differ there.
 */
#include <limits.h>

#define CLA_COARSE (-1)
#define CLA_FINE (-2)
#define CLA_STRONG_FINE (-3)
#define CLA_UNASSIGNED (-4)

static float cla_hash(int i) /* ourHash, strength_base.cu:41-55 */
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

/* computeStrongConnectionsAndWeightsKernel: AHAT rule; weights = hash(row) + number of rows that
 * strongly depend on it.  The reference accumulates the float weight with atomicAdd in an
 * unspecified order; here it is (float)count + hash (one rounding). */
ORC_API void orc_cla_strength(int n, const int *rp, const int *ci, const double *va, double alpha, double max_row_sum, unsigned char *s_con,
                              float *weights)
{
    const int compute_row_sum = max_row_sum < 1.0 && n > 0;
    int *cnt = (int *)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
    for (int row = 0; row < n; row++) {
        double diag = 0, minv = 0, maxv = 0, sum = 0, dsum = 0;
        for (int j = rp[row]; j < rp[row + 1]; j++) {
            const double v = va[j];
            if (ci[j] == row) diag = v;
            else { if (v < minv) minv = v; if (v > maxv) maxv = v; }
            sum += v;                          /* weighted_row_sum: include/specific_spmv.h:99-140 */
            if (ci[j] == row && v != 0) dsum = v;
        }
        const double row_sum = compute_row_sum ? fabs(sum / dsum) : -1.0;
        const double thr = ((diag < 0) ? maxv : minv) * alpha;
        for (int j = rp[row]; j < rp[row + 1]; j++) {
            int strong = 0;
            if (!(compute_row_sum && row_sum > max_row_sum))
                strong = ci[j] != row && ((diag < 0) ? va[j] > thr : va[j] < thr);
            s_con[j] = (unsigned char)strong;
            if (strong && ci[j] < n) cnt[ci[j]]++;
        }
    }
    for (int i = 0; i < n; i++) weights[i] = (float)cnt[i] + cla_hash(i);
    free(cnt);
}

/* computeWeightsKernel (Strength_All on S2): hash + in-degree, diagonal entries excluded */
static void cla_weights_pattern(int n, const int *rp, const int *ci, float *weights)
{
    int *cnt = (int *)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
    for (int i = 0; i < n; i++)
        for (int j = rp[i]; j < rp[i + 1]; j++)
            if (ci[j] != i) cnt[ci[j]]++;
    for (int i = 0; i < n; i++) weights[i] = (float)cnt[i] + cla_hash(i);
    free(cnt);
}

/* PMIS_Selector::markCoarseFinePoints_1x1 (device flow); s_con == NULL means "every entry strong" */
ORC_API void orc_cla_pmis(int n, const int *rp, const int *ci, const unsigned char *s_con, float *w, int *cf, int init)
{
    int *scratch = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int *mark = (int *)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
    for (int i = 0; init == 1 && i < n; i++) {   /* initialMarkingCfInitKernel (pmis.cu:316-360): cf comes from a previous pass (HMIS) */
        const int numj = rp[i + 1] - rp[i];
        if (numj == 0) cf[i] = CLA_FINE;
        else if (numj == 1 && ci[rp[i]] == i) cf[i] = CLA_FINE;
        else if (w[i] < 1) cf[i] = CLA_FINE;
        else if (cf[i] == CLA_STRONG_FINE) w[i] = 0.f;
        else if (cf[i] == CLA_FINE) { cf[i] = CLA_UNASSIGNED; mark[i] = 1; }
    }
    for (int i = 0; init != 1 && i < n; i++) {   /* initialMarkingKernel / initialMarkingCFInit3Kernel */
        const int numj = rp[i + 1] - rp[i];
        if (numj == 0) cf[i] = CLA_FINE;
        else if (numj == 1 && ci[rp[i]] == i) cf[i] = CLA_FINE;
        else if (w[i] < 1) cf[i] = CLA_FINE;
        else cf[i] = CLA_UNASSIGNED;
        int isolated = 1;
        for (int j = rp[i]; j < rp[i + 1]; j++)
            if (!s_con || s_con[j]) { isolated = 0; break; }
        if (isolated) { cf[i] = (init == 3) ? CLA_COARSE : CLA_STRONG_FINE; w[i] = 0.f; }
    }
    int iter = 0, num_unassigned;
    do {
        if (iter || !init) {
            for (int i = 0; i < n; i++) {   /* markUnassignedAsCoarseKernel */
                const int in = cf[i], un = (in == CLA_UNASSIGNED);
                mark[i] = un;
                scratch[i] = (w[i] > 1.f) ? (un ? CLA_COARSE : in) : in;
            }
            for (int i = 0; i < n; i++) {   /* markAdditionalCoarsePointsKernel (all writes store UNASSIGNED: order free) */
                if (mark[i] <= 0) continue;
                for (int j = rp[i]; j < rp[i + 1]; j++) {
                    if (s_con && !s_con[j]) continue;
                    const int jc = ci[j];
                    if (jc >= n) continue;
                    const float wc = w[jc], wr = w[i];
                    if (mark[jc] && wc > 1.0f) {
                        if (wr > wc) scratch[jc] = CLA_UNASSIGNED;
                        else if (wc > wr) scratch[i] = CLA_UNASSIGNED;
                    }
                }
            }
        } else {
            memcpy(scratch, cf, sizeof(int) * (size_t)n);
        }
        for (int i = 0; i < n; i++) {       /* markAdditionalFinePointsKernel */
            const int in = scratch[i];
            int fine = 0;
            if (in == CLA_UNASSIGNED)
                for (int j = rp[i]; !fine && j < rp[i + 1]; j++) {
                    if (s_con && !s_con[j]) continue;
                    if (ci[j] < n) fine = scratch[ci[j]] == CLA_COARSE;
                }
            cf[i] = fine ? CLA_FINE : in;
        }
        num_unassigned = 0;
        for (int i = 0; i < n; i++) num_unassigned += (cf[i] == CLA_UNASSIGNED);
        iter++;
    } while (num_unassigned != 0);
    free(scratch);
    free(mark);
}

/* RS_Selector<host>::markCoarseFinePoints_1x1 (src/classical/selectors/rs.cu:36-262): first pass of Ruge-Stueben coarsening,
 * sequential.  The reference keeps (weight, row) pairs in a std::set ordered by weight then by DEcreasing row, and always takes
 * rbegin(): the largest weight, the SMALLEST row among equals.  Here: a binary heap of (weight, row) with lazy deletion; a row
 * is in the set iff in_set[row], always with weight == iw[row] (the invariant of the reference's erase / insert pairs). */
typedef struct { int w, i; } rs_ent;
static int rs_before(rs_ent a, rs_ent b) { return a.w > b.w || (a.w == b.w && a.i < b.i); }   /* a leaves the heap before b */
typedef struct { rs_ent *e; int n, cap; } rs_heap;
static void rs_push(rs_heap *h, int w, int i)
{
    if (h->n == h->cap) { h->cap = h->cap ? 2 * h->cap : 1024; h->e = (rs_ent *)realloc(h->e, sizeof(rs_ent) * (size_t)h->cap); }
    int k = h->n++;
    rs_ent x = {w, i};
    while (k > 0) {
        const int p = (k - 1) / 2;
        if (!rs_before(x, h->e[p])) break;
        h->e[k] = h->e[p];
        k = p;
    }
    h->e[k] = x;
}
static rs_ent rs_pop(rs_heap *h)
{
    const rs_ent top = h->e[0], x = h->e[--h->n];
    int k = 0;
    for (;;) {
        int c = 2 * k + 1;
        if (c >= h->n) break;
        if (c + 1 < h->n && rs_before(h->e[c + 1], h->e[c])) c++;
        if (!rs_before(h->e[c], x)) break;
        h->e[k] = h->e[c];
        k = c;
    }
    if (h->n > 0) h->e[k] = x;
    return top;
}
ORC_API void orc_cla_rs(int n, const int *rp, const int *ci, const unsigned char *s_con, int *cf, int init)
{
#define RS_STRONG(k) ((!s_con || s_con[k]) && ci[k] < n)
    const size_t nn = (size_t)(n > 0 ? n : 1);
    int *stp = (int *)calloc((size_t)n + 2, sizeof(int)), *stc = (int *)malloc(sizeof(int) * (size_t)(rp[n] > 0 ? rp[n] : 1));
    for (int k = 0; k < rp[n]; k++) if (RS_STRONG(k)) stp[ci[k] + 1]++;
    for (int i = 0; i < n; i++) stp[i + 1] += stp[i];
    int *fillp = (int *)malloc(sizeof(int) * nn);
    for (int i = 0; i < n; i++) fillp[i] = stp[i];
    for (int i = 0; i < n; i++) for (int k = rp[i]; k < rp[i + 1]; k++) if (RS_STRONG(k)) stc[fillp[ci[k]]++] = i;   /* S^T, rows ascending */
    int *iw = (int *)malloc(sizeof(int) * nn);
    char *in_set = (char *)calloc(nn, 1);
    rs_heap h = {NULL, 0, 0};
#define RS_ERASE(i) (in_set[i] = 0)
#define RS_INSERT(i) do { in_set[i] = 1; rs_push(&h, iw[i], (i)); } while (0)
    for (int i = 0; i < n; i++) iw[i] = stp[i + 1] - stp[i];
    int num_left = 0;
    for (int j = 0; j < n; j++) {
        int isolated = 1;
        for (int k = rp[j]; k < rp[j + 1]; k++) if (RS_STRONG(k)) { isolated = 0; break; }
        if (isolated) { cf[j] = (init == 3) ? CLA_COARSE : CLA_STRONG_FINE; iw[j] = 0; }
        else { cf[j] = CLA_UNASSIGNED; num_left++; }
    }
    for (int j = 0; j < n; j++) {
        if (cf[j] == CLA_STRONG_FINE) continue;
        if (iw[j] > 0) { RS_INSERT(j); continue; }
        cf[j] = CLA_FINE;                                     /* nobody depends on j */
        for (int k = rp[j]; k < rp[j + 1]; k++) {
            if (!RS_STRONG(k)) continue;
            const int nb = ci[k];
            if (cf[nb] == CLA_STRONG_FINE) continue;
            if (nb < j) { if (iw[nb] > 0) RS_ERASE(nb); ++iw[nb]; RS_INSERT(nb); }
            else ++iw[nb];
        }
        --num_left;
    }
    while (num_left > 0) {
        int index = -1;
        while (h.n > 0) {                                     /* rbegin(): skip entries erased or re-weighted since they were pushed */
            const rs_ent t = rs_pop(&h);
            if (in_set[t.i] && iw[t.i] == t.w) { index = t.i; break; }
        }
        if (index < 0) break;                                 /* the reference would dereference an empty set here */
        cf[index] = CLA_COARSE;
        iw[index] = 0;
        --num_left;
        RS_ERASE(index);
        for (int j = stp[index]; j < stp[index + 1]; j++) {   /* rows that strongly depend on the new C point become F */
            const int nb = stc[j];
            if (cf[nb] != CLA_UNASSIGNED) continue;
            cf[nb] = CLA_FINE;
            RS_ERASE(nb);
            --num_left;
            for (int k = rp[nb]; k < rp[nb + 1]; k++) {
                if (!RS_STRONG(k)) continue;
                const int d2 = ci[k];
                if (cf[d2] == CLA_UNASSIGNED) { RS_ERASE(d2); ++iw[d2]; RS_INSERT(d2); }
            }
        }
        for (int j = rp[index]; j < rp[index + 1]; j++) {     /* points the new C point depends on lose one unit of measure */
            if (!RS_STRONG(j)) continue;
            const int nb = ci[j];
            if (cf[nb] != CLA_UNASSIGNED) continue;
            RS_ERASE(nb);
            const int wgt = --iw[nb];
            if (wgt > 0) { RS_INSERT(nb); continue; }
            cf[nb] = CLA_FINE;
            --num_left;
            for (int k = rp[nb]; k < rp[nb + 1]; k++) {
                if (!RS_STRONG(k)) continue;
                const int d2 = ci[k];
                if (cf[d2] == CLA_UNASSIGNED) { RS_ERASE(d2); ++iw[d2]; RS_INSERT(d2); }
            }
        }
    }
#undef RS_STRONG
#undef RS_ERASE
#undef RS_INSERT
    free(stp); free(stc); free(fillp); free(iw); free(in_set); free(h.e);
}

/* HMIS_Selector<device>::markCoarseFinePoints_1x1 (src/classical/selectors/hmis.cu:58-88): Ruge-Stueben first pass on the host
 * (always with cf_map_init = 0), then PMIS with cf_map_init = 1, whatever cf_map_init the caller asked for */
ORC_API void orc_cla_hmis(int n, const int *rp, const int *ci, const unsigned char *s_con, float *w, int *cf)
{
    orc_cla_rs(n, rp, ci, s_con, cf, 0);
    orc_cla_pmis(n, rp, ci, s_con, w, cf, 1);
}
static int g_cla_selector = 0;       /* 0 PMIS, 1 HMIS: selector of the NEXT classical setup (and of its aggressive levels) */
ORC_API void orc_set_classical_selector(int sel) { g_cla_selector = sel; }

/* renumberAndCountCoarsePoints: COARSE -> 0,1,2,... in row order */
ORC_API int orc_cla_renumber(int n, int *cf)
{
    int nc = 0;
    for (int i = 0; i < n; i++)
        if (cf[i] == CLA_COARSE) cf[i] = nc++;
    return nc;
}

static int cla_insert_sorted(int *a, int m, int key) /* sorted unique insert; returns new size */
{
    int lo = 0, hi = m;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    if (lo < m && a[lo] == key) return m;
    for (int k = m; k > lo; k--) a[k] = a[k - 1];
    a[lo] = key;
    return m + 1;
}
static int cla_find_sorted(const int *a, int m, int key)
{
    int lo = 0, hi = m;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return (lo < m && a[lo] == key) ? lo : -1;
}


/* ---- the reference's row order --------------------------------------------------------------------------------
 * A row of P leaves the reference in the slot order of its Hash_set / Hash_map (include/hash_containers_detail.inl):
 * 128 shared-memory slots, slot = ((key ^ c[f]) + c[4+f]) & 127 for the first of four hash functions f whose slot is
 * free or already holds the key; a warp inserts up to 32 keys at once and, when several lanes race for the same empty
 * slot, the LOWEST lane wins (observed: this rule reproduces every row of the reference dumps, tests/golden/
 * *classical*.npz); keys that lose all four rounds go to a global-memory table (same functions, mask gmem_size-1) and
 * are stored after the shared-memory ones.  The max-elements truncation keeps the FIRST of several equal weights, so
 * this order decides which coarse points survive on symmetric stencils.  The emulation below is sequential: one
 * "step" = one warp-wide insert call, lanes visited in ascending order inside every hash round. */
static const unsigned cla_hash_keys[8] = {3499211612u, 581869302u, 3890346734u, 3586334585u, 545404204u, 4161255391u, 3922919429u, 949333985u};
#define CLA_SLOTS 128
#define CLA_OVF 64
typedef struct { int tab[CLA_SLOTS]; int ovf_slot[CLA_OVF], ovf_key[CLA_OVF], n_ovf, gmem_mask, failed; } cla_slotset;
static void cla_ss_clear(cla_slotset *h, int gmem_size) { for (int s = 0; s < CLA_SLOTS; s++) h->tab[s] = -1; h->n_ovf = 0; h->gmem_mask = gmem_size - 1; h->failed = 0; }
static unsigned cla_ss_hash(int key, int f) { return ((unsigned)key ^ cla_hash_keys[f]) + cla_hash_keys[4 + f]; }
static void cla_ss_insert_step(cla_slotset *h, int *keys, int cnt)   /* keys[l] = key of lane l or -1; destroyed */
{
    for (int f = 0; f < 4; f++) {
        int any = 0;
        for (int l = 0; l < cnt; l++) {
            const int k = keys[l];
            if (k == -1) continue;
            const int s = (int)(cla_ss_hash(k, f) & (CLA_SLOTS - 1));
            if (h->tab[s] == -1) { h->tab[s] = k; keys[l] = -1; }
            else if (h->tab[s] == k) keys[l] = -1;
            else any = 1;
        }
        if (!any) return;
    }
    for (int f = 0; f < 4; f++) {
        int any = 0;
        for (int l = 0; l < cnt; l++) {
            const int k = keys[l];
            if (k == -1) continue;
            const int s = (int)(cla_ss_hash(k, f) & (unsigned)h->gmem_mask);
            int q = -1;
            for (int t = 0; t < h->n_ovf; t++) if (h->ovf_slot[t] == s) { q = t; break; }
            if (q < 0) {
                if (h->n_ovf < CLA_OVF) { h->ovf_slot[h->n_ovf] = s; h->ovf_key[h->n_ovf] = k; h->n_ovf++; keys[l] = -1; }
                else { h->failed = 1; keys[l] = -1; }
            } else if (h->ovf_key[q] == k) keys[l] = -1;
            else any = 1;
        }
        if (!any) return;
    }
    h->failed = 1;   /* the reference would double its global table and retry (status = 1); not emulated */
}
static int cla_ss_store(const cla_slotset *h, int *out)   /* shared-memory slots ascending, then global-memory slots ascending */
{
    int m = 0;
    for (int s = 0; s < CLA_SLOTS; s++) if (h->tab[s] != -1) out[m++] = h->tab[s];
    int idx[CLA_OVF];
    for (int t = 0; t < h->n_ovf; t++) idx[t] = t;
    for (int a = 1; a < h->n_ovf; a++) { const int v = idx[a]; int b = a - 1; while (b >= 0 && h->ovf_slot[idx[b]] > h->ovf_slot[v]) { idx[b + 1] = idx[b]; b--; } idx[b + 1] = v; }
    for (int t = 0; t < h->n_ovf; t++) out[m++] = h->ovf_key[idx[t]];
    return m;
}
static int cla_find_linear(const int *a, int m, int key) { for (int k = 0; k < m; k++) if (a[k] == key) return k; return -1; }

/* distance2::compute_c_hat_kernel (distance2.cu:848-1170): insertion sequence of a FINE row.  wide == 0: the
 * 8-lanes-per-row variant (avg nnz per row < 16), four rows of B in flight; wide == 1: one row of B per step, 32 lanes. */
static int cla_c_hat_fill_ref_order(int i, const int *rp, const int *ci, const unsigned char *s_con, const int *cf, int wide, int gmem_size, int *out)
{
    cla_slotset h;
    cla_ss_clear(&h, gmem_size);
    int keys[32], fines[32];
    for (int c0 = rp[i]; c0 < rp[i + 1]; c0 += 32) {
        int nf = 0;
        const int c1 = (c0 + 32 < rp[i + 1]) ? c0 + 32 : rp[i + 1];
        for (int l = 0; l < 32; l++) {
            keys[l] = -1;
            const int k = c0 + l;
            if (k >= c1) continue;
            const int c = ci[k];
            if (c == i || !s_con[k]) continue;
            if (cf[c] == CLA_FINE) fines[nf++] = c;
            else if (cf[c] != CLA_STRONG_FINE) keys[l] = c;
        }
        cla_ss_insert_step(&h, keys, 32);
        if (!wide) {
            for (int g0 = 0; g0 < nf; g0 += 4) {
                for (int t = 0;; t++) {
                    int any = 0;
                    for (int l = 0; l < 32; l++) {
                        keys[l] = -1;
                        const int gi = l >> 3, m = l & 7;
                        if (g0 + gi >= nf) continue;
                        const int b = fines[g0 + gi], k = rp[b] + m + 8 * t;
                        if (k >= rp[b + 1]) continue;
                        any = 1;
                        const int c = ci[k];
                        if (c != b && s_con[k] && cf[c] != CLA_FINE && cf[c] != CLA_STRONG_FINE) keys[l] = c;
                    }
                    if (!any) break;
                    cla_ss_insert_step(&h, keys, 32);
                }
            }
        } else {
            for (int g = 0; g < nf; g++) {
                const int b = fines[g];
                for (int k0 = rp[b]; k0 < rp[b + 1]; k0 += 32) {
                    for (int l = 0; l < 32; l++) {
                        keys[l] = -1;
                        const int k = k0 + l;
                        if (k >= rp[b + 1]) continue;
                        const int c = ci[k];
                        if (c != b && s_con[k] && cf[c] != CLA_FINE && cf[c] != CLA_STRONG_FINE) keys[l] = c;
                    }
                    cla_ss_insert_step(&h, keys, 32);
                }
            }
        }
    }
    return cla_ss_store(&h, out);
}

/* The distance-two coarse set of row i: strong coarse neighbours plus strong coarse neighbours of its
 * strong FINE neighbours (estimate_c_hat_size_kernel / compute_c_hat_kernel of selector.cu and
 * distance2.cu share this rule).  `cf` holds coarse ids >= 0, FINE, STRONG_FINE.  Fine-grid ids, sorted. */
static int cla_c_hat_upper(int i, const int *rp, const int *ci, const unsigned char *s_con, const int *cf)
{
    int count = 0;
    for (int j = rp[i]; j < rp[i + 1]; j++) {
        const int c = ci[j];
        if (c == i || !s_con[j]) continue;
        if (cf[c] == CLA_FINE) {
            for (int jj = rp[c]; jj < rp[c + 1]; jj++)
                if (ci[jj] != c && s_con[jj] && cf[ci[jj]] != CLA_FINE && cf[ci[jj]] != CLA_STRONG_FINE) count++;
        } else if (cf[c] != CLA_STRONG_FINE) count++;
    }
    return count;
}
static int cla_c_hat_fill(int i, const int *rp, const int *ci, const unsigned char *s_con, const int *cf, int *out)
{
    int m = 0;
    for (int j = rp[i]; j < rp[i + 1]; j++) {
        const int c = ci[j];
        if (c == i || !s_con[j]) continue;
        if (cf[c] == CLA_FINE) {
            for (int jj = rp[c]; jj < rp[c + 1]; jj++)
                if (ci[jj] != c && s_con[jj] && cf[ci[jj]] != CLA_FINE && cf[ci[jj]] != CLA_STRONG_FINE) m = cla_insert_sorted(out, m, ci[jj]);
        } else if (cf[c] != CLA_STRONG_FINE) m = cla_insert_sorted(out, m, c);
    }
    return m;
}

/* Aggressive_PMIS_Selector::markCoarseFinePoints: PMIS, S2 among the coarse points, PMIS on S2, correctCfMap */
ORC_API void orc_cla_aggressive_pmis(int n, const int *rp, const int *ci, const unsigned char *s_con, float *w, int *cf)
{
    /* Aggressive_PMIS / Aggressive_HMIS selectors differ only in the selector applied to A and to S2 (aggressive_hmis.cu:41,131) */
    if (g_cla_selector == 1) orc_cla_hmis(n, rp, ci, s_con, w, cf);
    else orc_cla_pmis(n, rp, ci, s_con, w, cf, 0);
    int *scanned = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    memcpy(scanned, cf, sizeof(int) * (size_t)n);
    const int nc = orc_cla_renumber(n, scanned);
    if (nc == 0) { free(scanned); return; }
    int *s2p = (int *)calloc((size_t)nc + 1, sizeof(int));
    long long ub = 0;
    for (int i = 0; i < n; i++) if (scanned[i] >= 0) ub += cla_c_hat_upper(i, rp, ci, s_con, scanned);
    int *s2c = (int *)malloc(sizeof(int) * (size_t)(ub > 0 ? ub : 1));
    int pos = 0;
    for (int i = 0; i < n; i++) {
        if (scanned[i] < 0) continue;
        const int m = cla_c_hat_fill(i, rp, ci, s_con, scanned, s2c + pos);
        for (int k = 0; k < m; k++) s2c[pos + k] = scanned[s2c[pos + k]];   /* fillS2ColIndices: coarse ids (monotone map keeps the order) */
        pos += m;
        s2p[scanned[i] + 1] = pos;
    }
    float *w2 = (float *)malloc(sizeof(float) * (size_t)nc);
    int *cf2 = (int *)malloc(sizeof(int) * (size_t)nc);
    cla_weights_pattern(nc, s2p, s2c, w2);
    if (g_cla_selector == 1) orc_cla_hmis(nc, s2p, s2c, NULL, w2, cf2);
    else orc_cla_pmis(nc, s2p, s2c, NULL, w2, cf2, 3);
    for (int i = 0; i < n; i++)      /* correctCfMapKernel */
        if (cf[i] == CLA_COARSE) { const int c2 = cf2[scanned[i]]; cf[i] = (c2 == CLA_STRONG_FINE) ? CLA_COARSE : c2; }
    free(scanned); free(s2p); free(s2c); free(w2); free(cf2);
}

static int cla_sign(double x) { return x >= 0.0; }   /* include/classical/interpolators/common.h:25-31 */

static void cla_diag(int n, const int *rp, const int *ci, const double *va, double *d)
{
    for (int i = 0; i < n; i++) { d[i] = 0; for (int j = rp[i]; j < rp[i + 1]; j++) if (ci[j] == i) { d[i] = va[j]; break; } }
}

typedef struct { int n, nc, nnz; int *rp, *ci; double *va; } cla_csr;
static void cla_csr_free(cla_csr *m) { free(m->rp); free(m->ci); free(m->va); m->rp = m->ci = NULL; m->va = NULL; }

/* Distance2_Interpolator (device): extended+i weights.  cf: coarse ids / FINE / STRONG_FINE. */
static void cla_interp_d2(int n, const int *rp, const int *ci, const double *va, const int *cf, const unsigned char *s_con, int nc, cla_csr *P)
{
    double *diag = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    cla_diag(n, rp, ci, va, diag);
    const int wide = !(n > 0 && rp[n] / n < 16);     /* kernel variant picked by the reference: avg nnz per row < 16 -> 8 lanes per row */
    int *prp = (int *)calloc((size_t)n + 1, sizeof(int));
    long long ub = 0;
    for (int i = 0; i < n; i++) ub += (cf[i] >= 0) ? 1 : (cf[i] == CLA_STRONG_FINE ? 0 : cla_c_hat_upper(i, rp, ci, s_con, cf));
    int *chat = (int *)malloc(sizeof(int) * (size_t)(ub > 0 ? ub : 1));      /* fine ids of the coarse set, later mapped to coarse ids */
    double *pv = (double *)malloc(sizeof(double) * (size_t)(ub > 0 ? ub : 1));
    int pos = 0;
    for (int i = 0; i < n; i++) {
        prp[i] = pos;
        if (cf[i] >= 0) { chat[pos] = i; pv[pos] = 1.0; pos++; continue; }
        if (cf[i] == CLA_STRONG_FINE) continue;
        int *ch = chat + pos;
        double *val = pv + pos;
        const int m = cla_c_hat_fill_ref_order(i, rp, ci, s_con, cf, wide, 512, ch);   /* gmem_size 512: distance2.cu:1912-1914 (sm >= 7) */
        for (int k = 0; k < m; k++) val[k] = 0.0;
        const int sign_i = cla_sign(diag[i]);
        double weak = 0.0;
        for (int j = rp[i]; j < rp[i + 1]; j++) {
            const int c = ci[j];
            const double a = va[j];
            const int offd = (c != i);
            const int strong = offd && s_con[j];
            const int p = cla_find_linear(ch, m, c);
            if (p >= 0) val[p] += a;
            if (offd && !strong && p < 0 && cf[c] != CLA_STRONG_FINE) weak += a;
            if (strong && cf[c] == CLA_FINE) {
                /* compute_inner_sum_kernel: a_ik / sum over l in C_hat(i) + {i} of the entries of row k whose sign differs from a_ii */
                double bottom = 0.0;
                for (int jj = rp[c]; jj < rp[c + 1]; jj++) {
                    const int l = ci[jj];
                    const int needed = (l == i) || cla_find_linear(ch, m, l) >= 0;
                    const double b = needed ? va[jj] : 0.0;
                    if (sign_i != cla_sign(b)) bottom += b;
                }
                const double inner = (bottom != 0.0) ? a / bottom : a;
                /* compute_interp_weight_kernel: distribute row k */
                const double dk = diag[c];
                double aki = 0.0;
                for (int jj = rp[c]; jj < rp[c + 1]; jj++) {
                    const int l = ci[jj];
                    double b = va[jj];
                    if (cla_sign(dk) == cla_sign(b)) b = 0.0;
                    if (l == i) aki = b;
                    const int q = cla_find_linear(ch, m, l);
                    if (q >= 0) { const double t = b * inner; val[q] += t; }
                }
                { const double t = aki * inner; weak += t; }
            }
        }
        weak += diag[i];
        const double scale = -1.0 / weak;
        for (int k = 0; k < m; k++) val[k] = scale * val[k];
        pos += m;
    }
    prp[n] = pos;
    for (int i = 0; i < n; i++)
        for (int k = prp[i]; k < prp[i + 1]; k++) chat[k] = cf[chat[k]];   /* store_map_keys: fine id -> coarse id */
    P->n = n; P->nc = nc; P->nnz = pos; P->rp = prp; P->ci = chat; P->va = pv;
    free(diag);
}

/* Multipass_Interpolator (device) */
/* Distance1_Interpolator<device>::generateInterpolationMatrix_1x1 (src/classical/interpolators/distance1.cu:353-867), the reference's
 * DEFAULT interpolator ("D1"): one thread per row, everything sequential inside a row.
 *   sets (categoriseEdgesKernel :690-745): for a non-coarse row, every off-diagonal entry is strong-coarse (1), weak-coarse (2),
 *     strong-fine (4: the two rows share a coarse neighbour, found by a two-pointer walk over both rows that assumes ascending
 *     columns, :645-687), strong-fine-without-common-C (8) or weak-fine (16);
 *   B_ij (calculateBKernel :426-548) = sum over strong-fine k of a_ik a_kj / sum_{m strong coarse} a_km, only entries of row k whose
 *     sign is opposite to a_kk count; a k whose denominator vanishes (|.| < 1e-10) goes to the diagonal once (first strong-coarse j);
 *   D_i (calculateDKernel :400-422) += weak-fine entries;  w_ij = -1 / (a_ii + D_i) * (a_ij + B_ij)  (calculateWKernel :577-612).
 * Quirks kept: a STRONG_FINE row gets ONE explicit entry (column 0, value 0) because numNonZerosVecKernel counts 1 for every row
 * that is not FINE (:374-386); calculateBKernel leaves a thread at its first coarse row ("return" inside the grid-stride loop, :448),
 * so with the reference's launch of 4096 x 64 threads a row i is skipped (B = 0, D not reset) when some row i - m * 262144 is coarse. */
#define D1_STRONG_COARSE 1
#define D1_WEAK_COARSE 2
#define D1_STRONG_FINE 4
#define D1_STRONG_FINE_NO_COMMON 8
#define D1_WEAK_FINE 16
#define D1_REF_THREADS 262144
static int d1_intersect(const int *ci, const unsigned char *mark, int b1, int e1, int b2, int e2)
{
    int i1 = b1, i2 = b2;
    if (b1 >= e1 || b2 >= e2) return 0;
    for (;;) {
        const int c1 = ci[i1], c2 = ci[i2];
        if (c1 == c2) {
            if (mark[i1] && mark[i2]) return 1;
            i1++; i2++;
            if (i1 >= e1 || i2 >= e2) return 0;
        } else if (c1 > c2) { if (++i2 >= e2) return 0; }
        else { if (++i1 >= e1) return 0; }
    }
}
static void cla_interp_d1(int n, const int *rp, const int *ci, const double *va, const int *cf, const unsigned char *s_con, int nc, cla_csr *P)
{
    const int nnz = rp[n];
    double *diag = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    cla_diag(n, rp, ci, va, diag);
    int *prp = (int *)calloc((size_t)n + 1, sizeof(int));
    for (int i = 0; i < n; i++) {                                   /* numNonZerosVecKernel + exclusive scan */
        int cnt = 0;
        if (cf[i] == CLA_FINE) { for (int j = rp[i]; j < rp[i + 1]; j++) if (s_con[j] && cf[ci[j]] >= 0) cnt++; }
        else cnt = 1;
        prp[i + 1] = prp[i] + cnt;
    }
    const int pnnz = prp[n];
    unsigned char *mark = (unsigned char *)calloc((size_t)(nnz > 0 ? nnz : 1), 1);
    int *set = (int *)calloc((size_t)(nnz > 0 ? nnz : 1), sizeof(int));
    for (int i = 0; i < n; i++)                                     /* markCoarseEdgesKernel */
        for (int j = rp[i]; j < rp[i + 1]; j++) if (ci[j] != i && cf[ci[j]] >= 0) mark[j] = 1;
    for (int i = 0; i < n; i++) {                                   /* categoriseEdgesKernel */
        if (cf[i] >= 0) continue;
        for (int j = rp[i]; j < rp[i + 1]; j++) {
            const int jc = ci[j];
            if (jc == i) continue;
            if (cf[jc] >= 0) set[j] |= s_con[j] ? D1_STRONG_COARSE : D1_WEAK_COARSE;
            else if (!s_con[j]) set[j] |= D1_WEAK_FINE;
            else set[j] |= d1_intersect(ci, mark, rp[i], rp[i + 1], rp[jc], rp[jc + 1]) ? D1_STRONG_FINE : D1_STRONG_FINE_NO_COMMON;
        }
    }
    int *pc = (int *)calloc((size_t)(pnnz > 0 ? pnnz : 1), sizeof(int));
    double *pv = (double *)calloc((size_t)(pnnz > 0 ? pnnz : 1), sizeof(double));
    double *B = (double *)calloc((size_t)(pnnz > 0 ? pnnz : 1), sizeof(double)), *D = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    for (int i = 0; i < n; i++) {                                   /* calculateBKernel */
        int skipped = 0;                                            /* the thread that owns row i returned at an earlier coarse row */
        for (int e = i - D1_REF_THREADS; e >= 0 && !skipped; e -= D1_REF_THREADS) skipped = cf[e] >= 0;
        if (skipped) continue;
        if (cf[i] >= 0) { B[prp[i]] = 1; continue; }
        D[i] = 0;
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
                if (fabs(bottom) < tol) { if (first_j_loop == 1) D[i] += va[k]; }
                else sum += top / bottom;
            }
            B[prp[i] + local] = sum;
            local++;
        }
    }
    for (int i = 0; i < n; i++) {                                   /* calculateDKernel */
        double sum = 0;
        for (int k = rp[i]; k < rp[i + 1]; k++) if (set[k] & D1_WEAK_FINE) sum += va[k];
        D[i] += sum;
    }
    for (int i = 0; i < n; i++) {                                   /* calculateWKernel */
        if (cf[i] >= 0) { pv[prp[i]] = 1.0; pc[prp[i]] = cf[i]; continue; }
        int local = 0;
        for (int j = rp[i]; j < rp[i + 1]; j++) {
            if (!(set[j] & D1_STRONG_COARSE)) continue;
            const double bottom = (fabs(diag[i] + D[i]) < 1e-10) ? 1. : diag[i] + D[i];
            pc[prp[i] + local] = cf[ci[j]];
            pv[prp[i] + local] = -1.0 / bottom * (va[j] + B[prp[i] + local]);
            local++;
        }
    }
    P->n = n; P->nc = nc; P->nnz = pnnz; P->rp = prp; P->ci = pc; P->va = pv;
    free(diag); free(mark); free(set); free(B); free(D);
}

static void cla_interp_multipass(int n, const int *rp, const int *ci, const double *va, const int *cf, const unsigned char *s_con, int nc, cla_csr *P)
{
    double *diag = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    cla_diag(n, rp, ci, va, diag);
    int *assigned = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int *ub = (int *)calloc((size_t)n + 1, sizeof(int));
    int num_unassigned = 0, num_sf = 0;
    for (int i = 0; i < n; i++) {          /* initializeAssignedArray */
        assigned[i] = -1;
        if (cf[i] >= 0) { assigned[i] = 0; ub[i] = 1; }
        else if (cf[i] == CLA_FINE) {
            int cc = 0;
            for (int j = rp[i]; j < rp[i + 1]; j++) if (ci[j] != i && s_con[j] && cf[ci[j]] >= 0) cc++;
            if (cc) { assigned[i] = 1; ub[i] = cc; }
        }
        if (assigned[i] < 0) num_unassigned++;
        if (cf[i] == CLA_STRONG_FINE) num_sf++;
    }
    int pass = 2;
    int remaining = num_unassigned - num_sf;
    while (remaining && pass < 10) {       /* fillAssignedArray: reads see values < pass only */
        for (int i = 0; i < n; i++) {
            if (assigned[i] != -1) continue;
            for (int j = rp[i]; j < rp[i + 1]; j++)
                if (ci[j] != i && s_con[j] && assigned[ci[j]] == pass - 1) { assigned[i] = -pass - 100; break; }
        }
        num_unassigned = 0;
        for (int i = 0; i < n; i++) { if (assigned[i] == -pass - 100) assigned[i] = pass; if (assigned[i] < 0) num_unassigned++; }
        remaining = num_unassigned - num_sf;
        pass++;
    }
    const int num_passes = pass;
    for (int p = 2; p < num_passes; p++)   /* estimate_c_hat_size_kernel: upper bounds chain through the passes */
        for (int i = 0; i < n; i++) {
            if (assigned[i] != p) continue;
            int c = 0;
            for (int j = rp[i]; j < rp[i + 1]; j++) if (ci[j] != i && s_con[j] && assigned[ci[j]] == p - 1) c += ub[ci[j]];
            ub[i] = c;
        }
    long long *off = (long long *)malloc(sizeof(long long) * ((size_t)n + 1));
    off[0] = 0;
    for (int i = 0; i < n; i++) off[i + 1] = off[i] + ub[i];
    int *cols = (int *)malloc(sizeof(int) * (size_t)(off[n] > 0 ? off[n] : 1));
    double *vals = (double *)malloc(sizeof(double) * (size_t)(off[n] > 0 ? off[n] : 1));
    int *len = (int *)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
    for (int i = 0; i < n; i++) {          /* coarse rows and first pass (compute_interp_weight_first_pass_kernel) */
        int *pc = cols + off[i];
        double *pvv = vals + off[i];
        if (assigned[i] == 0) { pc[0] = cf[i]; pvv[0] = 1.0; len[i] = 1; }
        else if (assigned[i] == 1) {
            double sum_N = 0.0, sum_C = 0.0;
            int m = 0;
            for (int j = rp[i]; j < rp[i + 1]; j++) {
                const int c = ci[j];
                if (c == i) continue;
                if (cf[c] != CLA_STRONG_FINE) sum_N += va[j];
                if (s_con[j] && assigned[c] == 0) { sum_C += va[j]; pc[m] = cf[c]; pvv[m] = va[j]; m++; }
            }
            const double sd = sum_C * diag[i];
            const double div = (fabs(sd) == 0.0) ? 1.0 : sd;
            const double alfa = -sum_N / div;
            for (int k = 0; k < m; k++) pvv[k] *= alfa;
            len[i] = m;
        }
    }
    for (int p = 2; p < num_passes; p++) { /* compute_c_hat_kernel + compute_interp_weight_kernel, one pass at a time */
        for (int i = 0; i < n; i++) {
            if (assigned[i] != p) continue;
            int *pc = cols + off[i];
            double *pvv = vals + off[i];
            /* multipass::compute_c_hat_kernel<8,...> (multipass.cu:762-905): union of the coarse sets of the strong
             * neighbours assigned in the previous pass, four neighbours in flight, 8 lanes each; keys = coarse ids */
            cla_slotset h;
            cla_ss_clear(&h, 2048);
            int keys[32], nb[32];
            for (int c0 = rp[i]; c0 < rp[i + 1]; c0 += 32) {
                int nn = 0;
                const int c1 = (c0 + 32 < rp[i + 1]) ? c0 + 32 : rp[i + 1];
                for (int j = c0; j < c1; j++) if (ci[j] != i && s_con[j] && assigned[ci[j]] == p - 1) nb[nn++] = ci[j];
                for (int g0 = 0; g0 < nn; g0 += 4)
                    for (int t = 0;; t++) {
                        int any = 0;
                        for (int l = 0; l < 32; l++) {
                            keys[l] = -1;
                            const int gi = l >> 3, mm = l & 7;
                            if (g0 + gi >= nn) continue;
                            const int b = nb[g0 + gi], idx = mm + 8 * t;
                            if (idx >= len[b]) continue;
                            any = 1;
                            keys[l] = cols[off[b] + idx];
                        }
                        if (!any) break;
                        cla_ss_insert_step(&h, keys, 32);
                    }
            }
            const int m = cla_ss_store(&h, pc);
            for (int q = 0; q < m; q++) pvv[q] = 0.0;
            double sum_N = 0.0, sum_C = 0.0;
            for (int j = rp[i]; j < rp[i + 1]; j++) {
                const int k = ci[j];
                if (k == i) continue;
                const int sa = s_con[j] && assigned[k] == p - 1;
                if (!sa) { if (cf[k] != CLA_STRONG_FINE) sum_N += va[j]; continue; }
                for (int q = 0; q < len[k]; q++) {
                    const double tmp = vals[off[k] + q] * va[j];
                    sum_C += tmp;
                    sum_N += tmp;
                    pvv[cla_find_linear(pc, m, cols[off[k] + q])] += tmp;
                }
            }
            const double sd = sum_C * diag[i];
            const double div = (fabs(sd) == 0.0) ? 1.0 : sd;
            const double alfa = -sum_N / div;
            for (int q = 0; q < m; q++) pvv[q] = alfa * pvv[q];
            len[i] = m;
        }
    }
    int *prp = (int *)malloc(sizeof(int) * ((size_t)n + 1));
    prp[0] = 0;
    for (int i = 0; i < n; i++) prp[i + 1] = prp[i] + len[i];
    int *pci = (int *)malloc(sizeof(int) * (size_t)(prp[n] > 0 ? prp[n] : 1));
    double *pva = (double *)malloc(sizeof(double) * (size_t)(prp[n] > 0 ? prp[n] : 1));
    for (int i = 0; i < n; i++) {
        memcpy(pci + prp[i], cols + off[i], sizeof(int) * (size_t)len[i]);
        memcpy(pva + prp[i], vals + off[i], sizeof(double) * (size_t)len[i]);
    }
    P->n = n; P->nc = nc; P->nnz = prp[n]; P->rp = prp; P->ci = pci; P->va = pva;
    free(diag); free(assigned); free(ub); free(off); free(cols); free(vals); free(len);
}

/* Truncate::truncateByMaxElements (device, non-NLARGEST path): keep the max_elmts largest |values| of each row
 * (earlier entries win ties), rescale to the original row sum.  Output order: descending |value| as stored by
 * the reference's kernel. */
static void cla_truncate(cla_csr *P, int max_elmts)
{
    const int n = P->n;
    int *nrp = (int *)malloc(sizeof(int) * ((size_t)n + 1));
    nrp[0] = 0;
    for (int i = 0; i < n; i++) { const int l = P->rp[i + 1] - P->rp[i]; nrp[i + 1] = nrp[i] + (l < max_elmts ? l : max_elmts); }
    int *nci = (int *)malloc(sizeof(int) * (size_t)(nrp[n] > 0 ? nrp[n] : 1));
    double *nva = (double *)malloc(sizeof(double) * (size_t)(nrp[n] > 0 ? nrp[n] : 1));
    for (int i = 0; i < n; i++) {
        const int s = P->rp[i], e = P->rp[i + 1], len = e - s;
        int *oc = nci + nrp[i];
        double *ov = nva + nrp[i];
        double orig = 0.0;
        for (int j = s; j < e; j++) orig += P->va[j];
        int m;
        if (len <= max_elmts) {
            m = len;
            for (int j = 0; j < len; j++) { oc[j] = P->ci[s + j]; ov[j] = P->va[s + j]; }
        } else {
            m = max_elmts;
            for (int j = 0; j < m; j++) { oc[j] = P->ci[s + j]; ov[j] = P->va[s + j]; }
            int nn = m;                     /* sortByFabs: bubble sort, strict < */
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
            for (int j = s + m; j < e; j++)
                for (int q = 0; q < m; q++)
                    if (fabs(P->va[j]) > fabs(ov[q])) {
                        for (int k = m - 1; k > q; k--) { ov[k] = ov[k - 1]; oc[k] = oc[k - 1]; }
                        ov[q] = P->va[j]; oc[q] = P->ci[j];
                        break;
                    }
        }
        double nsum = 0.0;
        for (int j = 0; j < m; j++) nsum += ov[j];
        const double mult = (fabs(nsum) == 0.0) ? 1.0 : orig / nsum;     /* scale_kernel */
        for (int j = 0; j < m; j++) ov[j] = ov[j] * mult;
    }
    free(P->rp); free(P->ci); free(P->va);
    P->rp = nrp; P->ci = nci; P->va = nva; P->nnz = nrp[n];
}

static void cla_transpose(const cla_csr *P, cla_csr *R)   /* R = P^T, rows ordered by ascending fine index */
{
    const int n = P->n, nc = P->nc;
    int *rp = (int *)calloc((size_t)nc + 1, sizeof(int));
    for (int k = 0; k < P->nnz; k++) rp[P->ci[k] + 1]++;
    for (int c = 0; c < nc; c++) rp[c + 1] += rp[c];
    int *ci = (int *)malloc(sizeof(int) * (size_t)(P->nnz > 0 ? P->nnz : 1));
    double *va = (double *)malloc(sizeof(double) * (size_t)(P->nnz > 0 ? P->nnz : 1));
    int *cur = (int *)malloc(sizeof(int) * (size_t)(nc > 0 ? nc : 1));
    memcpy(cur, rp, sizeof(int) * (size_t)nc);
    for (int i = 0; i < n; i++)
        for (int k = P->rp[i]; k < P->rp[i + 1]; k++) { const int q = cur[P->ci[k]]++; ci[q] = i; va[q] = P->va[k]; }
    free(cur);
    R->n = nc; R->nc = n; R->nnz = P->nnz; R->rp = rp; R->ci = ci; R->va = va;
}

/* C = A * B, rows accumulated in the storage order of A's row and of B's rows, product rounded before the add,
 * output columns ascending.  (csr_galerkin_product computes the same numbers in hash/atomic order.) */
static void cla_spgemm(int m, const int *arp, const int *aci, const double *ava, const int *brp, const int *bci, const double *bva, int ncols, cla_csr *Cm)
{
    double *acc = (double *)calloc((size_t)(ncols > 0 ? ncols : 1), sizeof(double));
    int *mark = (int *)malloc(sizeof(int) * (size_t)(ncols > 0 ? ncols : 1));
    for (int c = 0; c < ncols; c++) mark[c] = -1;
    int *list = (int *)malloc(sizeof(int) * (size_t)(ncols > 0 ? ncols : 1));
    int *crp = (int *)malloc(sizeof(int) * ((size_t)m + 1));
    size_t cap = 1024, nnz = 0;
    int *cci = (int *)malloc(sizeof(int) * cap);
    double *cva = (double *)malloc(sizeof(double) * cap);
    crp[0] = 0;
    for (int i = 0; i < m; i++) {
        int cnt = 0;
        for (int j = arp[i]; j < arp[i + 1]; j++) {
            const int k = aci[j];
            const double a = ava[j];
            for (int q = brp[k]; q < brp[k + 1]; q++) {
                const int c = bci[q];
                const double t = a * bva[q];
                if (mark[c] != i) { mark[c] = i; acc[c] = t; list[cnt++] = c; }
                else acc[c] += t;
            }
        }
        qsort(list, (size_t)cnt, sizeof(int), cmp_int);
        if (nnz + (size_t)cnt > cap) { while (nnz + (size_t)cnt > cap) cap *= 2; cci = (int *)realloc(cci, sizeof(int) * cap); cva = (double *)realloc(cva, sizeof(double) * cap); }
        for (int q = 0; q < cnt; q++) { cci[nnz] = list[q]; cva[nnz] = acc[list[q]]; nnz++; }
        crp[i + 1] = (int)nnz;
    }
    free(acc); free(mark); free(list);
    Cm->n = m; Cm->nc = ncols; Cm->nnz = (int)nnz; Cm->rp = crp; Cm->ci = cci; Cm->va = cva;
}

/* y = M x with the engine's per-row order (left to right, one FMA per entry) */
static void cla_spmv(const int n, const int *rp, const int *ci, const double *va, const double *x, double *y)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int j = rp[i]; j < rp[i + 1]; j++) s = fma(va[j], x[ci[j]], s);
        y[i] = s;
    }
}

/* ---- stage-level entry points used by the tests ------------------------------------------------ */
typedef struct { cla_csr m; } orc_cla_matrix;
ORC_API void orc_cla_matrix_free(orc_cla_matrix *h) { if (h) { cla_csr_free(&h->m); free(h); } }
ORC_API void orc_cla_matrix_sizes(const orc_cla_matrix *h, int *n, int *nc, int *nnz) { *n = h->m.n; *nc = h->m.nc; *nnz = h->m.nnz; }
ORC_API void orc_cla_matrix_get(const orc_cla_matrix *h, int *rp, int *ci, double *va)
{
    memcpy(rp, h->m.rp, sizeof(int) * ((size_t)h->m.n + 1));
    memcpy(ci, h->m.ci, sizeof(int) * (size_t)h->m.nnz);
    memcpy(va, h->m.va, sizeof(double) * (size_t)h->m.nnz);
}
/* interp: 0 = D2, 1 = MULTIPASS; cf holds renumbered coarse ids; max_elmts <= 0: no truncation */
ORC_API orc_cla_matrix *orc_cla_interpolate(int n, const int *rp, const int *ci, const double *va, const int *cf, const unsigned char *s_con, int nc,
                                            int interp, int max_elmts)
{
    orc_cla_matrix *h = (orc_cla_matrix *)calloc(1, sizeof(orc_cla_matrix));
    if (interp == 1) cla_interp_multipass(n, rp, ci, va, cf, s_con, nc, &h->m);
    else if (interp == 2) cla_interp_d1(n, rp, ci, va, cf, s_con, nc, &h->m);
    else cla_interp_d2(n, rp, ci, va, cf, s_con, nc, &h->m);
    if (max_elmts > 0 && n > 0) cla_truncate(&h->m, max_elmts);
    return h;
}
ORC_API orc_cla_matrix *orc_cla_galerkin(int n, const int *rp, const int *ci, const double *va, const orc_cla_matrix *P)
{
    cla_csr R, AP;
    cla_transpose(&P->m, &R);
    cla_spgemm(n, rp, ci, va, P->m.rp, P->m.ci, P->m.va, P->m.nc, &AP);
    orc_cla_matrix *h = (orc_cla_matrix *)calloc(1, sizeof(orc_cla_matrix));
    cla_spgemm(R.n, R.rp, R.ci, R.va, AP.rp, AP.ci, AP.va, P->m.nc, &h->m);
    cla_csr_free(&R);
    cla_csr_free(&AP);
    return h;
}
