/*
 * amg_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of the reference's algorithms on the AMG solve-phase hot path and of
 * the setup producers that decide the integer hierarchy.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the product (libamgx_b200.so)
 * never does.  Every function cites the reference file:line it follows (paths under
 * /root/reference).  The reference's device kernels run one thread per row / entry; here each
 * "kernel" is a loop with the same per-element arithmetic, and kernels that read and write the same
 * array concurrently in the reference are evaluated with snapshot semantics (reads see the state
 * before the kernel) -- the same choice the CUDA engine makes.
 *
 * Pinning: validated against outputs of the reference itself (oracle/_ref, generated on a B200 by
 * tests/golden/make_golden.py and committed under tests/golden/) -- see tests/test_oracle_golden.py.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fopenmp amg_oracle.c -o liboracle.so -lm
 * (-ffp-contract=off: FMAs appear only where written with fma()/fmaf(), mirroring where nvcc
 * contracts the reference's device expressions.)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------- */
/* helpers                                                                                     */
/* ------------------------------------------------------------------------------------------- */
static double guard_d(double d) /* isNotCloseToZero(d) ? d : epsilon(d): include/solvers/block_common_solver.h:22-100 */
{
    return fabs(d) < 1e-12 ? copysign(1e-12, d) : d;
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
ORC_API void orc_set_num_threads(int t)
{
#ifdef _OPENMP
    omp_set_num_threads(t);
#else
    (void)t;
#endif
}

/* ------------------------------------------------------------------------------------------- */
/* SpMV: y = A x.  Per row strictly left to right, y = a*x + y as one FMA per entry -- the order */
/* of the reference's csrmv kernel (src/amgx_cusparse.cu:983-1024) and of its host loop          */
/* (src/multiply.cu:753-852).                                                                    */
/* ------------------------------------------------------------------------------------------- */
ORC_API void orc_spmv(int n, const int *rp, const int *ci, const double *va, const double *x, double *y)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int k = rp[i]; k < rp[i + 1]; k++) s = fma(va[k], x[ci[k]], s);
        y[i] = s;
    }
}

/* r = b - A x : axmb = SpMV then r*(-1) + b*1 (src/blas.cu:601-623) */
ORC_API void orc_residual(int n, const int *rp, const int *ci, const double *va, const double *x, const double *b, double *r)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int k = rp[i]; k < rp[i + 1]; k++) s = fma(va[k], x[ci[k]], s);
        r[i] = b[i] - s;
    }
}

/* Jacobi sweep: x' = x + ((b - A x) * w) * (1/d) with the exact op order of jacobi_postsmooth_functor */
/* (src/solvers/block_jacobi_solver.cu:32-50): d = 1/d; b = b - y; b = b*w; return b*d + x  (FMA).    */
/* Same arithmetic for JACOBI_L1 with d = L1 row norm (src/solvers/jacobi_l1_solver.cu:27-44).        */
ORC_API void orc_jacobi_sweep(int n, const int *rp, const int *ci, const double *va, const double *d, const double *b, const double *x,
                              double *xout, double omega)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int k = rp[i]; k < rp[i + 1]; k++) s = fma(va[k], x[ci[k]], s);
        double dinv = 1.0 / guard_d(d[i]);
        double t = b[i] - s;
        t = t * omega;
        xout[i] = fma(t, dinv, x[i]);
    }
}

/* zero initial guess: x = b*w/d (jacobi_presmooth_functor, block_jacobi_solver.cu:24-30) */
ORC_API void orc_jacobi_zero(int n, const double *d, const double *b, double *x, double omega)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) x[i] = b[i] * omega / guard_d(d[i]);
}

/* d_i = A(i,i) -- first matching column (computeDiagonalKernelCSR, src/matrix.cu:101-126) */
ORC_API void orc_diag_index(int n, const int *rp, const int *ci, int *diag)
{
    for (int i = 0; i < n; i++) {
        int d = -1;
        for (int k = rp[i]; k < rp[i + 1]; k++)
            if (ci[k] == i) { d = k; break; }
        diag[i] = d;
    }
}
ORC_API void orc_extract_diag(int n, const int *rp, const int *ci, const double *va, double *d)
{
    for (int i = 0; i < n; i++) {
        d[i] = 0.0;
        for (int k = rp[i]; k < rp[i + 1]; k++)
            if (ci[k] == i) { d[i] = va[k]; break; }
    }
}
/* L1 row norms (compute_d_kernel, src/solvers/jacobi_l1_solver.cu:60-91) */
ORC_API void orc_l1_norms(int n, const int *rp, const int *ci, const double *va, double *d)
{
    for (int i = 0; i < n; i++) {
        double acc = 0.0;
        int npd = 0;
        for (int k = rp[i]; k < rp[i + 1]; k++) {
            double a = va[k];
            if (ci[k] == i && a < 0.) npd = 1;
            acc += fabs(a);
        }
        d[i] = npd ? -acc : acc;
    }
}

/* level-1 pieces in the reference's order */
ORC_API double orc_dot(int n, const double *x, const double *y)
{
    double s = 0.0;
    for (int i = 0; i < n; i++) s += x[i] * y[i];
    return s;
}
ORC_API double orc_nrm2(int n, const double *x) { return sqrt(orc_dot(n, x, x)); }
ORC_API double orc_nrm1(int n, const double *x)
{
    double s = 0.0;
    for (int i = 0; i < n; i++) s += fabs(x[i]);
    return s;
}
ORC_API double orc_nrmmax(int n, const double *x)
{
    double s = 0.0;
    for (int i = 0; i < n; i++) s = fmax(s, fabs(x[i]));
    return s;
}

/* ------------------------------------------------------------------------------------------- */
/* SIZE_2 aggregation                                                                          */
/* ------------------------------------------------------------------------------------------- */
static unsigned hash_val(unsigned a, unsigned seed) /* include/aggregation/selectors/common_selector.h:19-29 */
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

/* computeEdgeWeightsBlockDiaCsr_V2 (common_selector.h:63-137), WeightType = float, weight_formula 0/1 */
ORC_API void orc_edge_weights(int n, const int *rp, const int *ci, const double *va, int weight_formula, float *w)
{
    int *diag = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    orc_diag_index(n, rp, ci, diag);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        double aii = diag[i] >= 0 ? va[diag[i]] : 0.0;
        for (int k = rp[i]; k < rp[i + 1]; k++) {
            int j = ci[k];
            if (i == j || j >= n) { w[k] = -1.0f; continue; }
            double ajj = diag[j] >= 0 ? va[diag[j]] : 0.0;
            double mx = fabs(aii) > fabs(ajj) ? fabs(aii) : fabs(ajj);
            float den = (float)mx;
            double kval = 0.0;
            int found = 0;
            for (int kk = rp[j]; kk < rp[j + 1]; kk++)
                if (ci[kk] == i) { kval = va[kk]; found = 1; break; }
            float ew = 0.0f;
            if (found) {
                if (weight_formula == 0) {
                    double ssum = fabs(va[k]) + fabs(kval);
                    double t = 0.5 * ssum;
                    ew = (float)(t / (double)den);
                } else {
                    double rz = va[k] / aii + kval / ajj;
                    ew = (float)(-0.5 * (double)(float)rz);
                }
            }
            unsigned lo = (unsigned)(i < j ? i : j), hi = (unsigned)(i < j ? j : i);
            unsigned h = hash_val(lo, hi);
            float small_fraction = (1e-5f * (float)h) / 4294967296.0f; /* scaling_factor<float>() * hash / (float)UINT_MAX */
            ew = fmaf(small_fraction, ew, ew);                         /* ed_weight += small_fraction * ed_weight (contracted) */
            w[k] = ew;
        }
    }
    free(diag);
}

/* SIZE_4: Size4Selector::setAggregates_common_sqblock (src/aggregation/selectors/size4_selector.cu:96-224) with the kernels of
 * include/aggregation/selectors/common_selector.h:178-525, deterministic flow: handshake pairs (findStrongestNeighbour_NoMerge +
 * matchEdges), handshake pairs of pairs (_StoreWeight + agreeOnProposal + matchAggregates), candidate/join merge of the leftovers,
 * renumbering.  No kernel of this flow reads what another thread of the same launch writes, so the loops below may update in
 * place.  Returns #aggregates.  parity unpinned (no reference golden yet). */
static int s4_count(int n, const int *v) { int c = 0; for (int i = 0; i < n; i++) c += (v[i] == -1); return c; }
ORC_API int orc_size4_aggregates(int n, const int *rp, const int *ci, const double *va, int max_iterations, double max_unassigned, int weight_formula, int *agg)
{
    if (n == 0) return 0;
    float *w = (float *)malloc(sizeof(float) * (size_t)(rp[n] > 0 ? rp[n] : 1)), *wsn = (float *)malloc(sizeof(float) * (size_t)n);
    int *strongest = (int *)malloc(sizeof(int) * (size_t)n), *partner = (int *)malloc(sizeof(int) * (size_t)n);
    int *aggregated = (int *)malloc(sizeof(int) * (size_t)n), *cand = (int *)malloc(sizeof(int) * (size_t)n);
    orc_edge_weights(n, rp, ci, va, weight_formula, w);
    for (int i = 0; i < n; i++) { agg[i] = i; strongest[i] = -1; partner[i] = -1; }
    int unassigned = n, prev = n, icount = 0;
    do {                                                   /* ---- pairs ---- */
        for (int t = 0; t < n; t++) {                      /* findStrongestNeighbourBlockDiaCsr_NoMerge */
            if (partner[t] != -1) continue;
            float mw = 0.f; int best = -1;
            for (int j = rp[t]; j < rp[t + 1]; j++) {
                const int jc = ci[j];
                if (t == jc || jc >= n) continue;
                if (partner[jc] == -1 && (w[j] > mw || (w[j] == mw && jc > best))) { mw = w[j]; best = jc; }
            }
            if (best != -1) strongest[t] = best;
        }
        for (int t = 0; t < n; t++) {                      /* matchEdges */
            if (partner[t] != -1) continue;
            const int pm = strongest[t];
            if (pm != -1 && strongest[pm] == t) { partner[t] = pm; agg[t] = (pm > t ? t : pm); }     /* each thread writes its own entries only */
        }
        prev = unassigned;
        unassigned = s4_count(n, partner) + 2 * n;         /* the reference counts over its 3n-entry partner_index array */
        icount++;
    } while (!(unassigned == 0 || icount > max_iterations || 1.0 * unassigned / n < max_unassigned || prev == unassigned));
    for (int t = 0; t < n; t++) if (partner[t] == -1) partner[t] = t;      /* assignUnassignedVertices */
    for (int t = 0; t < n; t++) { wsn[t] = -1.f; aggregated[t] = -1; }
    icount = 0; unassigned = prev = n;
    do {                                                   /* ---- pairs of pairs ---- */
        for (int t = 0; t < n; t++) {                      /* findStrongestNeighbourBlockDiaCsr_StoreWeight */
            if (aggregated[t] != -1) continue;
            const int p = partner[t];
            float mw = 0.f; int best = -1;
            for (int j = rp[t]; j < rp[t + 1]; j++) {
                const int jc = ci[j];
                if (t == jc || jc >= n) continue;
                if (aggregated[jc] == -1 && jc != p && (w[j] > mw || (w[j] == mw && jc > best))) { mw = w[j]; best = jc; }
            }
            if (best != -1) { wsn[t] = mw; strongest[t] = agg[best]; }
        }
        {                                                  /* agreeOnProposal: decisions from the launch's inputs */
            int *sn_in = (int *)malloc(sizeof(int) * (size_t)n);
            memcpy(sn_in, strongest, sizeof(int) * (size_t)n);
            for (int t = 0; t < n; t++) {
                if (aggregated[t] != -1) continue;
                const int p = partner[t];
                const float mine = wsn[t], theirs = (p != -1) ? wsn[p] : -1.f;
                if (mine < 0.f && theirs < 0.f) { aggregated[t] = 1; strongest[t] = -1; }
                else if (mine < theirs) strongest[t] = sn_in[p];
            }
            free(sn_in);
        }
        {                                                  /* matchAggregates */
            int *ag_in = (int *)malloc(sizeof(int) * (size_t)n);
            memcpy(ag_in, agg, sizeof(int) * (size_t)n);
            for (int t = 0; t < n; t++) {
                if (aggregated[t] != -1) continue;
                const int pm = strongest[t];
                if (pm == -1) continue;
                const int mine = ag_in[t];
                if (strongest[pm] == mine) { aggregated[t] = 1; agg[t] = pm > mine ? mine : pm; }
            }
            free(ag_in);
        }
        prev = unassigned;
        unassigned = s4_count(n, aggregated);
        icount++;
    } while (!(unassigned == 0 || icount > max_iterations || 1.0 * unassigned / n < max_unassigned || prev == unassigned));
    for (int t = 0; t < n; t++) cand[t] = -1;
    while (unassigned != 0) {                              /* mergeWithExistingAggregatesBlockDiaCsr (deterministic) + joinExistingAggregates */
        for (int t = 0; t < n; t++) {
            if (aggregated[t] != -1) continue;
            float mw = 0.f; int best = -1;
            for (int j = rp[t]; j < rp[t + 1]; j++) {
                const int jc = ci[j];
                if (t == jc || jc >= n) continue;
                if (aggregated[jc] != -1 && (w[j] > mw || (w[j] == mw && jc > best))) { mw = w[j]; best = jc; }
            }
            cand[t] = best != -1 ? agg[best] : t;
        }
        for (int t = 0; t < n; t++) if (aggregated[t] == -1 && cand[t] != -1) { agg[t] = cand[t]; aggregated[t] = 1; }
        unassigned = s4_count(n, aggregated);
    }
    /* renumberAndCountAggregates */
    int *mark = (int *)calloc((size_t)n + 1, sizeof(int));
    for (int i = 0; i < n; i++) mark[agg[i]] = 1;
    int nagg = 0;
    for (int i = 0; i <= n; i++) { const int m = mark[i]; mark[i] = nagg; nagg += m; }
    for (int i = 0; i < n; i++) agg[i] = mark[agg[i]];
    free(mark); free(w); free(wsn); free(strongest); free(partner); free(aggregated); free(cand);
    return nagg;
}
static int g_agg_selector = 2;       /* 2 SIZE_2, 4 SIZE_4: selector of the NEXT aggregation setups */
ORC_API void orc_set_aggregation_selector(int size) { g_agg_selector = size; }

/* setAggregates_common_sqblocks (src/aggregation/selectors/size2_selector.cu:736-890), one-phase handshake,
 * deterministic leftover merge, then renumberAndCountAggregates (agg_selector.cu:18-43).  Returns #aggregates. */
ORC_API int orc_size2_aggregates(int n, const int *rp, const int *ci, const double *va, int max_iterations, double max_unassigned,
                                 int merge_singletons, int weight_formula, int *agg)
{
    if (n == 0) return 0;
    float *w = (float *)malloc(sizeof(float) * (size_t)(rp[n] > 0 ? rp[n] : 1));
    int *strongest = (int *)malloc(sizeof(int) * (size_t)n), *merge_to = (int *)malloc(sizeof(int) * (size_t)n);
    int *cand = (int *)malloc(sizeof(int) * (size_t)n);
    orc_edge_weights(n, rp, ci, va, weight_formula, w);
    for (int i = 0; i < n; i++) { agg[i] = -1; strongest[i] = -1; }
    /* The reference compiles this loop with EXPERIMENTAL_ITERATIVE_MATCHING (size2_selector.cu:22, 808-847): the
     * unaggregated count is taken after EVEN iterations only and consumed one iteration later, so the loop always
     * runs an even number of handshake steps and its exit test sees a count that is one step stale. */
    int unassigned = n, prev = n, icount = 0, s = 1, pending = n;
    do {
        /* findStrongestNeighbourBlockDiaCsr_V2 (size2_selector.cu:224-301), phase 1, snapshot reads of agg[] */
#pragma omp parallel for schedule(static)
        for (int t = 0; t < n; t++) {
            merge_to[t] = -1;
            if (agg[t] != -1) continue;
            int s_un = -1, s_ag = -1;
            float m_un = 0.f, m_ag = 0.f;
            for (int k = rp[t]; k < rp[t + 1]; k++) {
                int j = ci[k];
                float wt = w[k];
                if (j == t || j >= n) continue;
                if (agg[j] == -1 && (wt > m_un || (wt == m_un && j > s_un))) { m_un = wt; s_un = j; }
                else if (agg[j] != -1 && (wt > m_ag || (wt == m_ag && j > s_ag))) { m_ag = wt; s_ag = j; }
            }
            if (s_un == -1 && s_ag != -1) merge_to[t] = merge_singletons ? agg[s_ag] : t;
            else if (s_un != -1) strongest[t] = s_un;
            else strongest[t] = t;
        }
        /* matchEdges (size2_selector.cu:302-323) */
#pragma omp parallel for schedule(static)
        for (int t = 0; t < n; t++) {
            if (agg[t] != -1) continue;
            if (merge_to[t] != -1) { cand[t] = merge_to[t]; continue; }
            cand[t] = -1;
            int pm = strongest[t];
            if (pm < 0) continue;
            if (merge_to[pm] == -1 && strongest[pm] == t) cand[t] = (pm > t) ? t : pm;
        }
        int now = 0;
        for (int t = 0; t < n; t++) {
            if (agg[t] == -1 && cand[t] != -1) agg[t] = cand[t];
            if (agg[t] == -1) now++;
        }
        s = (icount & 1);
        if (s == 0) pending = now;
        else { prev = unassigned; unassigned = pending; }
        icount++;
    } while (s == 0 || !(unassigned == 0 || icount > max_iterations || 1.0 * unassigned / n < max_unassigned || unassigned == prev));

    if (merge_singletons) {
        /* mergeWithExistingAggregatesBlockDiaCsr_V2 + joinExistingAggregates, deterministic path (:508-566, 424-440) */
        for (int t = 0; t < n; t++) cand[t] = -1;
        while (unassigned != 0) {
#pragma omp parallel for schedule(static)
            for (int t = 0; t < n; t++) {
                if (agg[t] != -1) continue;
                float m_ag = 0.f;
                int s_ag = -1;
                for (int k = rp[t]; k < rp[t + 1]; k++) {
                    float wt = w[k];
                    int j = ci[k];
                    if (j == t || j >= n) continue;
                    if (agg[j] != -1 && (wt > m_ag || (wt == m_ag && j > s_ag))) { m_ag = wt; s_ag = j; }
                }
                cand[t] = (s_ag != -1) ? agg[s_ag] : t;
            }
            unassigned = 0;
            for (int t = 0; t < n; t++) {
                if (agg[t] == -1 && cand[t] != -1) agg[t] = cand[t];
                if (agg[t] == -1) unassigned++;
            }
        }
    } else {
        for (int t = 0; t < n; t++)
            if (agg[t] == -1) agg[t] = t;
    }
    /* renumber: mark used labels, exclusive scan, relabel */
    int *scratch = (int *)calloc((size_t)n + 1, sizeof(int));
    for (int t = 0; t < n; t++) scratch[agg[t]] = 1;
    int run = 0;
    for (int t = 0; t <= n; t++) { int v = scratch[t]; scratch[t] = run; run += v; }
    for (int t = 0; t < n; t++) agg[t] = scratch[agg[t]];
    int nagg = scratch[n];
    free(scratch); free(w); free(strongest); free(merge_to); free(cand);
    return nagg;
}

/* computeRestrictionOperator_common (src/aggregation/aggregation_amg_level.cu:237-299): stable sort of rows by aggregate */
ORC_API void orc_restriction(int n, int nagg, const int *agg, int *Rp, int *Rc)
{
    for (int I = 0; I <= nagg; I++) Rp[I] = 0;
    for (int i = 0; i < n; i++) Rp[agg[i] + 1]++;
    for (int I = 0; I < nagg; I++) Rp[I + 1] += Rp[I];
    int *pos = (int *)malloc(sizeof(int) * (size_t)(nagg > 0 ? nagg : 1));
    for (int I = 0; I < nagg; I++) pos[I] = Rp[I];
    for (int i = 0; i < n; i++) Rc[pos[agg[i]]++] = i;
    free(pos);
}

/* Galerkin product for piecewise-constant P: Ac(I,J) = sum_{i in I} sum_{j in J} a_ij
 * (LowDegCoarseAGenerator::computeAOperator, low_deg_coarse_A_generator.cu:1135-1320).  Entries of one coarse
 * coefficient are summed fine-row ascending, in-row order; coarse rows are emitted with ascending columns
 * (the reference emits hash-table order -- compare after sorting).  Two calls: counts, then fill. */
ORC_API int orc_galerkin_count(int n, const int *rp, const int *ci, const int *agg, int nagg, const int *Rp, const int *Rc, int *rpc)
{
    int *mark = (int *)malloc(sizeof(int) * (size_t)(nagg > 0 ? nagg : 1));
    for (int I = 0; I < nagg; I++) mark[I] = -1;
    int total = 0;
    rpc[0] = 0;
    for (int I = 0; I < nagg; I++) {
        int cnt = 0;
        for (int q = Rp[I]; q < Rp[I + 1]; q++) {
            int i = Rc[q];
            for (int k = rp[i]; k < rp[i + 1]; k++) {
                int J = agg[ci[k]];
                if (mark[J] != I) { mark[J] = I; cnt++; }
            }
        }
        total += cnt;
        rpc[I + 1] = total;
    }
    free(mark);
    (void)n;
    return total;
}
static int cmp_int(const void *a, const void *b) { return (*(const int *)a > *(const int *)b) - (*(const int *)a < *(const int *)b); }
ORC_API void orc_galerkin_fill(int n, const int *rp, const int *ci, const double *va, const int *agg, int nagg, const int *Rp, const int *Rc,
                               const int *rpc, int *cic, double *vac)
{
    int *slot = (int *)malloc(sizeof(int) * (size_t)(nagg > 0 ? nagg : 1));
    for (int I = 0; I < nagg; I++) slot[I] = -1;
    (void)n;
    for (int I = 0; I < nagg; I++) {
        int base = rpc[I], cnt = 0;
        /* collect distinct coarse columns */
        for (int q = Rp[I]; q < Rp[I + 1]; q++) {
            int i = Rc[q];
            for (int k = rp[i]; k < rp[i + 1]; k++) {
                int J = agg[ci[k]];
                if (slot[J] < base || slot[J] >= base + cnt || cic[slot[J]] != J) { cic[base + cnt] = J; slot[J] = base + cnt; cnt++; }
            }
        }
        qsort(cic + base, (size_t)cnt, sizeof(int), cmp_int);
        for (int p = 0; p < cnt; p++) { slot[cic[base + p]] = base + p; vac[base + p] = 0.0; }
        int *first = (int *)calloc((size_t)(cnt > 0 ? cnt : 1), sizeof(int));
        for (int q = Rp[I]; q < Rp[I + 1]; q++) {
            int i = Rc[q];
            for (int k = rp[i]; k < rp[i + 1]; k++) {
                int p = slot[agg[ci[k]]];
                if (!first[p - base]) { vac[p] = va[k]; first[p - base] = 1; }
                else vac[p] = vac[p] + va[k];
            }
        }
        free(first);
    }
    free(slot);
}

/* restrictResidualKernel / prolongateAndApplyCorrectionKernel (aggregation_amg_level.cu:91-111, 156-181) */
ORC_API void orc_restrict(int nagg, const int *Rp, const int *Rc, const double *r, double *rc)
{
#pragma omp parallel for schedule(static)
    for (int I = 0; I < nagg; I++) {
        double t = 0.0;
        for (int j = Rp[I]; j < Rp[I + 1]; j++) t = t + r[Rc[j]];
        rc[I] = t;
    }
}
ORC_API void orc_prolong_add(int n, const int *agg, const double *e, double *x)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) x[i] = x[i] + e[agg[i]];
}

/* ------------------------------------------------------------------------------------------- */
/* hierarchy + V-cycle + PCG, composed exactly like the reference (unfused)                    */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
    int n, nnz, nagg, coarsest;
    int *rp, *ci, *agg, *Rp, *Rc;
    double *va, *d;          /* d: smoother diagonal (Jacobi: a_ii, L1: row norms) */
    double *bc, *xc, *r, *tmp;
    int own;
    /* MULTICOLOR_DILU smoother data */
    int num_colors, *colors, *sorted_rows, *color_offsets;
    double *Einv, *delta, *Delta;
    /* classical AMG: explicit P (n x nagg) and R = P^T, C/F map of the level */
    int *Pp, *Pc, *Rtp, *Rtc, *cf;
    double *Pv, *Rtv;
    /* DENSE_LU_SOLVER on the coarsest level: column-major LU factors + pivots */
    double *lu;
    int *ipiv;
    /* Chebyshev smoothers: spectrum bounds, work vectors, damped-root step lengths */
    double lmax, lmin, *cheb_p, *cheb_z, *cheb_r, tau[10];
    /* error scaling: last computed scale and how many more corrections reuse it */
    double scale;
    int scale_counter;
} orc_level;

typedef struct {
    int num_levels;
    orc_level *lv;
    int presweeps, postsweeps, coarsest_sweeps, finest_sweeps, smoother; /* smoother: 0 BLOCK_JACOBI, 1 JACOBI_L1, 2 MULTICOLOR_DILU, 3 MULTICOLOR_GS */
    int symmetric_gs;        /* MULTICOLOR_GS: symmetric_GS = 1 sweeps the colours up, then down */
    /* smoother 4 CHEBYSHEV (src/solvers/cheb_solver.cu), 5 CHEBYSHEV_POLY (src/solvers/chebyshev_poly.cu) */
    int cheb_order, cheb_mode, cheb_precond;   /* cheb_precond: 0 none, 1 BLOCK_JACOBI, 2 JACOBI_L1 (one zero-guess sweep, weight cheb_inner_omega) */
    double cheb_inner_omega, cheb_user_max, cheb_user_min;
    double omega, uncolored_fraction;
    int dense_lu;            /* coarse_solver = DENSE_LU_SOLVER */
    int cycle;               /* 0 V, 1 W, 2 F, 3 CG, 4 CGF (src/cycles/{v,w,f,cg,cg_flex}_cycle.cu) */
    int cycle_iters;         /* CG / CGF cycles: CG iterations per visit of a level (default 2, src/core.cu:435) */
    /* error_scaling = 2, 3 on aggregation levels (src/aggregation/aggregation_amg_level.cu:700-824) */
    int error_scaling, scaling_smoother_steps, reuse_scale;
} orc_amg;

#include "classical_oracle.inc.c"

ORC_API int orc_color_min_max(int n, const int *rp, const int *ci, double max_uncolored_fraction, int *colors);
ORC_API int orc_color_parallel_greedy(int n, const int *rp, const int *ci, double max_uncolored_fraction, int *colors);
static int g_coloring_scheme;
ORC_API void orc_color_arrays(int n, int num_colors, const int *colors, int *sorted_rows, int *offsets);
ORC_API void orc_dilu_setup_1x1(int n, const int *rp, const int *ci, const double *va, int num_colors, const int *colors, const int *sorted_rows,
                                const int *offsets, double *Einv);
ORC_API void orc_dilu_sweep_1x1(int n, const int *rp, const int *ci, const double *va, int num_colors, const int *colors, const int *sorted_rows,
                                const int *offsets, const double *Einv, const double *b, double *x, double weight, double *delta, double *Delta);
static double g_uncolored_fraction = 0.15;
ORC_API void orc_set_uncolored_fraction(double f) { g_uncolored_fraction = f; }

static int g_cheb_precond = 0;       /* read by level_smoother_setup for smoother 4: which diagonal the inner Jacobi uses */
static void level_smoother_setup(orc_level *L, int smoother)
{
    L->d = (double *)malloc(sizeof(double) * (size_t)(L->n > 0 ? L->n : 1));
    if (smoother == 1 || (smoother == 4 && g_cheb_precond == 2)) orc_l1_norms(L->n, L->rp, L->ci, L->va, L->d);
    else orc_extract_diag(L->n, L->rp, L->ci, L->va, L->d);
    if (smoother == 4 || smoother == 5) {
        const size_t nn = (size_t)(L->n > 0 ? L->n : 1);
        L->cheb_p = (double *)calloc(nn, sizeof(double));
        L->cheb_z = (double *)calloc(nn, sizeof(double));
        L->cheb_r = (double *)calloc(nn, sizeof(double));
        double lam = 0.0;                 /* getLambdaEstimate + max_element: max_i sum_j |a_ij| */
        for (int i = 0; i < L->n; i++) {
            double cur = 0.0;
            for (int k = L->rp[i]; k < L->rp[i + 1]; k++) cur += fabs(L->va[k]);
            if (cur > lam) lam = cur;
        }
        L->lmax = lam;                    /* the preconditioned variants are filled in by orc_amg_set_chebyshev */
        L->lmin = lam * 0.125;
    }
    if (smoother == 2 || smoother == 3) {
        const size_t nn = (size_t)(L->n > 0 ? L->n : 1);
        L->colors = (int *)malloc(sizeof(int) * nn);
        L->sorted_rows = (int *)malloc(sizeof(int) * nn);
        L->num_colors = g_coloring_scheme == 1 ? orc_color_parallel_greedy(L->n, L->rp, L->ci, g_uncolored_fraction, L->colors)
                                               : orc_color_min_max(L->n, L->rp, L->ci, g_uncolored_fraction, L->colors);
        L->color_offsets = (int *)malloc(sizeof(int) * ((size_t)L->num_colors + 1));
        orc_color_arrays(L->n, L->num_colors, L->colors, L->sorted_rows, L->color_offsets);
        L->Einv = (double *)calloc(nn, sizeof(double));
        L->delta = (double *)calloc(nn, sizeof(double));
        L->Delta = (double *)calloc(nn, sizeof(double));
        if (smoother == 2) orc_dilu_setup_1x1(L->n, L->rp, L->ci, L->va, L->num_colors, L->colors, L->sorted_rows, L->color_offsets, L->Einv);
    }
}

/* AMGX_solver_resetup with structure_reuse_levels: the NEXT orc_amg_setup takes the aggregates of `from` for every coarsening
 * whose 1-based level index is < reuse_levels (all of them for -1), src/amg.cu:229-272; cleared by that setup. */
static const orc_amg *g_reuse_from = NULL;
static int g_reuse_levels = 0;
ORC_API void orc_amg_reuse_structure(const orc_amg *from, int reuse_levels) { g_reuse_from = from; g_reuse_levels = reuse_levels; }

ORC_API orc_amg *orc_amg_setup(int n, const int *rp, const int *ci, const double *va, int max_levels, int min_coarse_rows, double coarsen_threshold,
                               int presweeps, int postsweeps, int coarsest_sweeps, int finest_sweeps, int smoother, double omega,
                               int max_iterations, double max_unassigned, int merge_singletons, int weight_formula)
{
    /* AMG_Setup::setup level loop (src/amg.cu:201-418), single partition, coarse_solver = NOSOLVER */
    orc_amg *a = (orc_amg *)calloc(1, sizeof(orc_amg));
    a->cycle_iters = 2; a->scaling_smoother_steps = 2;
    a->lv = (orc_level *)calloc((size_t)(max_levels > 0 ? max_levels : 1) + 1, sizeof(orc_level));
    a->presweeps = presweeps; a->postsweeps = postsweeps; a->coarsest_sweeps = coarsest_sweeps; a->finest_sweeps = finest_sweeps;
    a->smoother = smoother; a->omega = omega;
    orc_level *L = &a->lv[0];
    L->n = n; L->nnz = rp[n]; L->rp = (int *)rp; L->ci = (int *)ci; L->va = (double *)va; L->own = 0;
    int num_levels = 1;
    for (;;) {
        L = &a->lv[num_levels - 1];
        level_smoother_setup(L, smoother);
        L->tmp = (double *)malloc(sizeof(double) * (size_t)(L->n > 0 ? L->n : 1));
        if (num_levels >= max_levels || L->n <= min_coarse_rows) { L->coarsest = 1; break; }
        L->agg = (int *)malloc(sizeof(int) * (size_t)L->n);
        int nagg;
        if (g_reuse_from && g_reuse_levels != 0 && (g_reuse_levels == -1 || g_reuse_levels > num_levels) && num_levels < g_reuse_from->num_levels &&
            g_reuse_from->lv[num_levels - 1].agg && g_reuse_from->lv[num_levels - 1].n == L->n) {
            memcpy(L->agg, g_reuse_from->lv[num_levels - 1].agg, sizeof(int) * (size_t)L->n);
            nagg = g_reuse_from->lv[num_levels - 1].nagg;
        } else {
            g_reuse_levels = 0;   /* the chain of reused levels ends at the first rebuilt one */
            nagg = g_agg_selector == 4 ? orc_size4_aggregates(L->n, L->rp, L->ci, L->va, max_iterations, max_unassigned, weight_formula, L->agg)
                                       : orc_size2_aggregates(L->n, L->rp, L->ci, L->va, max_iterations, max_unassigned, merge_singletons, weight_formula, L->agg);
        }
        if ((double)nagg <= coarsen_threshold * (double)L->n && nagg != L->n && nagg >= min_coarse_rows) {
            L->nagg = nagg;
            L->Rp = (int *)malloc(sizeof(int) * ((size_t)nagg + 1));
            L->Rc = (int *)malloc(sizeof(int) * (size_t)L->n);
            orc_restriction(L->n, nagg, L->agg, L->Rp, L->Rc);
            orc_level *N = &a->lv[num_levels];
            N->n = nagg;
            N->rp = (int *)malloc(sizeof(int) * ((size_t)nagg + 1));
            N->nnz = orc_galerkin_count(L->n, L->rp, L->ci, L->agg, nagg, L->Rp, L->Rc, N->rp);
            N->ci = (int *)malloc(sizeof(int) * (size_t)(N->nnz > 0 ? N->nnz : 1));
            N->va = (double *)malloc(sizeof(double) * (size_t)(N->nnz > 0 ? N->nnz : 1));
            orc_galerkin_fill(L->n, L->rp, L->ci, L->va, L->agg, nagg, L->Rp, L->Rc, N->rp, N->ci, N->va);
            N->own = 1;
            L->bc = (double *)calloc((size_t)nagg, sizeof(double));
            L->xc = (double *)calloc((size_t)nagg, sizeof(double));
            L->r = (double *)calloc((size_t)L->n, sizeof(double));
            num_levels++;
        } else {
            free(L->agg); L->agg = NULL; L->coarsest = 1;
            break;
        }
    }
    a->num_levels = num_levels;
    g_reuse_from = NULL; g_reuse_levels = 0;
    return a;
}


/* AMG_Setup::setup level loop with Classical_AMG_Level::createCoarseVertices / createCoarseMatrices
 * (src/classical/classical_amg_level.cu:213-299, 344-435).  interp / aggressive_interp: 0 = D2, 1 = MULTIPASS. */
ORC_API orc_amg *orc_amg_setup_classical(int n, const int *rp, const int *ci, const double *va, int max_levels, int min_coarse_rows, double coarsen_threshold,
                                         int presweeps, int postsweeps, int coarsest_sweeps, int finest_sweeps, int smoother, double omega,
                                         double strength_threshold, double max_row_sum, int interp, int aggressive_levels, int aggressive_interp,
                                         int max_elmts)
{
    orc_amg *a = (orc_amg *)calloc(1, sizeof(orc_amg));
    a->cycle_iters = 2; a->scaling_smoother_steps = 2;
    a->lv = (orc_level *)calloc((size_t)(max_levels > 0 ? max_levels : 1) + 1, sizeof(orc_level));
    a->presweeps = presweeps; a->postsweeps = postsweeps; a->coarsest_sweeps = coarsest_sweeps; a->finest_sweeps = finest_sweeps;
    a->smoother = smoother; a->omega = omega;
    orc_level *L = &a->lv[0];
    L->n = n; L->nnz = rp[n]; L->rp = (int *)rp; L->ci = (int *)ci; L->va = (double *)va; L->own = 0;
    int num_levels = 1;
    for (;;) {
        L = &a->lv[num_levels - 1];
        level_smoother_setup(L, smoother);
        L->tmp = (double *)malloc(sizeof(double) * (size_t)(L->n > 0 ? L->n : 1));
        if (num_levels >= max_levels || L->n <= min_coarse_rows) { L->coarsest = 1; break; }
        const int lvl = num_levels - 1;
        /* AMGX_solver_resetup with structure_reuse_levels (src/amg.cu:229-272, classical_amg_level.cu:274-291): a reused level keeps P and R --
         * pattern and values -- and only A_c = R A P follows the new matrix */
        if (g_reuse_from && g_reuse_levels != 0 && (g_reuse_levels == -1 || g_reuse_levels > num_levels) && num_levels < g_reuse_from->num_levels &&
            g_reuse_from->lv[lvl].Pp && g_reuse_from->lv[lvl].n == L->n) {
            const orc_level *O = &g_reuse_from->lv[lvl];
            const int nc = O->nagg, pn = O->Pp[O->n];
            cla_csr P, R, AP, Ac;
            P.n = L->n; P.nc = nc; P.nnz = pn;
            P.rp = (int *)malloc(sizeof(int) * ((size_t)L->n + 1)); memcpy(P.rp, O->Pp, sizeof(int) * ((size_t)L->n + 1));
            P.ci = (int *)malloc(sizeof(int) * (size_t)(pn > 0 ? pn : 1)); memcpy(P.ci, O->Pc, sizeof(int) * (size_t)pn);
            P.va = (double *)malloc(sizeof(double) * (size_t)(pn > 0 ? pn : 1)); memcpy(P.va, O->Pv, sizeof(double) * (size_t)pn);
            R.n = nc; R.nc = L->n; R.nnz = pn;
            R.rp = (int *)malloc(sizeof(int) * ((size_t)nc + 1)); memcpy(R.rp, O->Rtp, sizeof(int) * ((size_t)nc + 1));
            R.ci = (int *)malloc(sizeof(int) * (size_t)(pn > 0 ? pn : 1)); memcpy(R.ci, O->Rtc, sizeof(int) * (size_t)pn);
            R.va = (double *)malloc(sizeof(double) * (size_t)(pn > 0 ? pn : 1)); memcpy(R.va, O->Rtv, sizeof(double) * (size_t)pn);
            int *cf = (int *)malloc(sizeof(int) * (size_t)(L->n > 0 ? L->n : 1));
            memcpy(cf, O->cf, sizeof(int) * (size_t)L->n);
            cla_spgemm(L->n, L->rp, L->ci, L->va, P.rp, P.ci, P.va, nc, &AP);
            cla_spgemm(nc, R.rp, R.ci, R.va, AP.rp, AP.ci, AP.va, nc, &Ac);
            cla_csr_free(&AP);
            L->nagg = nc;
            L->Pp = P.rp; L->Pc = P.ci; L->Pv = P.va;
            L->Rtp = R.rp; L->Rtc = R.ci; L->Rtv = R.va;
            L->cf = cf;
            orc_level *N = &a->lv[num_levels];
            N->n = nc; N->nnz = Ac.nnz; N->rp = Ac.rp; N->ci = Ac.ci; N->va = Ac.va; N->own = 1;
            L->bc = (double *)calloc((size_t)nc, sizeof(double));
            L->xc = (double *)calloc((size_t)nc, sizeof(double));
            L->r = (double *)calloc((size_t)L->n, sizeof(double));
            num_levels++;
            continue;
        }
        g_reuse_levels = 0;   /* the chain of reused levels ends at the first rebuilt one */
        unsigned char *s_con = (unsigned char *)calloc((size_t)(L->nnz > 0 ? L->nnz : 1), 1);
        float *w = (float *)calloc((size_t)(L->n > 0 ? L->n : 1), sizeof(float));
        int *cf = (int *)calloc((size_t)(L->n > 0 ? L->n : 1), sizeof(int));
        orc_cla_strength(L->n, L->rp, L->ci, L->va, strength_threshold, max_row_sum, s_con, w);
        if (lvl < aggressive_levels) orc_cla_aggressive_pmis(L->n, L->rp, L->ci, s_con, w, cf);
        else if (g_cla_selector == 1) orc_cla_hmis(L->n, L->rp, L->ci, s_con, w, cf);
        else orc_cla_pmis(L->n, L->rp, L->ci, s_con, w, cf, 0);
        const int nc = orc_cla_renumber(L->n, cf);
        free(w);
        if ((double)nc <= coarsen_threshold * (double)L->n && nc != L->n && nc >= min_coarse_rows) {
            cla_csr P, R, AP, Ac;
            const int which = (lvl < aggressive_levels) ? aggressive_interp : interp;
            if (which == 1) cla_interp_multipass(L->n, L->rp, L->ci, L->va, cf, s_con, nc, &P);
            else if (which == 2) cla_interp_d1(L->n, L->rp, L->ci, L->va, cf, s_con, nc, &P);
            else cla_interp_d2(L->n, L->rp, L->ci, L->va, cf, s_con, nc, &P);
            if (max_elmts > 0 && L->n > 0) cla_truncate(&P, max_elmts);
            cla_transpose(&P, &R);
            cla_spgemm(L->n, L->rp, L->ci, L->va, P.rp, P.ci, P.va, nc, &AP);
            cla_spgemm(nc, R.rp, R.ci, R.va, AP.rp, AP.ci, AP.va, nc, &Ac);
            cla_csr_free(&AP);
            L->nagg = nc;
            L->Pp = P.rp; L->Pc = P.ci; L->Pv = P.va;
            L->Rtp = R.rp; L->Rtc = R.ci; L->Rtv = R.va;
            L->cf = cf;
            orc_level *N = &a->lv[num_levels];
            N->n = nc; N->nnz = Ac.nnz; N->rp = Ac.rp; N->ci = Ac.ci; N->va = Ac.va; N->own = 1;
            L->bc = (double *)calloc((size_t)nc, sizeof(double));
            L->xc = (double *)calloc((size_t)nc, sizeof(double));
            L->r = (double *)calloc((size_t)L->n, sizeof(double));
            free(s_con);
            num_levels++;
        } else {
            free(s_con); free(cf);
            L->coarsest = 1;
            break;
        }
    }
    a->num_levels = num_levels;
    g_reuse_from = NULL; g_reuse_levels = 0;
    return a;
}
ORC_API int orc_amg_level_classical(const orc_amg *a, int l, int *cf, int *Pp, int *Pc, double *Pv, int *pnnz)
{
    const orc_level *L = &a->lv[l];
    if (!L->Pp) return 0;
    if (pnnz) *pnnz = L->Pp[L->n];
    if (cf) memcpy(cf, L->cf, sizeof(int) * (size_t)L->n);
    if (Pp) memcpy(Pp, L->Pp, sizeof(int) * ((size_t)L->n + 1));
    if (Pc) memcpy(Pc, L->Pc, sizeof(int) * (size_t)L->Pp[L->n]);
    if (Pv) memcpy(Pv, L->Pv, sizeof(double) * (size_t)L->Pp[L->n]);
    return 1;
}

ORC_API void orc_amg_free(orc_amg *a)
{
    if (!a) return;
    for (int l = 0; l < a->num_levels; l++) {
        orc_level *L = &a->lv[l];
        if (L->own) { free(L->rp); free(L->ci); free(L->va); }
        free(L->agg); free(L->Rp); free(L->Rc); free(L->d); free(L->bc); free(L->xc); free(L->r); free(L->tmp);
        free(L->colors); free(L->sorted_rows); free(L->color_offsets); free(L->Einv); free(L->delta); free(L->Delta);
        free(L->Pp); free(L->Pc); free(L->Pv); free(L->Rtp); free(L->Rtc); free(L->Rtv); free(L->cf);
        free(L->lu); free(L->ipiv);
    }
    free(a->lv);
    free(a);
}

ORC_API int orc_amg_num_levels(const orc_amg *a) { return a->num_levels; }
ORC_API void orc_amg_level_sizes(const orc_amg *a, int l, int *n, int *nnz, int *nagg)
{
    *n = a->lv[l].n; *nnz = a->lv[l].nnz; *nagg = a->lv[l].nagg;
}
ORC_API int orc_amg_level_dilu(const orc_amg *a, int l, int *colors, double *Einv)
{
    const orc_level *L = &a->lv[l];
    if (!L->colors) return 0;
    if (colors) memcpy(colors, L->colors, sizeof(int) * (size_t)L->n);
    if (Einv) memcpy(Einv, L->Einv, sizeof(double) * (size_t)L->n);
    return L->num_colors;
}
ORC_API void orc_amg_level_arrays(const orc_amg *a, int l, int *rp, int *ci, double *va, int *agg, int *Rp, int *Rc, double *d)
{
    const orc_level *L = &a->lv[l];
    if (rp) memcpy(rp, L->rp, sizeof(int) * ((size_t)L->n + 1));
    if (ci) memcpy(ci, L->ci, sizeof(int) * (size_t)L->nnz);
    if (va) memcpy(va, L->va, sizeof(double) * (size_t)L->nnz);
    if (agg && L->agg) memcpy(agg, L->agg, sizeof(int) * (size_t)L->n);
    if (Rp && L->Rp) memcpy(Rp, L->Rp, sizeof(int) * ((size_t)L->nagg + 1));
    if (Rc && L->Rc) memcpy(Rc, L->Rc, sizeof(int) * (size_t)L->n);
    if (d) memcpy(d, L->d, sizeof(double) * (size_t)L->n);
}


/* ------------------------------------------------------------------------------------------- */
/* DENSE_LU_SOLVER (src/solvers/dense_lu_solver.cu:745-985): dense copy of the coarsest matrix,   */
/* LU with partial pivoting (cuSOLVER getrf in the reference), x = A^-1 b (getrs).  Right-looking */
/* elimination, pivot = first entry of largest magnitude, multipliers scaled by the reciprocal    */
/* pivot, one FMA per update -- the order of the engine's single-CTA kernels.                     */
/* ------------------------------------------------------------------------------------------- */
ORC_API void orc_dense_lu_factor(int n, double *a, int lda, int *ipiv)
{
    for (int k = 0; k < n; k++) {
        int p = k;
        double best = fabs(a[(size_t)k + (size_t)k * lda]);
        for (int i = k + 1; i < n; i++) { const double v = fabs(a[(size_t)i + (size_t)k * lda]); if (v > best) { best = v; p = i; } }
        ipiv[k] = p;
        if (p != k)
            for (int j = 0; j < n; j++) { const double t = a[(size_t)k + (size_t)j * lda]; a[(size_t)k + (size_t)j * lda] = a[(size_t)p + (size_t)j * lda]; a[(size_t)p + (size_t)j * lda] = t; }
        const double piv = a[(size_t)k + (size_t)k * lda];
        if (piv != 0.0) { const double r = 1.0 / piv; for (int i = k + 1; i < n; i++) a[(size_t)i + (size_t)k * lda] *= r; }
        for (int j = k + 1; j < n; j++)
            for (int i = k + 1; i < n; i++)
                a[(size_t)i + (size_t)j * lda] = fma(-a[(size_t)i + (size_t)k * lda], a[(size_t)k + (size_t)j * lda], a[(size_t)i + (size_t)j * lda]);
    }
}
ORC_API void orc_dense_lu_solve(int n, const double *lu, int lda, const int *ipiv, const double *rhs, double *x)
{
    for (int i = 0; i < n; i++) x[i] = rhs[i];
    for (int k = 0; k < n; k++) { const int p = ipiv[k]; if (p != k) { const double t = x[k]; x[k] = x[p]; x[p] = t; } }
    for (int k = 0; k < n - 1; k++) { const double xk = x[k]; for (int i = k + 1; i < n; i++) x[i] = fma(-lu[(size_t)i + (size_t)k * lda], xk, x[i]); }
    for (int k = n - 1; k >= 0; k--) {
        x[k] = x[k] / lu[(size_t)k + (size_t)k * lda];
        const double xk = x[k];
        for (int i = 0; i < k; i++) x[i] = fma(-lu[(size_t)i + (size_t)k * lda], xk, x[i]);
    }
}
/* call after a setup: turn the coarsest level into a direct solve */
ORC_API void orc_amg_enable_dense_lu(orc_amg *a)
{
    orc_level *L = &a->lv[a->num_levels - 1];
    const int n = L->n;
    L->lu = (double *)calloc((size_t)(n > 0 ? n : 1) * (size_t)(n > 0 ? n : 1), sizeof(double));
    L->ipiv = (int *)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
    for (int i = 0; i < n; i++)
        for (int k = L->rp[i]; k < L->rp[i + 1]; k++) L->lu[(size_t)i + (size_t)L->ci[k] * n] = L->va[k];   /* csr_to_dense (last duplicate wins, as the scatter does) */
    orc_dense_lu_factor(n, L->lu, n, L->ipiv);
    a->dense_lu = 1;
}

static double butterfly(double *v, int w);
/* MULTICOLOR_GS, scalar: multicolorGSSmoothCsrKernel_nPerRow (src/solvers/multicolor_gauss_seidel_solver.cu:496-551) colour by
 * colour, in place; N = 4 lanes per row, 32 when nnz/rows > 20 or nnz/colours < 500000 (:1004-1013); lane (k - row_begin) % N
 * accumulates -a_ik x_k with one FMA, shuffle-down tree, lane 0: x_i += w (sum + b_i) / a_ii.  symmetric: colours up, then down. */
ORC_API void orc_gs_sweep(int n, const int *rp, const int *ci, const double *va, int num_colors, const int *sorted_rows, const int *offsets,
                          const double *b, double *x, double weight, int symmetric)
{
    const int nnz = rp[n];
    int N = 4;
    if (n > 0 && nnz / n > 20) N = 32;
    if (num_colors > 0 && nnz / num_colors < 500000) N = 32;
    for (int pass = 0; pass < (symmetric ? 2 : 1); pass++)
        for (int cc = 0; cc < num_colors; cc++) {
            const int c = pass == 0 ? cc : num_colors - 1 - cc;
            for (int q = offsets[c]; q < offsets[c + 1]; q++) {
                const int i = sorted_rows[q];
                double lane[32], dia = 0.0;
                for (int l = 0; l < 32; l++) lane[l] = 0.0;
                int found = 0;
                for (int k = rp[i]; k < rp[i + 1]; k++) {
                    lane[(k - rp[i]) % N] = fma(-va[k], x[ci[k]], lane[(k - rp[i]) % N]);
                    if (!found && ci[k] == i) { dia = va[k]; found = 1; }
                }
                double acc = butterfly(lane, N);       /* lane 0 of the shuffle-down tree holds the same association */
                acc += b[i];
                acc /= guard_d(dia);
                x[i] = fma(weight, acc, x[i]);
            }
        }
}

/* smoother->solve(b, x, xIsZero) with max_iters = sweeps (Solver::solve loop without monitoring) */
static void smooth(const orc_amg *a, orc_level *L, const double *b, double *x, int x_is_zero, int sweeps)
{
    if (a->smoother == 4) {
        /* Chebyshev_Solver inside Solver::solve: r = b (x "is zero": x itself is NOT cleared, as in the reference) or b - A x;
         * solve_init: z = M^-1 r, p = z; every iteration runs cheb_order steps (cheb_solver.cu:243-330) */
        const int n = L->n;
        double *p = L->cheb_p, *z = L->cheb_z, *r = L->cheb_r;
        if (x_is_zero) memcpy(r, b, sizeof(double) * (size_t)n);
        else orc_residual(n, L->rp, L->ci, L->va, x, b, r);
#define CHEB_PRECOND() do { if (a->cheb_precond) orc_jacobi_zero(n, L->d, r, z, a->cheb_inner_omega); else memcpy(z, r, sizeof(double) * (size_t)n); } while (0)
        CHEB_PRECOND();
        memcpy(p, z, sizeof(double) * (size_t)n);
        double gamma = 0., beta = 0.;
        int first = 0;
        const double ca = (L->lmax + L->lmin) / 2, cc = (L->lmax - L->lmin) / 2;
        for (int it = 0; it < sweeps; it++)
            for (int i = 0; i < a->cheb_order; i++) {
                CHEB_PRECOND();
                if (first == 0) { gamma = 1. / ca; first = 1; }
                else {
                    beta = cc * cc * gamma * gamma / 4.;
                    if (gamma != 0.0 && (ca - (beta / gamma)) != 0.0) gamma = 1. / (ca - beta / gamma);
                    for (int k = 0; k < n; k++) p[k] = z[k] * 1.0 + p[k] * beta;
                }
                for (int k = 0; k < n; k++) x[k] = fma(gamma, p[k], x[k]);
                orc_residual(n, L->rp, L->ci, L->va, x, b, r);
            }
#undef CHEB_PRECOND
        return;
    }
    if (a->smoother == 5) {
        /* ChebyshevPolySolver::smooth_1x1 (chebyshev_poly.cu:288-312): x = x + tau_i (b - A x); x is read as it is */
        const int n = L->n;
        int order = a->cheb_order < 1 ? 1 : (a->cheb_order > 10 ? 10 : a->cheb_order);
        for (int it = 0; it < sweeps; it++)
            for (int i = 0; i < order; i++) {
                orc_residual(n, L->rp, L->ci, L->va, x, b, L->cheb_r);
                for (int k = 0; k < n; k++) x[k] = fma(L->tau[i], L->cheb_r[k], x[k]);
            }
        return;
    }
    if (a->smoother == 3) {
        for (int it = 0; it < sweeps; it++) {
            if (it == 0 && x_is_zero) memset(x, 0, sizeof(double) * (size_t)L->n);
            orc_gs_sweep(L->n, L->rp, L->ci, L->va, L->num_colors, L->sorted_rows, L->color_offsets, b, x, a->omega, a->symmetric_gs);
        }
        return;
    }
    if (a->smoother == 2) {
        for (int it = 0; it < sweeps; it++) {
            if (it == 0 && x_is_zero) memset(x, 0, sizeof(double) * (size_t)L->n);
            orc_dilu_sweep_1x1(L->n, L->rp, L->ci, L->va, L->num_colors, L->colors, L->sorted_rows, L->color_offsets, L->Einv, b, x, a->omega, L->delta, L->Delta);
        }
        return;
    }
    for (int it = 0; it < sweeps; it++) {
        if (it == 0 && x_is_zero) { orc_jacobi_zero(L->n, L->d, b, x, a->omega); continue; }
        orc_jacobi_sweep(L->n, L->rp, L->ci, L->va, L->d, b, x, L->tmp, a->omega);
        memcpy(x, L->tmp, sizeof(double) * (size_t)L->n);
    }
}

/* FixedCycle::cycle, V cycle (src/cycles/fixed_cycle.cu:25-248) */
static void vcycle_t(const orc_amg *a, int l, const double *b, double *x, int x_is_zero, int type);

/* CG_CycleDispatcher / CG_Flex_CycleDispatcher::dispatch (src/cycles/cg_cycle.cu:18-101, cg_flex_cycle.cu:18-103): cycle_iters
 * iterations of (flexible) PCG on level l, preconditioned by one CG(F) fixed cycle of that level with zero initial guess.
 * x_is_zero = the level's init-cycle flag. */
static void cg_cycle_dispatch(const orc_amg *a, int l, const double *b, double *x, int flex)
{
    const orc_level *L = &a->lv[l];
    const int n = L->n, type = flex ? 4 : 3;
    const size_t nb = sizeof(double) * (size_t)(n > 0 ? n : 1);
    double *y = (double *)malloc(nb), *z = (double *)malloc(nb), *r = (double *)malloc(nb), *p = (double *)malloc(nb), *d = (double *)malloc(nb);
    memset(x, 0, sizeof(double) * (size_t)n);                     /* the dispatcher is always entered with the init flag set */
    orc_spmv(n, L->rp, L->ci, L->va, x, y);
    for (int i = 0; i < n; i++) r[i] = b[i] * 1.0 + y[i] * -1.0;  /* axpby(b, y, r, 1, -1) */
    vcycle_t(a, l, r, z, 1, type);
    memcpy(p, z, nb);
    double rz = flex ? 0.0 : orc_dot(n, r, z);
    int k = 0;
    for (;;) {
        orc_spmv(n, L->rp, L->ci, L->va, p, y);
        if (flex) rz = orc_dot(n, r, z);
        const double alpha = rz / orc_dot(n, y, p);
        for (int i = 0; i < n; i++) x[i] = fma(alpha, p[i], x[i]);
        if (++k == a->cycle_iters) break;
        if (flex) memcpy(d, r, nb);
        { const double ma = alpha * -1.0; for (int i = 0; i < n; i++) r[i] = fma(ma, y[i], r[i]); }
        if (flex) for (int i = 0; i < n; i++) d[i] = r[i] * 1.0 + d[i] * -1.0;
        vcycle_t(a, l, r, z, 1, type);
        double beta;
        if (flex) beta = orc_dot(n, z, d) / rz;
        else { const double rz_old = rz; rz = orc_dot(n, r, z); beta = rz / rz_old; }
        for (int i = 0; i < n; i++) p[i] = z[i] * 1.0 + p[i] * beta;
    }
    free(y); free(z); free(r); free(p); free(d);
}

/* prolongateAndApplyCorrection with error_scaling = 2 (minimise the residual 2-norm) or 3 (minimise the error A-norm):
 * aggregation_amg_level.cu:700-824.  L->r holds the residual the restriction was computed from. */
static void smooth(const orc_amg *a, orc_level *L, const double *b, double *x, int x_is_zero, int sweeps);
static void scaled_correction(orc_amg *a, orc_level *L, double *x)
{
    const int n = L->n;
    if (L->scale_counter > 0) {
        for (int i = 0; i < n; i++) x[i] = fma(L->scale, L->xc[L->agg[i]], x[i]);
        L->scale_counter--;
        return;
    }
    double *ef = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1)), *Aef = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) ef[i] = L->xc[L->agg[i]];         /* prolongateVector */
    if (a->scaling_smoother_steps > 0) smooth(a, L, L->r, ef, 0, a->scaling_smoother_steps);   /* smooth the correction with rhs = residual */
    orc_spmv(n, L->rp, L->ci, L->va, ef, Aef);
    double nom, den;
    if (a->error_scaling == 2) { nom = orc_dot(n, L->r, Aef); den = orc_dot(n, Aef, Aef); }
    else { nom = orc_dot(n, L->r, ef); den = orc_dot(n, ef, Aef); }
    if (fabs(den) == 0.0) nom = den = 1.0;
    double alpha = nom / den;
    if (fabs(alpha) < .3) alpha = (alpha / fabs(alpha)) * .3;
    if (fabs(alpha) > 10) alpha = (alpha / fabs(alpha)) * 10.;
    for (int i = 0; i < n; i++) x[i] = fma(alpha, ef[i], x[i]);   /* applyCorrection */
    L->scale_counter = a->reuse_scale;
    L->scale = alpha;
    free(ef); free(Aef);
}

static void vcycle(const orc_amg *a, int l, const double *b, double *x, int x_is_zero) { vcycle_t(a, l, b, x, x_is_zero, a->cycle); }
static void vcycle_t(const orc_amg *a, int l, const double *b, double *x, int x_is_zero, int type)
{
    orc_level *L = &a->lv[l];
    int finest = (l == 0);
    int n_pre;
    if (L->coarsest && a->dense_lu) n_pre = 0;                    /* coarsest level with a coarse solver (fixed_cycle.cu:75-78) */
    else if (L->coarsest) n_pre = a->coarsest_sweeps;
    else if (finest && a->finest_sweeps != -1) n_pre = a->presweeps == 0 ? 0 : a->finest_sweeps;
    else n_pre = a->presweeps;
    if (n_pre > 0) smooth(a, L, b, x, x_is_zero, n_pre);
    else if (x_is_zero) memset(x, 0, sizeof(double) * (size_t)L->n);
    if (L->coarsest && a->dense_lu) { orc_dense_lu_solve(L->n, L->lu, L->n, L->ipiv, b, x); return; }   /* launchCoarseSolver */
    if (L->coarsest) return;
    orc_residual(L->n, L->rp, L->ci, L->va, x, b, L->r);          /* axmb */
    if (L->Pp) cla_spmv(L->nagg, L->Rtp, L->Rtc, L->Rtv, L->r, L->bc);   /* classical: rr = R r (classical_amg_level.cu:620-627) */
    else orc_restrict(L->nagg, L->Rp, L->Rc, L->r, L->bc);        /* restrictResidual */
    /* next level, initial guess zero; W: two W cycles, F: a W then a V cycle; a single fixed cycle when the next level is
     * the coarsest (fixed_cycle.cu:169-179); the second visit continues from the first one's xc */
    if (type == 0 || a->lv[l + 1].coarsest) vcycle_t(a, l + 1, L->bc, L->xc, 1, 0);
    else if (type == 1) { vcycle_t(a, l + 1, L->bc, L->xc, 1, 1); vcycle_t(a, l + 1, L->bc, L->xc, 0, 1); }
    else if (type == 2) { vcycle_t(a, l + 1, L->bc, L->xc, 1, 1); vcycle_t(a, l + 1, L->bc, L->xc, 0, 0); }
    else cg_cycle_dispatch(a, l + 1, L->bc, L->xc, type == 4);
    if (L->Pp) {                                                  /* classical: tmp = P e; x = x + tmp (classical_amg_level.cu:884-910) */
        cla_spmv(L->n, L->Pp, L->Pc, L->Pv, L->xc, L->tmp);
        for (int i = 0; i < L->n; i++) x[i] = x[i] + L->tmp[i];
    } else if (a->error_scaling >= 2) scaled_correction((orc_amg *)a, L, x);
    else orc_prolong_add(L->n, L->agg, L->xc, x);               /* prolongateAndApplyCorrection */
    int n_post;
    if (finest && a->finest_sweeps != -1) n_post = a->postsweeps == 0 ? 0 : a->finest_sweeps;
    else n_post = a->postsweeps;
    if (n_post > 0) smooth(a, L, b, x, 0, n_post);
}

ORC_API void orc_amg_vcycle(const orc_amg *a, const double *b, double *x, int x_is_zero) { vcycle(a, 0, b, x, x_is_zero); }
ORC_API void orc_amg_set_cycle(orc_amg *a, int type) { a->cycle = type; }
ORC_API void orc_amg_set_cycle_iters(orc_amg *a, int iters) { a->cycle_iters = iters; }
ORC_API void orc_amg_set_symmetric_gs(orc_amg *a, int sym) { a->symmetric_gs = sym; }
/* call BEFORE the setup: which diagonal the inner Jacobi of smoother 4 uses (0 none / 1 a_ii / 2 L1 row norm) */
ORC_API void orc_set_chebyshev_precond(int precond) { g_cheb_precond = precond; }
/* call AFTER a setup with smoother 4 or 5: spectrum bounds per chebyshev_lambda_estimate_mode (cheb_solver.cu:186-213) and the
 * damped-root step lengths of CHEBYSHEV_POLY (chebyshev_poly.cu:63-74, 199-208) */
ORC_API void orc_amg_set_chebyshev(orc_amg *a, int order, int mode, int precond, double inner_omega, double user_max, double user_min)
{
    a->cheb_order = order; a->cheb_mode = mode; a->cheb_precond = precond; a->cheb_inner_omega = inner_omega;
    a->cheb_user_max = user_max; a->cheb_user_min = user_min;
    for (int l = 0; l < a->num_levels; l++) {
        orc_level *L = &a->lv[l];
        if (a->smoother == 4 && precond) {
            if (mode == 2) { L->lmax = 0.9; L->lmin = L->lmax * 0.125; }
            else { L->lmax = user_max; L->lmin = user_min; }
        }
        if (a->smoother == 5) {
            const int m = order < 1 ? 1 : (order > 10 ? 10 : order);
            const double lambda = L->lmax, beta = M_PI / (4 * (double)m + 2);
            for (int i = 0; i < m; i++)
                L->tau[i] = (cos(beta) * cos(beta) / (cos(beta * (2 * i + 1)) * cos(beta * (2 * i + 1)) - sin(beta) * sin(beta))) / lambda;
        }
    }
}
ORC_API void orc_amg_level_lambda(const orc_amg *a, int l, double *lmax, double *lmin) { *lmax = a->lv[l].lmax; *lmin = a->lv[l].lmin; }
ORC_API void orc_amg_set_error_scaling(orc_amg *a, int error_scaling, int scaling_smoother_steps, int reuse_scale)
{
    a->error_scaling = error_scaling; a->scaling_smoother_steps = scaling_smoother_steps; a->reuse_scale = reuse_scale;
    for (int l = 0; l < a->num_levels; l++) { a->lv[l].scale = 0.0; a->lv[l].scale_counter = 0; }
}

/* RELATIVE_INI criterion (src/convergence/relative_ini.cu:22-45) */
static int conv_relative_ini(double nrm, double nrm_ini, double tol)
{
    const double eps = 1e-20;
    int conv = (nrm_ini <= eps) ? 1 : (nrm / nrm_ini <= tol);
    double thr = nrm_ini * 1.0e-12;
    if (thr < eps) thr = eps;
    int conv_abs = nrm <= thr;
    return conv_abs || conv;
}

static double norm_of(int n, const double *r, int norm_type)
{
    return norm_type == 0 ? orc_nrm1(n, r) : norm_type == 1 ? orc_nrm2(n, r) : orc_nrmmax(n, r);
}

/* Preconditioned CG: Solver::solve loop (src/solvers/solver.cu:585-970) around PCG_Solver::solve_init /
 * solve_iteration (src/solvers/pcg_solver.cu:77-190).  precond: 0 none, 1 one AMG V-cycle (zero guess),
 * 2 one Jacobi sweep (zero guess: z = w r/d).  RELATIVE_INI convergence.  res_hist[0..iters].  Returns iterations. */
ORC_API int orc_pcg(int n, const int *rp, const int *ci, const double *va, const orc_amg *amg, int precond, double jac_omega, const double *b,
                    double *x, int x_is_zero, double tol, int max_iters, int norm_type, double *res_hist, int *converged_out)
{
    double *r = (double *)malloc(sizeof(double) * (size_t)n), *z = (double *)calloc((size_t)n, sizeof(double));
    double *p = (double *)malloc(sizeof(double) * (size_t)n), *Ap = (double *)malloc(sizeof(double) * (size_t)n);
    double *dj = NULL;
    if (precond == 2) { dj = (double *)malloc(sizeof(double) * (size_t)n); orc_extract_diag(n, rp, ci, va, dj); }
    if (x_is_zero) memcpy(r, b, sizeof(double) * (size_t)n);
    else orc_residual(n, rp, ci, va, x, b, r);
    double nrm = norm_of(n, r, norm_type), nrm_ini = nrm;
    res_hist[0] = nrm;
    int done = conv_relative_ini(nrm, nrm_ini, tol), it = 0, conv = done;
    if (max_iters == 0) { conv = 0; goto finish; }
    if (!done) {
        /* solve_init */
        if (precond == 1) orc_amg_vcycle(amg, r, z, 1);
        else if (precond == 2) orc_jacobi_zero(n, dj, r, z, jac_omega);
        else memcpy(z, r, sizeof(double) * (size_t)n);
        memcpy(p, z, sizeof(double) * (size_t)n);
    }
    {
        double rz = done ? 0.0 : orc_dot(n, r, z);
        for (it = 0; it < max_iters && !done; it++) {
            orc_spmv(n, rp, ci, va, p, Ap);
            double dApp = orc_dot(n, Ap, p);
            double alpha = 0.0;
            if (dApp != 0.0) alpha = rz / dApp;
            for (int i = 0; i < n; i++) x[i] = fma(alpha, p[i], x[i]);       /* axpy(p, x, alpha)   */
            for (int i = 0; i < n; i++) r[i] = fma(-alpha, Ap[i], r[i]);     /* axpy(Ap, r, -alpha) */
            nrm = norm_of(n, r, norm_type);
            res_hist[it + 1] = nrm;
            if (conv_relative_ini(nrm, nrm_ini, tol)) { done = 1; conv = 1; it++; break; }
            if (it == max_iters - 1) { it++; break; }
            if (precond == 1) orc_amg_vcycle(amg, r, z, 1);
            else if (precond == 2) orc_jacobi_zero(n, dj, r, z, jac_omega);
            else memcpy(z, r, sizeof(double) * (size_t)n);
            double rz_old = rz;
            rz = orc_dot(n, r, z);
            double beta = 0.0;
            if (rz_old != 0.0) beta = rz / rz_old;
            for (int i = 0; i < n; i++) p[i] = z[i] * 1.0 + p[i] * beta;     /* axpby(z, p, p, 1, beta) */
        }
    }
finish:
    if (converged_out) *converged_out = conv;
    free(r); free(z); free(p); free(Ap); free(dj);
    return it;
}

/* ------------------------------------------------------------------------------------------- */
/* FGMRES(m): FGMRES_Solver::solve_iteration (src/solvers/fgmres_solver.cu:406-569) inside the   */
/* Solver::solve loop, scalar L2 norm (convergence estimate |s[m+1]|), krylov_dim == restart.   */
/* precond: 0 none, 1 AMG V-cycle (zero guess), 2 Jacobi zero-guess sweep.                      */
/* ------------------------------------------------------------------------------------------- */
static void gen_rot(double dx, double dy, double *cs, double *sn)
{
    if (dy < 0.0) { *cs = 1.0; *sn = 0.0; }
    else if (fabs(dy) > fabs(dx)) { double t = dx / dy; *sn = 1.0 / sqrt(1.0 + t * t); *cs = t * *sn; }
    else { double t = dy / dx; *cs = 1.0 / sqrt(1.0 + t * t); *sn = t * *cs; }
}

ORC_API int orc_fgmres(int n, const int *rp, const int *ci, const double *va, const orc_amg *amg, int precond, double jac_omega, const double *b,
                       double *x, int x_is_zero, double tol, int max_iters, int restart, double *res_hist, int *converged_out)
{
    const int R = restart;
    double **V = (double **)malloc(sizeof(double *) * (size_t)(R + 1)), **Z = (double **)malloc(sizeof(double *) * (size_t)R);
    for (int i = 0; i <= R; i++) V[i] = (double *)calloc((size_t)n, sizeof(double));
    for (int i = 0; i < R; i++) Z[i] = (double *)calloc((size_t)n, sizeof(double));
    double *H = (double *)calloc((size_t)(R + 2) * (R + 1), sizeof(double));
    double *s = (double *)calloc((size_t)R + 2, sizeof(double)), *cs = (double *)calloc((size_t)R + 1, sizeof(double)), *sn = (double *)calloc((size_t)R + 1, sizeof(double));
    double *r = (double *)malloc(sizeof(double) * (size_t)n), *dj = NULL;
#define HH(i, j) H[(size_t)(i) * (R + 1) + (j)]
    if (precond == 2) { dj = (double *)malloc(sizeof(double) * (size_t)n); orc_extract_diag(n, rp, ci, va, dj); }
    /* Solver::solve: initial residual + norm */
    if (x_is_zero) memcpy(r, b, sizeof(double) * (size_t)n);
    else orc_residual(n, rp, ci, va, x, b, r);
    double nrm = orc_nrm2(n, r), nrm_ini = nrm, beta = 0.0;
    res_hist[0] = nrm;
    int done = conv_relative_ini(nrm, nrm_ini, tol), it = 0, conv = done;
    if (max_iters == 0) { conv = 0; goto fin; }
    for (it = 0; it < max_iters && !done; it++) {
        const int m = it % R;
        if (m == 0) {
            orc_residual(n, rp, ci, va, x, b, V[0]);
            beta = orc_nrm2(n, V[0]);
            if (it == 0 && conv_relative_ini(beta, nrm_ini, tol)) { res_hist[it + 1] = beta; conv = 1; it++; break; }
            { const double a = 1.0 / beta; for (int i = 0; i < n; i++) V[0][i] = V[0][i] * a; }
            for (int i = 0; i < R + 2; i++) s[i] = 0.0;
            s[0] = beta;
        }
        if (precond == 1) orc_amg_vcycle(amg, V[m], Z[m], 1);
        else if (precond == 2) orc_jacobi_zero(n, dj, V[m], Z[m], jac_omega);
        else memcpy(Z[m], V[m], sizeof(double) * (size_t)n);
        orc_spmv(n, rp, ci, va, Z[m], V[m + 1]);
        for (int i = 0; i <= m; i++) {
            const double h = orc_dot(n, V[i], V[m + 1]);
            HH(i, m) = h;
            for (int k = 0; k < n; k++) V[m + 1][k] = fma(-h, V[i][k], V[m + 1][k]);
        }
        HH(m + 1, m) = orc_nrm2(n, V[m + 1]);
        { const double a = 1.0 / HH(m + 1, m); for (int k = 0; k < n; k++) V[m + 1][k] = V[m + 1][k] * a; }
        for (int k = 0; k < m; k++) {
            const double t = cs[k] * HH(k, m) + sn[k] * HH(k + 1, m);
            HH(k + 1, m) = -sn[k] * HH(k, m) + cs[k] * HH(k + 1, m);
            HH(k, m) = t;
        }
        gen_rot(HH(m, m), HH(m + 1, m), &cs[m], &sn[m]);
        HH(m, m) = cs[m] * HH(m, m) + sn[m] * HH(m + 1, m);
        HH(m + 1, m) = 0.0;
        { const double t = cs[m] * s[m]; s[m + 1] = -sn[m] * s[m]; s[m] = t; }
        beta = fabs(s[m + 1]);
        res_hist[it + 1] = beta;
        const int cv = conv_relative_ini(beta, nrm_ini, tol);
        if (m == R - 1 || it == max_iters - 1 || cv) {
            for (int j = m; j >= 0; j--) {
                s[j] /= HH(j, j);
                for (int k = j - 1; k >= 0; k--) s[k] -= HH(k, j) * s[j];
            }
            for (int j = 0; j <= m; j++) for (int k = 0; k < n; k++) x[k] = fma(s[j], Z[j][k], x[k]);
        }
        if (cv) { conv = 1; done = 1; it++; break; }
    }
fin:
    if (converged_out) *converged_out = conv;
    for (int i = 0; i <= R; i++) free(V[i]);
    for (int i = 0; i < R; i++) free(Z[i]);
    free(V); free(Z); free(H); free(s); free(cs); free(sn); free(r); free(dj);
#undef HH
    return it;
}

/* ------------------------------------------------------------------------------------------- */
/* MIN_MAX colouring, one ring (src/matrix_coloring/min_max.cu:103-160, 380-420).  Signed hash     */
/* comparison; a pass colours local maxima `c` and local minima `c+1`; repeat while more than       */
/* max_uncolored_fraction*n rows are uncoloured (0 with determinism_flag).  Returns num_colors.     */
/* ------------------------------------------------------------------------------------------- */
ORC_API int orc_color_min_max(int n, const int *rp, const int *ci, double max_uncolored_fraction, int *colors)
{
    int *snap = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) colors[i] = 0;
    const int max_uncolored = (int)(max_uncolored_fraction * (double)n);
    int num_colors = 1;
    for (int num_uncolored = n; num_uncolored > max_uncolored;) {
        memcpy(snap, colors, sizeof(int) * (size_t)n);
        for (int i = 0; i < n; i++) {
            if (snap[i] != 0) continue;
            const int hash_i = (int)hash_val((unsigned)i, 0);
            int max_i = 1, min_i = 1;
            for (int r = rp[i]; r < rp[i + 1]; r++) {
                const int j = ci[r];
                if (j >= n) continue;
                const int hash_j = (int)hash_val((unsigned)j, 0);
                const int cj = snap[j];
                if (hash_j > hash_i && (cj == 0 || cj == num_colors)) max_i = 0;
                if (hash_j < hash_i && (cj == 0 || cj == num_colors + 1)) min_i = 0;
            }
            if (max_i) colors[i] = num_colors;
            else if (min_i) colors[i] = num_colors + 1;
        }
        num_colors += 2;
        num_uncolored = 0;
        for (int i = 0; i < n; i++) num_uncolored += (colors[i] == 0);
    }
    int mx = 0;
    for (int i = 0; i < n; i++) if (colors[i] > mx) mx = colors[i];
    free(snap);
    return mx + 1;
}

/* createColorArrays (src/matrix_coloring/matrix_coloring.cu:230-281): rows stably sorted by colour */
/* PARALLEL_GREEDY, coloring_level 1 (src/matrix_coloring/parallel_greedy.cu:148-215, 665-790) in its synchronous form: every launch
 * reads the colours of the previous one (the reference updates in place and is not reproducible run to run; see csrc/coloring.cu).
 * Colours start at 1, 0 = uncoloured; stops when the count taken BEFORE a launch is <= max_uncolored, does not change, or 64 colours
 * are in use.  Returns num_colors = max colour + 1. */
ORC_API int orc_color_parallel_greedy(int n, const int *rp, const int *ci, double max_uncolored_fraction, int *colors)
{
    int *next = (int *)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
    for (int i = 0; i < n; i++) colors[i] = 0;
    const int max_uncolored = (int)(max_uncolored_fraction * (double)n);
    int prev = 0, maxc = 0;
    while (n > 0) {
        int num_uncolored = 0;
        for (int i = 0; i < n; i++) num_uncolored += (colors[i] == 0);
        for (int i = 0; i < n; i++) {
            int c = colors[i];
            if (c == 0) {
                const int hi = (int)hash_val((unsigned)i, 0);
                unsigned long long used = 0ull;
                int max_row = 1;
                for (int r = rp[i]; r < rp[i + 1]; r++) {
                    const int j = ci[r];
                    if (j >= n || j == i) continue;
                    const int cj = colors[j];
                    if (cj > 0 && cj <= 64) used |= 1ull << (64 - cj);
                    max_row &= (hi > (int)hash_val((unsigned)j, 0) || cj != 0);
                }
                if (max_row && ~used != 0ull) {
                    int p = 63;
                    while (!((~used >> p) & 1ull)) p--;          /* bfind(~used) */
                    c = 64 - p;
                }
            }
            next[i] = c;
            if (c > maxc) maxc = c;
        }
        memcpy(colors, next, sizeof(int) * (size_t)n);
        if (maxc + 1 >= 64 || prev == num_uncolored || num_uncolored <= max_uncolored) break;
        prev = num_uncolored;
    }
    free(next);
    return maxc + 1;
}
static int g_coloring_scheme = 0;    /* 0 MIN_MAX, 1 PARALLEL_GREEDY: scheme of the multicolour smoothers of the NEXT setups */
ORC_API void orc_set_coloring_scheme(int scheme) { g_coloring_scheme = scheme; }

ORC_API void orc_color_arrays(int n, int num_colors, const int *colors, int *sorted_rows, int *offsets)
{
    for (int c = 0; c <= num_colors; c++) offsets[c] = 0;
    for (int i = 0; i < n; i++) offsets[colors[i] + 1]++;
    for (int c = 0; c < num_colors; c++) offsets[c + 1] += offsets[c];
    int *pos = (int *)malloc(sizeof(int) * (size_t)(num_colors > 0 ? num_colors : 1));
    for (int c = 0; c < num_colors; c++) pos[c] = offsets[c];
    for (int i = 0; i < n; i++) sorted_rows[pos[colors[i]]++] = i;
    free(pos);
}

/* butterfly reduction over `w` lanes: v[l] += v[l ^ m] for m = w/2 .. 1 (all lanes end with the same sum) */
static double butterfly(double *v, int w)
{
    double t[32];
    for (int m = w / 2; m > 0; m >>= 1) {
        for (int l = 0; l < w; l++) t[l] = v[l] + v[l ^ m];
        for (int l = 0; l < w; l++) v[l] = t[l];
    }
    return v[0];
}

/* DILU_setup_1x1_kernel (src/solvers/multicolor_dilu_solver.cu:643-812): colour by colour;
 * lane (k - row_begin) % 32 accumulates e += (a_ji*Einv_j)*a_ij with one FMA, then a 32-lane butterfly. */
ORC_API void orc_dilu_setup_1x1(int n, const int *rp, const int *ci, const double *va, int num_colors, const int *colors, const int *sorted_rows,
                                const int *offsets, double *Einv)
{
    int *diag = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    orc_diag_index(n, rp, ci, diag);
    for (int i = 0; i < n; i++) Einv[i] = 0.0;
    for (int c = 0; c < num_colors; c++)
        for (int q = offsets[c]; q < offsets[c + 1]; q++) {
            const int i = sorted_rows[q];
            double lane[32];
            for (int l = 0; l < 32; l++) lane[l] = 0.0;
            if (c != 0)
                for (int k = rp[i]; k < rp[i + 1]; k++) {
                    const int j = ci[k];
                    if (j == i || colors[j] >= c) continue;
                    double a_ji = 0.0;
                    for (int kk = rp[j]; kk < rp[j + 1]; kk++)
                        if (ci[kk] == i) { a_ji = va[kk]; break; }
                    const int l = (k - rp[i]) % 32;
                    lane[l] = fma(a_ji * Einv[j], va[k], lane[l]);
                }
            const double e_out = butterfly(lane, 32);
            double res = (diag[i] >= 0 ? va[diag[i]] : 0.0) - e_out;
            if (res != 0.0) res = 1.0 / res;
            Einv[i] = res;
        }
    free(diag);
}

/* one DILU sweep, 1x1 (forward :1763-1901, backward :2772-2881, last colour :2885-2915), 8 lanes per row */
ORC_API void orc_dilu_sweep_1x1(int n, const int *rp, const int *ci, const double *va, int num_colors, const int *colors, const int *sorted_rows,
                                const int *offsets, const double *Einv, const double *b, double *x, double weight, double *delta, double *Delta)
{
    for (int c = 0; c < num_colors; c++)
        for (int q = offsets[c]; q < offsets[c + 1]; q++) {
            const int i = sorted_rows[q];
            double lane[8];
            for (int l = 0; l < 8; l++) lane[l] = 0.0;
            lane[0] = b[i];
            for (int k = rp[i]; k < rp[i + 1]; k++) {
                const int j = ci[k], l = (k - rp[i]) % 8;
                double xx = x[j];
                if (c != 0 && j < n && colors[j] < c) xx += delta[j];
                lane[l] = fma(-va[k], xx, lane[l]);
            }
            delta[i] = Einv[i] * butterfly(lane, 8);
        }
    for (int c = num_colors - 1; c >= 0; c--)
        for (int q = offsets[c]; q < offsets[c + 1]; q++) {
            const int i = sorted_rows[q];
            if (c == num_colors - 1) {
                const double v = delta[i];
                x[i] = fma(weight, v, x[i]);
                Delta[i] = v;
                continue;
            }
            double lane[8];
            for (int l = 0; l < 8; l++) lane[l] = 0.0;
            for (int k = rp[i]; k < rp[i + 1]; k++) {
                const int j = ci[k], l = (k - rp[i]) % 8;
                if (c != 0 && j < n && colors[j] > c) lane[l] = fma(va[k], Delta[j], lane[l]);
            }
            const double v = fma(-Einv[i], butterfly(lane, 8), delta[i]);   /* delta - Einv*acc, contracted on the device */
            x[i] = fma(weight, v, x[i]);
            Delta[i] = v;
        }
}

/* ------------------------------------------------------------------------------------------- */
/* 4x4 block kernels (values fp64 here; the fp32-matrix mode is checked with a tolerance)        */
/* ------------------------------------------------------------------------------------------- */
ORC_API void orc_bspmv4(int n, const int *rp, const int *ci, const double *va, const double *x, double *y)
{
    for (int i = 0; i < n; i++)
        for (int r = 0; r < 4; r++) {
            double s = 0.0;
            for (int k = rp[i]; k < rp[i + 1]; k++)
                for (int m = 0; m < 4; m++) s = fma(va[(size_t)k * 16 + r * 4 + m], x[(size_t)ci[k] * 4 + m], s);
            y[(size_t)i * 4 + r] = s;
        }
}

/* Gauss-Jordan without pivoting, order of compute_block_inverse_row_major (block_common_solver.h:106-137) */
ORC_API void orc_invert4x4(double *A)
{
    for (int row = 0; row < 4; row++) {
        const double diag = 1.0 / guard_d(A[row * 4 + row]);
        for (int j = 0; j < 4; j++) if (j != row) A[row * 4 + j] = A[row * 4 + j] * diag;
        for (int i = 0; i < 4; i++) if (i != row)
            for (int j = 0; j < 4; j++) if (j != row) A[i * 4 + j] = fma(-A[i * 4 + row], A[row * 4 + j], A[i * 4 + j]);
        for (int j = 0; j < 4; j++) A[j * 4 + row] = (j == row) ? diag : -(A[j * 4 + row] * diag);
    }
}

ORC_API void orc_bjacobi4_dinv(int n, const int *rp, const int *ci, const double *va, double *dinv)
{
    for (int i = 0; i < n; i++) {
        for (int m = 0; m < 16; m++) dinv[(size_t)i * 16 + m] = 0.0;
        for (int k = rp[i]; k < rp[i + 1]; k++)
            if (ci[k] == i) { memcpy(dinv + (size_t)i * 16, va + (size_t)k * 16, sizeof(double) * 16); break; }
        orc_invert4x4(dinv + (size_t)i * 16);
    }
}

/* jacobiSmooth4by4BlockDiaCsrKernel_NAIVE_tex_readDinv2 (block_jacobi_solver.cu:665-739) */
ORC_API void orc_bjacobi4_sweep(int n, const int *rp, const int *ci, const double *va, const double *dinv, const double *b, const double *x,
                                double *xout, double weight)
{
    for (int i = 0; i < n; i++) {
        double bm[4];
        for (int r = 0; r < 4; r++) {
            double acc = b[(size_t)i * 4 + r];
            for (int k = rp[i]; k < rp[i + 1]; k++)
                if (ci[k] == i) { for (int m = 0; m < 4; m++) acc = fma(-va[(size_t)k * 16 + r * 4 + m], x[(size_t)i * 4 + m], acc); break; }
            for (int k = rp[i]; k < rp[i + 1]; k++) {
                if (ci[k] == i) continue;
                for (int m = 0; m < 4; m++) acc = fma(-va[(size_t)k * 16 + r * 4 + m], x[(size_t)ci[k] * 4 + m], acc);
            }
            bm[r] = acc;
        }
        for (int r = 0; r < 4; r++) {
            double t = 0.0;
            for (int m = 0; m < 4; m++) t = fma(dinv[(size_t)i * 16 + r * 4 + m], bm[m], t);
            xout[(size_t)i * 4 + r] = fma(t, weight, x[(size_t)i * 4 + r]);
        }
    }
}

ORC_API void orc_bjacobi4_zero(int n, const double *dinv, const double *b, double *x, double weight)
{
    for (int i = 0; i < n; i++)
        for (int r = 0; r < 4; r++) {
            double t = 0.0;
            for (int m = 0; m < 4; m++) t = fma(dinv[(size_t)i * 16 + r * 4 + m], b[(size_t)i * 4 + m], t);
            x[(size_t)i * 4 + r] = t * weight;
        }
}

/* DILU 4x4: setup (:362-640) and one sweep (:1585-1759, 2604-2768) -- plain formulas, tolerance-checked */
ORC_API void orc_dilu_setup_4x4(int n, const int *rp, const int *ci, const double *va, int num_colors, const int *colors, const int *sorted_rows,
                                const int *offsets, double *Einv)
{
    for (size_t t = 0; t < (size_t)n * 16; t++) Einv[t] = 0.0;
    for (int c = 0; c < num_colors; c++)
        for (int q = offsets[c]; q < offsets[c + 1]; q++) {
            const int i = sorted_rows[q];
            double E[16];
            for (int m = 0; m < 16; m++) E[m] = 0.0;
            for (int k = rp[i]; k < rp[i + 1]; k++) if (ci[k] == i) { memcpy(E, va + (size_t)k * 16, sizeof(E)); break; }
            if (c != 0)
                for (int k = rp[i]; k < rp[i + 1]; k++) {
                    const int j = ci[k];
                    if (j == i || colors[j] >= c) continue;
                    int kji = -1;
                    for (int kk = rp[j]; kk < rp[j + 1]; kk++) if (ci[kk] == i) { kji = kk; break; }
                    double T[16];
                    for (int r = 0; r < 4; r++)
                        for (int cc = 0; cc < 4; cc++) {
                            double t = 0.0;
                            for (int m = 0; m < 4; m++) t += va[(size_t)k * 16 + r * 4 + m] * Einv[(size_t)j * 16 + m * 4 + cc];
                            T[r * 4 + cc] = t;
                        }
                    if (kji >= 0)
                        for (int r = 0; r < 4; r++)
                            for (int cc = 0; cc < 4; cc++)
                                for (int m = 0; m < 4; m++) E[r * 4 + cc] -= T[r * 4 + m] * va[(size_t)kji * 16 + m * 4 + cc];
                }
            orc_invert4x4(E);
            memcpy(Einv + (size_t)i * 16, E, sizeof(E));
        }
}

ORC_API void orc_dilu_sweep_4x4(int n, const int *rp, const int *ci, const double *va, int num_colors, const int *colors, const int *sorted_rows,
                                const int *offsets, const double *Einv, const double *b, double *x, double weight, double *delta, double *Delta)
{
    for (int c = 0; c < num_colors; c++)
        for (int q = offsets[c]; q < offsets[c + 1]; q++) {
            const int i = sorted_rows[q];
            double acc[4];
            for (int r = 0; r < 4; r++) {
                acc[r] = b[(size_t)i * 4 + r];
                for (int k = rp[i]; k < rp[i + 1]; k++) {
                    const int j = ci[k];
                    const int valid = c != 0 && j < n && colors[j] < c;
                    for (int m = 0; m < 4; m++) {
                        double xx = x[(size_t)j * 4 + m];
                        if (valid) xx += delta[(size_t)j * 4 + m];
                        acc[r] -= va[(size_t)k * 16 + r * 4 + m] * xx;
                    }
                }
            }
            for (int r = 0; r < 4; r++) {
                double y = 0.0;
                for (int m = 0; m < 4; m++) y += Einv[(size_t)i * 16 + r * 4 + m] * acc[m];
                delta[(size_t)i * 4 + r] = y;
            }
        }
    for (int c = num_colors - 1; c >= 0; c--)
        for (int q = offsets[c]; q < offsets[c + 1]; q++) {
            const int i = sorted_rows[q];
            double acc[4] = {0, 0, 0, 0};
            if (c != num_colors - 1)
                for (int r = 0; r < 4; r++)
                    for (int k = rp[i]; k < rp[i + 1]; k++) {
                        const int j = ci[k];
                        if (!(c != 0 && j < n && colors[j] > c)) continue;
                        for (int m = 0; m < 4; m++) acc[r] += va[(size_t)k * 16 + r * 4 + m] * Delta[(size_t)j * 4 + m];
                    }
            for (int r = 0; r < 4; r++) {
                double y = 0.0;
                for (int m = 0; m < 4; m++) y += Einv[(size_t)i * 16 + r * 4 + m] * acc[m];
                const double v = delta[(size_t)i * 4 + r] - y;
                x[(size_t)i * 4 + r] += weight * v;
                Delta[(size_t)i * 4 + r] = v;
            }
        }
}

/* AMG as the outer solver: Solver::solve loop (src/solvers/solver.cu:585-970) around
 * AlgebraicMultigrid_Solver::solve_iteration (one V-cycle, then residual + norm + RELATIVE_INI check). */
ORC_API int orc_amg_solve(const orc_amg *amg, int n, const int *rp, const int *ci, const double *va, const double *b, double *x, int x_is_zero,
                          double tol, int max_iters, int norm_type, double *res_hist, int *converged_out)
{
    double *r = (double *)malloc(sizeof(double) * (size_t)n);
    if (x_is_zero) memcpy(r, b, sizeof(double) * (size_t)n);
    else orc_residual(n, rp, ci, va, x, b, r);
    double nrm = norm_of(n, r, norm_type), nrm_ini = nrm;
    res_hist[0] = nrm;
    int done = conv_relative_ini(nrm, nrm_ini, tol), conv = done, it = 0;
    if (max_iters == 0) conv = 0;
    else
        for (it = 0; it < max_iters && !done; it++) {
            orc_amg_vcycle(amg, b, x, x_is_zero && it == 0);
            orc_residual(n, rp, ci, va, x, b, r);
            nrm = norm_of(n, r, norm_type);
            res_hist[it + 1] = nrm;
            if (conv_relative_ini(nrm, nrm_ini, tol)) { done = 1; conv = 1; it++; break; }
        }
    if (converged_out) *converged_out = conv;
    free(r);
    return it;
}

#include "krylov_oracle.inc.c"
