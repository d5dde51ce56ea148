/* krylov_oracle.inc.c -- TEST INFRASTRUCTURE (included by amg_oracle.c): CPU restatements of the remaining Krylov drivers
 * of the reference, each inside the Solver::solve loop (src/solvers/solver.cu:585-970), RELATIVE_INI convergence:
 *   CG          src/solvers/cg_solver.cu:41-103          (no preconditioner)
 *   PCGF        src/solvers/pcgf_solver.cu:76-175        (flexible PCG: beta = <z, r_new - r_old> / <r, z>)
 *   PBICGSTAB   src/solvers/pbicgstab_solver.cu:122-262
 *   GMRES       src/solvers/gmres_solver.cu:215-395      (right-preconditioned GMRES(m), one Z vector)
 * precond: 0 none, 1 one AMG cycle with zero initial guess, 2 one Jacobi sweep with zero initial guess.
 * axpy is an fma (cublasDaxpy), axpby / axpbypcz are x*a + y*b (+ z*c) (src/blas.cu:107-124), as in orc_pcg.
 * parity unpinned: no reference golden exists yet for these four (tests/golden/make_golden.py lists the cases to generate). */

typedef struct {
    int n, precond;
    const int *rp, *ci;
    const double *va;
    const orc_amg *amg;
    double jac_omega, *dj;
} kry_ctx;

static void kry_precond(const kry_ctx *c, const double *in, double *out)
{
    if (c->precond == 1) orc_amg_vcycle(c->amg, in, out, 1);
    else if (c->precond == 2) orc_jacobi_zero(c->n, c->dj, in, out, c->jac_omega);
    else memcpy(out, in, sizeof(double) * (size_t)c->n);
}
static void kry_axpy(int n, const double *x, double *y, double a) { for (int i = 0; i < n; i++) y[i] = fma(a, x[i], y[i]); }
static void kry_axpby(int n, const double *x, const double *y, double *o, double a, double b) { for (int i = 0; i < n; i++) o[i] = x[i] * a + y[i] * b; }
static void kry_axpbypcz(int n, const double *x, const double *y, const double *z, double *o, double a, double b, double c)
{
    for (int i = 0; i < n; i++) o[i] = x[i] * a + y[i] * b + z[i] * c;
}

/* kind: 0 CG, 1 PCGF, 2 PBICGSTAB, 3 GMRES(restart).  Returns the iteration count; res_hist[0..iters]. */
ORC_API int orc_krylov(int kind, int n, const int *rp, const int *ci, const double *va, const orc_amg *amg, int precond, double jac_omega,
                       const double *b, double *x, int x_is_zero, double tol, int max_iters, int restart, int norm_type, double *res_hist,
                       int *converged_out)
{
    kry_ctx c = {n, precond, rp, ci, va, amg, jac_omega, NULL};
    const size_t nb = sizeof(double) * (size_t)(n > 0 ? n : 1);
    if (precond == 2) { c.dj = (double *)malloc(nb); orc_extract_diag(n, rp, ci, va, c.dj); }
    if (kind == 0) c.precond = 0;
    double *r = (double *)malloc(nb);
    if (x_is_zero) memcpy(r, b, nb);
    else orc_residual(n, rp, ci, va, x, b, r);
    double nrm = norm_of(n, r, norm_type), nrm_ini = nrm;
    res_hist[0] = nrm;
    int done = conv_relative_ini(nrm, nrm_ini, tol), it = 0, conv = done;
    double *w[8] = {0};
    double **V = NULL, *H = NULL, *s = NULL, *cs = NULL, *sn = NULL;
    if (max_iters == 0) { conv = 0; goto fin; }
    if (done) goto fin;
    for (int i = 0; i < 8; i++) w[i] = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    if (kind == 0) {                         /* ---- CG ---- */
        double *p = w[0], *Ap = w[1];
        memcpy(p, r, nb);
        double rr = orc_dot(n, r, r);
        for (it = 0; it < max_iters; it++) {
            orc_spmv(n, rp, ci, va, p, Ap);
            const double alpha = rr / orc_dot(n, Ap, p);
            kry_axpy(n, p, x, alpha);
            kry_axpy(n, Ap, r, -alpha);
            nrm = norm_of(n, r, norm_type);
            res_hist[it + 1] = nrm;
            if (conv_relative_ini(nrm, nrm_ini, tol)) { conv = 1; it++; break; }
            if (it == max_iters - 1) { it++; break; }
            const double rr_old = rr;
            rr = orc_dot(n, r, r);
            kry_axpby(n, r, p, p, 1.0, rr / rr_old);
        }
    } else if (kind == 1) {                  /* ---- PCGF ---- */
        double *p = w[0], *z = w[1], *Ap = w[2], *d = w[3];
        kry_precond(&c, r, z);
        memcpy(p, z, nb);
        for (it = 0; it < max_iters; it++) {
            orc_spmv(n, rp, ci, va, p, Ap);
            const double rz = orc_dot(n, r, z);
            const double alpha = rz / orc_dot(n, Ap, p);
            kry_axpy(n, p, x, alpha);
            memcpy(d, r, nb);
            kry_axpy(n, Ap, r, -alpha);
            nrm = norm_of(n, r, norm_type);
            res_hist[it + 1] = nrm;
            if (conv_relative_ini(nrm, nrm_ini, tol)) { conv = 1; it++; break; }
            if (it == max_iters - 1) { it++; break; }
            kry_axpby(n, r, d, d, 1.0, -1.0);
            kry_precond(&c, r, z);
            const double beta = orc_dot(n, z, d) / rz;
            kry_axpby(n, z, p, p, 1.0, beta);
        }
    } else if (kind == 2) {                  /* ---- PBICGSTAB ---- */
        double *p = w[0], *Mp = w[1], *sv = w[2], *Ms = w[3], *t = w[4], *v = w[5], *rt = w[6];
        memcpy(rt, r, nb);
        double rho = orc_dot(n, rt, r);
        memcpy(p, r, nb);
        for (it = 0; it < max_iters; it++) {
            kry_precond(&c, p, Mp);
            orc_spmv(n, rp, ci, va, Mp, v);
            double red = orc_dot(n, rt, v);
            const double alpha = (red != 0.0) ? rho / red : 0.0;
            kry_axpby(n, r, v, sv, 1.0, -alpha);
            if (conv_relative_ini(norm_of(n, sv, norm_type), nrm_ini, tol)) {     /* early exit on ||s|| */
                kry_axpby(n, x, Mp, x, 1.0, alpha);
                orc_residual(n, rp, ci, va, x, b, r);
                nrm = norm_of(n, r, norm_type);
                res_hist[it + 1] = nrm;
                conv = 1; it++; break;
            }
            kry_precond(&c, sv, Ms);
            orc_spmv(n, rp, ci, va, Ms, t);
            red = orc_dot(n, t, t);
            double omega = orc_dot(n, t, sv);
            omega = (red == 0.0) ? 0.0 : omega / red;
            kry_axpbypcz(n, x, Mp, Ms, x, 1.0, alpha, omega);
            kry_axpby(n, sv, t, r, 1.0, -omega);
            nrm = norm_of(n, r, norm_type);
            res_hist[it + 1] = nrm;
            if (conv_relative_ini(nrm, nrm_ini, tol)) { conv = 1; it++; break; }
            if (it == max_iters - 1) { it++; break; }
            const double rho_new = orc_dot(n, rt, r);
            double beta = 0.0;
            if (rho != 0.0 && omega != 0.0) beta = (rho_new / rho) * (alpha / omega);
            rho = rho_new;
            kry_axpbypcz(n, r, p, v, p, 1.0, beta, -beta * omega);
        }
    } else {                                 /* ---- GMRES(R), max_iters > 1 branch (gmres_solver.cu:282-395) ---- */
        const int R = restart, K = (max_iters < R ? max_iters : R);
        double *Z = w[0];
        V = (double **)malloc(sizeof(double *) * (size_t)(K + 1));
        for (int i = 0; i <= K; i++) V[i] = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
        H = (double *)calloc((size_t)(K + 1) * (size_t)K, sizeof(double));
        s = (double *)calloc((size_t)K + 1, sizeof(double));
        cs = (double *)calloc((size_t)K, sizeof(double));
        sn = (double *)calloc((size_t)K, sizeof(double));
#define HH(i, j) H[(size_t)(i) + (size_t)(j) * (size_t)(K + 1)]
        if (max_iters == 1) {                /* solve_one_iteration (gmres_solver.cu:215-268): no second preconditioner application */
            orc_spmv(n, rp, ci, va, x, V[0]);
            kry_axpy(n, b, V[0], -1.0);
            const double beta = orc_nrm2(n, V[0]);
            { const double a = -1.0 / beta; for (int k = 0; k < n; k++) V[0][k] = V[0][k] * a; }
            s[0] = beta;
            kry_precond(&c, V[0], Z);
            orc_spmv(n, rp, ci, va, Z, V[1]);
            HH(0, 0) = orc_dot(n, V[1], V[0]);
            kry_axpy(n, V[0], V[1], -HH(0, 0));
            HH(1, 0) = orc_nrm2(n, V[1]);
            gen_rot(HH(0, 0), HH(1, 0), &cs[0], &sn[0]);
            { const double t = cs[0] * s[0]; s[1] = -sn[0] * s[0]; s[0] = t; }
            HH(0, 0) = cs[0] * HH(0, 0) + sn[0] * HH(1, 0);
            nrm = fabs(s[1]);
            res_hist[1] = nrm;
            s[0] = s[0] / HH(0, 0);
            kry_axpy(n, Z, x, s[0]);
            conv = conv_relative_ini(nrm, nrm_ini, tol);
            it = 1;
        } else
        for (it = 0; it < max_iters; it++) {
            const int i = it % R;
            if (i == 0) {
                orc_spmv(n, rp, ci, va, x, V[0]);                 /* V0 = A x      */
                kry_axpy(n, b, V[0], -1.0);                       /* V0 = V0 - b   */
                const double beta = orc_nrm2(n, V[0]);
                if (conv_relative_ini(beta, nrm_ini, tol)) { res_hist[it + 1] = beta; conv = 1; it++; break; }
                { const double a = -1.0 / beta; for (int k = 0; k < n; k++) V[0][k] = V[0][k] * a; }
                for (int k = 0; k <= K; k++) s[k] = 0.0;
                s[0] = beta;
            }
            kry_precond(&c, V[i], Z);
            orc_spmv(n, rp, ci, va, Z, V[i + 1]);
            for (int k = 0; k <= i; k++) {
                HH(k, i) = orc_dot(n, V[i + 1], V[k]);
                kry_axpy(n, V[k], V[i + 1], -HH(k, i));
            }
            HH(i + 1, i) = orc_nrm2(n, V[i + 1]);
            { const double a = 1.0 / HH(i + 1, i); for (int k = 0; k < n; k++) V[i + 1][k] = V[i + 1][k] * a; }
            for (int k = 0; k < i; k++) {
                const double t = cs[k] * HH(k, i) + sn[k] * HH(k + 1, i);
                HH(k + 1, i) = cs[k] * HH(k + 1, i) - sn[k] * HH(k, i);
                HH(k, i) = t;
            }
            gen_rot(HH(i, i), HH(i + 1, i), &cs[i], &sn[i]);
            { const double t = cs[i] * s[i]; s[i + 1] = -sn[i] * s[i]; s[i] = t; }
            HH(i, i) = cs[i] * HH(i, i) + sn[i] * HH(i + 1, i);
            HH(i + 1, i) = 0.0;
            nrm = fabs(s[i + 1]);
            res_hist[it + 1] = nrm;
            const int cv = conv_relative_ini(nrm, nrm_ini, tol);
            if (i == R - 1 || it == max_iters - 1 || cv) {
                for (int j = i; j >= 0; j--) {
                    s[j] = s[j] / HH(j, j);
                    for (int k = j - 1; k >= 0; k--) s[k] = s[k] - HH(k, j) * s[j];
                }
                memset(Z, 0, nb);
                for (int j = 0; j <= i; j++) kry_axpy(n, V[j], Z, s[j]);
                kry_precond(&c, Z, V[0]);
                kry_axpy(n, V[0], x, 1.0);
            }
            if (cv) { conv = 1; it++; break; }
        }
#undef HH
        if (V) { for (int i = 0; i <= K; i++) free(V[i]); free(V); }
        free(H); free(s); free(cs); free(sn);
    }
fin:
    if (converged_out) *converged_out = conv;
    for (int i = 0; i < 8; i++) free(w[i]);
    free(r); free(c.dj);
    return it;
}

/* ------------------------------------------------------------------------------------------- */
/* FGMRES with gmres_krylov_dim < gmres_n_restart: the reference's truncated variant ("DQGMRES"). */
/*   KrylovSubspaceBuffer          src/solvers/fgmres_solver.cu:17-211: rings of K+2 V- and K+1 Z-vectors (max_dimension = K + 1),        */
/*                                 get_smallest_m() = max(m - K, 0)                                                                       */
/*   solver_setup                  :284-298: update_x_every_iteration = update_r_every_iteration = (K < R) (with monitoring)              */
/*   solve_iteration               :406-569: truncated modified Gram-Schmidt over V(smallest..m); ALL Givens rotations of the cycle are   */
/*                                 applied to the new column (rows below `smallest` hold whatever the previous cycle left there: m_H is   */
/*                                 allocated once and never cleared -- kept here on purpose); p_m = (z_m - sum_{i>=smallest} h_im p_i) /  */
/*                                 h_mm, x += s_m p_m every iteration; the residual VECTOR follows the recursion of :520-533 and its L2    */
/*                                 norm (not |s[m+1]|) drives the convergence check (checkConvergenceGMRES :348-398).                      */
/* Same conventions as orc_fgmres above.  parity unpinned: no reference golden was generated for this variant (no shipped configuration  */
/* sets gmres_krylov_dim); the restatement is checked against an independent numpy formulation (tests/test_oracle_krylov.py).             */
/* ------------------------------------------------------------------------------------------- */
ORC_API int orc_fgmres_trunc(int n, const int *rp, const int *ci, const double *va, const orc_amg *amg, int precond, double jac_omega, const double *b,
                             double *x, int x_is_zero, double tol, int max_iters, int restart, int krylov_dim, double *res_hist, int *converged_out)
{
    const int R = restart;
    int K = max_iters < R ? max_iters : R;
    if (krylov_dim > 0 && krylov_dim < K) K = krylov_dim;
    const int NV = K + 2, NZ = K + 1;
    const size_t nn = (size_t)(n > 0 ? n : 1), nb = sizeof(double) * nn;
    kry_ctx c = {n, precond, rp, ci, va, amg, jac_omega, NULL};
    if (precond == 2) { c.dj = (double *)malloc(nb); orc_extract_diag(n, rp, ci, va, c.dj); }
    double **V = (double **)malloc(sizeof(double *) * (size_t)NV), **Z = (double **)malloc(sizeof(double *) * (size_t)NZ);
    for (int i = 0; i < NV; i++) V[i] = (double *)calloc(nn, sizeof(double));
    for (int i = 0; i < NZ; i++) Z[i] = (double *)calloc(nn, sizeof(double));
    double *H = (double *)calloc((size_t)(R + 2) * (size_t)(R + 1), sizeof(double));
    double *s = (double *)calloc((size_t)R + 2, sizeof(double)), *cs = (double *)calloc((size_t)R + 1, sizeof(double)), *sn = (double *)calloc((size_t)R + 1, sizeof(double));
    double *gamma = (double *)calloc((size_t)R + 2, sizeof(double));
    double *r = (double *)malloc(nb), *resid = (double *)calloc(nn, sizeof(double));
#define HH(i, j) H[(size_t)(i) * (size_t)(R + 1) + (size_t)(j)]
#define VV(i) V[(i) % NV]
#define ZZ(i) Z[(i) % NZ]
    if (x_is_zero) memcpy(r, b, nb);
    else orc_residual(n, rp, ci, va, x, b, r);
    double nrm = orc_nrm2(n, r), nrm_ini = nrm;
    res_hist[0] = nrm;
    int done = conv_relative_ini(nrm, nrm_ini, tol), it = 0, conv = done;
    if (max_iters == 0) { conv = 0; goto fin; }
    for (it = 0; it < max_iters && !done; it++) {
        const int m = it % R;
        if (m == 0) {
            orc_residual(n, rp, ci, va, x, b, VV(0));
            const double beta = orc_nrm2(n, VV(0));
            if (it == 0 && conv_relative_ini(beta, nrm_ini, tol)) { res_hist[it + 1] = beta; conv = 1; it++; break; }
            { const double a = 1.0 / beta; double *v0 = VV(0); for (int i = 0; i < n; i++) v0[i] = v0[i] * a; }
            for (int i = 0; i < R + 2; i++) s[i] = 0.0;
            s[0] = beta;
        }
        double *vm1 = VV(m + 1), *zm = ZZ(m);
        kry_precond(&c, VV(m), zm);
        orc_spmv(n, rp, ci, va, zm, vm1);
        const int sm = m > K ? m - K : 0;
        for (int i = sm; i <= m; i++) {
            const double h = orc_dot(n, VV(i), vm1);
            HH(i, m) = h;
            kry_axpy(n, VV(i), vm1, -h);
        }
        HH(m + 1, m) = orc_nrm2(n, vm1);
        { const double a = 1.0 / HH(m + 1, m); for (int k = 0; k < n; k++) vm1[k] = vm1[k] * a; }
        gamma[m] = s[m];
        for (int k = 0; k < m; k++) {
            const double t = cs[k] * HH(k, m) + sn[k] * HH(k + 1, m);
            HH(k + 1, m) = -sn[k] * HH(k, m) + cs[k] * HH(k + 1, m);
            HH(k, m) = t;
        }
        gen_rot(HH(m, m), HH(m + 1, m), &cs[m], &sn[m]);
        HH(m, m) = cs[m] * HH(m, m) + sn[m] * HH(m + 1, m);
        HH(m + 1, m) = 0.0;
        { const double t = cs[m] * s[m]; s[m + 1] = -sn[m] * s[m]; s[m] = t; }
        for (int i = sm; i < m; i++) kry_axpy(n, ZZ(i), zm, -HH(i, m));
        { const double a = 1.0 / HH(m, m); for (int k = 0; k < n; k++) zm[k] = zm[k] * a; }
        kry_axpy(n, zm, x, s[m]);
        if (m == 0) kry_axpby(n, VV(1), VV(0), resid, s[1] * cs[0], -1.0 * s[1] * sn[0]);
        else kry_axpby(n, vm1, resid, resid, s[m + 1] * cs[m], -1.0 * s[m + 1] * sn[m] / gamma[m]);
        nrm = orc_nrm2(n, resid);
        res_hist[it + 1] = nrm;
        if (conv_relative_ini(nrm, nrm_ini, tol)) { conv = 1; done = 1; it++; break; }
    }
fin:
    if (converged_out) *converged_out = conv;
    for (int i = 0; i < NV; i++) free(V[i]);
    for (int i = 0; i < NZ; i++) free(Z[i]);
    free(V); free(Z); free(H); free(s); free(cs); free(sn); free(gamma); free(r); free(resid); free(c.dj);
#undef HH
#undef VV
#undef ZZ
    return it;
}
