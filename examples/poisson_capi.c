/* poisson_capi.c -- a plain-C caller of the AMGX C API as exported by libamgx_b200.so (alias libamgxsh.so).
 * It uses only names that exist in the reference's include/amgx_c.h, so the same file builds against the reference:
 *     gcc examples/poisson_capi.c -Iinclude -Lamgx_b200 -lamgxsh -Wl,-rpath,$PWD/amgx_b200 -lm -o poisson_capi
 *     ./poisson_capi 64 amgx_b200/configs/PCG_AGGREGATION_JACOBI.json
 * Builds the 7-point Poisson matrix on an n^3 grid (diagonal first, the entry order of the reference's generator),
 * solves A x = 1 from x = 0 and prints the iteration count, the status and the residual history. */
#include <stdio.h>
#include <stdlib.h>
#include "amgx_b200.h"

#define CK(call)                                                                            \
    do {                                                                                    \
        AMGX_RC rc_ = (call);                                                               \
        if (rc_ != AMGX_RC_OK) {                                                            \
            char msg_[512];                                                                 \
            AMGX_get_error_string(rc_, msg_, 512);                                          \
            fprintf(stderr, "%s failed: %s\n", #call, msg_);                                \
            exit(1);                                                                        \
        }                                                                                   \
    } while (0)

static void print_cb(const char *msg, int length) { fwrite(msg, 1, (size_t)length, stdout); }

int main(int argc, char **argv)
{
    const int nx = argc > 1 ? atoi(argv[1]) : 32;
    const char *cfg_file = argc > 2 ? argv[2] : "amgx_b200/configs/PCG_AGGREGATION_JACOBI.json";
    const int n = nx * nx * nx;
    int *rp = (int *)malloc(sizeof(int) * ((size_t)n + 1)), *ci = (int *)malloc(sizeof(int) * (size_t)n * 7);
    double *va = (double *)malloc(sizeof(double) * (size_t)n * 7), *b = (double *)malloc(sizeof(double) * (size_t)n), *x = (double *)malloc(sizeof(double) * (size_t)n);
    int nnz = 0;
    for (int r = 0; r < n; r++) {
        const int i = r % nx, j = (r / nx) % nx, k = r / (nx * nx);
        rp[r] = nnz;
        ci[nnz] = r; va[nnz++] = 6.0;
        if (i > 0) { ci[nnz] = r - 1; va[nnz++] = -1.0; }
        if (i < nx - 1) { ci[nnz] = r + 1; va[nnz++] = -1.0; }
        if (j > 0) { ci[nnz] = r - nx; va[nnz++] = -1.0; }
        if (j < nx - 1) { ci[nnz] = r + nx; va[nnz++] = -1.0; }
        if (k > 0) { ci[nnz] = r - nx * nx; va[nnz++] = -1.0; }
        if (k < nx - 1) { ci[nnz] = r + nx * nx; va[nnz++] = -1.0; }
        b[r] = 1.0;
        x[r] = 0.0;
    }
    rp[n] = nnz;

    AMGX_config_handle cfg;
    AMGX_resources_handle rsrc;
    AMGX_matrix_handle A;
    AMGX_vector_handle vb, vx;
    AMGX_solver_handle solver;
    CK(AMGX_initialize());
    CK(AMGX_register_print_callback(&print_cb));
    CK(AMGX_config_create_from_file(&cfg, cfg_file));
    CK(AMGX_config_add_parameters(&cfg, "config_version=2, main:store_res_history=1, main:monitor_residual=1"));
    CK(AMGX_resources_create_simple(&rsrc, cfg));
    CK(AMGX_matrix_create(&A, rsrc, AMGX_mode_dDDI));
    CK(AMGX_vector_create(&vb, rsrc, AMGX_mode_dDDI));
    CK(AMGX_vector_create(&vx, rsrc, AMGX_mode_dDDI));
    CK(AMGX_solver_create(&solver, rsrc, AMGX_mode_dDDI, cfg));
    CK(AMGX_matrix_upload_all(A, n, nnz, 1, 1, rp, ci, va, NULL));
    CK(AMGX_vector_upload(vb, n, 1, b));
    CK(AMGX_vector_upload(vx, n, 1, x));
    CK(AMGX_solver_setup(solver, A));
    CK(AMGX_solver_solve(solver, vb, vx));
    AMGX_SOLVE_STATUS st;
    int iters = 0;
    CK(AMGX_solver_get_status(solver, &st));
    CK(AMGX_solver_get_iterations_number(solver, &iters));
    CK(AMGX_vector_download(vx, x));
    printf("status %d iterations %d\n", (int)st, iters);
    for (int it = 0; it <= iters; it++) {
        double r;
        if (AMGX_solver_get_iteration_residual(solver, it, 0, &r) == AMGX_RC_OK) printf("  %3d  %.6e\n", it, r);
    }
    printf("x[0] = %.12e  x[n/2] = %.12e\n", x[0], x[n / 2]);
    CK(AMGX_solver_destroy(solver));
    CK(AMGX_vector_destroy(vx));
    CK(AMGX_vector_destroy(vb));
    CK(AMGX_matrix_destroy(A));
    CK(AMGX_resources_destroy(rsrc));
    CK(AMGX_config_destroy(cfg));
    CK(AMGX_finalize());
    free(rp); free(ci); free(va); free(b); free(x);
    return st == AMGX_SOLVE_SUCCESS ? 0 : 2;
}
