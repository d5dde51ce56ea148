/* poisson_dist_capi.c -- a plain-C multi-GPU caller of the C API: one process per GPU, no MPI and no Python in the loop.
 * The reference's examples/amgx_mpi_poisson7.c hands an MPI_Comm* to AMGX_resources_create (examples/amgx_mpi_poisson7.c:158-362); this
 * library takes an AMGXB200_comm* instead (rank, world size, ncclUniqueId: include/amgx_b200.h).  Everything after that call is the
 * reference's sequence unchanged: generate the partitioned 7-point Poisson problem (z-slabs), setup, solve, read status / residuals.
 *
 * The 128-byte id has to travel from rank 0 to the other ranks by whatever means the application has (MPI_Bcast in an MPI program,
 * INTEGRATION.md).  This example needs nothing but a shared directory: rank 0 writes the id to $AMGXB_ID_FILE, the others wait for it.
 *
 *     gcc examples/poisson_dist_capi.c -Iinclude -Lamgx_b200 -lamgxsh -Wl,-rpath,$PWD/amgx_b200 -o poisson_dist_capi
 *     for r in 0 1; do RANK=$r WORLD_SIZE=2 LOCAL_RANK=$r AMGXB_ID_FILE=/tmp/amgxb.id ./poisson_dist_capi 128 \
 *         amgx_b200/configs/PCG_AGGREGATION_JACOBI.json & done; wait
 * Rank r owns the nx x nx x nx box number r of the nx x nx x (nx * WORLD_SIZE) grid. */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "amgx_b200.h"

#define CK(call)                                                                            \
    do {                                                                                    \
        AMGX_RC rc_ = (call);                                                               \
        if (rc_ != AMGX_RC_OK) {                                                            \
            char msg_[512];                                                                 \
            AMGX_get_error_string(rc_, msg_, 512);                                          \
            fprintf(stderr, "[rank %d] %s failed: %s\n", g_rank, #call, msg_);              \
            exit(1);                                                                        \
        }                                                                                   \
    } while (0)

static int g_rank = 0;

static void print_cb(const char *msg, int length)
{
    if (g_rank == 0) fwrite(msg, 1, (size_t)length, stdout);
}

static int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

/* rank 0 -> file (written under a temporary name, then renamed: readers never see half an id); the others poll for it */
static int share_id(char *id128, int rank, const char *path)
{
    if (rank == 0) {
        char tmp[1024];
        snprintf(tmp, sizeof(tmp), "%s.tmp", path);
        FILE *f = fopen(tmp, "wb");
        if (!f || fwrite(id128, 1, 128, f) != 128) return 1;
        fclose(f);
        return rename(tmp, path);
    }
    for (int tries = 0; tries < 600; tries++) {          /* up to a minute */
        FILE *f = fopen(path, "rb");
        if (f) {
            const size_t got = fread(id128, 1, 128, f);
            fclose(f);
            if (got == 128) return 0;
        }
        nanosleep(&(struct timespec){0, 100000000L}, NULL);
    }
    return 1;
}

int main(int argc, char **argv)
{
    const int nx = argc > 1 ? atoi(argv[1]) : 64;
    const char *cfg_file = argc > 2 ? argv[2] : "amgx_b200/configs/PCG_AGGREGATION_JACOBI.json";
    const int world = env_int("WORLD_SIZE", 1);
    int device = env_int("LOCAL_RANK", 0);
    const char *id_file = getenv("AMGXB_ID_FILE");
    g_rank = env_int("RANK", 0);
    if (world > 1 && !id_file) {
        fprintf(stderr, "set AMGXB_ID_FILE to a path every rank can read\n");
        return 1;
    }

    AMGXB200_comm comm;
    memset(&comm, 0, sizeof(comm));
    comm.rank = g_rank;
    comm.world_size = world;

    AMGX_config_handle cfg;
    AMGX_resources_handle rsrc;
    AMGX_matrix_handle A;
    AMGX_vector_handle b, x;
    AMGX_solver_handle solver;
    CK(AMGX_initialize());
    CK(AMGX_register_print_callback(&print_cb));
    if (world > 1) {
        if (g_rank == 0) CK(AMGXB200_get_nccl_unique_id(comm.nccl_unique_id));
        if (share_id(comm.nccl_unique_id, g_rank, id_file) != 0) {
            fprintf(stderr, "[rank %d] could not exchange the communicator id through %s\n", g_rank, id_file);
            return 1;
        }
    }
    CK(AMGX_config_create_from_file(&cfg, cfg_file));
    CK(AMGX_config_add_parameters(&cfg, "config_version=2, main:store_res_history=1, main:monitor_residual=1"));
    CK(AMGX_resources_create(&rsrc, cfg, world > 1 ? (void *)&comm : NULL, 1, &device));
    CK(AMGX_matrix_create(&A, rsrc, AMGX_mode_dDDI));
    CK(AMGX_vector_create(&b, rsrc, AMGX_mode_dDDI));
    CK(AMGX_vector_create(&x, rsrc, AMGX_mode_dDDI));
    CK(AMGX_solver_create(&solver, rsrc, AMGX_mode_dDDI, cfg));
    /* z-slabs: px = py = 1, pz = world (src/amgx_c.cu:1700-1740: rhs = 1, the generator fills sol with ones) */
    CK(AMGX_generate_distributed_poisson_7pt(A, b, x, 1, 1, nx, nx, nx, 1, 1, world));
    CK(AMGX_vector_bind(b, A));
    CK(AMGX_vector_bind(x, A));
    int n = 0, bx = 0, by = 0;
    CK(AMGX_matrix_get_size(A, &n, &bx, &by));
    CK(AMGX_vector_set_zero(x, n, bx));                  /* x0 = 0, as examples/amgx_mpi_poisson7.c:277-292 does */
    CK(AMGX_solver_setup(solver, A));
    CK(AMGX_solver_solve_with_0_initial_guess(solver, b, x));
    AMGX_SOLVE_STATUS st;
    int iters = 0;
    CK(AMGX_solver_get_status(solver, &st));
    CK(AMGX_solver_get_iterations_number(solver, &iters));
    if (g_rank == 0) {
        printf("ranks %d local rows %d status %d iterations %d\n", world, n, (int)st, iters);
        for (int it = 0; it <= iters; it++) {
            double r;
            if (AMGX_solver_get_iteration_residual(solver, it, 0, &r) == AMGX_RC_OK) printf("  %3d  %.6e\n", it, r);
        }
    }
    CK(AMGX_solver_destroy(solver));
    CK(AMGX_vector_destroy(x));
    CK(AMGX_vector_destroy(b));
    CK(AMGX_matrix_destroy(A));
    CK(AMGX_resources_destroy(rsrc));
    CK(AMGX_config_destroy(cfg));
    CK(AMGX_finalize());
    if (g_rank == 0 && world > 1) remove(id_file);
    return st == AMGX_SOLVE_SUCCESS ? 0 : 2;
}
