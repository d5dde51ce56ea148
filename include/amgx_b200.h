/*
 * amgx_b200.h -- C ABI of libamgx_b200.so, the B200-native AMG solve-phase engine.
 *
 * Every AMGX_* entry point below has the SAME name, argument list, argument meaning and
 * return-code behaviour as the reference's C API so that an application (or pyamgx /
 * AmgXWrapper / AMGX.jl style binding) built against the reference's amgx_c.h can be
 * re-linked against this library unchanged.  The reference declaration each one replaces is
 * cited as  [ref: include/amgx_c.h:<line>]  (paths relative to /root/reference).
 * Numeric values of the enums follow include/amgx_c.h:51-103 and include/amgx_config.h:13-124.
 *
 * Plain pointers and sizes only; no C++/torch types cross this boundary.
 * AMGXB200_* symbols are extensions of this library (multi-GPU bootstrap, introspection,
 * micro-benchmarks of the hot kernels); a caller of the reference never needs them.
 */
#ifndef AMGX_B200_H
#define AMGX_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__cplusplus)
extern "C" {
#endif

#define AMGX_API __attribute__((visibility("default")))

/* ---- return codes [ref: include/amgx_c.h:51-69] ---- */
typedef enum {
    AMGX_RC_OK = 0, AMGX_RC_BAD_PARAMETERS = 1, AMGX_RC_UNKNOWN = 2, AMGX_RC_NOT_SUPPORTED_TARGET = 3,
    AMGX_RC_NOT_SUPPORTED_BLOCKSIZE = 4, AMGX_RC_CUDA_FAILURE = 5, AMGX_RC_THRUST_FAILURE = 6,
    AMGX_RC_NO_MEMORY = 7, AMGX_RC_IO_ERROR = 8, AMGX_RC_BAD_MODE = 9, AMGX_RC_CORE = 10,
    AMGX_RC_PLUGIN = 11, AMGX_RC_BAD_CONFIGURATION = 12, AMGX_RC_NOT_IMPLEMENTED = 13,
    AMGX_RC_LICENSE_NOT_FOUND = 14, AMGX_RC_INTERNAL = 15
} AMGX_RC;

/* ---- solve status [ref: include/amgx_c.h:74-80] ---- */
typedef enum {
    AMGX_SOLVE_SUCCESS = 0, AMGX_SOLVE_FAILED = 1, AMGX_SOLVE_DIVERGED = 2, AMGX_SOLVE_NOT_CONVERGED = 3
} AMGX_SOLVE_STATUS;

/* [ref: include/amgx_c.h:85-91] */
typedef enum {
    AMGX_GET_PARAMS_DESC_JSON_TO_FILE = 0, AMGX_GET_PARAMS_DESC_JSON_TO_STRING = 1,
    AMGX_GET_PARAMS_DESC_TEXT_TO_FILE = 2, AMGX_GET_PARAMS_DESC_TEXT_TO_STRING = 3
} AMGX_GET_PARAMS_DESC_FLAG;

/* [ref: include/amgx_c.h:96-100] */
typedef enum { AMGX_DIST_PARTITION_VECTOR = 0, AMGX_DIST_PARTITION_OFFSETS = 1 } AMGX_DIST_PARTITION_INFO;

/* ---- modes [ref: include/amgx_config.h:102-124]: mem + 16*vec + 256*mat + 4096*ind ---- */
typedef enum {
    AMGX_unset    = -1,
    AMGX_mode_hDDI = 8192, AMGX_mode_hDFI = 8448, AMGX_mode_hFFI = 8464,
    AMGX_mode_dDDI = 8193, AMGX_mode_dDFI = 8449, AMGX_mode_dFFI = 8465
} AMGX_Mode;

/* ---- opaque handles [ref: include/amgx_c.h:105-124] ---- */
typedef void (*AMGX_print_callback)(const char *msg, int length);
typedef struct AMGX_config_handle_struct       *AMGX_config_handle;
typedef struct AMGX_resources_handle_struct    *AMGX_resources_handle;
typedef struct AMGX_matrix_handle_struct       *AMGX_matrix_handle;
typedef struct AMGX_vector_handle_struct       *AMGX_vector_handle;
typedef struct AMGX_solver_handle_struct       *AMGX_solver_handle;
typedef struct AMGX_distribution_handle_struct *AMGX_distribution_handle;

/* ---- build / init / system [ref: include/amgx_c.h:150-190] ---- */
AMGX_RC AMGX_API AMGX_get_api_version(int *major, int *minor);   /* :150 */
AMGX_RC AMGX_API AMGX_get_build_info_strings(char **version, char **date, char **time);   /* :154 */
AMGX_RC AMGX_API AMGX_get_error_string(AMGX_RC err, char *buf, int buf_len);   /* :159 */
AMGX_RC AMGX_API AMGX_initialize(void);   /* :165 */
AMGX_RC AMGX_API AMGX_initialize_plugins(void);   /* :167 */
AMGX_RC AMGX_API AMGX_finalize(void);   /* :169 */
AMGX_RC AMGX_API AMGX_finalize_plugins(void);   /* :171 */
void    AMGX_API AMGX_abort(AMGX_resources_handle rsrc, int err);   /* :173 */
AMGX_RC AMGX_API AMGX_pin_memory(void *ptr, unsigned int bytes);   /* :178 */
AMGX_RC AMGX_API AMGX_unpin_memory(void *ptr);   /* :182 */
AMGX_RC AMGX_API AMGX_install_signal_handler(void);   /* :185 */
AMGX_RC AMGX_API AMGX_reset_signal_handler(void);   /* :187 */
AMGX_RC AMGX_API AMGX_register_print_callback(AMGX_print_callback func);   /* :189 */

/* ---- config [ref: include/amgx_c.h:193-215] ---- */
AMGX_RC AMGX_API AMGX_config_create(AMGX_config_handle *cfg, const char *options);   /* :193 */
AMGX_RC AMGX_API AMGX_config_add_parameters(AMGX_config_handle *cfg, const char *options);   /* :197 */
AMGX_RC AMGX_API AMGX_config_create_from_file(AMGX_config_handle *cfg, const char *param_file);   /* :201 */
AMGX_RC AMGX_API AMGX_config_create_from_file_and_string(AMGX_config_handle *cfg, const char *param_file, const char *options);   /* :205 */
AMGX_RC AMGX_API AMGX_config_get_default_number_of_rings(AMGX_config_handle cfg, int *num_import_rings);   /* :210 */
AMGX_RC AMGX_API AMGX_config_destroy(AMGX_config_handle cfg);   /* :214 */

/* ---- resources [ref: include/amgx_c.h:218-230].
 * `comm`: the reference dereferences it as MPI_Comm*.  This image has no MPI; here `comm` is
 * NULL (single process, single GPU) or a pointer to an AMGXB200_comm (below): rank, world
 * size and an ncclUniqueId obtained with AMGXB200_get_nccl_unique_id on rank 0 and
 * broadcast by the launcher (torch.distributed in bench.py / tests). ---- */
AMGX_RC AMGX_API AMGX_resources_create(AMGX_resources_handle *rsc, AMGX_config_handle cfg, void *comm, int device_num, const int *devices);   /* :218 */
AMGX_RC AMGX_API AMGX_resources_create_simple(AMGX_resources_handle *rsc, AMGX_config_handle cfg);   /* :225 */
AMGX_RC AMGX_API AMGX_resources_destroy(AMGX_resources_handle rsc);   /* :229 */

/* ---- distribution [ref: include/amgx_c.h:235-259] ---- */
AMGX_RC AMGX_API AMGX_distribution_create(AMGX_distribution_handle *dist, AMGX_config_handle cfg);   /* :235 */
AMGX_RC AMGX_API AMGX_distribution_destroy(AMGX_distribution_handle dist);   /* :238 */
AMGX_RC AMGX_API AMGX_distribution_set_partition_data(AMGX_distribution_handle dist, AMGX_DIST_PARTITION_INFO info, const void *partition_data);   /* :251 */
AMGX_RC AMGX_API AMGX_distribution_set_32bit_colindices(AMGX_distribution_handle dist, int use32bit);   /* :258 */

/* ---- matrix [ref: include/amgx_c.h:262-333] ---- */
AMGX_RC AMGX_API AMGX_matrix_create(AMGX_matrix_handle *mtx, AMGX_resources_handle rsc, AMGX_Mode mode);   /* :262 */
AMGX_RC AMGX_API AMGX_matrix_destroy(AMGX_matrix_handle mtx);   /* :267 */
AMGX_RC AMGX_API AMGX_matrix_upload_all(AMGX_matrix_handle mtx, int n, int nnz, int block_dimx, int block_dimy,
                                        const int *row_ptrs, const int *col_indices, const void *data, const void *diag_data);   /* :270 */
AMGX_RC AMGX_API AMGX_matrix_replace_coefficients(AMGX_matrix_handle mtx, int n, int nnz, const void *data, const void *diag_data);   /* :281 */
AMGX_RC AMGX_API AMGX_matrix_get_size(const AMGX_matrix_handle mtx, int *n, int *block_dimx, int *block_dimy);   /* :288 */
AMGX_RC AMGX_API AMGX_matrix_get_nnz(const AMGX_matrix_handle mtx, int *nnz);   /* :294 */
AMGX_RC AMGX_API AMGX_matrix_download_all(const AMGX_matrix_handle mtx, int *row_ptrs, int *col_indices, void *data, void **diag_data);   /* :298 */
AMGX_RC AMGX_API AMGX_matrix_vector_multiply(AMGX_matrix_handle mtx, AMGX_vector_handle x, AMGX_vector_handle y);   /* :305 */
AMGX_RC AMGX_API AMGX_matrix_set_boundary_separation(AMGX_matrix_handle mtx, int boundary_separation);   /* :310 */
AMGX_RC AMGX_API AMGX_matrix_comm_from_maps(AMGX_matrix_handle mtx, int allocated_halo_depth, int num_import_rings, int max_num_neighbors,
                                            const int *neighbors, const int *send_ptrs, const int *send_maps, const int *recv_ptrs, const int *recv_maps);   /* :314 */
AMGX_RC AMGX_API AMGX_matrix_comm_from_maps_one_ring(AMGX_matrix_handle mtx, int allocated_halo_depth, int num_neighbors, const int *neighbors,
                                                     const int *send_sizes, const int **send_maps, const int *recv_sizes, const int **recv_maps);   /* :325 */

/* ---- vector [ref: include/amgx_c.h:336-370] ---- */
AMGX_RC AMGX_API AMGX_vector_create(AMGX_vector_handle *vec, AMGX_resources_handle rsc, AMGX_Mode mode);   /* :336 */
AMGX_RC AMGX_API AMGX_vector_destroy(AMGX_vector_handle vec);   /* :341 */
AMGX_RC AMGX_API AMGX_vector_upload(AMGX_vector_handle vec, int n, int block_dim, const void *data);   /* :344 */
AMGX_RC AMGX_API AMGX_vector_set_zero(AMGX_vector_handle vec, int n, int block_dim);   /* :350 */
AMGX_RC AMGX_API AMGX_vector_set_random(AMGX_vector_handle vec, int n);   /* :355 */
AMGX_RC AMGX_API AMGX_vector_download(const AMGX_vector_handle vec, void *data);   /* :359 */
AMGX_RC AMGX_API AMGX_vector_get_size(const AMGX_vector_handle vec, int *n, int *block_dim);   /* :363 */
AMGX_RC AMGX_API AMGX_vector_bind(AMGX_vector_handle vec, const AMGX_matrix_handle mtx);   /* :368 */

/* ---- solver [ref: include/amgx_c.h:373-605] ---- */
AMGX_RC AMGX_API AMGX_solver_create(AMGX_solver_handle *slv, AMGX_resources_handle rsc, AMGX_Mode mode, const AMGX_config_handle cfg_solver);   /* :373 */
AMGX_RC AMGX_API AMGX_solver_destroy(AMGX_solver_handle slv);   /* :379 */
AMGX_RC AMGX_API AMGX_solver_setup(AMGX_solver_handle slv, AMGX_matrix_handle mtx);   /* :382 */
AMGX_RC AMGX_API AMGX_solver_solve(AMGX_solver_handle slv, AMGX_vector_handle rhs, AMGX_vector_handle sol);   /* :386 */
AMGX_RC AMGX_API AMGX_solver_solve_with_0_initial_guess(AMGX_solver_handle slv, AMGX_vector_handle rhs, AMGX_vector_handle sol);   /* :391 */
AMGX_RC AMGX_API AMGX_solver_get_iterations_number(AMGX_solver_handle slv, int *n);   /* :396 */
AMGX_RC AMGX_API AMGX_solver_get_iteration_residual(AMGX_solver_handle slv, int it, int idx, double *res);   /* :400 */
AMGX_RC AMGX_API AMGX_solver_get_status(AMGX_solver_handle slv, AMGX_SOLVE_STATUS *st);   /* :406 */
AMGX_RC AMGX_API AMGX_solver_calculate_residual_norm(AMGX_solver_handle solver, AMGX_matrix_handle mtx, AMGX_vector_handle rhs, AMGX_vector_handle x, void *norm_vector);   /* :410 */
AMGX_RC AMGX_API AMGX_solver_resetup(AMGX_solver_handle slv, AMGX_matrix_handle mtx);   /* :603 */
AMGX_RC AMGX_API AMGX_solver_register_print_callback(AMGX_print_callback func);   /* :600 */

/* ---- utilities [ref: include/amgx_c.h:418-595] ---- */
AMGX_RC AMGX_API AMGX_read_system(AMGX_matrix_handle mtx, AMGX_vector_handle rhs, AMGX_vector_handle sol, const char *filename);   /* :435 */
AMGX_RC AMGX_API AMGX_write_system(const AMGX_matrix_handle mtx, const AMGX_vector_handle rhs, const AMGX_vector_handle sol, const char *filename);   /* :418 */
AMGX_RC AMGX_API AMGX_read_system_distributed(AMGX_matrix_handle mtx, AMGX_vector_handle rhs, AMGX_vector_handle sol, const char *filename,
                                              int allocated_halo_depth, int num_partitions, const int *partition_sizes, int partition_vector_size,
                                              const int *partition_vector);   /* :441 */
AMGX_RC AMGX_API AMGX_write_system_distributed(const AMGX_matrix_handle mtx, const AMGX_vector_handle rhs, const AMGX_vector_handle sol, const char *filename,
                                               int allocated_halo_depth, int num_partitions, const int *partition_sizes, int partition_vector_size,
                                               const int *partition_vector);   /* :424 */
AMGX_RC AMGX_API AMGX_read_system_maps_one_ring(int *n, int *nnz, int *block_dimx, int *block_dimy, int **row_ptrs, int **col_indices, void **data,
                                                void **diag_data, void **rhs, void **sol, int *num_neighbors, int **neighbors, int **send_sizes,
                                                int ***send_maps, int **recv_sizes, int ***recv_maps, AMGX_resources_handle rsc, AMGX_Mode mode,
                                                const char *filename, int allocated_halo_depth, int num_partitions, const int *partition_sizes,
                                                int partition_vector_size, const int *partition_vector);   /* :452 */
AMGX_RC AMGX_API AMGX_free_system_maps_one_ring(int *row_ptrs, int *col_indices, void *data, void *diag_data, void *rhs, void *sol, int num_neighbors,
                                                int *neighbors, int *send_sizes, int **send_maps, int *recv_sizes, int **recv_maps);   /* :478 */
AMGX_RC AMGX_API AMGX_read_system_global(int *n, int *nnz, int *block_dimx, int *block_dimy, int **row_ptrs, void **col_indices_global, void **data,
                                         void **diag_data, void **rhs, void **sol, AMGX_resources_handle rsc, AMGX_Mode mode, const char *filename,
                                         int allocated_halo_depth, int num_partitions, const int *partition_sizes, int partition_vector_size,
                                         const int *partition_vector);   /* :525 */
AMGX_RC AMGX_API AMGX_generate_distributed_poisson_7pt(AMGX_matrix_handle mtx, AMGX_vector_handle rhs, AMGX_vector_handle sol,
                                                       int allocated_halo_depth, int num_import_rings, int nx, int ny, int nz, int px, int py, int pz);   /* :492 */
AMGX_RC AMGX_API AMGX_write_parameters_description(char *filename, AMGX_GET_PARAMS_DESC_FLAG mode);   /* :505 */
AMGX_RC AMGX_API AMGX_matrix_attach_coloring(AMGX_matrix_handle mtx, int *row_coloring, int num_rows, int num_colors);   /* :512 */
AMGX_RC AMGX_API AMGX_matrix_attach_geometry(AMGX_matrix_handle mtx, double *geox, double *geoy, double *geoz, int n);   /* :518 */
AMGX_RC AMGX_API AMGX_matrix_upload_all_global(AMGX_matrix_handle mtx, int n_global, int n, int nnz, int block_dimx, int block_dimy,
                                               const int *row_ptrs, const void *col_indices_global, const void *data, const void *diag_data,
                                               int allocated_halo_depth, int num_import_rings, const int *partition_vector);   /* :545 */
AMGX_RC AMGX_API AMGX_matrix_upload_all_global_32(AMGX_matrix_handle mtx, int n_global, int n, int nnz, int block_dimx, int block_dimy,
                                                  const int *row_ptrs, const void *col_indices_global, const void *data, const void *diag_data,
                                                  int allocated_halo_depth, int num_import_rings, const int *partition_vector);   /* :560 */
AMGX_RC AMGX_API AMGX_matrix_upload_distributed(AMGX_matrix_handle mtx, int n_global, int n, int nnz, int block_dimx, int block_dimy,
                                                const int *row_ptrs, const void *col_indices_global, const void *data, const void *diag_data,
                                                AMGX_distribution_handle distribution);   /* :575 */
AMGX_RC AMGX_API AMGX_matrix_check_symmetry(AMGX_matrix_handle mtx, int *structurally_symmetric, int *symmetric);   /* :588 */
AMGX_RC AMGX_API AMGX_matrix_check_diag_dominant(const AMGX_matrix_handle mtx, int *diag_dominant);   /* :593 */

/* =====================================================================================
 * Extensions of this library (not in the reference).
 * ===================================================================================== */

/* What `void *comm` of AMGX_resources_create points to when world_size > 1. */
typedef struct {
    int  rank;
    int  world_size;
    char nccl_unique_id[128];   /* bytes of an ncclUniqueId (sizeof == 128) */
} AMGXB200_comm;

/* rank 0: fill `id` (128 bytes) with a fresh ncclUniqueId to be broadcast to all ranks. */
AMGX_RC AMGX_API AMGXB200_get_nccl_unique_id(char *id128);

/* Hierarchy introspection after AMGX_solver_setup (used by the parity tests).
 * level 0 = finest.  Arrays are copied to host buffers supplied by the caller (may be NULL to
 * query sizes only). */
AMGX_RC AMGX_API AMGXB200_solver_get_num_levels(AMGX_solver_handle slv, int *num_levels);
AMGX_RC AMGX_API AMGXB200_solver_get_level_info(AMGX_solver_handle slv, int level, int *n, int *nnz, int *block_dim, int *n_coarse);
AMGX_RC AMGX_API AMGXB200_solver_get_level_matrix(AMGX_solver_handle slv, int level, int *row_ptrs, int *col_indices, void *values);
/* aggregation: aggregates[n]; R_row_offsets[n_coarse+1]; R_column_indices[n] */
AMGX_RC AMGX_API AMGXB200_solver_get_level_aggregates(AMGX_solver_handle slv, int level, int *aggregates, int *R_row_offsets, int *R_column_indices);
/* classical: P (n x n_coarse) CSR and R (n_coarse x n) CSR; query nnz with NULL arrays */
AMGX_RC AMGX_API AMGXB200_solver_get_level_P(AMGX_solver_handle slv, int level, int *nnz, int *row_ptrs, int *col_indices, void *values);
AMGX_RC AMGX_API AMGXB200_solver_get_level_R(AMGX_solver_handle slv, int level, int *nnz, int *row_ptrs, int *col_indices, void *values);
/* classical: cf_map[n] (coarse index or <0 for fine points) */
AMGX_RC AMGX_API AMGXB200_solver_get_level_cf_map(AMGX_solver_handle slv, int level, int *cf_map);
/* smoother data of a level (Jacobi: d[n]; L1: d_L1[n]; DILU: Einv[n*b*b]) and colouring */
AMGX_RC AMGX_API AMGXB200_solver_get_level_smoother_data(AMGX_solver_handle slv, int level, void *data);
AMGX_RC AMGX_API AMGXB200_solver_get_level_coloring(AMGX_solver_handle slv, int level, int *num_colors, int *row_colors);

/* Timing of the last solve measured with CUDA events on the solve stream (seconds) and the number
 * of kernels this library launched in it. */
AMGX_RC AMGX_API AMGXB200_solver_get_last_solve_stats(AMGX_solver_handle slv, double *solve_seconds, long long *kernel_launches);

/* Does the engine provide every component `cfg` names (solvers, smoothers, cycles, selectors, colouring ...)?  AMGX_RC_OK, or the code and
 * message AMGX_solver_create would produce.  Pure host code: usable without a GPU to vet a configuration before moving to the engine. */
AMGX_RC AMGX_API AMGXB200_config_check(const AMGX_config_handle cfg, AMGX_Mode mode, char *msg, int msg_len);

/* Hot-kernel micro-benchmarks on an uploaded matrix (device-resident operands).
 * kind: 0 = SpMV y=A*x, 1 = fused Jacobi sweep x' = x + w*(b-A*x)/d, 2 = SpMV fused with dot.
 * Runs `reps` launches on the resource stream, returns average milliseconds per launch measured
 * with CUDA events on that stream, after `warmup` untimed launches. */
AMGX_RC AMGX_API AMGXB200_bench_kernel(AMGX_matrix_handle mtx, int kind, int warmup, int reps, int flush_l2, double *avg_ms);

/* Which scalar CSR kernel family an uploaded matrix was planned for: rows per tile of the TMA tile kernels (0: the fallback kernels),
 * tiles with coded column streams, with pair tables, with row patterns (k_spmv_enc.cu), and the ring size of the sliding x window of the
 * banded-matrix kernel (k_spmv_win.cu; 0 = not used).  Any pointer may be NULL.  Extension: the reference exposes no such query. */
AMGX_RC AMGX_API AMGXB200_matrix_get_kernel_plan(AMGX_matrix_handle mtx, int *tile_rows, int *coded_tiles, int *pair_tiles, int *row_pattern_tiles, int *window);

/* Partition planner (pure host code, needs no GPU): given this rank's rows of a global CSR
 * (global column ids, contiguous row partition by offsets[world+1]) computes the local
 * renumbering [interior | boundary | halo], the per-neighbour send maps (B2L) and the halo
 * layout.  Arrays are malloc'ed by the library; free with AMGXB200_partition_plan_free. */
typedef struct {
    int  n_owned, n_interior, n_halo, num_neighbors;
    int *neighbors;        /* [num_neighbors] ranks */
    int *send_offsets;     /* [num_neighbors+1] into send_maps */
    int *send_maps;        /* local (renumbered) owned row ids to pack for each neighbour */
    int *halo_offsets;     /* [num_neighbors+1] offsets (relative to n_owned) of each neighbour's halo segment */
    int64_t *halo_global;  /* [n_halo] global ids of halo columns in local order */
    int *perm_old_to_new;  /* [n_owned] local row i (partition order) -> renumbered id */
    int *local_cols;       /* [nnz] column ids in the renumbered local space (halo ids >= n_owned) */
} AMGXB200_partition_plan;
AMGX_RC AMGX_API AMGXB200_partition_plan_create(AMGXB200_partition_plan *plan, int rank, int world_size, const int64_t *offsets,
                                                int n, int nnz, const int *row_ptrs, const int64_t *col_indices_global);
void    AMGX_API AMGXB200_partition_plan_free(AMGXB200_partition_plan *plan);
/* Arbitrary partition vector (AMGX_DIST_PARTITION_VECTOR, include/amgx_c.h:241-259) -> the contiguous numbering the engine works
 * in, exactly the reference's ipartition_map (src/distributed/distributed_manager.cu:1175-1203): offsets[world_size+1],
 * new_global[n_global] (may be NULL).  Pure host code. */
/* Caller-supplied comm maps (AMGX_matrix_comm_from_maps_one_ring) -> global column ids of a matrix given in local numbering (owned
 * columns < n, halo columns >= n).  recv_global[q][k] = global id of the k-th value neighbour q sends (its send_maps row k, offset by
 * q's first global row).  Pure host code; returns AMGX_RC_BAD_PARAMETERS when the maps do not cover the halo columns. */
AMGX_RC AMGX_API AMGXB200_comm_maps_to_global_cols(int n, int nnz, const int *local_cols, int64_t my_offset, int num_neighbors, const int *recv_sizes,
                                                   const int *const *recv_maps, const int64_t *const *recv_global, int64_t *cols_out);
/* AMGX_read_system_maps_one_ring + AMGX_read_system_global for an explicit (rank, world_size) pair: no resources handle, no GPU.  Any of
 * col_indices_local (with the neighbour / map outputs) and col_indices_global may be NULL.  Free with AMGX_free_system_maps_one_ring (+ free()). */
AMGX_RC AMGX_API AMGXB200_read_system_partition(int rank, int world_size, AMGX_Mode mode, const char *filename, int num_partitions, const int *partition_sizes,
                                                int partition_vector_size, const int *partition_vector, int *n, int *nnz, int *block_dimx, int *block_dimy,
                                                int **row_ptrs, int **col_indices_local, int64_t **col_indices_global, void **data, void **diag_data,
                                                void **rhs, void **sol, int *num_neighbors, int **neighbors, int **send_sizes, int ***send_maps,
                                                int **recv_sizes, int ***recv_maps);
AMGX_RC AMGX_API AMGXB200_partition_vector_to_contiguous(int n_global, int world_size, const int *partition_vector, int64_t *offsets, int64_t *new_global);
/* The file writer behind AMGX_write_system[_distributed] for host arrays (fp64): writer = "matrixmarket" (the reference's layout,
 * src/matrix_io.cu:120-258) or "binary" (%%NVAMGBinary).  values holds nnz blocks followed by n diagonal blocks when ext_diag != 0;
 * rhs / sol may be NULL.  Pure host code, no GPU: lets the writer / reader pair be tested without a device. */
AMGX_RC AMGX_API AMGXB200_write_system_host(const char *filename, const char *writer, int n, int nnz, int block_dimx, int block_dimy, const int *row_ptrs,
                                            const int *col_indices, const double *values, int ext_diag, const double *rhs, const double *sol);

#if defined(__cplusplus)
}
#endif
#endif /* AMGX_B200_H */
