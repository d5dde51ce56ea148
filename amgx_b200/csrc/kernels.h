// kernels.h -- launchers of the hand-written sm_100a kernels (k_spmv.cu, k_blas.cu, k_transfer.cu,
// k_setup_agg.cu).  All launchers enqueue on the given stream and never synchronise.
#pragma once
#include "base.h"
#include "matrix.h"

namespace amgxb {

// -------------------------------------------------------------------------------------------
// Device-side scalar block shared by the Krylov solvers: inner products, step lengths and norms
// live in device memory so that an iteration needs no host round trip except the convergence
// check.  One KrylovScalars per solver instance.
// -------------------------------------------------------------------------------------------
enum ScalarSlot : int {
    S_RZ = 0, S_RZ_OLD = 1, S_DOT = 2, S_ALPHA = 3, S_NEG_ALPHA = 4, S_BETA = 5, S_NRM = 6, S_TMP0 = 7,
    S_TMP1 = 8, S_ONE = 9, S_ZERO = 10, S_COUNT = 32
};

// What the last CTA of a reduction kernel does with the finished sum.
enum FinOp : int {
    FIN_STORE = 0,      // scal[slot] = sum
    FIN_SQRT = 1,       // scal[slot] = sqrt(sum)
    FIN_PCG_ALPHA = 2,  // S_DOT = sum; S_ALPHA = S_DOT != 0 ? S_RZ / S_DOT : 0; S_NEG_ALPHA = -S_ALPHA   (pcg_solver.cu:128-137)
    FIN_PCG_BETA = 3,   // S_RZ_OLD = S_RZ; S_RZ = sum; S_BETA = S_RZ_OLD != 0 ? S_RZ / S_RZ_OLD : 0        (pcg_solver.cu:172-182)
    FIN_ABS = 4,        // scal[slot] = sum (L1 norm; identical to STORE, kept for readability)
    FIN_ADD = 5,        // scal[slot] += sum  (second row segment of a distributed matrix: interior + boundary partials)
};

struct ReduceCtx {          // one per solver; buffers sized for the largest grid we launch
    double *partials = nullptr;     // [max_blocks]
    unsigned *counter = nullptr;    // zero-initialised; reset by the last block
    double *scal = nullptr;         // KrylovScalars, S_COUNT doubles
    double *host_mirror = nullptr;  // pinned, mapped: last block also writes scal[slot] here if mirror_slot>=0
};

// -------------------------------------------------------------------------------------------
// Scalar CSR family (block size 1).  One kernel body, several epilogues:
//   y = A x | y = b - A x | y = x + w (b - A x)/d  (+ optional fused dot / norm reductions)
// `agg` != nullptr makes the kernel read x through a prolongation on the fly: x(j) := xc[agg[j]]
// (aggregation AMG: P is piecewise constant, aggregation_amg_level.cu:156-181).
// -------------------------------------------------------------------------------------------
enum CsrEpi : int {
    EPI_SPMV = 0,        // y_i = (A x)_i
    EPI_RESID = 1,       // y_i = b_i - (A x)_i
    EPI_JACOBI = 2,      // y_i = x_i + ((b_i - (A x)_i) * w) * (1/d_i)      (block_jacobi_solver.cu:32-50)
    EPI_SPMV_DOT = 3,    // EPI_SPMV and reduce sum_i y_i * x_i             (PCG: <Ap,p>)
    EPI_JACOBI_DOT = 4,  // EPI_JACOBI and reduce sum_i b_i * y_i           (PCG: <r,z> on the last sweep)
    EPI_RESID_NRM2 = 5,  // EPI_RESID and reduce sum_i y_i^2
    EPI_JACOBI_L1 = 6,   // same arithmetic as EPI_JACOBI with d = L1 row norm (jacobi_l1_solver.cu:27-44)
    EPI_ADD = 7,         // y_i = b_i + (A x)_i   (classical prolongation x + P e, classical_amg_level.cu:884-910)
};

struct CsrOpArgs {
    const void *x = nullptr;      // gather source (length n_cols, or n_coarse with agg)
    const int *agg = nullptr;
    const void *b = nullptr;
    const void *d = nullptr;      // diagonal (mat precision)
    void *y = nullptr;
    double omega = 1.0;
    // reduction
    ReduceCtx red;
    int fin_op = FIN_STORE;
    int fin_slot = S_TMP0;
    int mirror = 0;               // write final value to red.host_mirror[fin_slot]
};

void csr_op(const Matrix &A, CsrEpi epi, const CsrOpArgs &args, cudaStream_t s, int segment = 0);
// Build diag_idx and the tile plan of A (called once per matrix at upload / level creation).
void csr_build_plan(Matrix &A, cudaStream_t s);
int  csr_max_grid(const Matrix &A);     // number of CTAs csr_op launches (partials sizing)
// experimental compressed column stream (k_spmv_enc.cu; AMGXB_COLENC=1, default off)
void csr_build_colenc(Matrix &A, cudaStream_t s);                                                     // after csr_build_plan
void csr_values_changed(Matrix &A, cudaStream_t s);                                                   // after an in-place change of A.values (no-op unless value codes exist)
bool csr_op_enc(const Matrix &A, CsrEpi epi, const CsrOpArgs &args, cudaStream_t s, int segment);    // false: use the plain kernels
// sliding x window in shared memory for banded irregular matrices (k_spmv_win.cu; AMGXB_WINDOW=0 disables)
void csr_build_window(Matrix &A, cudaStream_t s);                                                     // after csr_build_colenc
bool csr_op_win(const Matrix &A, CsrEpi epi, const CsrOpArgs &args, cudaStream_t s, int segment);    // false: use the coded / plain kernels
void csr_window_values_changed(Matrix &A, cudaStream_t s);                                            // its sliced-ELL copy of the values follows in-place changes

// -------------------------------------------------------------------------------------------
// Level-1 kernels (k_blas.cu).  Vectors are VecT arrays of length n; scalars come from device
// memory (scal[slot]) so no host synchronisation is needed.
// -------------------------------------------------------------------------------------------
void vec_fill(void *x, Prec p, size_t n, double v, cudaStream_t s);
void vec_copy(void *dst, const void *src, Prec p, size_t n, cudaStream_t s);
// y = a*x + b*y style updates with host scalars (exact reference op order: x*a + y*b)
void vec_axpby(const void *x, const void *y, void *out, Prec p, size_t n, double a, double b, cudaStream_t s);
void vec_axpy(const void *x, void *y, Prec p, size_t n, double a, cudaStream_t s);          // y = a*x + y
void vec_axpbypcz(const void *x, const void *y, const void *z, void *out, Prec p, size_t n, double a, double b, double c, cudaStream_t s);   // out = x*a + y*b + z*c
void vec_scal(void *x, Prec p, size_t n, double a, cudaStream_t s);
// device-scalar variants: a = sign * scal[slot]
void vec_axpy_dev(const void *x, void *y, Prec p, size_t n, const double *scal, int slot, double sign, cudaStream_t s);
// scaling = DIAGONAL_SYMMETRIC (k_blas.cu)
int  diag_sym_scale_setup(const Matrix &A, DevVec &scale, cudaStream_t s);           // scale_i = 1/sqrt(a_ii); returns 1 if a diagonal entry is negative
void diag_sym_scale_matrix(Matrix &A, const DevVec &scale, bool unscale, cudaStream_t s);
void vec_scale_entrywise(void *v, const void *d, Prec p, size_t n, bool divide, cudaStream_t s);
void scalar_error_scale(const double *scal, int slot_nom, int slot_den, double *out, cudaStream_t s);   // error_scaling 2/3: clamped nom/den
// fused MGS step: y += sign*scal[slot]*x ; scal[fin_slot] = fin(<z, y>)  (z == nullptr: <y, y>)
void vec_axpy_dot_dev(const void *x, void *y, const void *z, Prec p, size_t n, const double *scal, int slot, double sign, const ReduceCtx &red, int fin_op,
                      int fin_slot, int mirror, cudaStream_t s);
void vec_axpby_dev(const void *x, const void *y, void *out, Prec p, size_t n, double a, const double *scal, int slot_b, cudaStream_t s);
void vec_scal_dev_inv(void *x, Prec p, size_t n, const double *scal, int slot, cudaStream_t s);  // x *= 1/scal[slot]
// reductions -> red.scal[fin_slot] through fin_op
void vec_dot(const void *x, const void *y, Prec p, size_t n, const ReduceCtx &red, int fin_op, int fin_slot, int mirror, cudaStream_t s);
void vec_nrm1(const void *x, Prec p, size_t n, const ReduceCtx &red, int fin_slot, int mirror, cudaStream_t s);
void vec_nrmmax(const void *x, Prec p, size_t n, const ReduceCtx &red, int fin_slot, int mirror, cudaStream_t s);
// PCG fused update: x += alpha p ; r -= alpha Ap ; nrm = ||r||_2 (or L1 / LMAX)  -- alpha = scal[S_ALPHA]
void pcg_update_xr(const void *p, const void *Ap, void *x, void *r, Prec pr, size_t n, const ReduceCtx &red, int norm_type,
                   int fin_slot, int mirror, cudaStream_t s, bool partial = false);   // partial: leave the un-finalised sum (distributed)
// Jacobi with zero initial guess: x = b*w/d   (block_jacobi_solver.cu:24-30)
void jacobi_zero_guess(const void *b, const void *d, void *x, Prec matp, Prec vecp, size_t n, double omega, cudaStream_t s);
int  blas_max_grid();

// -------------------------------------------------------------------------------------------
// Aggregation transfer operators (k_transfer.cu)
// -------------------------------------------------------------------------------------------
void agg_restrict(const int *R_row_offsets, const int *R_col, const void *r, void *rc, Prec p, int n_agg, int bsize, cudaStream_t s);
void agg_prolong_add(const int *aggregates, const void *e, const void *x, void *xout, Prec p, int n, int bsize, cudaStream_t s);   // xout = x + P e
void agg_prolong_set(const int *aggregates, const void *e, void *x, Prec p, int n, int bsize, cudaStream_t s);   // x = P e (x was zero)

// -------------------------------------------------------------------------------------------
// Aggregation setup (k_setup_agg.cu): SIZE_2 selector, R pattern, Galerkin product
// -------------------------------------------------------------------------------------------
struct AggSetupParams {
    int deterministic = 1;
    int max_iterations = 15;
    double max_unassigned = 0.05;
    int merge_singletons = 1;
    int weight_formula = 0;
    int edge_weight_component = 0;
    int two_phase = 0;
};
// aggregates[n] (renumbered), returns number of aggregates
int  size2_select(const Matrix &A, const AggSetupParams &prm, DevBuf<int> &aggregates, cudaStream_t s);
int  size4_select(const Matrix &A, const AggSetupParams &prm, DevBuf<int> &aggregates, cudaStream_t s);   // pairs of pairs (size4_selector.cu)
void build_restriction(const DevBuf<int> &aggregates, int n, int n_agg, DevBuf<int> &R_row_offsets, DevBuf<int> &R_col, cudaStream_t s);
void galerkin_aggregation(const Matrix &A, const DevBuf<int> &aggregates, int n_agg, Matrix &Ac, cudaStream_t s);
void extract_diagonal(const Matrix &A, DevVec &d, cudaStream_t s);   // d[i] = A(i,i) (mat precision); bs>1: diagonal blocks

}  // namespace amgxb
