// stubs.cu -- entry points whose real implementation lands in a later milestone.  Each fails loudly.
#include "solvers.h"
#include "dist.h"
namespace amgxb {
void block_norms(const DevVec &, int, int, int, const ReduceCtx &, ScalarBlock &, std::vector<double> &, cudaStream_t) { fatal(AMGX_RC_NOT_IMPLEMENTED, "block norms"); }
void block_jacobi_setup(const Matrix &, DevVec &, cudaStream_t) { fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "block Jacobi"); }
void block_jacobi_zero(const Matrix &, const DevVec &, const DevVec &, void *, double, cudaStream_t) { fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "block Jacobi"); }
void block_jacobi_sweep(const Matrix &, const DevVec &, const DevVec &, const void *, void *, double, cudaStream_t) { fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "block Jacobi"); }
void block_build_diag(Matrix &, cudaStream_t) { fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "block matrices"); }
void block_apply(const Matrix &, CsrEpi, const CsrOpArgs &, cudaStream_t) { fatal(AMGX_RC_NOT_SUPPORTED_BLOCKSIZE, "block SpMV"); }
std::unique_ptr<Solver> make_dilu_solver(Config &, const std::string &, std::shared_ptr<Resources>) { fatal(AMGX_RC_NOT_IMPLEMENTED, "MULTICOLOR_DILU"); }
std::unique_ptr<Solver> make_dense_lu_solver(Config &, const std::string &, std::shared_ptr<Resources>) { fatal(AMGX_RC_NOT_IMPLEMENTED, "DENSE_LU_SOLVER: set coarse_solver=NOSOLVER"); }
void classical_restrict(AMGLevel &, const DevVec &, cudaStream_t) { fatal(AMGX_RC_NOT_IMPLEMENTED, "classical AMG"); }
void classical_prolong_add(AMGLevel &, DevVec &, cudaStream_t) { fatal(AMGX_RC_NOT_IMPLEMENTED, "classical AMG"); }
void AMGSolver::setup_classical() { fatal(AMGX_RC_NOT_IMPLEMENTED, "classical AMG"); }
FGMRESSolver::FGMRESSolver(Config &cfg, const std::string &scope, std::shared_ptr<Resources> rsc) : Solver(cfg, scope, rsc) { fatal(AMGX_RC_NOT_IMPLEMENTED, "FGMRES"); }
void FGMRESSolver::solver_setup(bool) {}
void FGMRESSolver::solve_init(DevVec &, DevVec &, bool) {}
Status FGMRESSolver::solve_iteration(DevVec &, DevVec &, bool) { return ST_FAILED; }
}
