// stubs.cu -- entry points whose real implementation lands in a later milestone.  Each fails loudly.
#include "solvers.h"
#include "dist.h"
namespace amgxb {
std::unique_ptr<Solver> make_dense_lu_solver(Config &, const std::string &, std::shared_ptr<Resources>) { fatal(AMGX_RC_NOT_IMPLEMENTED, "DENSE_LU_SOLVER: set coarse_solver=NOSOLVER"); }
}
