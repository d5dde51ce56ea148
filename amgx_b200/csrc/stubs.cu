// stubs.cu -- entry points whose real implementation lands in a later milestone.  Each fails loudly.
// (currently none: DENSE_LU_SOLVER lives in dense_lu.cu, classical AMG in classical.cu)
#include "solvers.h"
namespace amgxb {
}
