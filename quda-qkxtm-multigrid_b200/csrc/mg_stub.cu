// temporary until coarse_op.cu lands
#include "dirac.h"
namespace qb {
void DiracTM::create_coarse_op(CoarseOperator &, const Transfer &) const { QB_ERROR("multigrid setup not built yet"); }
}
extern "C" {
void invertQuda(void *, void *, void *) { QB_ERROR("invertQuda not built yet"); }
void *newMultigridQuda(void *) { QB_ERROR("newMultigridQuda not built yet"); }
void destroyMultigridQuda(void *) {}
}
