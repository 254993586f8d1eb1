// Instantiations of the lean streaming lhs kernel for MODE = LHS_TERM (see lhs_fast.cuh).
#include "lhs_fast.cuh"

namespace ur {
namespace fast {

FastKernel fast_lookup_term(int kind, int kp, int r, int e, int rpt) UR_FAST_LOOKUP_BODY(LHS_TERM)

}  // namespace fast
}  // namespace ur
