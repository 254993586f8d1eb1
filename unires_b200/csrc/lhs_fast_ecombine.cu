// Instantiations of the lean streaming lhs kernel for MODE = LHS_ECOMBINE (see lhs_fast.cuh).
#include "lhs_fast.cuh"

namespace ur {
namespace fast {

FastKernel fast_lookup_ecombine(int kind, int kp, int r, int e, int rpt) UR_FAST_LOOKUP_BODY(LHS_ECOMBINE)

}  // namespace fast
}  // namespace ur
