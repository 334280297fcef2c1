// Throughput build of the timeline kernels: FMA contraction, shared reciprocals, short-range
// exp, MUFU normals. Same formulas as the strict build; agreement is tested to 1e-9 relative.
#define MCR_FAST 1
#include "mcr_kernels.cuh"
namespace mcr {
const Launchers& fast_launchers() {
  static const Launchers L = {launch_timeline, launch_search, launch_draw, launch_helper, launch_sweep};
  return L;
}
}  // namespace mcr
