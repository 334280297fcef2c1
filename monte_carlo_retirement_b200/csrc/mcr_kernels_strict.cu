// Parity build of the timeline kernels: compiled with -fmad=false so every product and sum
// rounds separately, in the reference's operation order.
#define MCR_FAST 0
#include "mcr_kernels.cuh"
namespace mcr {
const Launchers& strict_launchers() {
  static const Launchers L = {launch_timeline, launch_search, launch_draw, launch_helper, launch_sweep};
  return L;
}
}  // namespace mcr
