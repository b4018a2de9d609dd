// Plans of the warp-per-A-scan kernel (wrow_kernel.cuh), transform length 1280 (N = 1024 stays on the group-per-row-pair kernel: R = 16 would leave half the lanes idle in pass B, measured 3.3e8 against 3.6e8 A-scans/s) (see wrow_kernels.cu).
#include "plan_registry.cuh"

namespace abcoct {
static const WPlanEntry kWPlansB[] = {
    make_wentry<WPlan<1280, 16, 0>>(), make_wentry<WPlan<1280, 12, 0>>(), make_wentry<WPlan<1280, 16, 1>>(),
};
const WPlanEntry* wplans_b(int* n) {
  *n = (int)(sizeof(kWPlansB) / sizeof(kWPlansB[0]));
  return kWPlansB;
}
}  // namespace abcoct
