// Plans of the warp-per-A-scan kernel (wrow_kernel.cuh), transform lengths 1280 and 1024 (see wrow_kernels.cu).
#include "plan_registry.cuh"

namespace abcoct {
static const WPlanEntry kWPlansB[] = {
    make_wentry<WPlan<1280, 16, 0>>(), make_wentry<WPlan<1280, 12, 0>>(), make_wentry<WPlan<1280, 16, 1>>(),
    make_wentry<WPlan<1024, 16, 0>>(), make_wentry<WPlan<1024, 12, 0>>(),
};
const WPlanEntry* wplans_b(int* n) {
  *n = (int)(sizeof(kWPlansB) / sizeof(kWPlansB[0]));
  return kWPlansB;
}
}  // namespace abcoct
