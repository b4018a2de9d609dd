// Plans of the warp-per-A-scan kernel (wrow_kernel.cuh).
#include "plan_registry.cuh"

namespace abcoct {
// Transform lengths with N / 2 = 32 R, R even and <= 32: one warp holds the whole N/2-point complex transform.
// Two occupancy points per length (12 warps x 168 registers, 16 warps x 128 registers); abcoct_api.cpp picks.
static const WPlanEntry kWPlans[] = {
    make_wentry<WPlan<2048, 12>>(), make_wentry<WPlan<2048, 16>>(), make_wentry<WPlan<1920, 12>>(), make_wentry<WPlan<1920, 16>>(),
    make_wentry<WPlan<1280, 12>>(), make_wentry<WPlan<1280, 16>>(), make_wentry<WPlan<1024, 12>>(), make_wentry<WPlan<1024, 16>>(),
};
const WPlanEntry* find_wplan(int N, int nw) {
  for (const WPlanEntry& e : kWPlans)
    if (e.N == N && e.nw == nw) return &e;
  return nullptr;
}
}  // namespace abcoct
