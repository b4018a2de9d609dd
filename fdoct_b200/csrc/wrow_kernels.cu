// Plans of the warp-per-A-scan kernel (wrow_kernel.cuh).
#include "plan_registry.cuh"

namespace abcoct {
// Transform lengths with N / 2 = 32 R, R even and <= 32: one warp holds the whole N/2-point complex transform.
// Warps per CTA: as many as the shared memory takes (per warp: exchange / staging buffer + pixel row + mbarrier), plus a
// 12-warp point with 168 registers per thread; abcoct_api.cpp picks (ABCOCT_WROW_NW overrides).
static const WPlanEntry kWPlans[] = {
    make_wentry<WPlan<2048, 15>>(), make_wentry<WPlan<2048, 14>>(), make_wentry<WPlan<2048, 12>>(),
    make_wentry<WPlan<1920, 15>>(), make_wentry<WPlan<1920, 12>>(),
    make_wentry<WPlan<1280, 16>>(), make_wentry<WPlan<1280, 12>>(),
    make_wentry<WPlan<1024, 16>>(), make_wentry<WPlan<1024, 12>>(),
};
// nw > 0: that many warps per CTA exactly; nw == 0: the plan with the most warps
const WPlanEntry* find_wplan(int N, int nw) {
  const WPlanEntry* best = nullptr;
  for (const WPlanEntry& e : kWPlans)
    if (e.N == N && (nw == 0 ? (best == nullptr || e.nw > best->nw) : e.nw == nw)) best = &e;
  return best;
}
}  // namespace abcoct
