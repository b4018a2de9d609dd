// Plans of the resident-row kernel (wres_kernel.cuh).  Per transform length the FIRST entry is the default (abcoct_api.cpp;
// ABCOCT_WRES_NW selects another one for A/B measurements).
#include "plan_registry.cuh"

namespace abcoct {
static const WPlanEntry kRPlans[] = {
    make_rentry<RPlan<2048, 16>>(),
    make_rentry<RPlan<1920, 16>>(),
    make_rentry<RPlan<1280, 16>>(),
};
const WPlanEntry* find_rplan(int N, int nw) {
  for (const WPlanEntry& e : kRPlans)
    if (e.N == N && (nw == 0 || e.nw == nw)) return &e;
  return nullptr;
}
}  // namespace abcoct
