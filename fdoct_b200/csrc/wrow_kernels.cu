// Plans of the warp-per-A-scan kernel (wrow_kernel.cuh), transform lengths 2048 and 1920; wrow_kernels_b.cu holds 1280 / 1024.
#include "plan_registry.cuh"

namespace abcoct {
// Transform lengths with N / 2 = 32 R, R even and <= 32: one warp holds the whole N/2-point complex transform.  Per length a
// few (warps per CTA, load mode) points; the FIRST entry of a length is its default (abcoct_api.cpp; ABCOCT_WROW_NW /
// ABCOCT_WROW_LM select another one for A/B measurements).
static const WPlanEntry kWPlansA[] = {
    make_wentry<WPlan<2048, 16, 0>>(), make_wentry<WPlan<2048, 12, 0>>(), make_wentry<WPlan<2048, 12, 1>>(),
    make_wentry<WPlan<1920, 16, 0>>(), make_wentry<WPlan<1920, 12, 0>>(),
};
const WPlanEntry* wplans_a(int* n) {
  *n = (int)(sizeof(kWPlansA) / sizeof(kWPlansA[0]));
  return kWPlansA;
}
const WPlanEntry* wplans_b(int* n);
const WPlanEntry* find_wplan(int N, int nw, int lm) {
  for (int part = 0; part < 2; ++part) {
    int n = 0;
    const WPlanEntry* e = part ? wplans_b(&n) : wplans_a(&n);
    for (int i = 0; i < n; ++i)
      if (e[i].N == N && (nw == 0 || e[i].nw == nw) && (lm < 0 || e[i].lm == lm)) return &e[i];
  }
  return nullptr;
}
}  // namespace abcoct
