// Plans of the group-per-row-pair kernel (recon_kernel.cuh), transform lengths up to 1280.
#include "plan_registry.cuh"

namespace abcoct {
static const PlanEntry kPlansSmall[] = {
    make_entry<P128>(), make_entry<P256>(), make_entry<P512>(), make_entry<P640>(), make_entry<P1024>(), make_entry<P1280>(),
};
const PlanEntry* plans_small(int* n) {
  *n = (int)(sizeof(kPlansSmall) / sizeof(kPlansSmall[0]));
  return kPlansSmall;
}
}  // namespace abcoct
