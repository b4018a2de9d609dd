// Plans of the group-per-row-pair kernel (recon_kernel.cuh), transform lengths from 1920 (camera shapes / shipped .ini files).
#include "plan_registry.cuh"

namespace abcoct {
static const PlanEntry kPlansLarge[] = {
    make_entry<P1920>(), make_entry<P2048>(), make_entry<P2560>(), make_entry<P2880>(), make_entry<P3840>(), make_entry<P4096>(),
};
const PlanEntry* plans_large(int* n) {
  *n = (int)(sizeof(kPlansLarge) / sizeof(kPlansLarge[0]));
  return kPlansLarge;
}
}  // namespace abcoct
