// CUDA translation unit: the plan registry front end (the plans themselves are instantiated in plans_small.cu, plans_large.cu
// and wrow_kernels.cu so that they compile in parallel), the scheduler-reset kernel and its launcher.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "kernels.h"

namespace abcoct {

const PlanEntry* plans_small(int* n);
const PlanEntry* plans_large(int* n);

const PlanEntry* find_plan(int N) {
  for (int part = 0; part < 2; ++part) {
    int n = 0;
    const PlanEntry* e = part ? plans_large(&n) : plans_small(&n);
    for (int i = 0; i < n; ++i)
      if (e[i].d.N == N) return &e[i];
  }
  return nullptr;
}
int list_plans(int* out, int cap) {
  int k = 0;
  for (int part = 0; part < 2; ++part) {
    int n = 0;
    const PlanEntry* e = part ? plans_large(&n) : plans_small(&n);
    for (int i = 0; i < n; ++i) {
      if (k < cap) out[k] = e[i].d.N;
      ++k;
    }
  }
  return k;
}

// ------------------------------------------------------------------------------------------------ small kernels
// Reset the scheduler block of one launch: item ticket, per-B-scan min/max (order-preserving encodings of +inf / -inf)
// and the completion / hand-out counters.
__global__ void sched_init_kernel(int* sched, int nB) {
  const SchedView v = sched_view(sched, nB);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kSchedHeader) sched[i] = 0;
  if (i < nB) {
    v.minv[i] = float_to_ordered(__int_as_float(0x7f800000));  // +inf
    v.maxv[i] = float_to_ordered(__int_as_float(0xff800000));  // -inf
    v.cnt[i] = 0;
    sched[kSchedHeader + 3 * nB + 32 * i] = 0;  // the resident-row kernel's row count (one line per B-scan)
  }
}
cudaError_t launch_sched_init(int* sched, int nB, cudaStream_t st) {
  const int n = nB > kSchedHeader ? nB : kSchedHeader;
  sched_init_kernel<<<(n + 127) / 128, 128, 0, st>>>(sched, nB);
  return cudaGetLastError();
}

}  // namespace abcoct
