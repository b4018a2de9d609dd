// Host-visible interface of the CUDA translation units (plan registry + launchers).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "plan.h"
#include "recon_kernel.cuh"

namespace abcoct {

// One compiled FFT plan: how to build its table blob and how to launch it.
struct PlanEntry {
  PlanDesc d;
  int (*groups)(bool has_sub);  // thread groups (packed A-scan pairs in flight) per CTA, fixed at compile time
  // bytes of dynamic shared memory for G groups
  int (*smem_bytes)(int W, bool has_sub, int G);
  int (*table_bytes)(int W);
  // pack idx (N entries, already sentinel-remapped, values in [1, M]), weights (N), window (W) and the
  // inter-pass twiddles into the blob the kernel copies to shared memory
  void (*build_blob)(int W, const int* idx, const float* wq, const float* win, std::vector<unsigned char>& blob);
  cudaError_t (*launch)(const ReconArgs& a, bool has_sub, int grid, cudaStream_t st);  // picks the averages == 1 variant itself
  cudaError_t (*attrs)(bool has_sub, bool a1, int smem, int* regs);  // opt in to large smem, report registers/thread
};

const PlanEntry* find_plan(int N);
int list_plans(int* out, int cap);

cudaError_t launch_sched_init(int* sched, int nB, cudaStream_t st);
}  // namespace abcoct
