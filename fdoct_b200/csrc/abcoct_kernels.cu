// CUDA translation unit: plan instantiations of the fused kernel (and of its opt-in dual-pair variant), table builders,
// the scheduler-reset kernel and the launchers.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "kernels.h"

namespace abcoct {

// ------------------------------------------------------------------------------------------------ plan registry
template <class P>
static int smem_bytes_fn(int W, bool has_sub, int G) {
  return make_layout<P>(W, has_sub).total(G);
}
template <class P>
static int table_bytes_fn(int W) {
  return make_layout<P>(W, false).groups;
}

static void cossin_exact(long long num, long long den, int sgn, float2& out) {
  num %= den;
  const double ang = 2.0 * 3.14159265358979323846264338327950288 * double(num) / double(den);
  out.x = float(std::cos(ang));
  out.y = float(sgn * std::sin(ang));
}

template <class P>
static void build_blob_fn(int W, const int* idx, const float* wq, const float* win, std::vector<unsigned char>& blob) {
  const SmemLayout L = make_layout<P>(W, false);
  blob.assign(L.groups, 0);
  uint32_t* idxT = reinterpret_cast<uint32_t*>(blob.data() + L.idxT);
  float* wqT = reinterpret_cast<float*>(blob.data() + L.wqT);
  float* vwT = reinterpret_cast<float*>(blob.data() + L.vwT);
  float* winS = reinterpret_cast<float*>(blob.data() + L.win);
  float2* tw0 = reinterpret_cast<float2*>(blob.data() + L.tw0);
  float2* tw1 = reinterpret_cast<float2*>(blob.data() + L.tw1);
  for (int b = 0; b < P::N1; ++b) {
    for (int a = 0; a < P::R0P4; ++a) {
      const int q = P::N1 * a + b;
      const int i = a < P::R0 ? idx[q] : W;  // W = the zero sentinel slot (outside the swizzled range)
      const unsigned off1 = 8u * unsigned(i >= W ? W : stg_phys(i));
      const unsigned off0 = 8u * unsigned(i >= W ? W : stg_phys(i - 1));
      idxT[((a >> 2) * P::N1 + b) * 4 + (a & 3)] = off1 | (off0 << 16);
      // {weight, lerp(window)}: the window term carries the row-mean removal through the resampling (see phase_pre)
      const bool live = a < P::R0 && i < W;
      const double vw = live ? (double)win[i] + (double)wq[q] * ((double)win[i] - (double)win[i - 1]) : 0.0;
      wqT[((a >> 2) * P::N1 + b) * 4 + (a & 3)] = a < P::R0 ? wq[q] : 0.f;
      vwT[((a >> 2) * P::N1 + b) * 4 + (a & 3)] = (float)vw;
    }
    for (int c = 1; c < P::R0; ++c) cossin_exact((long long)b * c, P::N, kFftSign, tw0[(c - 1) * P::N1 + b]);
  }
  cal_swizzle_row(win, winS, W);
  if (P::THREE)
    for (int bp = 0; bp < P::N2; ++bp)
      for (int c1 = 1; c1 < P::R1; ++c1) cossin_exact((long long)bp * c1, P::N1, kFftSign, tw1[(c1 - 1) * P::N2 + bp]);
}

constexpr int kSmemBudget = 227 * 1024;

// Groups per CTA of a plan: limited by the thread budget and by shared memory at the largest row width (W = N).
template <class P, bool HAS_SUB>
struct PlanLimits {
  static constexpr SmemLayout L = make_layout<P>(P::N, HAS_SUB);
  static constexpr int by_threads = (P::MAXT / P::T) < 1 ? 1 : (P::MAXT / P::T);
  static constexpr int by_smem = (kSmemBudget - L.groups) / L.group_bytes;
  static constexpr int G = by_smem < 1 ? 1 : (by_smem < by_threads ? by_smem : by_threads);
};

template <class P>
static int groups_fn(bool has_sub) {
  return has_sub ? PlanLimits<P, true>::G : PlanLimits<P, false>::G;
}

template <class P, bool HAS_SUB, bool A1, bool IN_F32>
static cudaError_t launch_one(const ReconArgs& a, int grid, cudaStream_t st) {
  constexpr int G = PlanLimits<P, HAS_SUB>::G;
  const int smem = make_layout<P>(a.W, HAS_SUB).total(G);
  recon_kernel<P, G, HAS_SUB, A1, IN_F32><<<grid, P::T * G, smem, st>>>(a);
  return cudaGetLastError();
}
template <class P>
static cudaError_t launch_fn(const ReconArgs& a, bool has_sub, bool in_f32, int grid, cudaStream_t st) {
  const bool a1 = a.A == 1;
  if (in_f32) return a1 ? launch_one<P, false, true, true>(a, grid, st) : launch_one<P, false, false, true>(a, grid, st);
  if (has_sub) return a1 ? launch_one<P, true, true, false>(a, grid, st) : launch_one<P, true, false, false>(a, grid, st);
  return a1 ? launch_one<P, false, true, false>(a, grid, st) : launch_one<P, false, false, false>(a, grid, st);
}
template <class P, bool HAS_SUB, bool A1, bool IN_F32>
static cudaError_t attrs_one(int smem, int* regs) {
  const void* f = (const void*)recon_kernel<P, PlanLimits<P, HAS_SUB>::G, HAS_SUB, A1, IN_F32>;
  cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  cudaFuncAttributes fa;
  e = cudaFuncGetAttributes(&fa, f);
  if (e == cudaSuccess && regs) *regs = fa.numRegs;
  return e;
}
template <class P>
static cudaError_t attrs_fn(bool has_sub, bool a1, bool in_f32, int smem, int* regs) {
  if (in_f32) return a1 ? attrs_one<P, false, true, true>(smem, regs) : attrs_one<P, false, false, true>(smem, regs);
  if (has_sub) return a1 ? attrs_one<P, true, true, false>(smem, regs) : attrs_one<P, true, false, false>(smem, regs);
  return a1 ? attrs_one<P, false, true, false>(smem, regs) : attrs_one<P, false, false, false>(smem, regs);
}

// ------------------------------------------------------------------------------------------------ dual-pair variant
constexpr int kDualThreads = 256;  // 255 registers per thread: two row pairs live in every thread
template <class P, bool HAS_SUB>
struct PlanLimits2 {
  static constexpr SmemLayout2 L = make_layout2<P>(P::N, HAS_SUB);
  static constexpr int by_threads = (kDualThreads / P::T) < 1 ? 1 : (kDualThreads / P::T);
  static constexpr int by_smem = (kSmemBudget - L.groups) / L.group_bytes;
  static constexpr int G = by_smem < 1 ? 1 : (by_smem < by_threads ? by_smem : by_threads);
  static constexpr bool fits = by_smem >= 1 && 16 * P::N <= 65535;
};
template <class P>
static int groups2_fn(bool has_sub) {
  return has_sub ? PlanLimits2<P, true>::G : PlanLimits2<P, false>::G;
}
template <class P>
static int smem_bytes2_fn(int W, bool has_sub, int G) {
  return make_layout2<P>(W, has_sub).total(G);
}
template <class P>
static void build_blob2_fn(int W, const int* idx, const float* wq, const float* win, std::vector<unsigned char>& blob) {
  const SmemLayout2 L = make_layout2<P>(W, false);
  blob.assign(L.groups, 0);
  uint32_t* offT = reinterpret_cast<uint32_t*>(blob.data() + L.offT);
  float4* wvT = reinterpret_cast<float4*>(blob.data() + L.wvT);
  float* winS = reinterpret_cast<float*>(blob.data() + L.win);
  float4* tw0 = reinterpret_cast<float4*>(blob.data() + L.tw0);
  float4* tw1 = reinterpret_cast<float4*>(blob.data() + L.tw1);
  for (int b = 0; b < P::N1; ++b) {
    for (int a = 0; a < P::R0P4; ++a) {
      const int q = P::N1 * a + b;
      const int i = a < P::R0 ? idx[q] : W;  // W = the zero sentinel slot
      const unsigned off1 = 16u * unsigned(i >= W ? W : stg2_phys(i));
      const unsigned off0 = 16u * unsigned(i >= W ? W : stg2_phys(i - 1));
      offT[((a >> 2) * P::N1 + b) * 4 + (a & 3)] = off1 | (off0 << 16);
      const bool live = a < P::R0 && i < W;
      const float w = a < P::R0 ? wq[q] : 0.f;
      const float vw = live ? (float)((double)win[i] + (double)wq[q] * ((double)win[i] - (double)win[i - 1])) : 0.f;
      wvT[a * P::N1 + b] = make_float4(w, w, vw, vw);
    }
    for (int c = 1; c < P::R0; ++c) {
      float2 t;
      cossin_exact((long long)b * c, P::N, kFftSign, t);
      tw0[(c - 1) * P::N1 + b] = make_float4(t.x, t.x, t.y, t.y);
    }
  }
  cal_swizzle_row(win, winS, W);
  if (P::THREE)
    for (int bp = 0; bp < P::N2; ++bp)
      for (int c1 = 1; c1 < P::R1; ++c1) {
        float2 t;
        cossin_exact((long long)bp * c1, P::N1, kFftSign, t);
        tw1[(c1 - 1) * P::N2 + bp] = make_float4(t.x, t.x, t.y, t.y);
      }
}
template <class P, bool HAS_SUB, bool A1>
static cudaError_t launch2_one(const ReconArgs& a, int grid, cudaStream_t st) {
  constexpr int G = PlanLimits2<P, HAS_SUB>::G;
  const int smem = make_layout2<P>(a.W, HAS_SUB).total(G);
  recon2_kernel<P, G, HAS_SUB, A1><<<grid, P::T * G, smem, st>>>(a);
  return cudaGetLastError();
}
template <class P>
static cudaError_t launch2_fn(const ReconArgs& a, bool has_sub, int grid, cudaStream_t st) {
  const bool a1 = a.A == 1;
  if (has_sub) return a1 ? launch2_one<P, true, true>(a, grid, st) : launch2_one<P, true, false>(a, grid, st);
  return a1 ? launch2_one<P, false, true>(a, grid, st) : launch2_one<P, false, false>(a, grid, st);
}
template <class P, bool HAS_SUB, bool A1>
static cudaError_t attrs2_one(int smem, int* regs) {
  const void* f = (const void*)recon2_kernel<P, PlanLimits2<P, HAS_SUB>::G, HAS_SUB, A1>;
  cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  cudaFuncAttributes fa;
  e = cudaFuncGetAttributes(&fa, f);
  if (e == cudaSuccess && regs) *regs = fa.numRegs;
  return e;
}
template <class P>
static cudaError_t attrs2_fn(bool has_sub, bool a1, int smem, int* regs) {
  if (has_sub) return a1 ? attrs2_one<P, true, true>(smem, regs) : attrs2_one<P, true, false>(smem, regs);
  return a1 ? attrs2_one<P, false, true>(smem, regs) : attrs2_one<P, false, false>(smem, regs);
}

template <class P, bool DUAL>
static void add_dual(PlanEntry& e) {
  if constexpr (DUAL) {
    static_assert(PlanLimits2<P, false>::fits && PlanLimits2<P, true>::fits, "plan does not fit the dual-pair layout");
    e.groups2 = &groups2_fn<P>;
    e.smem_bytes2 = &smem_bytes2_fn<P>;
    e.build_blob2 = &build_blob2_fn<P>;
    e.launch2 = &launch2_fn<P>;
    e.attrs2 = &attrs2_fn<P>;
  } else {
    e.groups2 = nullptr;
    e.smem_bytes2 = nullptr;
    e.build_blob2 = nullptr;
    e.launch2 = nullptr;
    e.attrs2 = nullptr;
  }
}

template <class P, bool DUAL = false>
static PlanEntry make_entry() {
  PlanEntry e;
  e.d = PlanDesc{P::N, P::T, P::R0, P::R1, P::RL};
  e.groups = &groups_fn<P>;
  e.smem_bytes = &smem_bytes_fn<P>;
  e.table_bytes = &table_bytes_fn<P>;
  e.build_blob = &build_blob_fn<P>;
  e.launch = &launch_fn<P>;
  e.attrs = &attrs_fn<P>;
  add_dual<P, DUAL>(e);
  return e;
}

// The compiled transform lengths: powers of two for the sweep configs and 2^a*3^b*5^c lengths of the camera
// shapes / shipped .ini files (1280, 1920, 2560, 2880, 3840).  {N, T, R0, R1, RL}
using P128 = Plan<128, 32, 16, 1, 8>;  // tiny plan for the 128x96 reference fixtures
using P256 = Plan<256, 32, 16, 1, 16>;
using P512 = Plan<512, 32, 8, 8, 8>;
using P640 = Plan<640, 32, 10, 8, 8, 384>;
using P1024 = Plan<1024, 64, 16, 8, 8>;
using P1280 = Plan<1280, 64, 20, 8, 8, 384>;
using P1920 = Plan<1920, 128, 15, 16, 8>;
using P2048 = Plan<2048, 128, 16, 16, 8>;
using P2560 = Plan<2560, 128, 20, 16, 8, 384>;
using P2880 = Plan<2880, 96, 30, 12, 8>;
using P3840 = Plan<3840, 128, 30, 16, 8>;
using P4096 = Plan<4096, 128, 32, 16, 8>;

static const PlanEntry kPlans[] = {
    make_entry<P128>(),  make_entry<P256>(),  make_entry<P512>(),  make_entry<P640>(),
    make_entry<P1024, true>(), make_entry<P1280, true>(), make_entry<P1920, true>(), make_entry<P2048, true>(),
    make_entry<P2560>(), make_entry<P2880>(), make_entry<P3840>(), make_entry<P4096>(),
};

const PlanEntry* find_plan(int N) {
  for (const PlanEntry& e : kPlans)
    if (e.d.N == N) return &e;
  return nullptr;
}
int list_plans(int* out, int cap) {
  int n = 0;
  for (const PlanEntry& e : kPlans) {
    if (n < cap) out[n] = e.d.N;
    ++n;
  }
  return n;
}

// ------------------------------------------------------------------------------------------------ small kernels
// Reset the scheduler block of one launch: item ticket, per-B-scan min/max (order-preserving encodings of +inf / -inf)
// and the completion / hand-out counters.
__global__ void sched_init_kernel(int* sched, int nB) {
  const SchedView v = sched_view(sched, nB);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 32) sched[i] = 0;
  if (i < nB) {
    v.minv[i] = float_to_ordered(__int_as_float(0x7f800000));  // +inf
    v.maxv[i] = float_to_ordered(__int_as_float(0xff800000));  // -inf
    v.cnt[i] = 0;
  }
}
cudaError_t launch_sched_init(int* sched, int nB, cudaStream_t st) {
  const int n = nB > 32 ? nB : 32;
  sched_init_kernel<<<(n + 127) / 128, 128, 0, st>>>(sched, nB);
  return cudaGetLastError();
}

}  // namespace abcoct
