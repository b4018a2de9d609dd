// Templates that turn a compile-time plan into a registry entry (table builder, launcher, attribute setter).  Included by the
// plan translation units (plans_small.cu, plans_large.cu, wrow_kernels.cu) and by the host-side native tests, which only need the
// table builders and therefore do not instantiate any kernel.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>

#include "kernels.h"

namespace abcoct {

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

template <class P>
static PlanEntry make_entry() {
  PlanEntry e;
  e.d = PlanDesc{P::N, P::T, P::R0, P::R1, P::RL};
  e.groups = &groups_fn<P>;
  e.smem_bytes = &smem_bytes_fn<P>;
  e.table_bytes = &table_bytes_fn<P>;
  e.build_blob = &build_blob_fn<P>;
  e.launch = &launch_fn<P>;
  e.attrs = &attrs_fn<P>;
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

// ------------------------------------------------------------------------------------------------ warp-per-A-scan plans
template <class WP, bool HAS_SUB, bool A1, bool FULLD>
static cudaError_t wlaunch_one(const ReconArgs& a, int grid, cudaStream_t st) {
  wrow_kernel<WP, HAS_SUB, A1, FULLD><<<grid, WP::NW * 32, WP::SMEM_BYTES, st>>>(a);
  return cudaGetLastError();
}
template <class WP, bool HAS_SUB, bool A1, bool FULLD>
static cudaError_t wattrs_one(int* regs) {
  const void* f = (const void*)wrow_kernel<WP, HAS_SUB, A1, FULLD>;
  cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, WP::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  cudaFuncAttributes fa;
  e = cudaFuncGetAttributes(&fa, f);
  if (e == cudaSuccess && regs) *regs = fa.numRegs;
  return e;
}
// dispatch over the three compile-time switches
template <class WP, template <class, bool, bool, bool> class F, class... Args>
static cudaError_t wdispatch(bool has_sub, bool a1, bool fulld, Args... args) {
  const int key = (has_sub ? 4 : 0) | (a1 ? 2 : 0) | (fulld ? 1 : 0);
  switch (key) {
    case 0: return F<WP, false, false, false>::run(args...);
    case 1: return F<WP, false, false, true>::run(args...);
    case 2: return F<WP, false, true, false>::run(args...);
    case 3: return F<WP, false, true, true>::run(args...);
    case 4: return F<WP, true, false, false>::run(args...);
    case 5: return F<WP, true, false, true>::run(args...);
    case 6: return F<WP, true, true, false>::run(args...);
    default: return F<WP, true, true, true>::run(args...);
  }
}
template <class WP, bool S, bool A, bool D>
struct WLaunchF {
  static cudaError_t run(const ReconArgs& a, int grid, cudaStream_t st) { return wlaunch_one<WP, S, A, D>(a, grid, st); }
};
template <class WP, bool S, bool A, bool D>
struct WAttrsF {
  static cudaError_t run(int* regs) { return wattrs_one<WP, S, A, D>(regs); }
};
template <class WP>
static cudaError_t wlaunch_fn(const ReconArgs& a, bool has_sub, int grid, cudaStream_t st) {
  return wdispatch<WP, WLaunchF, const ReconArgs&, int, cudaStream_t>(has_sub, a.A == 1, a.D == WP::N2, a, grid, st);
}
template <class WP>
static cudaError_t wattrs_fn(bool has_sub, bool a1, bool fulld, int* regs) {
  return wdispatch<WP, WAttrsF, int*>(has_sub, a1, fulld, regs);
}
template <class WP>
static void wblob_fn(const WrowTablesHost& t, std::vector<unsigned char>& blob) {
  blob.assign(WP::TABLE_BYTES, 0);
  wrow_build_blob<WP>(t, blob.data());
}
template <class WP>
static WPlanEntry make_wentry() {
  WPlanEntry e;
  e.N = WP::N;
  e.R = WP::R;
  e.nw = WP::NW;
  e.lm = WP::LM;
  e.wmax = WP::WMAX;
  e.smem_bytes = WP::SMEM_BYTES;
  e.build_blob = &wblob_fn<WP>;
  e.permute_cal_row = &wrow_permute_cal_row<WP>;
  e.launch = &wlaunch_fn<WP>;
  e.attrs = &wattrs_fn<WP>;
  return e;
}

// ------------------------------------------------------------------------------------------------ resident-row plans (wres_kernel.cuh)
template <class RP, bool HAS_SUB, bool A1, bool FULLD>
static cudaError_t rlaunch_one(const ReconArgs& a, int grid, cudaStream_t st) {
  wres_kernel<RP, HAS_SUB, A1, FULLD><<<grid, RP::NW * 32, RP::SMEM_BYTES, st>>>(a);
  return cudaGetLastError();
}
template <class RP, bool HAS_SUB, bool A1, bool FULLD>
static cudaError_t rattrs_one(int* regs) {
  const void* f = (const void*)wres_kernel<RP, HAS_SUB, A1, FULLD>;
  cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, RP::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  cudaFuncAttributes fa;
  e = cudaFuncGetAttributes(&fa, f);
  if (e == cudaSuccess && regs) *regs = fa.numRegs;
  return e;
}
template <class RP, bool S, bool A, bool D>
struct RLaunchF {
  static cudaError_t run(const ReconArgs& a, int grid, cudaStream_t st) { return rlaunch_one<RP, S, A, D>(a, grid, st); }
};
template <class RP, bool S, bool A, bool D>
struct RAttrsF {
  static cudaError_t run(int* regs) { return rattrs_one<RP, S, A, D>(regs); }
};
template <class RP>
static cudaError_t rlaunch_fn(const ReconArgs& a, bool has_sub, int grid, cudaStream_t st) {
  return wdispatch<RP, RLaunchF, const ReconArgs&, int, cudaStream_t>(has_sub, a.A == 1, a.D == RP::N2, a, grid, st);
}
template <class RP>
static cudaError_t rattrs_fn(bool has_sub, bool a1, bool fulld, int* regs) {
  return wdispatch<RP, RAttrsF, int*>(has_sub, a1, fulld, regs);
}
template <class RP>
static WPlanEntry make_rentry() {
  WPlanEntry e;
  e.N = RP::N;
  e.R = RP::R;
  e.nw = RP::NW;
  e.lm = 0;
  e.wmax = RP::WMAX;
  e.smem_bytes = RP::SMEM_BYTES;
  e.resident = 1;
  e.slots = RP::K;
  e.teams = RP::NT;
  e.build_blob = &wblob_fn<RP>;
  e.permute_cal_row = &wrow_permute_cal_row<RP>;
  e.launch = &rlaunch_fn<RP>;
  e.attrs = &rattrs_fn<RP>;
  return e;
}

}  // namespace abcoct
