// CUDA translation unit: plan instantiations, the display-normalisation kernel and the launchers.
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
  uint16_t* idxT = reinterpret_cast<uint16_t*>(blob.data() + L.idxT);
  float* wqT = reinterpret_cast<float*>(blob.data() + L.wqT);
  float* winS = reinterpret_cast<float*>(blob.data() + L.win);
  float2* tw0 = reinterpret_cast<float2*>(blob.data() + L.tw0);
  float2* tw1 = reinterpret_cast<float2*>(blob.data() + L.tw1);
  for (int b = 0; b < P::N1; ++b) {
    for (int a = 0; a < P::R0P8; ++a) {
      const int q = P::N1 * a + b;
      idxT[((a >> 3) * P::N1 + b) * 8 + (a & 7)] = uint16_t(a < P::R0 ? idx[q] : W);
    }
    for (int a = 0; a < P::R0P4; ++a) {
      const int q = P::N1 * a + b;
      wqT[((a >> 2) * P::N1 + b) * 4 + (a & 3)] = a < P::R0 ? wq[q] : 0.f;
    }
    for (int c = 1; c < P::R0; ++c) cossin_exact((long long)b * c, P::N, kFftSign, tw0[(c - 1) * P::N1 + b]);
  }
  for (int i = 0; i < W; ++i) winS[i] = win[i];
  if (P::THREE)
    for (int bp = 0; bp < P::N2; ++bp)
      for (int c1 = 1; c1 < P::R1; ++c1) cossin_exact((long long)bp * c1, P::N1, kFftSign, tw1[(c1 - 1) * P::N2 + bp]);
}

template <class P>
struct PlanLimits {
  static constexpr int GMAX = (384 / P::T) < 1 ? 1 : (384 / P::T);
};

template <class P>
static cudaError_t launch_fn(const ReconArgs& a, bool has_sub, int G, int grid, cudaStream_t st) {
  constexpr int GMAX = PlanLimits<P>::GMAX;
  if (G < 1 || G > GMAX) return cudaErrorInvalidValue;
  const int smem = make_layout<P>(a.W, has_sub).total(G);
  if (has_sub)
    recon_kernel<P, GMAX, true, 1><<<grid, P::T * G, smem, st>>>(a, G);
  else
    recon_kernel<P, GMAX, false, 1><<<grid, P::T * G, smem, st>>>(a, G);
  return cudaGetLastError();
}
template <class P>
static cudaError_t attrs_fn(bool has_sub, int smem, int* regs) {
  constexpr int GMAX = PlanLimits<P>::GMAX;
  const void* f = has_sub ? (const void*)recon_kernel<P, GMAX, true, 1> : (const void*)recon_kernel<P, GMAX, false, 1>;
  cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  cudaFuncAttributes fa;
  e = cudaFuncGetAttributes(&fa, f);
  if (e == cudaSuccess && regs) *regs = fa.numRegs;
  return e;
}

template <class P>
static PlanEntry make_entry() {
  PlanEntry e;
  e.d = PlanDesc{P::N, P::T, P::R0, P::R1, P::RL};
  e.gmax = PlanLimits<P>::GMAX;
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
using P640 = Plan<640, 32, 10, 8, 8>;
using P1024 = Plan<1024, 64, 16, 8, 8>;
using P1280 = Plan<1280, 64, 20, 8, 8>;
using P1920 = Plan<1920, 128, 15, 16, 8>;
using P2048 = Plan<2048, 128, 16, 16, 8>;
using P2560 = Plan<2560, 128, 20, 16, 8>;
using P2880 = Plan<2880, 96, 30, 12, 8>;
using P3840 = Plan<3840, 128, 30, 16, 8>;
using P4096 = Plan<4096, 128, 32, 16, 8>;

static const PlanEntry kPlans[] = {
    make_entry<P128>(),  make_entry<P256>(),  make_entry<P512>(),  make_entry<P640>(),
    make_entry<P1024>(), make_entry<P1280>(), make_entry<P1920>(), make_entry<P2048>(),
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
__global__ void minmax_init_kernel(int* minmax, int nB) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nB) {
    minmax[2 * i] = float_to_ordered(__int_as_float(0x7f800000));      // +inf
    minmax[2 * i + 1] = float_to_ordered(__int_as_float(0xff800000));  // -inf
  }
}
cudaError_t launch_minmax_init(int* minmax, int nB, cudaStream_t st) {
  minmax_init_kernel<<<(nB + 127) / 128, 128, 0, st>>>(minmax, nB);
  return cudaGetLastError();
}

// threshold, global min-max normalise, transpose to depth-major, quantise (BscanFFT.cpp:1243-1255).
// tile = 128 A-scans x 32 depth bins; f64 arithmetic mirrors cv::normalize + convertTo(CV_8UC1, 255.0).
constexpr int NT_ROWS = 128, NT_BINS = 32;
__global__ void __launch_bounds__(256) normalise_kernel(const float* __restrict__ scratch, const int* __restrict__ minmax,
                                                        uint8_t* __restrict__ out8, float* __restrict__ outdb, int oph, int D,
                                                        float thr, int clamp55, float clamp_db) {
  __shared__ unsigned char t8[NT_BINS][NT_ROWS + 4];
  __shared__ float tdb[NT_BINS][NT_ROWS + 1];
  const int b = blockIdx.z, r0 = blockIdx.x * NT_ROWS, d0 = blockIdx.y * NT_BINS;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double smin = (double)ordered_to_float(minmax[2 * b]), smax = (double)ordered_to_float(minmax[2 * b + 1]);
  if (clamp55) {
    smin = fmin(smin, (double)clamp_db);
    smax = fmax(smax, (double)clamp_db);
  }
  const double scale = (smax - smin > 2.220446049250313e-16) ? 1.0 / (smax - smin) : 0.0;
  const double shift = 0.0 - smin * scale;
  const float* src = scratch + (size_t)b * oph * D;
  const bool want_db = outdb != nullptr;
  for (int rr = w; rr < NT_ROWS; rr += 8) {
    const int row = r0 + rr, d = d0 + lane;
    if (row < oph && d < D) {
      const float db = src[(size_t)row * D + d];
      double x = fmax((double)db, (double)thr);
      if (clamp55 && d == 5 && row == 5) x = (double)clamp_db;
      const double v = __dadd_rn(__dmul_rn(x, scale), shift);
      int q = __double2int_rn(v * 255.0);
      q = q < 0 ? 0 : (q > 255 ? 255 : q);
      t8[lane][rr] = (unsigned char)q;
      if (want_db) tdb[lane][rr] = db;
    }
  }
  __syncthreads();
  for (int dd = w; dd < NT_BINS; dd += 8) {
    const int d = d0 + dd;
    if (d >= D) continue;
    uint8_t* o8 = out8 + ((size_t)b * D + d) * oph + r0;
    const int rbase = 4 * lane;
    if (r0 + rbase + 3 < oph && ((((size_t)b * D + d) * oph + r0) & 3) == 0) {
      *reinterpret_cast<uint32_t*>(o8 + rbase) = *reinterpret_cast<const uint32_t*>(&t8[dd][rbase]);
    } else {
      for (int k = 0; k < 4; ++k)
        if (r0 + rbase + k < oph) o8[rbase + k] = t8[dd][rbase + k];
    }
    if (want_db) {
      float* od = outdb + ((size_t)b * D + d) * oph + r0;
      for (int k = 0; k < 4; ++k) {
        const int rr = lane + 32 * k;
        if (r0 + rr < oph) od[rr] = tdb[dd][rr];
      }
    }
  }
}
cudaError_t launch_normalise(const float* scratch, const int* minmax, uint8_t* out8, float* outdb, int nB, int oph, int D,
                             float thr, int clamp55, float clamp_db, cudaStream_t st) {
  dim3 grid((oph + NT_ROWS - 1) / NT_ROWS, (D + NT_BINS - 1) / NT_BINS, nB);
  normalise_kernel<<<grid, 256, 0, st>>>(scratch, minmax, out8, outdb, oph, D, thr, clamp55, clamp_db);
  return cudaGetLastError();
}

}  // namespace abcoct
