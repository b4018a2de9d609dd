// Consumers of a finished B-scan that the reference runs right after the reconstruction block, on data that is still on
// the GPU:
//   * the linear `bscan` Mat (BscanFFT.cpp:1220-1222) = exp(bscandb * 2.303 / 20), with the two DC rows - which the dB image has
//     overwritten by row 4 (:1239-1240) - taken from the side buffer the fused kernel fills on request
//   * J0 lock-in display  (BscanFFT.cpp:1225-1231, 1257-1267): positivediff = max(bscan - jscansave, 0) + 0.001 (linear scale),
//     ln(.) * 20 / 2.303, max(., bscanthreshold), global min-max normalise, convertTo(CV_8UC1, 255)
//   * applyColorMap(., COLORMAP_JET) (BscanFFT.cpp:1268, 1284): u8 -> BGR through OpenCV's own 256-entry table
// All of it is elementwise work plus one min/max per B-scan: HBM-bound streaming kernels, grid = B-scans x chunks.
#include <algorithm>
#include <cuda_runtime.h>

#include <cstdint>

#include "jet_lut.h"
#include "kernels.h"

namespace abcoct {

namespace {

__constant__ unsigned c_jet[256];  // B | G << 8 | R << 16

__device__ __forceinline__ int ford(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float unford(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// MUFU.EX2 / MUFU.LG2 directly: the arguments are far from the denormal / overflow ranges the library versions guard
// (dB / db_scale in [-17, 40]; positivediff >= 1e-3), and 2 ulp is far inside the tolerances
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// The linear image of one B-scan, either stored (lin) or derived on the fly from the dB image: 2^(dB / db_scale), rows 0 and 1
// from dc [oph][2] (their dB before the DC-row mask).
template <bool FROM_DB>
struct LinSrc {
  const float* src;
  const float* dc;
  int oph;
  float inv_db_scale;
  __device__ __forceinline__ float one(size_t i) const {
    if (!FROM_DB) return __ldcs(src + i);
    const float x = i < 2 * (size_t)oph ? dc[2 * (i % oph) + i / oph] : __ldcs(src + i);
    return ex2_approx(x * inv_db_scale);
  }
  __device__ __forceinline__ float4 four(size_t i) const {  // i % 4 == 0, src 16-byte aligned, i + 3 inside the image
    if (FROM_DB && i < 2 * (size_t)oph) return make_float4(one(i), one(i + 1), one(i + 2), one(i + 3));
    float4 v = __ldcs(reinterpret_cast<const float4*>(src + i));
    if (FROM_DB) {
      v.x = ex2_approx(v.x * inv_db_scale);
      v.y = ex2_approx(v.y * inv_db_scale);
      v.z = ex2_approx(v.z * inv_db_scale);
      v.w = ex2_approx(v.w * inv_db_scale);
    }
    return v;
  }
};
struct LinBatch {  // batch view; B-scan b = blockIdx.y
  const float* src;
  const float* dc01;
  int oph;
  float inv_db_scale;
  size_t px;
  template <bool FROM_DB>
  __device__ __forceinline__ LinSrc<FROM_DB> at(int b) const {
    return LinSrc<FROM_DB>{src + (size_t)b * px, FROM_DB ? dc01 + (size_t)b * oph * 2 : nullptr, oph, inv_db_scale};
  }
};

__device__ __forceinline__ float jsub_db(float lin, float j, float db_scale) {
  const float v = fmaxf(lin - j, 0.f) + 1e-3f;  // makeonlypositive (BscanFFT.cpp:173-178) + 0.001 (:1230)
  return __log2f(v) * db_scale;                  // log (:1260), 20 * . / 2.303 (:1261)
}

// linear image from the dB image (VEC: 4 pixels per thread and iteration, needs px % 4 == 0)
template <bool VEC>
__global__ void __launch_bounds__(256) lin_from_db_kernel(const LinBatch lb, float* __restrict__ lin) {
  const LinSrc<true> s = lb.at<true>(blockIdx.y);
  float* dst = lin + (size_t)blockIdx.y * lb.px;
  const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (VEC) {
    for (size_t i = 4 * t0; i < lb.px; i += 4 * stride) __stcs(reinterpret_cast<float4*>(dst + i), s.four(i));
  } else {
    for (size_t i = t0; i < lb.px; i += stride) dst[i] = s.one(i);
  }
}

__global__ void jsub_reset_kernel(int* mm, int nB) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nB) {
    mm[2 * i] = ford(__int_as_float(0x7f800000));
    mm[2 * i + 1] = ford(__int_as_float(0xff800000));
  }
}

// pass 1: min / max of the subtracted dB image of every B-scan (max(., thr) is monotone: applied to the two scalars later)
template <bool FROM_DB, bool VEC>
__global__ void __launch_bounds__(256) jsub_minmax_kernel(const LinBatch lb, const float* __restrict__ jscan, int* mm, float db_scale) {
  const int b = blockIdx.y;
  const LinSrc<FROM_DB> s = lb.at<FROM_DB>(b);
  float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
  const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (VEC) {
    for (size_t i = 4 * t0; i < lb.px; i += 4 * stride) {
      const float4 l = s.four(i), j = __ldg(reinterpret_cast<const float4*>(jscan + i));
      const float d0 = jsub_db(l.x, j.x, db_scale), d1 = jsub_db(l.y, j.y, db_scale), d2 = jsub_db(l.z, j.z, db_scale),
                  d3 = jsub_db(l.w, j.w, db_scale);
      mn = fminf(fminf(mn, fminf(d0, d1)), fminf(d2, d3));
      mx = fmaxf(fmaxf(mx, fmaxf(d0, d1)), fmaxf(d2, d3));
    }
  } else {
    for (size_t i = t0; i < lb.px; i += stride) {
      const float db = jsub_db(s.one(i), __ldg(jscan + i), db_scale);
      mn = fminf(mn, db);
      mx = fmaxf(mx, db);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __shared__ float smn[8], smx[8];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    smn[w] = mn;
    smx[w] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
      mn = fminf(mn, smn[k]);
      mx = fmaxf(mx, smx[k]);
    }
    atomicMin(mm + 2 * b, ford(mn));
    atomicMax(mm + 2 * b + 1, ford(mx));
  }
}

// pass 2: threshold, normalise, quantise (round-half-even via the 1.5 * 2^23 trick, as the fused kernel does)
template <bool FROM_DB, bool VEC>
__global__ void __launch_bounds__(256) jsub_quant_kernel(const LinBatch lb, const float* __restrict__ jscan, const int* __restrict__ mm,
                                                         uint8_t* __restrict__ out, float db_scale, float thr) {
  const int b = blockIdx.y;
  const LinSrc<FROM_DB> s = lb.at<FROM_DB>(b);
  const float mn = fmaxf(unford(mm[2 * b]), thr), mx = fmaxf(unford(mm[2 * b + 1]), thr);
  const float range = mx - mn;
  const float sc = range > 2.220446049250313e-16f ? 255.0f / range : 0.f;  // cv::normalize: scale 0 for a flat image
  uint8_t* o = out + (size_t)b * lb.px;
  auto q = [&](float lin, float j) -> unsigned {
    return __float_as_uint(fmaf(fmaxf(jsub_db(lin, j, db_scale), thr) - mn, sc, 12582912.0f)) & 0xffu;
  };
  const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (VEC) {
    for (size_t i = 4 * t0; i < lb.px; i += 4 * stride) {
      const float4 l = s.four(i), j = __ldg(reinterpret_cast<const float4*>(jscan + i));
      __stcs(reinterpret_cast<unsigned*>(o + i), q(l.x, j.x) | (q(l.y, j.y) << 8) | (q(l.z, j.z) << 16) | (q(l.w, j.w) << 24));
    }
  } else {
    for (size_t i = t0; i < lb.px; i += stride) o[i] = (uint8_t)q(s.one(i), __ldg(jscan + i));
  }
}

// u8 -> BGR through a shared-memory copy of the table.  Vector path: 16 pixels per thread and iteration (one 16-byte load,
// three 16-byte stores); the tail and unaligned images go pixel by pixel.
__global__ void __launch_bounds__(256) jet_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, size_t n) {
  __shared__ unsigned lut[256];
  lut[threadIdx.x] = c_jet[threadIdx.x];  // blockDim.x == 256
  __syncthreads();
  const size_t n16 = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 ? n / 16 : 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = t0; i < n16; i += stride) {
    const uint4 p = __ldcs(reinterpret_cast<const uint4*>(in) + i);
    const unsigned w[4] = {p.x, p.y, p.z, p.w};
    unsigned o[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // 4 pixels -> 3 words
      const unsigned c0 = lut[w[k] & 0xffu], c1 = lut[(w[k] >> 8) & 0xffu], c2 = lut[(w[k] >> 16) & 0xffu], c3 = lut[w[k] >> 24];
      o[3 * k] = c0 | (c1 << 24);
      o[3 * k + 1] = (c1 >> 8) | (c2 << 16);
      o[3 * k + 2] = (c2 >> 16) | (c3 << 8);
    }
    uint4* dst = reinterpret_cast<uint4*>(out) + 3 * i;
    __stcs(dst, make_uint4(o[0], o[1], o[2], o[3]));
    __stcs(dst + 1, make_uint4(o[4], o[5], o[6], o[7]));
    __stcs(dst + 2, make_uint4(o[8], o[9], o[10], o[11]));
  }
  for (size_t i = 16 * n16 + t0; i < n; i += stride) {
    const unsigned c = lut[in[i]];
    out[3 * i] = (uint8_t)c;
    out[3 * i + 1] = (uint8_t)(c >> 8);
    out[3 * i + 2] = (uint8_t)(c >> 16);
  }
}

int blocks_for(size_t n, int per_block, int cap) {
  const size_t b = (n + per_block - 1) / per_block;
  if (cap < 1) cap = 1;
  return (int)(b < 1 ? 1 : (b > (size_t)cap ? cap : b));
}

}  // namespace

cudaError_t post_init_device() {
  unsigned packed[256];
  for (int v = 0; v < 256; ++v) packed[v] = kJetBGR[3 * v] | (kJetBGR[3 * v + 1] << 8) | ((unsigned)kJetBGR[3 * v + 2] << 16);
  return cudaMemcpyToSymbol(c_jet, packed, sizeof(packed));
}

namespace {
// CTAs per B-scan: enough to fill the machine across the batch, at most one per `per_block` pixels
int per_bscan(size_t px, int nb, int sm_count, int per_block) { return blocks_for(px, per_block, (16 * sm_count + nb - 1) / nb); }
bool vec_ok(size_t px, const void* a, const void* b, const void* c) {
  return px % 4 == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
}
}  // namespace

__global__ void __launch_bounds__(256) add_const_kernel(float* __restrict__ x, size_t n, float v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] += v;
}
cudaError_t launch_add_const(float* a, size_t na, float* b, size_t nb, float v, int sm_count, cudaStream_t st, int* launched) {
  int n = 0;
  if (a && na) {
    add_const_kernel<<<(unsigned)std::min<size_t>((na + 255) / 256, (size_t)sm_count * 8), 256, 0, st>>>(a, na, v);
    ++n;
  }
  if (b && nb) {
    add_const_kernel<<<(unsigned)std::min<size_t>((nb + 255) / 256, (size_t)sm_count * 8), 256, 0, st>>>(b, nb, v);
    ++n;
  }
  if (launched) *launched = n;
  return cudaGetLastError();
}

cudaError_t launch_lin_from_db(const float* db, const float* dc01, float* lin, int oph, size_t px, int nB, float inv_db_scale,
                               int sm_count, cudaStream_t st, int* launched) {
  int n = 0;
  const bool vec = vec_ok(px, db, lin, nullptr);
  for (int b0 = 0; b0 < nB; b0 += 65535) {  // gridDim.y limit
    const int nb = nB - b0 < 65535 ? nB - b0 : 65535;
    const LinBatch lb{db + (size_t)b0 * px, dc01 + (size_t)b0 * oph * 2, oph, inv_db_scale, px};
    const dim3 grid(per_bscan(px, nb, sm_count, 4096), nb);
    if (vec)
      lin_from_db_kernel<true><<<grid, 256, 0, st>>>(lb, lin + (size_t)b0 * px);
    else
      lin_from_db_kernel<false><<<grid, 256, 0, st>>>(lb, lin + (size_t)b0 * px);
    ++n;
  }
  if (launched) *launched = n;
  return cudaGetLastError();
}

template <bool FROM_DB, bool VEC>
static void jsub_pair(const LinBatch& lb, const float* jscan, int* mm, uint8_t* out, int nb, float db_scale, float thr, int sm_count,
                      cudaStream_t st) {
  const dim3 grid(per_bscan(lb.px, nb, sm_count, 4096), nb);
  jsub_minmax_kernel<FROM_DB, VEC><<<grid, 256, 0, st>>>(lb, jscan, mm, db_scale);
  jsub_quant_kernel<FROM_DB, VEC><<<grid, 256, 0, st>>>(lb, jscan, mm, out, db_scale, thr);
}

cudaError_t launch_jsub(const float* lin, const float* db, const float* dc01, const float* jscan, int* mm, uint8_t* out, int oph,
                        size_t px, int nB, float db_scale, float inv_db_scale, float thr, int sm_count, cudaStream_t st, int* launched) {
  const bool from_db = lin == nullptr;
  const float* src = from_db ? db : lin;
  const bool vec = vec_ok(px, src, jscan, out);
  jsub_reset_kernel<<<(nB + 127) / 128, 128, 0, st>>>(mm, nB);
  int n = 1;
  for (int b0 = 0; b0 < nB; b0 += 65535) {  // gridDim.y limit
    const int nb = nB - b0 < 65535 ? nB - b0 : 65535;
    const LinBatch lb{src + (size_t)b0 * px, from_db ? dc01 + (size_t)b0 * oph * 2 : nullptr, oph, inv_db_scale, px};
    uint8_t* o = out + (size_t)b0 * px;
    if (from_db) {
      if (vec) jsub_pair<true, true>(lb, jscan, mm + 2 * b0, o, nb, db_scale, thr, sm_count, st);
      else jsub_pair<true, false>(lb, jscan, mm + 2 * b0, o, nb, db_scale, thr, sm_count, st);
    } else {
      if (vec) jsub_pair<false, true>(lb, jscan, mm + 2 * b0, o, nb, db_scale, thr, sm_count, st);
      else jsub_pair<false, false>(lb, jscan, mm + 2 * b0, o, nb, db_scale, thr, sm_count, st);
    }
    n += 2;
  }
  if (launched) *launched = n;
  return cudaGetLastError();
}

cudaError_t launch_jet(const uint8_t* in, uint8_t* out, size_t n, int sm_count, cudaStream_t st) {
  jet_kernel<<<blocks_for(n, 16 * 256, 16 * sm_count), 256, 0, st>>>(in, out, n);
  return cudaGetLastError();
}

}  // namespace abcoct
