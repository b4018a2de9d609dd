// Consumers of a finished B-scan that the reference runs right after the reconstruction block, on data that is still on
// the GPU:
//   * the linear `bscan` Mat (BscanFFT.cpp:1220-1222) = exp(bscandb * 2.303 / 20), with the two DC rows - which the dB image has
//     overwritten by row 4 (:1239-1240) - taken from the side buffer the fused kernel fills on request
//   * J0 lock-in display  (BscanFFT.cpp:1225-1231, 1257-1267): positivediff = max(bscan - jscansave, 0) + 0.001 (linear scale),
//     ln(.) * 20 / 2.303, max(., bscanthreshold), global min-max normalise, convertTo(CV_8UC1, 255)
//   * applyColorMap(., COLORMAP_JET) (BscanFFT.cpp:1268, 1284): u8 -> BGR through OpenCV's own 256-entry table
// All of it is elementwise work plus one min/max per B-scan: HBM-bound streaming kernels, grid = B-scans x chunks.
#include <cuda_runtime.h>

#include <cstdint>

#include "jet_lut.h"
#include "kernels.h"

namespace abcoct {

namespace {

__constant__ unsigned char c_jet[256 * 3];

__device__ __forceinline__ int ford(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float unford(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__device__ __forceinline__ float jsub_db(float lin, float j, float db_scale) {
  const float v = fmaxf(lin - j, 0.f) + 1e-3f;  // makeonlypositive (BscanFFT.cpp:173-178) + 0.001 (:1230)
  return log2f(v) * db_scale;                   // log (:1260), 20 * . / 2.303 (:1261)
}

// linear image from the dB image: lin = 2^(dB / db_scale); rows 0 and 1 from dc01 [nB][oph][2] (their dB before the mask)
__global__ void __launch_bounds__(256) lin_from_db_kernel(const float* __restrict__ db, const float* __restrict__ dc01,
                                                          float* __restrict__ lin, int oph, size_t px, float inv_db_scale) {
  const int b = blockIdx.y;
  const float* src = db + (size_t)b * px;
  float* dst = lin + (size_t)b * px;
  const float* dc = dc01 + (size_t)b * oph * 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < px; i += (size_t)gridDim.x * blockDim.x) {
    const float x = i < 2 * (size_t)oph ? dc[2 * (i % oph) + i / oph] : __ldcs(src + i);
    dst[i] = exp2f(x * inv_db_scale);
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
__global__ void __launch_bounds__(256) jsub_minmax_kernel(const float* __restrict__ lin, const float* __restrict__ jscan, int* mm,
                                                          size_t px, float db_scale) {
  const int b = blockIdx.y;
  const float* l = lin + (size_t)b * px;
  float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < px; i += (size_t)gridDim.x * blockDim.x) {
    const float db = jsub_db(__ldcs(l + i), __ldg(jscan + i), db_scale);
    mn = fminf(mn, db);
    mx = fmaxf(mx, db);
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
__global__ void __launch_bounds__(256) jsub_quant_kernel(const float* __restrict__ lin, const float* __restrict__ jscan,
                                                         const int* __restrict__ mm, uint8_t* __restrict__ out, size_t px,
                                                         float db_scale, float thr) {
  const int b = blockIdx.y;
  const float mn = fmaxf(unford(mm[2 * b]), thr), mx = fmaxf(unford(mm[2 * b + 1]), thr);
  const float range = mx - mn;
  const float sc = range > 2.220446049250313e-16f ? 255.0f / range : 0.f;  // cv::normalize: scale 0 for a flat image
  const float* l = lin + (size_t)b * px;
  uint8_t* o = out + (size_t)b * px;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < px; i += (size_t)gridDim.x * blockDim.x) {
    const float db = fmaxf(jsub_db(__ldcs(l + i), __ldg(jscan + i), db_scale), thr);
    o[i] = (uint8_t)(__float_as_uint(fmaf(db - mn, sc, 12582912.0f)) & 0xffu);
  }
}

// u8 -> BGR.  Each thread maps 4 pixels (one 32-bit load) to 12 bytes (three 32-bit stores) when the image allows it.
__global__ void __launch_bounds__(256) jet_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, size_t n) {
  const size_t n4 = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 3) == 0 ? n / 4 : 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = t0; i < n4; i += stride) {
    const unsigned p = __ldg(reinterpret_cast<const unsigned*>(in) + i);
    unsigned char c[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned v = (p >> (8 * k)) & 0xffu;
      c[3 * k] = c_jet[3 * v];
      c[3 * k + 1] = c_jet[3 * v + 1];
      c[3 * k + 2] = c_jet[3 * v + 2];
    }
    unsigned* o = reinterpret_cast<unsigned*>(out) + 3 * i;
#pragma unroll
    for (int k = 0; k < 3; ++k) o[k] = c[4 * k] | (c[4 * k + 1] << 8) | (c[4 * k + 2] << 16) | ((unsigned)c[4 * k + 3] << 24);
  }
  for (size_t i = 4 * n4 + t0; i < n; i += stride) {
    const unsigned v = in[i];
    out[3 * i] = c_jet[3 * v];
    out[3 * i + 1] = c_jet[3 * v + 1];
    out[3 * i + 2] = c_jet[3 * v + 2];
  }
}

int blocks_for(size_t n, int per_block, int cap) {
  const size_t b = (n + per_block - 1) / per_block;
  if (cap < 1) cap = 1;
  return (int)(b < 1 ? 1 : (b > (size_t)cap ? cap : b));
}

}  // namespace

cudaError_t post_init_device() { return cudaMemcpyToSymbol(c_jet, kJetBGR, sizeof(kJetBGR)); }

cudaError_t launch_lin_from_db(const float* db, const float* dc01, float* lin, int oph, size_t px, int nB, float inv_db_scale,
                               int sm_count, cudaStream_t st, int* launched) {
  int n = 0;
  for (int b0 = 0; b0 < nB; b0 += 65535) {  // gridDim.y limit
    const int nb = nB - b0 < 65535 ? nB - b0 : 65535;
    const int per = blocks_for(px, 2048, (8 * sm_count + nb - 1) / nb);
    lin_from_db_kernel<<<dim3(per, nb), 256, 0, st>>>(db + (size_t)b0 * px, dc01 + (size_t)b0 * oph * 2, lin + (size_t)b0 * px, oph, px,
                                                      inv_db_scale);
    ++n;
  }
  if (launched) *launched = n;
  return cudaGetLastError();
}

cudaError_t launch_jsub(const float* lin, const float* jscan, int* mm, uint8_t* out, size_t px, int nB, float db_scale, float thr,
                        int sm_count, cudaStream_t st, int* launched) {
  jsub_reset_kernel<<<(nB + 127) / 128, 128, 0, st>>>(mm, nB);
  int n = 1;
  for (int b0 = 0; b0 < nB; b0 += 65535) {  // gridDim.y limit
    const int nb = nB - b0 < 65535 ? nB - b0 : 65535;
    // enough CTAs to fill the machine across the batch, at most one per 2048 pixels of a B-scan
    const int per = blocks_for(px, 2048, (8 * sm_count + nb - 1) / nb);
    jsub_minmax_kernel<<<dim3(per, nb), 256, 0, st>>>(lin + (size_t)b0 * px, jscan, mm + 2 * b0, px, db_scale);
    jsub_quant_kernel<<<dim3(per, nb), 256, 0, st>>>(lin + (size_t)b0 * px, jscan, mm + 2 * b0, out + (size_t)b0 * px, px, db_scale, thr);
    n += 2;
  }
  if (launched) *launched = n;
  return cudaGetLastError();
}

cudaError_t launch_jet(const uint8_t* in, uint8_t* out, size_t n, int sm_count, cudaStream_t st) {
  jet_kernel<<<blocks_for(n, 4096, 16 * sm_count), 256, 0, st>>>(in, out, n);
  return cudaGetLastError();
}

}  // namespace abcoct
