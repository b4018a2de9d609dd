// General pre-processing path ("slow path") of the ABC-OCT reconstruction block: every optional stage the reference
// can switch on in front of the lambda->k resampling, as plain CUDA kernels that turn raw camera frames into the
// apodised (and optionally Fourier-upsampled) f32 rows the fused reconstruction kernel then resamples and transforms.
//
//   median_kernel   medianBlur(mraw, m, mediann)                   BscanFFT.cpp:953-956   (k = 3, 5; replicated border)
//   bin_kernel      resize(m, opm, ..., INTER_AREA) integer bins   BscanFFT.cpp:958, BscanFFTspinjnt.cpp:1553
//   rowprep_kernel  convertTo / smoothmovavg / dark subtract / normalizerows / normalize / (y - yp) / yb /
//                   row-mean removal / Bartlett-Hann window / zeropadrowwise
//                   BscanFFT.cpp:987-991, 1125-1147, 88-97, 180-245, 247-304; BscanDark.cpp:1269, 218-236
//
// The default configuration (16-bit frames, no binning / median / smoothing / normalisation, multiplier 1) never comes
// here: it is handled entirely inside recon_kernel.  This path trades speed for coverage (it is still two orders of
// magnitude above any camera's frame rate) and keeps the reference's integer rounding rules bit-exact.
#include <algorithm>
#include <cuda_runtime.h>

#include <cstdint>

#include "fft_regs.cuh"
#include "kernels.h"

namespace abcoct {

// ------------------------------------------------------------------------------------------------ median (k = 3, 5)
template <class T, int K>
__global__ void median_kernel(const T* __restrict__ in, T* __restrict__ out, int w, int h, size_t in_row_stride /*elements*/,
                              size_t in_frame_stride, int nframes) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int f = blockIdx.z;
  if (x >= w || f >= nframes) return;
  constexpr int R = K / 2, NV = K * K;
  const T* src = in + (size_t)f * in_frame_stride;
  T v[NV];
#pragma unroll
  for (int dy = -R; dy <= R; ++dy) {
    const int yy = min(max(y + dy, 0), h - 1);  // cv::medianBlur replicates the border
#pragma unroll
    for (int dx = -R; dx <= R; ++dx) {
      const int xx = min(max(x + dx, 0), w - 1);
      v[(dy + R) * K + (dx + R)] = src[(size_t)yy * in_row_stride + xx];
    }
  }
  // partial selection sort up to the middle element
#pragma unroll
  for (int i = 0; i <= NV / 2; ++i) {
#pragma unroll
    for (int j = i + 1; j < NV; ++j) {
      const T a = v[i], b = v[j];
      v[i] = a < b ? a : b;
      v[j] = a < b ? b : a;
    }
  }
  out[((size_t)f * h + y) * w + x] = v[NV / 2];
}

// ------------------------------------------------------------------------------------------------ INTER_AREA integer binning
// cv::resize(..., INTER_AREA) with integer scale factors (resizeAreaFast_): 2 x 2 uses (sum + 2) >> 2; every other
// factor pair multiplies the sum by the f32 reciprocal of the area and rounds half to even (saturate_cast<T>(float)).
template <class T>
__global__ void bin_kernel(const T* __restrict__ in, T* __restrict__ out, int opw, int oph, int bx, int by, size_t in_row_stride,
                           size_t in_frame_stride, int nframes) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int f = blockIdx.z;
  if (x >= opw || f >= nframes) return;
  const T* src = in + (size_t)f * in_frame_stride + (size_t)(y * by) * in_row_stride + (size_t)x * bx;
  unsigned sum = 0;
  for (int dy = 0; dy < by; ++dy)
    for (int dx = 0; dx < bx; ++dx) sum += src[(size_t)dy * in_row_stride + dx];
  unsigned r;
  if (bx == 2 && by == 2) {
    r = (sum + 2u) >> 2;
  } else {
    const float scale = 1.f / (float)(bx * by);
    r = (unsigned)__float2int_rn(__fmul_rn((float)sum, scale));
    const unsigned mx = sizeof(T) == 1 ? 255u : 65535u;
    r = r > mx ? mx : r;
  }
  out[((size_t)f * oph + y) * opw + x] = (T)r;
}

// ------------------------------------------------------------------------------------------------ generic block FFT
// Stockham autosort transform of one row in shared memory by a whole CTA (any n = 2^a 3^b 5^c; the host picks the
// radix list greedily from {16, 15, 12, 10, 9, 8, 6, 5, 4, 3, 2}, e.g. 1920 = 16 15 8, 3840 = 16 16 15):
//   pass with radix r, current length nc, stride s:  for p < nc / r, q < s
//     y[q + s (r p + c)] = ( sum_j x[q + s (p + (nc / r) j)] w_r^(jc) ) * w_n^(p c s)
// (tools/fft_plan_model.py holds the NumPy model of this index algebra).  tw[k] = exp(SGN 2 pi i k / n), k < n.
// The transform buffers carry one float2 of padding per 16: the first pass (s == 1) writes y[R p + c] from consecutive threads
// p, a stride of R float2 that lands 16 lanes on one bank pair (ncu: 58 % of the shared-memory wavefronts of the unpadded kernel
// were conflicts); with the padding the stride becomes R + R / 16.  Reads are consecutive in the thread index either way.
__host__ __device__ __forceinline__ int fpad(int i) { return i + (i >> 4); }

struct RadixList {
  int n;
  int count;
  int r[12];
  int nc[12], s[12];  // per pass: remaining length n / (r[0] .. r[i-1]) and stride r[0] .. r[i-1]
};

template <int R, int SGN>
__device__ __forceinline__ void stockham_pass(const float2* __restrict__ x, float2* __restrict__ y, int n, int nc, int s,
                                              const float2* __restrict__ tw) {
  const int mq = nc / R;
  const float inv_s = 1.0f / (float)s;
  for (int b = threadIdx.x; b < n / R; b += blockDim.x) {
    // b / s without an integer division: (b + 0.5) / s sits at least 0.5 / s away from an integer, far more than the f32 error
    const int p = __float2int_rz(((float)b + 0.5f) * inv_s), q = b - p * s;
    float2 in[R], out[R];
#pragma unroll
    for (int j = 0; j < R; ++j) in[j] = x[fpad(q + s * (p + mq * j))];
    Dft<R, SGN, 1, 1>::run(in, out);
    y[fpad(q + s * (R * p))] = out[0];
    const int tb = p * s;  // p < nc / R and s * nc == n: tb * c < n for every c < R, no reduction needed
#pragma unroll
    for (int c = 1; c < R; ++c) y[fpad(q + s * (R * p + c))] = cmul(out[c], tw[tb * c]);
  }
}

// in-place semantics for the caller: returns the buffer that holds the result (a or b)
template <int SGN>
__device__ float2* block_fft(float2* a, float2* b, const RadixList& rl, const float2* __restrict__ tw) {
  float2 *x = a, *y = b;
  for (int i = 0; i < rl.count; ++i) {
    const int r = rl.r[i], nc = rl.nc[i], s = rl.s[i];  // current length and stride of pass i, filled in by the host
    switch (r) {
      case 2: stockham_pass<2, SGN>(x, y, rl.n, nc, s, tw); break;
      case 3: stockham_pass<3, SGN>(x, y, rl.n, nc, s, tw); break;
      case 4: stockham_pass<4, SGN>(x, y, rl.n, nc, s, tw); break;
      case 5: stockham_pass<5, SGN>(x, y, rl.n, nc, s, tw); break;
      case 6: stockham_pass<6, SGN>(x, y, rl.n, nc, s, tw); break;
      case 8: stockham_pass<8, SGN>(x, y, rl.n, nc, s, tw); break;
      case 9: stockham_pass<9, SGN>(x, y, rl.n, nc, s, tw); break;
      case 10: stockham_pass<10, SGN>(x, y, rl.n, nc, s, tw); break;
      case 12: stockham_pass<12, SGN>(x, y, rl.n, nc, s, tw); break;
      case 15: stockham_pass<15, SGN>(x, y, rl.n, nc, s, tw); break;
      default: stockham_pass<16, SGN>(x, y, rl.n, nc, s, tw); break;
    }
    __syncthreads();
    float2* t = x;
    x = y;
    y = t;
  }
  return x;
}

// ------------------------------------------------------------------------------------------------ row preparation
struct PrepArgs {
  const void* binned;   // integer pixels (after median + binning), u8 or u16: [nframes] frames of oph rows of opw pixels
  size_t row_stride, frame_stride;  // in pixels (dense after the binning kernel; the caller's strides when that was skipped)
  int bpp;              // 8 or 16
  int opw, oph, nframes;
  int movavgn;          // smoothmovavg half width (0 = off)
  float px_scale;       // data_y = pixel * px_scale (BscanFFTwebcam.cpp:1036: 1 / 765 for the channel sum; 1 otherwise - exact)
  const float* yd;      // nullable: dark frame (DARK variant)
  int rowwise;          // normalizerows(data_y, 0, 1)
  int global_norm;      // normalize(data_y, 0, 1, NORM_MINMAX) over the frame: 0 off, 1 = reduce pass, 2 = apply pass
  float* frame_minmax;  // [nframes][2] order-preserving ints (global_norm)
  const float* yb;      // background (f32)
  const float* yp;      // nullable: pi-shifted frame
  const float* win;     // [opw]
  int m, M;             // Fourier upsample factor and output row length m * opw
  int bandpass;         // BscanDark.cpp:218-236
  RadixList rlW, rlM;
  const float2* twW;    // exp(-2 pi i k / opw)
  const float2* twM;    // exp(+2 pi i k / M)
  float* out;           // [nframes][oph][M]
  const float* pre;     // nullable: rows already apodised by rowprep64_kernel ([nframes][oph][opw]); only the Fourier upsample is left
};

__device__ __forceinline__ float block_reduce(float v, float* red, int op /*0 sum, 1 min, 2 max*/) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = op == 0 ? v + u : (op == 1 ? fminf(v, u) : fmaxf(v, u));
  }
  __syncthreads();  // red may still be read from a previous reduction
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < nw; ++i) r = op == 0 ? r + red[i] : (op == 1 ? fminf(r, red[i]) : fmaxf(r, red[i]));
  return r;
}

__device__ __forceinline__ int f2ord(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }


// Stages of one row up to the apodised samples in shared memory (x, and a second buffer behind it when movavgn > 0).
// Returns true when this was the reduce-only pass of the global normalisation.
__device__ bool rowprep_one(const PrepArgs& a, int row, int f, float* x, float* red) {
  const int W = a.opw;
  const size_t pix = (size_t)f * a.frame_stride + (size_t)row * a.row_stride;
  // convertTo(data_y, CV_64F)  (BscanFFT.cpp:987); integers up to 65535 are exact in f32
  if (a.bpp == 8) {
    const uint8_t* src = static_cast<const uint8_t*>(a.binned) + pix;
    for (int j = threadIdx.x; j < W; j += blockDim.x) x[j] = (float)src[j] * a.px_scale;
  } else {
    const uint16_t* src = static_cast<const uint16_t*>(a.binned) + pix;
    for (int j = threadIdx.x; j < W; j += blockDim.x) x[j] = (float)src[j] * a.px_scale;
  }
  __syncthreads();
  if (a.movavgn > 0) {  // smoothmovavg, BscanFFT.cpp:247-304: 2n+1 taps, centre counted twice, missing taps -> centre
    const int n = a.movavgn;
    float* y = x + ((W + 3) & ~3);  // second row buffer (allocated when movavgn > 0, see rowprep_smem_bytes)
    for (int j = threadIdx.x; j < W; j += blockDim.x) {
      const float c = x[j];
      float s = c;
      for (int k = -n; k <= n; ++k) {
        const int jj = j + k;
        s += (jj > -1 && jj < W) ? x[jj] : c;
      }
      y[j] = s / 2.f / (float)(n + 1);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < W; j += blockDim.x) x[j] = y[j];
    __syncthreads();
  }
  if (a.yd) {  // data_y = data_y - data_yd, BscanDark.cpp:1269
    const float* yd = a.yd + (size_t)row * W;
    for (int j = threadIdx.x; j < W; j += blockDim.x) x[j] -= yd[j];
    __syncthreads();
  }
  if (a.rowwise) {  // normalizerows(data_y, 0, 1), BscanFFT.cpp:88-97: dst = src * s + (0 - min * s), s = 1 / (max - min)
    float mn = 3.4e38f, mx = -3.4e38f;
    for (int j = threadIdx.x; j < W; j += blockDim.x) {
      mn = fminf(mn, x[j]);
      mx = fmaxf(mx, x[j]);
    }
    mn = block_reduce(mn, red, 1);
    mx = block_reduce(mx, red, 2);
    const float s = (mx - mn) > 2.220446049250313e-16f ? 1.f / (mx - mn) : 0.f;
    const float sh = 0.f - mn * s;
    for (int j = threadIdx.x; j < W; j += blockDim.x) x[j] = fmaf(x[j], s, sh);
    __syncthreads();
  }
  if (a.global_norm) {  // normalize(data_y, data_y, 0, 1, NORM_MINMAX), BscanFFT.cpp:1128-1129 (whole frame)
    if (a.global_norm == 1) {
      float mn = 3.4e38f, mx = -3.4e38f;
      for (int j = threadIdx.x; j < W; j += blockDim.x) {
        mn = fminf(mn, x[j]);
        mx = fmaxf(mx, x[j]);
      }
      mn = block_reduce(mn, red, 1);
      mx = block_reduce(mx, red, 2);
      if (threadIdx.x == 0) {
        atomicMin(reinterpret_cast<int*>(a.frame_minmax) + 2 * f, f2ord(mn));
        atomicMax(reinterpret_cast<int*>(a.frame_minmax) + 2 * f + 1, f2ord(mx));
      }
      return true;  // reduce pass only
    }
    const float mn = ord2f(reinterpret_cast<const int*>(a.frame_minmax)[2 * f]);
    const float mx = ord2f(reinterpret_cast<const int*>(a.frame_minmax)[2 * f + 1]);
    const float s = (mx - mn) > 2.220446049250313e-16f ? 1.f / (mx - mn) : 0.f;
    const float sh = 0.f - mn * s;
    for (int j = threadIdx.x; j < W; j += blockDim.x) x[j] = fmaf(x[j], s, sh);
    __syncthreads();
  }
  // data_y = (data_y - data_yp) / data_yb, BscanFFT.cpp:1132
  const float* yb = a.yb + (size_t)row * W;
  const float* yp = a.yp ? a.yp + (size_t)row * W : nullptr;
  float sum = 0.f;
  for (int j = threadIdx.x; j < W; j += blockDim.x) {
    const float t = (x[j] - (yp ? yp[j] : 0.f)) / yb[j];
    x[j] = t;
    sum += t;
  }
  // per-row mean removal and apodisation, BscanFFT.cpp:1135-1143
  const float mean = block_reduce(sum, red, 0) / (float)W;
  for (int j = threadIdx.x; j < W; j += blockDim.x) x[j] = (x[j] - mean) * a.win[j];
  __syncthreads();
  return false;
}

// The common case of the general path (no smoothing, no normalisation - e.g. C3, which is here only for the Fourier upsample):
// both rows in ONE vectorised pass - 8 pixels per 16-byte load, the calibration rows as float4 - straight to the divided
// samples and the two row sums.  Same arithmetic, in the same order per sample, as rowprep_one.
template <class PX>
__device__ __forceinline__ void rowpair_fast(const PrepArgs& a, const PX* src0, const PX* src1, int row0, int row1, float* x0, float* x1,
                                             float& sum0, float& sum1) {
  const int W = a.opw;
  const float* rows[2] = {a.yb + (size_t)row0 * W, a.yb + (size_t)row1 * W};
  const float* yps[2] = {a.yp ? a.yp + (size_t)row0 * W : nullptr, a.yp ? a.yp + (size_t)row1 * W : nullptr};
  const float* yds[2] = {a.yd ? a.yd + (size_t)row0 * W : nullptr, a.yd ? a.yd + (size_t)row1 * W : nullptr};
  const PX* srcs[2] = {src0, src1};
  float* xs[2] = {x0, x1};
  float sums[2] = {0.f, 0.f};
  for (int ch = threadIdx.x; ch < W / 8; ch += blockDim.x) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float v[8];
      if (sizeof(PX) == 2) {
        const uint4 p = __ldg(reinterpret_cast<const uint4*>(srcs[r]) + ch);
        const unsigned w32[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (float)((k & 1) ? (w32[k >> 1] >> 16) : (w32[k >> 1] & 0xffffu)) * a.px_scale;
      } else {
        const uint2 p = __ldg(reinterpret_cast<const uint2*>(srcs[r]) + ch);
        const unsigned w32[2] = {p.x, p.y};
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (float)((w32[k >> 2] >> (8 * (k & 3))) & 0xffu) * a.px_scale;
      }
      float b[8], s[8], d[8];
      *reinterpret_cast<float4*>(b) = __ldg(reinterpret_cast<const float4*>(rows[r]) + 2 * ch);
      *reinterpret_cast<float4*>(b + 4) = __ldg(reinterpret_cast<const float4*>(rows[r]) + 2 * ch + 1);
      if (yps[r]) {
        *reinterpret_cast<float4*>(s) = __ldg(reinterpret_cast<const float4*>(yps[r]) + 2 * ch);
        *reinterpret_cast<float4*>(s + 4) = __ldg(reinterpret_cast<const float4*>(yps[r]) + 2 * ch + 1);
      }
      if (yds[r]) {
        *reinterpret_cast<float4*>(d) = __ldg(reinterpret_cast<const float4*>(yds[r]) + 2 * ch);
        *reinterpret_cast<float4*>(d + 4) = __ldg(reinterpret_cast<const float4*>(yds[r]) + 2 * ch + 1);
      }
      float t[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float x = v[k];
        if (yds[r]) x -= d[k];                      // BscanDark.cpp:1269
        t[k] = (x - (yps[r] ? s[k] : 0.f)) / b[k];  // BscanFFT.cpp:1132
        sums[r] += t[k];
      }
      *reinterpret_cast<float4*>(xs[r] + 8 * ch) = make_float4(t[0], t[1], t[2], t[3]);
      *reinterpret_cast<float4*>(xs[r] + 8 * ch + 4) = make_float4(t[4], t[5], t[6], t[7]);
    }
  }
  sum0 = sums[0];
  sum1 = sums[1];
}

__host__ __device__ inline int fft_buf_slots(int M) { return (fpad(M) + 2) & ~1; }  // float2 slots of one padded transform buffer, 16-B multiple

__host__ __device__ inline size_t rowbuf_bytes(int opw, int movavgn) {
  const size_t one = (size_t)((opw + 3) & ~3) * sizeof(float);
  return (one * (movavgn > 0 ? 2 : 1) + 15) & ~(size_t)15;
}

// One CTA per (row pair, frame): the two rows share the Fourier upsample as the real and imaginary part of ONE complex
// transform each way.  Dynamic smem: m == 1: 2 x row buffers; m > 1: float2 a[M] | float2 b[M], with the two row buffers
// aliased onto b (they are dead once the rows are packed into a) - 62 KB instead of 77 KB for C3, i.e. 3 CTAs per SM.
__global__ void __launch_bounds__(256) rowprep_kernel(const PrepArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float red[8];
  const int W = a.opw, f = blockIdx.y;
  const int r0 = 2 * blockIdx.x;
  const bool has1 = r0 + 1 < a.oph;
  const int r1 = has1 ? r0 + 1 : r0;
  const size_t rb = rowbuf_bytes(W, a.movavgn);
  float2* bufa = reinterpret_cast<float2*>(smem_raw);
  float2* bufb = bufa + fft_buf_slots(a.M);
  unsigned char* rows_at = a.m > 1 ? reinterpret_cast<unsigned char*>(bufb) : smem_raw;
  float* x0 = reinterpret_cast<float*>(rows_at);
  float* x1 = reinterpret_cast<float*>(rows_at + rb);
  const size_t px = (size_t)(a.bpp == 8 ? 1 : 2);
  const unsigned char* s0 = static_cast<const unsigned char*>(a.binned) + ((size_t)f * a.frame_stride + (size_t)r0 * a.row_stride) * px;
  const unsigned char* s1 = static_cast<const unsigned char*>(a.binned) + ((size_t)f * a.frame_stride + (size_t)r1 * a.row_stride) * px;
  const bool fast = a.movavgn == 0 && !a.rowwise && !a.global_norm && (W & 7) == 0 &&
                    ((reinterpret_cast<uintptr_t>(s0) | reinterpret_cast<uintptr_t>(s1)) & (8 * px - 1)) == 0;
  if (a.pre) {
    const float* p0 = a.pre + ((size_t)f * a.oph + r0) * W;
    const float* p1 = a.pre + ((size_t)f * a.oph + r1) * W;
    for (int j = threadIdx.x; j < W; j += blockDim.x) {
      x0[j] = p0[j];
      x1[j] = p1[j];
    }
    __syncthreads();
  } else if (fast) {  // uniform over the CTA
    float sum0, sum1;
    if (a.bpp == 8)
      rowpair_fast<uint8_t>(a, s0, s1, r0, r1, x0, x1, sum0, sum1);
    else
      rowpair_fast<uint16_t>(a, reinterpret_cast<const uint16_t*>(s0), reinterpret_cast<const uint16_t*>(s1), r0, r1, x0, x1, sum0, sum1);
    const float mean0 = block_reduce(sum0, red, 0) / (float)W;  // BscanFFT.cpp:1135-1143; the barrier inside also publishes x0 / x1
    const float mean1 = block_reduce(sum1, red, 0) / (float)W;
    for (int j = threadIdx.x; j < W; j += blockDim.x) {
      const float wj = a.win[j];
      x0[j] = (x0[j] - mean0) * wj;
      x1[j] = (x1[j] - mean1) * wj;
    }
    __syncthreads();
  } else {
    const bool reduce_only = rowprep_one(a, r0, f, x0, red);
    if (has1) rowprep_one(a, r1, f, x1, red);
    if (reduce_only) return;
  }
  float* out0 = a.out + ((size_t)f * a.oph + r0) * a.M;
  float* out1 = a.out + ((size_t)f * a.oph + r1) * a.M;
  if (a.m <= 1) {
    for (int j = threadIdx.x; j < W; j += blockDim.x) {
      out0[j] = x0[j];
      if (has1) out1[j] = x1[j];
    }
    return;
  }
  // zeropadrowwise, BscanFFT.cpp:180-245: forward DFT scaled by 1 / opw, zero-pad the centred spectrum to M, inverse DFT
  // with DFT_REAL_OUTPUT, which only reads bins 0 .. M/2 - so the -opw/2 (Nyquist) bin is dropped.  For two real rows
  // z = row0 + i row1 this is: Z = DFT(z) / opw, keep Z[k] (0 <= k < opw/2) at k and Z[opw - k] (1 <= k < opw/2) at M - k,
  // z' = inverse DFT of length M; row0' = Re z', row1' = Im z'.
  for (int j = threadIdx.x; j < W; j += blockDim.x) bufa[fpad(j)] = make_float2(x0[j], has1 ? x1[j] : 0.f);
  __syncthreads();
  float2* X = block_fft<-1>(bufa, bufb, a.rlW, a.twW);
  float2* Y = (X == bufa) ? bufb : bufa;
  const float sc = 1.f / (float)W;
  const int half = W / 2;
  const int lo = a.bandpass ? 3 : 0;  // BscanDark.cpp:218-236 keeps bins [3, floor(opw / 10)) of either sign
  const int hi = a.bandpass ? W / 10 : half;
  for (int k = threadIdx.x; k < a.M; k += blockDim.x) {
    float2 v = make_float2(0.f, 0.f);
    if (k < half) {
      if (k >= lo && k < hi) v = make_float2(X[fpad(k)].x * sc, X[fpad(k)].y * sc);
    } else if (k > a.M - half) {
      const int kk = a.M - k;
      if (kk >= lo && kk < hi) v = make_float2(X[fpad(W - kk)].x * sc, X[fpad(W - kk)].y * sc);
    }
    Y[fpad(k)] = v;
  }
  __syncthreads();
  float2* R = block_fft<+1>(Y, X, a.rlM, a.twM);
  for (int j = threadIdx.x; j < a.M; j += blockDim.x) {
    const float2 v = R[fpad(j)];
    out0[j] = v.x;
    if (has1) out1[j] = v.y;
  }
}

// ------------------------------------------------------------------------------------------------ row preparation in f64
// The normalised-calibration regime (rowwisenormalize / !donotnormalize): data_yb is stretched to [1e-4, 1], 1 / data_yb spans four
// decades, and the mean removal cancels several digits.  The reference does all of this in CV_64F and rounds to f32 once, at
// Mat_<float>(data_ylin) (BscanFFT.cpp:1181); an f32 pipeline adds three to four roundings in front of the transform and ends up at
// 1.9e-4 of the floor against the reference's 0.7e-4.  Here the same stages run in double, one CTA per row, same order as
// rowprep_one; the rows go out as doubles (m == 1; generic_recon_kernel interpolates them in f64) or, in front of the Fourier
// upsample, as floats exactly where zeropadrowwise converts (BscanFFT.cpp:209).
struct PrepArgs64 {
  const void* binned;
  size_t row_stride, frame_stride;
  int bpp, opw, oph, nframes, movavgn;
  double px_scale;
  const double *yd, *yb, *yp, *win;  // yd / yp nullable
  int rowwise, global_norm;          // global_norm: 0 off, 1 reduce pass, 2 apply pass
  long long* frame_minmax;           // [nframes][2] order-preserving integers of the doubles
  double* out64;                     // m == 1: [nframes][oph][opw]
  float* out32;                      // m > 1
};
__device__ __forceinline__ long long d2ord(double d) {
  const long long i = __double_as_longlong(d);
  return i >= 0 ? i : i ^ 0x7fffffffffffffffLL;
}
__device__ __forceinline__ double ord2d(long long i) { return __longlong_as_double(i >= 0 ? i : i ^ 0x7fffffffffffffffLL); }
__device__ __forceinline__ double block_reduce64(double v, double* red, int op /*0 sum, 1 min, 2 max*/) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double u = __shfl_xor_sync(0xffffffffu, v, o);
    v = op == 0 ? v + u : (op == 1 ? fmin(v, u) : fmax(v, u));
  }
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double r = red[0];
  for (int i = 1; i < nw; ++i) r = op == 0 ? r + red[i] : (op == 1 ? fmin(r, red[i]) : fmax(r, red[i]));
  return r;
}
__global__ void minmax64_reset_kernel(long long* mm, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    mm[2 * i] = d2ord(__longlong_as_double(0x7ff0000000000000LL));
    mm[2 * i + 1] = d2ord(__longlong_as_double((long long)0xfff0000000000000ULL));
  }
}
__global__ void __launch_bounds__(256) rowprep64_kernel(const PrepArgs64 a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[8];
  double* x = reinterpret_cast<double*>(smem_raw);
  const int W = a.opw, row = blockIdx.x, f = blockIdx.y;
  const size_t pix = (size_t)f * a.frame_stride + (size_t)row * a.row_stride;
  if (a.bpp == 8) {  // convertTo(data_y, CV_64F), BscanFFT.cpp:987 (x 1 / 765 for the webcam channel sum)
    const uint8_t* src = static_cast<const uint8_t*>(a.binned) + pix;
    for (int j = threadIdx.x; j < W; j += blockDim.x) x[j] = (double)src[j] * a.px_scale;
  } else {
    const uint16_t* src = static_cast<const uint16_t*>(a.binned) + pix;
    for (int j = threadIdx.x; j < W; j += blockDim.x) x[j] = (double)src[j] * a.px_scale;
  }
  __syncthreads();
  if (a.movavgn > 0) {  // smoothmovavg, BscanFFT.cpp:247-304 (same summation order: taps -n .. n, then the centre once more)
    const int n = a.movavgn;
    double* y = x + W;
    for (int j = threadIdx.x; j < W; j += blockDim.x) {
      const double c = x[j];
      double s = 0.0;
      for (int k = -n; k <= n; ++k) {
        const int jj = j + k;
        s = s + ((jj > -1 && jj < W) ? x[jj] : c);
      }
      s = s + c;
      y[j] = s / 2 / (n + 1);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < W; j += blockDim.x) x[j] = y[j];
    __syncthreads();
  }
  if (a.yd) {  // BscanDark.cpp:1269
    const double* yd = a.yd + (size_t)row * W;
    for (int j = threadIdx.x; j < W; j += blockDim.x) x[j] -= yd[j];
    __syncthreads();
  }
  if (a.rowwise) {  // normalizerows(data_y, 0, 1): cv::normalize NORM_MINMAX = convertTo(scale, shift)
    double mn = 1.7e308, mx = -1.7e308;
    for (int j = threadIdx.x; j < W; j += blockDim.x) {
      mn = fmin(mn, x[j]);
      mx = fmax(mx, x[j]);
    }
    mn = block_reduce64(mn, red, 1);
    mx = block_reduce64(mx, red, 2);
    const double s = (mx - mn) > 2.220446049250313e-16 ? 1.0 / (mx - mn) : 0.0;
    const double sh = 0.0 - mn * s;
    for (int j = threadIdx.x; j < W; j += blockDim.x) x[j] = __dadd_rn(__dmul_rn(x[j], s), sh);
    __syncthreads();
  }
  if (a.global_norm) {  // normalize(data_y, data_y, 0, 1, NORM_MINMAX) over the frame, BscanFFT.cpp:1128-1129
    if (a.global_norm == 1) {
      double mn = 1.7e308, mx = -1.7e308;
      for (int j = threadIdx.x; j < W; j += blockDim.x) {
        mn = fmin(mn, x[j]);
        mx = fmax(mx, x[j]);
      }
      mn = block_reduce64(mn, red, 1);
      mx = block_reduce64(mx, red, 2);
      if (threadIdx.x == 0) {
        atomicMin(a.frame_minmax + 2 * f, d2ord(mn));
        atomicMax(a.frame_minmax + 2 * f + 1, d2ord(mx));
      }
      return;
    }
    const double mn = ord2d(a.frame_minmax[2 * f]), mx = ord2d(a.frame_minmax[2 * f + 1]);
    const double s = (mx - mn) > 2.220446049250313e-16 ? 1.0 / (mx - mn) : 0.0;
    const double sh = 0.0 - mn * s;
    for (int j = threadIdx.x; j < W; j += blockDim.x) x[j] = __dadd_rn(__dmul_rn(x[j], s), sh);
    __syncthreads();
  }
  const double* yb = a.yb + (size_t)row * W;
  const double* yp = a.yp ? a.yp + (size_t)row * W : nullptr;
  double sum = 0.0;
  for (int j = threadIdx.x; j < W; j += blockDim.x) {  // (data_y - data_yp) / data_yb, BscanFFT.cpp:1132 (cv::divide: x / 0 = 0)
    const double b = yb[j];
    const double t = b != 0.0 ? (x[j] - (yp ? yp[j] : 0.0)) / b : 0.0;
    x[j] = t;
    sum += t;
  }
  const double mean = block_reduce64(sum, red, 0) / (double)W;  // BscanFFT.cpp:1135-1143
  for (int j = threadIdx.x; j < W; j += blockDim.x) {
    const double v = __dmul_rn(x[j] - mean, a.win[j]);
    if (a.out64)
      a.out64[((size_t)f * a.oph + row) * W + j] = v;
    else
      a.out32[((size_t)f * a.oph + row) * W + j] = (float)v;
  }
}

// ------------------------------------------------------------------------------------------------ any transform length
// cv::dft takes any N (BscanFFT.cpp:1185); the fused kernels are compiled for twelve lengths, rows that are multiples of 8 samples
// and D <= N / 2.  Everything else (N = 2^a 3^b 5^c, any row width, D up to N) runs here, on the rows prepared by rowprep_kernel:
// one CTA per pair of A-scans of one B-scan - gather-lerp (BscanFFT.cpp:1151-1177) of both rows into one complex buffer, the
// shared-memory Stockham transform above (run-time radices), two-for-one split, magnitude (:1189-1190), accumulation over the
// frames (:1193-1209), dB + DC-row mask (:1221-1240) into the dB scratch and the B-scan's thresholded min / max; a second kernel
// normalises and transposes (:1243-1255).  Slow (a few 1e7 A-scans/s) and exact to the same tolerances; never used for the
// BASELINE configurations.
struct GenericArgs {
  const float* rows;  // [nB * A][oph][M] prepared rows
  const double* rows64;  // the same as doubles (single-row regime without Fourier upsample); rows is ignored then
  const double* wq64;    // [N] lerp weights in f64 (single-row regime: the interpolation runs in f64 like the reference's)
  int M, N, D, Dp, oph, A, nB;
  const int* idx;     // [N] source sample (1 .. M - 1), M = never written (zero)
  const float* wq;    // [N] lerp weights (the reference's quirk already applied)
  RadixList rl;
  const float2* tw;   // exp(+2 pi i k / N)
  float* scratch;     // [nB][oph][Dp]
  int *minv, *maxv;   // [nB] order-preserving ints
  float* dc01;        // nullable [nB][oph][2]
  float out_scale, db_scale_ln, thr;
  int clamp55;
  int single;  // one A-scan per transform (imaginary part zero): every row has its own f32 noise floor, like cv::dft's rows
};
__global__ void __launch_bounds__(256) generic_recon_kernel(const GenericArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float red[8];
  const int N = a.N, b = blockIdx.y, r0 = a.single ? blockIdx.x : 2 * blockIdx.x;
  const bool has1 = !a.single && r0 + 1 < a.oph;
  float2* bufa = reinterpret_cast<float2*>(smem_raw);
  float2* bufb = bufa + fft_buf_slots(N);
  float* acc0 = reinterpret_cast<float*>(bufb + fft_buf_slots(N));
  float* acc1 = acc0 + a.D;
  for (int k = threadIdx.x; k < a.D; k += blockDim.x) acc0[k] = acc1[k] = 0.f;
  for (int f = 0; f < a.A; ++f) {
    const size_t rowoff = (((size_t)b * a.A + f) * a.oph + r0) * a.M;
    const float* y0 = a.rows + rowoff;
    const float* y1 = y0 + (has1 ? a.M : 0);
    __syncthreads();  // the buffers of the previous frame have been consumed
    if (a.single) {  // data_ylin in f64, rounded to f32 once (Mat_<float>(data_ylin), BscanFFT.cpp:1181)
      const double* z0 = a.rows64 ? a.rows64 + rowoff : nullptr;
      for (int q = threadIdx.x; q < N; q += blockDim.x) {
        const int i = a.idx[q];
        float2 v = make_float2(0.f, 0.f);
        if (i < a.M) {
          const double yi = z0 ? z0[i] : (double)y0[i], ym = z0 ? z0[i - 1] : (double)y0[i - 1];
          v.x = (float)__dadd_rn(yi, __dmul_rn(a.wq64[q], yi - ym));
        }
        bufa[fpad(q)] = v;
      }
    } else
    for (int q = threadIdx.x; q < N; q += blockDim.x) {
      const int i = a.idx[q];
      float2 v = make_float2(0.f, 0.f);
      if (i < a.M) {  // data_ylin = y[i] + w (y[i] - y[i-1]); columns 0 and N - 1 are never written (BscanFFT.cpp:1164-1171)
        const float w = a.wq[q];
        v.x = fmaf(w, y0[i] - y0[i - 1], y0[i]);
        v.y = has1 ? fmaf(w, y1[i] - y1[i - 1], y1[i]) : 0.f;
      }
      bufa[fpad(q)] = v;
    }
    __syncthreads();
    const float2* Z = block_fft<+1>(bufa, bufb, a.rl, a.tw);  // unscaled inverse DFT of row0 + i row1 (BscanFFT.cpp:1185)
    for (int k = threadIdx.x; k < a.D; k += blockDim.x) {
      const float2 z = Z[fpad(k)], zc = Z[fpad(k == 0 ? 0 : N - k)];
      // X0[k] = (Z[k] + conj Z[N-k]) / 2, X1[k] = (Z[k] - conj Z[N-k]) / (2 i); the 1 / 2 lives in out_scale
      const float ar = z.x + zc.x, ai = z.y - zc.y, br = z.x - zc.x, bi = z.y + zc.y;
      acc0[k] += sqrtf(fmaf(ar, ar, ai * ai));
      acc1[k] += sqrtf(fmaf(br, br, bi * bi));
    }
  }
  __syncthreads();
  // /A, + 1e-5, ln, * 20 / 2.303 (BscanFFT.cpp:1221-1237); rows 0 and 1 take the value of row 4 (:1239-1240)
  float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
  for (int rr = 0; rr < (has1 ? 2 : 1); ++rr) {
    const float* acc = rr ? acc1 : acc0;
    const int row = r0 + rr;
    float* srow = a.scratch + ((size_t)b * a.oph + row) * a.Dp;
    const float db4 = logf(fmaf(acc[4], a.out_scale, 1e-5f)) * a.db_scale_ln;
    for (int k = threadIdx.x; k < a.D; k += blockDim.x) {
      float db = logf(fmaf(acc[k], a.out_scale, 1e-5f)) * a.db_scale_ln;
      if (k < 2) {
        if (a.dc01) a.dc01[2 * ((size_t)b * a.oph + row) + k] = db;
        db = db4;
      }
      srow[k] = db;
      if (!(a.clamp55 && k == 5 && row == 5)) {  // the forced element is excluded from the min / max of the data
        mn = fminf(mn, db);
        mx = fmaxf(mx, db);
      }
    }
  }
  mn = block_reduce(mn, red, 1);
  mx = block_reduce(mx, red, 2);
  if (threadIdx.x == 0 && mn <= mx) {  // max(., thr) commutes with min / max (BscanFFT.cpp:1247, 1254)
    atomicMin(a.minv + b, f2ord(fmaxf(mn, a.thr)));
    atomicMax(a.maxv + b, f2ord(fmaxf(mx, a.thr)));
  }
}
// threshold, global min-max normalise, round-half-even to u8 (BscanFFT.cpp:1243-1255), transposed to depth-major; 32 x 32 tiles
__global__ void generic_norm_kernel(const float* __restrict__ scratch, const int* __restrict__ minv, const int* __restrict__ maxv,
                                    uint8_t* __restrict__ out8, float* __restrict__ outdb, int oph, int D, int Dp, float thr, int clamp55,
                                    float clamp_db) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, bin0 = 32 * blockIdx.x, row0 = 32 * blockIdx.y;
  float mn = ord2f(minv[b]), mx = ord2f(maxv[b]);
  if (clamp55) {  // bscandisp.at<double>(5,5) = 50.0 before the min-max (BscanFFT.cpp:1248-1253)
    mn = fminf(mn, clamp_db);
    mx = fmaxf(mx, clamp_db);
  }
  const float sc = (mx - mn) > 2.220446049250313e-16f ? 255.0f / (mx - mn) : 0.f;  // cv::normalize: scale = 0 for a flat image
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int row = row0 + j, bin = bin0 + threadIdx.x;
    tile[j][threadIdx.x] = (row < oph && bin < D) ? scratch[((size_t)b * oph + row) * Dp + bin] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int bin = bin0 + j, row = row0 + threadIdx.x;
    if (bin < D && row < oph) {
      const float x = tile[threadIdx.x][j];
      const float v = (clamp55 && bin == 5 && row == 5) ? clamp_db : x;
      const float r = fmaf(fmaxf(v, thr) - mn, sc, 12582912.0f);  // round-half-even: 1.5 * 2^23 trick, result in the low byte
      const size_t o = ((size_t)b * D + bin) * oph + row;
      out8[o] = (uint8_t)(__float_as_uint(r) & 0xffu);
      if (outdb) outdb[o] = x;
    }
  }
}
__global__ void generic_minmax_reset_kernel(int* minv, int* maxv, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    minv[i] = f2ord(__int_as_float(0x7f800000));
    maxv[i] = f2ord(__int_as_float(0xff800000));
  }
}
size_t generic_smem_bytes(int N, int D) { return 2 * (size_t)fft_buf_slots(N) * sizeof(float2) + 2 * (size_t)D * sizeof(float); }
cudaError_t launch_generic(const GenericHost& h, cudaStream_t st, int* launched) {
  GenericArgs a{};
  a.rows = h.rows; a.M = h.M; a.N = h.N; a.D = h.D; a.Dp = h.Dp; a.oph = h.oph; a.A = h.A; a.nB = h.nB; a.idx = h.idx; a.wq = h.wq;
  a.tw = h.tw; a.scratch = h.scratch; a.minv = h.minv; a.maxv = h.maxv; a.dc01 = h.dc01; a.out_scale = h.out_scale;
  a.db_scale_ln = h.db_scale_ln; a.thr = h.thr; a.clamp55 = h.clamp55; a.single = h.single_row; a.rows64 = h.rows64; a.wq64 = h.wq64;
  a.rl.n = h.N; a.rl.count = h.nrad;
  int nc = h.N, stp = 1;
  for (int i = 0; i < h.nrad && i < 12; ++i) {
    a.rl.r[i] = h.rad[i];
    a.rl.nc[i] = nc;
    a.rl.s[i] = stp;
    nc /= h.rad[i];
    stp *= h.rad[i];
  }
  const size_t smem = generic_smem_bytes(h.N, h.D);
  cudaError_t e = cudaFuncSetAttribute(generic_recon_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  generic_minmax_reset_kernel<<<(h.nB + 127) / 128, 128, 0, st>>>(h.minv, h.maxv, h.nB);
  generic_recon_kernel<<<dim3(h.single_row ? h.oph : (h.oph + 1) / 2, h.nB), 256, smem, st>>>(a);
  generic_norm_kernel<<<dim3((h.D + 31) / 32, (h.oph + 31) / 32, h.nB), dim3(32, 8), 0, st>>>(h.scratch, h.minv, h.maxv, h.out8, h.outdb, h.oph,
                                                                                                h.D, h.Dp, h.thr, h.clamp55, h.clamp_db);
  if (launched) *launched = 3;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ debug tap
// data_ylin (BscanFFT.cpp:1151-1177) from prepared rows, for the stage-level parity test only: idx / wq are the kernel's
// remapped gather tables (idx >= 1, idx == M -> the never-written end points, weight quirk already applied).
__global__ void lerp_rows_kernel(const float* __restrict__ rows, const int* __restrict__ idx, const float* __restrict__ wq,
                                 float* __restrict__ ylin, int M, int N, int oph) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (q >= N || r >= oph) return;
  const float* y = rows + (size_t)r * M;
  const int i = idx[q];
  ylin[(size_t)r * N + q] = i >= M ? 0.f : fmaf(wq[q], y[i] - y[i - 1], y[i]);
}
// the same in f64 (normalised-calibration regime): rows as doubles (rows64) or floats, weights as doubles, one rounding at the end
__global__ void lerp_rows64_kernel(const float* __restrict__ rows, const double* __restrict__ rows64, const int* __restrict__ idx,
                                   const double* __restrict__ wq, float* __restrict__ ylin, int M, int N, int oph) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (q >= N || r >= oph) return;
  const int i = idx[q];
  float v = 0.f;
  if (i < M) {
    const size_t o = (size_t)r * M + i;
    const double yi = rows64 ? rows64[o] : (double)rows[o], ym = rows64 ? rows64[o - 1] : (double)rows[o - 1];
    v = (float)__dadd_rn(yi, __dmul_rn(wq[q], yi - ym));
  }
  ylin[(size_t)r * N + q] = v;
}
cudaError_t launch_lerp_rows64(const float* rows, const double* rows64, const int* idx, const double* wq, float* ylin, int M, int N, int oph,
                               cudaStream_t st) {
  lerp_rows64_kernel<<<dim3((N + 255) / 256, oph), 256, 0, st>>>(rows, rows64, idx, wq, ylin, M, N, oph);
  return cudaGetLastError();
}
cudaError_t launch_lerp_rows(const float* rows, const int* idx, const float* wq, float* ylin, int M, int N, int oph, cudaStream_t st) {
  lerp_rows_kernel<<<dim3((N + 255) / 256, oph), 256, 0, st>>>(rows, idx, wq, ylin, M, N, oph);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ webcam channel sum
// BscanFFTwebcam.cpp:1021-1037 (channelnum >= 3): mraw = (B + G + R) * 0.00130718954 as CV_64F.  The integer sum (<= 765) is
// exact in 16 bits; the scale is applied where the pixel becomes a float (PrepArgs::px_scale).
__global__ void bgr_sum_kernel(const uint8_t* __restrict__ in, uint16_t* __restrict__ out, int w, int h, size_t row_stride, size_t frame_stride) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
  if (x >= w) return;
  const uint8_t* p = in + (size_t)f * frame_stride + (size_t)y * row_stride + 3 * (size_t)x;
  out[((size_t)f * h + y) * w + x] = (uint16_t)((unsigned)p[0] + (unsigned)p[1] + (unsigned)p[2]);
}
cudaError_t launch_bgr_sum(const void* in, uint16_t* out, int w, int h, size_t row_stride_bytes, size_t frame_stride_bytes, int nframes,
                           cudaStream_t st) {
  dim3 grid((w + 127) / 128, h, nframes), block(128);
  bgr_sum_kernel<<<grid, block, 0, st>>>(static_cast<const uint8_t*>(in), out, w, h, row_stride_bytes, frame_stride_bytes);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ calibration captures
// accumulate(data_y, baccum) over the frames of a capture (keys b / p / o / r / t, BscanFFT.cpp:1041-1046): data_y is the binned
// integer pixel converted to CV_64F (x px_scale for the webcam channel sum) and smoothed by smoothmovavg (BscanFFT.cpp:247-304, f64,
// same order of additions as the reference: taps -n .. n with missing taps replaced by the centre, centre once more, / 2 / (n + 1)).
// One thread per binned pixel, frames in order: bit-identical to the host loop it replaces.
template <class T>
__global__ void cal_accum_kernel(const T* __restrict__ px, size_t rs, size_t fs, int nframes, int opw, int oph, int movavgn, double px_scale,
                                 double* __restrict__ acc) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= opw || y >= oph) return;
  double a = 0.0;
  for (int f = 0; f < nframes; ++f) {
    const T* row = px + (size_t)f * fs + (size_t)y * rs;
    const double c = (double)row[x] * px_scale;
    double v = c;
    if (movavgn > 0) {
      double s = 0.0;
      for (int k = -movavgn; k <= movavgn; ++k) {
        const int jj = x + k;
        s = s + ((jj > -1 && jj < opw) ? (double)row[jj] * px_scale : c);
      }
      s = s + c;
      v = s / 2 / (movavgn + 1);
    }
    a += v;
  }
  acc[(size_t)y * opw + x] = a;
}
cudaError_t launch_cal_accum(const void* px, int bpp, size_t row_stride_elems, size_t frame_stride_elems, int nframes, int opw, int oph,
                             int movavgn, double px_scale, double* acc, cudaStream_t st) {
  dim3 grid((opw + 127) / 128, oph), block(128);
  if (bpp == 8)
    cal_accum_kernel<uint8_t><<<grid, block, 0, st>>>(static_cast<const uint8_t*>(px), row_stride_elems, frame_stride_elems, nframes, opw, oph, movavgn, px_scale, acc);
  else
    cal_accum_kernel<uint16_t><<<grid, block, 0, st>>>(static_cast<const uint16_t*>(px), row_stride_elems, frame_stride_elems, nframes, opw, oph, movavgn, px_scale, acc);
  return cudaGetLastError();
}

// The once-per-capture tail on the accumulated frame (BscanFFT.cpp:1050-1057, 1092-1096; BscanDark.cpp:1070-1074), all f64 and in
// the reference's operation order, so the captures stay exact to the last bit of cv::normalize / Mat-by-scalar arithmetic:
//   normalizerows / normalize(., a, b, NORM_MINMAX): scale = (b - a) / (max - min) (0 for a flat image), shift = a - min * scale,
//   dst = src * scale + shift (two roundings);  Mat / n: one multiplication by 1 / n.
__global__ void __launch_bounds__(256) cal_rows_normalize_kernel(double* x, int cols, double a, double b) {
  __shared__ double red[8];
  double* row = x + (size_t)blockIdx.x * cols;
  double mn = 1.7e308, mx = -1.7e308;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) {
    mn = fmin(mn, row[j]);
    mx = fmax(mx, row[j]);
  }
  mn = block_reduce64(mn, red, 1);
  mx = block_reduce64(mx, red, 2);
  const double scale = __dmul_rn(b - a, (mx - mn) > 2.220446049250313e-16 ? 1.0 / (mx - mn) : 0.0);
  const double shift = a - __dmul_rn(mn, scale);
  for (int j = threadIdx.x; j < cols; j += blockDim.x) row[j] = __dadd_rn(__dmul_rn(row[j], scale), shift);
}
__global__ void __launch_bounds__(256) cal_minmax_kernel(const double* __restrict__ x, size_t n, long long* mm) {
  __shared__ double red[8];
  double mn = 1.7e308, mx = -1.7e308;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    mn = fmin(mn, x[i]);
    mx = fmax(mx, x[i]);
  }
  mn = block_reduce64(mn, red, 1);
  mx = block_reduce64(mx, red, 2);
  if (threadIdx.x == 0) {
    atomicMin(mm, d2ord(mn));
    atomicMax(mm + 1, d2ord(mx));
  }
}
// mode 0: x = x * scale + shift with the min / max in mm (global normalise to [a, b]); mode 1: x *= a (Mat / n)
__global__ void __launch_bounds__(256) cal_scale_kernel(double* x, size_t n, const long long* mm, double a, double b, int mode) {
  double scale = a, shift = 0.0;
  if (mode == 0) {
    const double mn = ord2d(mm[0]), mx = ord2d(mm[1]);
    scale = __dmul_rn(b - a, (mx - mn) > 2.220446049250313e-16 ? 1.0 / (mx - mn) : 0.0);
    shift = a - __dmul_rn(mn, scale);
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    x[i] = mode == 0 ? __dadd_rn(__dmul_rn(x[i], scale), shift) : __dmul_rn(x[i], scale);
}
// lpfilter, BscanDark.cpp:119-167: the row as f32, DFT scaled by 1 / cols, keep the centre 20 % of the shifted spectrum (bins
// |k| < floor(cols / 10)), inverse with DFT_REAL_OUTPUT (reads bins 0 .. cols / 2 only), back to f64 through f32:
//   out[n] = Re X[0] + 2 Re sum_{0 < k < keep} X[k] e^{2 pi i n k / cols}.
// Evaluated directly in f64 (keep * cols terms per row - a capture is ONE frame), one CTA per row; dynamic smem: cols + 2 keep doubles.
__global__ void __launch_bounds__(256) cal_lpfilter_kernel(double* x, int cols, int keep, const double* __restrict__ cs, const double* __restrict__ sn) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* v = reinterpret_cast<double*>(smem_raw);
  double* re = v + cols;
  double* im = re + keep;
  double* row = x + (size_t)blockIdx.x * cols;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) v[j] = (double)(float)row[j];  // convertTo(CV_32F)
  __syncthreads();
  for (int k = threadIdx.x; k < keep; k += blockDim.x) {
    double a = 0.0, b = 0.0;
    int t = 0;  // (j * k) mod cols, advanced incrementally
    for (int j = 0; j < cols; ++j) {
      a = __dadd_rn(a, __dmul_rn(v[j], cs[t]));
      b = __dadd_rn(b, -__dmul_rn(v[j], sn[t]));
      t += k;
      if (t >= cols) t -= cols;
    }
    re[k] = a / cols;
    im[k] = b / cols;
  }
  __syncthreads();
  for (int n = threadIdx.x; n < cols; n += blockDim.x) {
    double s = keep > 0 ? re[0] : 0.0;
    int t = 0;
    for (int k = 1; k < keep; ++k) {
      t += n;
      if (t >= cols) t -= cols;
      s = __dadd_rn(s, __dmul_rn(2.0, __dadd_rn(__dmul_rn(re[k], cs[t]), -__dmul_rn(im[k], sn[t]))));
    }
    row[n] = (double)(float)s;  // the inverse transform is f32
  }
}
cudaError_t launch_cal_tail(const CalTailHost& h, cudaStream_t st, int* launched) {
  int nl = 0;
  const size_t n = (size_t)h.rows * h.cols;
  const unsigned gridn = (unsigned)std::min<size_t>((n + 255) / 256, 1184);
  if (h.rowwise) {
    cal_rows_normalize_kernel<<<h.rows, 256, 0, st>>>(h.x, h.cols, h.lo, 1.0);
    ++nl;
  }
  if (h.global_norm) {
    minmax64_reset_kernel<<<1, 32, 0, st>>>(h.mm, 1);
    cal_minmax_kernel<<<gridn, 256, 0, st>>>(h.x, n, h.mm);
    cal_scale_kernel<<<gridn, 256, 0, st>>>(h.x, n, h.mm, h.lo, 1.0, 0);
    nl += 3;
  } else if (h.inv_n != 1.0) {
    cal_scale_kernel<<<gridn, 256, 0, st>>>(h.x, n, h.mm, h.inv_n, 0.0, 1);
    ++nl;
  }
  if (h.lowpass) {
    const int keep = h.cols / 10;
    const size_t smem = ((size_t)h.cols + 2 * (size_t)keep) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(cal_lpfilter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    cal_lpfilter_kernel<<<h.rows, 256, smem, st>>>(h.x, h.cols, keep, h.cs, h.sn);
    ++nl;
  }
  if (launched) *launched = nl;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ launchers
cudaError_t launch_median(const void* in, void* out, int bpp, int k, int w, int h, size_t row_stride_elems, size_t frame_stride_elems,
                          int nframes, cudaStream_t st) {
  dim3 grid((w + 127) / 128, h, nframes), block(128);
  if (bpp == 8 && k == 3)
    median_kernel<uint8_t, 3><<<grid, block, 0, st>>>((const uint8_t*)in, (uint8_t*)out, w, h, row_stride_elems, frame_stride_elems, nframes);
  else if (bpp == 8 && k == 5)
    median_kernel<uint8_t, 5><<<grid, block, 0, st>>>((const uint8_t*)in, (uint8_t*)out, w, h, row_stride_elems, frame_stride_elems, nframes);
  else if (bpp == 16 && k == 3)
    median_kernel<uint16_t, 3><<<grid, block, 0, st>>>((const uint16_t*)in, (uint16_t*)out, w, h, row_stride_elems, frame_stride_elems, nframes);
  else if (bpp == 16 && k == 5)
    median_kernel<uint16_t, 5><<<grid, block, 0, st>>>((const uint16_t*)in, (uint16_t*)out, w, h, row_stride_elems, frame_stride_elems, nframes);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_bin(const void* in, void* out, int bpp, int opw, int oph, int bx, int by, size_t row_stride_elems,
                       size_t frame_stride_elems, int nframes, cudaStream_t st) {
  dim3 grid((opw + 127) / 128, oph, nframes), block(128);
  if (bpp == 8)
    bin_kernel<uint8_t><<<grid, block, 0, st>>>((const uint8_t*)in, (uint8_t*)out, opw, oph, bx, by, row_stride_elems, frame_stride_elems, nframes);
  else
    bin_kernel<uint16_t><<<grid, block, 0, st>>>((const uint16_t*)in, (uint16_t*)out, opw, oph, bx, by, row_stride_elems, frame_stride_elems, nframes);
  return cudaGetLastError();
}

__global__ void minmax_reset_kernel(float* mm, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    reinterpret_cast<int*>(mm)[2 * i] = f2ord(__int_as_float(0x7f800000));
    reinterpret_cast<int*>(mm)[2 * i + 1] = f2ord(__int_as_float(0xff800000));
  }
}

size_t rowprep_smem_bytes(int opw, int M, int m, int movavgn) {
  const size_t rows = 2 * rowbuf_bytes(opw, movavgn);  // two rows per CTA
  if (m <= 1) return rows;
  const size_t one = (size_t)fft_buf_slots(M) * sizeof(float2);
  return one + (rows > one ? rows : one);  // the rows alias the second transform buffer
}

// returns the number of kernels launched (through *launched)
cudaError_t launch_rowprep(const PrepArgsHost& h, cudaStream_t st, int* launched) {
  PrepArgs a{};
  a.binned = h.binned; a.row_stride = h.row_stride; a.frame_stride = h.frame_stride; a.bpp = h.bpp; a.opw = h.opw; a.oph = h.oph; a.nframes = h.nframes; a.movavgn = h.movavgn; a.px_scale = h.px_scale;
  a.yd = h.yd; a.rowwise = h.rowwise; a.frame_minmax = h.frame_minmax; a.yb = h.yb; a.yp = h.yp; a.win = h.win;
  a.m = h.m; a.M = h.M; a.bandpass = h.bandpass; a.twW = h.twW; a.twM = h.twM; a.out = h.out; a.pre = h.pre;
  a.rlW.n = h.opw; a.rlW.count = h.nradW;
  a.rlM.n = h.M; a.rlM.count = h.nradM;
  for (int i = 0; i < 12; ++i) { a.rlW.r[i] = h.radW[i]; a.rlM.r[i] = h.radM[i]; }
  for (RadixList* rl : {&a.rlW, &a.rlM}) {
    int nc = rl->n, st = 1;
    for (int i = 0; i < rl->count && i < 12; ++i) {
      rl->nc[i] = nc;
      rl->s[i] = st;
      nc /= rl->r[i];
      st *= rl->r[i];
    }
  }
  const size_t smem = rowprep_smem_bytes(h.opw, h.M, h.m, h.movavgn);
  static bool attr_done = false;
  if (!attr_done || smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(rowprep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  dim3 grid((h.oph + 1) / 2, h.nframes);  // one CTA per row pair
  int n = 0;
  if (h.global_norm && !h.pre) {
    minmax_reset_kernel<<<(h.nframes + 127) / 128, 128, 0, st>>>(h.frame_minmax, h.nframes);
    a.global_norm = 1;
    rowprep_kernel<<<grid, 256, smem, st>>>(a);
    a.global_norm = 2;
    n += 2;
  }
  rowprep_kernel<<<grid, 256, smem, st>>>(a);
  ++n;
  if (launched) *launched = n;
  return cudaGetLastError();
}

cudaError_t launch_rowprep64(const PrepArgs64Host& h, cudaStream_t st, int* launched) {
  PrepArgs64 a{};
  a.binned = h.binned; a.row_stride = h.row_stride; a.frame_stride = h.frame_stride; a.bpp = h.bpp; a.opw = h.opw; a.oph = h.oph;
  a.nframes = h.nframes; a.movavgn = h.movavgn; a.px_scale = h.px_scale; a.yd = h.yd; a.yb = h.yb; a.yp = h.yp; a.win = h.win;
  a.rowwise = h.rowwise; a.frame_minmax = h.frame_minmax; a.out64 = h.out64; a.out32 = h.out32;
  const size_t smem = (size_t)h.opw * sizeof(double) * (h.movavgn > 0 ? 2 : 1);
  cudaError_t e = cudaFuncSetAttribute(rowprep64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  const dim3 grid(h.oph, h.nframes);
  int n = 0;
  if (h.global_norm) {
    minmax64_reset_kernel<<<(h.nframes + 127) / 128, 128, 0, st>>>(h.frame_minmax, h.nframes);
    a.global_norm = 1;
    rowprep64_kernel<<<grid, 256, smem, st>>>(a);
    a.global_norm = 2;
    n += 2;
  }
  rowprep64_kernel<<<grid, 256, smem, st>>>(a);
  ++n;
  if (launched) *launched = n;
  return cudaGetLastError();
}

}  // namespace abcoct
