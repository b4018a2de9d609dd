// Fused per-frame B-scan reconstruction kernel (sm_100a).
//
// Replaces the reference's inline OpenCV block BscanFFT.cpp:987-1209 + the dB part of 1220-1240 (and the
// BscanDark.cpp:1269 dark subtraction): for every pair of camera rows (two A-scans packed as the real and
// imaginary part of one complex transform) a group of T threads
//   1. reads the raw uint16 pixels once from HBM (128-bit coalesced loads, prefetched one frame ahead),
//   2. applies (y - yd - yp) / yb as one FFMA against calibration rows staged in shared memory by TMA
//      bulk copies (cp.async.bulk + mbarrier), removes the row mean (warp-shuffle + smem reduction) and
//      multiplies by the Bartlett-Hann window (BscanFFT.cpp:1132-1143),
//   3. resamples lambda -> k with the precomputed index/weight tables while loading the first FFT pass
//      from shared memory (BscanFFT.cpp:1151-1177),
//   4. runs the N-point transform as 2 or 3 in-register mixed-radix passes with in-place shared-memory
//      exchanges (replaces cv::dft, BscanFFT.cpp:1185), splits the two A-scans, takes magnitudes and
//      accumulates them over `averages` frames in registers (BscanFFT.cpp:1189-1209),
//   5. converts to dB, applies the DC-row mask (BscanFFT.cpp:1221-1240) and writes an A-scan-major f32
//      scratch image plus a per-B-scan min/max (needed by the global normalize at BscanFFT.cpp:1254).
// A second, tiny kernel (normalise_kernel in abcoct_kernels.cu) thresholds, normalises, transposes and
// quantises to the 8-bit display image (BscanFFT.cpp:1247-1255).
//
// The per-thread phases are plain __host__ __device__ functions so that tests/native/test_group_host.cu can
// run the very same index logic thread-by-thread on the CPU.
#pragma once
#include <cstdint>
#include <cstring>

#include "fft_regs.cuh"
#include "plan.h"

#define ABC_HD __host__ __device__ __forceinline__

namespace abcoct {

constexpr int kFftSign = +1;  // the reference calls dft(..., DFT_INVERSE) without scaling (BscanFFT.cpp:1185)

struct ReconArgs {
  const uint8_t* frames;  // device, uint16 pixels
  unsigned long long frame_stride, row_stride;  // bytes
  int W;        // samples per A-scan (opw); multiple of 8
  int oph;      // A-scans per frame
  int D;        // numdisplaypoints
  int A;        // averages (frames per B-scan)
  int nB;       // B-scans in this launch
  int Gb;       // B-scans per work item (calibration rows are reused for Gb*A frames)
  int npairs;   // ceil(oph / 2)
  int nitems;   // npairs * ceil(nB / Gb)
  const float* gain;  // [oph][W]  1 / yb
  const float* subg;  // [oph][W]  (yd + yp) / yb, or nullptr
  const uint16_t* idxT;
  const float* wqT;
  const float* win;
  const float2* tw0;
  const float2* tw1;
  float* scratch;  // [nB][oph][D] dB values, A-scan major
  int* minmax;     // [nB][2] order-preserving int encodings of min / max
  float inv_W;
  float out_scale;  // 0.5 / A
  float db_scale;   // ln(2) * 20 * (1 / 2.303)
  float thr;        // bscanthreshold
  int clamp55;      // clampupper: element (5,5) is excluded from the min/max here
};

// ------------------------------------------------------------------------------------------- shared memory map
struct SmemLayout {
  int idxT, wqT, win, tw0, tw1;  // CTA-wide tables (byte offsets)
  int groups;                    // start of the per-group blocks
  int g_gain, g_subg, g_stg, g_buf, g_red, g_mbar, group_bytes;
  __host__ __device__ int total(int G) const { return groups + G * group_bytes; }
};
__host__ __device__ constexpr int align16(int x) { return (x + 15) & ~15; }

template <class P>
__host__ __device__ inline SmemLayout make_layout(int W, bool has_sub) {
  SmemLayout L{};
  int o = 0;
  L.idxT = o; o = align16(o + P::R0P8 * P::N1 * 2);
  L.wqT = o;  o = align16(o + P::R0P4 * P::N1 * 4);
  L.win = o;  o = align16(o + W * 4);
  L.tw0 = o;  o = align16(o + (P::R0 - 1) * P::N1 * 8);
  L.tw1 = o;  o = align16(o + (P::THREE ? (P::R1 - 1) * P::N2 * 8 : 0));
  L.groups = o;
  int g = 0;
  L.g_gain = g; g = align16(g + 2 * W * 4);
  L.g_subg = g; g = align16(g + (has_sub ? 2 * W * 4 : 0));
  L.g_stg = g;  g = align16(g + (W + 1) * 8);
  L.g_buf = g;  g = align16(g + P::BUF * 8);
  L.g_red = g;  g = align16(g + 2 * P::NWARPS * 4);
  L.g_mbar = g; g = align16(g + 16);
  L.group_bytes = g;
  return L;
}

struct GroupSmem {  // resolved pointers of one group
  const uint16_t* idxT;
  const float* wqT;
  const float* win;
  const float2* tw0;
  const float2* tw1;
  float* gain;
  float* subg;
  float2* stg;
  float2* buf;
  float* red;
  unsigned long long* mbar;
};
template <class P>
__host__ __device__ inline GroupSmem resolve(unsigned char* base, const SmemLayout& L, int g) {
  unsigned char* gb = base + L.groups + g * L.group_bytes;
  GroupSmem s;
  s.idxT = reinterpret_cast<const uint16_t*>(base + L.idxT);
  s.wqT = reinterpret_cast<const float*>(base + L.wqT);
  s.win = reinterpret_cast<const float*>(base + L.win);
  s.tw0 = reinterpret_cast<const float2*>(base + L.tw0);
  s.tw1 = reinterpret_cast<const float2*>(base + L.tw1);
  s.gain = reinterpret_cast<float*>(gb + L.g_gain);
  s.subg = reinterpret_cast<float*>(gb + L.g_subg);
  s.stg = reinterpret_cast<float2*>(gb + L.g_stg);
  s.buf = reinterpret_cast<float2*>(gb + L.g_buf);
  s.red = reinterpret_cast<float*>(gb + L.g_red);
  s.mbar = reinterpret_cast<unsigned long long*>(gb + L.g_mbar);
  return s;
}

// ------------------------------------------------------------------------------------------- per-thread state
template <class P>
struct ThreadState {
  uint4 raw[2][P::NCH];          // prefetched pixels (8 x u16) of row a / row b
  float t[2][P::NCH][8];         // (y - sub) / yb
  float acc[P::NU][P::RL][2];    // 2 * sum over frames of |A[k]|, |B[k]|
};

ABC_HD float fast_sqrt(float x) {
#ifdef __CUDA_ARCH__
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return sqrtf(x);
#endif
}
ABC_HD float fast_log2(float x) {
#ifdef __CUDA_ARCH__
  return __log2f(x);
#else
  return log2f(x);
#endif
}
ABC_HD uint4 load_raw16(const uint8_t* p) {
#ifdef __CUDA_ARCH__
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
#else
  uint4 v;
  memcpy(&v, p, 16);
  return v;
#endif
}
// order-preserving float <-> int (for atomicMin / atomicMax on floats)
ABC_HD int float_to_ordered(float f) {
  int i;
  memcpy(&i, &f, 4);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
ABC_HD float ordered_to_float(int i) {
  int j = i >= 0 ? i : i ^ 0x7fffffff;
  float f;
  memcpy(&f, &j, 4);
  return f;
}

// ------------------------------------------------------------------------------------------- phases
// raw pixel prefetch of one frame's row pair into registers
template <class P>
ABC_HD void phase_load(int tid, const uint8_t* rowa, const uint8_t* rowb, int W8, ThreadState<P>& r) {
#pragma unroll
  for (int i = 0; i < P::NCH; ++i) {
    int ch = tid + P::T * i;
    if (ch < W8) {
      r.raw[0][i] = load_raw16(rowa + 16 * ch);
      r.raw[1][i] = load_raw16(rowb + 16 * ch);
    }
  }
}

// t = (y - yd - yp) / yb as y * gain - subg, and the per-thread partial row sums
template <class P, bool HAS_SUB>
ABC_HD void phase_pre1(int tid, const GroupSmem& s, int W, ThreadState<P>& r, float& sa, float& sb) {
  const int W8 = W >> 3;
  sa = 0.f;
  sb = 0.f;
#pragma unroll
  for (int i = 0; i < P::NCH; ++i) {
    int ch = tid + P::T * i;
    if (ch < W8) {
#pragma unroll
      for (int row = 0; row < 2; ++row) {
        const uint4 v = r.raw[row][i];
        const unsigned w32[4] = {v.x, v.y, v.z, v.w};
        const float4 g0 = *reinterpret_cast<const float4*>(s.gain + row * W + 8 * ch);
        const float4 g1 = *reinterpret_cast<const float4*>(s.gain + row * W + 8 * ch + 4);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        float q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if constexpr (HAS_SUB) {
          const float4 q0 = *reinterpret_cast<const float4*>(s.subg + row * W + 8 * ch);
          const float4 q1 = *reinterpret_cast<const float4*>(s.subg + row * W + 8 * ch + 4);
          q[0] = q0.x; q[1] = q0.y; q[2] = q0.z; q[3] = q0.w;
          q[4] = q1.x; q[5] = q1.y; q[6] = q1.z; q[7] = q1.w;
        }
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const unsigned px = (j & 1) ? (w32[j >> 1] >> 16) : (w32[j >> 1] & 0xffffu);
          const float y = static_cast<float>(px);
          float tv;
          if constexpr (HAS_SUB)
            tv = fmaf(y, g[j], -q[j]);
          else
            tv = y * g[j];
          r.t[row][i][j] = tv;
          acc += tv;
        }
        if (row == 0)
          sa += acc;
        else
          sb += acc;
      }
    }
  }
}

// DC removal + apodisation, written pairwise-interleaved (row a -> .x, row b -> .y) for the gather
template <class P>
ABC_HD void phase_pre2(int tid, const GroupSmem& s, int W, const ThreadState<P>& r, float ma, float mb) {
  const int W8 = W >> 3;
#pragma unroll
  for (int i = 0; i < P::NCH; ++i) {
    int ch = tid + P::T * i;
    if (ch < W8) {
      const float4 w0 = *reinterpret_cast<const float4*>(s.win + 8 * ch);
      const float4 w1 = *reinterpret_cast<const float4*>(s.win + 8 * ch + 4);
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      float4* dst = reinterpret_cast<float4*>(s.stg + 8 * ch);
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        float4 o;
        o.x = (r.t[0][i][j] - ma) * w[j];
        o.y = (r.t[1][i][j] - mb) * w[j];
        o.z = (r.t[0][i][j + 1] - ma) * w[j + 1];
        o.w = (r.t[1][i][j + 1] - mb) * w[j + 1];
        dst[j >> 1] = o;
      }
    }
  }
}

// pass 0: lambda->k gather-lerp (BscanFFT.cpp:1169-1171) fused into the first radix-R0 butterflies
template <class P>
ABC_HD void phase_pass0(int tid, const GroupSmem& s) {
#pragma unroll
  for (int i = 0; i < P::NB0; ++i) {
    const int b = tid + P::T * i;
    if (P::NB0 * P::T == P::N1 || b < P::N1) {
      float2 in[P::R0], out[P::R0];
      unsigned short idx[P::R0P8];
      float wq[P::R0P4];
#pragma unroll
      for (int c8 = 0; c8 < P::R0P8 / 8; ++c8) {
        const uint4 v = *reinterpret_cast<const uint4*>(s.idxT + (c8 * P::N1 + b) * 8);
        const unsigned w32[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 8; ++j)
          idx[c8 * 8 + j] = static_cast<unsigned short>((j & 1) ? (w32[j >> 1] >> 16) : (w32[j >> 1] & 0xffffu));
      }
#pragma unroll
      for (int c4 = 0; c4 < P::R0P4 / 4; ++c4) {
        const float4 v = *reinterpret_cast<const float4*>(s.wqT + (c4 * P::N1 + b) * 4);
        wq[c4 * 4 + 0] = v.x; wq[c4 * 4 + 1] = v.y; wq[c4 * 4 + 2] = v.z; wq[c4 * 4 + 3] = v.w;
      }
#pragma unroll
      for (int a = 0; a < P::R0; ++a) {
        const float2 y1 = s.stg[idx[a]];
        const float2 y0 = s.stg[idx[a] - 1];
        in[a].x = fmaf(wq[a], y1.x - y0.x, y1.x);
        in[a].y = fmaf(wq[a], y1.y - y0.y, y1.y);
      }
      Dft<P::R0, kFftSign, 1, 1>::run(in, out);
      s.buf[b] = out[0];
#pragma unroll
      for (int c = 1; c < P::R0; ++c) s.buf[P::ROWSTRIDE * c + b] = cmul(out[c], s.tw0[(c - 1) * P::N1 + b]);
    }
  }
}

// pass 1 (three-pass plans only), in place
template <class P>
ABC_HD void phase_pass1(int tid, const GroupSmem& s) {
  if constexpr (P::THREE) {
#pragma unroll
    for (int i = 0; i < P::NB1; ++i) {
      const int x = tid + P::T * i;
      if (P::NB1 * P::T == P::NBF1 || x < P::NBF1) {
        const int c = x % P::R0, bp = x / P::R0;
        float2* base = s.buf + P::ROWSTRIDE * c + bp;
        float2 in[P::R1], out[P::R1];
#pragma unroll
        for (int a = 0; a < P::R1; ++a) in[a] = base[P::N2 * a];
        Dft<P::R1, kFftSign, 1, 1>::run(in, out);
        base[0] = out[0];
#pragma unroll
        for (int c1 = 1; c1 < P::R1; ++c1) base[P::N2 * c1] = cmul(out[c1], s.tw1[(c1 - 1) * P::N2 + bp]);
      }
    }
  }
}

// unit -> output bins (shared by accumulate and finalise). Regular units u >= 1 pair butterflies (u, S-u);
// unit 0 carries the two self-conjugate butterflies k0 = 0 and k0 = S/2.
template <class P>
ABC_HD int unit_bin(int u, int j) {
  constexpr int RL = P::RL, S = P::S, JH = (RL + 1) / 2;
  if (u != 0) return j < JH ? u + S * j : (S - u) + S * (RL - 1 - j);
  return j < JH ? S * j : S / 2 + S * (j - JH);
}

// last pass + two-for-one split + magnitude + accumulation (BscanFFT.cpp:1185-1197)
template <class P>
ABC_HD void phase_passL(int tid, const GroupSmem& s, ThreadState<P>& r) {
  constexpr int RL = P::RL, S = P::S, JH = (RL + 1) / 2;
#pragma unroll
  for (int i = 0; i < P::NU; ++i) {
    const int u = tid + P::T * i;
    if (P::NU * P::T == P::NUNITS || u < P::NUNITS) {
      const int kA = u, kB = (u == 0) ? S / 2 : S - u;
      const float2* pa = s.buf + P::ROWSTRIDE * (kA % P::R0) + RL * (kA / P::R0);
      const float2* pb = s.buf + P::ROWSTRIDE * (kB % P::R0) + RL * (kB / P::R0);
      float2 za[RL], zb[RL], Za[RL], Zb[RL];
#pragma unroll
      for (int a = 0; a < RL; ++a) {
        za[a] = pa[a];
        zb[a] = pb[a];
      }
      Dft<RL, kFftSign, 1, 1>::run(za, Za);
      Dft<RL, kFftSign, 1, 1>::run(zb, Zb);
      if (u != 0) {
#pragma unroll
        for (int j = 0; j < RL; ++j) {
          const float2 Pz = Za[j], Qz = Zb[RL - 1 - j];
          const float sr = Pz.x + Qz.x, di = Pz.y - Qz.y, si = Pz.y + Qz.y, dr = Pz.x - Qz.x;
          r.acc[i][j][0] += fast_sqrt(fmaf(sr, sr, di * di));
          r.acc[i][j][1] += fast_sqrt(fmaf(si, si, dr * dr));
        }
      } else {
#pragma unroll
        for (int j = 0; j < RL; ++j) {
          float2 Pz, Qz;
          if (j < JH) {
            Pz = Za[j];
            Qz = Za[(RL - j) % RL];
          } else {
            Pz = Zb[j - JH];
            Qz = Zb[RL - 1 - (j - JH)];
          }
          const float sr = Pz.x + Qz.x, di = Pz.y - Qz.y, si = Pz.y + Qz.y, dr = Pz.x - Qz.x;
          r.acc[i][j][0] += fast_sqrt(fmaf(sr, sr, di * di));
          r.acc[i][j][1] += fast_sqrt(fmaf(si, si, dr * dr));
        }
      }
    }
  }
}

// average, +1e-5, ln -> dB, DC-row mask, min/max of the thresholded value (BscanFFT.cpp:1221-1247)
template <class P>
ABC_HD void phase_finalise(int tid, const ReconArgs& a, float* rowa_out, float* rowb_out, int row_a_index,
                           bool rowb_valid, ThreadState<P>& r, float& mn, float& mx) {
#pragma unroll
  for (int i = 0; i < P::NU; ++i) {
    const int u = tid + P::T * i;
    if (P::NU * P::T == P::NUNITS || u < P::NUNITS) {
#pragma unroll
      for (int j = 0; j < P::RL; ++j) {
        const int kk = unit_bin<P>(u, j);
#pragma unroll
        for (int row = 0; row < 2; ++row) {
          const float v = fmaf(r.acc[i][j][row], a.out_scale, 1e-5f);
          const float db = fast_log2(v) * a.db_scale;
          r.acc[i][j][row] = 0.f;
          float* dst = row ? rowb_out : rowa_out;
          const bool valid = (kk < a.D) && (row == 0 || rowb_valid);
          if (valid && kk >= 2) {
            dst[kk] = db;
            if (kk == 4) {  // bscandb.row(4).copyTo(row(1)), row(0): BscanFFT.cpp:1239-1240
              dst[0] = db;
              dst[1] = db;
            }
            const bool is55 = a.clamp55 && kk == 5 && (row_a_index + row) == 5;
            if (!is55) {
              const float c = fmaxf(db, a.thr);
              mn = fminf(mn, c);
              mx = fmaxf(mx, c);
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------- device-only glue
#ifdef __CUDACC__
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
template <int T>
__device__ __forceinline__ void group_sync(int g) {
  if constexpr (T == 32)
    __syncwarp();
  else
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(T) : "memory");
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <class P, int GMAX, bool HAS_SUB, int MINB>
__global__ void __launch_bounds__(P::T* GMAX, MINB) recon_kernel(const ReconArgs a, const int G) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout L = make_layout<P>(a.W, HAS_SUB);
  const int g = threadIdx.x / P::T;
  const int tid = threadIdx.x - g * P::T;
  const int lane = tid & 31, wrp = tid >> 5;
  const GroupSmem s = resolve<P>(smem, L, g);
  const int W = a.W, W8 = W >> 3;

  // ---- CTA-wide tables: global (L2-resident) -> shared, once per persistent CTA
  {
    const int n16 = L.groups >> 4;  // tables are laid out contiguously in the same order in global memory
    const uint4* src = reinterpret_cast<const uint4*>(a.idxT);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
  }
  if (tid == 0) {
    s.stg[W] = make_float2(0.f, 0.f);  // sentinel read by the never-written end points q = 0 and q = N-1
    mbar_init(s.mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  ThreadState<P> r;
#pragma unroll
  for (int i = 0; i < P::NU; ++i)
#pragma unroll
    for (int j = 0; j < P::RL; ++j) r.acc[i][j][0] = r.acc[i][j][1] = 0.f;

  unsigned cal_parity = 0;
  for (int item = blockIdx.x * G + g; item < a.nitems; item += gridDim.x * G) {
    const int pair = item % a.npairs, bg = item / a.npairs;
    const int ra = 2 * pair;
    const bool rowb_valid = (ra + 1) < a.oph;
    const int rb = rowb_valid ? ra + 1 : ra;
    const int b0 = bg * a.Gb;
    const int nb_here = min(a.Gb, a.nB - b0);

    group_sync<P::T>(g);  // every thread is done with the previous item's calibration rows and buffers
    if (tid == 0) {
      const unsigned rowbytes = static_cast<unsigned>(W) * 4u;
      mbar_expect_tx(s.mbar, rowbytes * (HAS_SUB ? 4u : 2u));
      bulk_g2s(s.gain, a.gain + static_cast<size_t>(ra) * W, rowbytes, s.mbar);
      bulk_g2s(s.gain + W, a.gain + static_cast<size_t>(rb) * W, rowbytes, s.mbar);
      if constexpr (HAS_SUB) {
        bulk_g2s(s.subg, a.subg + static_cast<size_t>(ra) * W, rowbytes, s.mbar);
        bulk_g2s(s.subg + W, a.subg + static_cast<size_t>(rb) * W, rowbytes, s.mbar);
      }
    }
    const uint8_t* f0 = a.frames + static_cast<size_t>(b0) * a.A * a.frame_stride;
    const uint8_t* pa = f0 + static_cast<size_t>(ra) * a.row_stride;
    const uint8_t* pb = f0 + static_cast<size_t>(rb) * a.row_stride;
    phase_load<P>(tid, pa, pb, W8, r);
    while (!mbar_try_wait(s.mbar, cal_parity)) {
    }
    cal_parity ^= 1u;

    const int nframes_item = nb_here * a.A;
    int fa = 0;  // frame index inside the current B-scan
    int bi = 0;
    for (int f = 0; f < nframes_item; ++f) {
      float sa, sb;
      phase_pre1<P, HAS_SUB>(tid, s, W, r, sa, sb);
      sa = warp_sum(sa);
      sb = warp_sum(sb);
      if constexpr (P::NWARPS > 1) {
        if (lane == 0) {
          s.red[2 * wrp] = sa;
          s.red[2 * wrp + 1] = sb;
        }
        group_sync<P::T>(g);
        sa = 0.f;
        sb = 0.f;
#pragma unroll
        for (int w = 0; w < P::NWARPS; ++w) {
          sa += s.red[2 * w];
          sb += s.red[2 * w + 1];
        }
      } else {
        group_sync<P::T>(g);
      }
      phase_pre2<P>(tid, s, W, r, sa * a.inv_W, sb * a.inv_W);
      if (f + 1 < nframes_item) {  // prefetch the next frame's rows while this one is transformed
        pa += a.frame_stride;
        pb += a.frame_stride;
        phase_load<P>(tid, pa, pb, W8, r);
      }
      group_sync<P::T>(g);
      phase_pass0<P>(tid, s);
      group_sync<P::T>(g);
      if constexpr (P::THREE) {
        phase_pass1<P>(tid, s);
        group_sync<P::T>(g);
      }
      phase_passL<P>(tid, s, r);

      if (++fa == a.A) {
        fa = 0;
        const int bscan = b0 + bi;
        ++bi;
        float* oa = a.scratch + (static_cast<size_t>(bscan) * a.oph + ra) * a.D;
        float* ob = oa + a.D;
        float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
        phase_finalise<P>(tid, a, oa, ob, ra, rowb_valid, r, mn, mx);
        mn = warp_min(mn);
        mx = warp_max(mx);
        if (lane == 0 && mn <= mx) {
          atomicMin(a.minmax + 2 * bscan, float_to_ordered(mn));
          atomicMax(a.minmax + 2 * bscan + 1, float_to_ordered(mx));
        }
      }
    }
  }
}
#endif  // __CUDACC__

}  // namespace abcoct
