// Fused per-frame B-scan reconstruction kernel (sm_100a): the whole block BscanFFT.cpp:987-1255 in one launch.
//
// Persistent grid, one CTA per SM, G independent thread groups per CTA.  A group of T threads takes "items" - one pair of
// camera rows (two A-scans packed as the real and imaginary part of one complex transform) of one B-scan - from a global
// ticket and, per frame of the item,
//   1. (phase_pre) takes the raw uint16 pixels read once from HBM (128-bit loads, prefetched one frame ahead, L2
//      evict-first), applies (y - yd - yp) / yb - 1 as one FFMA against calibration rows staged in shared memory by TMA bulk
//      copies (cp.async.bulk + mbarrier), multiplies by the Bartlett-Hann window and stages the result; the row-mean removal
//      (BscanFFT.cpp:1132-1143) is linear and is carried through the resampling as a second term,
//   2. (phase_gather) resamples lambda -> k with the precomputed offset / weight tables straight into the registers of the
//      first FFT pass (BscanFFT.cpp:1151-1177),
//   3. (phase_pass0 / 1 / L) runs the N-point transform as 2 or 3 in-register mixed-radix passes with in-place shared-memory
//      exchanges (replaces cv::dft, BscanFFT.cpp:1185), splits the two A-scans, takes magnitudes and accumulates them over
//      `averages` frames in registers (BscanFFT.cpp:1189-1209),
//   4. (phase_finalise, last frame) converts to dB, applies the DC-row mask (BscanFFT.cpp:1221-1240), writes an A-scan-major
//      f32 scratch row that is meant to live in L2, and folds its min / max into the B-scan's (BscanFFT.cpp:1247, 1254).
// The second half runs in the same kernel: once every pair of a B-scan has been published (a counter in global memory) the
// groups pick up its normalisation jobs (normalise_part): threshold, min-max normalise, transpose to depth-major, quantise
// to the 8-bit display image (BscanFFT.cpp:1243-1255), and drop the consumed scratch lines from L2.
//
// The per-thread phases are plain __host__ __device__ functions so that tests/native/test_group_host.cu can run the very
// same index logic thread-by-thread on the CPU.
#pragma once
#include <cstdint>
#include <cstring>

#include "fft_regs.cuh"
#include "plan.h"

#define ABC_HD __host__ __device__ __forceinline__
#ifndef ABC_PUBLISH_BATCH
#define ABC_PUBLISH_BATCH 2
#endif

namespace abcoct {

constexpr int kFftSign = +1;  // the reference calls dft(..., DFT_INVERSE) without scaling (BscanFFT.cpp:1185)

struct ReconArgs {
  const uint8_t* frames;  // device, uint16 pixels
  unsigned long long frame_stride, row_stride;  // bytes
  int W;        // samples per A-scan (opw); multiple of 8
  int oph;      // A-scans per frame
  int D;        // numdisplaypoints
  int Dp;       // scratch pitch in floats: D rounded up to 32 (whole 128-byte lines per A-scan)
  int A;        // averages (frames per B-scan)
  int nB;       // B-scans in this launch
  int npairs;   // ceil(oph / 2)
  int nitems;   // npairs * nB; item = bscan * npairs + pair
  int nparts;   // normalisation jobs per B-scan: ceil(oph / T), T = threads per group of the plan
  const float* gain;  // [oph][W]  1 / yb
  const float* subg;  // [oph][W]  (yd + yp) / yb, or nullptr
  const uint32_t* idxT;  // the table blob (shared-memory image), starts with the gather offsets
  float* scratch;  // [nB][oph][Dp] dB values, A-scan major; lives in L2 between the two halves of the kernel
  int* sched;      // [0] next item ticket; per B-scan arrays follow (see SchedView)
  uint8_t* out8;   // [nB][D][oph] display image (BscanFFT.cpp:1254-1255)
  float* outdb;    // nullable, [nB][D][oph] bscandb (BscanFFT.cpp:1237-1240)
  float* dc01;     // nullable, [nB][oph][2]: the dB value of bins 0 and 1 before the DC-row mask (for the linear `bscan` output)
  float inv_W;
  float out_scale;  // 0.5 / A
  float db_scale;   // ln(2) * 20 * (1 / 2.303)
  float thr;        // bscanthreshold
  float clamp_db;   // value forced into element (5,5) when clampupper
  int clamp55;      // clampupper: element (5,5) is excluded from the min/max of the data
  // warp-per-A-scan kernel (wrow_kernel.cuh) only
  int calpitch;     // floats per calibration row in its permuted layout
  int nsplit;       // depth-tile ranges a normalisation part is split into (small launches: more, shorter jobs)
  int ringB;        // wrow_kernel: the dB scratch holds this many B-scans and is reused round-robin (0: one region per B-scan)
  int hints;        // A/B switches (ABCOCT_HINTS).  L2 policies of the scratch kernel: 1 pixel-row prefetch evict_first, 2 scratch stores
                    // evict_last, 4 pixel loads evict_first (measured: no effect).  Resident-row kernel: 8 = every warp writes its own row's bytes
};

// Scheduler / per-B-scan state in global memory (ints).  Header of kSchedHeader ints - three 128-byte lines so that the hot
// words never share a line: [0] item ticket, [32] normalisation jobs handed out, [64] job frontier (wrow_kernel.cuh) - then nB
// each of minv, maxv (order-preserving int encodings) and cnt (rows / row pairs finished and published).
constexpr int kSchedHeader = 96;
constexpr int kNormBins = 32;   // depth bins per transposition tile
struct SchedView {
  int* ticket;
  int* minv;
  int* maxv;
  int* cnt;
};
__host__ __device__ inline SchedView sched_view(int* base, int nB) {
  SchedView v;
  v.ticket = base;
  v.minv = base + kSchedHeader;
  v.maxv = v.minv + nB;
  v.cnt = v.maxv + nB;
  return v;
}
// + one 128-byte line per B-scan: the row counts of the resident-row kernel (wres_kernel.cuh), polled from every SM
__host__ __device__ inline size_t sched_ints(int nB) { return kSchedHeader + 3 * (size_t)nB + 32 * (size_t)nB; }


// ------------------------------------------------------------------------------------------- shared memory map
struct SmemLayout {
  int idxT, wqT, vwT, win, tw0, tw1;  // CTA-wide tables (byte offsets)
  int groups;                    // start of the per-group blocks
  // g_buf is ONE region used in turn as the staging buffer of the apodised samples (phase_pre -> gather), as the FFT
  // exchange buffer (pass 0 -> last pass) and as the transposition tile of a normalisation job
  int g_gain, g_subg, g_buf, g_red, g_mbar, group_bytes;
  __host__ __device__ constexpr int total(int G) const { return groups + G * group_bytes; }
};
__host__ __device__ constexpr int align16(int x) { return (x + 15) & ~15; }

template <class P>
__host__ __device__ constexpr SmemLayout make_layout(int W, bool has_sub) {
  SmemLayout L{};
  int o = 0;
  L.idxT = o; o = align16(o + P::R0P4 * P::N1 * 4);
  L.wqT = o;  o = align16(o + P::R0P4 * P::N1 * 4);
  L.vwT = o;  o = align16(o + P::R0P4 * P::N1 * 4);
  L.win = o;  o = align16(o + W * 4);
  L.tw0 = o;  o = align16(o + (P::R0 - 1) * P::N1 * 8);
  L.tw1 = o;  o = align16(o + (P::THREE ? (P::R1 - 1) * P::N2 * 8 : 0));
  L.groups = o;
  int g = 0;
  L.g_gain = g; g = align16(g + 2 * W * 4);
  L.g_subg = g; g = align16(g + (has_sub ? 2 * W * 4 : 0));
  L.g_buf = g;  g = align16(g + cmax(cmax((W + 1) * 8, P::BUF * 8), kNormBins * (P::T + 4)));
  L.g_red = g;  g = align16(g + 2 * P::NWARPS * 4 + 48);
  L.g_mbar = g; g = align16(g + 16);
  L.group_bytes = g;
  return L;
}

struct GroupSmem {  // resolved pointers of one group
  const uint32_t* idxT;  // per gathered sample: byte offset of y[i] | byte offset of y[i-1] << 16 (swizzled staging)
  const float* wqT;      // per gathered sample: lerp weight
  const float* vwT;      // per gathered sample: the same lerp applied to the window (carries the row-mean removal)
  const float* win;
  const float2* tw0;
  const float2* tw1;
  float* gain;
  float* subg;
  float2* stg;
  float2* buf;
  float* red;
  int* slot;  // 12 ints: broadcast slots of the group leader (next items, normalise job) and its unpublished B-scans
  unsigned long long* mbar;
};
template <class P>
__host__ __device__ inline GroupSmem resolve(unsigned char* base, const SmemLayout& L, int g) {
  unsigned char* gb = base + L.groups + g * L.group_bytes;
  GroupSmem s;
  s.idxT = reinterpret_cast<const uint32_t*>(base + L.idxT);
  s.wqT = reinterpret_cast<const float*>(base + L.wqT);
  s.vwT = reinterpret_cast<const float*>(base + L.vwT);
  s.win = reinterpret_cast<const float*>(base + L.win);
  s.tw0 = reinterpret_cast<const float2*>(base + L.tw0);
  s.tw1 = reinterpret_cast<const float2*>(base + L.tw1);
  s.gain = reinterpret_cast<float*>(gb + L.g_gain);
  s.subg = reinterpret_cast<float*>(gb + L.g_subg);
  s.stg = reinterpret_cast<float2*>(gb + L.g_buf);  // same region, see SmemLayout
  s.buf = reinterpret_cast<float2*>(gb + L.g_buf);
  s.red = reinterpret_cast<float*>(gb + L.g_red);
  s.slot = reinterpret_cast<int*>(gb + L.g_red + 2 * P::NWARPS * 4);
  s.mbar = reinterpret_cast<unsigned long long*>(gb + L.g_mbar);
  return s;
}

// ------------------------------------------------------------------------------------------- bank-conflict swizzles
// Staging buffer of the apodised samples (float2 = rows a,b interleaved): every thread stores the four 16-byte units
// of its 8-sample chunk ch with one STS.128 each; lanes are 64 bytes apart, so unit j of chunk ch lives at unit
// j ^ ((ch >> 1) & 3) to spread a quarter-warp over all 32 banks.  In sample indices: bits 1..2 ^= bits 4..5.
ABC_HD int stg_phys(int i) { return i ^ (((i >> 4) & 3) << 1); }
// Calibration rows and window (f32, 8 floats per chunk read as two LDS.128, lanes 32 bytes apart): the two 16-byte
// halves of chunk ch are swapped when bit 2 of ch is set.  In float indices: bit 2 ^= bit 5.
ABC_HD int cal_phys(int i) { return i ^ (((i >> 5) & 1) << 2); }
inline void cal_swizzle_row(const float* in, float* out, int W) {
  for (int i = 0; i < W; ++i) out[cal_phys(i)] = in[i];
}

// ------------------------------------------------------------------------------------------- per-thread state
template <class P>
struct ThreadState {
  uint4 raw[2][P::NCH];          // prefetched pixels (8 x u16) of row a / row b
  float2 x[P::NB0][P::R0];       // lambda->k resampled inputs of this thread's first-pass butterflies
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
  float r;  // the argument is >= 1e-5 (BscanFFT.cpp:1222 adds it), never denormal: one MUFU.LG2, no range fix-up
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return log2f(x);
#endif
}
// raw pixels are read exactly once: bypass L1 and mark the lines evict-first in L2 so that they do not push the
// dB scratch (which must stay L2-resident between reconstruction and normalisation) out to HBM
ABC_HD uint4 load_raw16(const uint8_t* p, unsigned long long pol) {
#ifdef __CUDA_ARCH__
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p), "l"(pol));
  return v;
#else
  (void)pol;
  uint4 v;
  memcpy(&v, p, 16);
  return v;
#endif
}
// dB scratch store: the line must survive in L2 until the normalisation half of the kernel has consumed (and discarded)
// it, so it is written with an evict-last policy (`pol`, 0 on the host = plain store)
ABC_HD void store_scratch(float* p, float v, unsigned long long pol) {
#ifdef __CUDA_ARCH__
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
#else
  (void)pol;
  *p = v;
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
ABC_HD void phase_load(int tid, const uint8_t* rowa, const uint8_t* rowb, int W8, ThreadState<P>& r, unsigned long long pol = 0) {
#pragma unroll
  for (int i = 0; i < P::NCH; ++i) {
    int ch = tid + P::T * i;
    if (ch < W8) {
      r.raw[0][i] = load_raw16(rowa + 16 * ch, pol);
      r.raw[1][i] = load_raw16(rowb + 16 * ch, pol);
    }
  }
}

// Fused pre-processing of one frame's row pair (BscanFFT.cpp:987, 1132, 1141; BscanDark.cpp:1269):
//   t = (y - yd - yp) / yb  as  y * gain - subg,   staged value t * window,   partial row sums of t.
// The row-mean removal of BscanFFT.cpp:1135-1139 is linear, so it is applied after the resampling instead:
//   lerp((t - m) w) = lerp((t - 1) w) - (m - 1) lerp(w)      (phase_gather, lerp(w) precomputed per output sample)
// which keeps t out of the registers.  The constant 1 keeps the staged values small: t is the interferogram normalised
// by its background, so its DC level is close to 1 and the two terms do not cancel catastrophically in f32 (the
// identity is exact for any constant).  subg already contains the +1 when HAS_SUB.  Staged pairwise-interleaved (row a -> .x, row b -> .y).
template <class P, bool HAS_SUB>
ABC_HD void phase_pre(int tid, const GroupSmem& s, int W, const ThreadState<P>& r, float& sa, float& sb) {
  const int W8 = W >> 3;
  sa = 0.f;
  sb = 0.f;
  if (tid == 0) s.stg[W] = make_float2(0.f, 0.f);  // sentinel read by the never-written end points q = 0 and q = N-1
#pragma unroll
  for (int i = 0; i < P::NCH; ++i) {
    const int ch = tid + P::T * i;
    if (ch < W8) {
      const int h4 = (ch & 4);  // cal_phys: halves of the chunk swapped when bit 2 of ch is set
      const float4 w0 = *reinterpret_cast<const float4*>(s.win + 8 * ch + h4);
      const float4 w1 = *reinterpret_cast<const float4*>(s.win + 8 * ch + (h4 ^ 4));
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      float tw[2][8];
#pragma unroll
      for (int row = 0; row < 2; ++row) {
        const uint4 v = r.raw[row][i];
        const unsigned w32[4] = {v.x, v.y, v.z, v.w};
        const float4 g0 = *reinterpret_cast<const float4*>(s.gain + row * W + 8 * ch + h4);
        const float4 g1 = *reinterpret_cast<const float4*>(s.gain + row * W + 8 * ch + (h4 ^ 4));
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        float q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if constexpr (HAS_SUB) {
          const float4 q0 = *reinterpret_cast<const float4*>(s.subg + row * W + 8 * ch + h4);
          const float4 q1 = *reinterpret_cast<const float4*>(s.subg + row * W + 8 * ch + (h4 ^ 4));
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
            tv = fmaf(y, g[j], -1.0f);
          acc += tv;
          tw[row][j] = tv * w[j];
        }
        if (row == 0)
          sa += acc;
        else
          sb += acc;
      }
      float4* dst = reinterpret_cast<float4*>(s.stg + 8 * ch);
      const int sw = (ch >> 1) & 3;  // stg_phys: unit j of the chunk goes to unit j ^ sw
#pragma unroll
      for (int j = 0; j < 8; j += 2) dst[(j >> 1) ^ sw] = make_float4(tw[0][j], tw[1][j], tw[0][j + 1], tw[1][j + 1]);
    }
  }
}

// General path: the rows are already (t - mean) * window (and Fourier-upsampled) f32 values written by rowprep_kernel;
// stage them pairwise-interleaved exactly like phase_pre does.  W = samples per row (m * opw), a multiple of 8.
template <class P>
ABC_HD void phase_stage_f32(int tid, const GroupSmem& s, int W, const float* rowa, const float* rowb) {
  const int W8 = W >> 3;
  if (tid == 0) s.stg[W] = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < P::NCH; ++i) {
    const int ch = tid + P::T * i;
    if (ch < W8) {
      const float4 a0 = *reinterpret_cast<const float4*>(rowa + 8 * ch), a1 = *reinterpret_cast<const float4*>(rowa + 8 * ch + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(rowb + 8 * ch), b1 = *reinterpret_cast<const float4*>(rowb + 8 * ch + 4);
      float4* dst = reinterpret_cast<float4*>(s.stg + 8 * ch);
      const int sw = (ch >> 1) & 3;
      dst[0 ^ sw] = make_float4(a0.x, b0.x, a0.y, b0.y);
      dst[1 ^ sw] = make_float4(a0.z, b0.z, a0.w, b0.w);
      dst[2 ^ sw] = make_float4(a1.x, b1.x, a1.y, b1.y);
      dst[3 ^ sw] = make_float4(a1.z, b1.z, a1.w, b1.w);
    }
  }
}

// lambda->k gather-lerp (BscanFFT.cpp:1169-1171) into registers: the inputs of this thread's radix-R0 butterflies.
// Runs between two barriers because the exchange buffer written by pass 0 overlays the staging buffer read here.
template <class P>
ABC_HD void phase_gather(int tid, const GroupSmem& s, ThreadState<P>& r, float ma, float mb) {
  const unsigned char* stgb = reinterpret_cast<const unsigned char*>(s.stg);
#pragma unroll
  for (int i = 0; i < P::NB0; ++i) {
    const int b = tid + P::T * i;
    if (P::NB0 * P::T == P::N1 || b < P::N1) {
#pragma unroll
      for (int c4 = 0; c4 < P::R0P4 / 4; ++c4) {
        const uint4 v = *reinterpret_cast<const uint4*>(s.idxT + (c4 * P::N1 + b) * 4);
        const float4 f0 = *reinterpret_cast<const float4*>(s.wqT + (c4 * P::N1 + b) * 4);
        const float4 f1 = *reinterpret_cast<const float4*>(s.vwT + (c4 * P::N1 + b) * 4);
        const unsigned off[4] = {v.x, v.y, v.z, v.w};
        const float wq[4] = {f0.x, f0.y, f0.z, f0.w};
        const float vw[4] = {f1.x, f1.y, f1.z, f1.w};  // lerp(window) at this output sample
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int a = 4 * c4 + j;
          if (a < P::R0) {
            const float2 y1 = *reinterpret_cast<const float2*>(stgb + (off[j] & 0xffffu));
            const float2 y0 = *reinterpret_cast<const float2*>(stgb + (off[j] >> 16));
            r.x[i][a].x = fmaf(-ma, vw[j], fmaf(wq[j], y1.x - y0.x, y1.x));
            r.x[i][a].y = fmaf(-mb, vw[j], fmaf(wq[j], y1.y - y0.y, y1.y));
          }
        }
      }
    }
  }
}

// pass 0: first radix-R0 butterflies on the gathered inputs
template <class P>
ABC_HD void phase_pass0(int tid, const GroupSmem& s, ThreadState<P>& r) {
#pragma unroll
  for (int i = 0; i < P::NB0; ++i) {
    const int b = tid + P::T * i;
    if (P::NB0 * P::T == P::N1 || b < P::N1) {
      float2 out[P::R0];
      Dft<P::R0, kFftSign, 1, 1>::run(r.x[i], out);
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
template <class P, bool ACCUM = true>
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
          const float m0 = fast_sqrt(fmaf(sr, sr, di * di)), m1 = fast_sqrt(fmaf(si, si, dr * dr));
          r.acc[i][j][0] = ACCUM ? r.acc[i][j][0] + m0 : m0;
          r.acc[i][j][1] = ACCUM ? r.acc[i][j][1] + m1 : m1;
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
          const float m0 = fast_sqrt(fmaf(sr, sr, di * di)), m1 = fast_sqrt(fmaf(si, si, dr * dr));
          r.acc[i][j][0] = ACCUM ? r.acc[i][j][0] + m0 : m0;
          r.acc[i][j][1] = ACCUM ? r.acc[i][j][1] + m1 : m1;
        }
      }
    }
  }
}

// average, +1e-5, ln -> dB, DC-row mask, min/max of the dB values (BscanFFT.cpp:1221-1247).
// Bins 0, 1, 4 and the clampupper element (5,5) can only sit in slot j == 0 of a unit (every other slot holds a bin
// >= S / 2 >= 8), so only that slot pays for the special cases.
template <class P>
ABC_HD void phase_finalise(int tid, const ReconArgs& a, float* rowa_out, float* rowb_out, int row_a_index,
                           bool rowb_valid, ThreadState<P>& r, float& mn, float& mx, unsigned long long keep_pol = 0,
                           float* dc = nullptr /* [2 rows][2]: dB of bins 0, 1 before the mask, on request */) {
  static_assert(P::S / 2 >= 8, "special bins must all fall into slot 0");
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
          if (j == 0) {
            if (valid && kk >= 2) {
              store_scratch(dst + kk, db, keep_pol);
              if (kk == 4) {  // bscandb.row(4).copyTo(row(1)), row(0): BscanFFT.cpp:1239-1240
                store_scratch(dst, db, keep_pol);
                store_scratch(dst + 1, db, keep_pol);
              }
              const bool is55 = a.clamp55 && kk == 5 && (row_a_index + row) == 5;
              if (!is55) {  // min/max of the raw dB; max(., thr) is monotone and is applied to the two scalars afterwards
                mn = fminf(mn, db);
                mx = fmaxf(mx, db);
              }
            } else if (valid && dc != nullptr) {  // bins 0, 1: masked in every display image, kept only on request
              dc[2 * row + kk] = db;
            }
          } else if (valid) {
            store_scratch(dst + kk, db, keep_pol);
            mn = fminf(mn, db);
            mx = fmaxf(mx, db);
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
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// The only wait on another group's progress (the drain loop) is bounded: a scheduling bug must surface as a launch failure,
// not as a hung GPU.
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
constexpr unsigned long long kWatchdogNs = 20ull * 1000 * 1000 * 1000;
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_cg4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
// drop a consumed 128-byte scratch line from L2 without writing it back to HBM
__device__ __forceinline__ void l2_discard128(const void* p) { asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory"); }

// One normalisation job: A-scans [r0, r0 + T) of B-scan b, all D bins.  Threshold, global min-max normalise,
// transpose to depth-major, quantise (BscanFFT.cpp:1243-1255); optionally the transposed dB image.  The dB scratch
// is read from L2 (written moments ago by this kernel) and discarded line by line once consumed.
// Tile = job = T A-scans x 32 bins.  Thread (q = tid / 8, c4 = tid % 8) loads bins [4 c4, 4 c4 + 4) of the rows
// 4 q' + {0..3} for q' = q + (T / 8) k: four consecutive A-scans per bin become one packed 32-bit word of the
// transposed byte tile.  The loads of tile t + 1 are in flight while tile t is quantised and stored.
template <class P>
__device__ __forceinline__ void normalise_part(const ReconArgs& a, const SchedView& sv, int b, int part, int g, int tid, unsigned char* tile) {
  constexpr int T = P::T;
  constexpr int NR = T;          // A-scans per job and per tile: one row quad pair per thread whatever the plan
  constexpr int TP = NR + 4;     // byte pitch of the u8 tile [kNormBins][TP]
  constexpr int QPT = NR / 4;    // row quads per tile
  constexpr int NK = 2;          // row quads per thread (QPT * 8 / T)
  const int r0 = part * NR;
  const int nrows = min(NR, a.oph - r0);
  float mn = ordered_to_float(__ldcg(sv.minv + b)), mx = ordered_to_float(__ldcg(sv.maxv + b));
  if (a.clamp55) {  // bscandisp.at<double>(5,5) = 50.0 before the min-max (BscanFFT.cpp:1248-1253)
    mn = fminf(mn, a.clamp_db);
    mx = fmaxf(mx, a.clamp_db);
  }
  const float range = mx - mn;
  const float sc = range > 2.220446049250313e-16f ? 255.0f / range : 0.f;  // cv::normalize: scale = 0 for a flat image
  const float thr = a.thr;
  const float* src = a.scratch + ((size_t)b * a.oph + r0) * a.Dp;
  const bool has55 = a.clamp55 && r0 <= 5 && 5 < r0 + nrows;
  const bool vec_ok = (a.oph & 15) == 0 && nrows == NR && (NR % 16) == 0 && (reinterpret_cast<uintptr_t>(a.out8) & 15) == 0;
  const int c4 = tid & 7, q0 = tid >> 3;
  const int ntiles = (a.D + kNormBins - 1) / kNormBins;

  float4 va[NK][4], vb[NK][4];
  // row pointers and validity once per job; a tile only adds 128 bytes
  const float* rowp[NK][4];
  unsigned valid = 0;
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int q = q0 + (T / 8) * k;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int row = 4 * q + rr;
      const bool ok = q < QPT && row < nrows;
      rowp[k][rr] = src + (size_t)(ok ? row : 0) * a.Dp + 4 * c4;
      valid |= ok ? (1u << (4 * k + rr)) : 0u;
    }
  }
  auto load_tile = [&](int t, float4 (&v)[NK][4]) {
#pragma unroll
    for (int k = 0; k < NK; ++k)
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
        v[k][rr] = ((valid >> (4 * k + rr)) & 1u) ? ld_cg4(rowp[k][rr] + t * kNormBins) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto quant = [&](float x) -> unsigned {
    // round-half-even of (max(x, thr) - mn) * 255 / (mx - mn) in [0, 255]: 1.5 * 2^23 trick, result in the low byte
    return __float_as_uint(fmaf(fmaxf(x, thr) - mn, sc, 12582912.0f));
  };
  auto pack4 = [&](float x0, float x1, float x2, float x3) -> unsigned {
    const unsigned lo = __byte_perm(quant(x0), quant(x1), 0x0040);  // bytes: x0, x1
    const unsigned hi = __byte_perm(quant(x2), quant(x3), 0x0040);
    return __byte_perm(lo, hi, 0x5410);
  };
  auto process_tile = [&](int t, float4 (&v)[NK][4]) {
    const int d0 = t * kNormBins;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int q = q0 + (T / 8) * k;
      if (q < QPT) {
        unsigned* tw = reinterpret_cast<unsigned*>(tile + (4 * c4) * TP + 4 * q);
        tw[0 * (TP / 4)] = pack4(v[k][0].x, v[k][1].x, v[k][2].x, v[k][3].x);
        tw[1 * (TP / 4)] = pack4(v[k][0].y, v[k][1].y, v[k][2].y, v[k][3].y);
        tw[2 * (TP / 4)] = pack4(v[k][0].z, v[k][1].z, v[k][2].z, v[k][3].z);
        tw[3 * (TP / 4)] = pack4(v[k][0].w, v[k][1].w, v[k][2].w, v[k][3].w);
      }
    }
    group_sync<T>(g);
    if (has55 && d0 == 0) {  // uniform: the forced pixel lives in this tile
      if (tid == 0) tile[5 * TP + (5 - r0)] = (unsigned char)(quant(a.clamp_db) & 0xffu);
      group_sync<T>(g);
    }
    // ---- store: per bin, T consecutive A-scans = T contiguous bytes, 16 bytes per thread
    uint8_t* const obase = a.out8 + ((size_t)b * a.D + d0) * a.oph + r0;
#pragma unroll
    for (int i = tid; i < kNormBins * (NR / 16); i += T) {
      const int dd = i / (NR / 16), w16 = i % (NR / 16);
      const int d = d0 + dd;
      if (d < a.D) {
        uint8_t* o = obase + (size_t)dd * a.oph + 16 * w16;
        const unsigned char* tp = tile + dd * TP + 16 * w16;
        if (vec_ok) {
          const unsigned* t32 = reinterpret_cast<const unsigned*>(tp);
          __stcs(reinterpret_cast<uint4*>(o), make_uint4(t32[0], t32[1], t32[2], t32[3]));
        } else {
          for (int k = 0; k < 16 && 16 * w16 + k < nrows; ++k) o[k] = tp[k];
        }
      }
    }
    if (a.outdb != nullptr) {  // transposed dB image (rarely requested): straight from the registers, 16-byte row quads
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const int q = q0 + (T / 8) * k;
        if (q < QPT) {
          const float col[4][4] = {{v[k][0].x, v[k][1].x, v[k][2].x, v[k][3].x}, {v[k][0].y, v[k][1].y, v[k][2].y, v[k][3].y},
                                   {v[k][0].z, v[k][1].z, v[k][2].z, v[k][3].z}, {v[k][0].w, v[k][1].w, v[k][2].w, v[k][3].w}};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int d = d0 + 4 * c4 + j;
            if (d < a.D) {
              float* o = a.outdb + ((size_t)b * a.D + d) * a.oph + r0 + 4 * q;
#pragma unroll
              for (int rr = 0; rr < 4; ++rr)
                if (4 * q + rr < nrows) o[rr] = col[j][rr];
            }
          }
        }
      }
    }
    group_sync<T>(g);
    // ---- the lines of this tile are dead: drop them from L2 so that they are never written back
    for (int row = tid; row < nrows; row += T) l2_discard128(src + (size_t)row * a.Dp + d0);
  };

  load_tile(0, va);
  for (int t = 0; t < ntiles; t += 2) {
    if (t + 1 < ntiles) load_tile(t + 1, vb);
    process_tile(t, va);
    if (t + 1 < ntiles) {
      if (t + 2 < ntiles) load_tile(t + 2, va);
      process_tile(t + 1, vb);
    }
  }
}

// G thread groups per CTA (compile time: the register budget follows from T * G), A1 = averages == 1,
// IN_F32 = the frames are pre-processed f32 rows of the general path (a.W samples each, no calibration)
template <class P, int G, bool HAS_SUB, bool A1, bool IN_F32>
__global__ void __launch_bounds__(P::T* G, 1) recon_kernel(const ReconArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout L = make_layout<P>(a.W, HAS_SUB);
  const int g = threadIdx.x / P::T;
  const int tid = threadIdx.x - g * P::T;
  const int lane = tid & 31, wrp = tid >> 5;
  const GroupSmem s = resolve<P>(smem, L, g);
  const int W = a.W, W8 = W >> 3;
  const SchedView sv = sched_view(a.sched, a.nB);

  // ---- CTA-wide tables: global (L2-resident) -> shared, once per persistent CTA
  {
    const int n16 = L.groups >> 4;  // tables are laid out contiguously in the same order in global memory
    const uint4* src = reinterpret_cast<const uint4*>(a.idxT);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
  }
  if (tid == 0) {
    mbar_init(s.mbar, 1);
    mbar_init(s.mbar + 1, P::NWARPS);  // "exchange buffer free": one arrival per warp after its last-pass reads
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // dynamic schedule: every group claims items (one row pair of one B-scan) from a global ticket, two ahead;
    // the leader decodes ticket -> (pair, bscan) once so that nobody else divides
    const int t0 = atomicAdd(sv.ticket, 1), t1 = atomicAdd(sv.ticket, 1);
    s.slot[0] = t0 < a.nitems ? t0 % a.npairs : -1;
    s.slot[1] = t0 / a.npairs;
    s.slot[2] = t1 < a.nitems ? t1 % a.npairs : -1;
    s.slot[3] = t1 / a.npairs;
  }
  __syncthreads();

  unsigned long long pol, keep_pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep_pol));

  ThreadState<P> r;
#pragma unroll
  for (int i = 0; i < P::NU; ++i)
#pragma unroll
    for (int j = 0; j < P::RL; ++j) r.acc[i][j][0] = r.acc[i][j][1] = 0.f;

  auto issue_calibration = [&](int pair) {  // TMA bulk copies of the item's two calibration rows (tid 0 only)
    const int ra = 2 * pair, rb = (ra + 1 < a.oph) ? ra + 1 : ra;
    const unsigned rowbytes = static_cast<unsigned>(W) * 4u;
    mbar_expect_tx(s.mbar, rowbytes * (HAS_SUB ? 4u : 2u));
    bulk_g2s(s.gain, a.gain + static_cast<size_t>(ra) * W, rowbytes, s.mbar);
    bulk_g2s(s.gain + W, a.gain + static_cast<size_t>(rb) * W, rowbytes, s.mbar);
    if constexpr (HAS_SUB) {
      bulk_g2s(s.subg, a.subg + static_cast<size_t>(ra) * W, rowbytes, s.mbar);
      bulk_g2s(s.subg + W, a.subg + static_cast<size_t>(rb) * W, rowbytes, s.mbar);
    }
  };
  auto prefetch_rows = [&](int pair, int b, int f) {  // raw pixels of frame f of item (pair, b) into registers
    if constexpr (IN_F32) return;
    const int ra = 2 * pair, rb = (ra + 1 < a.oph) ? ra + 1 : ra;
    const uint8_t* fp = a.frames + (static_cast<size_t>(b) * a.A + f) * a.frame_stride;
    phase_load<P>(tid, fp + static_cast<size_t>(ra) * a.row_stride, fp + static_cast<size_t>(rb) * a.row_stride, W8, r, pol);
  };

  int pair = s.slot[0], bscan = s.slot[1], npair = s.slot[2], nbscan = s.slot[3];  // current and next item (pair < 0: none)
  if (pair >= 0) {
    if (!IN_F32 && tid == 0) issue_calibration(pair);
    prefetch_rows(pair, bscan, 0);
  }
  unsigned cal_parity = 0, free_parity = 0;
  bool buf_busy = false;  // a previous frame's last pass may still be reading the exchange / staging buffer

  // Normalisation jobs (B-scan b, block of T A-scans) are assigned statically: group `gid` owns jobs
  // gid, gid + ngroups, ...  A job may start once all npairs row pairs of its B-scan have been counted in sv.cnt.
  // Every global round trip of the group leader (ticket, completion poll, completion publish) is issued early and
  // consumed late so that no warp waits for L2 inside an item.
  const int ngroups = gridDim.x * G;
  const int njobs = a.nB * a.nparts;
  // The group's housekeeping is spread over its warps so that no single warp reaches the barriers late:
  constexpr int kTicketTid = 32 % P::T;   // claims and decodes the item tickets
  constexpr int kPublishTid = 64 % P::T;  // publishes finished pairs (fence + RED)
  constexpr int kJobTid = 96 % P::T;      // polls the completion counters and hands out this group's normalisation jobs
  int myjob = blockIdx.x * G + g;  // (kJobTid) next normalisation job of this group
  int myjob_b = myjob / a.nparts;  //           and its B-scan
  // (tid 0) B-scans whose finished pair has not been published yet.  Publishing needs a gpu-scope fence, which waits
  // for every memory operation the warp has in flight; it is therefore batched (one fence per kPublishBatch items)
  // and placed right after the pre-processing phase, when the previous items' scratch stores have long landed and
  // the next frame's pixel prefetch has not been issued yet.
  constexpr int kPublishBatch = ABC_PUBLISH_BATCH;
  int* const pend = s.slot + 8;  // kept in shared memory: only kPublishTid touches it
  int npend = 0;
  auto publish = [&]() {
    __threadfence();
    for (int i = 0; i < npend; ++i) atomicAdd(sv.cnt + pend[i], 1);  // result unused -> RED
    npend = 0;
  };

  while (pair >= 0) {
    const int ra = 2 * pair;
    const bool rowb_valid = (ra + 1) < a.oph;
    int t_next = 0, polled = 0;
    if (tid == kTicketTid) t_next = atomicAdd(sv.ticket, 1);  // the item after next
    if (tid == kJobTid && myjob < njobs) polled = *reinterpret_cast<volatile const int*>(sv.cnt + myjob_b);
    if constexpr (!IN_F32) {
      while (!mbar_try_wait(s.mbar, cal_parity)) {
      }
      cal_parity ^= 1u;
    }

    const int nA = A1 ? 1 : a.A;
    for (int f = 0; f < nA; ++f) {
      const bool last = A1 || (f + 1 == nA);
      // Split barrier: the staging buffer written next overlays the exchange buffer that slower warps may still be
      // reading in the previous frame's last pass.  Every warp signalled "done reading" right after that pass
      // (mbarrier arrive); the dB conversion, stores and scheduling in between hide the wait.
      if constexpr (P::NWARPS > 1) {
        if (buf_busy) {
          while (!mbar_try_wait(s.mbar + 1, free_parity)) {
          }
          free_parity ^= 1u;
        }
      }
      float sa = 0.f, sb = 0.f;
      if constexpr (IN_F32) {
        const int rb = rowb_valid ? ra + 1 : ra;
        const float* fp = reinterpret_cast<const float*>(a.frames) + (static_cast<size_t>(bscan) * a.A + f) * a.oph * static_cast<size_t>(W);
        phase_stage_f32<P>(tid, s, W, fp + static_cast<size_t>(ra) * W, fp + static_cast<size_t>(rb) * W);
      } else {
        phase_pre<P, HAS_SUB>(tid, s, W, r, sa, sb);
        sa = warp_sum(sa);
        sb = warp_sum(sb);
        if constexpr (P::NWARPS > 1) {
          if (lane == 0) {
            s.red[2 * wrp] = sa;
            s.red[2 * wrp + 1] = sb;
          }
        }
      }
      // prefetch while this frame is transformed: next frame of the item, else the first frame of the next item
      if (!last) {
        prefetch_rows(pair, bscan, f + 1);
      } else if (npair >= 0) {
        prefetch_rows(npair, nbscan, 0);
      }
      group_sync<P::T>(g);  // staged samples and row sums visible; the calibration rows have been consumed
      // Publish AFTER the barrier: it orders the other warps' scratch stores and min / max atomics of the finished items before
      // the publisher's gpu-scope fence (a count published before it could be seen ahead of the data it covers).
      if (f == 0 && tid == kPublishTid && npend == kPublishBatch) publish();
      if (!IN_F32 && last && npair >= 0 && tid == 0) issue_calibration(npair);
      if constexpr (P::NWARPS > 1 && !IN_F32) {
        sa = 0.f;
        sb = 0.f;
#pragma unroll
        for (int w = 0; w < P::NWARPS; ++w) {
          sa += s.red[2 * w];
          sb += s.red[2 * w + 1];
        }
      }
      phase_gather<P>(tid, s, r, sa * a.inv_W, sb * a.inv_W);
      group_sync<P::T>(g);  // the staging buffer has been consumed: pass 0 may overwrite it
      phase_pass0<P>(tid, s, r);
      if (last && tid == kTicketTid) {
        const int nb = t_next / a.npairs;
        s.slot[4] = t_next < a.nitems ? t_next - nb * a.npairs : -1;
        s.slot[5] = nb;
      }
      if (last && tid == kJobTid) {
        int job = -1;
        if (myjob < njobs && polled >= a.npairs) {
          __threadfence();  // acquire side of the completion count
          job = myjob;
          myjob += ngroups;
          myjob_b = myjob / a.nparts;
        }
        s.slot[6] = job;
      }
      group_sync<P::T>(g);
      if constexpr (P::THREE) {
        phase_pass1<P>(tid, s);
        group_sync<P::T>(g);
      }
      phase_passL<P, !A1>(tid, s, r);
      if constexpr (P::NWARPS > 1) {
        __syncwarp();
        if (lane == 0) mbar_arrive(s.mbar + 1);
        buf_busy = true;
      }
    }

    // ---- B-scan `bscan` of this pair is complete: dB to the L2 scratch, min/max
    {
      float* oa = a.scratch + (static_cast<size_t>(bscan) * a.oph + ra) * a.Dp;
      float* ob = oa + a.Dp;
      float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
      float* dc = a.dc01 != nullptr ? a.dc01 + 2 * (static_cast<size_t>(bscan) * a.oph + ra) : nullptr;
      phase_finalise<P>(tid, a, oa, ob, ra, rowb_valid, r, mn, mx, keep_pol, dc);
      mn = warp_min(mn);
      mx = warp_max(mx);
      if (lane == 0 && mn <= mx) {  // thresholded min/max (BscanFFT.cpp:1247): max(., thr) commutes with min/max
        atomicMin(sv.minv + bscan, float_to_ordered(fmaxf(mn, a.thr)));
        atomicMax(sv.maxv + bscan, float_to_ordered(fmaxf(mx, a.thr)));
      }
      if (tid == kPublishTid) pend[npend++] = bscan;
      const int job = s.slot[6];
      if (job >= 0) {
        group_sync<P::T>(g);  // every thread is done with the exchange buffer (it becomes the transposition tile)
        normalise_part<P>(a, sv, job / a.nparts, job % a.nparts, g, tid, reinterpret_cast<unsigned char*>(s.buf));
      }
    }
    pair = npair;
    bscan = nbscan;
    npair = s.slot[4];
    nbscan = s.slot[5];
  }

  // ---- drain: publish the last item, then finish this group's remaining normalisation jobs
  group_sync<P::T>(g);
  if (tid == kPublishTid && npend > 0) publish();
  for (;;) {
    if (tid == kJobTid) {
      int job = -1;
      if (myjob < njobs) {
        const unsigned long long t0 = global_ns();
        while (ld_acquire(sv.cnt + myjob / a.nparts) < a.npairs) {
          __nanosleep(200);
          if (global_ns() - t0 > kWatchdogNs) __trap();
        }
        job = myjob;
        myjob += ngroups;
      }
      s.slot[6] = job;
    }
    group_sync<P::T>(g);
    const int job = s.slot[6];
    if (job < 0) break;
    normalise_part<P>(a, sv, job / a.nparts, job % a.nparts, g, tid, reinterpret_cast<unsigned char*>(s.buf));
    group_sync<P::T>(g);
  }
}
#endif  // __CUDACC__

}  // namespace abcoct
