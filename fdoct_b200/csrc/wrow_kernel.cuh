// Warp-per-A-scan fused reconstruction kernel (sm_100a): the whole block BscanFFT.cpp:987-1255 in one launch, one WARP per
// camera row, no block-level barrier anywhere on the path.
//
// Why this shape (round-2 analysis): the round-1 kernel (recon_kernel.cuh: 128 threads per packed row
// pair, three FFT passes, two shared-memory exchanges, group barriers) was bound by the SM's L1 / shared-memory data pipe
// (70 % busy at 58 % issue utilisation, 1130 wavefronts per A-scan) and by barrier / short-scoreboard stalls.  Here
//   * a real row of N samples is ONE N/2-point complex transform (z[j] = x[2j] + i x[2j+1]) plus a split step, so a warp holds
//     the whole transform in registers (N/64 complex values per lane) and needs TWO in-register radix passes with ONE exchange
//     through shared memory (__syncwarp only); every row has its own f32 noise floor (no two-for-one packing of different rows);
//   * the reference's resampling weight is indexed by the SOURCE sample (BscanFFT.cpp:1170: fractionalk.at(nearestkindex)), so
//     y[i] + w[i] (y[i] - y[i-1]) depends on i only: it is evaluated once per source sample before staging and the
//     lambda -> k resampling itself becomes a pure 4-byte gather with one offset table (no weights, no second tap);
//   * the row mean (BscanFFT.cpp:1135-1139) is a warp-shuffle reduction over samples that never leave the registers;
//   * calibration rows are read straight from L2 into registers (no shared-memory staging), raw pixel rows are pulled into L2
//     ahead of time by the TMA unit (cp.async.bulk.prefetch.L2), re / im pairs are processed with the packed f32x2
//     instructions (fft_regs.cuh).
// Per A-scan this is ~2.0 k warp instructions and ~0.75 k data-pipe wavefronts against 3.7 k / 1.13 k before.
//
// Data flow of one row (W samples, N-point transform, R = N / 64 values per lane in pass A):
//   pre   : 8-sample runs (run = lane + 32 j): u16 -> f32, t = y * gain - subg (one FFMA; BscanFFT.cpp:987, 1132,
//           BscanDark.cpp:1269), warp sum -> mean, s = t - mean, v[i] = P[i] s[i] - Q[i] s[i-1] with P = (1 + F) win[i],
//           Q = F win[i-1] (window BscanFFT.cpp:1141 and resampling weight :1169-1171 folded), staged as two planes (even /
//           odd samples) so that the stride-2 gather below is bank-conflict free;
//   pass A: lane b gathers x[q] = v[nearestkindex[q]], q = 2 (b + 32 a) + {0, 1}, a < R, radix-R DFT over a, twiddle w^(b c);
//   xchg  : [c / 2][b][c % 2] float2, 16-byte stores, 8-byte loads, row pitch 33 * 16 bytes;
//   pass B: lane c < R: radix-32 DFT over b -> Z[c + R d], d < 32;
//   split : X[k] = (A + B) / 2, X[N/2 - k] = conj(A - B) / 2 with A = Z[k] + conj Z[N/2 - k], B = -i w_N^k (Z[k] - conj Z[N/2 - k]);
//           the partner values come from lane R - c by warp shuffle; magnitudes (BscanFFT.cpp:1189-1190), accumulation over
//           `averages` frames in registers (:1193-1209);
//   final : /A, + 1e-5, ln -> dB (2.303), DC-row mask (:1221-1240) -> f32 row in the L2 scratch, min / max by CREDUX;
//   norm  : when a B-scan is complete its jobs (32 A-scans x a range of 32-bin tiles) are picked up by the warps: threshold,
//           global min-max normalise, round-half-even to u8 (:1243-1255), transposed 32-byte-sector stores, scratch lines dropped.
#pragma once
#include <cstdint>
#include <cstring>

#include "fft_regs.cuh"
#include "plan.h"
#include "recon_kernel.cuh"
#include "wrow_prims.cuh"

namespace abcoct {

// LM_ (how a row's pixels and gain values reach the registers):
//   0  16-byte global loads at the start of the row (rows pulled into L2 ahead of time by the TMA unit's bulk prefetch)
//   1  TMA bulk copies of both rows into per-warp shared memory one row ahead (mbarrier), conflict-free 16-byte LDS
//   2  TMA bulk copy of the gain row into the (idle) exchange buffer, pixels by global loads issued before the split step
template <int N_, int NW_, int LM_ = 0>
struct WPlan {
  static constexpr int N = N_, N2 = N_ / 2, R = N_ / 64, NW = NW_, LM = LM_;
  static_assert(N_ % 128 == 0 && R <= 32 && R >= 8, "N must be 128 * even, 512 <= N <= 2048");
  static constexpr int NCH = (N / 8 + 31) / 32;  // 8-sample runs per lane
  static constexpr int WMAX = NCH * 256;         // padded row length (calibration pitch, staging planes)
  static constexpr int PO = WMAX / 2 + 4;        // word offset of the odd-sample plane; words [WMAX/2, WMAX/2 + 4) stay zero
  static constexpr int STAGE_BYTES = (PO + WMAX / 2) * 4;
  static constexpr int XPITCH = 33 * 16;  // exchange rows: 32 lanes x 16 bytes + 16 bytes of padding (conflict-free 8-byte reads)
  static constexpr int XCH_BYTES = (R / 2) * XPITCH;
  static constexpr int WBUF = ((cmax(STAGE_BYTES, XCH_BYTES) + 15) / 16) * 16;
  // table blob = shared-memory image (bytes)
  static constexpr int T_OFFS = 0;                          // uint4 [R/2][32]: byte offsets of the 4 samples of a = 2 a2, 2 a2 + 1
  static constexpr int T_PQ = T_OFFS + (R / 2) * 32 * 16;   // float4 [NCH][4][32]: P[0..3], P[4..7], Q[0..3], Q[4..7] of a run
  static constexpr int T_TWA = T_PQ + NCH * 4 * 32 * 16;    // float4 [R/2][32]: w^(b 2p), w^(b (2p+1))
  static constexpr int T_TWP = T_TWA + (R / 2) * 32 * 16;   // float4 [8][32]: T[c + R 2 d2], T[c + R (2 d2 + 1)]
  static constexpr int TABLE_BYTES = T_TWP + 8 * 32 * 16;
  // per warp: WBUF (gain row by TMA -> staged samples -> exchange rows, in turn), the raw pixel row (TMA) and its mbarrier
  static constexpr int RAWBUF = LM == 1 ? WMAX * 2 : 0;
  static constexpr int WSTRIDE = WBUF + RAWBUF + 128;  // + mbarrier, lane 0's bookkeeping words and the mailbox to the service warp
  static constexpr int NWK = NW - 1;  // worker warps; the last warp of the CTA is the service warp
  static constexpr int SMEM_BYTES = TABLE_BYTES + NW * WSTRIDE;
  static_assert(WMAX * 4 <= WBUF, "the calibration row must fit the warp buffer");
  static_assert(SMEM_BYTES <= 227 * 1024, "too many warps per CTA for the shared memory");
  static constexpr int ZERO_OFF = (WMAX / 2) * 4;  // byte offset of the zero sentinel inside a warp buffer
  static constexpr int MAXREG = cmax(32, ((65536 / (NW * 32)) / 8) * 8 > 255 ? 255 : ((65536 / (NW * 32)) / 8) * 8);  // one CTA per SM
};

// byte offset of staged sample i inside a warp buffer (two planes: even samples, then odd samples)
template <class WP>
__host__ __device__ constexpr int wrow_stage_off(int i) {
  return 4 * ((i & 1) ? WP::PO + (i >> 1) : (i >> 1));
}

struct WrowTablesHost {  // inputs of the blob builder (host)
  int W;
  const int* idx;     // N gather indices: source sample of output q, or -1 for "never written" (q = 0, N - 1: zero)
  const double* frac;  // fractionalk (BscanFFT.cpp:692-698), indexed by SOURCE sample (the reference's quirk), >= W entries
  const double* win;   // W window values
};

template <class WP>
inline void wrow_build_blob(const WrowTablesHost& t, unsigned char* blob /* WP::TABLE_BYTES, zeroed by the caller */) {
  constexpr int R = WP::R;
  uint32_t* offs = reinterpret_cast<uint32_t*>(blob + WP::T_OFFS);
  float* pq = reinterpret_cast<float*>(blob + WP::T_PQ);
  float* twa = reinterpret_cast<float*>(blob + WP::T_TWA);
  float* twp = reinterpret_cast<float*>(blob + WP::T_TWP);
  auto off_of = [&](int q) -> uint32_t {
    const int i = t.idx[q];
    return i < 0 ? (uint32_t)WP::ZERO_OFF : (uint32_t)wrow_stage_off<WP>(i);
  };
  for (int a2 = 0; a2 < R / 2; ++a2)
    for (int b = 0; b < 32; ++b) {
      uint32_t* o = offs + (a2 * 32 + b) * 4;
      const int q0 = 2 * (b + 32 * (2 * a2)), q1 = 2 * (b + 32 * (2 * a2 + 1));
      o[0] = off_of(q0);
      o[1] = off_of(q0 + 1);
      o[2] = off_of(q1);
      o[3] = off_of(q1 + 1);
    }
  for (int j = 0; j < WP::NCH; ++j)
    for (int lane = 0; lane < 32; ++lane)
      for (int e = 0; e < 8; ++e) {
        const int i = 8 * (lane + 32 * j) + e;
        double P = 0.0, Q = 0.0;
        if (i >= 1 && i < t.W) {  // v[i] = y[i] + F[i] (y[i] - y[i-1]) on the apodised samples y = s * win
          P = (1.0 + t.frac[i]) * t.win[i];
          Q = t.frac[i] * t.win[i - 1];
        }
        pq[((j * 4 + (e >> 2)) * 32 + lane) * 4 + (e & 3)] = (float)P;
        pq[((j * 4 + 2 + (e >> 2)) * 32 + lane) * 4 + (e & 3)] = (float)Q;
      }
  const double tau = 6.283185307179586476925286766559;
  for (int p = 0; p < R / 2; ++p)
    for (int b = 0; b < 32; ++b)
      for (int h = 0; h < 2; ++h) {
        const long long num = ((long long)b * (2 * p + h)) % WP::N2;
        const double ang = tau * (double)num / (double)WP::N2;
        twa[(p * 32 + b) * 4 + 2 * h] = (float)std::cos(ang);
        twa[(p * 32 + b) * 4 + 2 * h + 1] = (float)(kFftSign * std::sin(ang));
      }
  for (int d2 = 0; d2 < 8; ++d2)
    for (int c = 0; c < R; ++c)
      for (int h = 0; h < 2; ++h) {
        const int k = c + R * (2 * d2 + h);
        const double ang = tau * (double)k / (double)WP::N;
        // T_k = -i w_N^k (kFftSign = +1): sin - i cos;  for the forward sign it would be +i w^-k
        twp[(d2 * 32 + c) * 4 + 2 * h] = (float)std::sin(ang);
        twp[(d2 * 32 + c) * 4 + 2 * h + 1] = (float)(-kFftSign * std::cos(ang));
      }
}

// calibration rows as the kernel reads them: [row][j][h][lane][4] floats, pitch WP::WMAX (zeros beyond W)
template <class WP>
inline void wrow_permute_cal_row(const float* in, int W, float* out /* WP::WMAX */) {
  for (int j = 0; j < WP::NCH; ++j)
    for (int h = 0; h < 2; ++h)
      for (int lane = 0; lane < 32; ++lane)
        for (int e = 0; e < 4; ++e) {
          const int i = 8 * (lane + 32 * j) + 4 * h + e;
          out[((2 * j + h) * 32 + lane) * 4 + e] = i < W ? in[i] : 0.f;
        }
}

// ------------------------------------------------------------------------------------------------- normalisation job
// One job: A-scans [32 part, 32 part + 32) of B-scan b, tiles [t0, t1) of 32 depth bins.  Lane (q = lane / 4, cg = lane % 4)
// loads the rows 4 q .. 4 q + 3 at the bin quads cg and cg + 4 of the tile (16-byte L2 loads, 64 contiguous bytes per row and
// instruction) and packs the four rows of a bin into one 32-bit word, so that the eight lanes of a bin quad write one full
// 32-byte sector of the depth-major display image per bin - no shared-memory transposition.
struct WNormArgs {  // by value: the job runs as a real (non-inlined) device function, one copy in the instruction cache
  const float* scratch;
  uint8_t* out8;
  float* outdb;
  const int* minv;
  const int* maxv;
  int oph, D, Dp, nparts, nsplit, clamp55;
  float thr, clamp_db;
  int ring;  // ABC_WROW_RING: B-scans the scratch holds (B-scan b lives in slot b mod ring)
};
#ifdef ABC_WROW_HOST_EMU
#define WROW_NOINLINE static __host__ __device__
#else
#define WROW_NOINLINE static __device__ __noinline__
#endif
template <int DEPTH>
WROW_NOINLINE void wrow_normalise(const WNormArgs a, int job, int lane) {
  const int per_b = a.nparts * a.nsplit;
  const int b = job / per_b;
  const int rem = job - b * per_b;
  const int part = rem / a.nsplit, split = rem - part * a.nsplit;
  const int r0 = part * 32;
  const int nrows = (a.oph - r0) < 32 ? (a.oph - r0) : 32;
  const int ntiles = (a.D + 31) >> 5;
  const int tps = (ntiles + a.nsplit - 1) / a.nsplit;
  const int t0 = split * tps;
  const int t1 = (t0 + tps) < ntiles ? (t0 + tps) : ntiles;
  if (t0 >= t1) return;

  const float thr = a.thr;
  const int q = lane >> 2, cg = lane & 3;
#ifdef ABC_WROW_RING
  const float* src = a.scratch + ((size_t)(b % a.ring) * a.oph + r0) * a.Dp;
#else
  const float* src = a.scratch + ((size_t)b * a.oph + r0) * a.Dp;
#endif
  unsigned valid = 0;
  const float* rowp[4];
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int row = 4 * q + rr;
    const bool ok = row < nrows;
    rowp[rr] = src + (size_t)(ok ? row : 0) * a.Dp + 4 * cg + 32 * t0;
    valid |= ok ? (1u << rr) : 0u;
  }
  const bool has55 = a.clamp55 && part == 0 && nrows > 5;
  const bool word_ok = (a.oph & 3) == 0 && valid == 0xfu && (reinterpret_cast<uintptr_t>(a.out8) & 3) == 0;
  const bool db_vec_ok = a.outdb != nullptr && (a.oph & 3) == 0 && valid == 0xfu && (reinterpret_cast<uintptr_t>(a.outdb) & 15) == 0;
  // output pointers of this lane's first bin (tile t0, k = 0, j = 0); they advance by whole bins (oph elements)
  const size_t o_first = ((size_t)b * a.D + 32 * t0 + 4 * cg) * a.oph + r0 + 4 * q;
  uint8_t* o8 = a.out8 + o_first;
  float* odb = a.outdb != nullptr ? a.outdb + o_first : nullptr;
  const size_t oph = (size_t)a.oph;

  float mn = 0.f, mx = 0.f, sc = 0.f;  // loaded after the first tiles are in flight
  auto load_tile = [&](int i, float4 (&v)[2][4]) {  // tile t0 + i
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
        v[k][rr] = ((valid >> rr) & 1u) ? w_ld_cg16(rowp[rr] + 32 * i + 16 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto quant = [&](float x) -> unsigned {
    // round-half-even of (max(x, thr) - mn) * 255 / (mx - mn) in [0, 255]: 1.5 * 2^23 trick, result in the low byte
    float r = fmaf(fmaxf(x, thr) - mn, sc, 12582912.0f);
    unsigned u;
    memcpy(&u, &r, 4);
    return u;
  };
  auto pack4 = [&](float x0, float x1, float x2, float x3) -> unsigned {
    const unsigned lo = w_byte_perm(quant(x0), quant(x1), 0x0040);  // bytes: x0, x1
    const unsigned hi = w_byte_perm(quant(x2), quant(x3), 0x0040);
    return w_byte_perm(lo, hi, 0x5410);
  };
  auto process_tile = [&](int i, float4 (&v)[2][4]) {
    const int t = t0 + i;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int bin0 = 32 * t + 16 * k + 4 * cg;
      const float col[4][4] = {{v[k][0].x, v[k][1].x, v[k][2].x, v[k][3].x}, {v[k][0].y, v[k][1].y, v[k][2].y, v[k][3].y},
                               {v[k][0].z, v[k][1].z, v[k][2].z, v[k][3].z}, {v[k][0].w, v[k][1].w, v[k][2].w, v[k][3].w}};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int bin = bin0 + j;
        if (bin < a.D) {
          unsigned word = pack4(col[j][0], col[j][1], col[j][2], col[j][3]);
          if (has55 && bin == 5 && q == 1)  // element (5,5): row 5 = byte 1 of the quad 4..7
            word = (word & 0xffff00ffu) | ((quant(a.clamp_db) & 0xffu) << 8);
          uint8_t* o = o8 + (size_t)(32 * i + 16 * k + j) * oph;
          if (word_ok) {
            w_st_stream_u32(o, word);
          } else {
#pragma unroll
            for (int rr = 0; rr < 4; ++rr)
              if ((valid >> rr) & 1u) w_st_global_u8(o + rr, (word >> (8 * rr)) & 0xffu);
          }
          if (odb != nullptr) {  // transposed dB image (rarely requested)
            float* od = odb + (size_t)(32 * i + 16 * k + j) * oph;
            if (db_vec_ok) {
              w_st_stream_f4(od, make_float4(col[j][0], col[j][1], col[j][2], col[j][3]));
            } else {
#pragma unroll
              for (int rr = 0; rr < 4; ++rr)
                if ((valid >> rr) & 1u) w_st_keep(od + rr, col[j][rr]);
            }
          }
        }
      }
    }
    w_syncwarp();  // every lane has consumed its loads of this tile: the lines are dead, drop them from L2 without a write-back
    if (lane < nrows) w_discard128(src + (size_t)lane * a.Dp + 32 * t);
  };

  // DEPTH (2 or 3) tiles in flight: nothing else is live in this function, the registers are free for load latency
  const int nt = t1 - t0;
  float4 va[2][4], vb[2][4], vc[DEPTH == 3 ? 2 : 1][4];
  load_tile(0, va);
  if (DEPTH == 3 && 1 < nt) load_tile(1, vb);
  mn = ordered_to_float(w_ld_cg_i(a.minv + b));
  mx = ordered_to_float(w_ld_cg_i(a.maxv + b));
  if (a.clamp55) {  // bscandisp.at<double>(5,5) = 50.0 before the min-max (BscanFFT.cpp:1248-1253)
    mn = fminf(mn, a.clamp_db);
    mx = fmaxf(mx, a.clamp_db);
  }
  sc = (mx - mn) > 2.220446049250313e-16f ? 255.0f / (mx - mn) : 0.f;  // cv::normalize: scale = 0 for a flat image

  // Fast path (whole 32-row parts, whole tiles, aligned display image, no dB image): straight-line code per tile - packed
  // arithmetic on the (x, y) / (z, w) halves of every 16-byte load, one 32-bit streaming store per bin, no per-word branches.
  const bool fast = word_ok && odb == nullptr && 32 * t1 <= a.D;
  const float2 mn2 = make_float2(mn, mn), sc2 = make_float2(sc, sc), magic2 = make_float2(12582912.0f, 12582912.0f);
  const float thrm = thr - mn;  // max(x, thr) - mn == max(x - mn, thr - mn)
  auto fast_tile = [&](int i, float4 (&v)[2][4]) {
    uint8_t* o = o8 + (size_t)(32 * i) * oph;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      unsigned qv[4][4];  // [row][bin] quantised values in the low byte
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        float2 lo = pk_sub(make_float2(v[k][rr].x, v[k][rr].y), mn2), hi = pk_sub(make_float2(v[k][rr].z, v[k][rr].w), mn2);
        lo = pk_fma(make_float2(fmaxf(lo.x, thrm), fmaxf(lo.y, thrm)), sc2, magic2);
        hi = pk_fma(make_float2(fmaxf(hi.x, thrm), fmaxf(hi.y, thrm)), sc2, magic2);
        memcpy(&qv[rr][0], &lo.x, 4);
        memcpy(&qv[rr][1], &lo.y, 4);
        memcpy(&qv[rr][2], &hi.x, 4);
        memcpy(&qv[rr][3], &hi.y, 4);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned w01 = w_byte_perm(qv[0][j], qv[1][j], 0x0040), w23 = w_byte_perm(qv[2][j], qv[3][j], 0x0040);
        w_st_stream_u32(o + (size_t)(16 * k + j) * oph, w_byte_perm(w01, w23, 0x5410));
      }
    }
    if (has55 && t0 + i == 0 && q == 1 && cg == 1)  // element (5,5), written after (and by the thread of) the word that holds it
      w_st_global_u8(a.out8 + ((size_t)b * a.D + 5) * oph + 5, quant(a.clamp_db) & 0xffu);
    w_syncwarp();  // every lane has consumed its loads of this tile: the lines are dead, drop them from L2 without a write-back
    w_discard128(src + (size_t)lane * a.Dp + 32 * (t0 + i));
  };
  auto tile = [&](int i, float4 (&v)[2][4]) {
    if (fast)
      fast_tile(i, v);
    else
      process_tile(i, v);
  };
  if constexpr (DEPTH == 3) {
    for (int i = 0; i < nt; i += 3) {
      if (i + 2 < nt) load_tile(i + 2, vc);
      tile(i, va);
      if (i + 1 < nt) {
        if (i + 3 < nt) load_tile(i + 3, va);
        tile(i + 1, vb);
      }
      if (i + 2 < nt) {
        if (i + 4 < nt) load_tile(i + 4, vb);
        tile(i + 2, vc);
      }
    }
  } else {
    (void)vc;
    for (int i = 0; i < nt; i += 2) {
      if (i + 1 < nt) load_tile(i + 1, vb);
      tile(i, va);
      if (i + 1 < nt) {
        if (i + 2 < nt) load_tile(i + 2, va);
        tile(i + 1, vb);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------- the kernel body
constexpr unsigned long long kWrowWatchdogNs = 20ull * 1000 * 1000 * 1000;

template <class WP, bool HAS_SUB, bool A1, bool FULLD>
WROW_HD void wrow_body(const ReconArgs& a, unsigned char* smem) {
  constexpr int R = WP::R, NCH = WP::NCH, N2 = WP::N2;
  const int lane = w_lane(), warp = w_warp_in_cta();
  {  // tables: global (L2-resident) -> shared, once per persistent CTA
    const uint4* src = reinterpret_cast<const uint4*>(a.idxT);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = warp * 32 + lane; i < WP::TABLE_BYTES / 16; i += WP::NW * 32) dst[i] = src[i];
  }
  if (lane == 0) w_mbar_init(reinterpret_cast<unsigned long long*>(smem + WP::TABLE_BYTES + warp * WP::WSTRIDE + WP::WBUF + WP::RAWBUF));
  w_syncthreads();
  // every table is read as tbl + compile-time offset: one address register for all of them
  const unsigned char* const tbl = smem + 16 * lane;
  auto t_offs = [&](int i) { return *reinterpret_cast<const uint4*>(tbl + WP::T_OFFS + 512 * i); };
  auto t_pq = [&](int i) { return *reinterpret_cast<const float4*>(tbl + WP::T_PQ + 512 * i); };
  auto t_twa = [&](int i) { return *reinterpret_cast<const float4*>(tbl + WP::T_TWA + 512 * i); };
  auto t_twp = [&](int i) { return *reinterpret_cast<const float4*>(tbl + WP::T_TWP + 512 * i); };
  unsigned char* const wbuf = smem + WP::TABLE_BYTES + warp * WP::WSTRIDE;
  unsigned char* const rawbuf = wbuf + WP::WBUF;
  unsigned long long* const mbar = reinterpret_cast<unsigned long long*>(rawbuf + WP::RAWBUF);
  // scheduler words (recon_kernel.cuh::sched_view): computed from the kernel parameters where they are used
  auto sv_ticket = [&]() { return a.sched; };
  auto sv_minv = [&]() { return a.sched + kSchedHeader; };
  auto sv_maxv = [&]() { return a.sched + kSchedHeader + a.nB; };
  auto sv_cnt = [&]() { return a.sched + kSchedHeader + 2 * a.nB; };
  const int W8m1 = (a.W >> 3) - 1;
  const unsigned rowbytes = (unsigned)a.W * 2u;
  // L2 policies (ReconArgs::hints): the raw pixel rows pass through L2 once (prefetch + one read) and must not push out the dB
  // scratch, which is read back 10 - 25 us after it was written
  const unsigned long long pol_in = w_policy((a.hints & 1) ? 1 : 0), pol_scr = w_policy((a.hints & 2) ? 2 : 0);
  const bool ld_hint = (a.hints & 4) != 0;
  // samples of the padded runs (W < NCH * 256) see gain 0, i.e. t - 1 = -1 without a subtrahend row: taken out of the mean
  const float pad_corr = HAS_SUB ? 0.f : (float)(NCH * 256 - a.W);

  auto njobs = [&]() { return a.nB * a.nparts * a.nsplit; };
  auto per_b = [&]() { return a.nparts * a.nsplit; };
#ifdef ABC_WROW_RING
  // EXPERIMENT (compiled in with -DABC_WROW_RING, off in the product build): the dB scratch as a ring of `ring`
  // B-scans - B-scan b lives in slot b mod ring - so that the 4 KB per A-scan would be written and read back in L2 lines that stay
  // resident instead of streaming through DRAM once each way.  A row of B-scan b >= ring may only be stored once every
  // normalisation job of B-scan b - ring has finished: done_p(b) counts them (release by the job's warp, acquire by the writer).
  // Measured: correct (emulation cases 9 / 10, GPU parity), DRAM traffic 15.8 -> 10.5 GB per launch at 64 MB, but the writers
  // block (the write -> read distance of the scratch exceeds 16 B-scans) and even the dormant guard costs 2 % of the headline.
  const int ring = a.ringB > 0 ? a.ringB : a.nB;
  auto done_p = [&](int b) { return a.sched + kSchedHeader + 3 * a.nB + 32 * b; };  // one 128-byte line per B-scan, zeroed by sched_init
  auto job_done = [&](int j) {  // all lanes
    w_syncwarp();
    if (lane == 0) {
      w_fence_gpu();
      w_atomic_add(done_p(j / per_b()), 1);  // result unused -> RED
    }
  };
#else
  auto job_done = [&](int) {};
#endif
  auto norm_args = [&]() {  // built from the kernel parameters (constant bank) at the call, not kept in registers
    WNormArgs na;
    na.scratch = a.scratch;
    na.out8 = a.out8;
    na.outdb = a.outdb;
    na.minv = sv_minv();
    na.maxv = sv_maxv();
    na.oph = a.oph;
    na.D = a.D;
    na.Dp = a.Dp;
    na.nparts = a.nparts;
    na.nsplit = a.nsplit;
    na.clamp55 = a.clamp55;
    na.thr = a.thr;
    na.clamp_db = a.clamp_db;
#ifdef ABC_WROW_RING
    na.ring = ring;
#endif
    return na;
  };

  auto row_ptr = [&](int item, int f) -> const uint8_t* {
    const int b = item / a.oph;
    return a.frames + ((size_t)b * a.A + f) * a.frame_stride + (size_t)(item - b * a.oph) * a.row_stride;
  };
  auto claim = [&]() -> int {  // dynamic schedule: one ticket = one row of one B-scan; the value lives in lane 0 until broadcast
    int t = 0;
    if (lane == 0) t = w_atomic_add(sv_ticket(), 1);
    return t;
  };

  // partner lane of the split step and the validity of this lane's outputs
  const int pl = (lane == 0 || lane >= R) ? lane : R - lane;
  const bool lane_ok = (R == 32) || lane < R;
  const int cc = (R == 32) ? lane : (lane < R ? lane : R - 1);

  // Lane 0's bookkeeping and the mailbox to the service warp live in shared memory (ints at st[]):
  //   st[0]       rows posted by this warp (written with cta-scope release: the service warp acquires it)
  //   st[1]       rows consumed by the service warp            st[2]  this warp has finished (no more posts)
  //   st[4..6]    B-scan and bounds of the last min / max this warp pushed (skip atomics that cannot change anything)
  //   st[7..9]    the finished row whose min / max have not been pushed yet: B-scan (-1: none), ordered min, ordered max
  //   st[10..17]  ring of the B-scans of the posted rows
  int* const st = reinterpret_cast<int*>(rawbuf + WP::RAWBUF + 16);
  if (lane == 0) {
    st[0] = st[1] = st[2] = 0;
    st[4] = st[7] = -1;
    st[5] = st[6] = 0;
  }
  w_syncthreads();  // the service warp reads every mailbox

  // ---- the normalisation job queue (global words next to the row ticket): jobs are B-scan major;
  //   sched[64] = jready: jobs [0, jready) belong to COMPLETE B-scans (a frontier, advanced by whoever sees the next B-scan's
  //              row count reach oph); sched[32] = jnext: jobs handed out so far (atomicAdd).
  // A job may run once its number is below jready.  Worker warps only take a job when a fresh snapshot says one is ready, and
  // never wait for one inside the row loop (the missing row could be their own); service warps and finished workers wait.
  auto jnext_p = [&]() { return a.sched + 32; };
  auto jready_p = [&]() { return a.sched + 64; };
  auto advance_frontier = [&]() {  // lane 0
    const int jr = w_ld_acquire(jready_p());
    if (jr >= njobs()) return jr;
    const int fb = jr / per_b();
    if (w_ld_acquire(sv_cnt() + fb) >= a.oph) {
      w_fence_gpu();  // cumulative: the rows counted above are visible to whoever acquires the new frontier
      w_atomic_max(jready_p(), (fb + 1) * per_b());
      return (fb + 1) * per_b();
    }
    return jr;
  };
  // all lanes: wait until job j is ready (lane 0 polls and helps the frontier along), then run it.  `service` is called
  // between polls (the service warp keeps publishing while it waits).
  auto run_job_when_ready = [&](int j, auto&& service) {
    const unsigned long long t_start = w_now_ns();
    for (;;) {
      int ok = 0;
      if (lane == 0) ok = advance_frontier() > j ? 1 : 0;
      if (w_shfl_i(ok, 0)) break;
      service();
      if (w_now_ns() - t_start > kWrowWatchdogNs) w_trap();  // a scheduling bug must surface as a launch failure, not as a hung GPU
      w_backoff();
    }
    wrow_normalise<3>(norm_args(), j, lane);
    job_done(j);
  };
  // all lanes: hand out and run jobs until none is left (service warps when idle, worker warps after their last row)
  auto drain_jobs = [&](auto&& service) {
    for (;;) {
      int j = 0;
      if (lane == 0) j = w_ld_relaxed(jnext_p()) < njobs() ? w_atomic_add(jnext_p(), 1) : 0x7fffffff;
      j = w_shfl_i(j, 0);
      if (j >= njobs()) break;
      run_job_when_ready(j, service);
    }
  };

  if (warp == WP::NW - 1) {
    // ================================================================================================ service warp
    // Completion counting for the whole CTA.  Lane l watches the mailbox of worker warp l: rows whose dB values the worker
    // has stored are published to the other SMs with ONE gpu-scope fence for all of them (the fence is cumulative over the
    // cta-scope release / acquire of the mailboxes), so the workers never execute a MEMBAR.GPU.  In between it advances the
    // job frontier and runs normalisation jobs itself.
    int* const mb = reinterpret_cast<int*>(smem + WP::TABLE_BYTES + (lane < WP::NWK ? lane : 0) * WP::WSTRIDE + WP::WBUF + WP::RAWBUF + 16);
    int seen = 0;
    auto collect = [&]() -> bool {  // all lanes; true if anything was published
      const int wr = lane < WP::NWK ? w_ld_acquire_cta(mb + 0) : 0;
      const int fresh = wr - seen;
      if (w_ballot(fresh > 0) == 0u) return false;
      w_fence_gpu();
      for (int k = 0; k < fresh; ++k) w_atomic_add(sv_cnt() + mb[10 + ((seen + k) & 7)], 1);  // result unused -> RED
      if (fresh > 0) {
        seen = wr;
        w_st_release_cta(mb + 1, seen);  // frees the ring slots
      }
      return true;
    };
    unsigned long long idle_since = 0;
    for (;;) {
      const bool any = collect();
      // a ready job?  (snapshot; the hand-out itself is an atomicAdd, an overshoot waits in run_job_when_ready)
      int j = 0x7fffffff;
      if (lane == 0) {
        const int jr = advance_frontier();
        if (w_ld_relaxed(jnext_p()) < jr) j = w_atomic_add(jnext_p(), 1);
      }
      j = w_shfl_i(j, 0);
      if (j < njobs()) {
        run_job_when_ready(j, [&]() { collect(); });
        idle_since = 0;
        continue;
      }
      const bool done = lane < WP::NWK ? (w_ld_acquire_cta(mb + 2) != 0 && w_ld_acquire_cta(mb + 0) == seen) : true;
      if (w_ballot(!done) == 0u) break;  // every worker of this CTA has finished and all their rows are published
      if (any) {
        idle_since = 0;
      } else {
        const unsigned long long now = w_now_ns();
        if (idle_since == 0) idle_since = now;
        if (now - idle_since > kWrowWatchdogNs) w_trap();
        w_backoff();
      }
    }
    drain_jobs([&]() {});  // the tail: the last B-scans complete when the last rows have been published
    return;
  }

  // ==================================================================================================== worker warps
  auto post = [&](int b) {  // lane 0: hand a finished row to the service warp
    const int wr = st[0];
    while (wr - w_ld_acquire_cta(st + 1) >= 8) w_backoff();  // ring full (the service warp is inside a job): rare
    st[10 + (wr & 7)] = b;
    w_st_release_cta(st + 0, wr + 1);
  };
  auto housekeep = [&]() {  // lane 0: min / max and completion of the previous row (its stores are a whole row old)
    const int hb = st[7];
    if (hb < 0) return;
    const int imn = st[8], imx = st[9];
    const float fmn = ordered_to_float(imn), fmx = ordered_to_float(imx);
    if (fmn <= fmx) {
      // most rows do not move the B-scan's extrema: skip the atomics when this warp already pushed tighter bounds
      float cmn = w_inf(false), cmx = w_inf(true);
      if (st[4] == hb) {
        cmn = ordered_to_float(st[5]);
        cmx = ordered_to_float(st[6]);
      }
      if (fmn < cmn) {
        w_atomic_min(sv_minv() + hb, imn);
        cmn = fmn;
      }
      if (fmx > cmx) {
        w_atomic_max(sv_maxv() + hb, imx);
        cmx = fmx;
      }
      st[4] = hb;
      st[5] = float_to_ordered(cmn);
      st[6] = float_to_ordered(cmx);
    }
    post(hb);
    st[7] = -1;
  };

  // Work queue of this warp: it0 = the row being processed, it1 = the next one (its pixels and calibration rows are loaded into
  // registers while it0 is finalised), it2 = the one after (its pixel row is pulled into L2 by the TMA unit).  Tickets are
  // claimed one item ahead of their first use so that no warp ever waits for the atomic.
  int it0 = w_shfl_i(claim(), 0);
  int it1 = w_shfl_i(claim(), 0);

  // ---- input of one (item, frame), see WPlan::LM.  ahead(): issued by the previous row once it is done with the warp buffer
  // (LM 1, 2: the TMA unit copies while that row is finished; LM 2 also starts the pixel loads into registers); take(): the
  // values of this lane's 8-sample runs in registers.
  constexpr int LM = WP::LM;
  uint4 raw[NCH];
  float4 gq[NCH][2];
  unsigned tma_parity = 0;
  auto pixel_row = [&](int item, int f) -> const uint8_t* {
    const int b = item / a.oph;
    return a.frames + ((size_t)b * a.A + f) * a.frame_stride + (size_t)(item - b * a.oph) * a.row_stride;
  };
  auto ldg_pixels = [&](const uint8_t* rp) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int run = lane + 32 * j;
      const void* pp = rp + 16 * (run < W8m1 ? run : W8m1);  // padded runs re-read the last run (finite values)
      raw[j] = ld_hint ? w_ldg_stream16_pol(pp, pol_in) : w_ldg_stream16(pp);
    }
  };
  auto ahead = [&](int item, int f) {
    if constexpr (LM != 0) {
      if (lane == 0) {
        const unsigned gbytes = (unsigned)a.calpitch * 4u;
        w_tma_arm(mbar, gbytes + (LM == 1 ? rowbytes : 0u));
        if constexpr (LM == 1) w_tma_load(rawbuf, pixel_row(item, f), rowbytes, mbar);
        w_tma_load(wbuf, a.gain + (size_t)(item % a.oph) * a.calpitch, gbytes, mbar);
      }
      if constexpr (LM == 2) ldg_pixels(pixel_row(item, f));
    }
  };
  auto take = [&](int item, int f) {
    if constexpr (LM == 0) {
      ldg_pixels(pixel_row(item, f));
      const float* gp = a.gain + (size_t)(item % a.oph) * a.calpitch + 4 * lane;
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        gq[j][0] = w_ldg_cal16(gp + (2 * j) * 128);
        gq[j][1] = w_ldg_cal16(gp + (2 * j + 1) * 128);
      }
    } else {
      w_mbar_wait(mbar, tma_parity);
      tma_parity ^= 1u;
      if constexpr (LM == 1) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const int run = lane + 32 * j;
          raw[j] = *reinterpret_cast<const uint4*>(rawbuf + 16 * (run < W8m1 ? run : W8m1));
        }
      }
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        gq[j][0] = *reinterpret_cast<const float4*>(wbuf + ((2 * j) * 32 + lane) * 16);
        gq[j][1] = *reinterpret_cast<const float4*>(wbuf + ((2 * j + 1) * 32 + lane) * 16);
      }
      w_syncwarp();  // every lane holds its gain values: the buffer may now be overwritten by the staged samples
    }
  };
  // (item, frame) two steps ahead of (it0, f) in this warp's sequence, for the L2 prefetch
  auto prefetch_step2 = [&](int f, int it2) {
    if (lane != 0) return;
    const int nA = A1 ? 1 : a.A;
    const int k = f + 2;
    const int item = k < nA ? it0 : (k < 2 * nA ? it1 : it2);
    const int fr = k < nA ? k : (k < 2 * nA ? k - nA : k - 2 * nA);
    if (item < a.nitems && fr < nA) {
      w_prefetch_l2_pol(row_ptr(item, fr), rowbytes, pol_in);
      if (LM == 0 && fr == 0) w_prefetch_l2(a.gain + (size_t)(item % a.oph) * a.calpitch, (unsigned)a.calpitch * 4u);
    }
  };
  if (it0 < a.nitems) {
    ahead(it0, 0);
    if (lane == 0) {
      const int nA = A1 ? 1 : a.A;
      if (nA > 1)
        w_prefetch_l2_pol(row_ptr(it0, 1), rowbytes, pol_in);
      else if (it1 < a.nitems)
        w_prefetch_l2_pol(row_ptr(it1, 0), rowbytes, pol_in);
    }
  }

  int nx_raw = claim();  // (lane 0) the ticket after it1, broadcast in the middle of the first row
  float acc1[16], acc2[16];
  if constexpr (!A1) {
#pragma unroll
    for (int d = 0; d < 16; ++d) acc1[d] = acc2[d] = 0.f;
  }

  int myjob = -1;  // a job this warp owns but whose B-scan was not complete yet when it was handed out (overshoot)
  while (it0 < a.nitems) {
    const int bscan = it0 / a.oph;
    const int row = it0 - bscan * a.oph;
    int it2 = 0x7fffffff;
    // (lane 0) snapshot of the job queue, consumed after the pre-processing phase: no wait, a few microseconds stale
    int jn_snap = 0, jr_snap = 0;
    if (lane == 0) {
      jn_snap = w_ld_relaxed(jnext_p());
      jr_snap = w_ld_relaxed(jready_p());
    }

    const int nA = A1 ? 1 : a.A;
    for (int f = 0; f < nA; ++f) {
      const bool last = A1 || (f + 1 == nA);
      // ---------------------------------------------------------------- pre: pixels -> s = t - mean (registers)
      take(it0, f);
      float2 s[NCH][4];  // 8 samples of run j as 4 packed pairs
      float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        float4 q0 = make_float4(1.f, 1.f, 1.f, 1.f), q1 = q0;
        if constexpr (HAS_SUB) {
          const float* sp = a.subg + (size_t)row * a.calpitch + 4 * lane;
          q0 = w_ldg_cal16(sp + (2 * j) * 128);
          q1 = w_ldg_cal16(sp + (2 * j + 1) * 128);
        }
        const unsigned w32[4] = {raw[j].x, raw[j].y, raw[j].z, raw[j].w};
        const float2 gg[4] = {make_float2(gq[j][0].x, gq[j][0].y), make_float2(gq[j][0].z, gq[j][0].w), make_float2(gq[j][1].x, gq[j][1].y),
                              make_float2(gq[j][1].z, gq[j][1].w)};
        const float2 qq[4] = {make_float2(-q0.x, -q0.y), make_float2(-q0.z, -q0.w), make_float2(-q1.x, -q1.y), make_float2(-q1.z, -q1.w)};
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          // u16 -> f32 without a conversion instruction: 0x4B00hhll is the float 2^23 + pixel, the subtraction is exact
          unsigned lo = w_byte_perm(w32[e2], 0x4B000000u, 0x7610), hi = w_byte_perm(w32[e2], 0x4B000000u, 0x7632);
          float2 y;
          memcpy(&y.x, &lo, 4);
          memcpy(&y.y, &hi, 4);
          y = pk_sub(y, make_float2(8388608.f, 8388608.f));
          // t - 1 = y * gain - (subg + 1): the constant keeps the staged values small (t ~ 1 for a normalised interferogram)
          const float2 tv = pk_fma(y, gg[e2], qq[e2]);
          s[j][e2] = tv;
          sum2 = pk_add(sum2, tv);
        }
      }
      float sum = sum2.x + sum2.y;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += w_shfl(sum, lane ^ o);
      const float mean = (sum + pad_corr) * a.inv_W;
      const float2 mean2 = make_float2(mean, mean);
#pragma unroll
      for (int j = 0; j < NCH; ++j)
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) s[j][e2] = pk_sub(s[j][e2], mean2);
      // ---------------------------------------------------------------- stage v[i] = P[i] s[i] - Q[i] s[i-1]
      if (lane == 0) *reinterpret_cast<float4*>(wbuf + WP::ZERO_OFF) = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        // s[i-1] of the run's first sample lives in the previous lane (lane 0: in lane 31's previous run)
        const float sendv = (lane == 31) ? (j > 0 ? s[j > 0 ? j - 1 : 0][3].y : 0.f) : s[j][3].y;
        const float prev = w_shfl(sendv, (lane + 31) & 31);
        const float4 P0 = t_pq(j * 4 + 0), P1 = t_pq(j * 4 + 1);
        const float4 Q0 = t_pq(j * 4 + 2), Q1 = t_pq(j * 4 + 3);
        const float2 p01 = pk_mul(s[j][0], make_float2(P0.x, P0.y)), p23 = pk_mul(s[j][1], make_float2(P0.z, P0.w));
        const float2 p45 = pk_mul(s[j][2], make_float2(P1.x, P1.y)), p67 = pk_mul(s[j][3], make_float2(P1.z, P1.w));
        const float v0 = fmaf(-Q0.x, prev, p01.x), v1 = fmaf(-Q0.y, s[j][0].x, p01.y);
        const float v2 = fmaf(-Q0.z, s[j][0].y, p23.x), v3 = fmaf(-Q0.w, s[j][1].x, p23.y);
        const float v4 = fmaf(-Q1.x, s[j][1].y, p45.x), v5 = fmaf(-Q1.y, s[j][2].x, p45.y);
        const float v6 = fmaf(-Q1.z, s[j][2].y, p67.x), v7 = fmaf(-Q1.w, s[j][3].x, p67.y);
        const int run = lane + 32 * j;  // padded runs (beyond W) stage zeros: P = Q = 0 there
        *reinterpret_cast<float4*>(wbuf + 16 * run) = make_float4(v0, v2, v4, v6);
        *reinterpret_cast<float4*>(wbuf + 4 * WP::PO + 16 * run) = make_float4(v1, v3, v5, v7);
      }
      w_syncwarp();
      if (f == 0) {
        it2 = w_shfl_i(nx_raw, 0);  // claimed in the middle of the previous row
        if (lane == 0) housekeep();
        // ---- a normalisation job?  (here nothing but the staged row is live)  At most one per row; never waited for.
        int j = -1;
        if (lane == 0) {
          if (myjob >= 0) {
            if (jr_snap > myjob) j = myjob;
          } else if (jn_snap + w_ncta() < jr_snap && jn_snap < njobs()) {  // only when the service warps have a backlog
            const int t = w_atomic_add(jnext_p(), 1);
            if (t < njobs()) {
              if (t < jr_snap)
                j = t;
              else
                myjob = t;  // overshoot: owned, run at a later row once the frontier has passed it
            }
          }
          if (j >= 0) {
            myjob = -1;
            w_acquire_fence();
          }
        }
        j = w_shfl_i(j, 0);
        if (j >= 0) {
          wrow_normalise<(WP::NW <= 12 ? 3 : 2)>(norm_args(), j, lane);
          job_done(j);
        }
#ifdef ABC_WROW_RING
        if (bscan >= ring) {  // ring guard: the slot this row goes to must have been consumed (normally it was, B-scans ago)
          const unsigned long long t_start = w_now_ns();
          for (;;) {
            int ok = 0;
            if (lane == 0) ok = w_ld_acquire(done_p(bscan - ring)) >= per_b() ? 1 : 0;
            if (w_shfl_i(ok, 0)) break;
            int jj = -1;  // a job this warp still owns may be one of those it is waiting for
            if (lane == 0 && myjob >= 0 && advance_frontier() > myjob) {
              jj = myjob;
              myjob = -1;
              w_acquire_fence();
            }
            jj = w_shfl_i(jj, 0);
            if (jj >= 0) {
              wrow_normalise<(WP::NW <= 12 ? 3 : 2)>(norm_args(), jj, lane);
              job_done(jj);
            }
            if (w_now_ns() - t_start > kWrowWatchdogNs) w_trap();
            w_backoff();
          }
        }
#endif
        nx_raw = claim();  // the next ticket: issued here, far from the store burst at the end of a row
      }
      // ---------------------------------------------------------------- pass A: gather, radix-R, twiddle, exchange
      float2 x[R], y[R];
      uint4 o_next = t_offs(0);
#pragma unroll
      for (int a2 = 0; a2 < R / 2; ++a2) {
        const uint4 o = o_next;
        if (a2 + 1 < R / 2) o_next = t_offs(a2 + 1);  // one table row ahead of its use
        x[2 * a2].x = *reinterpret_cast<const float*>(wbuf + o.x);
        x[2 * a2].y = *reinterpret_cast<const float*>(wbuf + o.y);
        x[2 * a2 + 1].x = *reinterpret_cast<const float*>(wbuf + o.z);
        x[2 * a2 + 1].y = *reinterpret_cast<const float*>(wbuf + o.w);
      }
      Dft<R, kFftSign, 1, 1>::run(x, y);
      w_syncwarp();  // every lane has gathered: the exchange rows may overwrite the staging planes
      float4 tw_next = t_twa(0);
#pragma unroll
      for (int p = 0; p < R / 2; ++p) {
        const float4 tw = tw_next;
        if (p + 1 < R / 2) tw_next = t_twa(p + 1);  // one table row ahead of its use
        const float2 y0 = p == 0 ? y[0] : cmul(y[2 * p], make_float2(tw.x, tw.y));
        const float2 y1 = cmul(y[2 * p + 1], make_float2(tw.z, tw.w));
        *reinterpret_cast<float4*>(wbuf + p * WP::XPITCH + 16 * lane) = make_float4(y0.x, y0.y, y1.x, y1.y);
      }
      w_syncwarp();
      // ---------------------------------------------------------------- pass B: radix-32 over the lanes of pass A
      float2 u[32], Z[32];
      {
        const unsigned char* xb = wbuf + (cc >> 1) * WP::XPITCH + (cc & 1) * 8;
#pragma unroll
        for (int b = 0; b < 32; ++b) u[b] = *reinterpret_cast<const float2*>(xb + 16 * b);
      }
      w_syncwarp();  // the buffer is free: the TMA unit fetches the next row's pixels and gain while this row is finished
      if (!last)
        ahead(it0, f + 1);
      else if (it1 < a.nitems)
        ahead(it1, 0);
      Dft<32, kFftSign, 1, 1>::run(u, Z);
      // ---------------------------------------------------------------- split + magnitude (+ finalise on the last frame)
      // The dB conversion and the scratch stores are fused into the split loop: every Z register dies as soon as its
      // pair has been formed, nothing but the running min / max is carried (no magnitude array).
      prefetch_step2(f, it2);
#ifdef ABC_WROW_RING
      float* const srow = a.scratch + ((size_t)(bscan % ring) * a.oph + row) * a.Dp;
#else
      float* const srow = a.scratch + ((size_t)bscan * a.oph + row) * a.Dp;
#endif
      float* const s1 = srow + lane;         // bin k1 = lane + R d
      float* const s2 = srow + (N2 - lane);  // bin k2 = N/2 - lane - R d
      float mn = w_inf(false), mx = w_inf(true);
      auto split_pass = [&](auto fin_c) {
        constexpr bool FIN = decltype(fin_c)::value;
#pragma unroll
        for (int d = 0; d < 16; ++d) {
          const int e = 31 - d;
          // partner value Z[N/2 - k]: slot 31 - d of lane R - c; lane 0 pairs slot d with slot 32 - d of ITSELF (d = 0: slot 16)
          float2 sv2 = Z[e];
          if (lane == 0) sv2 = e < 31 ? Z[e < 31 ? e + 1 : e] : Z[16];
          float2 Rv;
          Rv.x = w_shfl(sv2.x, pl);
          Rv.y = w_shfl(sv2.y, pl);
          const float4 tq = t_twp(d >> 1);
          const float Tr = (d & 1) ? tq.z : tq.x, Ti = (d & 1) ? tq.w : tq.y;
          const float2 z = Z[d];
          const float2 Rc = make_float2(Rv.x, -Rv.y);
          const float2 Av = pk_add(z, Rc), Dv = pk_sub(z, Rc);  // A = Z + conj Z', D = Z - conj Z'
          const float2 Bt = pk_mul(Dv, make_float2(Tr, Tr));
          const float2 Bv = make_float2(fmaf(-Ti, Dv.y, Bt.x), fmaf(Ti, Dv.x, Bt.y));  // B = T D
          const float2 pv = pk_add(Av, Bv), qv = pk_sub(Av, Bv);
          float a1 = fast_sqrt(fmaf(pv.x, pv.x, pv.y * pv.y)), a2 = fast_sqrt(fmaf(qv.x, qv.x, qv.y * qv.y));
          if (d == 0 && lane == 0) {  // the two self-conjugate bins: X[0] = Re Z0 + Im Z0, |X[N/4]| = |Z[N/4]| (same 1/2 scale as the rest)
            a1 = 2.f * fabsf(z.x + z.y);
            a2 = 2.f * fast_sqrt(fmaf(Rv.x, Rv.x, Rv.y * Rv.y));
          }
          if constexpr (!A1) {  // accumulate(magI, bscantransposed) over the frames of the B-scan (BscanFFT.cpp:1193-1209)
            a1 += acc1[d];
            a2 += acc2[d];
            acc1[d] = FIN ? 0.f : a1;
            acc2[d] = FIN ? 0.f : a2;
          }
          if constexpr (FIN) {
            // /A, + 1e-5, ln, * 20 / 2.303 (BscanFFT.cpp:1221-1237)
            const float db1 = fast_log2(fmaf(a1, a.out_scale, 1e-5f)) * a.db_scale;
            const float db2 = fast_log2(fmaf(a2, a.out_scale, 1e-5f)) * a.db_scale;
            const int k1 = lane + R * d;
            int k2 = N2 - lane - R * d;
            bool ok1 = lane_ok && (FULLD || k1 < a.D);
            bool ok2 = lane_ok && (FULLD || k2 < a.D);
            if (d == 0) {
              // special bins live in slot 0 only: 0, 1 (masked), 4 (the mask source), 5 (clampupper), and lane 0's second
              // output is bin N/4 instead of the non-existent bin N/2
              if (lane == 0) k2 = N2 / 2;
              ok2 = lane_ok && (lane == 0 ? (FULLD || N2 / 2 < a.D) : ok2);
              if (lane < 2) {
                if (a.dc01 != nullptr && ok1) a.dc01[2 * ((size_t)bscan * a.oph + row) + lane] = db1;  // kept on request only
                ok1 = false;  // bscandb.row(4).copyTo(row(1)), row(0): BscanFFT.cpp:1239-1240
              }
              if (lane == 4 && ok1) {
                w_st_keep_pol(srow, db1, pol_scr);
                w_st_keep_pol(srow + 1, db1, pol_scr);
              }
              if (ok1) w_st_keep_pol(s1, db1, pol_scr);
              if (ok2) w_st_keep_pol(lane == 0 ? srow + N2 / 2 : s2, db2, pol_scr);
              const bool is55 = a.clamp55 && lane == 5 && row == 5;  // forced element: excluded from the min / max of the data
              if (ok1 && !is55) {
                mn = fminf(mn, db1);
                mx = fmaxf(mx, db1);
              }
              if (ok2) {
                mn = fminf(mn, db2);
                mx = fmaxf(mx, db2);
              }
            } else if (FULLD && R == 32) {
              w_st_keep_pol(s1 + R * d, db1, pol_scr);
              w_st_keep_pol(s2 - R * d, db2, pol_scr);
              mn = w_min3(mn, db1, db2);
              mx = w_max3(mx, db1, db2);
            } else {
              if (ok1) {
                w_st_keep_pol(s1 + R * d, db1, pol_scr);
                mn = fminf(mn, db1);
                mx = fmaxf(mx, db1);
              }
              if (ok2) {
                w_st_keep_pol(s2 - R * d, db2, pol_scr);
                mn = fminf(mn, db2);
                mx = fmaxf(mx, db2);
              }
            }
          }
        }
      };
      if (!last) {
        split_pass(std::false_type{});
        continue;
      }
      split_pass(std::true_type{});
      // thresholded min / max of the B-scan (BscanFFT.cpp:1247, 1254): max(., thr) commutes with min / max.  The atomics and the
      // completion bookkeeping of this row are done by lane 0 in the middle of the NEXT row (housekeep), when the reduction
      // results have long arrived.
      const int imn = w_redux_min(float_to_ordered(fmaxf(mn, a.thr)));
      const int imx = w_redux_max(float_to_ordered(fmaxf(mx, a.thr)));
      if (lane == 0) {
        st[7] = bscan;
        st[8] = imn;
        st[9] = imx;
      }
    }
    it0 = it1;
    it1 = it2;
  }

  // ---- the last row: hand it over, then tell the service warp that this worker is done
  w_syncwarp();  // every lane's stores of the last row are ordered before lane 0's post
  if (lane == 0) {
    housekeep();
    w_st_release_cta(st + 2, 1);
  }
  // ---- the tail: an owned job first, then help with whatever is left
  myjob = w_shfl_i(myjob, 0);
  if (myjob >= 0) run_job_when_ready(myjob, [&]() {});
  drain_jobs([&]() {});
}

#ifdef __CUDACC__
template <class WP, bool HAS_SUB, bool A1, bool FULLD>
__global__ void __maxnreg__(WP::MAXREG) wrow_kernel(const ReconArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  wrow_body<WP, HAS_SUB, A1, FULLD>(a, smem);
}
#endif

}  // namespace abcoct
