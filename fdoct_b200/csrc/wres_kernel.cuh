// Resident-row fused reconstruction kernel (sm_100a): the block BscanFFT.cpp:987-1255 in one launch, one WARP per camera row,
// and the dB values of a row never leave the SM - one HBM read of the raw pixels and one HBM write of the display pixels per
// A-scan, nothing else.
//
// Why (round-2 measurement, profiles/r02a_*): the warp-per-A-scan kernel with a global dB scratch (wrow_kernel.cuh) moves
// 15.8 GB through DRAM per 1024-frame launch against 5.4 GB algorithmic - the L2 does not keep 4 KB per A-scan of freshly written
// scratch for the 10 - 40 us until the B-scan's global min / max are known (L2 policies make no difference, measured), so the
// scratch is written back and fetched again, and a third of all stall samples are waits on global memory.  Here
//   * the kernel issues no MMA, so the SM's 256 KB of TENSOR MEMORY are free: a finished row's 32 dB values per lane are parked
//     in 32 TMEM columns of the warp's own lane quarter (tcgen05.st; K = 4 - 5 rows per warp) and come back into the same
//     registers (tcgen05.ld) when the B-scan is complete - no shared memory, no global scratch, no fence over bulk data;
//   * four consecutive warps form a TEAM that owns 4 adjacent A-scans (a block).  When the B-scan is complete each warp quantises
//     its own row (threshold, global min-max, round-half-even, BscanFFT.cpp:1243-1255) into a small byte tile [bin][4] in shared
//     memory; when all four rows are there each warp writes a quarter of the bins, 4 display pixels per 32-bit store, into the
//     depth-major image;
//   * blocks are assigned statically (block g -> team g mod nteams): every row costs the same, no ticket atomics, and all rows of
//     a B-scan are in flight at the same time, so a B-scan completes about one row time after its first row;
//   * the completion protocol in global memory is one RED per row (count) plus the occasional min / max atomic.
// The row pipeline itself (pre-processing, staging, the two in-register FFT passes, split, dB) is the one of wrow_kernel.cuh.
//
// Protocol (per warp: rounds k = 0, 1, ...; TMEM slot k mod K).  jS = next round this warp has to STAGE (TMEM -> bytes in the
// team's tile jS & 1), possible when the B-scan of that round is complete and the tile's previous user (round jS - 2) has been
// written out by all four warps; jW = next round whose quarter this warp has to WRITE, possible when all four rows of it are
// staged.  Both are tried, without waiting, twice per row (where no register is live).  The only wait: before the dB values of
// round k go into their slot, round k - K must have been staged (jS > k - K).  Progress needs K * nteams >= blocks per B-scan
// (otherwise round k - K and round k of a team could lie in the same, incomplete B-scan): checked on the host.
#pragma once
#include <cstdint>
#include <cstring>

#include "fft_regs.cuh"
#include "plan.h"
#include "recon_kernel.cuh"
#include "wrow_kernel.cuh"
#include "wrow_prims.cuh"

namespace abcoct {

template <int N_, int NW_>
struct RPlan {
  static constexpr int N = N_, N2 = N_ / 2, R = N_ / 64, NW = NW_, LM = 0;
  static constexpr int TR = 4, NT = NW_ / TR;  // rows per block = warps per team; teams per CTA
  static constexpr int K = 16 / NT;            // TMEM slots per warp: 512 columns / 32 per row, shared by the NT warps of a lane quarter
  static_assert(N_ % 128 == 0 && R <= 32 && R >= 8, "N must be 128 * even, 512 <= N <= 2048");
  static_assert(NW_ % TR == 0 && NT >= 1 && NT <= 8 && K >= 2, "whole teams, at least 2 slots per warp");
  static constexpr int NCH = (N / 8 + 31) / 32;
  static constexpr int WMAX = NCH * 256;
  static constexpr int PO = WMAX / 2 + 4;
  static constexpr int STAGE_BYTES = (PO + WMAX / 2) * 4;
  static constexpr int XPITCH = 33 * 16;
  static constexpr int XCH_BYTES = (R / 2) * XPITCH;
  static constexpr int WBUF = ((cmax(STAGE_BYTES, XCH_BYTES) + 15) / 16) * 16;
  // table blob = shared-memory image
  static constexpr int T_OFFS = 0;
  static constexpr int T_TWA = T_OFFS + (R / 2) * 32 * 16;
  static constexpr int T_TWP = T_TWA + (R / 2) * 32 * 16;
  static constexpr int T_PQ = T_TWP + 8 * 32 * 16;
  static constexpr int TABLE_BYTES = T_PQ + NCH * 4 * 32 * 16;
  static constexpr int TILE_BYTES = N2 * 4;  // display bytes of a block: [bin][4 rows]
  static constexpr int CTL_BYTES = 128;      // per team: staged[2] at int 0, done[2] at int 2, per-warp {B-scan, min, max} at int 8 + 4 wi
  static constexpr int OFF_WBUF = TABLE_BYTES;
  static constexpr int OFF_TILES = OFF_WBUF + NW * WBUF;
  static constexpr int OFF_CTL = OFF_TILES + NT * 2 * TILE_BYTES;
  static constexpr int OFF_TMEM = OFF_CTL + NT * CTL_BYTES;  // the TMEM base address written by tcgen05.alloc; + 4: frontier, + 8: time of the last poll
  static constexpr int SMEM_BYTES = OFF_TMEM + 16;
  static_assert(SMEM_BYTES <= 227 * 1024, "too many warps for the shared memory");
  static constexpr int ZERO_OFF = (WMAX / 2) * 4;
  static constexpr int MAXREG = cmax(32, ((65536 / (NW * 32)) / 8) * 8 > 255 ? 255 : ((65536 / (NW * 32)) / 8) * 8);
};

// ------------------------------------------------------------------------------------------------- the kernel body
template <class RP, bool HAS_SUB, bool A1, bool FULLD>
WROW_HD void wres_body(const ReconArgs& a, unsigned char* smem) {
  constexpr int R = RP::R, NCH = RP::NCH, N2 = RP::N2, K = RP::K;
  const int lane = w_lane(), warp = w_warp_in_cta();
  {  // tables: global (L2-resident) -> shared, once per persistent CTA; team counters start at zero
    const uint4* src = reinterpret_cast<const uint4*>(a.idxT);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = warp * 32 + lane; i < RP::TABLE_BYTES / 16; i += RP::NW * 32) dst[i] = src[i];
    int* ctl0 = reinterpret_cast<int*>(smem + RP::OFF_CTL);
    for (int i = warp * 32 + lane; i < RP::NT * RP::CTL_BYTES / 4; i += RP::NW * 32) ctl0[i] = ((i & 31) >= 8 && (i & 3) == 0) ? -1 : 0;
    if (warp == 0 && lane < 4) reinterpret_cast<int*>(smem + RP::OFF_TMEM)[lane] = 0;
  }
  const unsigned tmem_base = w_tmem_alloc512(reinterpret_cast<unsigned*>(smem + RP::OFF_TMEM));  // includes the CTA barrier
  const unsigned char* const tbl = smem + 16 * lane;
  auto t_offs = [&](int i) { return *reinterpret_cast<const uint4*>(tbl + RP::T_OFFS + 512 * i); };
  auto t_pq = [&](int i) { return *reinterpret_cast<const float4*>(tbl + RP::T_PQ + 512 * i); };
  auto t_twa = [&](int i) { return *reinterpret_cast<const float4*>(tbl + RP::T_TWA + 512 * i); };
  auto t_twp = [&](int i) { return *reinterpret_cast<const float4*>(tbl + RP::T_TWP + 512 * i); };
  unsigned char* const wbuf = smem + RP::OFF_WBUF + warp * RP::WBUF;
  const int team = warp >> 2, wi = warp & 3;
  // this warp's TMEM slots: lane quarter wi (the hardware rule: warp w reaches lanes 32 (w % 4) ..), 32 columns per slot
  const unsigned tslot0 = tmem_base + ((unsigned)(32 * wi) << 16) + (unsigned)(team * K * 32);
  unsigned char* const tiles = smem + RP::OFF_TILES + team * 2 * RP::TILE_BYTES;  // tile p: + p * TILE_BYTES
  int* const ctl = reinterpret_cast<int*>(smem + RP::OFF_CTL + team * RP::CTL_BYTES);
  int* const staged = ctl;    // [2] rows staged into tile p, all rounds
  int* const done = ctl + 2;  // [2] quarters written out of tile p, all rounds
  int* const mm = ctl + 8 + 4 * wi;  // {B-scan of the bounds below (-1: none), ordered min, ordered max} pushed by this warp so far

  auto sv_minv = [&]() { return a.sched + kSchedHeader; };
  auto sv_maxv = [&]() { return a.sched + kSchedHeader + a.nB; };
  auto sv_cnt = [&](int b) { return a.sched + kSchedHeader + 3 * a.nB + 32 * b; };  // one 128-byte line per B-scan
  // completion frontier of this CTA: B-scans [0, *frontier) are known to be complete.  It is advanced by whichever warp needs it,
  // at most one global poll per kPollNs and CTA - thousands of warps polling the counters themselves made the counts crawl
  int* const frontier = reinterpret_cast<int*>(smem + RP::OFF_TMEM) + 1;
  unsigned* const tpoll = reinterpret_cast<unsigned*>(smem + RP::OFF_TMEM) + 2;
  constexpr unsigned kPollNs = 400;
  auto bscan_complete = [&](int b) -> bool {  // lane 0
    int f = w_ld_acquire_cta(frontier);
    if (f > b) return true;
    const unsigned now = w_now_ns32(), last = *reinterpret_cast<volatile unsigned*>(tpoll);
    if (now - last < kPollNs) return false;
    if (w_cas_smem(tpoll, last, now) != last) return false;  // somebody else polls
    const int f0 = f;
    while (f < a.nB && f < f0 + 4 && w_ld_relaxed(sv_cnt(f)) >= a.oph) ++f;
    if (f == f0) return false;
    (void)w_ld_acquire(sv_cnt(f - 1));  // acquire: the min / max atomics of the rows counted are visible from here on ...
    w_acquire_fence();                  // ... for every B-scan below the new frontier
    w_max_release_smem(frontier, f);    // ... and for whoever acquires the frontier
    return f > b;
  };
  const int W8m1 = (a.W >> 3) - 1;
  const unsigned rowbytes = (unsigned)a.W * 2u;
  const float pad_corr = HAS_SUB ? 0.f : (float)(NCH * 256 - a.W);

  // ---- static schedule: block g = tg + k * nteams (k = 0, 1, ...: the rounds of this team), 4 rows per block
  const int nteams = w_ncta() * RP::NT;
  const int tg = team * w_ncta() + w_cta();
  const int nbb = (a.oph + 3) >> 2;  // blocks per B-scan
  const int G = a.nB * nbb;
  const int nk = tg < G ? (G - 1 - tg) / nteams + 1 : 0;
  struct RowId {
    int b, r0;
  };
  auto block_of = [&](int k) {
    const int g = tg + k * nteams;
    RowId r;
    r.b = g / nbb;
    r.r0 = 4 * (g - r.b * nbb);
    return r;
  };
  auto row_ptr = [&](const RowId& r, int f) -> const uint8_t* {
    return a.frames + ((size_t)r.b * a.A + f) * a.frame_stride + (size_t)(r.r0 + wi) * a.row_stride;
  };

  // ---- stage: this warp's row of round jS, TMEM -> display bytes in the team's tile (all lanes; conditions checked by lane 0)
  int jS = 0, jW = 0;
  const bool direct = (a.hints & 8) != 0;  // A/B: every warp writes the display bytes of its own row itself (no team tile, 1-byte stores)
  RowId sblk = block_of(0);  // block of round jS
  auto can_stage = [&]() -> bool {  // lane 0
    if (!direct && w_ld_acquire_cta(done + (jS & 1)) < 4 * (jS >> 1)) return false;  // the tile's previous block has been written out
    return bscan_complete(sblk.b);  // every row of the B-scan has been counted
  };
  auto stage_next = [&]() {  // all lanes
    const int row = sblk.r0 + wi;
    if (row < a.oph) {
      int imn = 0, imx = 0;
      if (lane == 0) {
        imn = w_ld_cg_i(sv_minv() + sblk.b);
        imx = w_ld_cg_i(sv_maxv() + sblk.b);
      }
      float mn = ordered_to_float(w_shfl_i(imn, 0)), mx = ordered_to_float(w_shfl_i(imx, 0));
      if (a.clamp55) {  // bscandisp.at<double>(5,5) = 50.0 before the min-max (BscanFFT.cpp:1248-1253)
        mn = fminf(mn, a.clamp_db);
        mx = fmaxf(mx, a.clamp_db);
      }
      const float sc = (mx - mn) > 2.220446049250313e-16f ? 255.0f / (mx - mn) : 0.f;  // cv::normalize: scale = 0 for a flat image
      const float thr = a.thr;
      auto quant = [&](float x) -> unsigned {  // round-half-even of (max(x, thr) - mn) * 255 / (mx - mn): 1.5 * 2^23 trick, low byte
        float r = fmaf(fmaxf(x, thr) - mn, sc, 12582912.0f);
        unsigned u;
        memcpy(&u, &r, 4);
        return u & 0xffu;
      };
      unsigned char* const tile = tiles + (jS & 1) * RP::TILE_BYTES + wi;
      const bool lane_okS = (R == 32) || lane < R;
      const bool is55row = a.clamp55 && row == 5;
      float* const od = a.outdb != nullptr ? a.outdb + (size_t)sblk.b * a.D * a.oph + row : nullptr;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v[16];
        w_tmem_ld16(tslot0 + (unsigned)((jS % K) * 32 + 16 * h), v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int d = 8 * h + (i >> 1);
          int bin = (i & 1) ? N2 - lane - R * d : lane + R * d;
          if ((i & 1) && d == 0 && lane == 0) bin = N2 / 2;
          if (lane_okS && bin < a.D) {
            unsigned q = quant(v[i]);
            if (is55row && bin == 5) q = quant(a.clamp_db);  // the forced element
            if (direct)
              w_st_global_u8(a.out8 + ((size_t)sblk.b * a.D + bin) * a.oph + row, q);
            else
              w_st_shared_u8(tile + 4 * bin, q);
            if (od != nullptr) w_st_keep(od + (size_t)bin * a.oph, v[i]);  // transposed dB image (on request)
          }
        }
      }
    }
    w_syncwarp();
    if (lane == 0 && !direct) w_red_release_cta(staged + (jS & 1), 1);
    ++jS;
    if (direct) jW = jS;
    sblk = block_of(jS);
  };
  // ---- write: this warp's quarter of the bins of round jW, 4 display pixels per store
  auto can_write = [&]() -> bool { return w_ld_acquire_cta(staged + (jW & 1)) >= 4 * ((jW >> 1) + 1); };  // lane 0
  auto write_next = [&]() {  // all lanes
    const RowId r = block_of(jW);
    const int nrows = (a.oph - r.r0) < 4 ? (a.oph - r.r0) : 4;
    const int ntiles = (a.D + 31) >> 5;
    const int tq = (ntiles + 3) >> 2;
    const int t0 = wi * tq;
    const int t1 = (t0 + tq) < ntiles ? (t0 + tq) : ntiles;
    const bool word_ok = nrows == 4 && (a.oph & 3) == 0 && (reinterpret_cast<uintptr_t>(a.out8) & 3) == 0;
    const unsigned char* const tile = tiles + (jW & 1) * RP::TILE_BYTES;
    uint8_t* const o_base = a.out8 + (size_t)r.b * a.D * a.oph + r.r0;
    for (int t = t0; t < t1; ++t) {
      const int bin = 32 * t + lane;
      if (bin < a.D) {
        const unsigned word = *reinterpret_cast<const unsigned*>(tile + 4 * bin);
        uint8_t* o = o_base + (size_t)bin * a.oph;
        if (word_ok) {
          w_st_global_u32(o, word);
        } else {
          for (int i = 0; i < nrows; ++i) w_st_global_u8(o + i, (word >> (8 * i)) & 0xffu);
        }
      }
    }
    w_syncwarp();
    if (lane == 0) w_red_release_cta(done + (jW & 1), 1);
    ++jW;
  };
  auto service = [&](int kmax) -> int {  // at most one stage and one write of rounds < kmax; never waits
    int go = 0;
    if (lane == 0) go = ((jS < kmax && can_stage()) ? 1 : 0) | ((jW < jS && can_write()) ? 2 : 0);
    go = w_shfl_i(go, 0);
    if (go & 1) stage_next();
    if (go & 2) write_next();
    return go;
  };
  auto acquire_slot = [&](int k) {  // before the dB values of round k go into TMEM slot k mod K: round k - K must have been staged
    unsigned spins = 0;
    unsigned long long t_start = 0;
    while (jS <= k - K) {
      if (service(k) == 0) {
        if ((++spins & 255u) == 0u) {  // a protocol bug must surface as a launch failure, not as a hung GPU
          const unsigned long long now = w_now_ns();
          if (t_start == 0) t_start = now;
          if (now - t_start > kWrowWatchdogNs) w_trap();
        }
        w_backoff_short();
      }
    }
  };

  // partner lane of the split step and the validity of this lane's outputs
  const int pl = (lane == 0 || lane >= R) ? lane : R - lane;
  const bool lane_ok = (R == 32) || lane < R;
  const int cc = (R == 32) ? lane : (lane < R ? lane : R - 1);

  uint4 raw[NCH];
  float4 gq[NCH][2];
  auto take = [&](const RowId& r, int f) {
    const uint8_t* rp = row_ptr(r, f);
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int run = lane + 32 * j;
      raw[j] = w_ldg_stream16(rp + 16 * (run < W8m1 ? run : W8m1));  // padded runs re-read the last run (finite values)
    }
    const float* gp = a.gain + (size_t)(r.r0 + wi) * a.calpitch + 4 * lane;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      gq[j][0] = w_ldg_cal16(gp + (2 * j) * 128);
      gq[j][1] = w_ldg_cal16(gp + (2 * j + 1) * 128);
    }
  };
  const int nA = A1 ? 1 : a.A;
  // pixel row `ahead` steps (frames) after (k, f) in this warp's sequence -> L2, by the TMA unit
  auto prefetch_step = [&](int k, int f, int ahead) {
    if (lane != 0) return;
    int kk = k, ff = f + ahead;
    while (ff >= nA) {
      ff -= nA;
      ++kk;
    }
    if (kk >= nk) return;
    const RowId r = block_of(kk);
    if (r.r0 + wi >= a.oph) return;
    w_prefetch_l2(row_ptr(r, ff), rowbytes);
    if (ff == 0) w_prefetch_l2(a.gain + (size_t)(r.r0 + wi) * a.calpitch, (unsigned)a.calpitch * 4u);
  };
  if (nk > 0) prefetch_step(0, 0, 1);

  float acc1[16], acc2[16];
  if constexpr (!A1) {
#pragma unroll
    for (int d = 0; d < 16; ++d) acc1[d] = acc2[d] = 0.f;
  }

  for (int k = 0; k < nk; ++k) {
    const RowId rid = block_of(k);
    const int bscan = rid.b;
    const int row = rid.r0 + wi;
    if (row >= a.oph) {  // the last block of a B-scan may be partial: this warp has no row, but it keeps the team protocol going
      service(k);
      acquire_slot(k);
      continue;
    }
    const unsigned tslot = tslot0 + (unsigned)((k % K) * 32);

    for (int f = 0; f < nA; ++f) {
      const bool last = A1 || (f + 1 == nA);
      // ---------------------------------------------------------------- pre: pixels -> s = t - mean (registers)
      take(rid, f);
      float2 s[NCH][4];
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
          // t - 1 = y * gain - (subg + 1)  (BscanFFT.cpp:987, 1132, BscanDark.cpp:1269)
          const float2 tv = pk_fma(y, gg[e2], qq[e2]);
          s[j][e2] = tv;
          sum2 = pk_add(sum2, tv);
        }
      }
      float sum = sum2.x + sum2.y;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += w_shfl(sum, lane ^ o);
      const float mean = (sum + pad_corr) * a.inv_W;  // BscanFFT.cpp:1135-1139
      const float2 mean2 = make_float2(mean, mean);
#pragma unroll
      for (int j = 0; j < NCH; ++j)
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) s[j][e2] = pk_sub(s[j][e2], mean2);
      // ---------------------------------------------------------------- stage v[i] = P[i] s[i] - Q[i] s[i-1]
      if (lane == 0) *reinterpret_cast<float4*>(wbuf + RP::ZERO_OFF) = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
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
        const int run = lane + 32 * j;
        *reinterpret_cast<float4*>(wbuf + 16 * run) = make_float4(v0, v2, v4, v6);
        *reinterpret_cast<float4*>(wbuf + 4 * RP::PO + 16 * run) = make_float4(v1, v3, v5, v7);
      }
      w_syncwarp();
      prefetch_step(k, f, 2);
      // ---- nothing but the staged row is live here: the cheap place to finish earlier rounds (never waits)
      if (f == 0) service(k);
      // ---------------------------------------------------------------- pass A: gather, radix-R, twiddle, exchange
      {
        float2 x[R], y[R];
        uint4 o_next = t_offs(0);
#pragma unroll
        for (int a2 = 0; a2 < R / 2; ++a2) {
          const uint4 o = o_next;
          if (a2 + 1 < R / 2) o_next = t_offs(a2 + 1);
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
          if (p + 1 < R / 2) tw_next = t_twa(p + 1);
          const float2 y0 = p == 0 ? y[0] : cmul(y[2 * p], make_float2(tw.x, tw.y));
          const float2 y1 = cmul(y[2 * p + 1], make_float2(tw.z, tw.w));
          *reinterpret_cast<float4*>(wbuf + p * RP::XPITCH + 16 * lane) = make_float4(y0.x, y0.y, y1.x, y1.y);
        }
      }
      w_syncwarp();
      // ---- between the passes no register is live either: a second chance, and the TMEM slot of this round is taken here
      // (with K slots per warp that waits only when some warp of the GPU is K - 1 rounds behind)
      if (last) {
        service(k);
        acquire_slot(k);
      }
      // ---------------------------------------------------------------- pass B: radix-32 over the lanes of pass A
      float2 u[32], Z[32];
      {
        const unsigned char* xb = wbuf + (cc >> 1) * RP::XPITCH + (cc & 1) * 8;
#pragma unroll
        for (int b = 0; b < 32; ++b) u[b] = *reinterpret_cast<const float2*>(xb + 16 * b);
      }
      w_syncwarp();
      Dft<32, kFftSign, 1, 1>::run(u, Z);
      // ---------------------------------------------------------------- split + magnitude (+ finalise on the last frame)
      float mn = w_inf(false), mx = w_inf(true);
      float dbv[16];  // the dB values of 8 split steps: value 2 d' = bin lane + R d, value 2 d' + 1 = bin N/2 - lane - R d
      auto split_pass = [&](auto fin_c) {
        constexpr bool FIN = decltype(fin_c)::value;
#pragma unroll
        for (int d = 0; d < 16; ++d) {
          const int e = 31 - d;
          float2 sv2 = Z[e];
          if (lane == 0) sv2 = e < 31 ? Z[e < 31 ? e + 1 : e] : Z[16];
          float2 Rv;
          Rv.x = w_shfl(sv2.x, pl);
          Rv.y = w_shfl(sv2.y, pl);
          const float4 tq = t_twp(d >> 1);
          const float Tr = (d & 1) ? tq.z : tq.x, Ti = (d & 1) ? tq.w : tq.y;
          const float2 z = Z[d];
          const float2 Rc = make_float2(Rv.x, -Rv.y);
          const float2 Av = pk_add(z, Rc), Dv = pk_sub(z, Rc);
          const float2 Bt = pk_mul(Dv, make_float2(Tr, Tr));
          const float2 Bv = make_float2(fmaf(-Ti, Dv.y, Bt.x), fmaf(Ti, Dv.x, Bt.y));
          const float2 pv = pk_add(Av, Bv), qv = pk_sub(Av, Bv);
          float a1 = fast_sqrt(fmaf(pv.x, pv.x, pv.y * pv.y)), a2 = fast_sqrt(fmaf(qv.x, qv.x, qv.y * qv.y));
          if (d == 0 && lane == 0) {
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
            float db1 = fast_log2(fmaf(a1, a.out_scale, 1e-5f)) * a.db_scale;
            const float db2 = fast_log2(fmaf(a2, a.out_scale, 1e-5f)) * a.db_scale;
            const int k1 = lane + R * d;
            int k2 = N2 - lane - R * d;
            bool ok1 = lane_ok && (FULLD || k1 < a.D);
            bool ok2 = lane_ok && (FULLD || k2 < a.D);
            if (d == 0) {
              if (lane == 0) k2 = N2 / 2;
              ok2 = lane_ok && (lane == 0 ? (FULLD || N2 / 2 < a.D) : ok2);
              if (lane < 2 && a.dc01 != nullptr && ok1) a.dc01[2 * ((size_t)bscan * a.oph + row) + lane] = db1;  // kept on request only
              // bscandb.row(4).copyTo(row(1)), row(0): BscanFFT.cpp:1239-1240 - bins 0 and 1 (lanes 0, 1) take the value of bin 4
              const float db4 = w_shfl(db1, 4);
              if (lane < 2) db1 = db4;
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
              mn = w_min3(mn, db1, db2);
              mx = w_max3(mx, db1, db2);
            } else {
              if (ok1) {
                mn = fminf(mn, db1);
                mx = fmaxf(mx, db1);
              }
              if (ok2) {
                mn = fminf(mn, db2);
                mx = fmaxf(mx, db2);
              }
            }
            dbv[2 * (d & 7)] = db1;
            dbv[2 * (d & 7) + 1] = db2;
            if ((d & 7) == 7) w_tmem_st16(tslot + (unsigned)(16 * (d >> 3)), dbv);  // 16 values -> 16 TMEM columns of this lane
          }
        }
      };
      if (!last) {
        split_pass(std::false_type{});
        continue;
      }
      split_pass(std::true_type{});
      // thresholded min / max of the B-scan (BscanFFT.cpp:1247, 1254): max(., thr) commutes with min / max
      const int imn = w_redux_min(float_to_ordered(fmaxf(mn, a.thr)));
      const int imx = w_redux_max(float_to_ordered(fmaxf(mx, a.thr)));
      w_tmem_st_wait();  // the row is in tensor memory before it is counted
      if (lane == 0) {
        const float fmn = ordered_to_float(imn), fmx = ordered_to_float(imx);
        bool pushed = false;
        if (fmn <= fmx) {
          // most rows do not move the B-scan's extrema: skip the atomics when this warp already pushed tighter bounds
          float cmn = w_inf(false), cmx = w_inf(true);
          if (mm[0] == bscan) {
            cmn = ordered_to_float(mm[1]);
            cmx = ordered_to_float(mm[2]);
          }
          if (fmn < cmn) {
            w_atomic_min(sv_minv() + bscan, imn);
            cmn = fmn;
            pushed = true;
          }
          if (fmx > cmx) {
            w_atomic_max(sv_maxv() + bscan, imx);
            cmx = fmx;
            pushed = true;
          }
          mm[0] = bscan;
          mm[1] = float_to_ordered(cmn);
          mm[2] = float_to_ordered(cmx);
        }
        // the count: released at gpu scope only when this row pushed a bound (the atomics must be visible before the count);
        // otherwise nothing of this row lives in global memory and a relaxed RED does
        if (pushed)
          w_release_add(sv_cnt(bscan), 1);
        else
          w_red_relaxed(sv_cnt(bscan), 1);
      }
    }
  }
  // ---- the tail: the rounds this warp has not staged / written yet
  {
    unsigned spins = 0;
    unsigned long long t_start = 0;
    while (jW < nk) {
      if (service(nk) == 0) {
        if ((++spins & 255u) == 0u) {
          const unsigned long long now = w_now_ns();
          if (t_start == 0) t_start = now;
          if (now - t_start > kWrowWatchdogNs) w_trap();
        }
        w_backoff_short();
      }
    }
  }
  w_tmem_free512(tmem_base);  // includes the CTA barrier: every warp has read its last row back
}

#ifdef __CUDACC__
template <class RP, bool HAS_SUB, bool A1, bool FULLD>
__global__ void __maxnreg__(RP::MAXREG) wres_kernel(const ReconArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  wres_body<RP, HAS_SUB, A1, FULLD>(a, smem);
}
#endif

}  // namespace abcoct
