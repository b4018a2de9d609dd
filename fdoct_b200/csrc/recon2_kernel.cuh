// Dual-pair variant of the fused reconstruction kernel (sm_100a): one thread group works on TWO row pairs (four
// A-scans) at once and keeps them side by side in the two lanes of 64-bit registers, so that every butterfly, twiddle
// multiply, lerp and magnitude of the FFT path is ONE packed FADD2 / FMUL2 / FFMA2 instead of two scalar instructions.
// The kernel is instruction-issue bound, so halving the FP instruction count - and sharing every table
// read, address computation and barrier between two pairs - is what moves it.  Semantics, scheduling (tickets, L2
// scratch, fused normalisation) and the reference citations are those of recon_kernel.cuh; only the data layout differs:
//
//   staging buffer   16 bytes per spectral sample: {rowA pair0, rowA pair1, rowB pair0, rowB pair1}, i.e. the packed
//                    real lane pair and the packed imaginary lane pair of the two complex transforms
//   exchange buffer  two planes (re, im) of 8-byte packed lanes, same element order as the single-pair kernel
//   tables           lerp weight / lerp(window) and the inter-pass twiddles are stored lane-duplicated ({w, w}) because
//                    the packed instructions take no scalar-broadcast register operand
#pragma once
#include "fft_v.cuh"
#include "recon_kernel.cuh"

namespace abcoct {

// ------------------------------------------------------------------------------------------- shared memory map
struct SmemLayout2 {
  int offT, wvT, win, tw0, tw1;  // CTA-wide tables (byte offsets)
  int groups;
  int g_gain, g_subg, g_buf, g_red, g_mbar, group_bytes;
  __host__ __device__ constexpr int total(int G) const { return groups + G * group_bytes; }
};
template <class P>
__host__ __device__ constexpr SmemLayout2 make_layout2(int W, bool has_sub) {
  SmemLayout2 L{};
  int o = 0;
  L.offT = o; o = align16(o + P::R0P4 * P::N1 * 4);
  L.wvT = o;  o = align16(o + P::R0P4 * P::N1 * 16);
  L.win = o;  o = align16(o + W * 4);
  L.tw0 = o;  o = align16(o + (P::R0 - 1) * P::N1 * 16);
  L.tw1 = o;  o = align16(o + (P::THREE ? (P::R1 - 1) * P::N2 * 16 : 0));
  L.groups = o;
  int g = 0;
  L.g_gain = g; g = align16(g + 4 * W * 4);
  L.g_subg = g; g = align16(g + (has_sub ? 4 * W * 4 : 0));
  L.g_buf = g;  g = align16(g + cmax(cmax((W + 1) * 16, P::BUF * 16), kNormBins * (P::T + 4)));
  L.g_red = g;  g = align16(g + 4 * P::NWARPS * 4 + 32 * 4);
  L.g_mbar = g; g = align16(g + 16);
  L.group_bytes = g;
  return L;
}

struct GroupSmem2 {
  const uint32_t* offT;  // per gathered sample: byte offset of y[i] | byte offset of y[i-1] << 16 in the staging buffer
  const float4* wvT;     // per gathered sample: {wq, wq, lerp(window), lerp(window)}
  const float* win;
  const float4* tw0;     // {wx, wx, wy, wy}
  const float4* tw1;
  float* gain;           // [4][W]: rowA pair0, rowB pair0, rowA pair1, rowB pair1
  float* subg;
  unsigned char* stg;    // staging view of g_buf (16 bytes per sample)
  V2* bre;               // exchange view of g_buf: real plane
  V2* bim;               //                         imaginary plane
  float* red;            // [NWARPS][4] partial row sums
  int* slot;             // 32 ints
  unsigned long long* mbar;
};
template <class P>
__host__ __device__ inline GroupSmem2 resolve2(unsigned char* base, const SmemLayout2& L, int g) {
  unsigned char* gb = base + L.groups + g * L.group_bytes;
  GroupSmem2 s;
  s.offT = reinterpret_cast<const uint32_t*>(base + L.offT);
  s.wvT = reinterpret_cast<const float4*>(base + L.wvT);
  s.win = reinterpret_cast<const float*>(base + L.win);
  s.tw0 = reinterpret_cast<const float4*>(base + L.tw0);
  s.tw1 = reinterpret_cast<const float4*>(base + L.tw1);
  s.gain = reinterpret_cast<float*>(gb + L.g_gain);
  s.subg = reinterpret_cast<float*>(gb + L.g_subg);
  s.stg = gb + L.g_buf;
  s.bre = reinterpret_cast<V2*>(gb + L.g_buf);
  s.bim = s.bre + P::BUF;
  s.red = reinterpret_cast<float*>(gb + L.g_red);
  s.slot = reinterpret_cast<int*>(gb + L.g_red + 4 * P::NWARPS * 4);
  s.mbar = reinterpret_cast<unsigned long long*>(gb + L.g_mbar);
  return s;
}

// staging swizzle: a thread stores the eight 16-byte units of chunk ch with lanes 128 bytes apart; unit j goes to
// j ^ (ch & 7) so that a quarter-warp covers all 32 banks.  In sample indices: bits 0..2 ^= bits 3..5.
ABC_HD int stg2_phys(int i) { return i ^ ((i >> 3) & 7); }

template <class P>
struct Thread2State {
  uint4 raw[2][2][P::NCH];     // [pair][row][chunk]
  Cx<V2> x[P::NB0][P::R0];     // resampled inputs of the first-pass butterflies, lane = pair
  V2 acc[P::NU][P::RL][2];     // 2 * sum over frames of |A[k]|, |B[k]|, lane = pair
};

// ------------------------------------------------------------------------------------------- phases
template <class P>
ABC_HD void phase2_load(int tid, const uint8_t* const (&rows)[4], int W8, Thread2State<P>& r, unsigned long long pol = 0) {
#pragma unroll
  for (int i = 0; i < P::NCH; ++i) {
    const int ch = tid + P::T * i;
    if (ch < W8) {
      r.raw[0][0][i] = load_raw16(rows[0] + 16 * ch, pol);
      r.raw[0][1][i] = load_raw16(rows[1] + 16 * ch, pol);
      r.raw[1][0][i] = load_raw16(rows[2] + 16 * ch, pol);
      r.raw[1][1][i] = load_raw16(rows[3] + 16 * ch, pol);
    }
  }
}

// see phase_pre in recon_kernel.cuh; sums[2 * pair + row]
template <class P, bool HAS_SUB>
ABC_HD void phase2_pre(int tid, const GroupSmem2& s, int W, const Thread2State<P>& r, float (&sums)[4]) {
  const int W8 = W >> 3;
  sums[0] = sums[1] = sums[2] = sums[3] = 0.f;
  if (tid == 0) *reinterpret_cast<float4*>(s.stg + 16 * W) = make_float4(0.f, 0.f, 0.f, 0.f);  // sentinel (q = 0, N-1)
#pragma unroll
  for (int i = 0; i < P::NCH; ++i) {
    const int ch = tid + P::T * i;
    if (ch < W8) {
      const int h4 = (ch & 4);
      const float4 w0 = *reinterpret_cast<const float4*>(s.win + 8 * ch + h4);
      const float4 w1 = *reinterpret_cast<const float4*>(s.win + 8 * ch + (h4 ^ 4));
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      float tw[2][2][8];
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int row = 0; row < 2; ++row) {
          const uint4 v = r.raw[q][row][i];
          const unsigned w32[4] = {v.x, v.y, v.z, v.w};
          const float* grow = s.gain + (2 * q + row) * W + 8 * ch;
          const float4 g0 = *reinterpret_cast<const float4*>(grow + h4);
          const float4 g1 = *reinterpret_cast<const float4*>(grow + (h4 ^ 4));
          const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          float qq[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          if constexpr (HAS_SUB) {
            const float* qrow = s.subg + (2 * q + row) * W + 8 * ch;
            const float4 q0 = *reinterpret_cast<const float4*>(qrow + h4);
            const float4 q1 = *reinterpret_cast<const float4*>(qrow + (h4 ^ 4));
            qq[0] = q0.x; qq[1] = q0.y; qq[2] = q0.z; qq[3] = q0.w;
            qq[4] = q1.x; qq[5] = q1.y; qq[6] = q1.z; qq[7] = q1.w;
          }
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const unsigned px = (j & 1) ? (w32[j >> 1] >> 16) : (w32[j >> 1] & 0xffffu);
            const float y = static_cast<float>(px);
            const float tv = HAS_SUB ? fmaf(y, g[j], -qq[j]) : fmaf(y, g[j], -1.0f);
            acc += tv;
            tw[q][row][j] = tv * w[j];
          }
          sums[2 * q + row] += acc;
        }
      float4* dst = reinterpret_cast<float4*>(s.stg) + 8 * ch;
      const int sw = ch & 7;
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j ^ sw] = make_float4(tw[0][0][j], tw[1][0][j], tw[0][1][j], tw[1][1][j]);
    }
  }
}

// gather-lerp for both pairs at once; nm = {-(mean - 1)} lanes: nm_re = rows A of the two pairs, nm_im = rows B
template <class P>
ABC_HD void phase2_gather(int tid, const GroupSmem2& s, Thread2State<P>& r, V2 nm_re, V2 nm_im) {
#pragma unroll
  for (int i = 0; i < P::NB0; ++i) {
    const int b = tid + P::T * i;
    if (P::NB0 * P::T == P::N1 || b < P::N1) {
#pragma unroll
      for (int c4 = 0; c4 < P::R0P4 / 4; ++c4) {
        const uint4 v = *reinterpret_cast<const uint4*>(s.offT + (c4 * P::N1 + b) * 4);
        const unsigned off[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int a = 4 * c4 + j;
          if (a < P::R0) {
            const ulonglong2 y1 = *reinterpret_cast<const ulonglong2*>(s.stg + (off[j] & 0xffffu));
            const ulonglong2 y0 = *reinterpret_cast<const ulonglong2*>(s.stg + (off[j] >> 16));
            const ulonglong2 wv = *reinterpret_cast<const ulonglong2*>(s.wvT + (a * P::N1 + b));
            const V2 wq{wv.x}, vw{wv.y};
            const V2 y1r{y1.x}, y1i{y1.y}, y0r{y0.x}, y0i{y0.y};
            r.x[i][a].x = vfma(nm_re, vw, vfma(wq, vsub(y1r, y0r), y1r));
            r.x[i][a].y = vfma(nm_im, vw, vfma(wq, vsub(y1i, y0i), y1i));
          }
        }
      }
    }
  }
}

// z * (wx + i wy), twiddle lanes duplicated in the table; no negated operand available -> 5 packed instructions
ABC_HD Cx<V2> cxmul_tab(Cx<V2> a, const float4* tw) {
  const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(tw);
  const V2 wx{t.x}, wy{t.y};
  return Cx<V2>{vsub(vmul(a.x, wx), vmul(a.y, wy)), vfma(a.y, wx, vmul(a.x, wy))};
}

template <class P>
ABC_HD void phase2_pass0(int tid, const GroupSmem2& s, Thread2State<P>& r) {
#pragma unroll
  for (int i = 0; i < P::NB0; ++i) {
    const int b = tid + P::T * i;
    if (P::NB0 * P::T == P::N1 || b < P::N1) {
      Cx<V2> out[P::R0];
      DftV<P::R0, kFftSign, 1, 1, V2>::run(r.x[i], out);
      s.bre[b] = out[0].x;
      s.bim[b] = out[0].y;
#pragma unroll
      for (int c = 1; c < P::R0; ++c) {
        const Cx<V2> z = cxmul_tab(out[c], s.tw0 + (c - 1) * P::N1 + b);
        s.bre[P::ROWSTRIDE * c + b] = z.x;
        s.bim[P::ROWSTRIDE * c + b] = z.y;
      }
    }
  }
}

template <class P>
ABC_HD void phase2_pass1(int tid, const GroupSmem2& s) {
  if constexpr (P::THREE) {
#pragma unroll
    for (int i = 0; i < P::NB1; ++i) {
      const int x = tid + P::T * i;
      if (P::NB1 * P::T == P::NBF1 || x < P::NBF1) {
        const int c = x % P::R0, bp = x / P::R0;
        V2* bre = s.bre + P::ROWSTRIDE * c + bp;
        V2* bim = s.bim + P::ROWSTRIDE * c + bp;
        Cx<V2> in[P::R1], out[P::R1];
#pragma unroll
        for (int a = 0; a < P::R1; ++a) in[a] = Cx<V2>{bre[P::N2 * a], bim[P::N2 * a]};
        DftV<P::R1, kFftSign, 1, 1, V2>::run(in, out);
        bre[0] = out[0].x;
        bim[0] = out[0].y;
#pragma unroll
        for (int c1 = 1; c1 < P::R1; ++c1) {
          const Cx<V2> z = cxmul_tab(out[c1], s.tw1 + (c1 - 1) * P::N2 + bp);
          bre[P::N2 * c1] = z.x;
          bim[P::N2 * c1] = z.y;
        }
      }
    }
  }
}

ABC_HD V2 v2_sqrt(V2 a) { return v2_make(fast_sqrt(v2_lo(a)), fast_sqrt(v2_hi(a))); }

template <class P, bool ACCUM>
ABC_HD void phase2_passL(int tid, const GroupSmem2& s, Thread2State<P>& r) {
  constexpr int RL = P::RL, S = P::S, JH = (RL + 1) / 2;
#pragma unroll
  for (int i = 0; i < P::NU; ++i) {
    const int u = tid + P::T * i;
    if (P::NU * P::T == P::NUNITS || u < P::NUNITS) {
      const int kA = u, kB = (u == 0) ? S / 2 : S - u;
      const int oa = P::ROWSTRIDE * (kA % P::R0) + RL * (kA / P::R0);
      const int ob = P::ROWSTRIDE * (kB % P::R0) + RL * (kB / P::R0);
      Cx<V2> za[RL], zb[RL], Za[RL], Zb[RL];
#pragma unroll
      for (int a = 0; a < RL; ++a) {
        za[a] = Cx<V2>{s.bre[oa + a], s.bim[oa + a]};
        zb[a] = Cx<V2>{s.bre[ob + a], s.bim[ob + a]};
      }
      DftV<RL, kFftSign, 1, 1, V2>::run(za, Za);
      DftV<RL, kFftSign, 1, 1, V2>::run(zb, Zb);
#pragma unroll
      for (int j = 0; j < RL; ++j) {
        Cx<V2> Pz, Qz;
        if (u != 0) {
          Pz = Za[j];
          Qz = Zb[RL - 1 - j];
        } else if (j < JH) {
          Pz = Za[j];
          Qz = Za[(RL - j) % RL];
        } else {
          Pz = Zb[j - JH];
          Qz = Zb[RL - 1 - (j - JH)];
        }
        const V2 sr = vadd(Pz.x, Qz.x), di = vsub(Pz.y, Qz.y), si = vadd(Pz.y, Qz.y), dr = vsub(Pz.x, Qz.x);
        const V2 m0 = v2_sqrt(vfma(sr, sr, vmul(di, di))), m1 = v2_sqrt(vfma(si, si, vmul(dr, dr)));
        r.acc[i][j][0] = ACCUM ? vadd(r.acc[i][j][0], m0) : m0;
        r.acc[i][j][1] = ACCUM ? vadd(r.acc[i][j][1], m1) : m1;
      }
    }
  }
}

// dB conversion and stores of both pairs; out[q][row], rows[q] = first camera row of pair q, valid[q][row]
template <class P>
ABC_HD void phase2_finalise(int tid, const ReconArgs& a, float* const (&out)[2][2], const int (&rowa)[2], const bool (&valid)[2][2],
                            Thread2State<P>& r, float (&mn)[2], float (&mx)[2], unsigned long long keep_pol = 0) {
  static_assert(P::S / 2 >= 8, "special bins must all fall into slot 0");
  const V2 scale = v2_make(a.out_scale, a.out_scale), eps = v2_make(1e-5f, 1e-5f);
#pragma unroll
  for (int i = 0; i < P::NU; ++i) {
    const int u = tid + P::T * i;
    if (P::NU * P::T == P::NUNITS || u < P::NUNITS) {
#pragma unroll
      for (int j = 0; j < P::RL; ++j) {
        const int kk = unit_bin<P>(u, j);
#pragma unroll
        for (int row = 0; row < 2; ++row) {
          const V2 v = vfma(r.acc[i][j][row], scale, eps);
          const float db2[2] = {fast_log2(v2_lo(v)) * a.db_scale, fast_log2(v2_hi(v)) * a.db_scale};
          r.acc[i][j][row] = v2_make(0.f, 0.f);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const float db = db2[q];
            float* dst = out[q][row];
            const bool ok = (kk < a.D) && valid[q][row];
            if (j == 0) {
              if (ok && kk >= 2) {
                store_scratch(dst + kk, db, keep_pol);
                if (kk == 4) {  // bscandb.row(4).copyTo(row(1)), row(0): BscanFFT.cpp:1239-1240
                  store_scratch(dst, db, keep_pol);
                  store_scratch(dst + 1, db, keep_pol);
                }
                const bool is55 = a.clamp55 && kk == 5 && (rowa[q] + row) == 5;
                if (!is55) {
                  mn[q] = fminf(mn[q], db);
                  mx[q] = fmaxf(mx[q], db);
                }
              } else if (ok && a.dc01 != nullptr) {
                a.dc01[2 * (size_t)((dst - a.scratch) / a.Dp) + kk] = db;
              }
            } else if (ok) {
              store_scratch(dst + kk, db, keep_pol);
              mn[q] = fminf(mn[q], db);
              mx[q] = fmaxf(mx[q], db);
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------- the kernel
#ifdef __CUDACC__
// A work unit = two consecutive items (row pairs); unit tickets come from the same global counter as in recon_kernel.
template <class P, int G, bool HAS_SUB, bool A1>
__global__ void __launch_bounds__(P::T* G, 1) recon2_kernel(const ReconArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout2 L = make_layout2<P>(a.W, HAS_SUB);
  const int g = threadIdx.x / P::T;
  const int tid = threadIdx.x - g * P::T;
  const int lane = tid & 31, wrp = tid >> 5;
  const GroupSmem2 s = resolve2<P>(smem, L, g);
  const int W = a.W, W8 = W >> 3;
  const SchedView sv = sched_view(a.sched, a.nB);
  const int nunits = (a.nitems + 1) >> 1;

  {
    const int n16 = L.groups >> 4;
    const uint4* src = reinterpret_cast<const uint4*>(a.idxT);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
  }
  // unit u -> items 2u, 2u+1 -> (pair, bscan) each; slot layout {p0, b0, p1, b1}, p < 0 = absent
  auto decode = [&](int u, int* dst4) {
    const int t0 = 2 * u, t1 = 2 * u + 1;
    const int b0 = t0 / a.npairs, b1 = t1 / a.npairs;
    dst4[0] = t0 < a.nitems ? t0 - b0 * a.npairs : -1;
    dst4[1] = b0;
    dst4[2] = t1 < a.nitems ? t1 - b1 * a.npairs : -1;
    dst4[3] = b1;
  };
  if (tid == 0) {
    mbar_init(s.mbar, 1);
    mbar_init(s.mbar + 1, P::NWARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    decode(atomicAdd(sv.ticket, 1), s.slot + 0);
    decode(atomicAdd(sv.ticket, 1), s.slot + 4);
  }
  __syncthreads();

  unsigned long long pol, keep_pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep_pol));

  Thread2State<P> r;
#pragma unroll
  for (int i = 0; i < P::NU; ++i)
#pragma unroll
    for (int j = 0; j < P::RL; ++j) r.acc[i][j][0] = r.acc[i][j][1] = v2_make(0.f, 0.f);

  // camera rows of a unit: {rowA pair0, rowB pair0, rowA pair1, rowB pair1}; an absent second pair repeats the first
  auto unit_rows = [&](int p0, int p1, int (&rows)[4]) {
    const int q1 = p1 >= 0 ? p1 : p0;
    rows[0] = 2 * p0;
    rows[1] = (2 * p0 + 1 < a.oph) ? 2 * p0 + 1 : 2 * p0;
    rows[2] = 2 * q1;
    rows[3] = (2 * q1 + 1 < a.oph) ? 2 * q1 + 1 : 2 * q1;
  };
  auto issue_calibration = [&](int p0, int p1) {  // tid 0
    int rows[4];
    unit_rows(p0, p1, rows);
    const unsigned rowbytes = static_cast<unsigned>(W) * 4u;
    mbar_expect_tx(s.mbar, rowbytes * (HAS_SUB ? 8u : 4u));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      bulk_g2s(s.gain + k * W, a.gain + static_cast<size_t>(rows[k]) * W, rowbytes, s.mbar);
      if constexpr (HAS_SUB) bulk_g2s(s.subg + k * W, a.subg + static_cast<size_t>(rows[k]) * W, rowbytes, s.mbar);
    }
  };
  auto prefetch_rows = [&](int p0, int b0, int p1, int b1, int f) {
    int rows[4];
    unit_rows(p0, p1, rows);
    const int bb1 = p1 >= 0 ? b1 : b0;
    const uint8_t* f0 = a.frames + (static_cast<size_t>(b0) * a.A + f) * a.frame_stride;
    const uint8_t* f1 = a.frames + (static_cast<size_t>(bb1) * a.A + f) * a.frame_stride;
    const uint8_t* const ptr[4] = {f0 + static_cast<size_t>(rows[0]) * a.row_stride, f0 + static_cast<size_t>(rows[1]) * a.row_stride,
                                   f1 + static_cast<size_t>(rows[2]) * a.row_stride, f1 + static_cast<size_t>(rows[3]) * a.row_stride};
    phase2_load<P>(tid, ptr, W8, r, pol);
  };

  int p0 = s.slot[0], b0 = s.slot[1], p1 = s.slot[2], b1 = s.slot[3];
  int np0 = s.slot[4], nb0 = s.slot[5], np1 = s.slot[6], nb1 = s.slot[7];
  if (p0 >= 0) {
    if (tid == 0) issue_calibration(p0, p1);
    prefetch_rows(p0, b0, p1, b1, 0);
  }
  unsigned cal_parity = 0, free_parity = 0;
  bool buf_busy = false;

  const int ngroups = gridDim.x * G;
  const int njobs = a.nB * a.nparts;
  constexpr int kTicketTid = 32 % P::T, kPublishTid = 64 % P::T, kJobTid = 96 % P::T;
  int myjob = blockIdx.x * G + g;
  int myjob_b = myjob / a.nparts;
  constexpr int kPublishBatch = 8;  // entries (two per unit)
  int* const pend = s.slot + 16;
  int* const jobslot = s.slot + 12;
  int npend = 0;
  auto publish = [&]() {
    __threadfence();
    for (int i = 0; i < npend; ++i) atomicAdd(sv.cnt + pend[i], 1);
    npend = 0;
  };
  (void)nunits;

  while (p0 >= 0) {
    int t_next = 0, polled = 0;
    if (tid == kTicketTid) t_next = atomicAdd(sv.ticket, 1);
    if (tid == kJobTid && myjob < njobs) polled = *reinterpret_cast<volatile const int*>(sv.cnt + myjob_b);
    while (!mbar_try_wait(s.mbar, cal_parity)) {
    }
    cal_parity ^= 1u;

    const int nA = A1 ? 1 : a.A;
    for (int f = 0; f < nA; ++f) {
      const bool last = A1 || (f + 1 == nA);
      if constexpr (P::NWARPS > 1) {
        if (buf_busy) {
          while (!mbar_try_wait(s.mbar + 1, free_parity)) {
          }
          free_parity ^= 1u;
        }
      }
      float sums[4];
      phase2_pre<P, HAS_SUB>(tid, s, W, r, sums);
#pragma unroll
      for (int k = 0; k < 4; ++k) sums[k] = warp_sum(sums[k]);
      if constexpr (P::NWARPS > 1) {
        if (lane == 0) *reinterpret_cast<float4*>(s.red + 4 * wrp) = make_float4(sums[0], sums[1], sums[2], sums[3]);
      }
      if (f == 0 && tid == kPublishTid && npend + 2 > kPublishBatch) publish();
      if (!last) {
        prefetch_rows(p0, b0, p1, b1, f + 1);
      } else if (np0 >= 0) {
        prefetch_rows(np0, nb0, np1, nb1, 0);
      }
      group_sync<P::T>(g);
      if (last && np0 >= 0 && tid == 0) issue_calibration(np0, np1);
      if constexpr (P::NWARPS > 1) {
        sums[0] = sums[1] = sums[2] = sums[3] = 0.f;
#pragma unroll
        for (int w = 0; w < P::NWARPS; ++w) {
          const float4 t = *reinterpret_cast<const float4*>(s.red + 4 * w);
          sums[0] += t.x;
          sums[1] += t.y;
          sums[2] += t.z;
          sums[3] += t.w;
        }
      }
      // lanes: lo = pair 0, hi = pair 1; re = rows A (sums[0], sums[2]), im = rows B (sums[1], sums[3])
      const V2 nm_re = v2_make(-sums[0] * a.inv_W, -sums[2] * a.inv_W), nm_im = v2_make(-sums[1] * a.inv_W, -sums[3] * a.inv_W);
      phase2_gather<P>(tid, s, r, nm_re, nm_im);
      group_sync<P::T>(g);
      phase2_pass0<P>(tid, s, r);
      if (last && tid == kTicketTid) decode(t_next, s.slot + 8);
      if (last && tid == kJobTid) {
        int job = -1;
        if (myjob < njobs && polled >= a.npairs) {
          __threadfence();
          job = myjob;
          myjob += ngroups;
          myjob_b = myjob / a.nparts;
        }
        *jobslot = job;
      }
      group_sync<P::T>(g);
      if constexpr (P::THREE) {
        phase2_pass1<P>(tid, s);
        group_sync<P::T>(g);
      }
      phase2_passL<P, !A1>(tid, s, r);
      if constexpr (P::NWARPS > 1) {
        __syncwarp();
        if (lane == 0) mbar_arrive(s.mbar + 1);
        buf_busy = true;
      }
    }

    {
      const bool have1 = p1 >= 0;
      const int q1 = have1 ? p1 : p0, bb1 = have1 ? b1 : b0;
      const int rowa[2] = {2 * p0, 2 * q1};
      const bool valid[2][2] = {{true, 2 * p0 + 1 < a.oph}, {have1, have1 && (2 * q1 + 1 < a.oph)}};
      float* const o00 = a.scratch + (static_cast<size_t>(b0) * a.oph + rowa[0]) * a.Dp;
      float* const o10 = a.scratch + (static_cast<size_t>(bb1) * a.oph + rowa[1]) * a.Dp;
      float* const out[2][2] = {{o00, o00 + a.Dp}, {o10, o10 + a.Dp}};
      float mn[2] = {__int_as_float(0x7f800000), __int_as_float(0x7f800000)};
      float mx[2] = {__int_as_float(0xff800000), __int_as_float(0xff800000)};
      phase2_finalise<P>(tid, a, out, rowa, valid, r, mn, mx, keep_pol);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        mn[q] = warp_min(mn[q]);
        mx[q] = warp_max(mx[q]);
      }
      if (lane == 0) {
        if (mn[0] <= mx[0]) {
          atomicMin(sv.minv + b0, float_to_ordered(fmaxf(mn[0], a.thr)));
          atomicMax(sv.maxv + b0, float_to_ordered(fmaxf(mx[0], a.thr)));
        }
        if (have1 && mn[1] <= mx[1]) {
          atomicMin(sv.minv + bb1, float_to_ordered(fmaxf(mn[1], a.thr)));
          atomicMax(sv.maxv + bb1, float_to_ordered(fmaxf(mx[1], a.thr)));
        }
      }
      if (tid == kPublishTid) {
        pend[npend++] = b0;
        if (have1) pend[npend++] = b1;
      }
      const int job = *jobslot;
      if (job >= 0) {
        group_sync<P::T>(g);
        normalise_part<P>(a, sv, job / a.nparts, job % a.nparts, g, tid, s.stg);
      }
    }
    p0 = np0; b0 = nb0; p1 = np1; b1 = nb1;
    np0 = s.slot[8]; nb0 = s.slot[9]; np1 = s.slot[10]; nb1 = s.slot[11];
  }

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
      *jobslot = job;
    }
    group_sync<P::T>(g);
    const int job = *jobslot;
    if (job < 0) break;
    normalise_part<P>(a, sv, job / a.nparts, job % a.nparts, g, tid, s.stg);
    group_sync<P::T>(g);
  }
}
#endif  // __CUDACC__

}  // namespace abcoct
