// Lock-step host emulation of the warp-per-A-scan kernel (fdoct_b200/csrc/wrow_kernel.cuh).
//
// The kernel body is compiled here as a __host__ __device__ function (ABC_WROW_HOST_EMU) and executed by 32 host threads per
// warp: shuffles, warp reductions, __syncwarp / __syncthreads and the global atomics are emulated with barriers and GCC
// atomics, everything else (table layouts, index algebra of both FFT passes, the lane pairing of the split step, the DC-row /
// clampupper special cases, the ticket scheduler, the completion protocol and the normalisation jobs) is the product code.
// The result is compared with a naive double-precision restatement of the block
//   (y * gain - subg) -> mean removal -> window -> y[i] + F[i] (y[i] - y[i-1]) -> gather -> unscaled inverse DFT -> |.| ->
//   average -> dB -> DC-row mask -> threshold -> min-max normalise -> u8
// (BscanFFT.cpp:1125-1255 with the source-indexed weight of :1170).  Test infrastructure only; built and run by
// tests/test_native_host.py.  No GPU needed.
#define ABC_WROW_HOST_EMU 1
#include <algorithm>
#include <cmath>
#include <complex>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <random>
#include <thread>
#include <vector>

#include "../../fdoct_b200/csrc/wrow_kernel.cuh"
#include "../../fdoct_b200/csrc/wres_kernel.cuh"

// ------------------------------------------------------------------------------------------------ the emulator
namespace wemu {
struct Barrier {
  std::mutex m;
  std::condition_variable cv;
  int count = 0, gen = 0, n = 0;
  void wait() {
    std::unique_lock<std::mutex> l(m);
    const int g = gen;
    if (++count == n) {
      count = 0;
      ++gen;
      cv.notify_all();
    } else {
      cv.wait(l, [&] { return gen != g; });
    }
  }
};
struct Warp {
  Barrier bar;
  unsigned slot[32];
};
struct Cta {
  Barrier bar;
  unsigned char* smem = nullptr;
  size_t smem_bytes = 0;
  std::vector<unsigned> tmem = std::vector<unsigned>(128 * 512, 0xdeadbeefu);  // tensor memory of the SM: [lane][column]
};
struct Ctx {
  int lane, warp, cta, ncta, nw;
  Warp* w;
  Cta* c;
};
thread_local Ctx* tl = nullptr;

int lane() { return tl->lane; }
int warp_in_cta() { return tl->warp; }
int cta() { return tl->cta; }
int ncta() { return tl->ncta; }
int nwarps() { return tl->nw; }
unsigned shfl_u32(unsigned v, int src) {
  tl->w->slot[tl->lane] = v;
  tl->w->bar.wait();
  const unsigned r = tl->w->slot[src & 31];
  tl->w->bar.wait();
  return r;
}
void syncwarp() { tl->w->bar.wait(); }
void syncthreads() { tl->c->bar.wait(); }
int redux_min(int v) {
  tl->w->slot[tl->lane] = (unsigned)v;
  tl->w->bar.wait();
  int r = (int)tl->w->slot[0];
  for (int i = 1; i < 32; ++i) r = std::min(r, (int)tl->w->slot[i]);
  tl->w->bar.wait();
  return r;
}
int redux_max(int v) {
  tl->w->slot[tl->lane] = (unsigned)v;
  tl->w->bar.wait();
  int r = (int)tl->w->slot[0];
  for (int i = 1; i < 32; ++i) r = std::max(r, (int)tl->w->slot[i]);
  tl->w->bar.wait();
  return r;
}
int atomic_add(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
void atomic_min(int* p, int v) {
  int cur = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v < cur && !__atomic_compare_exchange_n(p, &cur, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
}
void atomic_max(int* p, int v) {
  int cur = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v > cur && !__atomic_compare_exchange_n(p, &cur, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
}
int load_acquire(const int* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
void check_smem(const void*, int, int) {}
void backoff() { std::this_thread::yield(); }
unsigned* tmem() { return tl->c->tmem.data(); }
unsigned ballot(int pred) {
  tl->w->slot[tl->lane] = (unsigned)pred;
  tl->w->bar.wait();
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= (tl->w->slot[i] ? 1u : 0u) << i;
  tl->w->bar.wait();
  return r;
}
}  // namespace wemu

using namespace abcoct;

template <class T, class = void>
struct is_resident : std::false_type {};
template <class T>
struct is_resident<T, std::void_t<decltype(T::TR)>> : std::true_type {};

static int g_ring = 0;  // ReconArgs::ringB of the next case (scratch ring of the warp-per-A-scan kernel)
template <class WP, bool HAS_SUB, bool A1, bool FULLD>
static void run_grid(const ReconArgs& a, int ncta) {
  std::vector<wemu::Cta> ctas(ncta);
  std::vector<wemu::Warp> warps((size_t)ncta * WP::NW);
  std::vector<std::vector<unsigned char>> smem(ncta);
  for (int c = 0; c < ncta; ++c) {
    smem[c].assign(WP::SMEM_BYTES + 64, 0xcd);
    unsigned char* base = smem[c].data();
    while (reinterpret_cast<uintptr_t>(base) & 15) ++base;
    ctas[c].smem = base;
    ctas[c].bar.n = WP::NW * 32;
  }
  for (auto& w : warps) w.bar.n = 32;
  std::vector<wemu::Ctx> ctx((size_t)ncta * WP::NW * 32);
  std::vector<std::thread> th;
  for (int c = 0; c < ncta; ++c)
    for (int w = 0; w < WP::NW; ++w)
      for (int l = 0; l < 32; ++l) {
        wemu::Ctx& x = ctx[((size_t)c * WP::NW + w) * 32 + l];
        x = wemu::Ctx{l, w, c, ncta, WP::NW, &warps[(size_t)c * WP::NW + w], &ctas[c]};
        th.emplace_back([&a, &x] {
          wemu::tl = &x;
          if constexpr (is_resident<WP>::value)
            wres_body<WP, HAS_SUB, A1, FULLD>(a, x.c->smem);
          else
            wrow_body<WP, HAS_SUB, A1, FULLD>(a, x.c->smem);
        });
      }
  for (auto& t : th) t.join();
}

struct Case {
  int W, oph, D, A, nB, nsplit, ncta;
  bool clamp, want_db, want_dc;
};

template <class WP, bool HAS_SUB>
static double run_case(const Case& cs, unsigned seed) {
  const int N = WP::N, W = cs.W, oph = cs.oph, D = cs.D, A = cs.A, nB = cs.nB;
  std::mt19937 rng(seed);
  std::uniform_int_distribution<int> pix(0, 4000);
  std::uniform_real_distribution<double> ud(0.0, 1.0);
  // a monotone non-increasing gather like the real lambda -> k map (slope about -W/N with a slow drift), source-indexed weights
  std::vector<int> idx(N);
  std::vector<double> frac(N), win(W);
  for (int q = 0; q < N; ++q) {
    int i = (int)((double)(N - 1 - q) * (W - 1) / (N - 1) + 0.5 + 6.0 * std::sin(q * 0.004));
    idx[q] = std::min(std::max(i, 1), W - 1);
  }
  idx[0] = idx[N - 1] = -1;
  for (int i = 0; i < N; ++i) frac[i] = 0.001 + 0.999 * ud(rng);
  for (int i = 0; i < W; ++i) win[i] = 0.62 - 0.48 * std::fabs((double)i / (W - 1) - 0.5) + 0.38 * std::cos(6.283185307179586 * ((double)i / (W - 1) - 0.5));
  std::vector<unsigned char> blob(WP::TABLE_BYTES, 0);
  WrowTablesHost wt{W, idx.data(), frac.data(), win.data()};
  wrow_build_blob<WP>(wt, blob.data());

  std::vector<float> gain((size_t)oph * W), subg((size_t)oph * W);
  for (size_t i = 0; i < gain.size(); ++i) {
    gain[i] = (float)(1.0 / (20000.0 + 10000.0 * ud(rng)));
    subg[i] = HAS_SUB ? (float)((64.0 + 8.0 * ud(rng)) * gain[i] + 1.0) : 1.f;  // the kernel stages t - 1 (the +1 lives in subg)
  }
  std::vector<float> pg((size_t)oph * WP::WMAX), ps((size_t)oph * WP::WMAX);
  for (int r = 0; r < oph; ++r) {
    wrow_permute_cal_row<WP>(&gain[(size_t)r * W], W, &pg[(size_t)r * WP::WMAX]);
    wrow_permute_cal_row<WP>(&subg[(size_t)r * W], W, &ps[(size_t)r * WP::WMAX]);
  }
  const size_t row_stride = ((size_t)W * 2 + 15) / 16 * 16 + 32, frame_stride = row_stride * oph;
  std::vector<uint8_t> frames((size_t)nB * A * frame_stride + 64, 0);
  uint8_t* fbase = frames.data();
  while (reinterpret_cast<uintptr_t>(fbase) & 15) ++fbase;
  auto px = [&](int b, int f, int r, int i) -> uint16_t& {
    return *reinterpret_cast<uint16_t*>(fbase + ((size_t)b * A + f) * frame_stride + (size_t)r * row_stride + 2 * (size_t)i);
  };
  for (int b = 0; b < nB; ++b)
    for (int f = 0; f < A; ++f)
      for (int r = 0; r < oph; ++r)
        for (int i = 0; i < W; ++i) {
          const double ph = 0.02 * i * (1 + (r % 7)) + 0.3 * f + 0.11 * b, ph2 = 0.31 * i + 0.05 * r;
          px(b, f, r, i) = (uint16_t)(pix(rng) + 25000 + 11000 * std::sin(ph) + 3000 * std::cos(ph2));
        }

  const int Dp = (D + 31) / 32 * 32;
  std::vector<float> scratch((size_t)nB * oph * Dp, -777.f), outdb((size_t)nB * D * oph, -555.f), dc01((size_t)nB * oph * 2, -333.f);
  std::vector<uint8_t> out8((size_t)nB * D * oph + 16, 0x5a);
  std::vector<int> sched(sched_ints(nB));
  {
    SchedView v = sched_view(sched.data(), nB);
    for (int i = 0; i < kSchedHeader; ++i) sched[i] = 0;
    for (int b = 0; b < nB; ++b) {
      v.minv[b] = float_to_ordered(w_inf(false));
      v.maxv[b] = float_to_ordered(w_inf(true));
      v.cnt[b] = 0;
    }
  }
  ReconArgs a{};
  a.frames = fbase;
  a.frame_stride = frame_stride;
  a.row_stride = row_stride;
  a.W = W;
  a.oph = oph;
  a.D = D;
  a.Dp = Dp;
  a.A = A;
  a.nB = nB;
  a.nitems = nB * oph;
  a.nparts = (oph + 31) / 32;
  a.nsplit = cs.nsplit;
  a.calpitch = WP::WMAX;
  a.gain = pg.data();
  a.subg = ps.data();
  a.idxT = reinterpret_cast<const uint32_t*>(blob.data());
  a.scratch = scratch.data();
  a.ringB = g_ring;  // 0: one scratch region per B-scan
  a.sched = sched.data();
  uint8_t* o8 = out8.data();
  while (reinterpret_cast<uintptr_t>(o8) & 3) ++o8;
  a.out8 = o8;
  a.outdb = cs.want_db ? outdb.data() : nullptr;
  a.dc01 = cs.want_dc ? dc01.data() : nullptr;
  a.inv_W = 1.0f / W;
  a.out_scale = 0.5f / A;
  a.db_scale = (float)(0.6931471805599453 * 20.0 * (1.0 / 2.303));
  a.thr = -30.f;
  a.clamp_db = 50.f;
  a.clamp55 = cs.clamp ? 1 : 0;
  if (A == 1) {
    if (D == N / 2)
      run_grid<WP, HAS_SUB, true, true>(a, cs.ncta);
    else
      run_grid<WP, HAS_SUB, true, false>(a, cs.ncta);
  } else {
    if (D == N / 2)
      run_grid<WP, HAS_SUB, false, true>(a, cs.ncta);
    else
      run_grid<WP, HAS_SUB, false, false>(a, cs.ncta);
  }

  // ---- double reference
  std::vector<std::complex<double>> tw(N);
  for (int k = 0; k < N; ++k) tw[k] = std::polar(1.0, 2.0 * M_PI * k / N);
  double worst_db = 0;  // worst |mag - ref| / max(ref, 1e-3 * A-scan max), the tolerance of the GPU parity tests (1e-4)
  const double kDb = 20.0 * (1.0 / 2.303);
  auto mag_of = [&](double dbv) { return std::exp(dbv / kDb) - 1e-5; };
  int worst_lsb = 0;
  long ndiff = 0;
  for (int b = 0; b < nB; ++b) {
    std::vector<double> db((size_t)D * oph), raw01((size_t)oph * 2), rowmax(oph, 0.0);
    for (int r = 0; r < oph; ++r) {
      std::vector<double> accm(D, 0.0);
      for (int f = 0; f < A; ++f) {
        std::vector<double> t(W), y(W), v(W, 0.0), ylin(N, 0.0);
        double mean = 0;
        for (int i = 0; i < W; ++i) {
          t[i] = (double)px(b, f, r, i) * gain[(size_t)r * W + i] - ((double)subg[(size_t)r * W + i] - 1.0);
          mean += t[i];
        }
        mean /= W;
        for (int i = 0; i < W; ++i) y[i] = (t[i] - mean) * win[i];
        for (int i = 1; i < W; ++i) v[i] = y[i] + frac[i] * (y[i] - y[i - 1]);
        for (int q = 1; q < N - 1; ++q) ylin[q] = v[idx[q]];
        for (int k = 0; k < D; ++k) {
          std::complex<double> acc = 0;
          for (int q = 0; q < N; ++q) acc += ylin[q] * tw[((long long)q * k) % N];
          accm[k] += std::abs(acc);
        }
      }
      for (int k = 0; k < D; ++k) {
        db[(size_t)k * oph + r] = std::log(accm[k] / A + 1e-5) * kDb;
        rowmax[r] = std::max(rowmax[r], accm[k] / A);
      }
      raw01[2 * r] = db[r];
      raw01[2 * r + 1] = db[(size_t)oph + r];
      db[r] = db[(size_t)4 * oph + r];
      db[(size_t)oph + r] = db[(size_t)4 * oph + r];
    }
    std::vector<double> disp(db);
    for (double& x : disp) x = std::max(x, -30.0);
    if (cs.clamp) disp[(size_t)5 * oph + 5] = 50.0;
    const double mn = *std::min_element(disp.begin(), disp.end()), mx = *std::max_element(disp.begin(), disp.end());
    const double sc = (mx - mn) > 2.220446049250313e-16 ? 1.0 / (mx - mn) : 0.0;
    for (int k = 0; k < D; ++k)
      for (int r = 0; r < oph; ++r) {
        const double ref = std::nearbyint((disp[(size_t)k * oph + r] - mn) * sc * 255.0);
        const int got = a.out8[((size_t)b * D + k) * oph + r];
        const int dl = std::abs(got - (int)ref);
        worst_lsb = std::max(worst_lsb, dl);
        ndiff += dl != 0;
        if (cs.want_db) {
          const double rm = mag_of(db[(size_t)k * oph + r]), gm = mag_of((double)outdb[((size_t)b * D + k) * oph + r]);
          worst_db = std::max(worst_db, std::fabs(gm - rm) / std::max(rm, 1e-3 * rowmax[r]));
        }
      }
    if (cs.want_dc)
      for (int r = 0; r < oph; ++r)
        for (int k = 0; k < 2; ++k) {
          const double rm = mag_of(raw01[2 * r + k]), gm = mag_of((double)dc01[2 * ((size_t)b * oph + r) + k]);
          worst_db = std::max(worst_db, std::fabs(gm - rm) / std::max(rm, 1e-3 * rowmax[r]));
        }
  }
  std::printf("N=%d R=%d W=%d oph=%d D=%d A=%d nB=%d sub=%d clamp=%d nsplit=%d  max rel mag err=%.3g  display: worst %d LSB, %.4f%% differ\n", N,
              WP::R, W, oph, D, A, nB, (int)HAS_SUB, (int)cs.clamp, cs.nsplit, worst_db, worst_lsb, 100.0 * ndiff / ((double)nB * D * oph));
  std::fflush(stdout);
  return std::max(worst_db / 1e-4, worst_lsb > 1 ? 10.0 : 0.0);
}

int main(int argc, char** argv) {
  const bool quick = argc > 1 && std::string(argv[1]) == "quick";
  double worst = 0;
  //                                         W    oph  D    A nB nsplit ncta clamp db    dc
  worst = std::max(worst, run_case<WPlan<2048, 2, 1>, false>(Case{2048, 8, 1024, 1, 1, 1, 1, false, true, true}, 1));
  worst = std::max(worst, run_case<WPlan<1280, 2, 2>, true>(Case{1280, 9, 500, 2, 2, 3, 2, true, true, false}, 2));
  if (!quick) {
    worst = std::max(worst, run_case<WPlan<2048, 2>, true>(Case{1920, 37, 700, 2, 2, 3, 2, true, true, true}, 3));
    worst = std::max(worst, run_case<WPlan<1920, 2, 1>, false>(Case{1920, 12, 960, 1, 2, 2, 2, false, false, false}, 4));
    worst = std::max(worst, run_case<WPlan<1024, 2, 2>, false>(Case{1000, 33, 512, 3, 1, 1, 2, true, true, false}, 5));
    worst = std::max(worst, run_case<WPlan<1280, 2>, false>(Case{1280, 40, 640, 1, 3, 4, 3, false, true, false}, 6));
    // whole 32-row parts, whole tiles, no dB image: the straight-line path of the normalisation jobs (with the forced element)
    worst = std::max(worst, run_case<WPlan<1280, 2>, false>(Case{1280, 64, 640, 1, 2, 5, 2, true, false, false}, 7));
    // three worker warps per service warp, more B-scans than CTAs: mailboxes, one fence for several workers, job order
    worst = std::max(worst, run_case<WPlan<1280, 4>, false>(Case{1280, 36, 640, 1, 5, 2, 2, false, false, false}, 8));
  }
  if (!quick) {  // the dB scratch as a ring of two B-scans: slot reuse, the writer's guard on the finished-job count of B-scan b - 2
    g_ring = 2;
    worst = std::max(worst, run_case<WPlan<1280, 4>, false>(Case{1280, 36, 640, 1, 7, 2, 2, false, true, false}, 9));
    worst = std::max(worst, run_case<WPlan<2048, 2, 1>, true>(Case{2048, 8, 1024, 2, 5, 1, 1, true, true, true}, 10));
    g_ring = 0;
  }
  // ---- resident-row kernel (wres_kernel.cuh): teams of 4 warps, dB rows parked in (emulated) tensor memory, static schedule
  //   one team, several rounds per team (slot reuse), partial last block (oph % 4 != 0), forced element, dB image, DC rows
  worst = std::max(worst, run_case<RPlan<1280, 4>, true>(Case{1280, 9, 500, 2, 2, 1, 2, true, true, true}, 11));
  if (!quick) {
    worst = std::max(worst, run_case<RPlan<2048, 4>, false>(Case{2048, 8, 1024, 1, 3, 1, 1, false, true, true}, 12));   // 16 slots per warp
    worst = std::max(worst, run_case<RPlan<1920, 8>, false>(Case{1920, 12, 960, 1, 4, 1, 2, false, false, false}, 13));       // two teams per CTA, word stores
    worst = std::max(worst, run_case<RPlan<1280, 16>, false>(Case{1280, 38, 640, 1, 3, 1, 1, true, true, false}, 14));         // four teams per CTA (4 slots per warp), byte stores
    worst = std::max(worst, run_case<RPlan<1024, 4>, false>(Case{1000, 32, 512, 3, 2, 1, 4, true, false, false}, 15));        // averages, R = 16
  }
  std::printf("worst (in units of the tolerance) = %.3f\n", worst);
  return worst <= 1.0 ? 0 : 1;
}
