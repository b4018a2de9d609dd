// Host emulation of one thread group of recon_kernel: the __host__ __device__ phase functions are run
// thread-by-thread on the CPU and compared with a naive double-precision restatement of
// (y*gain - subg) -> mean removal -> window -> gather-lerp -> unscaled inverse DFT -> |.| -> average -> dB.
// This validates table layout, gather, pass/exchange index algebra, the two-for-one split and the output
// bin mapping of every compiled plan without a GPU.  Built and run by tests/test_native_host.py.
#include <array>
#include <complex>
#include <cstdio>
#include <random>
#include <vector>

#include "../../fdoct_b200/csrc/plan_registry.cuh"

using namespace abcoct;

template <class P, bool HAS_SUB>
static double run_plan(int W, int D, int A, unsigned seed) {
  const int N = P::N, T = P::T;
  std::mt19937 rng(seed);
  std::uniform_int_distribution<int> pix(1000, 60000);
  std::uniform_real_distribution<float> uf(0.f, 1.f);
  // tables: a monotone non-increasing gather like the real one, weights in (0, 1]
  std::vector<int> idx(N);
  std::vector<float> wq(N), win(W);
  for (int q = 0; q < N; ++q) {
    int i = int((double)(N - 1 - q) * (W - 1) / (N - 1) + 0.5) + (int)(3 * std::sin(q * 0.01));
    i = std::min(std::max(i, 1), W - 1);
    idx[q] = i;
    wq[q] = 0.001f + 0.999f * uf(rng);
  }
  idx[0] = W; wq[0] = 0.f; idx[N - 1] = W; wq[N - 1] = 0.f;
  for (int i = 0; i < W; ++i) win[i] = 0.62f - 0.48f * std::fabs(float(i) / (W - 1) - 0.5f) + 0.38f * std::cos(6.2831853f * (float(i) / (W - 1) - 0.5f));
  std::vector<unsigned char> blob;
  build_blob_fn<P>(W, idx.data(), wq.data(), win.data(), blob);

  const SmemLayout L = make_layout<P>(W, HAS_SUB);
  std::vector<unsigned char> smem(L.total(1) + 64, 0);
  unsigned char* base = smem.data();
  while (reinterpret_cast<uintptr_t>(base) & 15) ++base;
  memcpy(base, blob.data(), blob.size());
  GroupSmem s = resolve<P>(base, L, 0);
  std::vector<float> gain(2 * W), subg(2 * W);
  for (int i = 0; i < 2 * W; ++i) {
    gain[i] = 1.0f / (20000.f + 10000.f * uf(rng));
    subg[i] = HAS_SUB ? (64.f + 8.f * uf(rng)) * gain[i] + 1.f : 1.f;  // the kernel stages t - 1 (the +1 lives in subg)
  }
  for (int row = 0; row < 2; ++row) {  // calibration rows travel in the bank-conflict-free layout (cal_phys)
    cal_swizzle_row(gain.data() + row * W, s.gain + row * W, W);
    if (HAS_SUB) cal_swizzle_row(subg.data() + row * W, s.subg + row * W, W);
  }

  std::vector<std::vector<uint16_t>> frames(A, std::vector<uint16_t>(2 * W));
  for (auto& f : frames)
    for (int i = 0; i < 2 * W; ++i) {
      double ph = 0.05 * (i % W) * (1 + (i / W)) + 0.3 * (&f - &frames[0]);
      f[i] = uint16_t(pix(rng) / 8 + 25000 + 12000 * std::sin(ph));
    }

  ReconArgs a{};
  a.W = W; a.oph = 2; a.D = D; a.Dp = (D + 31) / 32 * 32; a.A = A; a.nB = 1; a.npairs = 1; a.nitems = 1;
  a.inv_W = 1.0f / W; a.out_scale = 0.5f / A; a.db_scale = float(0.6931471805599453 * 20.0 * (1.0 / 2.303));
  a.thr = -30.f; a.clamp55 = 0;
  std::vector<ThreadState<P>> st(T);
  for (auto& r : st) memset(&r, 0, sizeof(r));
  for (int f = 0; f < A; ++f) {
    const uint8_t* ra = reinterpret_cast<const uint8_t*>(frames[f].data());
    const uint8_t* rb = ra + 2 * W;
    std::vector<float> sa(T), sb(T);
    for (int t = 0; t < T; ++t) phase_load<P>(t, ra, rb, W / 8, st[t]);
    for (int t = 0; t < T; ++t) phase_pre<P, HAS_SUB>(t, s, W, st[t], sa[t], sb[t]);
    float ta = 0, tb = 0;
    for (int t = 0; t < T; ++t) { ta += sa[t]; tb += sb[t]; }
    for (int t = 0; t < T; ++t) phase_gather<P>(t, s, st[t], ta * a.inv_W, tb * a.inv_W);
    for (int t = 0; t < T; ++t) phase_pass0<P>(t, s, st[t]);
    for (int t = 0; t < T; ++t) phase_pass1<P>(t, s);
    for (int t = 0; t < T; ++t) phase_passL<P>(t, s, st[t]);
  }
  std::vector<float> oa(D, -999.f), ob(D, -999.f);
  float mn = 1e30f, mx = -1e30f;
  for (int t = 0; t < T; ++t) phase_finalise<P>(t, a, oa.data(), ob.data(), 0, true, st[t], mn, mx);
  mn = std::fmax(mn, a.thr);
  mx = std::fmax(mx, a.thr);

  // ---- double reference
  std::vector<std::vector<double>> accd(2, std::vector<double>(N / 2, 0.0));
  for (int f = 0; f < A; ++f)
    for (int row = 0; row < 2; ++row) {
      std::vector<double> t(W), y(W), ylin(N, 0.0);
      double mean = 0;
      for (int i = 0; i < W; ++i) {
        t[i] = (double)frames[f][row * W + i] * gain[row * W + i] - (subg[row * W + i] - 1.0);
        mean += t[i];
      }
      mean /= W;
      for (int i = 0; i < W; ++i) y[i] = (t[i] - mean) * win[i];
      for (int q = 1; q < N - 1; ++q) ylin[q] = y[idx[q]] + (double)wq[q] * (y[idx[q]] - y[idx[q] - 1]);
      for (int k = 0; k < D; ++k) {
        std::complex<double> acc = 0;
        for (int q = 0; q < N; ++q) acc += ylin[q] * std::polar(1.0, 2.0 * M_PI * double(((long long)q * k) % N) / N);
        accd[row][k] += std::abs(acc);
      }
    }
  double worst = 0, wmx = -1e30, wmn = 1e30;
  for (int row = 0; row < 2; ++row) {
    std::vector<double> db(D);
    for (int k = 0; k < D; ++k) db[k] = std::log(accd[row][k] / A + 1e-5) * (20.0 * (1.0 / 2.303));
    db[0] = db[4]; db[1] = db[4];
    for (int k = 0; k < D; ++k) {
      const float got = row ? ob[k] : oa[k];
      worst = std::fmax(worst, std::fabs(got - db[k]));
      wmx = std::fmax(wmx, std::fmax(db[k], -30.0)); wmn = std::fmin(wmn, std::fmax(db[k], -30.0));
    }
  }
  worst = std::fmax(worst, std::fabs(wmx - mx));
  worst = std::fmax(worst, std::fabs(wmn - mn));
  std::printf("N=%d T=%d radices=(%d,%d,%d) W=%d D=%d A=%d sub=%d  max|dB err|=%.3g  (min %.3f max %.3f)\n", N, T, P::R0, P::R1, P::RL, W, D,
              A, (int)HAS_SUB, worst, mn, mx);
  return worst;
}

int main(int argc, char** argv) {
  const bool quick = argc > 1;
  double w = 0;
  w = std::fmax(w, run_plan<P128, false>(128, 64, 1, 1));
  w = std::fmax(w, run_plan<P128, true>(96, 40, 2, 2));
  w = std::fmax(w, run_plan<P256, false>(256, 128, 1, 3));
  w = std::fmax(w, run_plan<P512, true>(512, 256, 2, 4));
  w = std::fmax(w, run_plan<P640, false>(640, 320, 1, 5));
  w = std::fmax(w, run_plan<P1024, false>(1024, 512, 2, 6));
  w = std::fmax(w, run_plan<P1280, true>(1280, 640, 2, 7));
  if (!quick) {
    w = std::fmax(w, run_plan<P1920, false>(1920, 960, 1, 8));
    w = std::fmax(w, run_plan<P2048, false>(2048, 1024, 1, 9));
    w = std::fmax(w, run_plan<P2048, true>(1280, 640, 2, 10));
    w = std::fmax(w, run_plan<P2560, false>(2560, 320, 1, 11));
    w = std::fmax(w, run_plan<P2880, false>(2880, 360, 1, 12));
    w = std::fmax(w, run_plan<P3840, false>(3840, 1024, 1, 13));
    w = std::fmax(w, run_plan<P4096, true>(4096, 2048, 1, 14));
  }
  std::printf("worst=%.3g dB\n", w);
  return w < 2e-3 ? 0 : 1;
}
