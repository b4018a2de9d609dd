// Host-side check of the in-register DFT templates (they are __host__ __device__): every radix the plans use,
// both directions, against a naive double-precision DFT.  Built and run by tests/test_native_host.py.
#include <cstdio>
#include <cmath>
#include <complex>
#include "../../fdoct_b200/csrc/fft_regs.cuh"
using namespace abcoct;
template <int R, int SGN>
static double check() {
  float2 z[R], o[R];
  for (int i = 0; i < R; ++i) z[i] = make_float2(std::sin(1.0f + 0.37f * i * i), std::cos(0.11f * i + 2.0f));
  Dft<R, SGN, 1, 1>::run(z, o);
  double worst = 0, scale = 0;
  for (int c = 0; c < R; ++c) {
    std::complex<double> acc = 0;
    for (int a = 0; a < R; ++a)
      acc += std::complex<double>(z[a].x, z[a].y) * std::polar(1.0, SGN * 2.0 * M_PI * double((a * c) % R) / R);
    worst = std::fmax(worst, std::abs(acc - std::complex<double>(o[c].x, o[c].y)));
    scale = std::fmax(scale, std::abs(acc));
  }
  double rel = worst / scale;
  std::printf("R=%d sgn=%d rel=%.3g\n", R, SGN, rel);
  return rel;
}
template <int R>
static double both() { return std::fmax(check<R, 1>(), check<R, -1>()); }
int main() {
  double w = 0;
  w = std::fmax(w, both<2>()); w = std::fmax(w, both<3>()); w = std::fmax(w, both<4>()); w = std::fmax(w, both<5>());
  w = std::fmax(w, both<6>()); w = std::fmax(w, both<8>()); w = std::fmax(w, both<9>()); w = std::fmax(w, both<10>());
  w = std::fmax(w, both<12>()); w = std::fmax(w, both<15>()); w = std::fmax(w, both<16>()); w = std::fmax(w, both<20>());
  w = std::fmax(w, both<30>()); w = std::fmax(w, both<32>()); w = std::fmax(w, both<40>());
  // trig table spot checks
  for (int den : {5, 7, 16, 20, 2048, 3840}) for (int n = 0; n < den; n += (den > 64 ? 97 : 1)) {
    CtCS cs = ct_cossin(n, den);
    w = std::fmax(w, 1e-7 * (std::fabs(cs.c - std::cos(2 * M_PI * n / den)) + std::fabs(cs.s - std::sin(2 * M_PI * n / den))) / 1e-15);
  }
  std::printf("worst=%.3g\n", w);
  return w < 2e-6 ? 0 : 1;
}
