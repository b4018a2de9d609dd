// FFT plan shapes shared by the host table builder and the device kernel.
//
// A plan factors the row FFT length N = R0 * R1 * RL (R1 == 1 -> two passes) and fixes the number of
// threads T that cooperate on one packed pair of A-scans.  The transform is an in-place decimation-in-
// frequency Stockham-free scheme (each butterfly writes back to the shared-memory slots it read, so only
// one barrier per exchange is needed):
//   pass 0 : butterfly b  in [0,N1)      elements at N1*a + b,            twiddle w_N ^(b*c)
//   pass 1 : butterfly (c,b')            elements at N1*c + N2*a' + b',   twiddle w_N1^(b'*c')
//   pass L : butterfly k0 = c + R0*c'    elements at N1*c + RL*c' + a'',  output bin k0 + S*c''
// with N1 = N/R0, N2 = N1/R1 (== RL), S = N/RL.  tools/fft_plan_model.py checks this algebra in NumPy.
#pragma once

namespace abcoct {

#ifdef __CUDACC__
#define ABC_CX __host__ __device__ constexpr
#else
#define ABC_CX constexpr
#endif
ABC_CX int ceil_div(int a, int b) { return (a + b - 1) / b; }
ABC_CX int cmax(int a, int b) { return a > b ? a : b; }

// MAXT_: thread budget of one CTA (sets the register budget: 65536 / MAXT_ per thread)
#ifndef ABC_DEFAULT_MAXT
#define ABC_DEFAULT_MAXT 512
#endif
template <int N_, int T_, int R0_, int R1_, int RL_, int MAXT_ = ABC_DEFAULT_MAXT>
struct Plan {
  static constexpr int N = N_, T = T_, R0 = R0_, R1 = R1_, RL = RL_, MAXT = MAXT_;
  static constexpr bool THREE = (R1_ > 1);
  static constexpr int N1 = N / R0;
  static constexpr int N2 = THREE ? N1 / R1 : N1;
  static constexpr int S = N / RL;
  static constexpr int PAD = 1;                 // float2 slots of padding per pass-0 row (bank spreading)
  static constexpr int ROWSTRIDE = N1 + PAD;
  static constexpr int BUF = R0 * ROWSTRIDE;    // float2 slots of the exchange buffer
  static constexpr int NB0 = ceil_div(N1, T);   // pass-0 butterflies per thread
  static constexpr int NBF1 = THREE ? R0 * N2 : 0;
  static constexpr int NB1 = THREE ? ceil_div(NBF1, T) : 0;
  static constexpr int NUNITS = S / 2;          // last-pass units (pairs of butterflies k0, S-k0)
  static constexpr int NU = ceil_div(NUNITS, T);
  static constexpr int NCH = ceil_div(N / 8, T);  // 8-sample input chunks per thread and row (W <= N)
  static constexpr int R0P8 = ceil_div(R0, 8) * 8;
  static constexpr int R0P4 = ceil_div(R0, 4) * 4;
  static constexpr int NWARPS = T / 32;
  static constexpr int E = cmax(cmax(NB0 * R0, NB1 * R1), NU * 2 * RL);  // complex registers live in one pass
  static_assert(R0 * (THREE ? R1 : 1) * RL == N, "radices must multiply to N");
  static_assert(N2 == RL, "last radix must equal N2");
  static_assert(S % 2 == 0, "N / RL must be even");
  static_assert(T % 32 == 0, "threads per transform must be whole warps");
  static_assert(N % 8 == 0, "N must be a multiple of 8");
};

// Runtime mirror used on the host.
struct PlanDesc {
  int N, T, R0, R1, RL;
  int N1() const { return N / R0; }
  int N2() const { return R1 > 1 ? N1() / R1 : N1(); }
  int S() const { return N / RL; }
  int rowstride() const { return N1() + 1; }
  int r0p8() const { return (R0 + 7) / 8 * 8; }
  int r0p4() const { return (R0 + 3) / 4 * 4; }
};

}  // namespace abcoct
