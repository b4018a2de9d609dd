// In-register DFT building blocks for the ABC-OCT row FFT (replaces cv::dft(..., DFT_ROWS), reference
// BscanFFT.cpp:1185 and the two dft() calls of zeropadrowwise, BscanFFT.cpp:211, 241).
//
// Everything is compile-time shaped: radix, direction, input/output strides and every internal twiddle
// are template constants, so after inlining the local arrays live in registers and the twiddles become
// FMUL/FFMA immediates.  Composite radices are built by one Cooley-Tukey step R = RA * RB; primes 2,3,5
// and radix 4 are hand-written.  No tensor cores: the transform is not a dense contraction.
#pragma once
#include <cuda_runtime.h>
#include <type_traits>
#include <cmath>

namespace abcoct {

// ---------------------------------------------------------------- constexpr trigonometry (double, compile time)
__host__ __device__ constexpr double ct_sin_small(double x) {  // |x| <= pi/4
  double x2 = x * x, term = x, sum = x;
  for (int i = 1; i < 14; ++i) {
    term *= -x2 / double((2 * i) * (2 * i + 1));
    sum += term;
  }
  return sum;
}
__host__ __device__ constexpr double ct_cos_small(double x) {
  double x2 = x * x, term = 1.0, sum = 1.0;
  for (int i = 1; i < 14; ++i) {
    term *= -x2 / double((2 * i - 1) * (2 * i));
    sum += term;
  }
  return sum;
}
constexpr double kPiD = 3.14159265358979323846264338327950288;

// cos / sin of 2*pi*num/den via octant reduction on the rational num/den.
struct CtCS {
  double c, s;
};
__host__ __device__ constexpr CtCS ct_cossin(long long num, long long den) {
  num %= den;
  if (num < 0) num += den;
  double sc = 1.0, ss = 1.0;
  if (2 * num > den) {  // f > 1/2 : f -> 1 - f
    num = den - num;
    ss = -ss;
  }
  // f in [0, 1/2]
  long long n2 = num, d2 = den;
  if (4 * n2 > d2) {  // f > 1/4 : f -> 1/2 - f
    n2 = d2 - 2 * n2;
    d2 = 2 * d2;
    sc = -sc;
  }
  // f in [0, 1/4]
  bool swap = false;
  if (8 * n2 > d2) {  // f > 1/8 : f -> 1/4 - f, swap cos/sin
    n2 = d2 - 4 * n2;
    d2 = 4 * d2;
    swap = true;
  }
  double x = 2.0 * kPiD * double(n2) / double(d2);
  double c = ct_cos_small(x), s = ct_sin_small(x);
  if (swap) {
    double t = c;
    c = s;
    s = t;
  }
  return CtCS{sc * c, ss * s};
}

// ---------------------------------------------------------------- complex helpers
// The re / im pair of ONE complex value is an aligned 64-bit register pair, so complex add / sub / scale are single packed
// sm_100 instructions (add / sub / mul / fma .f32x2 -> SASS FADD2 / FMUL2 / FFMA2, which also take a broadcast immediate):
// they halve the issue slots of the butterflies without needing any extra registers.
typedef unsigned long long abc_u64;
__host__ __device__ __forceinline__ float2 pk_add(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
  float2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<abc_u64*>(&r)) : "l"(*reinterpret_cast<abc_u64*>(&a)), "l"(*reinterpret_cast<abc_u64*>(&b)));
  return r;
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
__host__ __device__ __forceinline__ float2 pk_sub(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
  float2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<abc_u64*>(&r)) : "l"(*reinterpret_cast<abc_u64*>(&a)), "l"(*reinterpret_cast<abc_u64*>(&b)));
  return r;
#else
  return make_float2(a.x - b.x, a.y - b.y);
#endif
}
__host__ __device__ __forceinline__ float2 pk_mul(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
  float2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<abc_u64*>(&r)) : "l"(*reinterpret_cast<abc_u64*>(&a)), "l"(*reinterpret_cast<abc_u64*>(&b)));
  return r;
#else
  return make_float2(a.x * b.x, a.y * b.y);
#endif
}
__host__ __device__ __forceinline__ float2 pk_fma(float2 a, float2 b, float2 c) {
#ifdef __CUDA_ARCH__
  float2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(*reinterpret_cast<abc_u64*>(&r))
      : "l"(*reinterpret_cast<abc_u64*>(&a)), "l"(*reinterpret_cast<abc_u64*>(&b)), "l"(*reinterpret_cast<abc_u64*>(&c)));
  return r;
#else
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
__host__ __device__ __forceinline__ float2 pk_scale(float2 a, float s) { return pk_mul(a, make_float2(s, s)); }
__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) { return pk_add(a, b); }
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) { return pk_sub(a, b); }
__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 w) {
  return make_float2(fmaf(-a.y, w.y, a.x * w.x), fmaf(a.y, w.x, a.x * w.y));
}
// multiply by SGN * i
template <int SGN>
__host__ __device__ __forceinline__ float2 mul_i(float2 a) {
  if constexpr (SGN > 0)
    return make_float2(-a.y, a.x);
  else
    return make_float2(a.y, -a.x);
}

// z * exp(SGN * 2*pi*i * NUM / DEN) with a compile-time twiddle; trivial angles cost no multiply.
template <int NUM, int DEN, int SGN>
__host__ __device__ __forceinline__ float2 mul_w(float2 z) {
  constexpr int n = ((NUM % DEN) + DEN) % DEN;
  if constexpr (n == 0) {
    return z;
  } else if constexpr (4 * n == DEN) {
    return mul_i<SGN>(z);
  } else if constexpr (2 * n == DEN) {
    return make_float2(-z.x, -z.y);
  } else if constexpr (4 * n == 3 * DEN) {
    return mul_i<-SGN>(z);
  } else {
    constexpr CtCS cs = ct_cossin(n, DEN);
    constexpr float c = float(cs.c);
    constexpr float s = float(SGN > 0 ? cs.s : -cs.s);
    if constexpr (8 * n == DEN || 8 * n == 3 * DEN || 8 * n == 5 * DEN || 8 * n == 7 * DEN) {
      // |c| == |s| == sqrt(1/2): (c zx - s zy, s zx + c zy) = h * (+-zx -+ zy, ...)
      constexpr float h = 0.70710678118654752440f;
      constexpr bool cp = cs.c > 0, sp = (SGN > 0 ? cs.s : -cs.s) > 0;
      float re = (cp ? z.x : -z.x) - (sp ? z.y : -z.y);
      float im = (sp ? z.x : -z.x) + (cp ? z.y : -z.y);
      return pk_scale(make_float2(re, im), h);
    } else {
      const float2 t = pk_scale(z, c);
      return make_float2(fmaf(-s, z.y, t.x), fmaf(s, z.x, t.y));
    }
  }
}

// compile-time loop
template <int I, int N, class F>
__host__ __device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// ---------------------------------------------------------------- DFT<R>: out[OS*c] = sum_a in[IS*a] * w_R^(SGN*a*c)
template <int R>
struct Split {  // first Cooley-Tukey factor RA of a composite radix
  static constexpr int RA = (R % 4 == 0 && R > 4) ? 4 : (R % 2 == 0 && R > 2) ? 2 : (R % 3 == 0 && R > 3) ? 3 : (R % 5 == 0 && R > 5) ? 5 : R;
  static constexpr int RB = R / RA;
};

template <int R, int SGN, int IS, int OS>
struct Dft {
  static_assert(Split<R>::RA != R, "unsupported prime radix (only 2, 3, 5 and their products)");
  __host__ __device__ static __forceinline__ void run(const float2* in, float2* out) {
    constexpr int RA = Split<R>::RA, RB = Split<R>::RB;
    float2 t[R];  // t[a2*RA + c1]
    static_for<0, RB>([&](auto a2c) {
      constexpr int a2 = decltype(a2c)::value;
      Dft<RA, SGN, IS * RB, 1>::run(in + IS * a2, t + a2 * RA);
    });
    static_for<1, RB>([&](auto a2c) {
      constexpr int a2 = decltype(a2c)::value;
      static_for<1, RA>([&](auto c1c) {
        constexpr int c1 = decltype(c1c)::value;
        t[a2 * RA + c1] = mul_w<a2 * c1, R, SGN>(t[a2 * RA + c1]);
      });
    });
    static_for<0, RA>([&](auto c1c) {
      constexpr int c1 = decltype(c1c)::value;
      Dft<RB, SGN, RA, OS * RA>::run(t + c1, out + OS * c1);
    });
  }
};

template <int SGN, int IS, int OS>
struct Dft<1, SGN, IS, OS> {
  __host__ __device__ static __forceinline__ void run(const float2* in, float2* out) { out[0] = in[0]; }
};

template <int SGN, int IS, int OS>
struct Dft<2, SGN, IS, OS> {
  __host__ __device__ static __forceinline__ void run(const float2* in, float2* out) {
    float2 a = in[0], b = in[IS];
    out[0] = cadd(a, b);
    out[OS] = csub(a, b);
  }
};

template <int SGN, int IS, int OS>
struct Dft<4, SGN, IS, OS> {
  __host__ __device__ static __forceinline__ void run(const float2* in, float2* out) {
    float2 a0 = in[0], a1 = in[IS], a2 = in[2 * IS], a3 = in[3 * IS];
    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3);
    float2 t3 = mul_i<SGN>(csub(a1, a3));
    out[0] = cadd(t0, t2);
    out[OS] = cadd(t1, t3);
    out[2 * OS] = csub(t0, t2);
    out[3 * OS] = csub(t1, t3);
  }
};

template <int SGN, int IS, int OS>
struct Dft<3, SGN, IS, OS> {
  __host__ __device__ static __forceinline__ void run(const float2* in, float2* out) {
    constexpr float kS = 0.86602540378443864676f;  // sin(2 pi / 3)
    float2 a0 = in[0], a1 = in[IS], a2 = in[2 * IS];
    float2 s = cadd(a1, a2), d = csub(a1, a2);
    float2 m = pk_fma(s, make_float2(-0.5f, -0.5f), a0);
    float2 r = mul_i<SGN>(pk_scale(d, kS));
    out[0] = cadd(a0, s);
    out[OS] = cadd(m, r);
    out[2 * OS] = csub(m, r);
  }
};

template <int SGN, int IS, int OS>
struct Dft<5, SGN, IS, OS> {
  __host__ __device__ static __forceinline__ void run(const float2* in, float2* out) {
    constexpr float c1 = 0.30901699437494742410f;   // cos(2 pi / 5)
    constexpr float c2 = -0.80901699437494742410f;  // cos(4 pi / 5)
    constexpr float s1 = 0.95105651629515357212f;   // sin(2 pi / 5)
    constexpr float s2 = 0.58778525229247312917f;   // sin(4 pi / 5)
    float2 a0 = in[0], a1 = in[IS], a2 = in[2 * IS], a3 = in[3 * IS], a4 = in[4 * IS];
    float2 p1 = cadd(a1, a4), p2 = cadd(a2, a3), d1 = csub(a1, a4), d2 = csub(a2, a3);
    float2 m1 = pk_fma(p2, make_float2(c2, c2), pk_fma(p1, make_float2(c1, c1), a0));
    float2 m2 = pk_fma(p2, make_float2(c1, c1), pk_fma(p1, make_float2(c2, c2), a0));
    float2 r1 = mul_i<SGN>(pk_fma(d2, make_float2(s2, s2), pk_scale(d1, s1)));
    float2 r2 = mul_i<SGN>(pk_fma(d2, make_float2(-s1, -s1), pk_scale(d1, s2)));
    out[0] = cadd(cadd(a0, p1), p2);
    out[OS] = cadd(m1, r1);
    out[4 * OS] = csub(m1, r1);
    out[2 * OS] = cadd(m2, r2);
    out[3 * OS] = csub(m2, r2);
  }
};

// convenience: natural-order in-place transform of a register array
template <int R, int SGN>
__host__ __device__ __forceinline__ void dft_inplace(float2 (&z)[R]) {
  float2 o[R];
  Dft<R, SGN, 1, 1>::run(z, o);
#pragma unroll
  for (int i = 0; i < R; ++i) z[i] = o[i];
}

}  // namespace abcoct
