// In-register DFT building blocks over a generic lane type V (float, or the packed pair V2 = two independent f32
// lanes held in one 64-bit register and processed by the sm_100 f32x2 instructions FADD2 / FMUL2 / FFMA2).
// Same construction as fft_regs.cuh (compile-time Cooley-Tukey over hand-written radix 2/3/4/5 butterflies, every
// internal twiddle a constant), but written so that a multiplication by +-i never needs a negation: the packed
// instructions have no operand-negate modifier, so i*z is folded into the choice of add / sub that consumes it.
#pragma once
#include <cuda_runtime.h>
#include <cstring>
#include <type_traits>

#include "fft_regs.cuh"  // ct_cossin, static_for, Split

namespace abcoct {

// ---------------------------------------------------------------- V2: two f32 lanes in one 64-bit register
struct V2 {
  unsigned long long v;
};
__host__ __device__ __forceinline__ V2 v2_make(float lo, float hi) {
  V2 r;
#ifdef __CUDA_ARCH__
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
#else
  float t[2] = {lo, hi};
  memcpy(&r.v, t, 8);
#endif
  return r;
}
__host__ __device__ __forceinline__ float v2_lo(V2 a) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(static_cast<unsigned>(a.v));
#else
  float t[2];
  memcpy(t, &a.v, 8);
  return t[0];
#endif
}
__host__ __device__ __forceinline__ float v2_hi(V2 a) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(static_cast<unsigned>(a.v >> 32));
#else
  float t[2];
  memcpy(t, &a.v, 8);
  return t[1];
#endif
}

// ---------------------------------------------------------------- lane arithmetic: float and V2 share one vocabulary
__host__ __device__ __forceinline__ float vadd(float a, float b) { return a + b; }
__host__ __device__ __forceinline__ float vsub(float a, float b) { return a - b; }
__host__ __device__ __forceinline__ float vmul(float a, float b) { return a * b; }
__host__ __device__ __forceinline__ float vfma(float a, float b, float c) { return fmaf(a, b, c); }
__host__ __device__ __forceinline__ V2 vadd(V2 a, V2 b) {
#ifdef __CUDA_ARCH__
  V2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
#else
  return v2_make(v2_lo(a) + v2_lo(b), v2_hi(a) + v2_hi(b));
#endif
}
__host__ __device__ __forceinline__ V2 vsub(V2 a, V2 b) {
#ifdef __CUDA_ARCH__
  V2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
#else
  return v2_make(v2_lo(a) - v2_lo(b), v2_hi(a) - v2_hi(b));
#endif
}
__host__ __device__ __forceinline__ V2 vmul(V2 a, V2 b) {
#ifdef __CUDA_ARCH__
  V2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
#else
  return v2_make(v2_lo(a) * v2_lo(b), v2_hi(a) * v2_hi(b));
#endif
}
__host__ __device__ __forceinline__ V2 vfma(V2 a, V2 b, V2 c) {
#ifdef __CUDA_ARCH__
  V2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
#else
  return v2_make(fmaf(v2_lo(a), v2_lo(b), v2_lo(c)), fmaf(v2_hi(a), v2_hi(b), v2_hi(c)));
#endif
}
// broadcast of a scalar (compile-time constants become two MOV-immediates that the compiler hoists)
template <class V>
__host__ __device__ __forceinline__ V vbc(float s) {
  if constexpr (std::is_same<V, float>::value)
    return s;
  else
    return v2_make(s, s);
}

template <class V>
struct Cx {
  V x, y;
};
template <class V>
__host__ __device__ __forceinline__ Cx<V> cxadd(Cx<V> a, Cx<V> b) { return Cx<V>{vadd(a.x, b.x), vadd(a.y, b.y)}; }
template <class V>
__host__ __device__ __forceinline__ Cx<V> cxsub(Cx<V> a, Cx<V> b) { return Cx<V>{vsub(a.x, b.x), vsub(a.y, b.y)}; }
// a + SGN * i * b   and   a - SGN * i * b   without negating anything
template <int SGN, class V>
__host__ __device__ __forceinline__ Cx<V> cxadd_i(Cx<V> a, Cx<V> b) {
  if constexpr (SGN > 0)
    return Cx<V>{vsub(a.x, b.y), vadd(a.y, b.x)};
  else
    return Cx<V>{vadd(a.x, b.y), vsub(a.y, b.x)};
}
template <int SGN, class V>
__host__ __device__ __forceinline__ Cx<V> cxsub_i(Cx<V> a, Cx<V> b) { return cxadd_i<-SGN>(a, b); }
// z * (wx + i wy) with run-time twiddle lanes
template <class V>
__host__ __device__ __forceinline__ Cx<V> cxmul(Cx<V> a, V wx, V wy, V nwy) {  // nwy = -wy
  return Cx<V>{vfma(a.y, nwy, vmul(a.x, wx)), vfma(a.y, wx, vmul(a.x, wy))};
}

// z * exp(SGN * 2 pi i NUM / DEN), compile-time angle
template <int NUM, int DEN, int SGN, class V>
__host__ __device__ __forceinline__ Cx<V> cxmul_w(Cx<V> z) {
  constexpr int n = ((NUM % DEN) + DEN) % DEN;
  if constexpr (n == 0) {
    return z;
  } else if constexpr (4 * n == DEN) {  // * (SGN i): only reachable through callers that cannot fold it
    if constexpr (SGN > 0)
      return Cx<V>{vsub(vbc<V>(0.f), z.y), z.x};
    else
      return Cx<V>{z.y, vsub(vbc<V>(0.f), z.x)};
  } else if constexpr (2 * n == DEN) {
    return Cx<V>{vsub(vbc<V>(0.f), z.x), vsub(vbc<V>(0.f), z.y)};
  } else if constexpr (4 * n == 3 * DEN) {
    if constexpr (SGN > 0)
      return Cx<V>{z.y, vsub(vbc<V>(0.f), z.x)};
    else
      return Cx<V>{vsub(vbc<V>(0.f), z.y), z.x};
  } else {
    constexpr CtCS cs = ct_cossin(n, DEN);
    constexpr float c = float(cs.c);
    constexpr float s = float(SGN > 0 ? cs.s : -cs.s);
    return Cx<V>{vfma(z.y, vbc<V>(-s), vmul(z.x, vbc<V>(c))), vfma(z.y, vbc<V>(c), vmul(z.x, vbc<V>(s)))};
  }
}

// ---------------------------------------------------------------- DftV<R>: out[OS*c] = sum_a in[IS*a] * w_R^(SGN*a*c)
template <int R, int SGN, int IS, int OS, class V>
struct DftV {
  static_assert(Split<R>::RA != R, "unsupported prime radix (only 2, 3, 5 and their products)");
  __host__ __device__ static __forceinline__ void run(const Cx<V>* in, Cx<V>* out) {
    constexpr int RA = Split<R>::RA, RB = Split<R>::RB;
    Cx<V> t[R];  // t[a2*RA + c1]
    static_for<0, RB>([&](auto a2c) {
      constexpr int a2 = decltype(a2c)::value;
      DftV<RA, SGN, IS * RB, 1, V>::run(in + IS * a2, t + a2 * RA);
    });
    static_for<1, RB>([&](auto a2c) {
      constexpr int a2 = decltype(a2c)::value;
      static_for<1, RA>([&](auto c1c) {
        constexpr int c1 = decltype(c1c)::value;
        t[a2 * RA + c1] = cxmul_w<a2 * c1, R, SGN, V>(t[a2 * RA + c1]);
      });
    });
    static_for<0, RA>([&](auto c1c) {
      constexpr int c1 = decltype(c1c)::value;
      DftV<RB, SGN, RA, OS * RA, V>::run(t + c1, out + OS * c1);
    });
  }
};
template <int SGN, int IS, int OS, class V>
struct DftV<1, SGN, IS, OS, V> {
  __host__ __device__ static __forceinline__ void run(const Cx<V>* in, Cx<V>* out) { out[0] = in[0]; }
};
template <int SGN, int IS, int OS, class V>
struct DftV<2, SGN, IS, OS, V> {
  __host__ __device__ static __forceinline__ void run(const Cx<V>* in, Cx<V>* out) {
    const Cx<V> a = in[0], b = in[IS];
    out[0] = cxadd(a, b);
    out[OS] = cxsub(a, b);
  }
};
template <int SGN, int IS, int OS, class V>
struct DftV<4, SGN, IS, OS, V> {
  __host__ __device__ static __forceinline__ void run(const Cx<V>* in, Cx<V>* out) {
    const Cx<V> a0 = in[0], a1 = in[IS], a2 = in[2 * IS], a3 = in[3 * IS];
    const Cx<V> t0 = cxadd(a0, a2), t1 = cxsub(a0, a2), t2 = cxadd(a1, a3), d = cxsub(a1, a3);
    out[0] = cxadd(t0, t2);
    out[OS] = cxadd_i<SGN>(t1, d);
    out[2 * OS] = cxsub(t0, t2);
    out[3 * OS] = cxsub_i<SGN>(t1, d);
  }
};
template <int SGN, int IS, int OS, class V>
struct DftV<3, SGN, IS, OS, V> {
  __host__ __device__ static __forceinline__ void run(const Cx<V>* in, Cx<V>* out) {
    constexpr float kS = 0.86602540378443864676f;  // sin(2 pi / 3)
    const Cx<V> a0 = in[0], a1 = in[IS], a2 = in[2 * IS];
    const Cx<V> s = cxadd(a1, a2), d = cxsub(a1, a2);
    const Cx<V> m{vfma(vbc<V>(-0.5f), s.x, a0.x), vfma(vbc<V>(-0.5f), s.y, a0.y)};
    const Cx<V> r{vmul(vbc<V>(kS), d.x), vmul(vbc<V>(kS), d.y)};
    out[0] = cxadd(a0, s);
    out[OS] = cxadd_i<SGN>(m, r);
    out[2 * OS] = cxsub_i<SGN>(m, r);
  }
};
template <int SGN, int IS, int OS, class V>
struct DftV<5, SGN, IS, OS, V> {
  __host__ __device__ static __forceinline__ void run(const Cx<V>* in, Cx<V>* out) {
    constexpr float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    constexpr float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    const Cx<V> a0 = in[0], a1 = in[IS], a2 = in[2 * IS], a3 = in[3 * IS], a4 = in[4 * IS];
    const Cx<V> p1 = cxadd(a1, a4), p2 = cxadd(a2, a3), d1 = cxsub(a1, a4), d2 = cxsub(a2, a3);
    const Cx<V> m1{vfma(vbc<V>(c2), p2.x, vfma(vbc<V>(c1), p1.x, a0.x)), vfma(vbc<V>(c2), p2.y, vfma(vbc<V>(c1), p1.y, a0.y))};
    const Cx<V> m2{vfma(vbc<V>(c1), p2.x, vfma(vbc<V>(c2), p1.x, a0.x)), vfma(vbc<V>(c1), p2.y, vfma(vbc<V>(c2), p1.y, a0.y))};
    const Cx<V> r1{vfma(vbc<V>(s2), d2.x, vmul(vbc<V>(s1), d1.x)), vfma(vbc<V>(s2), d2.y, vmul(vbc<V>(s1), d1.y))};
    const Cx<V> r2{vfma(vbc<V>(-s1), d2.x, vmul(vbc<V>(s2), d1.x)), vfma(vbc<V>(-s1), d2.y, vmul(vbc<V>(s2), d1.y))};
    out[0] = Cx<V>{vadd(vadd(a0.x, p1.x), p2.x), vadd(vadd(a0.y, p1.y), p2.y)};
    out[OS] = cxadd_i<SGN>(m1, r1);
    out[4 * OS] = cxsub_i<SGN>(m1, r1);
    out[2 * OS] = cxadd_i<SGN>(m2, r2);
    out[3 * OS] = cxsub_i<SGN>(m2, r2);
  }
};

}  // namespace abcoct
