// c4_math.cuh — the f32 math of the self-play path, bit-exact with the reference on this platform.
//
// The reference computes UCT, softmax and temperature with Rust's f32::ln / f32::exp
// (rust/src/mcts.rs:379, 430, 451-453), which lower to glibc's logf / expf.  Tree shape depends on
// an f32 argmax over those values, so "close" is not enough: the device must return the same
// bits.  glibc >= 2.27 implements both with the ARM "optimized routines" algorithm (one table
// lookup + a degree-3 polynomial evaluated in double, then one rounding to float), and on x86-64
// hosts with FMA it dispatches to the FMA-contracted build of that code.  c4_logf / c4_expf
// restate exactly that data flow — same tables, same constants, the same five fused
// multiply-adds — in double precision, which IEEE-754 makes reproducible on any hardware.
// tests/test_math_host.py compares the host build against libm over tens of millions of inputs
// (and DESIGN.md records the exhaustive 2^32 sweep); tests/test_gpu_engine.py::test_device_logf_expf_match_libm does the same on the
// device build.
//
// softmax          mcts.rs:416-434      (max; exp(x - max); left-fold sum; divide)
// apply_temperature mcts.rs:439-454
// Compile with -fmad=false: the reference never contracts f32 a*b+c.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "c4_rules.cuh"

namespace c4 {

// 2^(i/32) with the exponent contribution of i folded out (glibc __exp2f_data.tab)
#define C4_EXP2F_TAB_INIT                                                                         \
  {0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL,   \
   0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL,   \
   0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL,   \
   0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL,   \
   0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL,   \
   0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL,   \
   0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL,   \
   0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL}
// {1/c, log(c)} for 16 sub-intervals of [0x1.66p-1, 0x1.66p0) (glibc __logf_data.tab), as bits
#define C4_LOGF_TAB_INIT                                                                          \
  {0x3ff661ec79f8f3beULL, 0xbfd57bf7808caadeULL, 0x3ff571ed4aaf883dULL, 0xbfd2bef0a7c06ddbULL,   \
   0x3ff49539f0f010b0ULL, 0xbfd01eae7f513a67ULL, 0x3ff3c995b0b80385ULL, 0xbfcb31d8a68224e9ULL,   \
   0x3ff30d190c8864a5ULL, 0xbfc6574f0ac07758ULL, 0x3ff25e227b0b8ea0ULL, 0xbfc1aa2bc79c8100ULL,   \
   0x3ff1bb4a4a1a343fULL, 0xbfba4e76ce8c0e5eULL, 0x3ff12358f08ae5baULL, 0xbfb1973c5a611cccULL,   \
   0x3ff0953f419900a7ULL, 0xbfa252f438e10c1eULL, 0x3ff0000000000000ULL, 0x0000000000000000ULL,   \
   0x3fee608cfd9a47acULL, 0x3faaa5aa5df25984ULL, 0x3feca4b31f026aa0ULL, 0x3fbc5e53aa362eb4ULL,   \
   0x3feb2036576afce6ULL, 0x3fc526e57720db08ULL, 0x3fe9c2d163a1aa2dULL, 0x3fcbc2860d224770ULL,   \
   0x3fe886e6037841edULL, 0x3fd1058bc8a07ee1ULL, 0x3fe767dcf5534862ULL, 0x3fd4043057b6ee09ULL}

static const uint64_t H_EXP2F_TAB[32] = C4_EXP2F_TAB_INIT;
static const uint64_t H_LOGF_TAB[32] = C4_LOGF_TAB_INIT;
#if defined(__CUDACC__)
static __constant__ uint64_t D_EXP2F_TAB[32] = C4_EXP2F_TAB_INIT;
static __constant__ uint64_t D_LOGF_TAB[32] = C4_LOGF_TAB_INIT;
#endif
#if defined(__CUDA_ARCH__)
#define EXP2F_TAB D_EXP2F_TAB
#define LOGF_TAB D_LOGF_TAB
#else
#define EXP2F_TAB H_EXP2F_TAB
#define LOGF_TAB H_LOGF_TAB
#endif

C4_HD uint32_t f32_bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}
C4_HD float bits_f32(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
C4_HD uint64_t f64_bits(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t u;
  memcpy(&u, &d, 8);
  return u;
#endif
}
C4_HD double bits_f64(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double d;
  memcpy(&d, &u, 8);
  return d;
#endif
}
C4_HD float f32_inf() { return bits_f32(0x7f800000u); }
C4_HD float f32_nan() { return bits_f32(0x7fc00000u); }

// glibc logf (sysdeps/ieee754/flt-32/e_logf.c, FMA build).
C4_HD float c4_logf(float x) {
  uint32_t ix = f32_bits(x);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix * 2u == 0u) return -f32_inf();                        // log(+-0) = -inf
    if (ix == 0x7f800000u) return x;                             // log(inf) = inf
    if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u) return f32_nan();
    ix = f32_bits(x * 8388608.0f);                               // subnormal: scale by 2^23
    ix -= 23u << 23;
  }
  uint32_t tmp = ix - 0x3f330000u;
  int i = (int)((tmp >> 19) & 15u);
  int k = (int32_t)tmp >> 23;
  uint32_t iz = ix - (tmp & 0xff800000u);
  double invc = bits_f64(LOGF_TAB[2 * i]);
  double logc = bits_f64(LOGF_TAB[2 * i + 1]);
  double z = (double)bits_f32(iz);
  const double Ln2 = bits_f64(0x3fe62e42fefa39efULL);
  const double A0 = bits_f64(0xbfd00ea348b88334ULL);
  const double A1 = bits_f64(0x3fd5575b0be00b6aULL);
  const double A2 = bits_f64(0xbfdffffef20a4123ULL);
  double r = fma(z, invc, -1.0);
  double y0 = fma((double)k, Ln2, logc);
  double y = fma(r, A1, A2);
  double r2 = r * r;
  double t = r + y0;
  y = fma(r2, A0, y);
  y = fma(r2, y, t);
  return (float)y;
}

// glibc expf (sysdeps/ieee754/flt-32/e_expf.c, FMA build).
C4_HD float c4_expf(float x) {
  uint32_t ix = f32_bits(x);
  uint32_t abstop = (ix >> 20) & 0x7ffu;
  if (abstop > 0x42au) {  // |x| >= 88 or inf/nan
    if (ix == 0xff800000u) return 0.0f;
    if (abstop > 0x7f7u) return x + x;
    if (x > 88.72283172607421875f) return f32_inf();          // 0x1.62e42ep6: overflow
    if (x < -103.97207641601562500f) return 0.0f;             // -0x1.9fe368p6: underflow
    if (x < -103.27892303466796875f) return bits_f32(1u);     // -0x1.9d1d9ep6: 0x1.4p-75^2 -> 2^-149
  }
  double xd = (double)x;
  const double SHIFT = bits_f64(0x4338000000000000ULL);
  const double InvLn2N = bits_f64(0x40471547652b82feULL);
  const double C0 = bits_f64(0x3ebc6af84b912394ULL);
  const double C1 = bits_f64(0x3f2ebfce50fac4f3ULL);
  const double C2 = bits_f64(0x3f962e42ff0c52d6ULL);
  double kd = fma(InvLn2N, xd, SHIFT);
  uint64_t ki = f64_bits(kd);
  kd = kd - SHIFT;
  double r = fma(InvLn2N, xd, -kd);
  uint64_t t = EXP2F_TAB[ki & 31u] + (ki << 47);
  double s = bits_f64(t);
  double z = fma(r, C0, C1);
  double r2 = r * r;
  double y = fma(r, C2, 1.0);
  y = fma(z, r2, y);
  y = y * s;
  return (float)y;
}

// mcts.rs:416-434.  x[i] = -inf for masked entries.  Returns false if every entry is -inf
// (the reference panics there).
C4_HD bool softmax7(const float x[7], float out[7]) {
  float mx = -f32_inf();
  for (int i = 0; i < 7; i++) mx = (x[i] > mx) ? x[i] : mx;
  if (mx == -f32_inf()) return false;
  float e[7], s = 0.0f;
  for (int i = 0; i < 7; i++) e[i] = c4_expf(x[i] - mx);
  for (int i = 0; i < 7; i++) s = s + e[i];
  for (int i = 0; i < 7; i++) out[i] = e[i] / s;
  return true;
}

// mcts.rs:439-454
C4_HD void apply_temperature7(const float p[7], float temperature, float out[7]) {
  bool all_eq = true;
  for (int i = 1; i < 7; i++) all_eq = all_eq && (p[i] == p[0]);
  if (temperature == 1.0f || all_eq) {
    for (int i = 0; i < 7; i++) out[i] = p[i];
    return;
  }
  if (temperature == 0.0f) {
    float mx = -f32_inf();
    for (int i = 0; i < 7; i++) mx = (p[i] > mx) ? p[i] : mx;
    float r[7], s = 0.0f;
    for (int i = 0; i < 7; i++) r[i] = (p[i] == mx) ? 1.0f : 0.0f;
    for (int i = 0; i < 7; i++) s = s + r[i];
    for (int i = 0; i < 7; i++) out[i] = r[i] / s;
    return;
  }
  float l[7], s = 0.0f;
  for (int i = 0; i < 7; i++) l[i] = c4_logf(p[i]) / temperature;
  for (int i = 0; i < 7; i++) s = s + c4_expf(l[i]);
  float lse = c4_logf(s);
  for (int i = 0; i < 7; i++) {
    float v = c4_expf(l[i] - lse);
    if (v < 0.0f) v = 0.0f;
    if (v > 1.0f) v = 1.0f;
    out[i] = v;
  }
}

// self_play.rs:294-299
C4_HD float temperature_for_ply(int ply) { return ply < 4 ? 4.0f : (ply < 8 ? 2.0f : 1.0f); }

}  // namespace c4
