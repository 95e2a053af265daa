// xc_math_spec.cuh — the "SPEC" transcendentals (DESIGN.md §SPEC math), device side.
//
// binary32 in / binary32 out, binary64 inside.  Only IEEE-754 round-to-nearest +, -, *,
// fma and integer operations are used, so the same sequence of roundings can be (and is,
// independently, in oracle/xcape_oracle.cpp) reproduced on a CPU: results are bit-identical
// across the two by construction.  Against the correctly-rounded binary32 function they
// differ with probability ~2^-20 per call (internal relative error < 2^-44), i.e. they are
// "valid libm" replacements for gfortran's expf/logf/powf in the reference
// (CAPE_CODE_model_lev.f90:236,438-462,570-620).
//
// B200 note: FP64 runs at half the FP32 rate on sm_100a, which is what makes a
// double-precision core cheaper here than a float-float one.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace xc {

#define XC_SP_L2E    0x1.71547652b82fep+0   /* log2(e)            */
#define XC_SP_LN2_HI 0x1.62e42fee00000p-1   /* fdlibm split, ln 2 */
#define XC_SP_LN2_LO 0x1.a39ef35793c76p-33
#define XC_SP_MAGIC  6755399441055744.0     /* 1.5 * 2^52         */

// Polynomial / reduction constants live in constant memory: DFMA takes a c[bank][offset]
// operand directly, whereas an immediate binary64 costs two UMOV issue slots per use
// (ncu r1a: 40 of the 195 instructions of the moist-iteration body were such UMOVs).
__constant__ double kExp[16] = {
    0x1.71547652b82fep+0,    // [0]  log2(e)
    0x1.62e42fee00000p-1,    // [1]  ln2 hi
    0x1.a39ef35793c76p-33,   // [2]  ln2 lo
    0x1.ae64567f544e4p-26,   // [3]  1/11!
    0x1.27e4fb7789f5cp-22,   // [4]  1/10!
    0x1.71de3a556c734p-19,   // [5]  1/9!
    0x1.a01a01a01a01ap-16,   // [6]  1/8!
    0x1.a01a01a01a01ap-13,   // [7]  1/7!
    0x1.6c16c16c16c17p-10,   // [8]  1/6!
    0x1.1111111111111p-7,    // [9]  1/5!
    0x1.5555555555555p-5,    // [10] 1/4!
    0x1.5555555555555p-3,    // [11] 1/3!
    0.5, 1.0, 6755399441055744.0, -6755399441055744.0};
__constant__ double kLog[12] = {
    0x1.1111111111111p-4,    // 1/15
    0x1.3b13b13b13b14p-4,    // 1/13
    0x1.745d1745d1746p-4,    // 1/11
    0x1.c71c71c71c71cp-4,    // 1/9
    0x1.2492492492492p-3,    // 1/7
    0x1.999999999999ap-3,    // 1/5
    0x1.5555555555555p-2,    // 1/3
    -0.2391, 0.98525, 0x1.6a09e667f3bcdp+0, 0x1.62e42fee00000p-1, 0x1.a39ef35793c76p-33};

// exp of a double already known to lie in (-700, 700]; returns double.
__device__ __forceinline__ double spec_exp_core(double x) {
  const double tm = __dadd_rn(__dmul_rn(x, kExp[0]), kExp[14]);
  const double nd = __dadd_rn(tm, kExp[15]);
  const int n = __double2loint(tm);          // low word of t+1.5*2^52 is rint(t) in two's complement
  double r = __fma_rn(-nd, kExp[1], x);
  r = __fma_rn(-nd, kExp[2], r);
  double p = kExp[3];                        // 1/11!
  p = __fma_rn(p, r, kExp[4]);
  p = __fma_rn(p, r, kExp[5]);
  p = __fma_rn(p, r, kExp[6]);
  p = __fma_rn(p, r, kExp[7]);
  p = __fma_rn(p, r, kExp[8]);
  p = __fma_rn(p, r, kExp[9]);
  p = __fma_rn(p, r, kExp[10]);
  p = __fma_rn(p, r, kExp[11]);
  p = __fma_rn(p, r, kExp[12]);
  p = __fma_rn(p, r, kExp[13]);
  p = __fma_rn(p, r, kExp[13]);
  return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

// exp of a double with |x| <= 2^-3: no range reduction (n == 0), Taylor degree 8
// (truncation x^9/9! <= 2^-45 relative).  Used for theta2 = theta1*exp(small) (f90:460-462).
__device__ __forceinline__ double spec_exp_small_core(double x) {
  double p = kExp[6];                        // 1/8!
  p = __fma_rn(p, x, kExp[7]);
  p = __fma_rn(p, x, kExp[8]);
  p = __fma_rn(p, x, kExp[9]);
  p = __fma_rn(p, x, kExp[10]);
  p = __fma_rn(p, x, kExp[11]);
  p = __fma_rn(p, x, kExp[12]);
  p = __fma_rn(p, x, kExp[13]);
  p = __fma_rn(p, x, kExp[13]);
  return p;
}

__device__ __forceinline__ double spec_exp_d(double x) {
  if (!(x > -700.0)) return (x != x) ? x : 0.0;
  if (x > 700.0) return CUDART_INF;
  return spec_exp_core(x);
}

// log of a positive, finite, normal double.
__device__ __forceinline__ double spec_log_core(double x) {
  const int hi = __double2hiint(x);
  int e = (hi >> 20) - 1023;
  double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));   // [1,2)
  if (m > kLog[9]) { m = __dmul_rn(m, 0.5); e += 1; }
  const double f = __dadd_rn(m, -1.0);
  const double d = __dadd_rn(m, 1.0);
  double y = __fma_rn(kLog[7], d, kLog[8]);
  double t = __fma_rn(-d, y, 1.0); y = __fma_rn(y, t, y);
  t = __fma_rn(-d, y, 1.0); y = __fma_rn(y, t, y);
  t = __fma_rn(-d, y, 1.0); y = __fma_rn(y, t, y);
  const double s = __dmul_rn(f, y);
  const double z = __dmul_rn(s, s);
  double q = kLog[0];                              // 1/15
  q = __fma_rn(q, z, kLog[1]);
  q = __fma_rn(q, z, kLog[2]);
  q = __fma_rn(q, z, kLog[3]);
  q = __fma_rn(q, z, kLog[4]);
  q = __fma_rn(q, z, kLog[5]);
  q = __fma_rn(q, z, kLog[6]);
  q = __fma_rn(q, z, 1.0);
  const double lm = __dmul_rn(__dadd_rn(s, s), q);
  const double ed = (double)e;
  const double r = __fma_rn(ed, kLog[11], lm);
  return __fma_rn(ed, kLog[10], r);
}

__device__ __forceinline__ double spec_log_d(double x) {
  if (!(x > 0.0)) return (x == 0.0) ? -CUDART_INF : CUDART_NAN;
  if (x == CUDART_INF) return x;
  return spec_log_core(x);
}

// ---- binary32 front ends -------------------------------------------------------------
// The range guards are evaluated on the binary32 argument (FP32 pipe) — equivalent to the
// binary64 guards of the spec because +-700 are exact in binary32.
__device__ __forceinline__ float spec_expf(float x) {
  if (fabsf(x) <= 700.0f) return __double2float_rn(spec_exp_core((double)x));   // one compare on the hot path
  return (x != x) ? x : (x < 0.0f ? 0.0f : CUDART_INF_F);
}
// exp for arguments that are almost always tiny (|x| <= 2^-3 takes the reduction-free path)
__device__ __forceinline__ float spec_expf_small(float x) {
  if (fabsf(x) <= 0.125f) return __double2float_rn(spec_exp_small_core((double)x));
  return spec_expf(x);
}
__device__ __forceinline__ float spec_logf(float x) {
  if (!(x > 0.0f)) return (x == 0.0f) ? -CUDART_INF_F : CUDART_NAN_F;
  if (x == CUDART_INF_F) return x;
  return __double2float_rn(spec_log_core((double)x));
}
__device__ __forceinline__ float spec_powf(float x, float y) {
  return __double2float_rn(spec_exp_d(__dmul_rn((double)y, spec_log_d((double)x))));
}

}  // namespace xc
