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

// 2^(j/128), j = 0..127, correctly rounded to binary64.  Read through the read-only data path
// (__ldg): 1 KB, L1-resident; the index differs per lane, which constant memory would serialise.
__device__ const double kExpT[128] = {
    0x1.0000000000000p+0, 0x1.0163da9fb3335p+0, 0x1.02c9a3e778061p+0, 0x1.04315e86e7f85p+0,
    0x1.059b0d3158574p+0, 0x1.0706b29ddf6dep+0, 0x1.0874518759bc8p+0, 0x1.09e3ecac6f383p+0,
    0x1.0b5586cf9890fp+0, 0x1.0cc922b7247f7p+0, 0x1.0e3ec32d3d1a2p+0, 0x1.0fb66affed31bp+0,
    0x1.11301d0125b51p+0, 0x1.12abdc06c31ccp+0, 0x1.1429aaea92de0p+0, 0x1.15a98c8a58e51p+0,
    0x1.172b83c7d517bp+0, 0x1.18af9388c8deap+0, 0x1.1a35beb6fcb75p+0, 0x1.1bbe084045cd4p+0,
    0x1.1d4873168b9aap+0, 0x1.1ed5022fcd91dp+0, 0x1.2063b88628cd6p+0, 0x1.21f49917ddc96p+0,
    0x1.2387a6e756238p+0, 0x1.251ce4fb2a63fp+0, 0x1.26b4565e27cddp+0, 0x1.284dfe1f56381p+0,
    0x1.29e9df51fdee1p+0, 0x1.2b87fd0dad990p+0, 0x1.2d285a6e4030bp+0, 0x1.2ecafa93e2f56p+0,
    0x1.306fe0a31b715p+0, 0x1.32170fc4cd831p+0, 0x1.33c08b26416ffp+0, 0x1.356c55f929ff1p+0,
    0x1.371a7373aa9cbp+0, 0x1.38cae6d05d866p+0, 0x1.3a7db34e59ff7p+0, 0x1.3c32dc313a8e5p+0,
    0x1.3dea64c123422p+0, 0x1.3fa4504ac801cp+0, 0x1.4160a21f72e2ap+0, 0x1.431f5d950a897p+0,
    0x1.44e086061892dp+0, 0x1.46a41ed1d0057p+0, 0x1.486a2b5c13cd0p+0, 0x1.4a32af0d7d3dep+0,
    0x1.4bfdad5362a27p+0, 0x1.4dcb299fddd0dp+0, 0x1.4f9b2769d2ca7p+0, 0x1.516daa2cf6642p+0,
    0x1.5342b569d4f82p+0, 0x1.551a4ca5d920fp+0, 0x1.56f4736b527dap+0, 0x1.58d12d497c7fdp+0,
    0x1.5ab07dd485429p+0, 0x1.5c9268a5946b7p+0, 0x1.5e76f15ad2148p+0, 0x1.605e1b976dc09p+0,
    0x1.6247eb03a5585p+0, 0x1.6434634ccc320p+0, 0x1.6623882552225p+0, 0x1.68155d44ca973p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6c012750bdabfp+0, 0x1.6dfb23c651a2fp+0, 0x1.6ff7df9519484p+0,
    0x1.71f75e8ec5f74p+0, 0x1.73f9a48a58174p+0, 0x1.75feb564267c9p+0, 0x1.780694fde5d3fp+0,
    0x1.7a11473eb0187p+0, 0x1.7c1ed0130c132p+0, 0x1.7e2f336cf4e62p+0, 0x1.80427543e1a12p+0,
    0x1.82589994cce13p+0, 0x1.8471a4623c7adp+0, 0x1.868d99b4492edp+0, 0x1.88ac7d98a6699p+0,
    0x1.8ace5422aa0dbp+0, 0x1.8cf3216b5448cp+0, 0x1.8f1ae99157736p+0, 0x1.9145b0b91ffc6p+0,
    0x1.93737b0cdc5e5p+0, 0x1.95a44cbc8520fp+0, 0x1.97d829fde4e50p+0, 0x1.9a0f170ca07bap+0,
    0x1.9c49182a3f090p+0, 0x1.9e86319e32323p+0, 0x1.a0c667b5de565p+0, 0x1.a309bec4a2d33p+0,
    0x1.a5503b23e255dp+0, 0x1.a799e1330b358p+0, 0x1.a9e6b5579fdbfp+0, 0x1.ac36bbfd3f37ap+0,
    0x1.ae89f995ad3adp+0, 0x1.b0e07298db666p+0, 0x1.b33a2b84f15fbp+0, 0x1.b59728de5593ap+0,
    0x1.b7f76f2fb5e47p+0, 0x1.ba5b030a1064ap+0, 0x1.bcc1e904bc1d2p+0, 0x1.bf2c25bd71e09p+0,
    0x1.c199bdd85529cp+0, 0x1.c40ab5fffd07ap+0, 0x1.c67f12e57d14bp+0, 0x1.c8f6d9406e7b5p+0,
    0x1.cb720dcef9069p+0, 0x1.cdf0b555dc3fap+0, 0x1.d072d4a07897cp+0, 0x1.d2f87080d89f2p+0,
    0x1.d5818dcfba487p+0, 0x1.d80e316c98398p+0, 0x1.da9e603db3285p+0, 0x1.dd321f301b460p+0,
    0x1.dfc97337b9b5fp+0, 0x1.e264614f5a129p+0, 0x1.e502ee78b3ff6p+0, 0x1.e7a51fbc74c83p+0,
    0x1.ea4afa2a490dap+0, 0x1.ecf482d8e67f1p+0, 0x1.efa1bee615a27p+0, 0x1.f252b376bba97p+0,
    0x1.f50765b6e4540p+0, 0x1.f7bfdad9cbe14p+0, 0x1.fa7c1819e90d8p+0, 0x1.fd3c22b8f71f1p+0};
__constant__ double kExp2[8] = {
    0x1.71547652b82fep+7,     // [0] 128 log2(e)
    0x1.62e42fee00000p-8,     // [1] ln2/128 hi (32 significant bits)
    0x1.a39ef35793c76p-40,    // [2] ln2/128 lo
    0x1.5555555555555p-5,     // [3] 1/24
    0x1.5555555555555p-3,     // [4] 1/6
    0.5, 6755399441055744.0, -6755399441055744.0};

// exp of a double already known to lie in [-700, 700]; returns double.
//   n = rint(128 x log2 e), r = x - n ln2/128 (|r| <= 0.0028), j = n mod 128, e = floor(n/128)
//   exp(x) = 2^e * T[j] * (1 + r(1 + r(1/2 + r(1/6 + r/24))))          truncation < 2^-49
// 10 FP64 issue slots against 16 for the table-free degree-11 form it replaces (ncu r1c: FP64 and
// XU instructions, at 2+ dispatch cycles each, cap the faithful kernel's IPC).
__device__ __forceinline__ double spec_exp_core(double x) {
  const double tm = __dadd_rn(__dmul_rn(x, kExp2[0]), kExp2[6]);
  const double nd = __dsub_rn(tm, kExp2[6]);   // same value as tm + (-magic): one constant fewer to load
  const int n = __double2loint(tm);          // low word of t+1.5*2^52 is rint(t) in two's complement
  double r = __fma_rn(-nd, kExp2[1], x);
  r = __fma_rn(-nd, kExp2[2], r);
  double q = kExp2[3];
  q = __fma_rn(q, r, kExp2[4]);
  q = __fma_rn(q, r, 0.5);                      // literal: fits the instruction's 32-bit immediate
  q = __fma_rn(q, r, 1.0);
  q = __fma_rn(q, r, 1.0);
  const double s = __dmul_rn(__ldg(&kExpT[n & 127]), q);
  return __hiloint2double(__double2hiint(s) + ((n >> 7) << 20), __double2loint(s));
}

// exp of a double with |x| <= 2^-3: no range reduction (n == 0).  |x| <= 2^-6 (the usual case for
// theta2 = theta1*exp(.), f90:460-462): Taylor degree 5 (truncation < 2^-45); else degree 8.
__device__ __forceinline__ double spec_exp_small_core(double x, bool tiny) {
  double p;
  if (tiny) {
    p = kExp[9];                             // 1/5!
  } else {
    p = kExp[6];                             // 1/8!
    p = __fma_rn(p, x, kExp[7]);
    p = __fma_rn(p, x, kExp[8]);
    p = __fma_rn(p, x, kExp[9]);
  }
  p = __fma_rn(p, x, kExp[10]);
  p = __fma_rn(p, x, kExp[11]);
  p = __fma_rn(p, x, 0.5);
  p = __fma_rn(p, x, 1.0);
  p = __fma_rn(p, x, 1.0);
  return p;
}

__device__ __forceinline__ double spec_exp_d(double x) {
  if (!(x >= -700.0)) return (x != x) ? x : 0.0;
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
  const float ax = fabsf(x);
  // three separate returns: written as one call with a `tiny` flag, nvcc if-converts the tiers and the
  // usual (tiny) case pays for the three extra DFMAs of the degree-8 tier plus two selects
  if (ax <= 0.015625f) return __double2float_rn(spec_exp_small_core((double)x, true));
  if (ax <= 0.125f) return __double2float_rn(spec_exp_small_core((double)x, false));
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
