// cape_kernel2.cuh — the FAITHFUL CAPE/CIN kernel for sm_100a: TWO columns per thread, moist iteration in
// packed binary32 arithmetic (FFMA2 / FADD2 / FMUL2 — instructions that exist from sm_100 on).
//
// Why (ncu r2a): with the exp in binary32 the one-column kernel is bound by instruction ISSUE
// (84 % of the slots busy, 4 eligible warps per scheduler), and 75 % of what it issues are FP32 adds, multiplies
// and fmas of the moist fixed-point body.  A packed instruction does the same IEEE-rounded operation on two
// independent values for ONE issue slot (measured: FFMA2 sustains the same flop rate as FFMA with half the
// instructions, profiles/lab/f2_probe.cu), so a thread that carries the two adjacent columns 2t and 2t+1 through
// the iteration together needs ~60 issue slots per column-pass instead of ~91, and the bound moves from the
// issue port to the FMA pipe itself.
//
// Every packed operation is the same round-to-nearest binary32 operation the one-column kernel (and the
// reference) performs, applied lane-wise, so results are bit-identical to cape_kernel.cuh — the parity tests do
// not know which kernel ran.  Everything that is not the moist iteration (source parcel, per-level environment,
// sub-step prologue, buoyancy integration, the general fall-back loop) is the scalar code of cape_kernel.cuh
// run once per column.
//
// Structure of the ascent (differences from the one-column kernel, none of which changes a result):
//  * layers are walked by ABSOLUTE level index, so lanes (and the two columns of a thread) whose parcels start
//    at different levels meet the same layer — and therefore the same number of sub-steps on pressure grids —
//    in the same loop trip (profiles/divergence_model.py: 2501 -> 2398 loop trips per warp);
//  * the pressure window that licenses the unguarded log / pow / division of the sub-step prologue and the
//    temperature bound tmax of the moist window are evaluated once per LAYER, at the layer's top pressure,
//    instead of once per sub-step.
#pragma once
#include "cape_kernel.cuh"

namespace xc {

using f2 = float2;
__device__ __forceinline__ f2 splat(float x) { return make_float2(x, x); }
__device__ __forceinline__ f2 vneg(f2 a) { return make_float2(-a.x, -a.y); }            // folds into an operand modifier
__device__ __forceinline__ f2 vadd(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 vsub(f2 a, f2 b) { return __fadd2_rn(a, vneg(b)); }
__device__ __forceinline__ f2 vmul(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 vfma(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
// ptxas 12.9 contracts a packed multiply whose only use is a packed add into FFMA2 even under --fmad=false and with
// explicit .rn on both PTX instructions (it does not for the scalar forms).  Where the reference rounds the
// product first (a*b + c written as two operations), the add is therefore done as two scalar FADDs, which ptxas
// leaves alone; tests/test_gpu_parity.py's bit-exact comparisons would expose any contraction that slipped in.
__device__ __forceinline__ f2 sadd(f2 a, f2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ f2 sadd(f2 a, float b) { return make_float2(__fadd_rn(a.x, b), __fadd_rn(a.y, b)); }
__device__ __forceinline__ f2 ssub(f2 a, f2 b) { return make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)); }
// The packed form of such an add: fma(a, 1, b) rounds a*1 + b = a + b once, i.e. IS the IEEE add, for one issue slot
// instead of two scalar FADDs — provided the compiler cannot simplify it back into an add and contract that with the
// multiply that produced `a`: the 1 is a kernel argument (CapeArgs::one2), opaque at compile time.
__device__ __forceinline__ f2 vaddx(f2 a, f2 b, f2 one) { return __ffma2_rn(a, one, b); }
__device__ __forceinline__ f2 vaddx(f2 a, float b, f2 one) { return __ffma2_rn(a, one, splat(b)); }
__device__ __forceinline__ f2 vsubx(f2 a, f2 b, f2 one) { return __ffma2_rn(b, vneg(one), a); }   // a - b
__device__ __forceinline__ f2 vadd(f2 a, float b) { return __fadd2_rn(a, splat(b)); }
__device__ __forceinline__ f2 vmul(f2 a, float b) { return __fmul2_rn(a, splat(b)); }
__device__ __forceinline__ f2 vfma(f2 a, float b, float c) { return __ffma2_rn(a, splat(b), splat(c)); }
__device__ __forceinline__ f2 vfma(f2 a, f2 b, float c) { return __ffma2_rn(a, b, splat(c)); }
__device__ __forceinline__ f2 vfma(f2 a, float b, f2 c) { return __ffma2_rn(a, splat(b), c); }

// fdiv_fast (cape_kernel.cuh) on both halves: two MUFU.RCP, five packed fmas
__device__ __forceinline__ f2 vdiv_fast(f2 a, f2 b) {
  f2 r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(b.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(b.y));
  const f2 e = vfma(vneg(b), r, 1.0f);
  r = vfma(r, e, r);
  const f2 q = vfma(a, r, 0.0f);
  const f2 rem = vfma(vneg(b), q, a);
  return vfma(r, rem, q);
}

// spec32_exp_core (xc_math_spec.cuh) on both halves; |x| <= 87
__device__ __forceinline__ f2 vexp32_core(f2 x) {
  const f2 t = vfma(x, e32::kKL, e32::kMagic);
  const f2 nf = vadd(t, -e32::kMagic);
  f2 r = vfma(nf, -e32::kL1, x);
  r = vfma(nf, -e32::kL2, r);
  const int bx = __float_as_int(t.x), by = __float_as_int(t.y);
  // table addresses and exponent insertion in PTX: ptxas turns `and + mad.lo` into LOP3 + IMAD (two issue slots per
  // half each), where the compiler's own canonical form (shift, mask, add) takes three
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sExp32);
  unsigned ax, ay;
  asm("{ .reg .b32 j; and.b32 j, %1, 1023; mad.lo.u32 %0, j, 4, %2; }" : "=r"(ax) : "r"(bx), "r"(sbase));
  asm("{ .reg .b32 j; and.b32 j, %1, 1023; mad.lo.u32 %0, j, 4, %2; }" : "=r"(ay) : "r"(by), "r"(sbase));
  f2 Th, Tl;
  asm("ld.shared.f32 %0, [%1];" : "=f"(Th.x) : "r"(ax));
  asm("ld.shared.f32 %0, [%1];" : "=f"(Th.y) : "r"(ay));
  asm("ld.shared.f32 %0, [%1+4096];" : "=f"(Tl.x) : "r"(ax));
  asm("ld.shared.f32 %0, [%1+4096];" : "=f"(Tl.y) : "r"(ay));
  const f2 q = vfma(r, e32::kC3, 0.5f);
  const f2 v = vmul(r, r);
  const f2 p = vfma(q, v, r);
  const f2 y = vadd(Th, vfma(Th, p, Tl));
  f2 o;
  // exponent += n div 1024: bits(t) = 0x4B400000 + n and the low 19 bits of 0x4B400000 are zero, so
  // ((n >> 10) << 23) == (bits(t) & ~1023) * 2^13 (mod 2^32)
  unsigned ox, oy;
  asm("{ .reg .b32 m; and.b32 m, %1, 0xFFFFFC00; mad.lo.u32 %0, m, 8192, %2; }" : "=r"(ox) : "r"(bx), "r"(__float_as_int(y.x)));
  asm("{ .reg .b32 m; and.b32 m, %1, 0xFFFFFC00; mad.lo.u32 %0, m, 8192, %2; }" : "=r"(oy) : "r"(by), "r"(__float_as_int(y.y)));
  o.x = __int_as_float((int)ox); o.y = __int_as_float((int)oy);
  return o;
}
// spec32_exp_tiny on both halves; |x| <= 2^-6
__device__ __forceinline__ f2 vexp32_tiny(f2 x) {
  const f2 h = vadd(x, 1.0f);
  const f2 e = vsub(x, vadd(h, -1.0f));
  const f2 v = vmul(x, x);
  f2 u = vfma(x, 0x1.111112p-7f, 0x1.555556p-5f);
  u = vfma(u, x, e32::kC3);
  u = vfma(u, x, 0.5f);
  return vadd(h, vfma(v, u, e));
}
__device__ __forceinline__ f2 vqsat(f2 p, f2 t, float a, float b, f2 one) {   // getqvs / getqvi inside the window (f90:570-598)
  const f2 x = vdiv_fast(vmul(vadd(t, -273.15f), a), vadd(t, -b));
  const f2 es = vmul(vexp32_core(x), 611.2f);
  return vdiv_fast(vmul(es, cc::eps_q), vsubx(p, es, one));
}

// moist_arg<M, ICE, true> of cape_kernel.cuh on both halves (same operations in the same order)
template <bool ICE, bool PSEUDO>
__device__ __forceinline__ f2 vmoist_arg(f2 t2, f2 p2, f2 qt, f2 t1, f2 qv1, f2 ql1, f2 qi1, f2 logp, f2& qv2, f2& ql2, f2& qi2, f2 one) {
  f2 fice;
  if (ICE) {
    f2 fl = vdiv_fast(vadd(t2, -233.15f), splat(273.15f - 233.15f));
    fl.x = fmaxf(fminf(fl.x, 1.0f), 0.0f);
    fl.y = fmaxf(fminf(fl.y, 1.0f), 0.0f);
    fice = vsub(splat(1.0f), fl);
    const f2 qs = vaddx(vmul(fl, vqsat(p2, t2, 17.67f, 29.65f, one)), vmul(fice, vqsat(p2, t2, 21.8745584f, 7.66f, one)), one);
    qv2 = make_float2(fminf(qt.x, qs.x), fminf(qt.y, qs.y));
    f2 w = vmul(fice, vsub(qt, qv2));
    qi2 = make_float2(fmaxf(w.x, 0.0f), fmaxf(w.y, 0.0f));
    w = vsub(vsub(qt, qv2), qi2);
    ql2 = make_float2(fmaxf(w.x, 0.0f), fmaxf(w.y, 0.0f));
  } else {
    const f2 qs = vqsat(p2, t2, 17.67f, 29.65f, one);
    qv2 = make_float2(fminf(qt.x, qs.x), fminf(qt.y, qs.y));
    qi2 = splat(0.0f);
    const f2 w = vsub(qt, qv2);
    ql2 = make_float2(fmaxf(w.x, 0.0f), fmaxf(w.y, 0.0f));
  }
  // The reference's layer means tbar = 0.5 (t1 + t2), qvbar, qlbar, qibar are not formed: a multiplication by 0.5 is
  // exact, so it commutes with every rounding that follows — rv * (0.5 s) and (0.5 rv) * s are the same real number, hence
  // the same binary32; lv1 - lv2 * tbar is half of 2 lv1 - lv2 * s; and in lhv * dql / (cpm * tbar) numerator and
  // denominator are both doubled, which leaves the correctly rounded quotient unchanged.  (The window admits the
  // sub-step only when no product here can be subnormal: 180 <= s_t <= 800, qt = 0 or qt >= 1e-18.)
  const f2 st = vadd(t1, t2);
  const f2 sq = vadd(qv1, qv2);
  const f2 sl = PSEUDO ? ql2 : vadd(ql1, ql2);                                 // ql1 = 0: 0 + x = x
  const f2 lhv2 = vaddx(vmul(st, -cc::lv2), 2.0f * cc::lv1, one);              // 2 (lv1 - lv2*tbar)
  const f2 rm = vaddx(vmul(sq, 0.5f * cc::rv), cc::rd, one);
  f2 cpm = vaddx(vmul(sl, 0.5f * cc::cpl), vaddx(vmul(sq, 0.5f * cc::cpv), cc::cp, one), one);   // (cp + cpv qvbar) + cpl qlbar
  const f2 dql = PSEUDO ? ql2 : vsub(ql2, ql1);
  f2 arg;
  if (ICE) {
    const f2 si = PSEUDO ? qi2 : vadd(qi1, qi2);
    const f2 lhs2 = vaddx(vmul(st, -cc::ls2), 2.0f * cc::ls1, one);
    cpm = vaddx(vmul(si, 0.5f * cc::cpi), cpm, one);
    const f2 den2 = vmul(cpm, st);                                             // 2 cpm tbar
    const f2 dqi = PSEUDO ? qi2 : vsub(qi2, qi1);
    arg = vadd(vdiv_fast(vmul(lhv2, dql), den2), vdiv_fast(vmul(lhs2, dqi), den2));
  } else {
    arg = vdiv_fast(vmul(lhv2, dql), vmul(cpm, st));
  }
  return vaddx(vmul(vadd(vdiv_fast(rm, cpm), -cc::rddcp), logp), arg, one);
}

// ---------------------------------------------------------------------------------------------------------
// per-column state of the ascent
struct Col2 {
  int64_t c;            // column index
  int ks, nk;           // first 3-D level used (1-based), levels of the assembled column
  int k;                // assembled index of the level the parcel is at
  int lev_next;         // absolute 0-based index of the 3-D level that tops the next layer
  bool live;            // the column has a result to write
  bool active;          // still ascending
  int st, iters, mulvl;
  float zout, cape, cin, narea, z, b2;
  float th2, pi2, p2, t2, qv2, ql2, qi2, qt;
  float prev_p, prev_pi, prev_thv;
};
struct Layer2 {         // what a column knows about the layer it is crossing
  bool in;              // the column takes part in this layer
  bool fastp;           // every pressure of the layer lies in [2, 5e6] Pa: unguarded log / pow / division in the prologue
  float tmax;           // upper end of the moist window's temperature range for this layer (< 0: no window)
  float dp, b1;
  int nloop;
  float cur_p, cur_pi, cur_thv;
};
struct Sub2 {           // start state of a sub-step
  float p1, t1, th1, qv1, ql1, qi1, logp;
  bool window;
};

template <class M, int SOURCE, bool P1D>
__device__ __forceinline__ void col2_init(const CapeArgs& a, int64_t c, Col2& C) {
  C.c = c; C.live = (c < a.ncol); C.active = false;
  C.st = 0; C.iters = 0; C.mulvl = 0; C.zout = 0.0f; C.cape = 0.0f; C.cin = 0.0f; C.narea = 0.0f; C.z = 0.0f; C.b2 = 0.0f;
  C.th2 = C.pi2 = C.p2 = C.t2 = C.qv2 = C.ql2 = C.qi2 = C.qt = 0.0f;
  C.prev_p = C.prev_pi = C.prev_thv = 0.0f; C.ks = 1; C.nk = 1; C.k = 1; C.lev_next = 0;
  if (!C.live) return;
  if (!(a.ts[c] > 0.0f)) { C.st = 1; return; }      // model_lev.f90:77,83-88 (degC gate): zeros, MUlvl 0
  int ks = a.start ? a.start[c] : 1;                // pressure_lev.f90:154-160
  if ((ks < 1 || ks > a.nlev) && a.more_levels) { C.st = 4; return; }   // start level outside the shipped part: needs the full column
  ks = ks < 1 ? 1 : (ks > a.nlev ? a.nlev : ks);
  C.ks = ks; C.nk = a.nlev - ks + 2;
  const Parcel P = select_source<M, SOURCE, P1D>(a, c, ks, C.nk);
  C.k = P.k; C.z = P.zk; C.mulvl = P.mulvl; C.zout = -999999.0f;
  C.th2 = P.th2; C.pi2 = P.pi2; C.p2 = P.p2; C.t2 = P.t2; C.qv2 = P.qv2; C.b2 = P.b2; C.qt = P.qv2;
  C.prev_p = P.prev.p; C.prev_pi = P.prev.pi; C.prev_thv = P.prev.thv;
  C.lev_next = ks + C.k - 2;
  C.active = C.k < C.nk;
  if (!C.active && a.more_levels) C.st = 4;         // the parcel starts on the last level shipped: the column is taller
}

// sorted execution (cape_sort.cuh): the state col2_init left in the column's record
__device__ __forceinline__ void col2_load(const CapeArgs& a, int64_t c, Col2& C) {
  C.c = c; C.live = (c >= 0 && c < a.ncol); C.active = false;
  C.st = 0; C.iters = 0; C.mulvl = 0; C.zout = 0.0f; C.cape = 0.0f; C.cin = 0.0f; C.narea = 0.0f; C.z = 0.0f; C.b2 = 0.0f;
  C.th2 = C.pi2 = C.p2 = C.t2 = C.qv2 = C.ql2 = C.qi2 = C.qt = 0.0f;
  C.prev_p = C.prev_pi = C.prev_thv = 0.0f; C.ks = 1; C.nk = 1; C.k = 1; C.lev_next = 0;
  if (!C.live) return;
  const int4 ri = a.sorted.rec_i[c];
  const float4 ra = a.sorted.rec_a[c];
  const float4 rb = a.sorted.rec_b[c];
  const float2 rc = a.sorted.rec_c[c];
  C.ks = ri.x; C.nk = a.nlev - ri.x + 2; C.k = ri.y; C.mulvl = ri.z;
  C.st = ri.w & 15; C.active = (ri.w & 16) != 0; C.zout = (ri.w & 32) ? -999999.0f : 0.0f;
  C.th2 = ra.x; C.pi2 = ra.y; C.p2 = ra.z; C.t2 = ra.w;
  C.qv2 = rb.x; C.qt = rb.x; C.b2 = rb.y; C.z = rb.z; C.prev_p = rb.w;
  C.prev_pi = rc.x; C.prev_thv = rc.y;
  C.lev_next = C.ks + C.k - 2;
}

// f90:403-415 — the layer's environment, sub-step count, and the per-layer windows
template <class M, bool P1D, bool ICE>
__device__ __forceinline__ void col2_layer_begin(const CapeArgs& a, Col2& C, Layer2& Y) {
  C.k = C.k + 1;
  const Env cur = load_env<M, P1D>(a, C.c, C.ks, C.k);
  Y.cur_p = cur.p; Y.cur_pi = cur.pi; Y.cur_thv = cur.thv;
  Y.b1 = C.b2;
  float dp = C.prev_p - cur.p;
  int nloop = 1;
  if (!(dp < a.pinc)) {
    const float r = dp / a.pinc;
    if (!(r < (float)cc::nloop_cap)) { C.st = 3; C.active = false; Y.in = false; return; }   // see cape_kernel.cuh
    nloop = 1 + (int)r;
    dp = dp / (float)nloop;
  }
  Y.dp = dp; Y.nloop = nloop;
  // p2 starts within a few ulps of prev_p and ends within a few ulps of cur_p
  const float plo = fminf(C.prev_p, cur.p), phi = fmaxf(C.prev_p, cur.p);
  Y.fastp = (plo >= 2.0f) && (phi <= 5e6f) && (C.p2 >= 2.0f) && (C.p2 <= 5e6f);
  Y.tmax = -1.0f;
  if (Y.fastp) {
    // es(T) = 0.3 p inverted with approximate math at (just below) the layer's lowest pressure: tmax only chooses
    // between two code paths with identical results, and a smaller p gives a smaller (safer) tmax
    const float lg = __logf((0.999f * plo) * (0.3f / 611.2f));
    Y.tmax = fminf(__fdividef(4826.5605f - 29.65f * lg, 17.67f - lg), 400.0f);
  }
}

// sub-step prologue (f90:417-434): shift the state, step the pressure, Exner function and ln(p2/p1)
template <bool PSEUDO>
__device__ __forceinline__ void col2_sub_shift(Col2& C, const Layer2& Y, Sub2& S) {
  S.p1 = C.p2; S.t1 = C.t2; S.th1 = C.th2; S.qv1 = C.qv2;
  S.ql1 = PSEUDO ? 0.0f : C.ql2;                   // pseudo-adiabats reset condensate each sub-step (f90:487-491)
  S.qi1 = PSEUDO ? 0.0f : C.qi2;
  C.p2 = C.p2 - Y.dp;
}
__device__ __forceinline__ bool col2_window(const Col2& C, const Layer2& Y, const Sub2& S) {
  return (Y.tmax >= 100.0f) && (S.t1 >= 90.0f) && (S.t1 <= 400.0f) && (C.qt == 0.0f || C.qt >= 1e-18f) && (C.qt <= 1.0f) &&
         (S.ql1 <= 1.0f) && (S.qi1 <= 1.0f) && (fabsf(S.th1) < CUDART_INF_F);
}
template <class M, bool PSEUDO>
__device__ __forceinline__ void col2_sub_begin(Col2& C, const Layer2& Y, Sub2& S) {
  col2_sub_shift<PSEUDO>(C, Y, S);
  if (Y.fastp && C.p2 >= 2.0f) {
    // all operands normal and far from the range ends: the unguarded cores give what M::pow / M::log / `/` give
    C.pi2 = __double2float_rn(spec_exp_core(__dmul_rn((double)cc::rddcp, spec_log_core((double)(C.p2 * cc::rp00)))));
    S.logp = __double2float_rn(spec_log_core((double)fdiv_fast(C.p2, S.p1)));
    S.window = col2_window(C, Y, S);
  } else {
    C.pi2 = M::pow(C.p2 * cc::rp00, cc::rddcp);
    S.logp = M::log(C.p2 / S.p1);
    S.window = false;
  }
}
// both columns of the thread at once: the four binary64 chains (two logs + one exp per column) are independent, and
// written in one basic block they interleave — the FP64 pipe's latency, which a single chain exposes (ncu r2k: half of
// the prologue's samples are fixed-latency waits), overlaps.  Same operations per column as col2_sub_begin.
template <class M, bool PSEUDO>
__device__ __forceinline__ void col2_sub_begin_pair(Col2& A, const Layer2& YA, Sub2& SA, Col2& B, const Layer2& YB, Sub2& SB) {
  col2_sub_shift<PSEUDO>(A, YA, SA);
  col2_sub_shift<PSEUDO>(B, YB, SB);
  if (YA.fastp && A.p2 >= 2.0f && YB.fastp && B.p2 >= 2.0f) {
    const double la = spec_log_core((double)(A.p2 * cc::rp00)), lb = spec_log_core((double)(B.p2 * cc::rp00));
    const double ga = spec_log_core((double)fdiv_fast(A.p2, SA.p1)), gb = spec_log_core((double)fdiv_fast(B.p2, SB.p1));
    const double ea = spec_exp_core(__dmul_rn((double)cc::rddcp, la)), eb = spec_exp_core(__dmul_rn((double)cc::rddcp, lb));
    A.pi2 = __double2float_rn(ea); B.pi2 = __double2float_rn(eb);
    SA.logp = __double2float_rn(ga); SB.logp = __double2float_rn(gb);
    SA.window = col2_window(A, YA, SA); SB.window = col2_window(B, YB, SB);
  } else {
    A.pi2 = M::pow(A.p2 * cc::rp00, cc::rddcp); SA.logp = M::log(A.p2 / SA.p1); SA.window = false;
    B.pi2 = M::pow(B.p2 * cc::rp00, cc::rddcp); SB.logp = M::log(B.p2 / SB.p1); SB.window = false;
  }
}

// the reference's loop verbatim (slow divisions, guarded exp): columns outside the windows.  Everything by value:
// a reference parameter of a non-inlined function would pin the caller's column state to local memory.
struct SubOut { float t2, th2, qv2, ql2, qi2; int i, st; };
template <class M, bool ICE>
__device__ __noinline__ SubOut col2_sub_general(float pi2, float p2, float qt, float t1, float th1, float qv1, float ql1, float qi1,
                                                float logp, int st) {
  SubOut o;
  o.st = st;
  int i = 0;
  float thlast = th1;
  for (;;) {
    i = i + 1;
    o.t2 = thlast * pi2;
    o.th2 = moist_body<M, ICE, false>(o.t2, p2, qt, t1, th1, qv1, ql1, qi1, logp, o.qv2, o.ql2, o.qi2);
    if (i > 100) { o.st = 2; break; }             // f90:464-474 lack of convergence
    if (fabsf(o.th2 - thlast) > cc::converge) thlast = thlast + 0.3f * (o.th2 - thlast);
    else break;
  }
  o.i = i;
  return o;
}

template <bool PSEUDO>
__device__ __forceinline__ void col2_sub_end(Col2& C, int i) {
  C.iters += i;
  if (C.iters > cc::iter_budget) C.st = 3;
  if (C.st) { C.active = false; return; }
  if (PSEUDO) { C.qt = C.qv2; C.ql2 = 0.0f; C.qi2 = 0.0f; }
}

// f90:501-558 — buoyancy, trapezoid CAPE / CIN with zero-crossing split, stop rule
__device__ __forceinline__ void col2_layer_end(Col2& C, const Layer2& Y, bool more_levels) {
  const float thv2 = C.th2 * (1.0f + cc::reps * C.qv2) / (1.0f + C.qv2 + C.ql2 + C.qi2);
  const float b1 = Y.b1;
  const float b2 = cc::g * (thv2 - Y.cur_thv) / Y.cur_thv;
  const float dz = -cc::cpdg * 0.5f * (Y.cur_thv + C.prev_thv) * (Y.cur_pi - C.prev_pi);
  float parea;
  if (b2 >= 0.0f && b1 < 0.0f) {
    const float frac = b2 / (b2 - b1);
    parea = 0.5f * b2 * dz * frac;
    C.narea = C.narea - 0.5f * b1 * dz * (1.0f - frac);
    C.cin = C.cin + C.narea;
    C.narea = 0.0f;
  } else if (b2 < 0.0f && b1 > 0.0f) {
    const float frac = b1 / (b1 - b2);
    parea = 0.5f * b1 * dz * frac;
    C.narea = -0.5f * b2 * dz * (1.0f - frac);
  } else if (b2 < 0.0f) {
    parea = 0.0f;
    C.narea = C.narea - 0.5f * dz * (b1 + b2);
  } else {
    parea = 0.5f * dz * (b1 + b2);
    C.narea = 0.0f;
  }
  C.cape = C.cape + fmax_(0.0f, parea);
  C.b2 = b2;
  C.z = C.z + dz;
  C.zout = C.z;
  C.prev_p = Y.cur_p; C.prev_pi = Y.cur_pi; C.prev_thv = Y.cur_thv;
  C.lev_next = C.lev_next + 1;
  if (Y.cur_p <= 10000.0f && b2 < 0.0f) C.active = false;         // f90:554-557
  else if (!(C.k < C.nk)) { C.active = false; if (more_levels) C.st = 4; }   // ran out of (shipped) levels while ascending
}

__device__ __forceinline__ void col2_store(const CapeArgs& a, const Col2& C) {
  if (!C.live) return;
  const bool failed = C.st >= 2;
  a.cape[C.c] = failed ? 0.0f : C.cape;
  a.cin[C.c] = failed ? 0.0f : C.cin;
  a.zout[C.c] = C.zout;
  a.mulvl[C.c] = C.mulvl;
  if (a.status) a.status[C.c] = C.st;
  if (a.n_iter) a.n_iter[C.c] = C.iters;
}

// The window loop of one sub-step for both halves (f90:436-474 inside the window): returns the last pass's values, per
// half the `left` counter at the last pass that ended with the half still moving, and whether a window was left.
struct WinOut { f2 t2, th2, qv2, ql2, qi2; int lx, ly; bool left_window; };
template <bool ICE, bool PSEUDO>
__device__ __forceinline__ WinOut window_loop(f2 pi2, f2 p2, f2 qt, f2 t1, f2 th1, f2 qv1, f2 ql1, f2 qi1, f2 logp, float tmx, f2 one) {
  WinOut o;
  f2 thlast = th1;
  // the window's accumulators are shared by the two halves (leaving it is rare, and the general loop is always right):
  // 90 <= every t2 <= the smaller tmax, every |arg| <= 2^-6
  float t_hi = 90.0f, t_lo = 400.0f, a_hi = 0.0f;
  int lx = 101, ly = 101;                     // `left` at the last pass that ended with the half still moving
  int left = 100;
  bool mx, my;
  do {
    o.t2 = vmul(thlast, pi2);
    t_hi = fmaxf(t_hi, fmaxf(o.t2.x, o.t2.y)); t_lo = fminf(t_lo, fminf(o.t2.x, o.t2.y));      // FMNMX3
    const f2 arg = vmoist_arg<ICE, PSEUDO>(o.t2, p2, qt, t1, qv1, ql1, qi1, logp, o.qv2, o.ql2, o.qi2, one);
    a_hi = fmaxf(a_hi, fmaxf(fabsf(arg.x), fabsf(arg.y)));
    o.th2 = vmul(th1, vexp32_tiny(arg));
    const f2 d = vsubx(o.th2, thlast, one);
    const f2 step = vmul(d, 0.3f);
    mx = fabsf(d.x) > cc::converge; my = fabsf(d.y) > cc::converge;
    // a half that has converged keeps its thlast: the following passes recompute its final pass unchanged
    if (mx) { thlast.x = thlast.x + step.x; lx = left; }
    if (my) { thlast.y = thlast.y + step.y; ly = left; }
    left = left - 1;
  } while ((mx || my) && left != 0);
  // pass k runs with left = 101 - k: the half took (101 - l) + 1 passes, or was still moving at pass 100 (l == 1)
  o.lx = lx; o.ly = ly;
  o.left_window = !(t_lo >= 90.0f) || !(t_hi <= tmx) || !(a_hi <= 0.015625f);
  return o;
}

#ifndef XC_CAPE2_THREADS
#define XC_CAPE2_THREADS 128
#endif
#ifndef XC_CAPE2_MIN_BLOCKS
#define XC_CAPE2_MIN_BLOCKS 5    // 96 registers.  Measured per ERA5 field, sorted execution (ms, whole call; profiles/r2l_lab_launch_geometry*.txt):
#endif                           // 5 x 128 threads 7.58; 4 x 128 (128 regs, no spills) 7.93; 6 x 128 (80 regs) 7.80; 9 x 64 / 6 x 96 / 3 x 192 at
                                 // 112 regs 7.73 / 8.06 / 8.58: bound by FMA-pipe throughput and issue, not by occupancy or the spills
// SORTED: the columns come in the order a.sorted.perm with their source parcels in records (cape_sort.cuh); SOURCE is
// then irrelevant (instantiated with 1 only)
template <class M, int SOURCE, int ADIABAT, bool P1D, bool SORTED>
__global__ void __launch_bounds__(XC_CAPE2_THREADS, XC_CAPE2_MIN_BLOCKS) cape_kernel2(const CapeArgs a) {
  exp32_smem_fill();
  constexpr bool ICE = (ADIABAT == 3 || ADIABAT == 4);
  constexpr bool PSEUDO = (ADIABAT == 1 || ADIABAT == 3);
  const int64_t c0 = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  if (c0 >= a.ncol) return;

  Col2 A, B;
  if (SORTED) {                                     // positions 2t, 2t+1 of the order, state from the records
    col2_load(a, a.sorted.perm[c0], A);
    col2_load(a, (c0 + 1 < a.ncol) ? a.sorted.perm[c0 + 1] : -1, B);
  } else {
    col2_init<M, SOURCE, P1D>(a, c0, A);
    col2_init<M, SOURCE, P1D>(a, c0 + 1, B);
  }

  for (int L = 0; L < a.nlev; ++L) {
    if (!(A.active || B.active)) break;
    Layer2 YA, YB;
    YA.in = A.active && A.lev_next == L;
    YB.in = B.active && B.lev_next == L;
    if (!(YA.in || YB.in)) continue;
    YA.nloop = 0; YB.nloop = 0;
    if (YA.in) col2_layer_begin<M, P1D, ICE>(a, A, YA);
    if (YB.in) col2_layer_begin<M, P1D, ICE>(a, B, YB);
    const int nmax = max(YA.in ? YA.nloop : 0, YB.in ? YB.nloop : 0);

    for (int n = 1; n <= nmax; ++n) {
      const bool doA = YA.in && A.active && n <= YA.nloop;
      const bool doB = YB.in && B.active && n <= YB.nloop;
      if (!(doA || doB)) break;
      Sub2 SA, SB;
      SA.window = false; SB.window = false;
      if (doA && doB) col2_sub_begin_pair<M, PSEUDO>(A, YA, SA, B, YB, SB);
      else if (doA) col2_sub_begin<M, PSEUDO>(A, YA, SA);
      else col2_sub_begin<M, PSEUDO>(B, YB, SB);
      bool genA = doA, genB = doB;
      int iA = 0, iB = 0;

      if (SA.window && SB.window) {
        // ---- the usual case after sorting: both columns inside their windows.  Operands packed without selects; when no
        // window was left and neither half ran into the cap, the sub-step's epilogue is straight-line code.
        const WinOut w = window_loop<ICE, PSEUDO>(make_float2(A.pi2, B.pi2), make_float2(A.p2, B.p2), make_float2(A.qt, B.qt),
                                                  make_float2(SA.t1, SB.t1), make_float2(SA.th1, SB.th1), make_float2(SA.qv1, SB.qv1),
                                                  make_float2(SA.ql1, SB.ql1), make_float2(SA.qi1, SB.qi1),
                                                  make_float2(SA.logp, SB.logp), fminf(YA.tmax, YB.tmax), a.one2);
        if (!w.left_window && w.lx != 1 && w.ly != 1) {
          A.t2 = w.t2.x; A.th2 = w.th2.x; A.qv2 = w.qv2.x; A.ql2 = w.ql2.x; A.qi2 = w.qi2.x;
          B.t2 = w.t2.y; B.th2 = w.th2.y; B.qv2 = w.qv2.y; B.ql2 = w.ql2.y; B.qi2 = w.qi2.y;
          col2_sub_end<PSEUDO>(A, 102 - w.lx);
          col2_sub_end<PSEUDO>(B, 102 - w.ly);
          continue;
        }
        // a window was left or a half is still moving after 100 passes: both columns take the reference's loop verbatim
      } else if (SA.window || SB.window) {
        // ---- one column inside its window: it is lifted in packed arithmetic with a copy of itself in the other half
        // (identical arithmetic, so the copy neither delays convergence nor matters)
        const bool wa = SA.window;
        const Col2& C = wa ? A : B;
        const Sub2& S = wa ? SA : SB;
        const WinOut w = window_loop<ICE, PSEUDO>(splat(C.pi2), splat(C.p2), splat(C.qt), splat(S.t1), splat(S.th1), splat(S.qv1),
                                                  splat(S.ql1), splat(S.qi1), splat(S.logp), wa ? YA.tmax : YB.tmax, a.one2);
        if (!w.left_window) {
          if (wa) {
            genA = false;
            A.t2 = w.t2.x; A.th2 = w.th2.x; A.qv2 = w.qv2.x; A.ql2 = w.ql2.x; A.qi2 = w.qi2.x;
            iA = 102 - w.lx;
            if (w.lx == 1) { iA = 101; A.st = 2; }    // the reference runs pass 101 and gives up there (f90:464-474)
          } else {
            genB = false;
            B.t2 = w.t2.x; B.th2 = w.th2.x; B.qv2 = w.qv2.x; B.ql2 = w.ql2.x; B.qi2 = w.qi2.x;
            iB = 102 - w.lx;
            if (w.lx == 1) { iB = 101; B.st = 2; }
          }
        }
      }
      if (genA) {
        const SubOut o = col2_sub_general<M, ICE>(A.pi2, A.p2, A.qt, SA.t1, SA.th1, SA.qv1, SA.ql1, SA.qi1, SA.logp, A.st);
        A.t2 = o.t2; A.th2 = o.th2; A.qv2 = o.qv2; A.ql2 = o.ql2; A.qi2 = o.qi2; A.st = o.st; iA = o.i;
      }
      if (genB) {
        const SubOut o = col2_sub_general<M, ICE>(B.pi2, B.p2, B.qt, SB.t1, SB.th1, SB.qv1, SB.ql1, SB.qi1, SB.logp, B.st);
        B.t2 = o.t2; B.th2 = o.th2; B.qv2 = o.qv2; B.ql2 = o.ql2; B.qi2 = o.qi2; B.st = o.st; iB = o.i;
      }
      if (doA) col2_sub_end<PSEUDO>(A, iA);
      if (doB) col2_sub_end<PSEUDO>(B, iB);
    }
    if (YA.in && A.st == 0) col2_layer_end(A, YA, a.more_levels != 0);
    if (YB.in && B.st == 0) col2_layer_end(B, YB, a.more_levels != 0);
  }
  col2_store(a, A);
  col2_store(a, B);
}

}  // namespace xc
