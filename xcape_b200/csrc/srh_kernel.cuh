// srh_kernel.cuh — fused storm-relative-helicity column kernel (one thread = one column).
//
// One upward pass over the column replaces three reference routines and the two full-size
// float64 height arrays they exchange (core.py:516-535, srh.py:41-61):
//   stdheight_ml / stdheight_pl   (stdheight_2D_model_lev.f90:74-160)    hypsometric AGL height, binary64
//   bunkers_calc_ml / _pl         (Bunkers_model_lev.f90:75-180, DINTERP2DZ :188-236)  binary32
//   DCALRELHL_ml / _pl            (SREH_model_lev.f90:61-129)            binary64
//
// Streaming formulation (DESIGN.md §SRH kernel):
//  * heights are produced level by level and consumed immediately;
//  * the 12 Bunkers sample heights (500..6000 m) are met in increasing order while heights
//    increase, so a single cursor replaces DINTERP2DZ's 24 top-down searches;
//  * the SRH sum  -sum[(u_k-c_u)dv_k - (v_k-c_v)du_k]  needs the storm motion c, which is only
//    known after the 6 km level; it is accumulated as three c-independent binary64 sums
//    S1 = sum(u_k dv_k - v_k du_k), S2 = sum dv_k, S3 = sum du_k and combined at the end
//    (differs from the reference's summation order by O(1e-13) m2/s2);
//  * levels above max(6 km, depth) are only checked for monotone pressure (loads, no math).
//    If pressure is not strictly decreasing anywhere (heights not increasing — invalid input
//    for the reference too) the column is redone by the EXACT path, which reproduces
//    DINTERP2DZ's "highest bracket wins" search and its orientation switch literally.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "xc_math_spec.cuh"

namespace xc {

template <class T>
struct SrhArgs {
  const T* __restrict__ p;      // P1D: [nlev]; else level-major [nlev][ld]
  const T* __restrict__ t;
  const T* __restrict__ td;
  const T* __restrict__ u;
  const T* __restrict__ v;
  const T* __restrict__ ps;     // [ncol]
  const T* __restrict__ ts;
  const T* __restrict__ tds;
  const T* __restrict__ us;
  const T* __restrict__ vs;
  const T* __restrict__ aglh;   // heights given (srh.srh signature): level-major [nlev][ld]; else nullptr
  const T* __restrict__ aglhs;  // [ncol] height of the surface level when aglh is given
  const int32_t* __restrict__ start;   // 1-based or nullptr
  int64_t ncol, ld;
  int64_t lev_stride, col_stride;      // element (lev, c) of a 3-D field at lev*lev_stride + c*col_stride
  int nlev;
  double depth, aglh0;
  double* __restrict__ srh_rm;
  double* __restrict__ srh_lm;
  float* __restrict__ rm;       // [2*ncol] or nullptr
  float* __restrict__ lm;
  float* __restrict__ mean6;
  int fast_heights;                  // precision == XCAPE_FAST: binary32 height chain
  int32_t* __restrict__ work_list;   // [ncol] scratch: columns deferred to the EXACT kernel
  int* __restrict__ work_count;      // zeroed by the launcher
};

// stdheight_2D_model_lev.f90:90-125 — binary64 arithmetic on single-precision literals
// (the SRH goldens pin this detail).
namespace hc {
constexpr double R = (double)287.04f, g = (double)-9.80665f, eps = (double)0.6219800858985514f;
constexpr double t0 = (double)273.15f, c1 = (double)6.112f, c2 = (double)53.49f, c3 = (double)5.09f;
}  // namespace hc

// binary64 quotient to <= 1 ulp: MUFU.RCP64H seed + two Newton steps (6 FP64 issue slots against
// ~20 for the IEEE-rounded `/`).  The SRH chain is compared with the reference at 1e-6 m2/s2
// (values ~1e2), i.e. it needs ~1e-9 relative accuracy, not a particular rounding.
__device__ __forceinline__ double ddiv_fast(double a, double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  r = __fma_rn(__fma_rn(-b, r, 1.0), r, r);
  r = __fma_rn(__fma_rn(-b, r, 1.0), r, r);
  const double q = a * r;
  return __fma_rn(__fma_rn(-b, q, a), r, q);
}

// Virtual temperature, stdheight_2D_model_lev.f90:104-125:
//   E = 6.112 exp(53.49 - 6808/Td - 5.09 ln Td),  w = eps E/(P-E),  Tv = T (w+eps)/(eps (1+w)).
// The last two lines are evaluated in the algebraically identical one-division form
//   Tv = T P / (P - (1-eps) E);  exp / log are the SPEC binary64 cores (relative error < 2^-44).
__device__ __forceinline__ double tvirt(double T, double Td, double P) {
  const double Tin = T + hc::t0;
  const double Tdin = Td + hc::t0;
  const double x = hc::c2 - ddiv_fast(6808.0, Tdin) - hc::c3 * spec_log_d(Tdin);
  const double E = hc::c1 * spec_exp_d(x);
  return ddiv_fast(Tin * P, P - (1.0 - hc::eps) * E);
}
// precision = XCAPE_FAST: the same chain in binary32 on the FP32/XU pipes (MUFU.LG2/EX2/RCP).  The
// SURVEY's probe (§8d) puts an all-binary32 height -> Bunkers -> SRH chain within 1.2e-3 m2/s2 of the
// reference — three orders inside max(1, 1e-4 rel); heights are still accumulated in binary64.
__device__ __forceinline__ float rcpf_(float b) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b)); return r; }
__device__ __forceinline__ float tvirt_f(float T, float Td, float P) {
  const float Tin = T + 273.15f, Tdin = Td + 273.15f;
  const float x = 53.49f - 6808.0f * rcpf_(Tdin) - 5.09f * __logf(Tdin);
  const float E = 6.112f * __expf(x);
  return Tin * P * rcpf_(P - (1.0f - 0.6219800858985514f) * E);
}
__device__ __forceinline__ double hyps_step_f(double Hp, float Tv, float Tvp, float P, float Pp) {
  const float dz = (287.04f * (0.5f * (Tv + Tvp)) * (1.0f / -9.80665f)) * __logf(P * rcpf_(Pp));
  return Hp + (double)dz;
}
// one hypsometric step, f90:143,152:  H = Hp + (R (Tv+Tvp)/2 / g) ln(P/Pp), with ln(P/Pp) = ln P - ln Pp
__device__ __forceinline__ double hyps_step(double Hp, double Tv, double Tvp, double lnP, double lnPp) {
  return Hp + (hc::R * ((Tv + Tvp) * 0.5) * (1.0 / hc::g)) * (lnP - lnPp);
}

__device__ __forceinline__ double interp1(double y1, double y3, double x1, double x2, double x3) {  // SREH_model_lev.f90:123-129
  if (x3 == x1) x1 = x1 - (double)0.01f;
  return y1 + ((y3 - y1) * ((x2 - x1) / (x3 - x1)));
}

struct SrhOut { double srm, slm; float rmu, rmv, lmu, lmv, m6u, m6v; };

__device__ __forceinline__ void bunkers_finish(float mu, float mv, float s1u, float s1v, float s2u, float s2v,
                                               float s12u, float s12v, float s13u, float s13v, SrhOut& o) {
  // Bunkers_model_lev.f90:128-177 (accumulators start from 0: SURVEY App. B-8)
  mu = mu / 13.0f; mv = mv / 13.0f;
  const float uu = ((0.0f + s12u) + s13u) / 2.0f, vu = ((0.0f + s12v) + s13v) / 2.0f;
  const float ud = ((0.0f + s1u) + s2u) / 2.0f, vd = ((0.0f + s1v) + s2v) / 2.0f;
  const float ushr = uu - ud, vshr = vu - vd;
  const float nrm = sqrtf(ushr * ushr + vshr * vshr);        // (ushr**2+vshr**2)**0.5
  o.rmu = mu + 7.5f * vshr / nrm;
  o.rmv = mv - 7.5f * ushr / nrm;
  o.lmu = mu - 7.5f * vshr / nrm;
  o.lmv = mv + 7.5f * ushr / nrm;
  o.m6u = mu; o.m6v = mv;
}

// Element (level lev, column c) of a 3-D field: level-major (lev_stride = ld, col_stride = 1) or the
// reference's level-last layout (lev_stride = 1, col_stride = nlev; strided, kept for generality — the
// API re-lays level-last input out first: a warp-cooperative shared-memory-tile reader of that layout
// was measured at 1.85-2.07 ms per HRRR field against 1.44 ms for relayout + this kernel).
template <class T>
__device__ __forceinline__ int64_t off3(const SrhArgs<T>& a, int64_t c, int lev) {
  return (int64_t)lev * a.lev_stride + c * a.col_stride;
}
template <class T, bool P1D>
__device__ __forceinline__ double ld_p(const SrhArgs<T>& a, int64_t c, int lev) {
  return (double)(P1D ? __ldg(a.p + lev) : a.p[off3(a, c, lev)]);
}

// Per-column state of the streaming pass.  EXACT: literal DINTERP2DZ search (all 13 samples kept,
// later brackets overwrite); HG: heights given; FH: binary32 height chain.
template <bool EXACT, bool HG, bool FH>
struct SrhState {
  double Tvp, Pp, lnPp, Hp, Hs, up, vp, S1, S2, S3;
  float upf, vpf, zpf, s1u, s1v, s2u, s2v, s12u, s12v, s13u, s13v, mu, mv;
  float su[EXACT ? 13 : 1], sv[EXACT ? 13 : 1];
  int j;
  bool found_top, mono, descending;

  // surface level: (ps, ts, tds) or the given surface height, and the 10 m wind
  __device__ __forceinline__ void init(double Ps, double Ts, double Tds, double hs_given, double aglh0, double us, double vs,
                                       float usf, float vsf) {
    Tvp = 0.0; Pp = 0.0; lnPp = 0.0;
    if (HG) {
      Hp = hs_given;
    } else {
      Tvp = FH ? (double)tvirt_f((float)Ts, (float)Tds, (float)Ps) : tvirt(Ts, Tds, Ps);
      Hp = aglh0; Pp = Ps; lnPp = FH ? 0.0 : spec_log_d(Ps);
    }
    Hs = Hp;
    up = us; vp = vs; upf = usf; vpf = vsf; zpf = (float)Hp;
    s1u = upf; s1v = vpf;
    s2u = s2v = s12u = s12v = s13u = s13v = -999999.0f;
    mu = 0.0f + upf; mv = 0.0f + vpf;          // running in-order sum (streaming path)
    j = 1;
    if (EXACT) {
      su[0] = upf; sv[0] = vpf;
      for (int i = 1; i < 13; ++i) { su[i] = -999999.0f; sv[i] = -999999.0f; }
    }
    found_top = false; mono = true; descending = false;
    S1 = S2 = S3 = 0.0;
  }

  __device__ __forceinline__ bool math_done() const { return !EXACT && found_top && j >= 13; }

  // levels above max(6 km, depth): only the monotonicity of pressure (heights) is checked
  __device__ __forceinline__ void tail(double PorH) {
    if (HG) { if (!(PorH > Hp)) mono = false; Hp = PorH; }
    else { if (!(PorH < Pp)) mono = false; Pp = PorH; }
  }

  // one level: PorH = pressure (hPa) or the given height; t, td in degC; winds in both precisions
  __device__ __forceinline__ void step(double PorH, double Tk, double Tdk, double uk, double vk, float ukf, float vkf,
                                       double depth) {
    double Tv = 0.0, lnP = 0.0, H;
    if (HG) {
      H = PorH;
      if (!(H > Hp)) mono = false;
    } else {
      const double P = PorH;
      if (!(P < Pp)) mono = false;
      if (FH) {
        Tv = (double)tvirt_f((float)Tk, (float)Tdk, (float)P);
        H = hyps_step_f(Hp, (float)Tv, (float)Tvp, (float)P, (float)Pp);
      } else {
        Tv = tvirt(Tk, Tdk, P);
        lnP = spec_log_d(P);
        H = hyps_step(Hp, Tv, Tvp, lnP, lnPp);                  // stdheight_2D_model_lev.f90:143,152
      }
      Pp = P;
    }
    const float zkf = (float)H;

    // ---- Bunkers samples (DINTERP2DZ, Bunkers_model_lev.f90:188-236) ----
    if (EXACT) {
      const float zlo = descending ? zkf : zpf, zhi = descending ? zpf : zkf;
      const float vlo_u = descending ? ukf : upf, vhi_u = descending ? upf : ukf;
      const float vlo_v = descending ? vkf : vpf, vhi_v = descending ? vpf : vkf;
      for (int q = 1; q < 13; ++q) {               // later (higher) brackets overwrite: == first hit of the top-down search
        const float h = 500.0f * (float)q;
        if (zlo <= h && zhi > h) {
          const float w2 = (h - zlo) / (zhi - zlo);
          const float w1 = (float)(1.0 - (double)w2);
          su[q] = w1 * vlo_u + w2 * vhi_u;
          sv[q] = w1 * vlo_v + w2 * vhi_v;
        }
      }
    } else {
      while (j < 13) {
        const float h = 500.0f * (float)j;
        if (!(zkf > h)) break;                     // sample j is above this layer
        float qu = -999999.0f, qv = -999999.0f;    // below the layer's base: never bracketed (VMSG)
        if (zpf <= h) {
          const float w2 = (h - zpf) / (zkf - zpf);
          const float w1 = (float)(1.0 - (double)w2);   // 1.D0 - W2 (f90:228)
          qu = w1 * upf + w2 * ukf;
          qv = w1 * vpf + w2 * vkf;
        }
        mu = mu + qu; mv = mv + qv;
        if (j == 1) { s2u = qu; s2v = qv; }
        if (j == 11) { s12u = qu; s12v = qv; }
        if (j == 12) { s13u = qu; s13v = qv; }
        ++j;
      }
    }

    // ---- SRH partial sums (DCALRELHL, SREH_model_lev.f90:86-114) ----
    if (!found_top) {
      double ue = uk, ve = vk;
      if (H > depth) {
        found_top = true;
        ue = interp1(uk, up, H, depth, Hp);
        ve = interp1(vk, vp, H, depth, Hp);
      }
      const double du = ue - up, dv = ve - vp;
      S1 = S1 + (ue * dv - ve * du);
      S2 = S2 + dv;
      S3 = S3 + du;
    }
    Hp = H; Tvp = Tv; lnPp = lnP; up = uk; vp = vk; upf = ukf; vpf = vkf; zpf = zkf;
  }

  // returns false if the column must be redone by the EXACT path
  __device__ __forceinline__ bool finish(SrhOut& o) {
    if (!EXACT) {
      if (!mono) return false;
      for (; j < 13; ++j) { mu = mu + -999999.0f; mv = mv + -999999.0f; }   // column top below the sample: VMSG enters the mean
    } else {
      mu = 0.0f; mv = 0.0f;
      for (int q = 0; q < 13; ++q) { mu = mu + su[q]; mv = mv + sv[q]; }
      s1u = su[0]; s1v = sv[0]; s2u = su[1]; s2v = sv[1];
      s12u = su[11]; s12v = sv[11]; s13u = su[12]; s13v = sv[12];
    }
    bunkers_finish(mu, mv, s1u, s1v, s2u, s2v, s12u, s12v, s13u, s13v, o);
    if (found_top) {
      o.srm = -(S1 - (double)o.rmu * S2 + (double)o.rmv * S3);
      o.slm = -(S1 - (double)o.lmu * S2 + (double)o.lmv * S3);
    } else {
      o.srm = 0.0; o.slm = 0.0;                    // ktop == 0: empty sum (SREH_model_lev.f90:97-103)
    }
    return true;
  }
};

// Per-thread driver over global memory (level-major: coalesced; level-last: strided, EXACT kernel only).
// Returns false when the column needs the EXACT path.
template <class T, bool P1D, bool EXACT, bool HG, bool FH = false>
__device__ __forceinline__ bool srh_column(const SrhArgs<T>& a, int64_t c, int ks, SrhOut& o) {
  const int n3 = a.nlev - ks + 1;                 // 3-D levels used
  SrhState<EXACT, HG, FH> st;
  st.init(HG ? 0.0 : (double)a.ps[c], HG ? 0.0 : (double)a.ts[c], HG ? 0.0 : (double)a.tds[c],
          HG ? (double)a.aglhs[c] : 0.0, a.aglh0, (double)a.us[c], (double)a.vs[c], (float)a.us[c], (float)a.vs[c]);
  if (EXACT) {
    // DINTERP2DZ orientation test Z(1) > Z(NZ) needs the top height first (f90:211-216)
    double H;
    if (HG) {
      H = (double)a.aglh[off3(a, c, ks - 1 + n3 - 1)];
    } else {
      SrhState<false, false, FH> pre;
      pre.init((double)a.ps[c], (double)a.ts[c], (double)a.tds[c], 0.0, a.aglh0, 0.0, 0.0, 0.0f, 0.0f);
      for (int i = 0; i < n3; ++i) {
        const int64_t off = off3(a, c, ks - 1 + i);
        pre.step(ld_p<T, P1D>(a, c, ks - 1 + i), (double)a.t[off], (double)a.td[off], 0.0, 0.0, 0.0f, 0.0f, 1e300);
      }
      H = pre.Hp;
    }
    st.descending = ((float)st.Hs > (float)H);
  }
  int i = 0;
  // software prefetch: the next level's loads are issued before this level's (long, FP64) step, so the
  // memory latency hides behind arithmetic instead of stalling the first use
  T pn = T(0), tn = T(0), tdn = T(0), un = T(0), vn = T(0);
  auto fetch = [&](int k) {
    const int64_t off = off3(a, c, ks - 1 + k);
    un = a.u[off]; vn = a.v[off];
    if (HG) pn = a.aglh[off];
    else { pn = (T)ld_p<T, P1D>(a, c, ks - 1 + k); tn = a.t[off]; tdn = a.td[off]; }
  };
  if (n3 > 0) fetch(0);
  for (; i < n3; ++i) {
    const T pc = pn, tc = tn, tdc = tdn, uin = un, vin = vn;
    if (i + 1 < n3) fetch(i + 1);
    if (HG) st.step((double)pc, 0.0, 0.0, (double)uin, (double)vin, (float)uin, (float)vin, a.depth);
    else st.step((double)pc, (double)tc, (double)tdc, (double)uin, (double)vin, (float)uin, (float)vin, a.depth);
    if (st.math_done()) { ++i; break; }            // nothing above can matter if p keeps decreasing
  }
  if (!EXACT)
    for (; i < n3; ++i) st.tail(HG ? (double)a.aglh[off3(a, c, ks - 1 + i)] : ld_p<T, P1D>(a, c, ks - 1 + i));
  return st.finish(o);
}

template <class T>
__device__ __forceinline__ void srh_store(const SrhArgs<T>& a, int64_t c, const SrhOut& o) {
  a.srh_rm[c] = o.srm; a.srh_lm[c] = o.slm;
  if (a.rm) { a.rm[2 * c] = o.rmu; a.rm[2 * c + 1] = o.rmv; }
  if (a.lm) { a.lm[2 * c] = o.lmu; a.lm[2 * c + 1] = o.lmv; }
  if (a.mean6) { a.mean6[2 * c] = o.m6u; a.mean6[2 * c + 1] = o.m6v; }
}

// Streaming pass over level-major input.  Columns that need the literal DINTERP2DZ search
// (non-monotone pressure / heights) are appended to a work list and left to srh_exact_kernel, so
// that the EXACT code's registers (13 + 13 samples, second height pass) never limit this kernel.
template <class T, bool P1D, bool HG, bool FH>
__global__ void __launch_bounds__(128) srh_kernel(const SrhArgs<T> a) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.ncol) return;
  int ks = a.start ? a.start[c] : 1;
  ks = ks < 1 ? 1 : (ks > a.nlev ? a.nlev : ks);
  SrhOut o;
  if (srh_column<T, P1D, false, HG, FH>(a, c, ks, o)) srh_store(a, c, o);
  else a.work_list[atomicAdd(a.work_count, 1)] = (int32_t)c;
}

template <class T, bool P1D, bool HG, bool FH>
__global__ void __launch_bounds__(128) srh_exact_kernel(const SrhArgs<T> a) {
  const int n = *a.work_count;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int64_t c = a.work_list[k];
    int ks = a.start ? a.start[c] : 1;
    ks = ks < 1 ? 1 : (ks > a.nlev ? a.nlev : ks);
    SrhOut o;
    srh_column<T, P1D, true, HG, FH>(a, c, ks, o);
    srh_store(a, c, o);
  }
}

// Heights only: loop_stdheight_ml / loop_stdheight_pl1d.  Output strides let the caller pick
// level-major or level-last.
template <class T>
struct HeightArgs {
  const T* __restrict__ p; const T* __restrict__ t; const T* __restrict__ td;
  const T* __restrict__ ps; const T* __restrict__ ts; const T* __restrict__ tds;
  const int32_t* __restrict__ start;
  int64_t ncol, ld;
  int nlev;
  double aglh0;
  double* __restrict__ h; int64_t h_col_stride, h_lev_stride;
  double* __restrict__ hs;
};

template <class T, bool P1D>
__global__ void __launch_bounds__(128) stdheight_kernel(const HeightArgs<T> a) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.ncol) return;
  int ks = a.start ? a.start[c] : 1;
  ks = ks < 1 ? 1 : (ks > a.nlev ? a.nlev : ks);
  for (int lev = 0; lev < ks - 1; ++lev) a.h[c * a.h_col_stride + lev * a.h_lev_stride] = -999999.0;   // pressure_lev.f90:85-87
  const double Ps = (double)a.ps[c];
  double Tvp = tvirt((double)a.ts[c], (double)a.tds[c], Ps), Hp = a.aglh0, lnPp = spec_log_d(Ps);
  a.hs[c] = a.aglh0;
  for (int lev = ks - 1; lev < a.nlev; ++lev) {
    const int64_t off = (int64_t)lev * a.ld + c;
    const double P = (double)(P1D ? __ldg(a.p + lev) : a.p[off]);
    const double Tv = tvirt((double)a.t[off], (double)a.td[off], P);
    const double lnP = spec_log_d(P);
    const double H = hyps_step(Hp, Tv, Tvp, lnP, lnPp);
    a.h[c * a.h_col_stride + lev * a.h_lev_stride] = H;
    Hp = H; Tvp = Tv; lnPp = lnP;
  }
}

}  // namespace xc
