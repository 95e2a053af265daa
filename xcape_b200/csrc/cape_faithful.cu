// cape_faithful.cu — instantiations of the CAPE kernel with the FAITHFUL math policy
// (and, compiled a second time with -DXC_FAST_TU as build/cape_fast.o, with the FAST moist body).
// Compiled with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false: every binary32
// operation is an individually rounded IEEE operation, exactly like the reference built by
// gfortran -O3 on x86-64 (no FMA contraction).
#include "xc_common.cuh"
#include "xc_math_spec.cuh"
#include "cape_kernel.cuh"
#if !defined(XC_FAST_TU) && !defined(XC_FAST_RELAXED_TU)
#include "cape_kernel2.cuh"
#include "cape_sort.cuh"
#include <cstdlib>
#include <mutex>
#endif

namespace xc {

struct MathSpec {
  static constexpr bool kFastBody = false;
  static constexpr bool kSecant = false;
#if defined(XC_EXP64)
  // round-1 arithmetic: binary64 exp cores behind F2F conversions (kept for A/B timing only)
  static __device__ __forceinline__ float exp(float x) { return spec_expf(x); }
  static __device__ __forceinline__ float exp_in_range(float x) { return __double2float_rn(spec_exp_core((double)x)); }
  static __device__ __forceinline__ float exp_small(float x) { return spec_expf_small(x); }
  static __device__ __forceinline__ float exp_tiny(float x) { return __double2float_rn(spec_exp_small_core((double)x, true)); }
#else
  static __device__ __forceinline__ float exp(float x) { return spec32_expf(x); }
  // caller guarantees |x| <= 88 (same value as exp(x): only the range guard is dropped)
  static __device__ __forceinline__ float exp_in_range(float x) { return spec32_exp_core(x); }
  // the theta update's exp (f90:460-462); exp_tiny: caller guarantees |x| <= 2^-6 (same value as exp_small(x))
  static __device__ __forceinline__ float exp_small(float x) { return spec32_expf_small(x); }
  static __device__ __forceinline__ float exp_tiny(float x) { return spec32_exp_tiny(x); }
#endif
  static __device__ __forceinline__ float log(float x) { return spec_logf(x); }
  static __device__ __forceinline__ float pow(float x, float y) { return spec_powf(x, y); }
};

#if defined(XC_FAST_TU)
// Same SPEC prep arithmetic (so source selection / MU index stay bit-exact), FAST moist body,
// secant-accelerated sub-step solve.
struct MathFast : MathSpec { static constexpr bool kFastBody = true; static constexpr bool kSecant = true; };
using MathPolicy = MathFast;
#define XC_LAUNCH_NAME launch_cape_fast
#elif defined(XC_FAST_RELAXED_TU)
// FAST moist body, but the reference's own damped iteration (same pass counts as the reference).
struct MathFastRelaxed : MathSpec { static constexpr bool kFastBody = true; };
using MathPolicy = MathFastRelaxed;
#define XC_LAUNCH_NAME launch_cape_fast_relaxed
#else
using MathPolicy = MathSpec;
#define XC_LAUNCH_NAME launch_cape_faithful
#endif

template <int SOURCE, int ADIABAT, bool P1D>
static int launch(const CapeArgs& a, cudaStream_t s) {
#if !defined(XC_FAST_TU) && !defined(XC_FAST_RELAXED_TU)
  // faithful arithmetic: the two-column packed kernel (cape_kernel2.cuh); XCAPE_B200_CAPE_KERNEL=1 selects the
  // one-column kernel (same results bit for bit; kept for A/B timing)
  const char* which = getenv("XCAPE_B200_CAPE_KERNEL");      // read per call: tests flip it to cross-check the two kernels
  if (!(which && which[0] == '1')) {
    const int threads2 = XC_CAPE2_THREADS;
    const int64_t pairs = (a.ncol + 1) / 2;
    const int64_t blocks2 = (pairs + threads2 - 1) / threads2;
    if (blocks2 <= 0) return XCAPE_OK;
    if (a.sort_scratch) {
      // sorted execution (cape_sort.cuh): source parcels + keys, per-window sort, then the ascent in key order
      SortBufs b = sort_carve(a.sort_scratch, a.ncol, a.nlev);
      const char* tb = getenv("XCAPE_B200_SORT_TBIN");               // lab knob: theta-e bin width of the window key, K
      const float tbin = tb ? (float)atof(tb) : 4.0f;
      b.inv_tbin = 1.0f / (tbin > 0.01f ? tbin : 4.0f);
      // global order (counting sort over the whole call) where the gathers it causes are cheap — a shared pressure
      // axis (two gathered fields) of at most 64 levels; windows otherwise.  XCAPE_B200_SORT_MODE=window|global overrides.
      const char* sm = getenv("XCAPE_B200_SORT_MODE");
      const bool global_order = sm ? (sm[0] == 'g') : (P1D && a.nlev <= 64);
      if (global_order) XC_CUDA(cudaMemsetAsync(b.hist, 0, sizeof(uint32_t) * (size_t)b.nbins_padded, s));
      else b.hist = nullptr;
      cape_source_kernel<MathPolicy, SOURCE, P1D><<<(unsigned)((a.ncol + 127) / 128), 128, 0, s>>>(a, b);
      XC_LAUNCH_CHECK();
      if (global_order) {
        const unsigned ntiles = (unsigned)(b.nbins_padded / kSortScanTile);
        cape_scan_totals_kernel<<<ntiles, 1024, 0, s>>>(b.hist, b.tile_total);
        XC_LAUNCH_CHECK();
        cape_scan_kernel<<<ntiles, 1024, 0, s>>>(b.hist, b.tile_total);
        XC_LAUNCH_CHECK();
        cape_scatter_kernel<<<(unsigned)((a.ncol + 255) / 256), 256, 0, s>>>(b.key, b.hist, b.perm, a.ncol, (uint32_t)(b.nbins - 1));
        XC_LAUNCH_CHECK();
      } else {
        static std::once_flag smem_once[64];
        int dev = 0;
        XC_CUDA(cudaGetDevice(&dev));
        cudaError_t attr_err = cudaSuccess;
        std::call_once(smem_once[dev & 63], [&] {
          attr_err = cudaFuncSetAttribute(cape_window_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortWindow * 8);
        });
        XC_CUDA(attr_err);
        cape_window_sort_kernel<<<(unsigned)((a.ncol + kSortWindow - 1) / kSortWindow), kSortThreads, kSortWindow * 8, s>>>(b.key, b.perm, a.ncol);
        XC_LAUNCH_CHECK();
      }
      CapeArgs as = a;
      as.sorted.perm = b.perm; as.sorted.rec_i = b.rec_i; as.sorted.rec_a = b.rec_a; as.sorted.rec_b = b.rec_b; as.sorted.rec_c = b.rec_c;
      kernel_timer_begin(s);
      cape_kernel2<MathPolicy, 1, ADIABAT, P1D, true><<<(unsigned)blocks2, threads2, 0, s>>>(as);
      kernel_timer_end(s);
      XC_LAUNCH_CHECK();
      return XCAPE_OK;
    }
    kernel_timer_begin(s);
    cape_kernel2<MathPolicy, SOURCE, ADIABAT, P1D, false><<<(unsigned)blocks2, threads2, 0, s>>>(a);
    kernel_timer_end(s);
    XC_LAUNCH_CHECK();
    return XCAPE_OK;
  }
#endif
  const int threads = XC_CAPE_THREADS;
  const int64_t blocks = (a.ncol + threads - 1) / threads;
  if (blocks <= 0) return XCAPE_OK;
  kernel_timer_begin(s);
  cape_kernel<MathPolicy, SOURCE, ADIABAT, P1D><<<(unsigned)blocks, threads, 0, s>>>(a);
  kernel_timer_end(s);
  XC_LAUNCH_CHECK();
  return XCAPE_OK;
}

template <int SOURCE, bool P1D>
static int launch_adiabat(const CapeArgs& a, int adiabat, cudaStream_t s) {
  switch (adiabat) {
    case 1: return launch<SOURCE, 1, P1D>(a, s);
    case 2: return launch<SOURCE, 2, P1D>(a, s);
    case 3: return launch<SOURCE, 3, P1D>(a, s);
    case 4: return launch<SOURCE, 4, P1D>(a, s);
  }
  return fail(XCAPE_ERR_ARG, "adiabat must be 1..4");
}

#if !defined(XC_FAST_TU) && !defined(XC_FAST_RELAXED_TU)
// bytes of scratch the sorted execution of the faithful kernel wants for this call; 0 = run in storage order.
// XCAPE_B200_SORT=0 disables it, =1 forces it for any size (tests); by default calls of >= 8192 columns are sorted.
size_t cape_sort_scratch_bytes(int64_t ncol, int nlev) {
  const char* e = getenv("XCAPE_B200_SORT");
  const char* which = getenv("XCAPE_B200_CAPE_KERNEL");
  if ((which && which[0] == '1') || (e && e[0] == '0')) return 0;
  if (ncol >= (int64_t)1 << 31 || ncol < 2) return 0;
  if (!(e && e[0] == '1') && ncol < 8192) return 0;
  return sort_scratch_bytes(ncol, nlev);
}
int launch_exner_table(const float* p_hpa, float* pi, int nlev, cudaStream_t s) {
  exner_table_kernel<MathSpec><<<(nlev + 127) / 128, 128, 0, s>>>(p_hpa, pi, nlev);
  XC_LAUNCH_CHECK();
  return XCAPE_OK;
}
#endif

int XC_LAUNCH_NAME(const CapeArgs& a, int source, int adiabat, bool p1d, cudaStream_t s) {
  switch (source) {
    case 1: return p1d ? launch_adiabat<1, true>(a, adiabat, s) : launch_adiabat<1, false>(a, adiabat, s);
    case 2: return p1d ? launch_adiabat<2, true>(a, adiabat, s) : launch_adiabat<2, false>(a, adiabat, s);
    case 3: return p1d ? launch_adiabat<3, true>(a, adiabat, s) : launch_adiabat<3, false>(a, adiabat, s);
  }
  return fail(XCAPE_ERR_ARG, "source must be 1..3");
}

}  // namespace xc
