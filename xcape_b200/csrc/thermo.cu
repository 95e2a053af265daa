// thermo.cu — dew point from specific humidity: the step immediately upstream of the
// column kernels for archives that ship q instead of Td (ERA5; doc/tutorial.rst:19-23 leaves it to an
// external script and the fixtures carry an unused `q`).
//
// The reference has no such routine, so there is nothing to be bit-compatible with ("parity
// unpinned"): the formula is the exact inverse of the saturation law the CAPE kernel itself
// uses (getqvs, CAPE_CODE_model_lev.f90:570-581: es = 611.2 exp(17.67 (T-273.15)/(T-29.65)) Pa, Bolton
// 1980 eq. 10), so that a parcel's mixing ratio computed by the kernel from this Td is the q it came from:
//     r  = q / (1 - q)                      mixing ratio from specific humidity
//     e  = p r / (eps + r)                  vapour pressure, hPa        (eps = 287.04 / 461.5)
//     L  = ln(e / 6.112)
//     Td = 243.5 L / (17.67 - L)            degC
// Arithmetic in binary64 with the library's deterministic log (xc_math_spec.cuh); one pass, HBM-bound:
// 2 loads + 1 store per element.
#include <algorithm>

#include "xc_common.cuh"
#include "xc_math_spec.cuh"
#include "thermo.cuh"

namespace xc {

__device__ __forceinline__ double dewpoint_c(double p_hpa, double q, double q_min) {
  if (q < q_min) q = q_min;                      // spectral noise / stratospheric zeros (q_min = 0: keep as is)
  const double r = q / (1.0 - q);
  const double e = p_hpa * r / (287.04 / 461.5 + r);
  const double L = spec_log_d(e / 6.112);
  return 243.5 * L / (17.67 - L);
}

// LM: level-major [nlev][ncol] (blockIdx.y = level, no index arithmetic); else level-last [ncol][nlev].
template <class T, bool P1D, bool LM>
__global__ void __launch_bounds__(256) dewpoint_kernel(const T* __restrict__ p, const T* __restrict__ q, T* __restrict__ td,
                                                       int64_t ncol, int nlev, double q_min) {
  if (LM) {
    const int k = blockIdx.y;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double pk = P1D ? (double)__ldg(p + k) : 0.0;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncol; c += stride) {
      const int64_t i = (int64_t)k * ncol + c;
      td[i] = (T)dewpoint_c(P1D ? pk : (double)p[i], (double)q[i], q_min);
    }
  } else {
    const int64_t n = ncol * nlev, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const int k = (int)(i % nlev);
      td[i] = (T)dewpoint_c(P1D ? (double)__ldg(p + k) : (double)p[i], (double)q[i], q_min);
    }
  }
}

template <class T>
static int launch_t(const void* p, const void* q, void* td, int64_t ncol, int nlev, bool p1d, bool lm, double q_min, cudaStream_t s) {
  const int64_t per = lm ? ncol : ncol * nlev;
  const unsigned gx = (unsigned)std::min<int64_t>((per + 255) / 256, 148 * 16);
  const dim3 grid(gx, lm ? nlev : 1);
  const T *p_ = (const T*)p, *q_ = (const T*)q;
  T* o = (T*)td;
  if (p1d && lm) dewpoint_kernel<T, true, true><<<grid, 256, 0, s>>>(p_, q_, o, ncol, nlev, q_min);
  else if (p1d) dewpoint_kernel<T, true, false><<<grid, 256, 0, s>>>(p_, q_, o, ncol, nlev, q_min);
  else if (lm) dewpoint_kernel<T, false, true><<<grid, 256, 0, s>>>(p_, q_, o, ncol, nlev, q_min);
  else dewpoint_kernel<T, false, false><<<grid, 256, 0, s>>>(p_, q_, o, ncol, nlev, q_min);
  XC_LAUNCH_CHECK();
  return XCAPE_OK;
}

int launch_dewpoint(const void* p, const void* q, void* td, int dtype, int64_t ncol, int nlev, bool p1d, bool level_major,
                    double q_min, cudaStream_t s) {
  if (ncol <= 0) return XCAPE_OK;
  if (level_major && nlev > 65535) return fail(XCAPE_ERR_ARG, "nlev too large");
  return dtype == XCAPE_F64 ? launch_t<double>(p, q, td, ncol, nlev, p1d, level_major, q_min, s)
                            : launch_t<float>(p, q, td, ncol, nlev, p1d, level_major, q_min, s);
}

}  // namespace xc
