// srh.cu — instantiations of the fused SRH kernel and the heights-only kernel.
// Compiled with -fmad=false so that the binary64 height chain and the binary32 Bunkers
// chain round operation by operation like the reference.
#include <algorithm>

#include "xc_common.cuh"
#include "srh_kernel.cuh"
#include "srh_launch.cuh"
#include "srh_tile.cuh"

namespace xc {

template <class T>
int launch_srh_t(const SrhArgs<T>& a, bool p1d, cudaStream_t s) {
  if (a.ncol <= 0) return XCAPE_OK;
  const unsigned blocks = (unsigned)((a.ncol + 127) / 128);
  if (a.ncol >= (int64_t)1 << 31) return fail(XCAPE_ERR_ARG, "srh: more than 2^31-1 columns per call");
  XC_CUDA(cudaMemsetAsync(a.work_count, 0, sizeof(int), s));
  const bool fh = a.fast_heights != 0;
  kernel_timer_begin(s);
  if (a.aglh) srh_kernel<T, false, true, false><<<blocks, 128, 0, s>>>(a);
  else if (p1d && fh) srh_kernel<T, true, false, true><<<blocks, 128, 0, s>>>(a);
  else if (p1d) srh_kernel<T, true, false, false><<<blocks, 128, 0, s>>>(a);
  else if (fh) srh_kernel<T, false, false, true><<<blocks, 128, 0, s>>>(a);
  else srh_kernel<T, false, false, false><<<blocks, 128, 0, s>>>(a);
  kernel_timer_end(s);
  XC_LAUNCH_CHECK();
  const unsigned eb = (unsigned)std::min<int64_t>(blocks, 148 * 4);   // grid-stride over the (usually empty) work list
  if (a.aglh) srh_exact_kernel<T, false, true, false><<<eb, 128, 0, s>>>(a);
  else if (p1d && fh) srh_exact_kernel<T, true, false, true><<<eb, 128, 0, s>>>(a);
  else if (p1d) srh_exact_kernel<T, true, false, false><<<eb, 128, 0, s>>>(a);
  else if (fh) srh_exact_kernel<T, false, false, true><<<eb, 128, 0, s>>>(a);
  else srh_exact_kernel<T, false, false, false><<<eb, 128, 0, s>>>(a);
  XC_LAUNCH_CHECK();
  return XCAPE_OK;
}
int launch_srh(const SrhArgs<float>& a, bool p1d, cudaStream_t s) { return launch_srh_t(a, p1d, s); }
int launch_srh(const SrhArgs<double>& a, bool p1d, cudaStream_t s) { return launch_srh_t(a, p1d, s); }

// level-last (reference-layout) input read in place: srh_tile.cuh.  a.lev_stride = 1, a.col_stride = nlev.
template <class T>
int launch_srh_tile_t(const SrhArgs<T>& a, bool p1d, cudaStream_t s) {
  if (a.ncol <= 0) return XCAPE_OK;
  if (a.ncol >= (int64_t)1 << 31) return fail(XCAPE_ERR_ARG, "srh: more than 2^31-1 columns per call");
  const unsigned blocks = (unsigned)((a.ncol + kTileCols - 1) / kTileCols);
  XC_CUDA(cudaMemsetAsync(a.work_count, 0, sizeof(int), s));
  const bool fh = a.fast_heights != 0;
  kernel_timer_begin(s);
  if (p1d && fh) srh_tile_kernel<T, true, true><<<blocks, kTileCols, 0, s>>>(a);
  else if (p1d) srh_tile_kernel<T, true, false><<<blocks, kTileCols, 0, s>>>(a);
  else if (fh) srh_tile_kernel<T, false, true><<<blocks, kTileCols, 0, s>>>(a);
  else srh_tile_kernel<T, false, false><<<blocks, kTileCols, 0, s>>>(a);
  kernel_timer_end(s);
  XC_LAUNCH_CHECK();
  const unsigned eb = (unsigned)std::min<int64_t>(blocks, 148 * 4);
  if (p1d && fh) srh_exact_kernel<T, true, false, true><<<eb, 128, 0, s>>>(a);
  else if (p1d) srh_exact_kernel<T, true, false, false><<<eb, 128, 0, s>>>(a);
  else if (fh) srh_exact_kernel<T, false, false, true><<<eb, 128, 0, s>>>(a);
  else srh_exact_kernel<T, false, false, false><<<eb, 128, 0, s>>>(a);
  XC_LAUNCH_CHECK();
  return XCAPE_OK;
}
int launch_srh_tile(const SrhArgs<float>& a, bool p1d, cudaStream_t s) { return launch_srh_tile_t(a, p1d, s); }
int launch_srh_tile(const SrhArgs<double>& a, bool p1d, cudaStream_t s) { return launch_srh_tile_t(a, p1d, s); }

template <class T>
int launch_height_t(const HeightArgs<T>& a, bool p1d, cudaStream_t s) {
  if (a.ncol <= 0) return XCAPE_OK;
  const unsigned blocks = (unsigned)((a.ncol + 127) / 128);
  if (p1d) stdheight_kernel<T, true><<<blocks, 128, 0, s>>>(a);
  else stdheight_kernel<T, false><<<blocks, 128, 0, s>>>(a);
  XC_LAUNCH_CHECK();
  return XCAPE_OK;
}
int launch_stdheight(const HeightArgs<float>& a, bool p1d, cudaStream_t s) { return launch_height_t(a, p1d, s); }
int launch_stdheight(const HeightArgs<double>& a, bool p1d, cudaStream_t s) { return launch_height_t(a, p1d, s); }

}  // namespace xc
