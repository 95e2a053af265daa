// relayout.cu — HBM-bound glue kernels in front of the column kernels:
//   * transpose_cast: level-last [ncol][nlev] (the reference's f2py layout, core.py:44-50)
//     -> level-major [nlev][ld] binary32, fused with the f2py float64->float32 down-cast
//     (SURVEY §8b "Ownership"); column-block shared-memory staging, both sides coalesced;
//   * cast_copy: dtype cast of already level-major / 1-D arrays;
//   * pres_lev_pos: core.py:286-289 (numpy masked argmin) evaluated in the input dtype.
#include "xc_common.cuh"
#include "relayout.cuh"

namespace xc {

// Column-block transpose.  A CTA owns TC consecutive columns: in the level-last input they are
// ONE contiguous run of TC*nlev elements, read with fully coalesced loads whatever nlev is
// (37, 50, 137 ...); the run is parked in shared memory with an odd row stride (no bank
// conflicts on the transposed read) and written out as nlev rows of TC consecutive columns.
constexpr int kTC = 64;
template <class T, class TO>
__global__ void __launch_bounds__(256) transpose_cast_kernel(const T* __restrict__ in, TO* __restrict__ out,
                                                             int64_t ncol, int nlev, int64_t ld) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TO* tile = reinterpret_cast<TO*>(smem_raw);
  const int S = nlev | 1;                                  // odd stride
  const int64_t c0 = (int64_t)blockIdx.x * kTC;
  const int nc = (int)min((int64_t)kTC, ncol - c0);
  const int nelem = nc * nlev;
  const T* src = in + c0 * nlev;
  for (int i = threadIdx.x; i < nelem; i += blockDim.x) {
    const int c = i / nlev, l = i - c * nlev;
    tile[c * S + l] = (TO)src[i];
  }
  __syncthreads();
  // write: consecutive threads -> consecutive columns of one level row
  for (int i = threadIdx.x; i < nlev * kTC; i += blockDim.x) {
    const int l = i / kTC, c = i - l * kTC;
    if (c < nc) out[(int64_t)l * ld + c0 + c] = tile[c * S + l];
  }
}

template <class T>
__global__ void __launch_bounds__(256) cast_copy_kernel(const T* __restrict__ in, float* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (float)in[i];
}

template <class T>
__global__ void __launch_bounds__(256) pres_lev_pos_kernel(const T* __restrict__ p, const T* __restrict__ ps,
                                                           int64_t ncol, int nlev, int32_t* __restrict__ start) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncol) return;
  const T s = ps[c];
  int best = 0;
  bool have = false;
  T bestd = 0;
  for (int k = 0; k < nlev; ++k) {
    const T d = s - __ldg(p + k);          // temp_index = p_s1d - p_2d          (core.py:286)
    if (!(d < (T)0)) {                     // masked_less(temp_index, 0)          (core.py:287)
      if (!have || d < bestd) { have = true; bestd = d; best = k; }   // argmin, first minimum
    }
  }
  start[c] = best + 1;                     // Fortran convention                  (core.py:289)
}

static inline unsigned grid_for(int64_t n, int threads, int64_t cap = 148 * 32) {
  int64_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (unsigned)b;
}

int launch_transpose_cast(const void* in, int dtype, float* out, int64_t ncol, int nlev, int64_t ld, cudaStream_t s) {
  if (ncol <= 0) return XCAPE_OK;
  const unsigned grid = (unsigned)((ncol + kTC - 1) / kTC);
  const size_t smem = (size_t)kTC * (nlev | 1) * sizeof(float);
  if (smem > 200 * 1024) return fail(XCAPE_ERR_ARG, "nlev too large for the relayout kernel (max ~790 levels)");
  if (dtype == XCAPE_F64) {
    XC_CUDA(cudaFuncSetAttribute(transpose_cast_kernel<double, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    transpose_cast_kernel<double, float><<<grid, 256, smem, s>>>((const double*)in, out, ncol, nlev, ld);
  } else {
    XC_CUDA(cudaFuncSetAttribute(transpose_cast_kernel<float, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    transpose_cast_kernel<float, float><<<grid, 256, smem, s>>>((const float*)in, out, ncol, nlev, ld);
  }
  XC_LAUNCH_CHECK();
  return XCAPE_OK;
}

int launch_transpose_same(const void* in, int dtype, void* out, int64_t ncol, int nlev, int64_t ld, cudaStream_t s) {
  if (ncol <= 0) return XCAPE_OK;
  const unsigned grid = (unsigned)((ncol + kTC - 1) / kTC);
  const size_t smem = (size_t)kTC * (nlev | 1) * esize(dtype);
  if (smem > 200 * 1024) return fail(XCAPE_ERR_ARG, "nlev too large for the relayout kernel");
  if (dtype == XCAPE_F64) {
    XC_CUDA(cudaFuncSetAttribute(transpose_cast_kernel<double, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    transpose_cast_kernel<double, double><<<grid, 256, smem, s>>>((const double*)in, (double*)out, ncol, nlev, ld);
  } else {
    XC_CUDA(cudaFuncSetAttribute(transpose_cast_kernel<float, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    transpose_cast_kernel<float, float><<<grid, 256, smem, s>>>((const float*)in, (float*)out, ncol, nlev, ld);
  }
  XC_LAUNCH_CHECK();
  return XCAPE_OK;
}

int launch_cast_copy(const void* in, int dtype, float* out, int64_t n, cudaStream_t s) {
  if (n <= 0) return XCAPE_OK;
  if (dtype == XCAPE_F64) cast_copy_kernel<double><<<grid_for(n, 256), 256, 0, s>>>((const double*)in, out, n);
  else cast_copy_kernel<float><<<grid_for(n, 256), 256, 0, s>>>((const float*)in, out, n);
  XC_LAUNCH_CHECK();
  return XCAPE_OK;
}

int launch_pres_lev_pos(const void* p, const void* ps, int dtype, int64_t ncol, int nlev, int32_t* start, cudaStream_t s) {
  if (ncol <= 0) return XCAPE_OK;
  const unsigned blocks = (unsigned)((ncol + 255) / 256);
  if (dtype == XCAPE_F64) pres_lev_pos_kernel<double><<<blocks, 256, 0, s>>>((const double*)p, (const double*)ps, ncol, nlev, start);
  else pres_lev_pos_kernel<float><<<blocks, 256, 0, s>>>((const float*)p, (const float*)ps, ncol, nlev, start);
  XC_LAUNCH_CHECK();
  return XCAPE_OK;
}

}  // namespace xc
