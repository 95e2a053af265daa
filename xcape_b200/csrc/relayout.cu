// relayout.cu — HBM-bound glue kernels in front of the column kernels:
//   * transpose_cast: level-last [ncol][nlev] (the reference's f2py layout, core.py:44-50)
//     -> level-major [nlev][ld] binary32, fused with the f2py float64->float32 down-cast
//     (SURVEY §8b "Ownership"); column-block shared-memory staging, both sides coalesced;
//   * cast_copy: dtype cast of already level-major / 1-D arrays;
//   * reverse_copy: flips a shared 1-D pressure axis stored top-first (XCAPE_LEVELS_TOP_FIRST);
//   * pres_lev_pos: core.py:286-289 (numpy masked argmin) evaluated in the input dtype.
#include "xc_common.cuh"
#include <atomic>
#include "relayout.cuh"

namespace xc {

// Column-block transpose.  A CTA owns TC consecutive columns: in the level-last input they are
// ONE contiguous run of TC*nlev elements, read with fully coalesced loads whatever nlev is
// (37, 50, 137 ...); the run is parked in shared memory with an odd row stride (no bank
// conflicts on the transposed read) and written out as nlev rows of TC consecutive columns.
constexpr int kTC = 64;

template <class T> struct Vec4 { T x, y, z, w; };

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
  // ---- read: the block's columns are one contiguous run; 4 elements per load when the run is
  // 4-element aligned (always for full blocks: kTC*nlev is a multiple of 4 and cudaMalloc'ed
  // bases are 256-byte aligned), one index division per 4 elements.
  const bool vec_in = ((reinterpret_cast<uintptr_t>(src) & (4 * sizeof(T) - 1)) == 0) && ((nelem & 3) == 0);
  if (vec_in) {
    const Vec4<T>* src4 = reinterpret_cast<const Vec4<T>*>(src);
    for (int q = threadIdx.x; q < (nelem >> 2); q += blockDim.x) {
      const Vec4<T> v = src4[q];
      const int i = q << 2;
      int c = i / nlev, l = i - c * nlev;
      tile[c * S + l] = (TO)v.x; if (++l == nlev) { l = 0; ++c; }
      tile[c * S + l] = (TO)v.y; if (++l == nlev) { l = 0; ++c; }
      tile[c * S + l] = (TO)v.z; if (++l == nlev) { l = 0; ++c; }
      tile[c * S + l] = (TO)v.w;
    }
  } else {
    for (int i = threadIdx.x; i < nelem; i += blockDim.x) {
      const int c = i / nlev, l = i - c * nlev;
      tile[c * S + l] = (TO)src[i];
    }
  }
  __syncthreads();
  // ---- write: one warp stores one level row of the block (kTC = 64 consecutive columns) per
  // instruction, two columns per lane, when the row start is 2-element aligned.
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const bool vec_out = (nc == kTC) && ((ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & (2 * sizeof(TO) - 1)) == 0);
  if (vec_out) {
    struct alignas(2 * sizeof(TO)) Pair { TO a, b; };
    for (int l = warp; l < nlev; l += nwarp) {
      Pair pr;
      pr.a = tile[(2 * lane) * S + l];
      pr.b = tile[(2 * lane + 1) * S + l];
      *reinterpret_cast<Pair*>(out + (int64_t)l * ld + c0 + 2 * lane) = pr;
    }
  } else {
    for (int l = warp; l < nlev; l += nwarp)
      for (int c = lane; c < nc; c += 32) out[(int64_t)l * ld + c0 + c] = tile[c * S + l];
  }
}

template <class T>
__global__ void __launch_bounds__(256) cast_copy_kernel(const T* __restrict__ in, float* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (float)in[i];
}

template <class T>
__global__ void reverse_copy_kernel(const T* __restrict__ in, T* __restrict__ out, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = in[n - 1 - i];
}

template <class T>
__global__ void __launch_bounds__(256) pres_lev_pos_kernel(const T* __restrict__ p, const T* __restrict__ ps,
                                                           int64_t ncol, int nlev, int32_t* __restrict__ start, int none_value) {
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
  // every level masked: argmin of an all-masked array is 0 -> level 1.  When only the lower part of
  // the column was shipped (api.cu), "no level found" is not known yet: none_value = 0 makes the CAPE kernel hand the
  // column back for a second pass with all levels
  start[c] = have ? best + 1 : none_value; // Fortran convention                  (core.py:289)
}

// Dynamic shared memory above the 48 KB default has to be allowed per function AND per device.  The attribute is
// set to the maximum this file ever asks for (200 KB), once per (function, device): setting the exact size on every
// launch would let two host threads with different nlev interleave set(A), set(B < A), launch(A) -> launch failure.
constexpr int kMaxRelayoutSmem = 200 * 1024;
template <class K>
cudaError_t allow_big_smem(K kernel) {
  static std::atomic<uint64_t> done{0};           // one bit per device ordinal (per template instantiation)
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const uint64_t bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxRelayoutSmem);
  if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
  return e;
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
    XC_CUDA(allow_big_smem(transpose_cast_kernel<double, float>));
    transpose_cast_kernel<double, float><<<grid, 256, smem, s>>>((const double*)in, out, ncol, nlev, ld);
  } else {
    XC_CUDA(allow_big_smem(transpose_cast_kernel<float, float>));
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
    XC_CUDA(allow_big_smem(transpose_cast_kernel<double, double>));
    transpose_cast_kernel<double, double><<<grid, 256, smem, s>>>((const double*)in, (double*)out, ncol, nlev, ld);
  } else {
    XC_CUDA(allow_big_smem(transpose_cast_kernel<float, float>));
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

int launch_reverse_copy(const void* in, int dtype, void* out, int n, cudaStream_t s) {
  if (n <= 0) return XCAPE_OK;
  if (dtype == XCAPE_F64) reverse_copy_kernel<double><<<1, 256, 0, s>>>((const double*)in, (double*)out, n);
  else reverse_copy_kernel<float><<<1, 256, 0, s>>>((const float*)in, (float*)out, n);
  XC_LAUNCH_CHECK();
  return XCAPE_OK;
}

int launch_pres_lev_pos(const void* p, const void* ps, int dtype, int64_t ncol, int nlev, int32_t* start, cudaStream_t s, int none_value) {
  if (ncol <= 0) return XCAPE_OK;
  const unsigned blocks = (unsigned)((ncol + 255) / 256);
  if (dtype == XCAPE_F64) pres_lev_pos_kernel<double><<<blocks, 256, 0, s>>>((const double*)p, (const double*)ps, ncol, nlev, start, none_value);
  else pres_lev_pos_kernel<float><<<blocks, 256, 0, s>>>((const float*)p, (const float*)ps, ncol, nlev, start, none_value);
  XC_LAUNCH_CHECK();
  return XCAPE_OK;
}

}  // namespace xc
