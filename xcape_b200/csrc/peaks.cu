// peaks.cu — measured FP32 / FP64 FMA peaks of the current device.
// MEASURED_PEAKS.json (driver-written) carries HBM and bf16-tensor peaks only; the CAPE kernel
// is bound by the FP32 / FP64 pipes, so its roofline denominator is measured here: 8
// independent FMA chains per thread (enough ILP to cover the 4-cycle pipe latency with 8
// resident warps per scheduler), grid = 148 SMs x 8 CTAs x 256 threads.
#include "xc_common.cuh"
#include "peaks.cuh"

namespace xc {

template <class T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters, T a, T b) {
  T x0 = (T)threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const T s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == (T)123456789) out[0] = s;   // never true; keeps the chains alive
}

// FFMA with three distinct per-thread register operands and no operand reuse — what a real kernel issues (the
// classic microbenchmark above feeds two of the three operands from uniform registers).  The register file cannot
// deliver three fresh operands per lane per cycle: this sustains ~0.70 of the nominal FP32 peak on B200, and the
// packed FFMA2 form (three register PAIRS) ~0.60 (profiles/lab/f3_probe.cu).
__global__ void __launch_bounds__(128) ffma_rrr_kernel(float* out, const float* in, int iters) {
  float a[8], b[8], c[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = in[threadIdx.x + j]; b[j] = in[threadIdx.x + 8 + j]; c[j] = in[threadIdx.x + 16 + j]; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = __fmaf_rn(b[j], c[(j + 1) & 7], a[j]);
  }
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += a[j];
  if (s == 123456789.0f) out[0] = s;
}

int measure_fp32_rrr(int reps, double* tflops) {
  if (reps < 1) reps = 1;
  cudaDeviceProp prop;
  int dev;
  XC_CUDA(cudaGetDevice(&dev));
  XC_CUDA(cudaGetDeviceProperties(&prop, dev));
  const int blocks = prop.multiProcessorCount * 8, threads = 128, iters = 20000;
  float *out, *in;
  XC_CUDA(cudaMalloc(&out, sizeof(float)));
  XC_CUDA(cudaMalloc(&in, 4096 * sizeof(float)));
  XC_CUDA(cudaMemset(in, 0, 4096 * sizeof(float)));
  cudaEvent_t e0, e1;
  XC_CUDA(cudaEventCreate(&e0));
  XC_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int r = 0; r < reps + 1; ++r) {
    XC_CUDA(cudaEventRecord(e0));
    ffma_rrr_kernel<<<blocks, threads>>>(out, in, iters);
    XC_LAUNCH_CHECK();
    XC_CUDA(cudaEventRecord(e1));
    XC_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    XC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
    if (r > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out); cudaFree(in);
  *tflops = best;
  return XCAPE_OK;
}

template <class T>
static int peak_of(int reps, int iters, double* tflops) {
  cudaDeviceProp prop;
  int dev;
  XC_CUDA(cudaGetDevice(&dev));
  XC_CUDA(cudaGetDeviceProperties(&prop, dev));
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  T* out;
  XC_CUDA(cudaMalloc(&out, sizeof(T)));
  cudaEvent_t e0, e1;
  XC_CUDA(cudaEventCreate(&e0));
  XC_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int r = 0; r < reps + 1; ++r) {
    XC_CUDA(cudaEventRecord(e0));
    fma_peak_kernel<T><<<blocks, threads>>>(out, iters, (T)0.999, (T)0.001);
    XC_LAUNCH_CHECK();
    XC_CUDA(cudaEventRecord(e1));
    XC_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    XC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
    if (r > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  *tflops = best;
  return XCAPE_OK;
}

int measure_peaks(int reps, double* fp32_tflops, double* fp64_tflops) {
  if (reps < 1) reps = 1;
  int rc;
  if (fp32_tflops && (rc = peak_of<float>(reps, 4096, fp32_tflops))) return rc;
  if (fp64_tflops && (rc = peak_of<double>(reps, 2048, fp64_tflops))) return rc;
  return XCAPE_OK;
}

}  // namespace xc
