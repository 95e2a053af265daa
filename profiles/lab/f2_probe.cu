#include <cuda_runtime.h>
#include <cstdio>
__global__ void k_scalar(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4=a0+4,a5=a0+5,a6=a0+6,a7=a0+7;
  const float b = 1.0001f, c = 1e-4f;
  for (int i = 0; i < iters; ++i) {
    a0 = __fmaf_rn(a0, b, c); a1 = __fmaf_rn(a1, b, c); a2 = __fmaf_rn(a2, b, c); a3 = __fmaf_rn(a3, b, c);
    a4 = __fmaf_rn(a4, b, c); a5 = __fmaf_rn(a5, b, c); a6 = __fmaf_rn(a6, b, c); a7 = __fmaf_rn(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3+a4+a5+a6+a7;
}
__global__ void k_packed(float* out, int iters) {
  float2 a0 = make_float2(threadIdx.x * 1e-3f, 1.f), a1 = make_float2(a0.x + 1, 2.f), a2 = make_float2(a0.x + 2, 3.f), a3 = make_float2(a0.x + 3, 4.f);
  float2 b = make_float2(1.0001f, 1.0002f), c = make_float2(1e-4f, 2e-4f);
  for (int i = 0; i < iters; ++i) {
    a0 = __ffma2_rn(a0, b, c); a1 = __ffma2_rn(a1, b, c); a2 = __ffma2_rn(a2, b, c); a3 = __ffma2_rn(a3, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0.x + a1.x + a2.x + a3.x + a0.y + a1.y + a2.y + a3.y;
}
// mixed: packed add/mul/fma chain, dependent (latency probe)
__global__ void k_packed_dep(float* out, int iters) {
  float2 a = make_float2(threadIdx.x * 1e-3f, 1.f);
  float2 b = make_float2(1.0001f, 1.0002f), c = make_float2(1e-4f, 2e-4f);
  for (int i = 0; i < iters; ++i) { a = __ffma2_rn(a, b, c); a = __fadd2_rn(a, c); a = __fmul2_rn(a, b); a = __ffma2_rn(a, b, c); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a.x + a.y;
}
__global__ void k_scalar_dep(float* out, int iters) {
  float a = threadIdx.x * 1e-3f; const float b = 1.0001f, c = 1e-4f;
  for (int i = 0; i < iters; ++i) { a = __fmaf_rn(a, b, c); a = __fadd_rn(a, c); a = __fmul_rn(a, b); a = __fmaf_rn(a, b, c); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, blocks = 148 * 8, threads = 256;
  for (int rep = 0; rep < 2; ++rep) {
    float ms;
    cudaEventRecord(e0); k_scalar<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("scalar FFMA  x8 : %.3f ms  %.1f TFLOP/s\n", ms, 2.0 * 8 * iters * blocks * threads / ms * 1e-9);
    cudaEventRecord(e0); k_packed<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("packed FFMA2 x4 : %.3f ms  %.1f TFLOP/s\n", ms, 2.0 * 8 * iters * blocks * threads / ms * 1e-9);
    cudaEventRecord(e0); k_scalar_dep<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("scalar dependent chain (8 warps/SMSP): %.3f ms  %.2f cycles/instr/warp-equivalent\n", ms, ms * 1e-3 * 1.965e9 / (4.0 * iters));
    cudaEventRecord(e0); k_packed_dep<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("packed dependent chain (8 warps/SMSP): %.3f ms\n", ms);
  }
  k_scalar_dep<<<148*4, 32>>>(out, iters); cudaDeviceSynchronize();
  cudaEventRecord(e0); k_scalar_dep<<<148 * 4, 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("scalar dep, 1 warp/SMSP: %.2f cycles per instr\n", ms * 1e-3 * 1.965e9 / (4.0 * iters));
  cudaEventRecord(e0); k_packed_dep<<<148 * 4, 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  printf("packed dep, 1 warp/SMSP: %.2f cycles per instr\n", ms * 1e-3 * 1.965e9 / (4.0 * iters));
  return 0;
}
