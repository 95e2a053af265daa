#include <cuda_runtime.h>
#include <cstdio>
// throughput of FFMA / FFMA2 with three distinct per-thread register operands (no uniform / immediate operands)
template <int V>
__global__ void k_scalar3(float* out, const float* in, int iters) {
  float a[8], b[8], c[8];
  for (int j = 0; j < 8; ++j) { a[j] = in[threadIdx.x + j]; b[j] = in[threadIdx.x + 8 + j]; c[j] = in[threadIdx.x + 16 + j]; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = __fmaf_rn(b[j], c[(j + V) & 7], a[j]);
  }
  float s = 0; for (int j = 0; j < 8; ++j) s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int V>
__global__ void k_packed3(float* out, const float2* in, int iters) {
  float2 a[4], b[4], c[4];
  for (int j = 0; j < 4; ++j) { a[j] = in[threadIdx.x + j]; b[j] = in[threadIdx.x + 8 + j]; c[j] = in[threadIdx.x + 16 + j]; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = __ffma2_rn(b[j], c[(j + V) & 3], a[j]);
  }
  float s = 0; for (int j = 0; j < 4; ++j) s += a[j].x + a[j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// dependent chain with 3-register operands, 1 chain per thread: latency under load
__global__ void k_packed_chain(float* out, const float2* in, int iters) {
  float2 a = in[threadIdx.x], b = in[threadIdx.x + 8], c = in[threadIdx.x + 16], d = in[threadIdx.x + 24];
  for (int i = 0; i < iters; ++i) { a = __ffma2_rn(b, c, a); a = __ffma2_rn(a, d, b); a = __ffma2_rn(c, a, d); a = __ffma2_rn(a, b, c); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a.x + a.y;
}
__global__ void k_scalar_chain(float* out, const float* in, int iters) {
  float a = in[threadIdx.x], b = in[threadIdx.x + 8], c = in[threadIdx.x + 16], d = in[threadIdx.x + 24];
  for (int i = 0; i < iters; ++i) { a = __fmaf_rn(b, c, a); a = __fmaf_rn(a, d, b); a = __fmaf_rn(c, a, d); a = __fmaf_rn(a, b, c); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
int main() {
  float *out, *in; cudaMalloc(&out, 148 * 16 * 256 * 4); cudaMalloc(&in, 4096 * 8); cudaMemset(in, 0, 4096 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, threads = 128;
  float ms;
#define RUN(name, call, flops) for (int r = 0; r < 2; ++r) { cudaEventRecord(e0); call; cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); } \
  printf("%-46s %.3f ms  %.1f TFLOP/s\n", name, ms, (flops) / ms * 1e-9);
  for (int wps = 2; wps <= 8; wps *= 2) {     // warps per SMSP
    const int blocks = 148 * wps;              // 128 threads = 4 warps per block -> wps blocks/SM = wps warps per SMSP
    printf("-- %d warps per SMSP\n", wps);
    RUN("scalar FFMA r,r,r x8 independent", (k_scalar3<1><<<blocks, threads>>>(out, in, iters)), 2.0 * 8 * iters * blocks * threads);
    RUN("packed FFMA2 r,r,r x4 independent", (k_packed3<1><<<blocks, threads>>>(out, (float2*)in, iters)), 2.0 * 8 * iters * blocks * threads);
    RUN("scalar FFMA dependent chain", (k_scalar_chain<<<blocks, threads>>>(out, in, iters)), 2.0 * 4 * iters * blocks * threads);
    RUN("packed FFMA2 dependent chain", (k_packed_chain<<<blocks, threads>>>(out, (float2*)in, iters)), 2.0 * 8 * iters * blocks * threads);
  }
  return 0;
}
