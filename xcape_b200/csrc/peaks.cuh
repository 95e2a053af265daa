// peaks.cuh — arithmetic-peak microbenchmarks (roofline denominators for the FP-bound kernels)
#pragma once
namespace xc {
int measure_peaks(int reps, double* fp32_tflops, double* fp64_tflops);
int measure_fp32_rrr(int reps, double* tflops);
}
