// relayout.cuh — launchers of the glue kernels in relayout.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
namespace xc {
// level-last [ncol][nlev] (dtype) -> level-major [nlev][ld] float32.  ld may be negative (out then points
// at the row of input level 0): the level axis is flipped on the way
int launch_transpose_cast(const void* in, int dtype, float* out, int64_t ncol, int nlev, int64_t ld, cudaStream_t s);
// level-last [ncol][nlev] -> level-major [nlev][ld], dtype preserved
int launch_transpose_same(const void* in, int dtype, void* out, int64_t ncol, int nlev, int64_t ld, cudaStream_t s);
// flat cast copy (dtype -> float32)
int launch_cast_copy(const void* in, int dtype, float* out, int64_t n, cudaStream_t s);
// out[i] = in[n-1-i], dtype preserved (n = nlev: one small CTA)
int launch_reverse_copy(const void* in, int dtype, void* out, int n, cudaStream_t s);
// core.py:286-289
int launch_pres_lev_pos(const void* p, const void* ps, int dtype, int64_t ncol, int nlev, int32_t* start, cudaStream_t s, int none_value = 1);
}  // namespace xc
