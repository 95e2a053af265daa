// thermo.cuh — launcher of thermo.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
namespace xc {
// Td (degC) from p (hPa; [nlev] when p1d, else like q) and specific humidity q (kg/kg); q and td are dense
// 3-D fields of `dtype` in the same layout (level-major [nlev][ncol] or level-last [ncol][nlev]).
int launch_dewpoint(const void* p, const void* q, void* td, int dtype, int64_t ncol, int nlev, bool p1d, bool level_major,
                    double q_min, cudaStream_t s);
}  // namespace xc
