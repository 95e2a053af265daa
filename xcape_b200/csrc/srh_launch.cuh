// srh_launch.cuh — launchers defined in srh.cu
#pragma once
#include "srh_kernel.cuh"
namespace xc {
int launch_srh(const SrhArgs<float>& a, bool p1d, cudaStream_t s);
int launch_srh(const SrhArgs<double>& a, bool p1d, cudaStream_t s);
int launch_srh_tile(const SrhArgs<float>& a, bool p1d, cudaStream_t s);
int launch_srh_tile(const SrhArgs<double>& a, bool p1d, cudaStream_t s);
int launch_stdheight(const HeightArgs<float>& a, bool p1d, cudaStream_t s);
int launch_stdheight(const HeightArgs<double>& a, bool p1d, cudaStream_t s);
}  // namespace xc
