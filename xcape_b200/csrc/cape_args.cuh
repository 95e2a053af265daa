// cape_args.cuh — argument block of the CAPE kernels (shared by the launchers in api.cu and the kernel TUs).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace xc {

// Sorted execution of the faithful kernel (cape_sort.cuh): source parcels found in storage order and written to
// per-column records; the ascent kernel then takes its columns in the order `perm`, which groups columns whose parcels
// start on the same level with (almost) the same theta-e — lanes of a warp then need the same pass counts.
struct CapeSorted {
  const int32_t* __restrict__ perm;   // [ncol] column handled at position i; nullptr = storage order, no records
  const int4* __restrict__ rec_i;     // {ks, k, mulvl, st | active << 4 | zout_is_sentinel << 5}
  const float4* __restrict__ rec_a;   // {th2, pi2, p2, t2}
  const float4* __restrict__ rec_b;   // {qv2, b2, z, prev_p}
  const float2* __restrict__ rec_c;   // {prev_pi, prev_thv}
};

struct CapeArgs {
  const float* __restrict__ p;     // P1D: [nlev] hPa; else level-major [nlev][ld]
  const float* __restrict__ t;     // level-major [nlev][ld], degC
  const float* __restrict__ td;
  const float* __restrict__ ps;    // [ncol]
  const float* __restrict__ ts;
  const float* __restrict__ tds;
  const int32_t* __restrict__ start;   // [ncol] 1-based first level used, or nullptr (=1)
  int64_t ncol;
  int64_t ld;                      // distance (elements) between consecutive levels (negative: the level axis is stored top first)
  int64_t cs;                      // distance (elements) between consecutive columns: 1 = level-major; nlev = the reference's
                                   // level-last layout read in place (element (lev, c) at lev * ld + c * cs, ld = +-1)
  int nlev;
  float pinc;
  float ml_depth;
  float* __restrict__ cape;
  float* __restrict__ cin;
  float* __restrict__ zout;
  int32_t* __restrict__ mulvl;
  int32_t* __restrict__ status;    // nullable
  int32_t* __restrict__ n_iter;    // nullable: moist iterations executed (roofline work counter)
  int keep_secant;                 // fast mode only: keep the secant solve's value where the reference's iteration gives up
  int more_levels;                 // the 3-D arrays hold only the lowest `nlev` levels of a taller column (host path, see api.cu):
                                   // a column still ascending after the last one gets status 4 (internal) and is redone
  const float* __restrict__ pl_pi; // P1D only, nullable: Exner function of the nlev pressure levels, precomputed once per
                                   // call by exner_table_kernel with the same SPEC pow (bit-identical, saves a pow per level)
  float2 one2;                     // {1.0f, 1.0f}, set by the launcher: a multiplier the compiler cannot see through (cape_kernel2.cuh, vaddx)
  void* sort_scratch;              // faithful kernel only, nullable: cape_sort_scratch_bytes() bytes -> sorted execution
  CapeSorted sorted;               // filled by the launcher from sort_scratch
};

}  // namespace xc
