// srh_tile.cuh — the fused SRH kernel reading the REFERENCE layout (level-last: [ncol][nlev], each column
// contiguous; core.py:44-50) directly.
//
// Why: through the public calc_srh the 3-D fields arrive level-last.  Re-laying five fields out first
// (transpose_cast_kernel x 5) moves 2 x the field through HBM before the 0.58 ms SRH kernel reads it a third time:
// 1.36 ms per HRRR field = 0.22 of HBM peak, more than half of it relayout.
// Here a CTA owns a tile of 128 consecutive columns — for each field one contiguous run of 128 x nlev elements —
// and walks it in chunks of KC levels: all threads copy the chunk (for every column a 32-byte piece, read by KC
// consecutive lanes) into shared memory transposed, then thread c runs the streaming SRH step of srh_kernel.cuh
// on its column's KC levels out of shared memory.  Shared memory holds two chunks (the next one arrives by cp.async
// while this one is computed; 42 KB), not the column, so five CTAs stay resident per SM and the FP64 chain keeps its
// latency hiding; every input byte crosses HBM once, and
// levels above the last one any column of the tile needs (6 km / depth) are not read at all (only their pressures,
// for the monotonicity check).  A level axis stored top first (ERA5 downloads) is walked backwards in place.
#pragma once
#include "srh_kernel.cuh"

namespace xc {

#ifndef XC_SRH_TILE_COLS
#define XC_SRH_TILE_COLS 128
#endif
constexpr int kTileCols = XC_SRH_TILE_COLS;

#ifndef XC_SRH_TILE_KC
#define XC_SRH_TILE_KC 4          // levels per chunk for 4-byte elements (16-byte pieces); half as many for 8-byte elements.  Measured per HRRR
                                  // field (ms, device path; profiles/r2p_lab_srh_tile_geometry.txt): KC 8 x 5 CTAs 1.31, 4 x 5 0.94, 4 x 6 0.87, 4 x 7 0.90, 4 x 8 1.10, 2 x 8 1.03
#endif
#ifndef XC_SRH_TILE_MIN_BLOCKS
#define XC_SRH_TILE_MIN_BLOCKS 6
#endif
template <class T> struct TileCfg { static constexpr int KC = (sizeof(T) == 4) ? XC_SRH_TILE_KC : XC_SRH_TILE_KC / 2; };
constexpr int kTilePad = 4;           // row stride 132: the transposing stores of a warp hit 32 distinct banks

// asynchronous copy (LDGSTS) of levels [k0, k0 + KC) of the tile's columns of one level-last field into s[kk][col];
// elements outside the tile / above the top level are not written (and never read)
template <class T> __device__ __forceinline__ void cp_async_elem(T* smem_dst, const T* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (sizeof(T) == 4) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
// `ls` = +1: levels stored surface first; -1: stored top first, `g` then points at the LAST stored level of column 0
// (the level axis is walked backwards, as everywhere else in the library; pieces stay contiguous, only reversed)
template <class T, int KC>
__device__ __forceinline__ void tile_load_async(const T* __restrict__ g, int64_t c0, int ncols, int nlev, int k0, int ls,
                                                T (*s)[kTileCols + kTilePad]) {
  const T* base = g + c0 * nlev;
#pragma unroll
  for (int e = threadIdx.x; e < kTileCols * KC; e += kTileCols) {
    const int col = e / KC, kk = e % KC;
    const int k = k0 + kk;
    if (col < ncols && k < nlev) cp_async_elem(&s[kk][col], base + (int64_t)col * nlev + k * ls);
  }
}

// level-last SrhArgs: a.p/t/td/u/v point at [ncol][nlev] arrays (P1D: a.p is [nlev]); a.ld unused.
// Two chunk buffers: the copies of chunk i+1 are in flight while chunk i is computed.
template <class T, bool P1D, bool FH>
__global__ void __launch_bounds__(kTileCols, XC_SRH_TILE_MIN_BLOCKS) srh_tile_kernel(const SrhArgs<T> a) {
  constexpr int KC = TileCfg<T>::KC;
  constexpr int W = kTileCols + kTilePad;
  __shared__ T sP[2][P1D ? 1 : KC][W];
  __shared__ T sT[2][KC][W];
  __shared__ T sTd[2][KC][W];
  __shared__ T sU[2][KC][W];
  __shared__ T sV[2][KC][W];
  const int64_t c0 = (int64_t)blockIdx.x * kTileCols;
  const int ncols = (int)min((int64_t)kTileCols, a.ncol - c0);
  const int tid = threadIdx.x;
  const int64_t c = c0 + tid;
  const bool live = tid < ncols;

  bool all_done = false;           // no column of the tile needs t / td / u / v any more
  auto issue = [&](int k0, int buf) {
    if (k0 < a.nlev) {
      const int ls = (int)a.lev_stride;                 // +1 / -1
      if (!P1D) tile_load_async<T, KC>(a.p, c0, ncols, a.nlev, k0, ls, sP[buf]);
      if (!all_done) {
        tile_load_async<T, KC>(a.t, c0, ncols, a.nlev, k0, ls, sT[buf]);
        tile_load_async<T, KC>(a.td, c0, ncols, a.nlev, k0, ls, sTd[buf]);
        tile_load_async<T, KC>(a.u, c0, ncols, a.nlev, k0, ls, sU[buf]);
        tile_load_async<T, KC>(a.v, c0, ncols, a.nlev, k0, ls, sV[buf]);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue(0, 0);

  int ks = 1;
  SrhState<false, false, FH> st;
  if (live) {
    ks = a.start ? a.start[c] : 1;
    ks = ks < 1 ? 1 : (ks > a.nlev ? a.nlev : ks);
    st.init((double)a.ps[c], (double)a.ts[c], (double)a.tds[c], 0.0, a.aglh0, (double)a.us[c], (double)a.vs[c],
            (float)a.us[c], (float)a.vs[c]);
  }
  bool math_done = !live;
  int buf = 0;
  for (int k0 = 0; k0 < a.nlev; k0 += KC, buf ^= 1) {
    issue(k0 + KC, buf ^ 1);                            // next chunk (an empty group past the top)
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    if (live) {
#pragma unroll 1
      for (int kk = 0; kk < KC; ++kk) {
        const int k = k0 + kk;
        if (k >= a.nlev) break;
        if (k < ks - 1) continue;                       // below the column's first level (pressure grids)
        const double P = (double)(P1D ? __ldg(a.p + k) : sP[buf][kk][tid]);
        if (!math_done) {
          // math_done is false only while every chunk so far carried all fields: all_done needs math_done of all columns
          const T uin = sU[buf][kk][tid], vin = sV[buf][kk][tid];
          st.step(P, (double)sT[buf][kk][tid], (double)sTd[buf][kk][tid], (double)uin, (double)vin, (float)uin, (float)vin, a.depth);
          math_done = st.math_done();
        } else {
          st.tail(P);
        }
      }
    }
    all_done = __syncthreads_and(math_done) != 0;       // also the barrier that frees this chunk's buffer
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (live) {
    SrhOut o;
    if (st.finish(o)) srh_store(a, c, o);
    else a.work_list[atomicAdd(a.work_count, 1)] = (int32_t)c;
  }
}

}  // namespace xc
