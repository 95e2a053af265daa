// cape_sort.cuh — sorted execution of the faithful CAPE kernel.
//
// Why (profiles/divergence_sorted_model.py): the moist iteration of a sub-step runs until the slowest
// of a warp's 64 columns has converged, and a layer takes as many sub-steps as its thickest column needs.  On an
// ERA5-shape field in storage order a warp executes 1.17 x the mean passes of its columns (1.30 x when neighbouring
// columns are unrelated).  The pass counts of a column are a function of the parcel it lifts — the level it starts
// from, its theta-e (which names the moist adiabat it follows above the LCL) — and of the surface pressure (which sets
// the sub-step count of every layer on sigma grids, of the first layer on pressure grids).  Columns grouped by
// (start level, theta-e in 4 K bins, surface pressure) converge together: 1.06-1.07 x the mean, whatever the storage
// order was.  Columns that do not ascend at all (ts <= 0 degC gate, parcel on the top level) go to the end of their
// window, so whole warps of them leave at once instead of idling beside working lanes.
//
// Two ways to establish the order (the launcher in cape_faithful.cu picks; XCAPE_B200_SORT_MODE overrides):
//  * GLOBAL, pressure grids (a shared pressure axis of <= 64 levels): counting sort over the whole call by (start level,
//    theta-e in 0.1 K bins) — histogram by atomics in the source kernel, a two-kernel exclusive scan, an atomic scatter.
//    Positions inside a bin follow the atomics' order: grouping (and timing) may differ from run to run, results cannot.
//  * WINDOWS of 16384 consecutive columns, sigma grids: one CTA per window sorts (key << 32 | index) in shared memory
//    (bitonic network; deterministic).  Windows exist because the ascent kernel's per-level loads become gathers (a
//    32-byte sector for 4 bytes): the other seven columns of a sector must be lifted by CTAs running at the same time for
//    the sector to come from L2 instead of crossing HBM eight times.  Measured: a global order costs HRRR (50 sigma
//    levels x 3 fields) 15.8 -> 17.2 ms where windows give 14.8; ERA5 (<= 27 levels x 2 fields reached) gains 0.28 ms from the
//    global order over windows (profiles/r2m_lab_sort_modes.txt).
//
// Pipeline (all on the caller's stream, scratch from the caller):
//   1. cape_source_kernel   storage order, coalesced: gate, start level, source parcel (the scalar code of
//                           cape_kernel.cuh, run once per column) -> parcel record + 32-bit key (+ histogram)
//   2. the order            cape_scan_totals_kernel + cape_scan_kernel + cape_scatter_kernel, or cape_window_sort_kernel -> perm[]
//   3. cape_kernel2<SORTED> thread t lifts columns perm[2t], perm[2t+1] from their records
// The key only decides which columns share a warp; results are per-column and bit-identical to the unsorted run
// (tests/test_gpu_parity.py::test_sorted_execution_matches_storage_order, both orders).
#pragma once
#include "cape_kernel2.cuh"

namespace xc {

constexpr int kSortWindow = 16384;             // columns per window: 128 KB of shared memory for the 64-bit sort elements
constexpr int kSortThreads = 1024;

constexpr int kSortThetaBins = 4096;           // global mode: 0.1 K bins of theta-e over [200, 609.6) K
constexpr int kSortScanTile = 4096;            // bins per scan tile (1024 threads x int4)

struct SortBufs {
  int4* rec_i; float4* rec_a; float4* rec_b; float2* rec_c;
  uint32_t* key; int32_t* perm;
  float inv_tbin;                               // window mode: 1 / (theta-e bin width, K)
  uint32_t* hist;                               // global mode (counting sort over the whole call), else nullptr
  uint32_t* tile_total;                         // global mode: one total per scan tile
  int nbins, nbins_padded;                      // nlev * kSortThetaBins + 1 (last bin: no ascent), padded to the scan tile
};
inline int sort_nbins(int nlev) { return nlev * kSortThetaBins + 1; }
inline int sort_nbins_padded(int nlev) { return (sort_nbins(nlev) + kSortScanTile - 1) / kSortScanTile * kSortScanTile; }

inline size_t sort_align(size_t x) { return (x + 255) & ~(size_t)255; }
inline size_t sort_scratch_bytes(int64_t ncol, int nlev) {
  const size_t n = (size_t)ncol;
  return sort_align(16 * n) * 3 + sort_align(8 * n) + sort_align(4 * n) * 2 + sort_align(4 * (size_t)sort_nbins_padded(nlev)) +
         sort_align(4 * (size_t)(sort_nbins_padded(nlev) / kSortScanTile));
}
inline SortBufs sort_carve(void* blob, int64_t ncol, int nlev) {
  char* q = (char*)blob;
  const size_t n = (size_t)ncol;
  SortBufs b;
  b.rec_i = (int4*)q; q += sort_align(16 * n);
  b.rec_a = (float4*)q; q += sort_align(16 * n);
  b.rec_b = (float4*)q; q += sort_align(16 * n);
  b.rec_c = (float2*)q; q += sort_align(8 * n);
  b.key = (uint32_t*)q; q += sort_align(4 * n);
  b.perm = (int32_t*)q; q += sort_align(4 * n);
  b.hist = (uint32_t*)q; q += sort_align(4 * (size_t)sort_nbins_padded(nlev));
  b.tile_total = (uint32_t*)q;
  b.nbins = sort_nbins(nlev); b.nbins_padded = sort_nbins_padded(nlev);
  b.inv_tbin = 0.25f;
  return b;
}

// Bolton's theta-e of the parcel (p Pa, t K, q kg/kg) with approximate intrinsics: a grouping key, not a result
__device__ __forceinline__ float thetae_key(float p, float t, float q) {
  q = fminf(fmaxf(q, 1e-9f), 0.2f);
  const float e = __fdividef(q * p, 0.622f + q) * 0.01f;                       // vapour pressure, hPa
  const float tl = __fdividef(2840.0f, 3.5f * __logf(t) - __logf(e) - 4.805f) + 55.0f;
  return t * __powf(__fdividef(100000.0f, p), 0.2854f * (1.0f - 0.28f * q)) *
         __expf((__fdividef(3376.0f, tl) - 2.54f) * q * (1.0f + 0.81f * q));
}

// key: start level (8 bits) | theta-e bin from 200 K (10 bits; bin width 1 / inv_tbin K) | surface pressure in Pa / 8 (14 bits);
// 0xFFFFFFFF = no ascent.  NaN / out-of-range inputs clamp into some bin — any grouping is a valid grouping.
template <class M, int SOURCE, bool P1D>
__global__ void __launch_bounds__(128) cape_source_kernel(const CapeArgs a, const SortBufs b) {
  exp32_smem_fill();
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.ncol) return;
  Col2 C;
  col2_init<M, SOURCE, P1D>(a, c, C);
  const int flags = C.st | (C.active ? 16 : 0) | (C.zout != 0.0f ? 32 : 0);
  b.rec_i[c] = make_int4(C.ks, C.k, C.mulvl, flags);
  b.rec_a[c] = make_float4(C.th2, C.pi2, C.p2, C.t2);
  b.rec_b[c] = make_float4(C.qv2, C.b2, C.z, C.prev_p);
  b.rec_c[c] = make_float2(C.prev_pi, C.prev_thv);
  if (b.hist) {                                  // global mode: key = histogram bin (start level, theta-e in 0.1 K bins)
    uint32_t bin = (uint32_t)(b.nbins - 1);
    if (C.active) {
      const float x = (thetae_key(C.p2, C.t2, C.qv2) - 200.0f) * 10.0f;
      const int tb = (int)fminf(fmaxf(x, 0.0f), (float)(kSortThetaBins - 1));
      bin = (uint32_t)(min(max(C.lev_next, 0), a.nlev - 1) * kSortThetaBins + tb);
    }
    b.key[c] = bin;
    // the "no ascent" bin can hold a large share of the field (40 % of the global-mix field): one atomic per warp for it
    const unsigned am = __activemask();
    const unsigned im = __ballot_sync(am, !C.active);
    if (C.active) atomicAdd(b.hist + bin, 1u);
    else if ((threadIdx.x & 31) == (unsigned)(__ffs(im) - 1)) atomicAdd(b.hist + bin, (uint32_t)__popc(im));
    return;
  }
  uint32_t key = 0xFFFFFFFFu;
  if (C.active) {
    const float x = (thetae_key(C.p2, C.t2, C.qv2) - 200.0f) * b.inv_tbin;
    const uint32_t bin = (uint32_t)fminf(fmaxf(x, 0.0f), 1023.0f);             // NaN -> 0
    const uint32_t lev = (uint32_t)min(max(C.lev_next, 0), 255);
    const uint32_t psq = (uint32_t)fminf(fmaxf(a.ps[c] * 12.5f, 0.0f), 16383.0f);
    key = (lev << 24) | (bin << 14) | psq;
    if (key == 0xFFFFFFFFu) key = 0xFFFFFFFEu;
  }
  b.key[c] = key;
}

// global mode: exclusive prefix sum of hist[0 .. nbins_padded) in place, two launches of one CTA per tile of 4096 bins:
// tile totals, then every CTA adds up the totals of the tiles before its own (<= 138 values) and scans its tile.
// (One CTA walking all tiles took 47 us per call — 2 % of the GPU time of the blocked host path.)
__device__ __forceinline__ uint32_t block_sum_1024(uint32_t v, uint32_t* warp_part) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if (lane == 0) warp_part[w] = v;
  __syncthreads();
  uint32_t t = warp_part[lane];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
  __syncthreads();
  return t;                                     // the block total, in every thread
}
__global__ void __launch_bounds__(1024) cape_scan_totals_kernel(const uint32_t* __restrict__ hist, uint32_t* __restrict__ tile_total) {
  __shared__ uint32_t warp_part[32];
  const uint4 v = reinterpret_cast<const uint4*>(hist + (size_t)blockIdx.x * kSortScanTile)[threadIdx.x];
  const uint32_t t = block_sum_1024(v.x + v.y + v.z + v.w, warp_part);
  if (threadIdx.x == 0) tile_total[blockIdx.x] = t;
}
__global__ void __launch_bounds__(1024) cape_scan_kernel(uint32_t* __restrict__ hist, const uint32_t* __restrict__ tile_total) {
  __shared__ uint32_t warp_part[32];
  __shared__ uint32_t warp_sum[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t before = 0;                          // totals of the tiles before this one
  for (int i = threadIdx.x; i < (int)blockIdx.x; i += 1024) before += tile_total[i];
  const uint32_t carry = block_sum_1024(before, warp_part);
  uint32_t* tile = hist + (size_t)blockIdx.x * kSortScanTile;
  const uint4 v = reinterpret_cast<const uint4*>(tile)[threadIdx.x];
  const uint32_t s1 = v.x, s2 = s1 + v.y, s3 = s2 + v.z, s4 = s3 + v.w;
  uint32_t incl = s4;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) warp_sum[w] = incl;
  __syncthreads();
  if (w == 0) {
    const uint32_t ws = warp_sum[lane];
    uint32_t wi = ws;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += o;
    }
    warp_sum[lane] = wi - ws;                   // exclusive prefix of the warp totals
  }
  __syncthreads();
  const uint32_t excl = carry + warp_sum[w] + (incl - s4);
  reinterpret_cast<uint4*>(tile)[threadIdx.x] = make_uint4(excl, excl + s1, excl + s2, excl + s3);
}
// global mode: counting-sort scatter.  Positions inside a bin follow the order of the atomics and may differ from run
// to run — grouping (and therefore timing) only, never a result.
__global__ void __launch_bounds__(256) cape_scatter_kernel(const uint32_t* __restrict__ key, uint32_t* __restrict__ cursor,
                                                           int32_t* __restrict__ perm, int64_t ncol, uint32_t idle_bin) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncol) return;
  const uint32_t k = key[c];
  const unsigned am = __activemask();
  const unsigned im = __ballot_sync(am, k == idle_bin);          // the "no ascent" bin: one atomic per warp
  uint32_t pos;
  if (k == idle_bin) {
    const int lane = threadIdx.x & 31, leader = __ffs(im) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(cursor + k, (uint32_t)__popc(im));
    base = __shfl_sync(im, base, leader);
    pos = base + (uint32_t)__popc(im & ((1u << lane) - 1u));
  } else {
    pos = atomicAdd(cursor + k, 1u);
  }
  perm[pos] = (int32_t)c;
}

// one CTA sorts one window: elements (key << 32 | index in window), bitonic network in shared memory
__global__ void __launch_bounds__(kSortThreads) cape_window_sort_kernel(const uint32_t* __restrict__ key, int32_t* __restrict__ perm,
                                                                       int64_t ncol) {
  extern __shared__ unsigned long long sk[];
  const int64_t w0 = (int64_t)blockIdx.x * kSortWindow;
  const int count = (int)min((int64_t)kSortWindow, ncol - w0);
  for (int i = threadIdx.x; i < kSortWindow; i += kSortThreads)
    sk[i] = (i < count) ? (((unsigned long long)key[w0 + i] << 32) | (unsigned)i) : ~0ull;
  __syncthreads();
  for (int k = 2; k <= kSortWindow; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < kSortWindow / 2; t += kSortThreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i + j;
        const unsigned long long x = sk[i], y = sk[l];
        const bool asc = (i & k) == 0;
        if ((x > y) == asc) { sk[i] = y; sk[l] = x; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < count; i += kSortThreads) perm[w0 + i] = (int32_t)(w0 + (int64_t)(sk[i] & 0xFFFFFFFFull));
}

}  // namespace xc
