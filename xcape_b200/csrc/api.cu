// api.cu — the C ABI of libxcape_b200.so (include/xcape_b200.h): argument validation,
// canonicalisation of dtype/layout on the device, kernel dispatch, and the host-pointer
// path (column blocks staged H2D / computed / D2H on a ring of streams).
#include <algorithm>
#include <cstdint>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "xc_common.cuh"
#include "relayout.cuh"
#include "cape_args.cuh"
#include "srh_launch.cuh"
#include "peaks.cuh"
#include "thermo.cuh"

namespace xc {

thread_local std::string g_last_error;
std::atomic<int64_t> g_launches{0};
// per-thread event pair of xcape_cuda_time_kernels (events belong to the device they were created on)
thread_local bool t_time_kernels = false, t_timer_valid = false;
thread_local cudaEvent_t t_timer_ev[2] = {nullptr, nullptr};
thread_local int t_timer_dev = -1;
void kernel_timer_begin(cudaStream_t s) {
  if (!t_time_kernels) return;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  if (t_timer_dev != dev) {
    for (auto& e : t_timer_ev) { if (e) cudaEventDestroy(e); e = nullptr; }
    if (cudaEventCreate(&t_timer_ev[0]) != cudaSuccess || cudaEventCreate(&t_timer_ev[1]) != cudaSuccess) { t_timer_dev = -1; return; }
    t_timer_dev = dev;
  }
  t_timer_valid = false;
  cudaEventRecord(t_timer_ev[0], s);
}
void kernel_timer_end(cudaStream_t s) {
  if (!t_time_kernels || t_timer_dev < 0) return;
  t_timer_valid = cudaEventRecord(t_timer_ev[1], s) == cudaSuccess;
}
std::atomic<int64_t> g_redone{0};       // columns the host path handed back for a second pass with all levels

namespace {

// Stream-ordered scratch: cudaMallocAsync / cudaFreeAsync on the call's stream, so the
// device-pointer entry points never synchronise.
// The library owns one memory pool per device whose release threshold is "never": the
// default pool hands freed blocks back to the OS at the next synchronisation, and re-mapping
// ~300 MB of relayout scratch on every call cost ~5 ms per ERA5 field (r1a bench).
cudaMemPool_t pool_for_current_device() {
  static std::mutex mu;
  static cudaMemPool_t pools[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  if (!pools[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool = nullptr;
    if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    uint64_t keep = UINT64_MAX;
    if (const char* e = getenv("XCAPE_B200_POOL_KEEP_BYTES")) keep = strtoull(e, nullptr, 10);
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    pools[dev] = pool;
  }
  return pools[dev];
}

cudaError_t pool_alloc(void** q, size_t bytes, cudaStream_t s) {
  cudaMemPool_t pool = pool_for_current_device();
  if (pool) return cudaMallocFromPoolAsync(q, bytes, pool, s);
  return cudaMallocAsync(q, bytes, s);
}

struct Scratch {
  cudaStream_t s;
  std::vector<void*> ptrs;
  explicit Scratch(cudaStream_t st) : s(st) {}
  template <class T> cudaError_t alloc(T** p, size_t n) {
    void* q = nullptr;
    cudaError_t e = pool_alloc(&q, std::max<size_t>(n, 1) * sizeof(T), s);
    if (e == cudaSuccess) ptrs.push_back(q);
    *p = (T*)q;
    return e;
  }
  ~Scratch() { for (void* q : ptrs) cudaFreeAsync(q, s); }
};

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// `layout` = storage order of the 3-D fields (low byte) | XCAPE_LEVELS_TOP_FIRST
inline int base_layout(int layout) { return layout & 0xff; }
inline bool top_first(int layout) { return (layout & XCAPE_LEVELS_TOP_FIRST) != 0; }

int check_common(int64_t ncol, int nlev, int dtype, int layout, int mem) {
  if (ncol < 0) return fail(XCAPE_ERR_ARG, "ncol < 0");
  if (nlev < 1) return fail(XCAPE_ERR_ARG, "nlev < 1");
  if (dtype != XCAPE_F32 && dtype != XCAPE_F64) return fail(XCAPE_ERR_ARG, "dtype must be XCAPE_F32 or XCAPE_F64");
  if ((layout & ~(0xff | XCAPE_LEVELS_TOP_FIRST)) || (base_layout(layout) != XCAPE_LEVEL_LAST && base_layout(layout) != XCAPE_LEVEL_MAJOR))
    return fail(XCAPE_ERR_ARG, "bad layout");
  if (mem != XCAPE_MEM_HOST && mem != XCAPE_MEM_DEVICE) return fail(XCAPE_ERR_ARG, "bad mem");
  return XCAPE_OK;
}

// 3-D field (dtype, layout, leading dimension ld_in for level-major) -> level-major float32 with
// level 0 at the surface.  Zero-copy when already float32 level-major: a top-first level axis is
// then walked backwards (pointer at the last stored level, negative level stride); the relayout
// kernels flip it for free by writing through a negative leading dimension.
int canon3d(const void* in, int dtype, int layout, int64_t ncol, int nlev, int64_t ld_in,
            Scratch& sc, const float** out, int64_t* ld_out, cudaStream_t s) {
  const bool rev = top_first(layout);
  const int lay = base_layout(layout);
  if (dtype == XCAPE_F32 && lay == XCAPE_LEVEL_MAJOR) {
    *out = (const float*)in + (rev ? (int64_t)(nlev - 1) * ld_in : 0);
    *ld_out = rev ? -ld_in : ld_in;
    return XCAPE_OK;
  }
  float* buf;
  XC_CUDA(sc.alloc(&buf, (size_t)ncol * nlev));
  *out = buf; *ld_out = ncol;
  if (lay == XCAPE_LEVEL_LAST)
    return rev ? launch_transpose_cast(in, dtype, buf + (size_t)(nlev - 1) * ncol, ncol, nlev, -ncol, s)
               : launch_transpose_cast(in, dtype, buf, ncol, nlev, ncol, s);
  if (ld_in == ncol && !rev) return launch_cast_copy(in, dtype, buf, ncol * nlev, s);
  for (int k = 0; k < nlev; ++k) {   // flipped or strided level-major float64
    const int kin = rev ? nlev - 1 - k : k;
    int rc = launch_cast_copy((const char*)in + (size_t)kin * ld_in * esize(dtype), dtype, buf + (size_t)k * ncol, ncol, s);
    if (rc) return rc;
  }
  return XCAPE_OK;
}

// shared 1-D pressure axis in its own dtype, surface first (a flipped copy of nlev elements if needed)
int canon_p1d(const void* p, int dtype, int layout, int nlev, Scratch& sc, const void** out, cudaStream_t s) {
  if (!top_first(layout)) { *out = p; return XCAPE_OK; }
  char* buf;
  XC_CUDA(sc.alloc(&buf, (size_t)nlev * esize(dtype)));
  *out = buf;
  return launch_reverse_copy(p, dtype, buf, nlev, s);
}

int canon1d(const void* in, int dtype, int64_t n, Scratch& sc, const float** out, cudaStream_t s) {
  if (dtype == XCAPE_F32) { *out = (const float*)in; return XCAPE_OK; }
  float* buf;
  XC_CUDA(sc.alloc(&buf, (size_t)n));
  *out = buf;
  return launch_cast_copy(in, dtype, buf, n, s);
}

// ---------------------------------------------------------------------------------------
// CAPE, device pointers, asynchronous on stream s.  ld_in: level stride of level-major input.
// ---------------------------------------------------------------------------------------
int cape_device(const void* p, const void* t, const void* td, const void* ps, const void* ts, const void* tds,
                int64_t ncol, int nlev, int p_is_1d, int dtype, int layout, int64_t ld_in,
                int source, int adiabat, float ml_depth, float pinc, const int32_t* start_3d,
                float* cape, float* cin, int32_t* mulev, float* zmulev, int32_t* status, int32_t* n_iter,
                int precision, cudaStream_t s, bool more_levels = false) {
  if (ncol == 0) return XCAPE_OK;
  Scratch sc(s);
  CapeArgs a{};
  a.more_levels = more_levels ? 1 : 0;
  a.one2 = make_float2(1.0f, 1.0f);
  int rc;
  int64_t ld = ncol, ldp = ncol;
  const float* q;
  a.cs = 1;
  // The reference's level-last float32 layout is read IN PLACE by the sorted execution of the faithful kernel: there every
  // thread walks its own column, whose levels are contiguous, so the loads that are gathers on a level-major copy become
  // sector-sequential reads served by L1 — and the two or three relayout launches (2 x the field through HBM) disappear.
  // (Storage-order execution keeps the level-major copy: adjacent threads = adjacent columns = coalesced loads.)
  const char* direct_env = getenv("XCAPE_B200_CAPE_DIRECT");
  const bool direct = dtype == XCAPE_F32 && base_layout(layout) == XCAPE_LEVEL_LAST && precision == XCAPE_FAITHFUL &&
                      cape_sort_scratch_bytes(ncol, nlev) > 0 && !(direct_env && direct_env[0] == '0');
  if (direct) {
    const int64_t o = top_first(layout) ? nlev - 1 : 0;     // stored top first: start at each column's last stored level, stride -1
    a.t = (const float*)t + o; a.td = (const float*)td + o;
    ld = top_first(layout) ? -1 : 1; a.cs = nlev;
  } else {
    if ((rc = canon3d(t, dtype, layout, ncol, nlev, ld_in, sc, &q, &ld, s))) return rc; a.t = q;
    int64_t ld2 = ncol;
    if ((rc = canon3d(td, dtype, layout, ncol, nlev, ld_in, sc, &q, &ld2, s))) return rc; a.td = q;
    if (ld2 != ld) return fail(XCAPE_ERR_ARG, "internal: inconsistent leading dimensions");
  }
  if (p_is_1d) {
    const void* pv;
    if ((rc = canon_p1d(p, dtype, layout, nlev, sc, &pv, s))) return rc; p = pv;
    if ((rc = canon1d(p, dtype, nlev, sc, &q, s))) return rc; a.p = q;
  } else if (direct) {
    a.p = (const float*)p + (top_first(layout) ? nlev - 1 : 0);
  } else {
    if ((rc = canon3d(p, dtype, layout, ncol, nlev, ld_in, sc, &q, &ldp, s))) return rc; a.p = q;
    if (ldp != ld) return fail(XCAPE_ERR_ARG, "internal: inconsistent leading dimensions");
  }
  if ((rc = canon1d(ps, dtype, ncol, sc, &q, s))) return rc; a.ps = q;
  if ((rc = canon1d(ts, dtype, ncol, sc, &q, s))) return rc; a.ts = q;
  if ((rc = canon1d(tds, dtype, ncol, sc, &q, s))) return rc; a.tds = q;
  a.start = start_3d;
  if (p_is_1d && !start_3d) {
    int32_t* st;
    XC_CUDA(sc.alloc(&st, (size_t)ncol));
    if ((rc = launch_pres_lev_pos(p, ps, dtype, ncol, nlev, st, s, more_levels ? 0 : 1))) return rc;   // in the inputs' own dtype
    a.start = st;
  }
  if (p_is_1d) {          // Exner function of the shared pressure axis, once per call
    float* pl;
    XC_CUDA(sc.alloc(&pl, (size_t)nlev));
    if ((rc = launch_exner_table(a.p, pl, nlev, s))) return rc;
    a.pl_pi = pl;
  }
  a.ncol = ncol; a.ld = ld; a.nlev = nlev; a.pinc = pinc; a.ml_depth = ml_depth;
  a.cape = cape; a.cin = cin; a.zout = zmulev; a.mulvl = mulev; a.status = status; a.n_iter = n_iter;
  if (precision == XCAPE_FAITHFUL) {
    if (const size_t nb = cape_sort_scratch_bytes(ncol, nlev)) {    // sorted execution (cape_sort.cuh)
      char* blob;
      XC_CUDA(sc.alloc(&blob, nb));
      a.sort_scratch = blob;
    }
    return launch_cape_faithful(a, source, adiabat, p_is_1d != 0, s);
  }
  if (precision == XCAPE_FAST || precision == XCAPE_FAST_OPTIMISTIC) {
    a.keep_secant = (precision == XCAPE_FAST_OPTIMISTIC) ? 1 : 0;
    return launch_cape_fast(a, source, adiabat, p_is_1d != 0, s);
  }
  if (precision == XCAPE_FAST_RELAXED) return launch_cape_fast_relaxed(a, source, adiabat, p_is_1d != 0, s);
  return fail(XCAPE_ERR_ARG, "unknown precision mode");
}

// ---------------------------------------------------------------------------------------
// SRH / heights, device pointers, asynchronous on stream s.
// ---------------------------------------------------------------------------------------
// 3-D field -> level-major in its OWN dtype (the SRH chain consumes binary64 inputs as such:
// stdheight and SREH are double-precision routines).
int canon3d_same(const void* in, int dtype, int layout, int64_t ncol, int nlev, int64_t ld_in, Scratch& sc,
                 const void** out, int64_t* ld_out, cudaStream_t s) {
  const bool rev = top_first(layout);
  if (base_layout(layout) == XCAPE_LEVEL_MAJOR) {
    *out = (const char*)in + (rev ? (size_t)(nlev - 1) * ld_in * esize(dtype) : 0);
    *ld_out = rev ? -ld_in : ld_in;
    return XCAPE_OK;
  }
  char* buf;
  XC_CUDA(sc.alloc(&buf, (size_t)ncol * nlev * esize(dtype)));
  *out = buf; *ld_out = ncol;
  return rev ? launch_transpose_same(in, dtype, buf + (size_t)(nlev - 1) * ncol * esize(dtype), ncol, nlev, -ncol, s)
             : launch_transpose_same(in, dtype, buf, ncol, nlev, ncol, s);
}

template <class T>
int srh_device_t(const void* p, const void* t, const void* td, const void* u, const void* v, const void* ps,
                 const void* ts, const void* tds, const void* us, const void* vs, int64_t ncol, int nlev, int p_is_1d,
                 int dtype, int layout, int64_t ld_in, double depth, double aglh0, const int32_t* start_3d,
                 double* srh_rm, double* srh_lm, float* rm, float* lm, float* mean6, int precision, cudaStream_t s) {
  Scratch sc(s);
  SrhArgs<T> a{};
  a.fast_heights = (precision == XCAPE_FAST);
  int rc;
  const void* q;
  int64_t ld = ncol, l2 = ncol;
  // reference layout (level-last), either level order: read in place by the tile kernel (srh_tile.cuh) — no relayout.
  // XCAPE_B200_SRH_TILE=0 selects relayout + level-major kernel (A/B, tests).
  const char* tile_env = getenv("XCAPE_B200_SRH_TILE");
  const bool tile = base_layout(layout) == XCAPE_LEVEL_LAST && !(tile_env && tile_env[0] == '0');
  if (tile) {
    // stored top first: point at each column's last stored level and walk the level axis backwards (stride -1)
    const int64_t o = top_first(layout) ? nlev - 1 : 0;
    a.t = (const T*)t + o; a.td = (const T*)td + o; a.u = (const T*)u + o; a.v = (const T*)v + o;
    if (p_is_1d) { if ((rc = canon_p1d(p, dtype, layout, nlev, sc, &q, s))) return rc; p = q; a.p = (const T*)q; }   // flipped copy of nlev elements
    else a.p = (const T*)p + o;
    a.lev_stride = top_first(layout) ? -1 : 1; a.col_stride = nlev;
  } else {
    if ((rc = canon3d_same(t, dtype, layout, ncol, nlev, ld_in, sc, &q, &ld, s))) return rc; a.t = (const T*)q;
    if ((rc = canon3d_same(td, dtype, layout, ncol, nlev, ld_in, sc, &q, &l2, s))) return rc; a.td = (const T*)q;
    if ((rc = canon3d_same(u, dtype, layout, ncol, nlev, ld_in, sc, &q, &l2, s))) return rc; a.u = (const T*)q;
    if ((rc = canon3d_same(v, dtype, layout, ncol, nlev, ld_in, sc, &q, &l2, s))) return rc; a.v = (const T*)q;
    if (p_is_1d) { if ((rc = canon_p1d(p, dtype, layout, nlev, sc, &q, s))) return rc; p = q; a.p = (const T*)q; }
    else { if ((rc = canon3d_same(p, dtype, layout, ncol, nlev, ld_in, sc, &q, &l2, s))) return rc; a.p = (const T*)q; }
    a.lev_stride = ld; a.col_stride = 1;
  }
  a.ps = (const T*)ps; a.ts = (const T*)ts; a.tds = (const T*)tds; a.us = (const T*)us; a.vs = (const T*)vs;
  a.start = start_3d;
  if (p_is_1d && !start_3d) {
    int32_t* st;
    XC_CUDA(sc.alloc(&st, (size_t)ncol));
    if ((rc = launch_pres_lev_pos(p, ps, dtype, ncol, nlev, st, s))) return rc;
    a.start = st;
  }
  a.ncol = ncol; a.ld = ld; a.nlev = nlev; a.depth = depth; a.aglh0 = aglh0;
  a.srh_rm = srh_rm; a.srh_lm = srh_lm; a.rm = rm; a.lm = lm; a.mean6 = mean6;
  { int32_t* wl; int* wc; XC_CUDA(sc.alloc(&wl, (size_t)ncol)); XC_CUDA(sc.alloc(&wc, (size_t)1)); a.work_list = wl; a.work_count = wc; }
  return tile ? launch_srh_tile(a, p_is_1d != 0, s) : launch_srh(a, p_is_1d != 0, s);
}

// srh.srh signature: heights are given (float64 from stdheight in the reference)
template <class T>
int srh_heights_device_t(const void* u, const void* v, const void* aglh, const void* us, const void* vs, const void* aglhs,
                         int64_t ncol, int nlev, int dtype, int layout, int64_t ld_in, double depth, const int32_t* start_3d,
                         double* srh_rm, double* srh_lm, float* rm, float* lm, float* mean6, cudaStream_t s) {
  Scratch sc(s);
  SrhArgs<T> a{};
  int rc;
  const void* q;
  int64_t ld = ncol, l2 = ncol;
  if ((rc = canon3d_same(u, dtype, layout, ncol, nlev, ld_in, sc, &q, &ld, s))) return rc; a.u = (const T*)q;
  if ((rc = canon3d_same(v, dtype, layout, ncol, nlev, ld_in, sc, &q, &l2, s))) return rc; a.v = (const T*)q;
  if ((rc = canon3d_same(aglh, dtype, layout, ncol, nlev, ld_in, sc, &q, &l2, s))) return rc; a.aglh = (const T*)q;
  a.us = (const T*)us; a.vs = (const T*)vs; a.aglhs = (const T*)aglhs;
  a.start = start_3d;
  a.lev_stride = ld; a.col_stride = 1;
  a.ncol = ncol; a.ld = ld; a.nlev = nlev; a.depth = depth; a.aglh0 = 0.0;
  a.srh_rm = srh_rm; a.srh_lm = srh_lm; a.rm = rm; a.lm = lm; a.mean6 = mean6;
  { int32_t* wl; int* wc; XC_CUDA(sc.alloc(&wl, (size_t)ncol)); XC_CUDA(sc.alloc(&wc, (size_t)1)); a.work_list = wl; a.work_count = wc; }
  return launch_srh(a, false, s);
}

template <class T>
int height_device_t(const void* p, const void* t, const void* td, const void* ps, const void* ts, const void* tds,
                    int64_t ncol, int nlev, int p_is_1d, int dtype, int layout, int64_t ld_in, double aglh0,
                    const int32_t* start_3d, double* h, double* hs, cudaStream_t s) {
  Scratch sc(s);
  HeightArgs<T> a{};
  int rc;
  const void* q;
  int64_t ld = ncol, l2 = ncol;
  if ((rc = canon3d_same(t, dtype, layout, ncol, nlev, ld_in, sc, &q, &ld, s))) return rc; a.t = (const T*)q;
  if ((rc = canon3d_same(td, dtype, layout, ncol, nlev, ld_in, sc, &q, &l2, s))) return rc; a.td = (const T*)q;
  if (p_is_1d) { if ((rc = canon_p1d(p, dtype, layout, nlev, sc, &q, s))) return rc; p = q; a.p = (const T*)q; }
  else { if ((rc = canon3d_same(p, dtype, layout, ncol, nlev, ld_in, sc, &q, &l2, s))) return rc; a.p = (const T*)q; }
  a.ps = (const T*)ps; a.ts = (const T*)ts; a.tds = (const T*)tds;
  a.start = start_3d;
  if (p_is_1d && !start_3d) {
    int32_t* st;
    XC_CUDA(sc.alloc(&st, (size_t)ncol));
    if ((rc = launch_pres_lev_pos(p, ps, dtype, ncol, nlev, st, s))) return rc;
    a.start = st;
  }
  a.ncol = ncol; a.ld = ld; a.nlev = nlev; a.aglh0 = aglh0; a.h = h; a.hs = hs;
  if (base_layout(layout) == XCAPE_LEVEL_LAST) { a.h_col_stride = nlev; a.h_lev_stride = 1; }
  else { a.h_col_stride = 1; a.h_lev_stride = ncol; }
  if (top_first(layout)) { a.h = h + (int64_t)(nlev - 1) * a.h_lev_stride; a.h_lev_stride = -a.h_lev_stride; }   // heights go back in the caller's order
  return launch_stdheight(a, p_is_1d != 0, s);
}

// ---------------------------------------------------------------------------------------
// Host-pointer plumbing: a ring of streams, each owning device buffers for one column block.
// The block's inputs go H2D, the caller-supplied `launch` enqueues the device work, the
// outputs come D2H — all stream-ordered, so blocks on different streams overlap copy and
// compute.  Pageable host memory works (the driver stages it); pinned memory overlaps fully.
// ---------------------------------------------------------------------------------------
constexpr int kMaxStreams = 8;
inline int64_t env_i64(const char* name, int64_t dflt, int64_t lo, int64_t hi) {
  if (const char* e = getenv(name)) {
    const long long v = atoll(e);
    if (v >= lo && v <= hi) return v;
  }
  return dflt;
}
// Tunables of the host-pointer path (columns per staged block, streams in the ring).  Blocks
// grow geometrically from XCAPE_B200_FIRST_CHUNK_COLS to XCAPE_B200_CHUNK_COLS: a small first block
// gets the GPU busy after ~0.3 ms of H2D, large later blocks keep per-kernel tails rare.
inline int64_t chunk_cols() { return env_i64("XCAPE_B200_CHUNK_COLS", 1 << 17, 1024, 1 << 26); }   // r2 sweep (profiles/r2_probe_e2e_plans2.txt): 131072 -> 10.0 ms per ERA5 field, 262144 -> 10.7, 524288 -> 11.0
inline int64_t first_chunk_cols() { return env_i64("XCAPE_B200_FIRST_CHUNK_COLS", 1 << 16, 1024, 1 << 26); }
inline int ring_streams() { return (int)env_i64("XCAPE_B200_STREAMS", 4, 1, kMaxStreams); }

struct HostIn3 { const void* host; };                    // [ncol][nlev] or [nlev][ncol], es bytes/element
struct HostIn1 { const void* host; size_t es; };          // [ncol]
struct HostOut { void* host; size_t bytes_per_col; int is3d; };   // per-column outputs; is3d: a 3-D field in the inputs' layout, bytes_per_col = bytes per ELEMENT

struct Block {
  std::vector<void*> in3, in1, out;
  void* p1d = nullptr;
};

// Stored levels [l0, l0 + nl) of columns [c0, c0 + n) of a 3-D host field -> dense [n][nl] / [nl][n] block on the device.
cudaError_t h2d_field(void* dst, const void* src, int layout, int64_t ncol, int nlev, int64_t c0, int64_t n, size_t es,
                      cudaStream_t s, int l0, int nl) {
  if (layout == XCAPE_LEVEL_LAST) {
    const char* from = (const char*)src + ((size_t)c0 * nlev + l0) * es;
    if (nl == nlev) return cudaMemcpyAsync(dst, from, (size_t)n * nlev * es, cudaMemcpyHostToDevice, s);
    return cudaMemcpy2DAsync(dst, (size_t)nl * es, from, (size_t)nlev * es, (size_t)nl * es, (size_t)n, cudaMemcpyHostToDevice, s);
  }
  return cudaMemcpy2DAsync(dst, (size_t)n * es, (const char*)src + ((size_t)l0 * ncol + c0) * es, (size_t)ncol * es, (size_t)n * es,
                           (size_t)nl, cudaMemcpyHostToDevice, s);
}
cudaError_t d2h_field(void* dst, const void* src, int layout, int64_t ncol, int nlev, int64_t c0, int64_t n, size_t es,
                      cudaStream_t s) {
  if (layout == XCAPE_LEVEL_LAST)
    return cudaMemcpyAsync((char*)dst + (size_t)c0 * nlev * es, src, (size_t)n * nlev * es, cudaMemcpyDeviceToHost, s);
  return cudaMemcpy2DAsync((char*)dst + (size_t)c0 * es, (size_t)ncol * es, src, (size_t)n * es, (size_t)n * es, (size_t)nlev,
                           cudaMemcpyDeviceToHost, s);
}

// Pinned host staging buffers, cached process-wide (cudaHostAlloc costs milliseconds).  Best fit with a slack limit
// (a small request must not occupy a buffer several times its size), and a cap on what stays cached: idle buffers
// are freed least-recently-used first once the total exceeds XCAPE_B200_PINNED_CACHE_BYTES (default 8 GiB), so a
// long-running worker that sees many shapes does not grow page-locked memory without bound.
struct PinnedCache {
  struct Ent { void* p; size_t bytes; bool used; uint64_t stamp; };
  std::mutex mu;
  std::vector<Ent> ents;
  uint64_t clock = 0;
  size_t total = 0;
  static size_t cap() { return (size_t)env_i64("XCAPE_B200_PINNED_CACHE_BYTES", (int64_t)8 << 30, 0, (int64_t)1 << 46); }
  void* acquire(size_t bytes) {
    bytes = std::max<size_t>(bytes, 1);
    std::lock_guard<std::mutex> lk(mu);
    Ent* best = nullptr;
    for (auto& e : ents)
      if (!e.used && e.bytes >= bytes && e.bytes <= 2 * bytes + ((size_t)1 << 20) && (!best || e.bytes < best->bytes)) best = &e;
    if (best) { best->used = true; best->stamp = ++clock; return best->p; }
    evict_locked(bytes);
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    ents.push_back({p, bytes, true, ++clock});
    total += bytes;
    return p;
  }
  void release(void* p) {
    std::lock_guard<std::mutex> lk(mu);
    for (auto& e : ents)
      if (e.p == p) { e.used = false; e.stamp = ++clock; }
    evict_locked(0);
  }
  // free idle entries, oldest first, until `incoming` more bytes fit under the cap
  void evict_locked(size_t incoming) {
    const size_t limit = cap();
    while (total + incoming > limit) {
      int victim = -1;
      for (int i = 0; i < (int)ents.size(); ++i)
        if (!ents[i].used && (victim < 0 || ents[i].stamp < ents[victim].stamp)) victim = i;
      if (victim < 0) break;
      cudaFreeHost(ents[victim].p);
      total -= ents[victim].bytes;
      ents.erase(ents.begin() + victim);
    }
  }
  void trim() {                                   // give idle buffers back (xcape_cuda_release_memory)
    std::lock_guard<std::mutex> lk(mu);
    std::vector<Ent> kept;
    for (auto& e : ents) {
      if (e.used) kept.push_back(e);
      else { cudaFreeHost(e.p); total -= e.bytes; }
    }
    ents.swap(kept);
  }
  size_t cached_bytes() { std::lock_guard<std::mutex> lk(mu); return total; }
};
PinnedCache g_pinned;

// Multi-threaded host copies into a pinned staging buffer (pageable caller memory would otherwise go
// through the driver's single bounce buffer at ~12 GB/s and block the enqueuing thread).
inline int copy_threads() {
  static const int nthr = (int)env_i64("XCAPE_B200_COPY_THREADS", std::min<int64_t>(8, std::max<int64_t>(1, std::thread::hardware_concurrency() / 2)), 1, 64);
  return nthr;
}
// fn(part) for part in [0, nparts) on up to copy_threads() host threads (the caller's included)
template <class F>
void parallel_parts(int nparts, F fn) {
  const int nthr = std::min(copy_threads(), nparts);
  if (nthr <= 1) { for (int q = 0; q < nparts; ++q) fn(q); return; }
  std::atomic<int> next{0};
  auto work = [&] { for (int q; (q = next.fetch_add(1)) < nparts;) fn(q); };
  std::vector<std::thread> th;
  for (int t = 1; t < nthr; ++t) th.emplace_back(work);
  work();
  for (auto& x : th) x.join();
}
void parallel_memcpy(void* dst, const void* src, size_t bytes) {
  if (bytes < ((size_t)4 << 20) || copy_threads() == 1) { memcpy(dst, src, bytes); return; }
  const size_t part = (size_t)1 << 20;
  const int nparts = (int)((bytes + part - 1) / part);
  parallel_parts(nparts, [=](int q) {
    const size_t a = (size_t)q * part;
    memcpy((char*)dst + a, (const char*)src + a, std::min(part, bytes - a));
  });
}

// stored levels [l0, l0 + nl) of columns [c0, c0+n) of a 3-D host field -> dense block of the same layout in `dst` (host)
void stage_field(void* dst, const void* src, int layout, int64_t ncol, int nlev, int64_t c0, int64_t n, size_t es, int l0, int nl) {
  if (layout == XCAPE_LEVEL_LAST) {
    const char* from = (const char*)src + ((size_t)c0 * nlev + l0) * es;
    if (nl == nlev) { parallel_memcpy(dst, from, (size_t)n * nlev * es); return; }
    const int64_t per = 8192;                      // columns per work item: compact each column's window
    parallel_parts((int)((n + per - 1) / per), [=](int q) {
      const int64_t a = (int64_t)q * per, b = std::min<int64_t>(n, a + per);
      for (int64_t c = a; c < b; ++c) memcpy((char*)dst + (size_t)c * nl * es, from + (size_t)c * nlev * es, (size_t)nl * es);
    });
    return;
  }
  auto row = [=](int k) { memcpy((char*)dst + (size_t)k * n * es, (const char*)src + ((size_t)(l0 + k) * ncol + c0) * es, (size_t)n * es); };
  if ((size_t)n * nl * es < ((size_t)4 << 20)) { for (int k = 0; k < nl; ++k) row(k); return; }
  parallel_parts(nl, row);                         // level-major: one row of the block per level
}

bool is_pageable_host(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return a.type == cudaMemoryTypeUnregistered;
}

// Streams and events of the host-path ring are created once and recycled through a process-wide pool
// (cudaStreamCreate costs ~0.1 ms, a visible share of a 3-17 ms call).  A pool rather than thread-local
// storage: callers such as dask workers or this package's own multi-GPU sharding use short-lived threads,
// which would otherwise create (and leak) a ring per call.
struct Ring {
  int dev = -1, n = 0;
  cudaStream_t st[kMaxStreams] = {};
  cudaEvent_t ev[kMaxStreams] = {};
  int ensure(int want) {
    for (; n < want; ++n) {
      XC_CUDA(cudaStreamCreateWithFlags(&st[n], cudaStreamNonBlocking));
      XC_CUDA(cudaEventCreateWithFlags(&ev[n], cudaEventDisableTiming));
    }
    return XCAPE_OK;
  }
};
struct RingPool {
  std::mutex mu;
  std::vector<Ring*> idle;
  Ring* acquire(int device) {
    std::lock_guard<std::mutex> lk(mu);
    for (size_t i = 0; i < idle.size(); ++i)
      if (idle[i]->dev == device) { Ring* r = idle[i]; idle.erase(idle.begin() + i); return r; }
    Ring* r = new Ring();
    r->dev = device;
    return r;
  }
  void release(Ring* r) { std::lock_guard<std::mutex> lk(mu); idle.push_back(r); }
};
RingPool g_rings;

// A D2H copy into PAGEABLE memory blocks the host until the producing kernel has finished, which
// would serialise the ring (r1c probe: one block per field was faster than four).  Outputs — per-column
// ones and 3-D fields (stdheight's h, dewpoint_from_q's td) alike — are therefore landed in a pinned
// staging slot per stream and copied to the caller's arrays by the host once the slot's event fires,
// while later blocks are still running.
// `l0`, `nl`: window of STORED levels of the 3-D inputs (and of the shared pressure axis) that is shipped; the launch
// callback then sees nl-level blocks.  Calls with 3-D outputs ship every level.
template <class F>
int run_staged(int64_t ncol, int nlev, int layout, size_t es, const void* p1d_host, const std::vector<HostIn3>& in3,
               const std::vector<HostIn1>& in1, const std::vector<HostOut>& outs, F launch, int l0 = 0, int nl = -1,
               int64_t pitch = -1, int64_t col0 = 0) {
  // `pitch`, `col0`: the call works on columns [col0, col0 + ncol) of host arrays that hold `pitch` columns (a shard of a
  // larger field, xcape_cuda_*_multi): every host address below is formed with col0 + c and, for level-major rows, pitch.
  if (nl < 0) nl = nlev;
  if (pitch < 0) pitch = ncol;
  const int64_t chunk = std::min<int64_t>(ncol, chunk_cols());          // capacity of a slot
  const int64_t first = std::min<int64_t>(chunk, first_chunk_cols());
  // block plan: geometric ramp, then full blocks; a short remainder is shared with its predecessor so that
  // the call does not end on a kernel too small to fill the GPU
  std::vector<int64_t> plan;
  for (int64_t c = 0, n = first; c < ncol; n = std::min<int64_t>(chunk, n * 2)) {
    const int64_t m = std::min<int64_t>(n, ncol - c);
    plan.push_back(m);
    c += m;
  }
  if (plan.size() >= 2 && plan.back() * 2 < plan[plan.size() - 2]) {
    const int64_t tot = plan.back() + plan[plan.size() - 2];
    const int64_t a = ((tot / 2 + 127) / 128) * 128;
    plan[plan.size() - 2] = a;
    plan.back() = tot - a;
  }
  const int nblocks = (int)plan.size();
  const int nstream = std::min<int>(ring_streams(), nblocks);
  cudaStream_t st[kMaxStreams] = {};
  cudaEvent_t done[kMaxStreams] = {};
  Ring* ring = nullptr;
  Block b[kMaxStreams];
  struct Pending { bool on = false; int64_t c0 = 0, n = 0; } pend[kMaxStreams];
  std::vector<void*> stage[kMaxStreams];          // pinned staging per output (nullptr = direct copy)
  std::vector<void*> stage3[kMaxStreams], stage1[kMaxStreams];   // pinned staging per pageable input
  std::vector<char> in3_pageable(in3.size(), 0), in1_pageable(in1.size(), 0);
  const bool stage_inputs = env_i64("XCAPE_B200_STAGE_PAGEABLE", 1, 0, 1) != 0;
  for (size_t k = 0; k < in3.size(); ++k) in3_pageable[k] = stage_inputs && is_pageable_host(in3[k].host);
  for (size_t k = 0; k < in1.size(); ++k) in1_pageable[k] = stage_inputs && is_pageable_host(in1[k].host);
  std::vector<char> staged(outs.size(), 0);
  for (size_t k = 0; k < outs.size(); ++k)
    staged[k] = (outs[k].host && is_pageable_host(outs[k].host)) ? 1 : 0;        // per-column AND 3-D outputs

  const bool trace = getenv("XCAPE_B200_TRACE") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_begin = now();
  auto drain = [&](int i) -> int {                // wait for slot i's block, hand its outputs to the caller
    if (!pend[i].on) return XCAPE_OK;
    XC_CUDA(cudaEventSynchronize(done[i]));
    for (size_t k = 0; k < outs.size(); ++k) {
      if (!staged[k]) continue;
      const size_t bpc = outs[k].bytes_per_col;
      const int64_t c0 = pend[i].c0 + col0, n = pend[i].n;
      if (!outs[k].is3d) { memcpy((char*)outs[k].host + (size_t)c0 * bpc, stage[i][k], (size_t)n * bpc); continue; }
      // a 3-D field in the inputs' layout: the slot holds the dense [n][nlev] / [nlev][n] block
      if (layout == XCAPE_LEVEL_LAST) parallel_memcpy((char*)outs[k].host + (size_t)c0 * nlev * bpc, stage[i][k], (size_t)n * nlev * bpc);
      else parallel_parts(nlev, [&](int lev) {
        memcpy((char*)outs[k].host + ((size_t)lev * pitch + c0) * bpc, (const char*)stage[i][k] + (size_t)lev * n * bpc, (size_t)n * bpc);
      });
    }
    pend[i].on = false;
    return XCAPE_OK;
  };

  // XCAPE_B200_TRACE: device-side timeline of every block (timing events; diagnostics only)
  struct BlockEv { cudaEvent_t e[4] = {}; int64_t n = 0; int slot = 0; };
  std::vector<BlockEv> tl;
  cudaEvent_t tl_base = nullptr;
  auto mark = [&](cudaEvent_t* e, cudaStream_t s) {
    if (cudaEventCreate(e) == cudaSuccess) cudaEventRecord(*e, s);
  };

  auto body = [&]() -> int {
    int cur_dev = 0;
    XC_CUDA(cudaGetDevice(&cur_dev));
    ring = g_rings.acquire(cur_dev);
    { int r = ring->ensure(nstream); if (r) return r; }
    for (int i = 0; i < nstream; ++i) {
      st[i] = ring->st[i];
      done[i] = ring->ev[i];
      for (size_t k = 0; k < in3.size(); ++k) {
        void* q; XC_CUDA(pool_alloc(&q, (size_t)chunk * nl * es, st[i])); b[i].in3.push_back(q);
        void* h = nullptr;
        if (in3_pageable[k] && !(h = g_pinned.acquire((size_t)chunk * nl * es))) return fail(XCAPE_ERR_CUDA, "cudaHostAlloc failed (input staging)");
        stage3[i].push_back(h);
      }
      for (size_t k = 0; k < in1.size(); ++k) {
        void* q; XC_CUDA(pool_alloc(&q, (size_t)chunk * in1[k].es, st[i])); b[i].in1.push_back(q);
        void* h = nullptr;
        if (in1_pageable[k] && !(h = g_pinned.acquire((size_t)chunk * in1[k].es))) return fail(XCAPE_ERR_CUDA, "cudaHostAlloc failed (input staging)");
        stage1[i].push_back(h);
      }
      for (size_t k = 0; k < outs.size(); ++k) {
        void* q; XC_CUDA(pool_alloc(&q, (size_t)chunk * (outs[k].is3d ? (size_t)nlev : 1) * outs[k].bytes_per_col, st[i]));
        b[i].out.push_back(q);
        void* h = nullptr;
        if (staged[k]) {
          h = g_pinned.acquire((size_t)chunk * (outs[k].is3d ? (size_t)nlev : 1) * outs[k].bytes_per_col);
          if (!h) return fail(XCAPE_ERR_CUDA, "cudaHostAlloc failed for the output staging buffer");
        }
        stage[i].push_back(h);
      }
      if (p1d_host) {
        XC_CUDA(pool_alloc(&b[i].p1d, (size_t)nl * es, st[i]));
        XC_CUDA(cudaMemcpyAsync(b[i].p1d, (const char*)p1d_host + (size_t)l0 * es, (size_t)nl * es, cudaMemcpyHostToDevice, st[i]));
      }
    }
    if (trace) fprintf(stderr, "[xcape_b200]   setup done at %.3f ms\n", now() - t_begin);
    int i = 0;
    int64_t c0 = 0;
    for (size_t blk = 0; blk < plan.size(); ++blk, i = (i + 1) % nstream) {
      const int64_t n = plan[blk];
      cudaStream_t s = st[i];
      int r = drain(i);                            // slot reuse: its previous block must have left
      if (r) return r;
      if (trace) {
        if (!tl_base) mark(&tl_base, s);
        tl.emplace_back(); tl.back().n = n; tl.back().slot = i;
        mark(&tl.back().e[0], s);
      }
      for (size_t k = 0; k < in3.size(); ++k) {
        if (stage3[i][k]) {        // pageable: host threads fill the slot's pinned buffer, then a true async H2D
          stage_field(stage3[i][k], in3[k].host, layout, pitch, nlev, c0 + col0, n, es, l0, nl);
          XC_CUDA(cudaMemcpyAsync(b[i].in3[k], stage3[i][k], (size_t)n * nl * es, cudaMemcpyHostToDevice, s));
        } else {
          XC_CUDA(h2d_field(b[i].in3[k], in3[k].host, layout, pitch, nlev, c0 + col0, n, es, s, l0, nl));
        }
      }
      for (size_t k = 0; k < in1.size(); ++k) {
        const void* src = (const char*)in1[k].host + (size_t)(c0 + col0) * in1[k].es;
        if (stage1[i][k]) { memcpy(stage1[i][k], src, (size_t)n * in1[k].es); src = stage1[i][k]; }
        XC_CUDA(cudaMemcpyAsync(b[i].in1[k], src, (size_t)n * in1[k].es, cudaMemcpyHostToDevice, s));
      }
      if (trace) mark(&tl.back().e[1], s);
      if ((r = launch(b[i], n, s))) return r;
      if (trace) mark(&tl.back().e[2], s);
      for (size_t k = 0; k < outs.size(); ++k) {
        if (!outs[k].host) continue;
        if (outs[k].is3d && staged[k]) XC_CUDA(cudaMemcpyAsync(stage[i][k], b[i].out[k], (size_t)n * nlev * outs[k].bytes_per_col, cudaMemcpyDeviceToHost, s));
        else if (outs[k].is3d) XC_CUDA(d2h_field(outs[k].host, b[i].out[k], layout, pitch, nlev, c0 + col0, n, outs[k].bytes_per_col, s));
        else if (staged[k]) XC_CUDA(cudaMemcpyAsync(stage[i][k], b[i].out[k], (size_t)n * outs[k].bytes_per_col, cudaMemcpyDeviceToHost, s));
        else XC_CUDA(cudaMemcpyAsync((char*)outs[k].host + (size_t)(c0 + col0) * outs[k].bytes_per_col, b[i].out[k], (size_t)n * outs[k].bytes_per_col, cudaMemcpyDeviceToHost, s));
      }
      if (trace) mark(&tl.back().e[3], s);
      XC_CUDA(cudaEventRecord(done[i], s));
      pend[i].on = true; pend[i].c0 = c0; pend[i].n = n;
      c0 += n;
    }
    if (trace) fprintf(stderr, "[xcape_b200]   all blocks enqueued at %.3f ms\n", now() - t_begin);
    for (int k = 0; k < nstream; ++k) {            // oldest block first
      int r = drain((i + k) % nstream);
      if (r) return r;
      if (trace) fprintf(stderr, "[xcape_b200]   slot %d drained at %.3f ms\n", (i + k) % nstream, now() - t_begin);
    }
    for (int k = 0; k < nstream; ++k) XC_CUDA(cudaStreamSynchronize(st[k]));
    return XCAPE_OK;
  };
  int rc = body();
  const double t_body = now();
  std::string keep = g_last_error;
  for (int i = 0; i < nstream; ++i) {
    if (st[i]) {
      for (void* q : b[i].in3) cudaFreeAsync(q, st[i]);
      for (void* q : b[i].in1) cudaFreeAsync(q, st[i]);
      for (void* q : b[i].out) cudaFreeAsync(q, st[i]);
      if (b[i].p1d) cudaFreeAsync(b[i].p1d, st[i]);
      if (rc) cudaStreamSynchronize(st[i]);       // success: body() already synchronised every stream
    }
    for (void* h : stage[i]) if (h) g_pinned.release(h);
    for (void* h : stage3[i]) if (h) g_pinned.release(h);
    for (void* h : stage1[i]) if (h) g_pinned.release(h);
  }
  if (ring) g_rings.release(ring);
  if (trace && tl_base) {
    for (size_t q = 0; q < tl.size(); ++q) {
      float t[4] = {-1, -1, -1, -1};
      for (int j = 0; j < 4; ++j)
        if (tl[q].e[j]) { if (!rc) cudaEventElapsedTime(&t[j], tl_base, tl[q].e[j]); cudaEventDestroy(tl[q].e[j]); }
      fprintf(stderr, "[xcape_b200]   block %2zu slot %d n=%7lld: h2d %.3f..%.3f ms, kernels ..%.3f ms, d2h ..%.3f ms (device clock)\n",
              q, tl[q].slot, (long long)tl[q].n, t[0], t[1], t[2], t[3]);
    }
    cudaEventDestroy(tl_base);
  }
  if (rc) { cudaGetLastError(); g_last_error = keep; }
  if (trace) fprintf(stderr, "[xcape_b200] run_staged ncol=%lld blocks=%d streams=%d: body %.3f ms, teardown %.3f ms\n",
                     (long long)ncol, nblocks, nstream, t_body - t_begin, now() - t_body);
  return rc;
}

}  // namespace
}  // namespace xc

using namespace xc;

// block r of `ndevices` contiguous blocks of whole 128-column units (xcape_b200/sharding.py::column_blocks)
static void shard_block(int64_t ncol, int ndevices, int r, int64_t* c0, int64_t* c1) {
  const int64_t units = (ncol + 127) / 128, base = units / ndevices, extra = units % ndevices;
  const int64_t u0 = (int64_t)r * base + std::min<int64_t>(r, extra), u1 = u0 + base + (r < extra ? 1 : 0);
  *c0 = std::min(u0 * 128, ncol); *c1 = std::min(u1 * 128, ncol);
}
// fn(r, c0, c1) on one host thread per non-empty block, each bound to its device; first failure wins
template <class F>
static int run_sharded(int64_t ncol, const int* devices, int ndevices, F fn) {
  std::vector<int> rcs((size_t)ndevices, XCAPE_OK);
  std::vector<std::string> msgs((size_t)ndevices);
  std::vector<std::thread> th;
  for (int r = 0; r < ndevices; ++r) {
    int64_t c0, c1;
    shard_block(ncol, ndevices, r, &c0, &c1);
    if (c1 <= c0) continue;
    auto work = [&, r, c0, c1] {
      DeviceGuard dg(devices[r]);
      if (!dg.ok) { rcs[(size_t)r] = XCAPE_ERR_NODEV; msgs[(size_t)r] = "cudaSetDevice failed"; return; }
      rcs[(size_t)r] = fn(r, c0, c1);
      if (rcs[(size_t)r]) msgs[(size_t)r] = g_last_error;      // thread-local: carry it to the caller's thread
    };
    try {
      th.emplace_back(work);
    } catch (...) {                                          // no thread to be had: run the block here (no exception may cross the C ABI)
      work();
    }
  }
  for (auto& x : th) x.join();
  for (int r = 0; r < ndevices; ++r)
    if (rcs[(size_t)r]) return fail(rcs[(size_t)r], "device " + std::to_string(devices[r]) + ": " + msgs[(size_t)r]);
  return XCAPE_OK;
}

extern "C" {

const char* xcape_cuda_last_error(void) { return g_last_error.c_str(); }
const char* xcape_cuda_version(void) { return "xcape_b200 0.1.0 sm_100a"; }
int64_t xcape_cuda_kernel_launches(void) { return g_launches.load(); }
int xcape_cuda_time_kernels(int enable) { xc::t_time_kernels = enable != 0; return XCAPE_OK; }
int xcape_cuda_last_kernel_ms(double* ms) {
  using namespace xc;
  if (!ms) return fail(XCAPE_ERR_ARG, "null pointer");
  if (!t_timer_valid) return fail(XCAPE_ERR_ARG, "no timed kernel on this thread (call xcape_cuda_time_kernels(1) first)");
  DeviceGuard dg(t_timer_dev);
  XC_CUDA(cudaEventSynchronize(t_timer_ev[1]));
  float f = 0.0f;
  XC_CUDA(cudaEventElapsedTime(&f, t_timer_ev[0], t_timer_ev[1]));
  *ms = (double)f;
  return XCAPE_OK;
}
int64_t xcape_cuda_columns_redone(void) { return g_redone.load(); }
int xcape_cuda_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return -1; }
  return n;
}

int xcape_cuda_release_memory(int device) {
  DeviceGuard dg(device);
  if (!dg.ok) return fail(XCAPE_ERR_NODEV, "cudaSetDevice failed");
  XC_CUDA(cudaDeviceSynchronize());
  if (cudaMemPool_t pool = pool_for_current_device()) XC_CUDA(cudaMemPoolTrimTo(pool, 0));
  g_pinned.trim();
  return XCAPE_OK;
}

int xcape_cuda_measure_peaks(int device, int reps, double* fp32_tflops, double* fp64_tflops) {
  DeviceGuard dg(device);
  if (!dg.ok) return fail(XCAPE_ERR_NODEV, "cudaSetDevice failed");
  return measure_peaks(reps, fp32_tflops, fp64_tflops);
}

int xcape_cuda_measure_fp32_rrr(int device, int reps, double* fp32_tflops) {
  if (!fp32_tflops) return fail(XCAPE_ERR_ARG, "null pointer");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(XCAPE_ERR_NODEV, "cudaSetDevice failed");
  return measure_fp32_rrr(reps, fp32_tflops);
}

int xcape_cuda_pres_lev_pos(const void* p, const void* ps, int64_t ncol, int nlev, int dtype, int mem,
                            int32_t* start_3d, int device, void* stream) {
  int rc = check_common(ncol, nlev, dtype, XCAPE_LEVEL_LAST, mem);
  if (rc) return rc;
  if (ncol == 0) return XCAPE_OK;
  if (!p || !ps || !start_3d) return fail(XCAPE_ERR_ARG, "null pointer");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(XCAPE_ERR_NODEV, "cudaSetDevice failed");
  if (mem == XCAPE_MEM_DEVICE) return launch_pres_lev_pos(p, ps, dtype, ncol, nlev, start_3d, (cudaStream_t)stream);
  return run_staged(ncol, nlev, XCAPE_LEVEL_LAST, esize(dtype), p, {}, {{ps, esize(dtype)}}, {{start_3d, 4, 0}},
                    [&](Block& b, int64_t n, cudaStream_t s) {
                      return launch_pres_lev_pos(b.p1d, b.in1[0], dtype, n, nlev, (int32_t*)b.out[0], s);
                    });
}

// Host-pointer CAPE on columns [col0, col0 + ncol) of arrays that hold `pitch` columns (the whole call: col0 = 0, pitch = ncol;
// a device's shard in xcape_cuda_cape_multi).  The current device is the one the work runs on.
static int cape_host(const void* p, const void* t, const void* td, const void* ps, const void* ts, const void* tds,
                     int64_t pitch, int64_t col0, int64_t ncol, int nlev, int p_is_1d, int dtype, int layout,
                     int source, int adiabat, float ml_depth, float pinc, const int32_t* start_3d,
                     float* cape, float* cin, int32_t* mulev, float* zmulev, int32_t* status, int32_t* n_iter, int precision) {
  const size_t es = esize(dtype);
  const int lay = base_layout(layout);
  const bool rev = top_first(layout);
  auto host_call = [&](const void* p_, const void* t_, const void* td_, const void* ps_, const void* ts_, const void* tds_,
                       const int32_t* st3_, int64_t n, int64_t pitch_, int64_t col0_, int lay_flags, int l0, int nl, float* cape_,
                       float* cin_, int32_t* mulev_, float* zmulev_, int32_t* status_, int32_t* n_iter_) -> int {
    std::vector<HostIn3> in3 = {{t_}, {td_}};
    if (!p_is_1d) in3.push_back({p_});
    std::vector<HostIn1> in1 = {{ps_, es}, {ts_, es}, {tds_, es}};
    if (st3_) in1.push_back({st3_, 4});
    std::vector<HostOut> outs = {{cape_, 4, 0}, {cin_, 4, 0}, {mulev_, 4, 0}, {zmulev_, 4, 0}, {status_, 4, 0}, {n_iter_, 4, 0}};
    const bool more = nl < nlev;
    return run_staged(n, nlev, base_layout(lay_flags), es, p_is_1d ? p_ : nullptr, in3, in1, outs,
                      [&](Block& b, int64_t m, cudaStream_t s) {
                        return cape_device(p_is_1d ? b.p1d : b.in3[2], b.in3[0], b.in3[1], b.in1[0], b.in1[1], b.in1[2], m, nl,
                                           p_is_1d, dtype, lay_flags, m, source, adiabat, ml_depth, pinc,
                                           st3_ ? (const int32_t*)b.in1[3] : nullptr, (float*)b.out[0], (float*)b.out[1],
                                           (int32_t*)b.out[2], (float*)b.out[3], status_ ? (int32_t*)b.out[4] : nullptr,
                                           n_iter_ ? (int32_t*)b.out[5] : nullptr, precision, s, more);
                      }, l0, nl, pitch_, col0_);
  };

  // Ship only the levels the ascent can reach.  The parcel stops at the first level with p <= 100 hPa where it is
  // negatively buoyant (f90:554-557), which for almost every column is the 100 hPa level itself — the 10 ERA5 levels
  // above it (27 % of the bytes) are never read.  On a shared pressure axis that level is known before anything is
  // copied: levels up to and including the first one with p <= 100 hPa go to the device, a column that is still
  // ascending there comes back with status 4 and is redone below with every level.
  // Only level-major fields: they lose whole rows, which is free.  Level-last (reference-layout) fields would need a
  // 2-D copy of ~100-byte rows, which the copy engines do at a tenth of their bandwidth (measured: 135 ms instead of 13
  // per ERA5 field), and compacting them column by column in the pageable staging copy costs more host time than the
  // bytes saved (17.6 vs 15.8 ms).
  const bool can_window = lay == XCAPE_LEVEL_MAJOR;
  int nl = nlev, l0 = 0;
  if (p_is_1d && can_window && env_i64("XCAPE_B200_SHIP_ALL_LEVELS", 0, 0, 1) == 0) {
    int need = nlev;
    for (int k = 0; k < nlev; ++k) {               // k counts from the surface
      const int ks = rev ? nlev - 1 - k : k;
      const double pk = (dtype == XCAPE_F64) ? ((const double*)p)[ks] : (double)((const float*)p)[ks];
      if (pk <= 100.0) { need = k + 1; break; }
    }
    if (need + 2 <= nlev) { nl = need; l0 = rev ? nlev - need : 0; }
  }
  if (nl == nlev)
    return host_call(p, t, td, ps, ts, tds, start_3d, ncol, pitch, col0, layout, 0, nlev, cape, cin, mulev, zmulev, status, n_iter);

  static thread_local std::vector<int32_t> own_status;      // reused: a fresh 4 MB vector per call costs 0.3 ms of page faults
  // status words of this shard, indexed like the caller's arrays (col0 + c)
  int32_t* stat = status;
  if (!stat) {
    if (own_status.size() < (size_t)ncol) own_status.resize((size_t)ncol);
    // run_staged addresses every host array at (col0 + c): bias the shard-local buffer so that element col0 is its first
    stat = (int32_t*)((uintptr_t)own_status.data() - (uintptr_t)col0 * sizeof(int32_t));
  }
  int rc = host_call(p, t, td, ps, ts, tds, start_3d, ncol, pitch, col0, layout, l0, nl, cape, cin, mulev, zmulev, stat, n_iter);
  if (rc) return rc;
  std::vector<int64_t> redo;
  {
    int32_t seen = 0;                              // valid status words are 0..4: OR-ing them finds "any 4" at memory speed
    for (int64_t c = col0; c < col0 + ncol; ++c) seen |= stat[c];
    if (seen & 4)
      for (int64_t c = col0; c < col0 + ncol; ++c)
        if (stat[c] == 4) redo.push_back(c);
  }
  if (redo.empty()) return XCAPE_OK;
  g_redone.fetch_add((int64_t)redo.size(), std::memory_order_relaxed);
  // gather the unfinished columns (all levels, stored order) into a compact level-last batch and run it again
  const int64_t m = (int64_t)redo.size();
  std::vector<char> gt((size_t)m * nlev * es), gtd((size_t)m * nlev * es), gs((size_t)m * es * 3);
  std::vector<int32_t> gst3(start_3d ? (size_t)m : 0), o_mu((size_t)m), o_st((size_t)m), o_it((size_t)m);
  std::vector<float> o_cape((size_t)m), o_cin((size_t)m), o_z((size_t)m);
  auto gather3 = [&](const void* src, char* dst) {
    for (int64_t j = 0; j < m; ++j) {
      const int64_t c = redo[(size_t)j];
      if (lay == XCAPE_LEVEL_LAST) memcpy(dst + (size_t)j * nlev * es, (const char*)src + (size_t)c * nlev * es, (size_t)nlev * es);
      else for (int k = 0; k < nlev; ++k) memcpy(dst + ((size_t)j * nlev + k) * es, (const char*)src + ((size_t)k * pitch + c) * es, es);
    }
  };
  gather3(t, gt.data()); gather3(td, gtd.data());
  const void* srf[3] = {ps, ts, tds};
  for (int f = 0; f < 3; ++f)
    for (int64_t j = 0; j < m; ++j) memcpy(gs.data() + ((size_t)f * m + j) * es, (const char*)srf[f] + (size_t)redo[(size_t)j] * es, es);
  for (int64_t j = 0; j < m && start_3d; ++j) gst3[(size_t)j] = start_3d[redo[(size_t)j]];
  rc = host_call(p, gt.data(), gtd.data(), gs.data(), gs.data() + (size_t)m * es, gs.data() + (size_t)2 * m * es,
                 start_3d ? gst3.data() : nullptr, m, m, 0, XCAPE_LEVEL_LAST | (rev ? XCAPE_LEVELS_TOP_FIRST : 0), 0, nlev,
                 o_cape.data(), o_cin.data(), o_mu.data(), o_z.data(), o_st.data(), o_it.data());
  if (rc) return rc;
  for (int64_t j = 0; j < m; ++j) {
    const int64_t c = redo[(size_t)j];
    cape[c] = o_cape[(size_t)j]; cin[c] = o_cin[(size_t)j]; mulev[c] = o_mu[(size_t)j]; zmulev[c] = o_z[(size_t)j];
    if (status) status[c] = o_st[(size_t)j];
    if (n_iter) n_iter[c] = o_it[(size_t)j];
  }
  return XCAPE_OK;
}

int xcape_cuda_cape(const void* p, const void* t, const void* td, const void* ps, const void* ts, const void* tds,
                    int64_t ncol, int nlev, int p_is_1d, int dtype, int layout, int mem,
                    int source, int adiabat, float ml_depth, float pinc, const int32_t* start_3d,
                    float* cape, float* cin, int32_t* mulev, float* zmulev, int32_t* status, int32_t* n_iter,
                    int precision, int device, void* stream) {
  int rc = check_common(ncol, nlev, dtype, layout, mem);
  if (rc) return rc;
  if (source < 1 || source > 3) return fail(XCAPE_ERR_ARG, "source must be 1 (surface), 2 (most-unstable) or 3 (mixed-layer)");
  if (adiabat < 1 || adiabat > 4) return fail(XCAPE_ERR_ARG, "adiabat must be 1..4");
  if (!(pinc > 0.0f)) return fail(XCAPE_ERR_ARG, "pinc must be > 0");
  if (precision < XCAPE_FAITHFUL || precision > XCAPE_FAST_OPTIMISTIC) return fail(XCAPE_ERR_ARG, "unknown precision mode");
  if (ncol == 0) return XCAPE_OK;
  if (!p || !t || !td || !ps || !ts || !tds || !cape || !cin || !mulev || !zmulev) return fail(XCAPE_ERR_ARG, "null pointer");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(XCAPE_ERR_NODEV, "cudaSetDevice failed");

  if (mem == XCAPE_MEM_DEVICE)
    return cape_device(p, t, td, ps, ts, tds, ncol, nlev, p_is_1d, dtype, layout, ncol, source, adiabat, ml_depth, pinc,
                       start_3d, cape, cin, mulev, zmulev, status, n_iter, precision, (cudaStream_t)stream);

  return cape_host(p, t, td, ps, ts, tds, ncol, 0, ncol, nlev, p_is_1d, dtype, layout, source, adiabat, ml_depth, pinc, start_3d,
                   cape, cin, mulev, zmulev, status, n_iter, precision);
}

// One host call, several GPUs: contiguous 128-aligned column blocks, one host thread per device, no collective
// (SURVEY 8e).  Host memory only; both layouts (a level-major shard is addressed through the field's pitch).
int xcape_cuda_cape_multi(const void* p, const void* t, const void* td, const void* ps, const void* ts, const void* tds,
                          int64_t ncol, int nlev, int p_is_1d, int dtype, int layout,
                          int source, int adiabat, float ml_depth, float pinc, const int32_t* start_3d,
                          float* cape, float* cin, int32_t* mulev, float* zmulev, int32_t* status, int32_t* n_iter,
                          int precision, const int* devices, int ndevices) {
  int rc = check_common(ncol, nlev, dtype, layout, XCAPE_MEM_HOST);
  if (rc) return rc;
  if (!devices || ndevices < 1) return fail(XCAPE_ERR_ARG, "devices: need at least one device index");
  if (source < 1 || source > 3) return fail(XCAPE_ERR_ARG, "source must be 1 (surface), 2 (most-unstable) or 3 (mixed-layer)");
  if (adiabat < 1 || adiabat > 4) return fail(XCAPE_ERR_ARG, "adiabat must be 1..4");
  if (!(pinc > 0.0f)) return fail(XCAPE_ERR_ARG, "pinc must be > 0");
  if (precision < XCAPE_FAITHFUL || precision > XCAPE_FAST_OPTIMISTIC) return fail(XCAPE_ERR_ARG, "unknown precision mode");
  if (ncol == 0) return XCAPE_OK;
  if (!p || !t || !td || !ps || !ts || !tds || !cape || !cin || !mulev || !zmulev) return fail(XCAPE_ERR_ARG, "null pointer");
  return run_sharded(ncol, devices, ndevices, [&](int, int64_t c0, int64_t c1) {
    return cape_host(p, t, td, ps, ts, tds, ncol, c0, c1 - c0, nlev, p_is_1d, dtype, layout, source, adiabat, ml_depth, pinc,
                     start_3d, cape, cin, mulev, zmulev, status, n_iter, precision);
  });
}

// SRH on device pointers (mem == DEVICE: col0 = 0, pitch = ncol) or on columns [col0, col0 + ncol) of host arrays that hold
// `pitch` columns.  The current device is the one the work runs on.
static int srh_any(const void* p, const void* t, const void* td, const void* u, const void* v,
                   const void* ps, const void* ts, const void* tds, const void* us, const void* vs,
                   int64_t pitch, int64_t col0, int64_t ncol, int nlev, int p_is_1d, int dtype, int layout, int mem,
                   double depth, double aglh0, const int32_t* start_3d,
                   double* srh_rm, double* srh_lm, float* rm, float* lm, float* mean6, int precision, cudaStream_t stream) {
  auto dev = [&](const void* p_, const void* t_, const void* td_, const void* u_, const void* v_, const void* ps_,
                 const void* ts_, const void* tds_, const void* us_, const void* vs_, int64_t n, int64_t ld_in,
                 const int32_t* st_, double* srm_, double* slm_, float* rm_, float* lm_, float* m6_, cudaStream_t s) {
    if (dtype == XCAPE_F64)
      return srh_device_t<double>(p_, t_, td_, u_, v_, ps_, ts_, tds_, us_, vs_, n, nlev, p_is_1d, dtype, layout, ld_in, depth,
                                  aglh0, st_, srm_, slm_, rm_, lm_, m6_, precision, s);
    return srh_device_t<float>(p_, t_, td_, u_, v_, ps_, ts_, tds_, us_, vs_, n, nlev, p_is_1d, dtype, layout, ld_in, depth,
                               aglh0, st_, srm_, slm_, rm_, lm_, m6_, precision, s);
  };
  if (mem == XCAPE_MEM_DEVICE)
    return dev(p, t, td, u, v, ps, ts, tds, us, vs, ncol, ncol, start_3d, srh_rm, srh_lm, rm, lm, mean6, stream);

  const size_t es = esize(dtype);
  std::vector<HostIn3> in3 = {{t}, {td}, {u}, {v}};
  if (!p_is_1d) in3.push_back({p});
  std::vector<HostIn1> in1 = {{ps, es}, {ts, es}, {tds, es}, {us, es}, {vs, es}};
  if (start_3d) in1.push_back({start_3d, 4});
  std::vector<HostOut> outs = {{srh_rm, 8, 0}, {srh_lm, 8, 0}, {rm, 8, 0}, {lm, 8, 0}, {mean6, 8, 0}};
  return run_staged(ncol, nlev, base_layout(layout), es, p_is_1d ? p : nullptr, in3, in1, outs,
                    [&](Block& b, int64_t n, cudaStream_t s) {
                      return dev(p_is_1d ? b.p1d : b.in3[4], b.in3[0], b.in3[1], b.in3[2], b.in3[3], b.in1[0], b.in1[1],
                                 b.in1[2], b.in1[3], b.in1[4], n, n, start_3d ? (const int32_t*)b.in1[5] : nullptr,
                                 (double*)b.out[0], (double*)b.out[1], rm ? (float*)b.out[2] : nullptr,
                                 lm ? (float*)b.out[3] : nullptr, mean6 ? (float*)b.out[4] : nullptr, s);
                    }, 0, -1, pitch, col0);
}

static int srh_check(int64_t ncol, int nlev, int dtype, int layout, int mem, int precision, const void* p, const void* t,
                     const void* td, const void* u, const void* v, const void* ps, const void* ts, const void* tds,
                     const void* us, const void* vs, const double* srh_rm, const double* srh_lm) {
  int rc = check_common(ncol, nlev, dtype, layout, mem);
  if (rc) return rc;
  if (precision != XCAPE_FAITHFUL && precision != XCAPE_FAST) return fail(XCAPE_ERR_ARG, "srh: precision must be XCAPE_FAITHFUL or XCAPE_FAST");
  if (ncol > 0 && (!p || !t || !td || !u || !v || !ps || !ts || !tds || !us || !vs || !srh_rm || !srh_lm)) return fail(XCAPE_ERR_ARG, "null pointer");
  return XCAPE_OK;
}

int xcape_cuda_srh(const void* p, const void* t, const void* td, const void* u, const void* v,
                   const void* ps, const void* ts, const void* tds, const void* us, const void* vs,
                   int64_t ncol, int nlev, int p_is_1d, int dtype, int layout, int mem,
                   double depth, double aglh0, const int32_t* start_3d,
                   double* srh_rm, double* srh_lm, float* rm, float* lm, float* mean6,
                   int precision, int device, void* stream) {
  int rc = srh_check(ncol, nlev, dtype, layout, mem, precision, p, t, td, u, v, ps, ts, tds, us, vs, srh_rm, srh_lm);
  if (rc) return rc;
  if (ncol == 0) return XCAPE_OK;
  DeviceGuard dg(device);
  if (!dg.ok) return fail(XCAPE_ERR_NODEV, "cudaSetDevice failed");
  return srh_any(p, t, td, u, v, ps, ts, tds, us, vs, ncol, 0, ncol, nlev, p_is_1d, dtype, layout, mem, depth, aglh0, start_3d,
                 srh_rm, srh_lm, rm, lm, mean6, precision, (cudaStream_t)stream);
}

int xcape_cuda_srh_multi(const void* p, const void* t, const void* td, const void* u, const void* v,
                         const void* ps, const void* ts, const void* tds, const void* us, const void* vs,
                         int64_t ncol, int nlev, int p_is_1d, int dtype, int layout,
                         double depth, double aglh0, const int32_t* start_3d,
                         double* srh_rm, double* srh_lm, float* rm, float* lm, float* mean6,
                         int precision, const int* devices, int ndevices) {
  int rc = srh_check(ncol, nlev, dtype, layout, XCAPE_MEM_HOST, precision, p, t, td, u, v, ps, ts, tds, us, vs, srh_rm, srh_lm);
  if (rc) return rc;
  if (!devices || ndevices < 1) return fail(XCAPE_ERR_ARG, "devices: need at least one device index");
  if (ncol == 0) return XCAPE_OK;
  return run_sharded(ncol, devices, ndevices, [&](int, int64_t c0, int64_t c1) {
    return srh_any(p, t, td, u, v, ps, ts, tds, us, vs, ncol, c0, c1 - c0, nlev, p_is_1d, dtype, layout, XCAPE_MEM_HOST, depth, aglh0,
                   start_3d, srh_rm, srh_lm, rm, lm, mean6, precision, nullptr);
  });
}

int xcape_cuda_srh_from_heights(const void* u, const void* v, const void* aglh, const void* us, const void* vs,
                                const void* aglhs, int64_t ncol, int nlev, int dtype, int layout, int mem, double depth,
                                const int32_t* start_3d, double* srh_rm, double* srh_lm, float* rm, float* lm,
                                float* mean6, int device, void* stream) {
  int rc = check_common(ncol, nlev, dtype, layout, mem);
  if (rc) return rc;
  if (ncol == 0) return XCAPE_OK;
  if (!u || !v || !aglh || !us || !vs || !aglhs || !srh_rm || !srh_lm) return fail(XCAPE_ERR_ARG, "null pointer");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(XCAPE_ERR_NODEV, "cudaSetDevice failed");
  auto dev = [&](const void* u_, const void* v_, const void* h_, const void* us_, const void* vs_, const void* hs_, int64_t n,
                 const int32_t* st_, double* srm_, double* slm_, float* rm_, float* lm_, float* m6_, cudaStream_t s) {
    if (dtype == XCAPE_F64)
      return srh_heights_device_t<double>(u_, v_, h_, us_, vs_, hs_, n, nlev, dtype, layout, n, depth, st_, srm_, slm_, rm_, lm_, m6_, s);
    return srh_heights_device_t<float>(u_, v_, h_, us_, vs_, hs_, n, nlev, dtype, layout, n, depth, st_, srm_, slm_, rm_, lm_, m6_, s);
  };
  if (mem == XCAPE_MEM_DEVICE)
    return dev(u, v, aglh, us, vs, aglhs, ncol, start_3d, srh_rm, srh_lm, rm, lm, mean6, (cudaStream_t)stream);
  const size_t es = esize(dtype);
  std::vector<HostIn3> in3 = {{u}, {v}, {aglh}};
  std::vector<HostIn1> in1 = {{us, es}, {vs, es}, {aglhs, es}};
  if (start_3d) in1.push_back({start_3d, 4});
  std::vector<HostOut> outs = {{srh_rm, 8, 0}, {srh_lm, 8, 0}, {rm, 8, 0}, {lm, 8, 0}, {mean6, 8, 0}};
  return run_staged(ncol, nlev, base_layout(layout), es, nullptr, in3, in1, outs,
                    [&](Block& b, int64_t n, cudaStream_t s) {
                      return dev(b.in3[0], b.in3[1], b.in3[2], b.in1[0], b.in1[1], b.in1[2], n,
                                 start_3d ? (const int32_t*)b.in1[3] : nullptr, (double*)b.out[0], (double*)b.out[1],
                                 rm ? (float*)b.out[2] : nullptr, lm ? (float*)b.out[3] : nullptr,
                                 mean6 ? (float*)b.out[4] : nullptr, s);
                    });
}

int xcape_cuda_stdheight(const void* p, const void* t, const void* td, const void* ps, const void* ts, const void* tds,
                         int64_t ncol, int nlev, int p_is_1d, int dtype, int layout, int mem, double aglh0,
                         const int32_t* start_3d, double* h, double* hs, int device, void* stream) {
  int rc = check_common(ncol, nlev, dtype, layout, mem);
  if (rc) return rc;
  if (ncol == 0) return XCAPE_OK;
  if (!p || !t || !td || !ps || !ts || !tds || !h || !hs) return fail(XCAPE_ERR_ARG, "null pointer");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(XCAPE_ERR_NODEV, "cudaSetDevice failed");
  auto dev = [&](const void* p_, const void* t_, const void* td_, const void* ps_, const void* ts_, const void* tds_,
                 int64_t n, const int32_t* st_, double* h_, double* hs_, cudaStream_t s) {
    if (dtype == XCAPE_F64)
      return height_device_t<double>(p_, t_, td_, ps_, ts_, tds_, n, nlev, p_is_1d, dtype, layout, n, aglh0, st_, h_, hs_, s);
    return height_device_t<float>(p_, t_, td_, ps_, ts_, tds_, n, nlev, p_is_1d, dtype, layout, n, aglh0, st_, h_, hs_, s);
  };
  if (mem == XCAPE_MEM_DEVICE) return dev(p, t, td, ps, ts, tds, ncol, start_3d, h, hs, (cudaStream_t)stream);
  const size_t es = esize(dtype);
  std::vector<HostIn3> in3 = {{t}, {td}};
  if (!p_is_1d) in3.push_back({p});
  std::vector<HostIn1> in1 = {{ps, es}, {ts, es}, {tds, es}};
  if (start_3d) in1.push_back({start_3d, 4});
  std::vector<HostOut> outs = {{h, 8, 1}, {hs, 8, 0}};
  return run_staged(ncol, nlev, base_layout(layout), es, p_is_1d ? p : nullptr, in3, in1, outs,
                    [&](Block& b, int64_t n, cudaStream_t s) {
                      return dev(p_is_1d ? b.p1d : b.in3[2], b.in3[0], b.in3[1], b.in1[0], b.in1[1], b.in1[2], n,
                                 start_3d ? (const int32_t*)b.in1[3] : nullptr, (double*)b.out[0], (double*)b.out[1], s);
                    });
}

int xcape_cuda_dewpoint_from_q(const void* p, const void* q, int64_t ncol, int nlev, int p_is_1d, int dtype, int layout,
                               int mem, double q_min, void* td, int device, void* stream) {
  int rc = check_common(ncol, nlev, dtype, layout, mem);
  if (rc) return rc;
  if (!(q_min >= 0.0) || q_min >= 1.0) return fail(XCAPE_ERR_ARG, "q_min must lie in [0, 1)");
  if (ncol == 0) return XCAPE_OK;
  if (!p || !q || !td) return fail(XCAPE_ERR_ARG, "null pointer");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(XCAPE_ERR_NODEV, "cudaSetDevice failed");
  const bool lm = base_layout(layout) == XCAPE_LEVEL_MAJOR;    // elementwise: the level ORDER is irrelevant here
  if (mem == XCAPE_MEM_DEVICE) return launch_dewpoint(p, q, td, dtype, ncol, nlev, p_is_1d != 0, lm, q_min, (cudaStream_t)stream);
  const size_t es = esize(dtype);
  std::vector<HostIn3> in3 = {{q}};
  if (!p_is_1d) in3.push_back({p});
  std::vector<HostOut> outs = {{td, es, 1}};
  return run_staged(ncol, nlev, base_layout(layout), es, p_is_1d ? p : nullptr, in3, {}, outs,
                    [&](Block& b, int64_t n, cudaStream_t s) {
                      return launch_dewpoint(p_is_1d ? b.p1d : b.in3[1], b.in3[0], b.out[0], dtype, n, nlev, p_is_1d != 0, lm, q_min, s);
                    });
}

}  // extern "C"
