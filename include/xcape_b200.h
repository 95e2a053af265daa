/* xcape_b200.h — C ABI of libxcape_b200.so (B200 / sm_100a column kernels for xcape).
 *
 * This is the drop-in boundary for the reference's per-column hot path.  The reference
 * (xgcm/xcape v0.1.4) crosses from Python to native code through eight f2py routines
 * (src/xcape/fortran/\*.pyf); every entry point below names the routine(s) it replaces.
 * Plain pointers and sizes only: no torch / cupy / numpy types.  Every function returns an
 * int status (0 = XCAPE_OK), never throws and never aborts the process (the reference's
 * Fortran `stop` on a bad enum, CAPE_CODE_model_lev.f90:343-352,492-497, becomes
 * XCAPE_ERR_ARG).  All functions are thread-safe (the reference routines are marked
 * `threadsafe`, e.g. CAPE_CODE_model_lev.pyf:7).
 *
 * Units follow the reference: pressure hPa, temperature / dew point degC, wind m/s,
 * heights m AGL, pinc Pa, ml_depth / depth m.
 *
 * Array layouts for the 3-D fields (t, td, u, v, and p when p_is_1d == 0):
 *   XCAPE_LEVEL_LAST   element (level k, column i) at [i*nlev + k]  — each column contiguous.
 *                      This is the reference's (nk, n2) Fortran-order f2py layout and what
 *                      core._reshape_inputs (core.py:44-50) produces from [..., nlev] arrays.
 *   XCAPE_LEVEL_MAJOR  element (k, i) at [k*ncol + i] — structure-of-arrays by level, the
 *                      native (time, level, lat, lon) order of ERA5/HRRR files; zero-copy
 *                      into the kernels when dtype is XCAPE_F32.
 * Level index 0 is the level nearest the surface (reference assumption).
 * OR-ing XCAPE_LEVELS_TOP_FIRST into `layout` declares the opposite storage order (index 0 = model
 * top, the order of ERA5 downloads): the library then walks the level axis backwards on the device
 * (no copy for level-major float32 input).  Level INDICES that cross the boundary — start_3d in,
 * mulev out — always count from the surface (1 = lowest level), whatever the storage order; 3-D
 * outputs (xcape_cuda_stdheight's h) are written in the caller's storage order.
 */
#ifndef XCAPE_B200_H
#define XCAPE_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define XCAPE_API __attribute__((visibility("default")))
#else
#define XCAPE_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define XCAPE_OK 0
#define XCAPE_ERR_ARG 1     /* invalid enum / size / null pointer */
#define XCAPE_ERR_CUDA 2    /* a CUDA call failed; see xcape_cuda_last_error() */
#define XCAPE_ERR_NODEV 3   /* no usable CUDA device */

enum { XCAPE_F32 = 0, XCAPE_F64 = 1 };                 /* dtype of ALL input arrays of a call */
enum { XCAPE_LEVEL_LAST = 0, XCAPE_LEVEL_MAJOR = 1,   /* layout of the 3-D input arrays */
       XCAPE_LEVELS_TOP_FIRST = 0x100 };               /* flag OR-ed into layout: level axis stored top -> surface */
enum { XCAPE_MEM_HOST = 0, XCAPE_MEM_DEVICE = 1 };     /* where input AND output pointers live */
enum { XCAPE_SOURCE_SURFACE = 1, XCAPE_SOURCE_MOST_UNSTABLE = 2, XCAPE_SOURCE_MIXED_LAYER = 3 }; /* core.py:302 */
enum { XCAPE_ADIABAT_PSEUDO_LIQUID = 1, XCAPE_ADIABAT_REVERSIBLE_LIQUID = 2,
       XCAPE_ADIABAT_PSEUDO_ICE = 3, XCAPE_ADIABAT_REVERSIBLE_ICE = 4 };                          /* core.py:303-304 */
/* XCAPE_FAITHFUL: IEEE binary32 chain without FMA contraction + the deterministic "SPEC"
 * transcendentals of DESIGN.md (bit-identical to oracle tmode=SPEC).  Default. */
enum { XCAPE_FAITHFUL = 0,
/* XCAPE_FAST (CAPE only): identical prep / source selection (MU level index stays exact); the
 * moist fixed-point body runs on the FP32 pipe with FMAs, a ~1-ulp expf and MUFU reciprocals.
 * CAPE / CIN then agree with the reference within max(1 J/kg, 1e-4 relative) except on
 * ill-conditioned columns (non-convergence limit cycles, CIN sign flips). */
       XCAPE_FAST = 1,
/* XCAPE_FAST additionally replaces the reference's damped fixed-point iteration of a sub-step
 * (x += 0.3 g, ~10 passes) by a safeguarded secant solve of the same equation with the same stopping
 * rule (3-5 passes) and advances the Exner function incrementally between level anchors.  Sub-steps on
 * which the reference's own iteration cannot converge (slope of the map below ~-6: limit cycle, status 2,
 * cape = cin = 0) are detected from the secant slope and decided by the reference's
 * iteration, so the non-convergence semantics are the reference's.
 * XCAPE_FAST_RELAXED keeps the reference's iteration everywhere (identical pass counts) and only uses the
 * fast arithmetic.  XCAPE_FAST_OPTIMISTIC is XCAPE_FAST without that detection: where the reference gives
 * up, the converged secant value is returned with status 0. */
       XCAPE_FAST_RELAXED = 2,
       XCAPE_FAST_OPTIMISTIC = 3 };
/* per-column status word (optional output) */
enum { XCAPE_ST_OK = 0, XCAPE_ST_SKIPPED = 1 /* ts <= 0 degC gate, f90:77 */,
       XCAPE_ST_NONCONVERGED = 2 /* > 100 moist iterations, f90:464-474: cape = cin = 0 */,
       XCAPE_ST_INVALID = 3 /* non-finite / absurd pressure step (fill values, NaN) or > 2^22 moist passes:
                               cape = cin = 0; the reference has undefined behaviour here (int overflow).
                               Guarantees that no input can make a thread spin. */ };

/* ---------------------------------------------------------------------------------------
 * CAPE / CIN.  Replaces loopcape_ml (CAPE_CODE_model_lev.pyf:6-24, f90:4-91) when
 * p_is_1d == 0 and loopcape_pl1d (CAPE_CODE_pressure_lev.pyf:26-45, f90:88-169) when
 * p_is_1d == 1, including getcape_ml/getcape_pl, getqvs, getqvi, getthe (f90:97-620), and
 * the numpy `pres_lev_pos` pre-step of core.py:286-289 when start_3d == NULL.
 *
 *   p        p_is_1d ? [nlev] : 3-D field           ps, ts, tds   [ncol]
 *   t, td    3-D fields                             start_3d      [ncol] 1-based first used
 *                                                   level (int32) or NULL (computed on device
 *                                                   for p_is_1d, 1 otherwise)
 *   outputs  cape, cin, zmulev float32 [ncol]; mulev int32 [ncol]; status int32 [ncol] or NULL;
 *            n_iter int32 [ncol] or NULL = number of moist-adiabat iterations the column ran
 *            (the work counter behind the roofline figure).
 *            mulev / zmulev follow the reference's conventions.
 *   mem      XCAPE_MEM_DEVICE: every pointer is a device pointer on `device`; the work is
 *            enqueued on `stream` (a cudaStream_t / CUstream, NULL = default stream) and the
 *            call returns without synchronising.  XCAPE_MEM_HOST: pointers are host memory;
 *            the call stages column blocks through pinned buffers, overlaps copies with
 *            compute, and returns when the outputs are complete (`stream` is ignored).
 * ------------------------------------------------------------------------------------- */
XCAPE_API int xcape_cuda_cape(const void* p, const void* t, const void* td,
                    const void* ps, const void* ts, const void* tds,
                    int64_t ncol, int nlev, int p_is_1d, int dtype, int layout, int mem,
                    int source, int adiabat, float ml_depth, float pinc,
                    const int32_t* start_3d,
                    float* cape, float* cin, int32_t* mulev, float* zmulev, int32_t* status,
                    int32_t* n_iter, int precision, int device, void* stream);

/* xcape_cuda_cape on HOST memory over several GPUs of one box: [0, ncol) is split into contiguous blocks of whole
 * 128-column units, one per entry of `devices` (an index may repeat), each block runs the host path of xcape_cuda_cape on
 * its device from its own host thread; no collective, no peer traffic (columns are independent).  Replaces
 * the chunk-level parallelism the reference gets from dask (core.py:237-258) for one in-memory field.  Returns the
 * first failing device's code; outputs of the other blocks are complete. */
XCAPE_API int xcape_cuda_cape_multi(const void* p, const void* t, const void* td,
                    const void* ps, const void* ts, const void* tds,
                    int64_t ncol, int nlev, int p_is_1d, int dtype, int layout,
                    int source, int adiabat, float ml_depth, float pinc,
                    const int32_t* start_3d,
                    float* cape, float* cin, int32_t* mulev, float* zmulev, int32_t* status,
                    int32_t* n_iter, int precision, const int* devices, int ndevices);

/* ---------------------------------------------------------------------------------------
 * Storm-relative helicity, fused.  Replaces, in one pass over the column,
 *   loop_stdheight_ml / loop_stdheight_pl1d  (stdheight_2D_model_lev.pyf:6-19,
 *                                             stdheight_2D_pressure_lev.pyf:21-35),
 *   bunkers_loop_ml / bunkers_loop_pl        (Bunkers_model_lev.pyf:8-21, Bunkers_pressure_lev.pyf:6-20),
 *   loop_sreh_ml / loop_sreh_pl              (SREH_model_lev.pyf:6-23, SREH_pressure_lev.pyf:6-24)
 * as chained by core._calc_srh_numpy (core.py:516-535) and srh.srh (srh.py:41-61).
 *   srh_rm, srh_lm  float64 [ncol]
 *   rm, lm, mean6   float32 [2*ncol] in the reference's (2, n2) Fortran order (component c of
 *                   column i at [2*i + c]) or NULL (output_var == 'srh')
 *   aglh0           height of the surface level (core.py passes 2.0)
 * ------------------------------------------------------------------------------------- */
XCAPE_API int xcape_cuda_srh(const void* p, const void* t, const void* td, const void* u, const void* v,
                   const void* ps, const void* ts, const void* tds, const void* us, const void* vs,
                   int64_t ncol, int nlev, int p_is_1d, int dtype, int layout, int mem,
                   double depth, double aglh0, const int32_t* start_3d,
                   double* srh_rm, double* srh_lm, float* rm, float* lm, float* mean6,
                   int precision, int device, void* stream);

/* xcape_cuda_srh on HOST memory over several GPUs of one box (see xcape_cuda_cape_multi). */
XCAPE_API int xcape_cuda_srh_multi(const void* p, const void* t, const void* td, const void* u, const void* v,
                   const void* ps, const void* ts, const void* tds, const void* us, const void* vs,
                   int64_t ncol, int nlev, int p_is_1d, int dtype, int layout,
                   double depth, double aglh0, const int32_t* start_3d,
                   double* srh_rm, double* srh_lm, float* rm, float* lm, float* mean6,
                   int precision, const int* devices, int ndevices);

/* The reference's two-call form: heights already computed (by xcape_cuda_stdheight or by the
 * caller).  Replaces bunkers_loop_ml / bunkers_loop_pl + loop_sreh_ml / loop_sreh_pl exactly as
 * srh.srh chains them (srh.py:41-61): u, v, aglh 3-D fields, us, vs, aglhs [ncol]; levels below
 * start_3d (NULL = 1) are ignored.  Heights are used as binary64 by the helicity sum and
 * down-cast to binary32 by the Bunkers part, like the f2py casts do. */
XCAPE_API int xcape_cuda_srh_from_heights(const void* u, const void* v, const void* aglh,
                                const void* us, const void* vs, const void* aglhs,
                                int64_t ncol, int nlev, int dtype, int layout, int mem,
                                double depth, const int32_t* start_3d,
                                double* srh_rm, double* srh_lm, float* rm, float* lm, float* mean6,
                                int device, void* stream);

/* Heights only (loop_stdheight_ml / loop_stdheight_pl1d).  h is float64, same layout as the
 * inputs' `layout`; levels below start_3d are -999999 (stdheight_2D_pressure_lev.f90:85-87);
 * hs [ncol] = aglh0. */
XCAPE_API int xcape_cuda_stdheight(const void* p, const void* t, const void* td,
                         const void* ps, const void* ts, const void* tds,
                         int64_t ncol, int nlev, int p_is_1d, int dtype, int layout, int mem,
                         double aglh0, const int32_t* start_3d,
                         double* h, double* hs, int device, void* stream);

/* core.py:286-289 on the device: start[i] = 1 + argmin_k { ps[i] - p[k] : ps[i] - p[k] >= 0 }
 * (first minimum; 1 when every level is masked), evaluated in the inputs' own dtype. */
XCAPE_API int xcape_cuda_pres_lev_pos(const void* p, const void* ps, int64_t ncol, int nlev, int dtype,
                            int mem, int32_t* start_3d, int device, void* stream);

/* Diagnostics. */
XCAPE_API const char* xcape_cuda_last_error(void);      /* thread-local message of the last failure */
XCAPE_API int xcape_cuda_device_count(void);            /* < 0 on error */
XCAPE_API const char* xcape_cuda_version(void);         /* "xcape_b200 <semver> sm_100a" */
XCAPE_API int64_t xcape_cuda_kernel_launches(void);     /* kernels launched by this library so far (process-wide) */
/* xcape_cuda_cape with HOST pointers on a shared pressure axis copies only the levels the ascent can reach (up to the
 * first one with p <= 100 hPa, where the reference stops a negatively buoyant parcel, CAPE_CODE_model_lev.f90:554-557);
 * columns still ascending there are redone with every level.  This counts those columns (process-wide).
 * XCAPE_B200_SHIP_ALL_LEVELS=1 in the environment disables the optimisation. */
XCAPE_API int64_t xcape_cuda_columns_redone(void);
/* The library keeps its stream-ordered scratch (relayout buffers, staging blocks) in a private
 * cudaMemPool per device and never returns it to the driver on its own (re-mapping ~300 MB per call cost
 * 5 ms per ERA5 field).  This call synchronises `device` and hands the cached memory back, along
 * with the idle pinned host staging buffers of the host-pointer path. */
XCAPE_API int xcape_cuda_release_memory(int device);
/* Dew point (degC) from pressure (hPa) and specific humidity q (kg/kg), elementwise over a 3-D field —
 * the conversion ERA5 users run ahead of calc_cape (the reference ships none: doc/tutorial.rst:19-23;
 * SURVEY 8f-3).  The inverse of the kernels' own saturation law (getqvs, CAPE_CODE_model_lev.f90:570-581):
 * r = q/(1-q), e = p r/(eps + r), L = ln(e/6.112), Td = 243.5 L/(17.67 - L).  q and td: `dtype`, same
 * layout; p: [nlev] when p_is_1d, else like q.  q below q_min is raised to q_min first (0 keeps q as is;
 * q <= 0 then yields NaN / -inf like the formula). */
XCAPE_API int xcape_cuda_dewpoint_from_q(const void* p, const void* q, int64_t ncol, int nlev, int p_is_1d, int dtype,
                                         int layout, int mem, double q_min, void* td, int device, void* stream);
/* Measured arithmetic peaks of `device` (roofline denominators the driver's MEASURED_PEAKS.json
 * lacks): dependent-chain-free FFMA / DFMA loops, 2 flop per FMA, best of `reps` launches. */
XCAPE_API int xcape_cuda_measure_peaks(int device, int reps, double* fp32_tflops, double* fp64_tflops);
/* FFMA rate with three distinct register operands per instruction (no uniform / immediate operands, no reuse): the
 * rate the register file lets a real kernel sustain — ~0.70 of the figure above on B200.  Reported beside it. */
XCAPE_API int xcape_cuda_measure_fp32_rrr(int device, int reps, double* fp32_tflops);

/* Device time of the DOMINANT kernel of a call, for roofline accounting (bench.py): after xcape_cuda_time_kernels(1)
 * the calling thread's device-pointer calls record a CUDA event pair on their stream around the column kernel alone
 * (the CAPE ascent kernel, without the source-parcel / ordering kernels of its sorted execution; the streaming SRH
 * kernel); xcape_cuda_last_kernel_ms waits for the pair recorded last by this thread and returns the elapsed
 * milliseconds.  Off by default (no events are recorded). */
XCAPE_API int xcape_cuda_time_kernels(int enable);
XCAPE_API int xcape_cuda_last_kernel_ms(double* ms);

#ifdef __cplusplus
}
#endif
#endif /* XCAPE_B200_H */
