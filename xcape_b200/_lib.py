"""ctypes binding of ``libxcape_b200.so`` (C ABI: ``include/xcape_b200.h``).

There is NO CPU fallback: if the CUDA extension has not been built this module raises, and
every ``method='cuda'`` call fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('XCAPE_B200_LIB') or os.path.join(_HERE, 'libxcape_b200.so')   # env override: A/B experiments

# enums of include/xcape_b200.h
F32, F64 = 0, 1
LEVEL_LAST, LEVEL_MAJOR = 0, 1
LEVELS_TOP_FIRST = 0x100      # flag OR-ed into the layout: level axis stored top -> surface
MEM_HOST, MEM_DEVICE = 0, 1
FAITHFUL, FAST, FAST_RELAXED, FAST_OPTIMISTIC = 0, 1, 2, 3
PRECISION = {'faithful': FAITHFUL, 'fast': FAST, 'fast-relaxed': FAST_RELAXED, 'fast-optimistic': FAST_OPTIMISTIC}
OK, ERR_ARG, ERR_CUDA, ERR_NODEV = 0, 1, 2, 3

# every symbol include/xcape_b200.h declares (tests check the library exports all of them)
SYMBOLS = ('xcape_cuda_cape', 'xcape_cuda_cape_multi', 'xcape_cuda_srh', 'xcape_cuda_srh_multi', 'xcape_cuda_srh_from_heights', 'xcape_cuda_stdheight', 'xcape_cuda_pres_lev_pos',
           'xcape_cuda_last_error', 'xcape_cuda_device_count', 'xcape_cuda_version',
           'xcape_cuda_kernel_launches', 'xcape_cuda_measure_peaks', 'xcape_cuda_release_memory',
           'xcape_cuda_dewpoint_from_q', 'xcape_cuda_columns_redone', 'xcape_cuda_measure_fp32_rrr',
           'xcape_cuda_time_kernels', 'xcape_cuda_last_kernel_ms')

_lib = None


class XcapeCudaError(RuntimeError):
    pass


def build():
    """Compile the extension in-tree with nvcc for sm_100a (works without a GPU)."""
    import subprocess
    subprocess.check_call(['make', '-C', os.path.join(_HERE, 'csrc'), '-s', '-j4'])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f'{LIB_PATH} is missing: the CUDA extension is not built. Run '
                '`python -c "import __graft_entry__ as g; g.build()"` (or `make -C xcape_b200/csrc`). '
                'xcape_b200 has no CPU fallback for method="cuda".')
        L = C.CDLL(LIB_PATH)
        vp, i64, i32, f32, f64 = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_double
        L.xcape_cuda_cape.restype = i32
        L.xcape_cuda_cape.argtypes = [vp] * 6 + [i64, i32, i32, i32, i32, i32, i32, i32, f32, f32, vp,
                                               vp, vp, vp, vp, vp, vp, i32, i32, vp]
        L.xcape_cuda_cape_multi.restype = i32
        L.xcape_cuda_cape_multi.argtypes = [vp] * 6 + [i64, i32, i32, i32, i32, i32, i32, f32, f32, vp,
                                                     vp, vp, vp, vp, vp, vp, i32, C.POINTER(i32), i32]
        L.xcape_cuda_srh.restype = i32
        L.xcape_cuda_srh.argtypes = [vp] * 10 + [i64, i32, i32, i32, i32, i32, f64, f64, vp,
                                               vp, vp, vp, vp, vp, i32, i32, vp]
        L.xcape_cuda_srh_multi.restype = i32
        L.xcape_cuda_srh_multi.argtypes = [vp] * 10 + [i64, i32, i32, i32, i32, f64, f64, vp,
                                                     vp, vp, vp, vp, vp, i32, C.POINTER(i32), i32]
        L.xcape_cuda_srh_from_heights.restype = i32
        L.xcape_cuda_srh_from_heights.argtypes = [vp] * 6 + [i64, i32, i32, i32, i32, f64, vp, vp, vp, vp, vp, vp, i32, vp]
        L.xcape_cuda_stdheight.restype = i32
        L.xcape_cuda_stdheight.argtypes = [vp] * 6 + [i64, i32, i32, i32, i32, i32, f64, vp, vp, vp, i32, vp]
        L.xcape_cuda_pres_lev_pos.restype = i32
        L.xcape_cuda_pres_lev_pos.argtypes = [vp, vp, i64, i32, i32, i32, vp, i32, vp]
        L.xcape_cuda_dewpoint_from_q.restype = i32
        L.xcape_cuda_dewpoint_from_q.argtypes = [vp, vp, i64, i32, i32, i32, i32, i32, f64, vp, i32, vp]
        L.xcape_cuda_last_error.restype = C.c_char_p
        L.xcape_cuda_version.restype = C.c_char_p
        L.xcape_cuda_device_count.restype = i32
        L.xcape_cuda_kernel_launches.restype = i64
        L.xcape_cuda_columns_redone.restype = i64
        L.xcape_cuda_release_memory.restype = i32
        L.xcape_cuda_release_memory.argtypes = [i32]
        L.xcape_cuda_measure_peaks.restype = i32
        L.xcape_cuda_measure_peaks.argtypes = [i32, i32, C.POINTER(f64), C.POINTER(f64)]
        L.xcape_cuda_measure_fp32_rrr.restype = i32
        L.xcape_cuda_measure_fp32_rrr.argtypes = [i32, i32, C.POINTER(f64)]
        L.xcape_cuda_time_kernels.restype = i32
        L.xcape_cuda_time_kernels.argtypes = [i32]
        L.xcape_cuda_last_kernel_ms.restype = i32
        L.xcape_cuda_last_kernel_ms.argtypes = [C.POINTER(f64)]
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        msg = lib().xcape_cuda_last_error().decode(errors='replace')
        if rc == ERR_ARG:
            raise ValueError(f'xcape_b200: {msg}')
        raise XcapeCudaError(f'xcape_b200 (code {rc}): {msg}')


def kernel_launches():
    return int(lib().xcape_cuda_kernel_launches())


def columns_redone():
    """Columns the host path of ``xcape_cuda_cape`` redid with all levels so far (see include/xcape_b200.h)."""
    return int(lib().xcape_cuda_columns_redone())


def measure_peaks(device=0, reps=5):
    """(fp32_tflops, fp64_tflops) of FMA microbenchmarks on ``device`` (2 flop per FMA)."""
    a, b = C.c_double(0.0), C.c_double(0.0)
    check(lib().xcape_cuda_measure_peaks(int(device), int(reps), C.byref(a), C.byref(b)))
    return a.value, b.value


def measure_fp32_rrr(device=0, reps=5):
    """FFMA TFLOP/s with three distinct register operands per instruction (what the register file sustains)."""
    a = C.c_double(0.0)
    check(lib().xcape_cuda_measure_fp32_rrr(int(device), int(reps), C.byref(a)))
    return a.value


def time_kernels(enable=True):
    """Record a CUDA event pair around the dominant column kernel of this thread's device-pointer calls."""
    check(lib().xcape_cuda_time_kernels(1 if enable else 0))


def last_kernel_ms():
    """Device milliseconds of the dominant kernel of this thread's last timed call (waits for it)."""
    a = C.c_double(0.0)
    check(lib().xcape_cuda_last_kernel_ms(C.byref(a)))
    return a.value


def release_memory(device=0):
    """Return the library's cached device scratch on ``device`` to the CUDA driver."""
    check(lib().xcape_cuda_release_memory(int(device)))


def device_count():
    return int(lib().xcape_cuda_device_count())
