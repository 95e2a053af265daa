"""Streamed multi-time-step execution (SURVEY.md §8f-4, §8e "map blocks round-robin to devices").

Reanalysis and forecast archives hold one ``(level, lat, lon)`` block per time step, far more steps
than fit in host memory at once (C5: 24 steps x 1.04 M columns x 137 levels = 41 GB).  ``stream_cape``
/ ``stream_srh`` walk such a sequence with a bounded pipeline:

    reader threads   page step k+1, k+2 ... in from disk (``np.load(..., mmap_mode='r')`` arrays,
                     raw ``np.memmap`` views, or any zero-argument loader) into host memory
    device workers   one host thread per GPU, each pulling the next loaded step and calling
                     ``calc_cape`` / ``calc_srh`` (whose C call releases the GIL and runs its own
                     copy/compute ring), so steps are dealt round-robin to the GPUs with no collective
    the caller       receives results strictly in step order through a generator

Nothing here touches the numerics: every step goes through ``xcape_b200.core`` exactly as a direct
call would.
"""
import queue
import threading

import numpy as np

from . import core

_STOP = object()


class _PinnedPool:
    """Page-locked host buffers for the reader threads, recycled by shape and dtype.  A step read
    straight into pinned memory is DMA-ed by the library as is; an ordinary array would first be copied
    into the library's own pinned staging ring — a second pass over host memory per field."""

    def __init__(self):
        self.free, self.lock = {}, threading.Lock()

    def take(self, shape, dtype):
        import torch
        key = (tuple(shape), np.dtype(dtype).str)
        with self.lock:
            if self.free.get(key):
                return self.free[key].pop()
        t = torch.empty(tuple(shape), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
        t._xcape_pool_key = key
        return t

    def give(self, t):
        with self.lock:
            self.free.setdefault(t._xcape_pool_key, []).append(t)


    def clear(self):
        with self.lock:
            self.free.clear()


_POOL = _PinnedPool()      # process-wide: page-locking ~150 MB costs tens of ms, a stream re-uses the same few shapes


def release_buffers():
    """Drop the idle pinned read buffers kept from earlier streams."""
    _POOL.clear()


def _is_mapped(a):
    return isinstance(a, np.memmap) or (isinstance(a, np.ndarray) and isinstance(getattr(a, 'base', None), np.memmap))


def _materialise(step, pool=None, min_pinned_bytes=1 << 20):
    """A step is a tuple of arrays or a zero-argument callable returning one.  Memory-mapped
    arrays are read into host memory here, on the reader thread — that is the disk I/O.  With a
    ``pool`` the large ones land in recycled pinned buffers; returns ``(arrays, borrowed buffers)``."""
    if callable(step):
        step = step()
    out, borrowed = [], []
    for a in step:
        if _is_mapped(a):
            if pool is not None and a.nbytes >= min_pinned_bytes and a.dtype in (np.float32, np.float64):
                t = pool.take(a.shape, a.dtype)
                borrowed.append(t)
                dst = t.numpy()
                np.copyto(dst, a)
                a = dst
            else:
                a = np.array(a)                    # dense copy in ordinary memory
        out.append(a)
    return tuple(out), borrowed


def _stream(fn, steps, devices, prefetch, readers, kwargs):
    devices = list(devices) if devices is not None else [kwargs.get('device', 0)]
    kwargs.pop('device', None)
    pinned = kwargs.pop('pinned', None)
    if pinned is None:                             # default: pin when a CUDA-enabled torch is there to do it
        try:
            import torch
            pinned = torch.cuda.is_available()
        except ImportError:
            pinned = False
    if not devices:
        raise ValueError('devices must name at least one GPU')
    if prefetch < 1 or readers < 1:
        raise ValueError('prefetch and readers must be >= 1')
    steps = list(steps)                            # cheap: tuples of views / loaders, not data
    return _run(fn, steps, devices, prefetch, readers, kwargs, _POOL if pinned else None)


def _run(fn, steps, devices, prefetch, readers, kwargs, pool):
    todo = queue.Queue()                           # (index, step) for the readers
    for item in enumerate(steps):
        todo.put(item)
    loaded = queue.Queue(maxsize=prefetch)         # (index, arrays): bounds the host memory in flight
    results, cond = {}, threading.Condition()
    failed = []
    state = {'readers_left': readers}

    def fail(e):
        with cond:
            failed.append(e)
            cond.notify_all()

    def put(item):                                 # a bounded put that gives up once the run has failed
        while not failed:
            try:
                loaded.put(item, timeout=0.1)
                return
            except queue.Full:
                pass

    def reader():
        try:
            while not failed:
                try:
                    k, s = todo.get_nowait()
                except queue.Empty:
                    break
                put((k, *_materialise(s, pool)))
        except BaseException as e:  # noqa: BLE001 - handed to the consumer
            fail(e)
        finally:
            with cond:
                state['readers_left'] -= 1
                last = state['readers_left'] == 0
            if last:                               # every step is loaded: release the device workers
                for _ in devices:
                    put(_STOP)

    def worker(dev):
        try:
            while not failed:
                try:
                    item = loaded.get(timeout=0.1)
                except queue.Empty:
                    continue
                if item is _STOP:
                    break
                k, arrays, borrowed = item
                r = fn(*arrays, device=dev, **kwargs)
                for t in borrowed:                 # the call has returned: its inputs are free again
                    pool.give(t)
                with cond:
                    results[k] = r
                    cond.notify_all()
        except BaseException as e:  # noqa: BLE001
            fail(e)

    ths = [threading.Thread(target=reader, daemon=True) for _ in range(readers)]
    ths += [threading.Thread(target=worker, args=(d,), daemon=True) for d in devices]
    for t in ths:
        t.start()
    try:
        for k in range(len(steps)):
            with cond:
                cond.wait_for(lambda: k in results or failed)
                if failed:
                    raise failed[0]
                r = results.pop(k)
            yield r
    finally:
        if not failed:
            fail(GeneratorExit())                  # consumer stopped early: wind the threads down
        for t in ths:
            t.join()


def stream_cape(steps, *, devices=None, prefetch=2, readers=1, **kwargs):
    """Run ``calc_cape`` over a sequence of time steps; yields each step's result tuple in order.

    ``steps``: iterable of ``(p, t, td, ps, ts, tds)`` tuples — numpy arrays, ``np.memmap`` /
    ``np.load(mmap_mode='r')`` views (read from disk by the reader threads), or zero-argument callables
    returning such a tuple.  ``devices``: GPUs the steps are dealt to (one host thread each; default
    the single ``device``).  ``prefetch``: steps held in host memory ahead of the GPUs.  ``readers``:
    threads paging steps in.  ``pinned`` (default: on when CUDA is available): read memory-mapped fields
    straight into recycled page-locked buffers, which the library then DMA-s without its own staging
    copy.  Remaining keyword arguments go to ``calc_cape`` unchanged.
    """
    return _stream(core.calc_cape, steps, devices, prefetch, readers, dict(kwargs))


def stream_srh(steps, *, devices=None, prefetch=2, readers=1, **kwargs):
    """As ``stream_cape`` for ``calc_srh``: steps are ``(p, t, td, u, v, ps, ts, tds, us, vs)``."""
    return _stream(core.calc_srh, steps, devices, prefetch, readers, dict(kwargs))
