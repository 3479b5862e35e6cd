"""Device context: one per GPU, owns the stream, scratch and the launch/profile counters."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, dptr

_contexts = {}


class Context:
    def __init__(self, device=0):
        self.lib = _lib.load()
        self.device = int(device)
        h = C.c_void_p()
        check(self.lib.gpd_create(self.device, C.byref(h)))
        self.handle = h

    def close(self):
        if self.handle:
            self.lib.gpd_destroy(self.handle)
            self.handle = None
            _contexts.pop(self.device, None)

    # ---- knobs / accounting -------------------------------------------------------------------
    def set_scratch_bytes(self, nbytes):
        check(self.lib.gpd_set_scratch_bytes(self.handle, int(nbytes)))

    def set_option(self, name, value):
        check(self.lib.gpd_set_option(self.handle, name.encode(), int(value)))

    def check_scratch(self):
        """(number of red zones, overwritten bytes) of the current scratch layout; needs set_option('scratch_redzones', 1)."""
        z, b = C.c_int64(), C.c_int64()
        check(self.lib.gpd_debug_check_scratch(self.handle, C.byref(z), C.byref(b)))
        return int(z.value), int(b.value)

    def launch_count(self):
        return int(self.lib.gpd_launch_count(self.handle))

    def profile_enable(self, on=True):
        check(self.lib.gpd_profile_enable(self.handle, 1 if on else 0))

    def profile_read(self):
        ms = np.zeros(_lib.NFAMILIES)
        fl = np.zeros(_lib.NFAMILIES)
        nl = (C.c_uint64 * _lib.NFAMILIES)()
        check(self.lib.gpd_profile_read(self.handle, dptr(ms), nl, dptr(fl)))
        return {name: {'ms': float(ms[i]), 'launches': int(nl[i]), 'flops': float(fl[i])}
                for i, name in enumerate(_lib.FAMILIES)}

    def timer_start(self):
        check(self.lib.gpd_timer_start(self.handle))

    def timer_stop(self):
        ms = C.c_double()
        check(self.lib.gpd_timer_stop(self.handle, C.byref(ms)))
        return ms.value

    def sync(self):
        check(self.lib.gpd_sync(self.handle))

    def probe_fp64_peak(self):
        v = C.c_double()
        check(self.lib.gpd_probe_fp64_peak(self.handle, C.byref(v)))
        return v.value

    def flush_l2(self):
        check(self.lib.gpd_flush_l2(self.handle))

    # ---- device buffers -----------------------------------------------------------------------
    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        check(self.lib.gpd_dev_alloc(self.handle, int(nbytes), C.byref(p)))
        return p

    def dev_free(self, p):
        check(self.lib.gpd_dev_free(self.handle, p))

    def dev_upload(self, dst, arr):
        arr = np.ascontiguousarray(arr)
        check(self.lib.gpd_dev_upload(self.handle, dst, arr.ctypes.data_as(C.c_void_p), arr.nbytes))

    def dev_download(self, src, shape, dtype=np.float64):
        out = np.empty(shape, dtype=dtype)
        check(self.lib.gpd_dev_download(self.handle, out.ctypes.data_as(C.c_void_p), src, out.nbytes))
        return out

    def pinned_empty(self, shape, dtype=np.float64):
        """Page-locked host array (freed with the context's process; for end-to-end staging)."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        check(self.lib.gpd_host_alloc_pinned(self.handle, nbytes, C.byref(p)))
        buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        return arr

    def pooled_pinned(self, shape, dtype=np.float64):
        """Page-locked array from a per-context pool: the buffer returns to the pool when the array (and every view
        of it) has been garbage collected, so the drop-in functions can hand out pinned results call after call
        without paying cudaHostAlloc each time (a D2H copy into pageable memory is staged and ~4x slower)."""
        import weakref
        count = int(np.prod(shape))
        nbytes = max(count * np.dtype(dtype).itemsize, 1)
        cap = 1 << max(16, (nbytes - 1).bit_length())
        pool = self.__dict__.setdefault('_pin_pool', {})
        free = pool.setdefault(cap, [])
        if free:
            addr = free.pop()
        else:
            p = C.c_void_p()
            check(self.lib.gpd_host_alloc_pinned(self.handle, cap, C.byref(p)))
            addr = p.value
        buf = (C.c_char * cap).from_address(addr)
        weakref.finalize(buf, free.append, addr)           # numpy keeps `buf` alive as the base of every view
        return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)

    def fill_uniform(self, X_dev, N, D, seed, first_row=0, lo=None, hi=None):
        lo = np.zeros(D) if lo is None else _lib.as_f64(lo)
        hi = np.ones(D) if hi is None else _lib.as_f64(hi)
        check(self.lib.gpd_dev_fill_uniform(self.handle, X_dev, int(N), int(D), int(seed), int(first_row),
                                            dptr(lo), dptr(hi)))

    def argmax(self, dc):
        dc = _lib.as_f64(dc).ravel()
        v = C.c_double()
        i = C.c_int64()
        check(self.lib.gpd_argmax(self.handle, dptr(dc), dc.size, C.byref(v), C.byref(i)))
        return v.value, int(i.value)


_default_device = None


def set_default_device(device):
    """Device used by the entry points that take no model (design / discrimination criteria, marginal assembly)."""
    global _default_device
    _default_device = int(device)


def default_device():
    """set_default_device() value, else LOCAL_RANK (one process per GPU under torchrun), else 0."""
    if _default_device is not None:
        return _default_device
    import os
    return int(os.environ.get('LOCAL_RANK', 0))


def get_context(device=None):
    """Process-wide context of a device (created on first use; raises if there is no B200)."""
    if device is None:
        device = default_device()
    ctx = _contexts.get(int(device))
    if ctx is None:
        ctx = Context(device)
        _contexts[int(device)] = ctx
    return ctx
