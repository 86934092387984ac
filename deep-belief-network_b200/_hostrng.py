"""np.random.randn from the global legacy stream, bit for bit, several times faster (csrc/host_rng.c).

The reference draws every RBM's parameters with `np.random.randn` (dbn/models.py:55-69) and this backend keeps
that contract -- same seed, same initial weights.  `legacy_randn(*shape)` reads the state of NumPy's global
`RandomState` through the public `get_state()`, lets libdbn_hostrng.so produce exactly the values (and the final
state) `np.random.randn(*shape)` would, and writes the state back with `set_state()`.  If the helper library is
missing, the global generator is not MT19937, or the self-check against NumPy fails, it simply calls NumPy: this is
host-side convenience, not part of the compute path (which has no fallback)."""
import ctypes as C
import os
import warnings

import numpy as np

_LIB = None        # None = not tried, False = unavailable
_LIB_F32 = None    # dbn_host_legacy_randn_f32 of the same library
_THREADS = max(1, min(8, (os.cpu_count() or 2) - 1))   # worker threads (the calling thread produces the MT19937 words meanwhile)
_PAR_MIN = int(os.environ.get("DBN_B200_RANDN_PAR_MIN", 1 << 18))   # draws from which the threads pay for their start-up (tools/randn_probe.py)


def _threads_for(n):
    return _THREADS if n >= _PAR_MIN else 1


def _raw(fn, rs_state, n):
    key = np.ascontiguousarray(rs_state[1], dtype=np.uint32).copy()
    pos, has_gauss, gauss = C.c_int(int(rs_state[2])), C.c_int(int(rs_state[3])), C.c_double(float(rs_state[4]))
    out = np.empty(n, dtype=np.float64)
    rc = fn(key.ctypes.data, C.byref(pos), C.byref(has_gauss), C.byref(gauss), out.ctypes.data, n, _threads_for(n))
    if rc != 0:
        return None, None
    return out, ('MT19937', key, pos.value, has_gauss.value, gauss.value)


def _self_check(fn):
    """The helper must reproduce NumPy on a private generator: values and state, odd and even lengths, with and
    without a cached Gaussian, across 624-word block boundaries."""
    a, b = np.random.RandomState(20240531), np.random.RandomState(20240531)
    for n in (1, 2, 311, 4097, 20000, 3):
        ref = a.randn(n)
        got, st = _raw(fn, b.get_state(), n)
        if got is None or not np.array_equal(ref.view(np.uint64), got.view(np.uint64)):
            return False
        b.set_state(st)
        sa, sb = a.get_state(), b.get_state()
        if not (np.array_equal(sa[1], sb[1]) and tuple(sa[2:]) == tuple(sb[2:])):
            return False
    return True


def _load():
    global _LIB
    if _LIB is None:
        _LIB = False
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdbn_hostrng.so")
        if os.environ.get("DBN_B200_FAST_RANDN", "1") != "0" and os.path.exists(path):
            try:
                fn = C.CDLL(path).dbn_host_legacy_randn
                fn.restype = C.c_int
                fn.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_void_p,
                               C.c_size_t, C.c_int]
                if _self_check(fn):
                    _LIB = fn
                    global _LIB_F32
                    f32 = C.CDLL(path).dbn_host_legacy_randn_f32
                    f32.restype = C.c_int
                    f32.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_void_p,
                                    C.c_size_t, C.c_double, C.c_int]
                    _LIB_F32 = f32
                else:
                    warnings.warn("dbn_b200: libdbn_hostrng.so does not reproduce np.random.randn here; using NumPy")
            except (OSError, AttributeError):
                pass
    return _LIB


def available():
    return bool(_load())


def legacy_randn(*shape):
    """Same values, same effect on the global np.random state as np.random.randn(*shape)."""
    fn = _load()
    n = int(np.prod(shape)) if shape else 1
    if not fn or n < 4096:                      # small draws: NumPy is as fast
        return np.random.randn(*shape)
    st = np.random.get_state()
    if st[0] != 'MT19937':
        return np.random.randn(*shape)
    out, new_state = _raw(fn, st, n)
    if out is None:
        return np.random.randn(*shape)
    np.random.set_state(new_state)
    return out.reshape(shape) if shape else float(out[0])


def legacy_randn_f32(out, divisor):
    """out[...] = float32(np.random.randn(*out.shape) / divisor) with the same effect on the global np.random state:
    initial weights in the form the device stores them, written straight into `out` (a C-contiguous float32 array,
    typically pinned memory about to be uploaded) -- no float64 array of the layer's size on the way."""
    assert out.dtype == np.float32 and out.flags.c_contiguous
    n = int(out.size)
    _load()
    st = np.random.get_state() if (_LIB_F32 and n >= 4096) else None
    if st is not None and st[0] == 'MT19937':
        key = np.ascontiguousarray(st[1], dtype=np.uint32).copy()
        pos, has_gauss, gauss = C.c_int(int(st[2])), C.c_int(int(st[3])), C.c_double(float(st[4]))
        rc = _LIB_F32(key.ctypes.data, C.byref(pos), C.byref(has_gauss), C.byref(gauss), out.ctypes.data, n, float(divisor),
                      _threads_for(n))
        if rc == 0:
            np.random.set_state(('MT19937', key, pos.value, has_gauss.value, gauss.value))
            return out
        if rc != -1:
            raise RuntimeError("dbn_host_legacy_randn_f32 failed after consuming part of the stream")
    out[...] = np.random.randn(*out.shape) / divisor
    return out
