"""Host side of the operand upload (csrc/host_stage.c in libdbn_hostrng.so): rows of a host dataset, gathered in epoch
order (the reference's `data = _data[idx]`, dbn/models.py:106-107) and rounded to the 16-bit operand type exactly as the
device would round them, written straight into pinned memory in the device's operand layout -- so a tensor-core fit
moves half the bytes over PCIe and needs no staging kernel.  Host-side helper: when the library is missing the
callers keep the fp32 upload + device cast (same numbers)."""
import ctypes as C
import os

import numpy as np

_FN = None         # None = not tried, False = unavailable
THREADS = max(1, min(8, (os.cpu_count() or 2) - 1))
F32, BF16, F16 = 0, 1, 2


def _load():
    global _FN
    if _FN is None:
        _FN = False
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdbn_hostrng.so")
        if os.environ.get("DBN_B200_HOST_STAGE", "1") != "0" and os.path.exists(path):
            try:
                fn = C.CDLL(path).dbn_host_stage_rows
                fn.restype = C.c_int
                fn.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t,
                               C.c_int, C.c_int, C.c_int]
                _FN = fn
            except (OSError, AttributeError):
                pass
    return _FN


def available():
    return bool(_load())


def kind_of(torch_dtype):
    import torch
    return {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}[torch_dtype]


def stage_rows(X, rows, out_ptr, ldo, kind, ones, threads=None):
    """out[i] = [cast(X[rows[i]]) | 1 if ones | 0...] for i < len(rows) (rows None: every row of X in order).
    X: C-contiguous float32 [n, cols]; rows: int64 array of valid row numbers; out_ptr: address of a buffer of
    len(rows) x ldo elements of `kind`.  Releases the GIL (ctypes), so it overlaps with the caller's other threads."""
    fn = _load()
    if not fn:
        raise RuntimeError("libdbn_hostrng.so (host_stage) is not available")
    assert X.dtype == np.float32 and X.ndim == 2 and X.flags.c_contiguous
    if rows is None:
        m, rp = int(X.shape[0]), None
    else:
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        m, rp = int(rows.shape[0]), rows.ctypes.data
    rc = fn(X.ctypes.data, int(X.shape[1]), rp, m, int(X.shape[1]), int(out_ptr), int(ldo), int(kind), int(bool(ones)),
            int(threads or THREADS))
    if rc != 0:
        raise ValueError("dbn_host_stage_rows: bad argument")
    return m
