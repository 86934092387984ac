"""CPU-side checks of the boundary: the library loads, exports every symbol include/dbn_b200.h
declares, agrees with the ctypes struct layouts, fails loudly without a device, and its host
Philox matches the NumPy restatement used by the tests.  No compute kernels are called."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gpu_harness as G
from dbn_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "dbn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(dbn_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    lib = G.lib()
    declared = header_symbols()
    assert declared, "no symbols parsed from the header"
    for name in sorted(declared):
        assert hasattr(lib, name), "libdbn_b200.so does not export %s" % name
    # and the binding covers the header exactly
    assert declared == set(L.SYMBOLS.keys())


def test_host_helper_library_exports_its_header():
    """include/dbn_hostrng.h <-> libdbn_hostrng.so (host-side C helper, no CUDA)."""
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "dbn_hostrng.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b(dbn_[a-z0-9_]+)\s*\(", src))
    assert declared == {"dbn_host_legacy_randn"}
    lib = C.CDLL(os.path.join(ROOT, "deep-belief-network_b200", "libdbn_hostrng.so"))
    for name in declared:
        assert hasattr(lib, name)


def test_struct_layouts_match():
    lib = G.lib()
    for i, st in enumerate((L.Mat, L.Rng, L.Operands, L.ActEpilogue, L.UpdateEpilogue, L.RbmStepArgs, L.MlpStepArgs,
                            L.ColSum, L.DpExchange)):
        assert lib.dbn_struct_size(i) == C.sizeof(st), st.__name__


def test_host_philox_matches_numpy():
    lib = G.lib()
    for seed, stream, epoch, row0, rows, cols in [(0, 0, 0, 0, 3, 4), (2 ** 63 + 12345, 7, 3, 1000, 9, 13),
                                                  (0xDEADBEEFCAFE, 1, 99, 2 ** 31, 4, 257)]:
        out = np.zeros((rows, cols), dtype=np.float32)
        assert lib.dbn_philox_uniforms_host(seed, stream, epoch, row0, rows, cols, out.ctypes.data_as(C.c_void_p)) == 0
        ref = G.philox_uniforms(seed, stream, epoch, row0 & 0xFFFFFFFF, rows, cols)
        assert np.array_equal(out, ref)
        assert out.min() >= 0.0 and out.max() < 1.0


def test_philox_is_uniform_enough():
    u = G.philox_uniforms(42, 0, 0, 0, 512, 512).astype(np.float64)
    assert abs(u.mean() - 0.5) < 5e-3
    assert abs(u.var() - 1.0 / 12) < 2e-3
    # rows and columns decorrelated
    assert abs(np.corrcoef(u[:-1].ravel(), u[1:].ravel())[0, 1]) < 1e-2


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    lib = G.lib()
    assert lib.dbn_device_count() == 0
    ops, epi = L.Operands(), L.ActEpilogue()
    rc = lib.dbn_gemm_act(None, L.PREC_FP32, 4, 4, C.byref(ops), 1, C.byref(epi))
    assert rc == L.ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.dbn_last_error()
    import dbn_b200
    with pytest.raises(L.DbnB200Error):
        dbn_b200.BinaryRBM(n_epochs=1, verbose=False).fit(np.zeros((4, 3), dtype=np.float32))


def test_invalid_arguments_are_rejected_before_any_launch():
    lib = G.lib()
    assert lib.dbn_rbm_fit_epoch(None, None, 10, 4) == L.ERR_INVALID_ARG
    assert lib.dbn_philox_uniforms_host(0, 0, 0, 0, 1, 1, None) == L.ERR_INVALID_ARG
