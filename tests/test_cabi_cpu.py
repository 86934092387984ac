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
    assert declared == {"dbn_host_legacy_randn", "dbn_host_legacy_randn_f32", "dbn_host_stage_rows"}
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


def test_persistent_kernel_plans_are_pure_host_functions():
    """The tile plans of the two persistent tcgen05 kernels (csrc/rbm_tc_epoch.cuh, csrc/mlp_tc_epoch.cuh) are computed on
    the host: every phase fits the 148 SMs of a B200 and 227 KB of shared memory; shapes outside the plan report 0."""
    import ctypes as C
    lib = L.load()
    out = (C.c_int32 * 16)()
    for (V, H, B) in [(784, 500, 256), (500, 500, 256), (500, 2000, 256), (256, 256, 32), (64, 128, 16)]:
        m = lib.dbn_rbm_tc_epoch_plan_info(V, H, B, out, 16)
        assert m >= 16, (V, H, B)
        ctas, hid_tiles, vis_tiles, dw_tiles, smem = out[0], out[3], out[8], out[10], out[11]
        assert 0 < ctas <= 148 and max(hid_tiles, vis_tiles, dw_tiles) == ctas and 0 < smem <= 232448
        assert out[1] in (32, 64) and out[2] % 8 == 0 and out[5] % 8 == 0 and out[9] % 8 == 0 and out[9] <= 128
        assert lib.dbn_rbm_tc_epoch_plan(L.PREC_BF16, V, H, B) == ctas and lib.dbn_rbm_tc_epoch_plan(L.PREC_FP32, V, H, B) == 0
    assert lib.dbn_rbm_tc_epoch_plan_info(4096, 4096, 256, out, 16) == 0      # config 5: CTA-pair kernels
    assert lib.dbn_rbm_tc_epoch_plan_info(784, 500, 512, out, 16) == 0        # mini-batch beyond one tile of the dW contraction
    assert lib.dbn_rbm_tc_epoch_plan_info(13, 100, 16, out, 16) == 0          # config 4: FP32 cluster kernel
    dims = (C.c_int32 * 3)(500, 500, 2000)
    assert 0 < lib.dbn_mlp_tc_epoch_plan(L.PREC_BF16, 784, dims, 3, 10, 256) <= 148
    assert lib.dbn_mlp_tc_epoch_plan(L.PREC_FP32, 784, dims, 3, 10, 256) == 0
    assert lib.dbn_mlp_tc_epoch_plan(L.PREC_BF16, 784, dims, 3, 100, 256) == 0   # wide head: tiled path
    five = (C.c_int32 * 5)(128, 128, 128, 128, 128)
    assert lib.dbn_mlp_tc_epoch_plan(L.PREC_BF16, 784, five, 5, 10, 256) == 0     # more than 4 hidden layers
    prev = lib.dbn_set_tc_finetune(0)
    assert lib.dbn_mlp_tc_epoch_plan(L.PREC_BF16, 784, dims, 3, 10, 256) == 0
    lib.dbn_set_tc_finetune(prev)


def test_scheduling_switches_are_host_state():
    """dbn_set_tc_epoch / dbn_set_tc_finetune / dbn_set_tail_split flip process-wide host flags and return the previous
    value (tests and A/B runs rely on restoring them); none of them needs a device."""
    lib = L.load()
    for name in ("dbn_set_tc_epoch", "dbn_set_tc_finetune", "dbn_set_tail_split"):
        fn = getattr(lib, name)
        prev = fn(0)
        assert prev in (0, 1)
        assert fn(1) == 0 and fn(prev) == 1
        assert fn(prev) == prev
    assert lib.dbn_tail_split_count() >= 0
