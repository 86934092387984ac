"""Device time of single contractions at given shapes (tile-plan experiments):
    python tools/shape_bench.py            # the per-rank shapes of config 5 at 1 / 2 / 4 / 8 GPUs
Env: DBN_B200_2CTA_BN=128|256 forces the CTA-pair tile width."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dbn_b200 import _lib as L, engine as E  # noqa: E402

lib = E.require_device()
dev = torch.device("cuda")
V = H = 4096


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for rows in (32768, 16384, 8192, 4096):
    dt = torch.float16
    X = (torch.rand(rows, E.bf16_pitch(V), device=dev) < 0.5).to(dt)
    W = (torch.randn(H, E.bf16_pitch(V), device=dev) / 64).to(dt)
    out = torch.empty(rows, E.bf16_pitch(H), dtype=dt, device=dev)
    c = torch.zeros(H, device=dev)
    # G1: rows x H, K = V
    ops = L.Operands()
    ops.A, ops.B, ops.K = E.mat(X), E.mat(W), V
    epi = L.ActEpilogue()
    epi.bias, epi.act = c.data_ptr(), L.ACT_SIGMOID
    epi.rng.keep_p = 1.0
    epi.out = E.mat(out)
    t1 = timeit(lambda: L.check(lib.dbn_gemm_act(E.stream_ptr(), L.PREC_F16, rows, H, C.byref(ops), 1, C.byref(epi))))
    # G4: H x V, K = 2 x rows (raw output)
    raw = torch.empty(H, V, device=dev)
    o2 = (L.Operands * 2)()
    for o in o2:
        o.A, o.B, o.transA, o.transB, o.K = E.mat(out), E.mat(X), 1, 1, rows
    u = L.UpdateEpilogue()
    u.decay, u.scale, u.raw = 1.0, 1.0, E.mat(raw)
    t4 = timeit(lambda: L.check(lib.dbn_gemm_update(E.stream_ptr(), L.PREC_F16, H, V, 0, 0, o2, 2, C.byref(u))))
    print("rows %6d  G1 %.4f ms (%.0f TF)   G4 %.4f ms (%.0f TF)   bn=%s" %
          (rows, t1, 2.0 * rows * V * H / t1 / 1e9, t4, 4.0 * rows * V * H / t4 / 1e9, os.environ.get("DBN_B200_2CTA_BN", "auto")), flush=True)
