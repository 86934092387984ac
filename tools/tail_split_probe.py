"""CTA-pair contraction at the per-rank size of config 5 on 8 GPUs (4096 rows): time of the three chain contractions and
of the dW contraction with the tail split on and off (single GPU; CUDA events around 20 launches each)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import dbn_b200
from dbn_b200 import _lib as L, engine as E
lib = L.load()
dev = torch.device("cuda")
V = H = 4096


def operand(rows, cols):
    t = torch.zeros(rows, E.bf16_pitch(cols), dtype=torch.float16, device=dev)
    t[:, :cols] = (torch.rand(rows, cols, device=dev) < 0.5).to(torch.float16)
    return t


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for rows in (4096, 8192, 32768):
    W = operand(H, V)
    X = operand(rows, V)
    P = operand(rows, H)
    out = torch.zeros(rows, E.bf16_pitch(H), dtype=torch.float16, device=dev)
    cvec = torch.zeros(H, device=dev)
    raw = torch.zeros(H + 1, V + 8, device=dev)
    o1 = L.Operands(); o1.A, o1.B, o1.transA, o1.transB, o1.K = E.mat(X), E.mat(W), 0, 0, V
    e = L.ActEpilogue(); e.bias = cvec.data_ptr(); e.act = E.ACT_IDS["sigmoid"]; e.out = E.mat(out)
    ops = (L.Operands * 2)()
    for o in ops:
        o.A, o.B, o.transA, o.transB, o.K = E.mat(P), E.mat(X), 1, 1, rows
    u = L.UpdateEpilogue(); u.decay, u.scale, u.raw = 1.0, 1.0, E.mat(raw)
    for split in (0, 1, 0, 1):
        lib.dbn_set_tail_split(split)
        c0 = lib.dbn_tail_split_count()
        t_act = timeit(lambda: L.check(lib.dbn_gemm_act(E.stream_ptr(), L.PREC_F16, rows, H, C.byref(o1), 1, C.byref(e))))
        c1 = lib.dbn_tail_split_count()
        t_upd = timeit(lambda: L.check(lib.dbn_gemm_update(E.stream_ptr(), L.PREC_F16, H, V, 0, 0, ops, 2, C.byref(u))))
        c2 = lib.dbn_tail_split_count()
        fl_act, fl_upd = 2.0 * rows * H * V, 2.0 * H * V * 2 * rows
        print("rows %5d  tail split %d: v->h %.1f us (%.0f TFLOP/s, split launches %d)   dW %.1f us (%.0f TFLOP/s, split launches %d)"
              % (rows, split, t_act, fl_act / t_act / 1e6, c1 - c0, t_upd, fl_upd / t_upd / 1e6, c2 - c1), flush=True)
lib.dbn_set_tail_split(1)
