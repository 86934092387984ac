"""Phase-latency probe of the tcgen05 contraction kernel on the config-2 shapes (debug tool).
Prints, per kernel of one CD-1 step, the median over CTAs of the cycles spent in each phase."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import gpu_harness as G
from dbn_b200 import _lib as L, engine as E

lib = G.lib()
dev = torch.device("cuda")
names = ["setup", "first operands", "mainloop issue", "acc ready", "epilogue", "teardown", "total"]


def probe(label, fn, n_cta_hint=148):
    buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    buf2 = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    lib.dbn_debug_set_trace(C.c_void_p(buf.data_ptr()))
    fn()
    lib.dbn_debug_set_trace(C.c_void_p(buf2.data_ptr()))
    fn()          # dependent launch right behind the first: measures the inter-kernel gap
    torch.cuda.synchronize()
    lib.dbn_debug_set_trace(None)
    t = buf.cpu().numpy().reshape(148, 16)
    t2 = buf2.cpu().numpy().reshape(148, 16)
    t = t[t[:, 7] != 0]
    t2 = t2[t2[:, 7] != 0]
    if len(t) == 0 or len(t2) == 0:
        print("%-34s (kernel variant without trace marks)" % label)
        return
    a_exit = t[:, 9].max()
    gap_entry = t2[:, 0].min() - a_exit
    gap_wait = t2[:, 8].min() - a_exit
    dur_a = a_exit - t[:, 0].min()
    d = np.stack([t[:, 2] - t[:, 1], t[:, 3] - t[:, 2], t[:, 4] - t[:, 3], t[:, 5] - t[:, 4], t[:, 6] - t[:, 5],
                  t[:, 7] - t[:, 6], t[:, 7] - t[:, 1]], axis=1)
    med = np.median(d, axis=0)
    span_ns = t[:, 0].max() - t[:, 0].min()
    extra = ""
    if np.any(t[:, 10] != 0):   # split-K kernel: slot 10 = partial tiles exchanged (after the cluster barrier)
        md = lambda a, b: np.median(t[:, a] - t[:, b])
        extra = ("  [split-K: tmem->staging=%d, fence+bar+copy issue=%d, wait partials=%d, cluster arrive=%d, sum=%d, "
                 "epilogue math+stores=%d]" % (md(12, 5), md(13, 12), md(14, 13), md(10, 14), md(11, 10), md(6, 11)))
    print("%-34s ctas=%3d  " % (label, len(t)) + "  ".join("%s=%d" % (n, m) for n, m in zip(names, med)) + extra +
          "  | entry skew %d ns | kernel %d ns, next kernel entered %+d ns, passed its dependency wait %+d ns after this one's exit"
          % (span_ns, dur_a, gap_entry, gap_wait))


rng = np.random.default_rng(0)
for (V, H) in [(784, 500), (500, 500), (500, 2000)]:
    B = 256
    W = rng.standard_normal((H, V)) / np.sqrt(V)
    layer = E.DeviceRbm(W, np.zeros(V), np.zeros(H), "sigmoid", "bf16", dev)
    X = (rng.random((B, V)) < 0.13).astype(np.float32)
    Xop = G.stage_operand(X, "bf16")
    ws = layer.make_workspace(B)
    a = layer._step_args(Xop, 1, 0.01 / 256, ws, 1, None)
    a.B, a.row0 = B, 0
    L.check(lib.dbn_rbm_cd_step(E.stream_ptr(), C.byref(a)))
    hm = torch.zeros(B, E.bf16_pitch(H), dtype=torch.bfloat16, device=dev)
    vm = torch.zeros(B, E.bf16_pitch(V), dtype=torch.bfloat16, device=dev)
    raw = torch.zeros(H + 1, (V + 1 + 3) // 4 * 4, dtype=torch.float32, device=dev)
    a.raw_delta = E.mat(raw)
    probe("v->h %dx%d K=%d" % (B, H, V), lambda: layer.hidden_means(Xop, 0, B, hm))
    probe("h->v %dx%d K=%d" % (B, V, H), lambda: layer.visible_means(hm, B, vm))
    probe("dW %dx%d K=2x%d" % (H + 1, V + 1, B), lambda: L.check(lib.dbn_rbm_delta_rows(E.stream_ptr(), C.byref(a), 0, H)))

    # wall time of the same launches back to back (no trace)
    for label, fn in [("v->h", lambda: layer.hidden_means(Xop, 0, B, hm)), ("h->v", lambda: layer.visible_means(hm, B, vm)),
                      ("dW", lambda: L.check(lib.dbn_rbm_delta_rows(E.stream_ptr(), C.byref(a), 0, H))),
                      ("cd_step", lambda: L.check(lib.dbn_rbm_cd_step(E.stream_ptr(), C.byref(a))))]:
        a.raw_delta = E.mat(raw) if label != "cd_step" else E.mat(None)
        g = torch.cuda.CUDAGraph()
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for _ in range(50):
                fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        print("   graph of 50 x %-8s: %.2f us per call" % (label, e0.elapsed_time(e1) * 1e3 / 200))
