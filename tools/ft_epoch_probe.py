"""Step time and in-kernel phase trace of the persistent fine-tune kernel (csrc/mlp_tc_epoch.cuh) on config 3
(784-[500,500,2000]-10, batch 256) against the tiled path.      python tools/ft_epoch_probe.py [--rows 8192]"""
import argparse, ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_harness as G
from dbn_b200 import _lib as L, engine as E

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=8192)
ap.add_argument("--prec", default="bf16")
args = ap.parse_args()
lib = G.lib()
dev = torch.device("cuda")
rng = np.random.default_rng(0)
dims = [784, 500, 500, 2000]
n, bs = args.rows, 256
layers = [E.DeviceRbm(rng.standard_normal((h, v)) / np.sqrt(v), np.zeros(v), rng.standard_normal(h) * 0.1, "sigmoid", args.prec, dev)
          for v, h in zip(dims[:-1], dims[1:])]
mlp = E.DeviceMlp(layers, rng.standard_normal((10, 2000)) / np.sqrt(2000), np.zeros(10), "sigmoid", "softmax", args.prec, dev)
Xop = G.stage_operand((rng.random((n, 784)) < 0.13).astype(np.float32), args.prec)
Yd = torch.from_numpy(np.eye(10, dtype=np.float32)[rng.integers(0, 10, n)]).to(dev)
ws = mlp.make_workspace(bs)
loss = torch.zeros(n, device=dev)
a = mlp.step_args(Xop, Yd, 0.1 / bs, 1.0 - 0.1 / n, 1.0, ws, 0, None, loss, None)
steps = (n + bs - 1) // bs
for tc in (1, 0):
    lib.dbn_set_tc_finetune(tc)
    for _ in range(2):
        L.check(lib.dbn_mlp_fit_epoch(E.stream_ptr(), C.byref(a), n, bs))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        L.check(lib.dbn_mlp_fit_epoch(E.stream_ptr(), C.byref(a), n, bs))
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * steps)
    print("%-30s %7.2f us per mini-batch  (%.2f M samples/s)" % ("persistent kernel" if tc else "tiled path (eager launches)", us, bs / us))
lib.dbn_set_tc_finetune(1)
dd = (C.c_int32 * 3)(500, 500, 2000)
ctas = lib.dbn_mlp_tc_epoch_plan(1, 784, dd, 3, 10, bs)
NP = 13
tr = torch.zeros(ctas * 8 * NP * 8, dtype=torch.int64, device=dev)
lib.dbn_debug_set_trace(C.c_void_p(tr.data_ptr()))
L.check(lib.dbn_mlp_fit_epoch(E.stream_ptr(), C.byref(a), n, bs))
torch.cuda.synchronize()
lib.dbn_debug_set_trace(None)
t = tr.cpu().numpy().reshape(ctas, 8, NP, 8).astype(np.float64)
names = ["F0 784->500", "F1 500->500", "F2 500->2000", "HEAD", "B1 + U_head", "B0 + U2", "U1 + U0"]
print("phase           operands   ->acc(1st)  ->acc(2nd)  ->epi done   cta-join   release   barrier (mean / min)   total   (cycles, CTAs with stamps, steps 2..7)")
tot = 0.0
for p in range(7):
    sel = t[:, 2:8, p]
    prev = t[:, 2:8, p - 1, 5] if p > 0 else t[:, 1:7, 6, 5]
    ok = (sel[..., 5] > 0) & (prev > 0)
    def m(x, y, need=None):
        msk = ok & (x > 0) & (y > 0)
        return float((x - y)[msk].mean()) if msk.any() else float("nan")
    ops, acc1 = m(sel[..., 0], prev), m(sel[..., 1], sel[..., 0])
    acc2 = m(sel[..., 6], sel[..., 1])
    last_acc = np.where(sel[..., 6] > 0, sel[..., 6], sel[..., 1])
    epi = m(sel[..., 2], last_acc)
    join, rel = m(sel[..., 3], sel[..., 2]), m(sel[..., 4], sel[..., 3])
    bar = (sel[..., 5] - sel[..., 4])[ok]
    total = float((sel[..., 5] - prev)[ok].mean())
    tot += total
    print("%-14s %9.0f %11.0f %11.0f %11.0f %10.0f %9.0f %10.0f / %-8.0f %7.0f" % (names[p], ops, acc1, acc2, epi, join, rel, bar.mean(), bar.min(), total))
print("sum %.0f cycles = %.2f us at 1.965 GHz" % (tot, tot / 1965.0))
