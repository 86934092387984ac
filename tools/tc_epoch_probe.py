"""Step time and in-kernel phase trace of the persistent tcgen05 epoch kernel (csrc/rbm_tc_epoch.cuh) against the
tiled four-launch path, on the three layer shapes of config 2.

    python tools/tc_epoch_probe.py [--rows 8192] [--prec fp16]

Per shape: us per mini-batch of both paths (CUDA events around whole epochs) and, from the kernel's debug trace
(dbn_debug_set_trace), the cycles a CTA spends per phase in: waiting for operands after the previous barrier, the
MMAs, the epilogue, and the grid barrier (averaged over CTAs with a tile, steps 2..7)."""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_harness as G  # noqa: E402
from dbn_b200 import _lib as L, engine as E  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=8192)
    ap.add_argument("--prec", default="fp16")
    ap.add_argument("--bs", type=int, default=256)
    ap.add_argument("--shapes", default="784x500,500x500,500x2000")
    args = ap.parse_args()
    lib = G.lib()
    rng = np.random.default_rng(0)
    for shp in args.shapes.split(","):
        V, H = [int(x) for x in shp.split("x")]
        n, bs = args.rows, args.bs
        plan = G.tc_epoch_plan(V, H, bs)
        print("== %d -> %d, batch %d, %d rows, %s   plan %s" % (V, H, bs, n, args.prec, plan))
        W = rng.standard_normal((H, V)) * 0.01
        layer = E.DeviceRbm(W, np.zeros(V), np.zeros(H), "sigmoid", args.prec, torch.device("cuda"))
        Xop = G.stage_operand((rng.random((n, V)) < 0.13).astype(np.float32), args.prec)
        ws = layer.make_workspace(bs)
        ctr = torch.zeros(1, dtype=torch.int32, device="cuda")
        a = layer._step_args(Xop, 1, 0.01 / bs, ws, 3, ctr, None)
        steps = (n + bs - 1) // bs
        for tc in (1, 0):
            lib.dbn_set_tc_epoch(tc)
            for _ in range(2):
                L.check(lib.dbn_rbm_fit_epoch(E.stream_ptr(), C.byref(a), n, bs))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                L.check(lib.dbn_rbm_fit_epoch(E.stream_ptr(), C.byref(a), n, bs))
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (reps * steps)
            print("   %-28s %7.2f us per mini-batch  (%.2f M samples/s)" % ("persistent kernel" if tc else "tiled path (eager launches)", us, bs / us))
        lib.dbn_set_tc_epoch(1)
        if not plan:
            continue
        ctas = plan["ctas"]
        tr = torch.zeros(ctas * 8 * 128, dtype=torch.int64, device="cuda")
        lib.dbn_debug_set_trace(C.c_void_p(tr.data_ptr()))
        L.check(lib.dbn_rbm_fit_epoch(E.stream_ptr(), C.byref(a), n, bs))
        torch.cuda.synchronize()
        lib.dbn_debug_set_trace(None)
        t = tr.cpu().numpy().reshape(ctas, 8, 8, 16).astype(np.float64)
        names = ["v->h + sample", "h->v", "v->h (k)", "dW + update"]
        tiles = [plan["hid_tiles"], plan["vis_tiles"], plan["hid_tiles"], plan["dw_tiles"]]
        print("   cycles, mean over the CTAs with a tile, steps 2..7; columns: loads issued after the barrier | operands landed | MMAs issued |"
              " MMAs complete | (prefetch done) | accumulator seen by the epilogue | first chunk in registers | first chunk stored |"
              " epilogue done | CTA joined | release + arrive | barrier passed (mean / min over CTAs)")
        tot = 0.0
        for p in range(4):
            sel = t[:tiles[p], 2:8, p]
            prev_bar = t[:tiles[p], 2:8, p - 1, 5] if p > 0 else t[:tiles[p], 1:7, 3, 5]
            d = lambda x, y: float((x - y).mean())
            cols = [d(sel[..., 11], prev_bar) if p > 0 else float("nan"), d(sel[..., 0], prev_bar), d(sel[..., 10], sel[..., 0]), d(sel[..., 6], sel[..., 10]),
                    d(sel[..., 7], prev_bar), d(sel[..., 1], sel[..., 6]), d(sel[..., 8], sel[..., 1]), d(sel[..., 9], sel[..., 8]),
                    d(sel[..., 2], sel[..., 9]), d(sel[..., 3], sel[..., 2]), d(sel[..., 4], sel[..., 3])]
            bar = sel[..., 5] - sel[..., 4]
            tot += float((sel[..., 5] - prev_bar).mean())
            print("   %-14s %4d | %5.0f | %5.0f | %5.0f | %5.0f | (%5.0f) | %5.0f | %5.0f | %5.0f | %5.0f | %5.0f | %5.0f | %5.0f / %.0f" %
                  tuple([names[p], tiles[p]] + cols + [float(bar.mean()), float(bar.min(axis=0).mean())]))
        print("   sum %.0f cycles = %.2f us at 1.965 GHz" % (tot, tot / 1965.0))


if __name__ == "__main__":
    main()
