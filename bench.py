#!/usr/bin/env python
"""bench.py -- throughput of the CD-k hot path on B200, one JSON line (contract in the task brief).

Workloads (BASELINE.json `configs`):
  c2  (default at --gpus 1)  784-[500,500,2000] greedy CD-1 pre-training, batch 256, N = 60000
      step = one epoch of every layer of the stack over the synthetic matrix (3 epoch graphs of 235
      mini-batch CD-1 updates each + the 2 inter-layer transforms); value = rows * steps / time.
  c5  (default at --gpus > 1) 4096 x 4096 RBM, CD-1, GLOBAL batch 32768 sharded over the ranks,
      raw-delta all-reduce (NCCL) per step; step = one mini-batch update; strong scaling.
  c3  (--workload c3)        784-[500,500,2000]-10 backprop fine-tune, batch 256, one iteration/step.

`value`   device-timed (CUDA events, barrier + synchronize on both sides, max over ranks), inputs
          resident in HBM.
`e2e`     the same metric through the public estimator API with HOST (pinned) buffers: H2D of the
          inputs and D2H of the trained parameters inside the timed region.
`roofline` dominant kernel (tcgen05 contraction) timed alone, back to back, with CUDA events.
`cpu_baseline` / `--impl reference`: the reference's per-sample float64 NumPy algorithm
          (oracle/dbn_oracle.py per-sample port; the Python reference cannot travel to the GPU box)
          on the host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "RBM CD-1 samples/sec"
C2_DIMS = [784, 500, 500, 2000]
C2_ROWS, C2_BS = 60000, 256
C5_V, C5_H, C5_BS = 4096, 4096, 32768


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"],
                "source": "measured"}
    except Exception:
        return {"hbm": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


# ---------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))
        except Exception:
            pass

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        # samples taken inside the timed region; if it was shorter than the sampling period, the samples
        # of the (identical-load) warm-up that immediately precedes it are used and the window says so
        inside = [r for (t, r) in self.rows if self.t0 is not None and self.t0 - 0.05 <= t <= self.t1 + 0.05]
        window = "timed region"
        if len(inside) < 2:
            inside = [r for (_t, r) in self.rows][-10:]
            window = "warm-up + timed region"
        sm = [float(r[0]) for r in inside if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in inside if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in inside if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": window}


# ---------------------------------------------------------------------------------------------------
# CPU baseline: the reference's per-sample algorithm (float64 NumPy) on a bounded sample
# ---------------------------------------------------------------------------------------------------
def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"]
        return int(max(n)) if n else 1
    except Exception:
        return 1


def cpu_c2_pass(rows, seed=0):
    """`rows` synthetic rows through one CD-1 epoch of each of the three layers, per-sample loops as in
    dbn/models.py:105-146.  Returns seconds."""
    from oracle import dbn_oracle as orc
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    t = 0.0
    for v, h in zip(C2_DIMS[:-1], C2_DIMS[1:]):
        X = (rng.random((rows, v)) < 0.13).astype(np.float32)
        W, b, c = orc.rbm_init(v, h, "sigmoid")
        t0 = time.perf_counter()
        orc.rbm_epoch_per_sample(W, b, c, X, 1, "sigmoid", 0.01, C2_BS)
        t += time.perf_counter() - t0
    return t


def cpu_c5_pass(rows, seed=0):
    from oracle import dbn_oracle as orc
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    X = (rng.random((rows, C5_V)) < 0.5).astype(np.float32)
    W, b, c = orc.rbm_init(C5_V, C5_H, "sigmoid")
    t0 = time.perf_counter()
    orc.rbm_epoch_per_sample(W, b, c, X, 1, "sigmoid", 0.01, C5_BS)
    return time.perf_counter() - t0


def cpu_c3_pass(rows, seed=0):
    from oracle import dbn_oracle as orc
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    X = (rng.random((rows, 784)) < 0.13).astype(np.float32)
    Y = np.eye(10)[rng.integers(0, 10, rows)]
    Ws, cs = [], []
    for v, h in zip(C2_DIMS[:-1], C2_DIMS[1:]):
        W, _b, c = orc.rbm_init(v, h, "sigmoid")
        Ws.append(W); cs.append(c)
    Wo, bo = orc.head_init(10, 2000)
    t0 = time.perf_counter()
    orc.finetune_epoch_per_sample(Ws, cs, Wo, bo, X, Y, "sigmoid", "softmax", 0.1, 1.0, C2_BS, 0.0)
    return time.perf_counter() - t0


CPU_PASS = {"c2": (cpu_c2_pass, 96), "c5": (cpu_c5_pass, 8), "c3": (cpu_c3_pass, 128)}


def cpu_baseline(workload, budget_s=12.0):
    fn, rows = CPU_PASS[workload]
    t = fn(rows)
    reps = 1
    while t * reps < budget_s / 2 and reps < 4:   # a second look if the first pass was short
        t = min(t, fn(rows))
        reps += 1
    return {"value": rows / t, "unit": "samples/s", "cores": cpu_threads(), "kind": "port",
            "sample": "%d rows, one epoch of the per-sample float64 NumPy loop of every layer (best of %d)" % (rows, reps)}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload or ("c2" if args.gpus == 1 else "c5")
    fn, rows = CPU_PASS[wl]
    rows = max(8, rows // 3)
    for _ in range(args.warmup):
        fn(rows)
    t0 = time.perf_counter()
    for i in range(args.steps):
        fn(rows, seed=i)
    dt = time.perf_counter() - t0
    val = rows * args.steps / dt
    out = {"impl": "reference", "metric": METRIC if wl != "c3" else "backprop samples/sec", "value": val,
           "unit": "samples/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
           "higher_is_better": True, "scaling": "weak" if wl != "c5" else "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": workload_config(wl, args.gpus),
           "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cpu_threads(), "kind": "port",
                            "sample": "%d rows per step through the per-sample float64 NumPy loop (oracle port of "
                                      "dbn/models.py:96-146; the Python reference does not travel to the GPU box)" % rows},
           "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def workload_config(wl, gpus):
    if wl == "c2":
        return {"workload": "c2: UnsupervisedDBN 784-[500,500,2000] greedy CD-1 pre-training, batch 256, "
                            "60000 synthetic rows (13% ones), sigmoid, lr 0.01",
                "rows": C2_ROWS, "batch_size": C2_BS, "layers": C2_DIMS, "cd_k": 1,
                "l2_policy": "inputs larger than L2 (282 MB of fp32+bf16 rows per layer-1 epoch)",
                "parallelism": "replicas only" if gpus == 1 else "%d independent replicas" % gpus}
    if wl == "c5":
        return {"workload": "c5: BinaryRBM 4096x4096 CD-1, global batch 32768 sharded data-parallel, sigmoid, lr 0.01",
                "global_batch": C5_BS, "visible": C5_V, "hidden": C5_H, "cd_k": 1,
                "l2_policy": "inputs larger than L2 (268 MB of bf16 rows + 96 MB of weights per step)",
                "parallelism": "dp%d, NCCL sum all-reduce of dW|db|dc in row blocks overlapped with the dW contraction" % gpus}
    return {"workload": "c3: SupervisedDBNClassification 784-[500,500,2000]-10 backprop fine-tune, batch 256, "
                        "60000 synthetic rows, sigmoid, lr 0.1, l2 1.0",
            "rows": C2_ROWS, "batch_size": C2_BS, "layers": C2_DIMS + [10],
            "l2_policy": "inputs larger than L2", "parallelism": "replicas only"}


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default=None, choices=[None, "c2", "c3", "c5"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--global-batch", type=int, default=None, help="override the c5 global batch (experiments)")
    args = ap.parse_args()
    if args.global_batch:
        global C5_BS
        C5_BS = args.global_batch
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from dbn_b200 import _lib as L, engine as E
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = args.workload or ("c2" if world == 1 else "c5")
    lib = E.require_device()
    pk = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, finish=None):
        sampler = ClockSampler(local)
        sampler.start()
        for _ in range(warmup):
            fn()
        if finish is not None:
            finish()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = lib.dbn_launch_count()
        sampler.mark_start()
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()   # the main stream joins outstanding all-reduce / apply work before the stop event
        e1.record()
        barrier()
        sampler.mark_end()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop()
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, clocks, int(lib.dbn_launch_count() - n0)

    rng = np.random.default_rng(0)
    np.random.seed(0)
    extra = {}

    if wl == "c2":
        # ---------------- device-resident stack ----------------
        Xh = torch.empty(C2_ROWS, 784, dtype=torch.float32).pin_memory()
        Xh.numpy()[:] = (rng.random((C2_ROWS, 784)) < 0.13)
        Xd = Xh.to(dev)
        layers, feats, runners = [], [Xd], []
        for v, h in zip(C2_DIMS[:-1], C2_DIMS[1:]):
            W = np.random.randn(h, v) / np.sqrt(v)
            layers.append(E.DeviceRbm(W, np.random.randn(v) / np.sqrt(v), np.random.randn(h) / np.sqrt(v), "sigmoid",
                                      args.precision, dev))
            feats.append(torch.empty(C2_ROWS, h, dtype=torch.float32, device=dev))
        for i, l in enumerate(layers):
            l.transform(feats[i], out=feats[i + 1])
            runners.append(E.RbmEpochRunner(l, feats[i], C2_BS, 1, 0.01, seed=1234 + i))

        def step():
            for i, r in enumerate(runners):
                r.run()
                if i + 1 < len(runners):
                    layers[i].transform(feats[i], out=feats[i + 1])

        ms, clocks, eager = timed(step, args.steps, args.warmup)
        launches = eager + args.steps * sum(r.launches_per_epoch for r in runners) * (1 if E._use_graphs() else 0)
        rows_per_step = C2_ROWS
        flops_per_step = sum(10.0 * v * h for v, h in zip(C2_DIMS[:-1], C2_DIMS[1:])) * C2_ROWS
        scaling = "weak"

        # dominant kernel alone: the v->h contraction + Bernoulli sampling of layer 1 (256 x 500, K = 784), the
        # largest share of the step in the ncu launch list (profiles/r01_c2_splitk_launches.txt)
        l1 = layers[0]
        ws = l1.make_workspace(C2_BS)
        a = l1._step_args(runners[0].Xp, 1, 0.0, ws, 1, None)
        a.B, a.row0 = C2_BS, 0
        reps = 200

        def dom():
            L.check(lib.dbn_rbm_first_hidden(E.stream_ptr(), C.byref(a), 0, l1.H))
        for _ in range(10):
            dom()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if E._use_graphs():
            # back to back inside a CUDA graph, like the epoch it comes from (an eager loop of 5 us kernels would time
            # the host's launch rate instead)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(50):
                    dom()
            g.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps // 50):
                g.replay()
            e1.record()
        else:
            e0.record()
            for _ in range(reps):
                dom()
            e1.record()
        torch.cuda.synchronize()
        dom_ms = e0.elapsed_time(e1) / reps
        dom_flops = 2.0 * C2_BS * l1.H * l1.V
        roof = {"bound": "tensor", "kernel": "umma_gemm_sk_kernel<64, act> (split-K, clusters of 4, 64 CTAs): v->h + sample "
                                             "of layer 1 (256x500, K=784)",
                "achieved": dom_flops / (dom_ms * 1e-3) / 1e12, "peak": pk["bf16_burst"], "unit": "TFLOP/s",
                "peak_source": pk["source"] + " bf16 burst",
                "traffic": 1.3e6, "traffic_source": "ncu dram__bytes_read+write per launch, profiles/r01_c2_splitk_full.txt "
                                                    "(algorithmic: 1.2 MB of operands; the outputs stay in L2)",
                "note": "latency-bound at batch 256: 0.2 GFLOP per launch and a chain of four dependent launches per "
                        "mini-batch; see step_roofline and profiles/r01_notes.md for the phase breakdown"}

        # ---------------- e2e through the estimator API, host buffers ----------------
        def e2e_step():
            from dbn_b200 import UnsupervisedDBN
            m = UnsupervisedDBN(hidden_layers_structure=C2_DIMS[1:], learning_rate_rbm=0.01, n_epochs_rbm=1,
                                batch_size=C2_BS, verbose=False, precision=args.precision)
            m.fit(Xh.numpy())
            return m
        h2d = Xh.numel() * 4
        d2h = sum((v * h + v + h) * 4 for v, h in zip(C2_DIMS[:-1], C2_DIMS[1:]))
    elif wl == "c5":
        steps_per_epoch = 4
        n_rows = C5_BS * steps_per_epoch
        gen = torch.Generator(device="cpu").manual_seed(0)
        Xh = torch.empty(n_rows, C5_V, dtype=torch.float32).pin_memory()
        Xh.copy_((torch.rand(n_rows, C5_V, generator=gen) < 0.5).float())
        Xd = Xh.to(dev)
        W = (torch.randn(C5_H, C5_V, generator=gen) / np.sqrt(C5_V)).numpy()
        layer = E.DeviceRbm(W, np.zeros(C5_V), np.zeros(C5_H), "sigmoid", args.precision, dev)
        order = np.arange(n_rows)
        if world > 1:
            runner = E.DpRbmEpochRunner(layer, Xd, C5_BS, 1, 0.01, 99, (rank, world))
            local_idx = order.reshape(runner.steps, world, runner.bl)[:, rank, :].reshape(-1)
            runner.feeder.push(local_idx)
            layer.stage(Xd, idx=runner.feeder.dev, out=runner.Xp)
            state = {"s": 0}

            def step():
                runner.step(state["s"] % runner.steps)
                state["s"] += 1
            drain = runner.drain
        else:
            runner = E.RbmEpochRunner(layer, Xd, C5_BS, 1, 0.01, 99)
            layer.stage(Xd, out=runner.Xp)
            a = runner.args
            state = {"s": 0}
            drain = None

            def step():
                a.B = C5_BS
                a.row0 = (state["s"] % steps_per_epoch) * C5_BS
                a.rng.row0 = a.row0
                L.check(lib.dbn_rbm_cd_step(E.stream_ptr(), C.byref(a)))
                state["s"] += 1
        ms, clocks, launches = timed(step, args.steps, args.warmup, finish=drain)
        rows_per_step = C5_BS
        flops_per_step = 10.0 * C5_V * C5_H * C5_BS
        scaling = "strong"
        # dominant kernel alone: v -> h contraction on this rank's rows
        bl = C5_BS // world
        Xop = runner.Xp
        out = torch.empty(bl, E.bf16_pitch(C5_H), dtype=torch.bfloat16, device=dev) if args.precision == "bf16" else \
            torch.empty(bl, C5_H, dtype=torch.float32, device=dev)

        def dom():
            layer.hidden_means(Xop, 0, bl, out)
        for _ in range(3):
            dom()
        torch.cuda.synchronize()
        reps = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            dom()
        e1.record()
        torch.cuda.synchronize()
        dom_ms = e0.elapsed_time(e1) / reps
        dom_flops = 2.0 * bl * C5_V * C5_H
        roof = {"bound": "tensor", "kernel": "umma_gemm2_kernel<256, act> v->h (%dx4096, K=4096), cta_group::2" % bl,
                "achieved": dom_flops / (dom_ms * 1e-3) / 1e12, "peak": pk["bf16_burst"], "unit": "TFLOP/s",
                "peak_source": pk["source"] + " bf16 burst (kernel timed alone)",
                "traffic": 835e6 * bl / 32768.0,
                "traffic_source": "ncu dram__bytes_read+write per 32768-row launch, profiles/r01_c5_umma2cta_full.txt"}

        # the same workload on ONE GPU of this box (rank 0 alone, after the timed region), so that the
        # N > 1 line carries its own scaling denominator: the N = 1 default command benches config 2
        if world > 1:
            barrier()
            if rank == 0:
                l1 = E.DeviceRbm(W, np.zeros(C5_V), np.zeros(C5_H), "sigmoid", args.precision, dev)
                r1 = E.RbmEpochRunner(l1, Xd, C5_BS, 1, 0.01, 99)
                l1.stage(Xd, out=r1.Xp)
                a1 = r1.args

                def step1(i):
                    a1.B = C5_BS
                    a1.row0 = (i % steps_per_epoch) * C5_BS
                    a1.rng.row0 = a1.row0
                    L.check(lib.dbn_rbm_cd_step(E.stream_ptr(), C.byref(a1)))
                for i in range(3):
                    step1(i)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(10):
                    step1(i)
                e1.record()
                torch.cuda.synchronize()
                t1 = e0.elapsed_time(e1) / 10
                extra["same_workload_1gpu"] = {"value": C5_BS / (t1 * 1e-3), "unit": "samples/s", "ms_per_step": t1,
                                               "note": "rank 0 alone, full 32768-row batch, same box, 10 steps"}
                del l1, r1
                torch.cuda.empty_cache()
            barrier()

        def e2e_step():
            from dbn_b200 import BinaryRBM
            os.environ["DBN_B200_DATA_PARALLEL"] = "1"
            np.random.seed(5)
            m = BinaryRBM(n_hidden_units=C5_H, learning_rate=0.01, n_epochs=1, batch_size=C5_BS, verbose=False,
                          precision=args.precision)
            m.fit(Xh.numpy())
            return m
        h2d = Xh.numel() * 4
        d2h = (C5_V * C5_H + C5_V + C5_H) * 4
        rows_per_e2e = n_rows
    else:  # c3
        Xh = torch.empty(C2_ROWS, 784, dtype=torch.float32).pin_memory()
        Xh.numpy()[:] = (rng.random((C2_ROWS, 784)) < 0.13)
        y = rng.integers(0, 10, C2_ROWS)
        Yh = np.eye(10, dtype=np.float32)[y]
        Xd, Yd = Xh.to(dev), torch.from_numpy(Yh).to(dev)
        layers = []
        for v, h in zip(C2_DIMS[:-1], C2_DIMS[1:]):
            layers.append(E.DeviceRbm(np.random.randn(h, v) / np.sqrt(v), np.zeros(v), np.random.randn(h) / np.sqrt(v),
                                      "sigmoid", args.precision, dev))
        mlp = E.DeviceMlp(layers, np.random.randn(10, 2000) / np.sqrt(2000), np.zeros(10), "sigmoid", "softmax",
                          args.precision, dev)
        runner = E.MlpIterRunner(mlp, Xd, Yd, C2_BS, 0.1, 1.0, 0.0, 77)

        def step():
            runner.run()
        ms, clocks, eager = timed(step, args.steps, args.warmup)
        launches = eager + args.steps * runner.launches_per_iter * (1 if E._use_graphs() else 0)
        rows_per_step = C2_ROWS
        P = [v * h for v, h in zip(C2_DIMS[:-1], C2_DIMS[1:])] + [2000 * 10]
        flops_per_step = (2.0 * sum(P) + 2.0 * sum(P[1:]) + 2.0 * sum(P)) * C2_ROWS
        scaling = "weak"
        roof = {"bound": "tensor", "kernel": "whole step (see step_roofline)", "achieved": None,
                "peak": pk["bf16_burst"], "unit": "TFLOP/s", "traffic": None}
        dom_ms = None

        def e2e_step():
            from dbn_b200 import SupervisedDBNClassification
            m = SupervisedDBNClassification(hidden_layers_structure=C2_DIMS[1:], learning_rate=0.1, n_epochs_rbm=0,
                                            n_iter_backprop=1, batch_size=C2_BS, verbose=False,
                                            precision=args.precision)
            m.fit(Xh.numpy(), y)
            return m
        h2d = Xh.numel() * 4 + C2_ROWS * 8
        d2h = sum((v * h + h) * 4 for v, h in zip(C2_DIMS[:-1], C2_DIMS[1:])) + 2000 * 10 * 4

    value = rows_per_step * args.steps / (ms * 1e-3)
    if roof.get("achieved") is not None:
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["kernel_ms"] = dom_ms
    step_tf = flops_per_step * args.steps / (ms * 1e-3) / 1e12 / world
    extra["step_roofline"] = {"bound": "tensor", "achieved": step_tf, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                              "frac": step_tf / pk["bf16_sustained"], "per_gpu": True,
                              "algorithmic_flops_per_step": flops_per_step}

    # ---- e2e: public API, host buffers, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, 5))
        e2e_step()  # warm-up (module load, allocator)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            m = e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        rows = rows_per_e2e if wl == "c5" else rows_per_step
        e2e = {"value": rows * e2e_steps / dt, "unit": "samples/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
               "api": "estimator.fit(X_host) with one epoch per call; H2D of X and D2H of W,b,c per call"}

    if rank == 0:
        out = {"metric": METRIC if wl != "c3" else "backprop samples/sec", "value": value, "unit": "samples/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
               "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
               "dtype": "bf16" if args.precision == "bf16" else "f32",
               "data": "synthetic", "config": dict(workload_config(wl, world), arithmetic=(
                   "bf16 operands on tcgen05, fp32 accumulation in TMEM, fp32 master weights"
                   if args.precision == "bf16" else "strict fp32 FFMA")), "clocks": clocks, "e2e": e2e,
               "gpu_launches": int(launches), "roofline": roof}
        out.update(extra)
        if not args.no_cpu and world == 1:
            out["cpu_baseline"] = cpu_baseline(wl)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
