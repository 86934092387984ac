#!/usr/bin/env python
"""bench.py -- throughput of the CD-k / backprop hot path on B200, one JSON line (contract in the task brief).

The line's headline workload is the SAME at every GPU count, so the driver's 1 -> 8 curve is one workload:

  c5  4096 x 4096 RBM, CD-1, GLOBAL batch 32768 (BASELINE.json config 5).  step = one mini-batch update.
      N = 1: the whole batch on one GPU.  N > 1: rows sharded over the ranks, dW reduce-scattered from the
      contraction's epilogue over NVLink peer memory, sharded update, bf16 rows written into every rank's shadow
      (dp_peer.DpPeerRunner; NCCL collectives only if CUDA IPC is unavailable).  Strong scaling.

At N = 1 the line additionally carries `sub`-records for the other configurations of BASELINE.json, each with its own
value / e2e / roofline / cpu_baseline (the metric is "RBM CD-1 samples/sec AND backprop samples/sec"):

  c2  784-[500,500,2000] greedy CD-1 pre-training, batch 256, 60000 rows; step = one epoch of every layer
  c3  784-[500,500,2000]-10 backprop fine-tune, batch 256; step = one iteration over the 60000 rows
  c1  README example 64-[256,256]-10 (digits-shaped, relu, dropout 0.2, batch 32) through the estimator API
  c4  13-[100]-1 regression, CD-10, relu, batch 16 through the estimator API

`value`    device-timed (CUDA events, barrier + synchronize on both sides, max over ranks), inputs resident in HBM.
`e2e`      the same metric through the public estimator API with HOST (pinned) buffers: H2D of the inputs and D2H
           of the trained parameters inside the timed region.
`roofline` dominant kernel (tcgen05 contraction) timed alone with CUDA events.
`cpu_baseline` / `--impl reference`: the reference's own NumPy implementation (oracle/_ref = the unmodified package,
           installed by __graft_entry__.build(); the oracle's per-sample port if that is absent) on the host cores,
           on a bounded sample of the same workload.
`--workload c2|c3` makes one of the sub-workloads the headline instead (N = 1 only)."""
import os
import sys

if "--impl" in sys.argv and "reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is entitled to every host core
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.pop(_k, None)

import argparse
import json
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "RBM CD-1 samples/sec"
C2_DIMS = [784, 500, 500, 2000]
C2_ROWS, C2_BS = 60000, 256
C5_V, C5_H, C5_BS = 4096, 4096, 32768
C5_STEPS_PER_EPOCH = 4


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
    """SM clock and throttle reasons of this rank's GPU sampled every 5 ms through NVML (in-process, so that even a
    30 ms timed region holds several samples); `nvidia-smi -lms 100` is the fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None
        self._stop_flag = False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _nvml_row(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        act = lambda bit: "Active" if (r & bit) else "Not Active"
        return [str(sm), str(self.max_sm), "0",
                act(n.nvmlClocksThrottleReasonHwSlowdown), act(n.nvmlClocksThrottleReasonHwThermalSlowdown),
                act(n.nvmlClocksThrottleReasonSwThermalSlowdown), act(n.nvmlClocksThrottleReasonSwPowerCap)]

    def run(self):
        if self.nvml is not None:
            while not self._stop_flag:
                try:
                    self.rows.append((time.perf_counter(), self._nvml_row()))
                except Exception:
                    break
                time.sleep(0.005)
            return
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
        self._stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        # samples taken inside the timed region; if it was shorter than the sampling period, the samples
        # of the (identical-load) warm-up that immediately precedes it are used and the window says so
        inside = [r for (t, r) in self.rows if self.t0 is not None and self.t0 <= t <= self.t1]
        window = "timed region"
        if len(inside) < 2:
            inside = [r for (_t, r) in self.rows][-10:]
            window = "warm-up + timed region"
        sm = [float(r[0]) for r in inside if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in inside if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in inside if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": window,
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference's own NumPy backend on a bounded sample of each workload
# ---------------------------------------------------------------------------------------------------
def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"]
        return int(max(n)) if n else 1
    except Exception:
        return 1


_REF = None


def reference_backend():
    """The unmodified reference package if __graft_entry__.build() installed it under oracle/_ref (it travels to the
    GPU box with the snapshot), else None -> the oracle's per-sample port."""
    global _REF
    if _REF is None:
        _REF = False
        path = os.path.join(ROOT, "oracle", "_ref")
        if os.path.isdir(os.path.join(path, "dbn")):
            sys.path.insert(0, path)
            try:
                import dbn as ref  # noqa: F401
                _REF = ref
            except Exception:
                _REF = False
            finally:
                sys.path.remove(path)
    return _REF or None


def cpu_kind():
    return "reference" if reference_backend() is not None else "port"


def _rbm_pass(v, h, X, bs, k, act, lr):
    """One CD-k epoch of one layer on the CPU; returns seconds."""
    ref = reference_backend()
    if ref is not None:
        m = ref.BinaryRBM(n_hidden_units=h, activation_function=act, learning_rate=lr, n_epochs=1,
                          contrastive_divergence_iter=k, batch_size=bs, verbose=False)
        t0 = time.perf_counter()
        m.fit(X)
        return time.perf_counter() - t0
    from oracle import dbn_oracle as orc
    W, b, c = orc.rbm_init(v, h, act)
    t0 = time.perf_counter()
    orc.rbm_epoch_per_sample(W, b, c, X, k, act, lr, bs)
    return time.perf_counter() - t0


def cpu_c2_pass(rows, seed=0):
    """`rows` synthetic rows through one CD-1 epoch of each of the three layers (dbn/models.py:105-146)."""
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    t = 0.0
    for v, h in zip(C2_DIMS[:-1], C2_DIMS[1:]):
        X = (rng.random((rows, v)) < 0.13).astype(np.float32)
        t += _rbm_pass(v, h, X, C2_BS, 1, "sigmoid", 0.01)
    return t


def cpu_c5_pass(rows, seed=0):
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    X = (rng.random((rows, C5_V)) < 0.5).astype(np.float32)
    return _rbm_pass(C5_V, C5_H, X, C5_BS, 1, "sigmoid", 0.01)


def _finetune_pass(dims, n_out, X, y, bs, act, lr, dropout, clf):
    ref = reference_backend()
    if ref is not None:
        cls = ref.SupervisedDBNClassification if clf else ref.SupervisedDBNRegression
        m = cls(hidden_layers_structure=list(dims[1:]), learning_rate=lr, n_epochs_rbm=0, n_iter_backprop=1,
                batch_size=bs, activation_function=act, dropout_p=dropout, verbose=False)
        m.pre_train(X)                 # 0 epochs: initialises the weights only (dbn/models.py:105)
        t0 = time.perf_counter()
        m.fit(X, y, pre_train=False)
        return time.perf_counter() - t0
    from oracle import dbn_oracle as orc
    Y = np.eye(n_out)[y] if clf else np.asarray(y, dtype=np.float64).reshape(len(y), -1)
    Ws, cs = [], []
    for v, h in zip(dims[:-1], dims[1:]):
        W, _b, c = orc.rbm_init(v, h, act)
        Ws.append(W); cs.append(c)
    Wo, bo = orc.head_init(n_out, dims[-1])
    t0 = time.perf_counter()
    orc.finetune_epoch_per_sample(Ws, cs, Wo, bo, X, Y, act, "softmax" if clf else "linear", lr, 1.0, bs, dropout)
    return time.perf_counter() - t0


def cpu_c3_pass(rows, seed=0):
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    X = (rng.random((rows, 784)) < 0.13).astype(np.float32)
    y = rng.integers(0, 10, rows)
    y[:10] = np.arange(10)
    return _finetune_pass(C2_DIMS, 10, X, y, C2_BS, "sigmoid", 0.1, 0.0, True)


CPU_PASS = {"c2": (cpu_c2_pass, 96), "c5": (cpu_c5_pass, 8), "c3": (cpu_c3_pass, 128)}


def cpu_baseline(workload, budget_s=12.0):
    fn, rows = CPU_PASS[workload]
    t = fn(rows)
    reps = 1
    while t * reps < budget_s / 2 and reps < 4:   # a second look if the first pass was short
        t = min(t, fn(rows))
        reps += 1
    return {"value": rows / t, "unit": "samples/s", "cores": cpu_threads(), "kind": cpu_kind(),
            "sample": "%d rows, one epoch of the reference's per-sample float64 NumPy loop (best of %d)" % (rows, reps)}


def workload_config(wl, gpus):
    if wl == "c2":
        return {"workload": "c2: UnsupervisedDBN 784-[500,500,2000] greedy CD-1 pre-training, batch 256, "
                            "60000 synthetic rows (13% ones), sigmoid, lr 0.01",
                "rows": C2_ROWS, "batch_size": C2_BS, "layers": C2_DIMS, "cd_k": 1,
                "l2_policy": "inputs larger than L2 (282 MB of fp32+bf16 rows per layer-1 epoch)",
                "parallelism": "replicas only"}
    if wl == "c5":
        return {"workload": "c5: BinaryRBM 4096x4096 CD-1, global batch 32768 sharded data-parallel, sigmoid, lr 0.01",
                "global_batch": C5_BS, "visible": C5_V, "hidden": C5_H, "cd_k": 1,
                "l2_policy": "inputs larger than L2 (268 MB of bf16 rows + 96 MB of weights per step)",
                "parallelism": "dp%d" % gpus}
    return {"workload": "c3: SupervisedDBNClassification 784-[500,500,2000]-10 backprop fine-tune, batch 256, "
                        "60000 synthetic rows, sigmoid, lr 0.1, l2 1.0",
            "rows": C2_ROWS, "batch_size": C2_BS, "layers": C2_DIMS + [10],
            "l2_policy": "inputs larger than L2", "parallelism": "replicas only"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation, each step a bounded sample of the workload.  Under
    torchrun rank 0 alone works; the other ranks exit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload or "c5"
    fn, rows = CPU_PASS[wl]
    rows = max(8, rows // 3)
    for _ in range(args.warmup):
        fn(rows)
    t0 = time.perf_counter()
    for i in range(args.steps):
        fn(rows, seed=i)
    dt = time.perf_counter() - t0
    val = rows * args.steps / dt
    kind = cpu_kind()
    what = ("the unmodified reference package (oracle/_ref, `dbn` NumPy backend)" if kind == "reference" else
            "oracle port of dbn/models.py:96-146 (the reference package was not installed under oracle/_ref)")
    out = {"impl": "reference", "metric": METRIC if wl != "c3" else "backprop samples/sec", "value": val,
           "unit": "samples/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
           "higher_is_better": True, "scaling": "strong" if wl == "c5" else "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": workload_config(wl, args.gpus),
           "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cpu_threads(), "kind": kind,
                            "sample": "%d rows per step through the per-sample float64 NumPy loop: %s" % (rows, what)},
           "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------------
class Ctx:
    pass


def make_ctx(args):
    import torch
    import torch.distributed as dist
    from dbn_b200 import engine as E
    c = Ctx()
    c.args = args
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = int(os.environ.get("RANK", "0"))
    c.local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(c.local)
    c.dev = torch.device("cuda", c.local)
    if c.world > 1:
        dist.init_process_group("nccl", device_id=c.dev)
    c.lib = E.require_device()
    c.pk = peaks()
    c.dist = dist
    return c


def barrier(c):
    import torch
    if c.world > 1:
        c.dist.barrier()
    torch.cuda.synchronize()


def timed(c, fn, steps, warmup, finish=None):
    import torch
    sampler = ClockSampler(c.local)
    sampler.start()
    for _ in range(warmup):
        fn()
    if finish is not None:
        finish()
    barrier(c)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = c.lib.dbn_launch_count()
    sampler.mark_start()
    e0.record()
    for _ in range(steps):
        fn()
    if finish is not None:
        finish()   # the main stream joins the outstanding exchange before the stop event
    e1.record()
    barrier(c)
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    if c.world > 1:
        # every rank's own time and median SM clock travel with the line: a straggler (a GPU held at low clocks, a
        # starved host thread) shows up here instead of silently setting the max
        mine = torch.tensor([ms, float(clocks.get("sm_mhz") or 0.0)], device=c.dev)
        every = [torch.zeros_like(mine) for _ in range(c.world)]
        c.dist.all_gather(every, mine)
        clocks["ms_per_rank"] = [round(float(t[0].item()), 3) for t in every]
        clocks["sm_mhz_per_rank"] = [float(t[1].item()) for t in every]
        ms = max(float(t[0].item()) for t in every)
    return ms, clocks, int(c.lib.dbn_launch_count() - n0)


def time_alone(fn, reps, graph_batch=0):
    """Average device time of fn() launched back to back (inside a CUDA graph when the kernel is a few microseconds:
    an eager loop would time the host's launch rate)."""
    import torch
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if graph_batch:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(graph_batch):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(max(1, reps // graph_batch)):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (max(1, reps // graph_batch) * graph_batch)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def e2e_measure(c, e2e_step, rows, steps, h2d, d2h, api):
    import torch
    for _ in range(2):
        e2e_step()  # warm-up (module load, allocator, pinned staging pool: the second fit still grows the pools)
    barrier(c)
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    barrier(c)
    dt = time.perf_counter() - t0
    if c.world > 1:
        t = torch.tensor([dt], device=c.dev)
        c.dist.all_reduce(t, op=c.dist.ReduceOp.MAX)
        dt = float(t.item())
    return {"value": rows * steps / dt, "unit": "samples/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "steps": steps, "api": api}


def step_roofline(c, flops_per_step, steps, ms):
    tf = flops_per_step * steps / (ms * 1e-3) / 1e12 / c.world
    return {"bound": "tensor", "achieved": tf, "peak": c.pk["bf16_sustained"], "unit": "TFLOP/s",
            "frac": tf / c.pk["bf16_sustained"], "per_gpu": True, "algorithmic_flops_per_step": flops_per_step,
            "peak_source": c.pk["source"] + " bf16 sustained (whole step, launch gaps included)"}


# ---------------------------------------------------------------------------------------------------
def pick_precision(args, wl):
    if args.precision != "auto":
        return args.precision
    return "bf16" if wl == "c3" else "fp16"


# ncu dram__bytes_read.sum + dram__bytes_write.sum of one rbm_tc_epoch_kernel launch per layer (profiles/r02_c2_tc_epoch_full.txt)
C2_TC_TRAFFIC = {0: 108.5e6, 1: 69.1e6, 2: 75.1e6}   # algorithmic: 99.8 / 61.4 / 61.4 MB of input rows + 2.4 / 1.5 / 6.0 MB of W in and out


def bench_c2(c, steps, warmup, with_e2e=True):
    import ctypes as C
    import torch
    from dbn_b200 import _lib as L, engine as E
    args, dev, lib = c.args, c.dev, c.lib
    precision = pick_precision(args, "c2")
    rng = np.random.default_rng(0)
    np.random.seed(0)
    Xh = torch.empty(C2_ROWS, 784, dtype=torch.float32).pin_memory()
    Xh.numpy()[:] = (rng.random((C2_ROWS, 784)) < 0.13)
    Xd = Xh.to(dev)
    layers, feats, runners = [], [Xd], []
    for v, h in zip(C2_DIMS[:-1], C2_DIMS[1:]):
        W = np.random.randn(h, v) / np.sqrt(v)
        layers.append(E.DeviceRbm(W, np.random.randn(v) / np.sqrt(v), np.random.randn(h) / np.sqrt(v), "sigmoid",
                                  precision, dev))
        feats.append(torch.empty(C2_ROWS, h, dtype=torch.float32, device=dev))
    for i, l in enumerate(layers):
        l.transform(feats[i], out=feats[i + 1])
        runners.append(E.RbmEpochRunner(l, feats[i], C2_BS, 1, 0.01, seed=1234 + i))

    def step():
        for i, r in enumerate(runners):
            r.run()
            if i + 1 < len(runners):
                layers[i].transform(feats[i], out=feats[i + 1])

    ms, clocks, eager = timed(c, step, steps, warmup)
    launches = eager + steps * sum(r.launches_per_epoch for r in runners if not (r.tc_epoch or r.persistent)) * (1 if E._use_graphs() else 0)
    flops_per_step = sum(10.0 * v * h for v, h in zip(C2_DIMS[:-1], C2_DIMS[1:])) * C2_ROWS
    if all(r.tc_epoch for r in runners):
        # dominant kernel: rbm_tc_epoch_kernel, ONE cooperative launch per layer-epoch (csrc/rbm_tc_epoch.cuh).  Each layer's
        # launch is timed alone with CUDA events on the launching stream; the longest one is reported.
        per_layer = []
        for i, r in enumerate(runners):
            def one(r=r):
                L.check(lib.dbn_rbm_fit_epoch(E.stream_ptr(), C.byref(r.args), r.n, r.bs))
            per_layer.append(time_alone(one, 5))
        i_dom = int(np.argmax(per_layer))
        dom_ms = per_layer[i_dom]
        v_d, h_d = C2_DIMS[i_dom], C2_DIMS[i_dom + 1]
        dom_tf = 10.0 * v_d * h_d * C2_ROWS / (dom_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "rbm_tc_epoch_kernel: persistent tcgen05 CD-1 epoch of layer %d (%d -> %d, %d mini-batches of "
                                             "%d rows in one cooperative launch, grid barrier between phases)"
                                             % (i_dom + 1, v_d, h_d, -(-C2_ROWS // C2_BS), C2_BS),
                "achieved": dom_tf, "peak": c.pk["bf16_burst"], "unit": "TFLOP/s", "frac": dom_tf / c.pk["bf16_burst"],
                "kernel_ms": dom_ms, "kernel_ms_per_layer": per_layer, "us_per_minibatch": [1e3 * t / -(-C2_ROWS // C2_BS) for t in per_layer],
                "peak_source": c.pk["source"] + " bf16 burst (kernel timed alone)",
                "algorithmic_flops_per_launch": 10.0 * v_d * h_d * C2_ROWS,
                "traffic": C2_TC_TRAFFIC.get(i_dom), "traffic_source": "ncu dram__bytes_read+write of this launch, profiles/r02_c2_tc_epoch_full.txt "
                                                                     "(algorithmic: the epoch's input rows once + W in and out; P0 / samples / V1 / Pk stay in L2)",
                "note": "latency-bound by construction at batch 256: per mini-batch four dependent phases of < 0.3 us tensor time each, "
                        "separated by grid barriers (~1.3 us) and operand refills out of L2 (~0.8 us); phase trace in profiles/r02_notes.md"}
    else:
        # dominant kernel alone: the v->h contraction + Bernoulli sampling of layer 1 (256 x 500, K = 784), the
        # largest share of the step in the ncu launch list (profiles/)
        l1 = layers[0]
        ws = l1.make_workspace(C2_BS)
        a = l1._step_args(runners[0].Xp, 1, 0.0, ws, 1, None)
        a.B, a.row0 = C2_BS, 0

        def dom():
            L.check(lib.dbn_rbm_first_hidden(E.stream_ptr(), C.byref(a), 0, l1.H))
        dom_ms = time_alone(dom, 200, graph_batch=50 if E._use_graphs() else 0)
        dom_tf = 2.0 * C2_BS * l1.H * l1.V / (dom_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "umma_gemm_sk_kernel<64, act> (split-K, clusters of 4, 64 CTAs): v->h + sample "
                                             "of layer 1 (256x500, K=784)",
                "achieved": dom_tf, "peak": c.pk["bf16_burst"], "unit": "TFLOP/s", "frac": dom_tf / c.pk["bf16_burst"],
                "kernel_ms": dom_ms, "peak_source": c.pk["source"] + " bf16 burst",
                "traffic": 1.3e6, "traffic_source": "ncu dram__bytes_read+write per launch, profiles/r01_c2_splitk_full.txt "
                                                    "(algorithmic: 1.2 MB of operands; the outputs stay in L2)",
                "note": "latency-bound at batch 256: 0.2 GFLOP per launch and a chain of four dependent launches per "
                        "mini-batch; see step_roofline and profiles/ for the phase breakdown"}
    rec = {"metric": METRIC, "value": C2_ROWS * steps / (ms * 1e-3), "unit": "samples/s", "steps": steps, "warmup": warmup,
           "ms_per_step": ms / steps, "config": workload_config("c2", c.world), "clocks": clocks,
           "gpu_launches": int(launches), "roofline": roof, "step_roofline": step_roofline(c, flops_per_step, steps, ms)}
    if with_e2e:
        def e2e_step():
            from dbn_b200 import UnsupervisedDBN
            m = UnsupervisedDBN(hidden_layers_structure=C2_DIMS[1:], learning_rate_rbm=0.01, n_epochs_rbm=1,
                                batch_size=C2_BS, verbose=False, precision=args.precision)
            m.fit(Xh.numpy())
            return m
        d2h = sum((v * h + v + h) * 4 for v, h in zip(C2_DIMS[:-1], C2_DIMS[1:]))
        rec["e2e"] = e2e_measure(c, e2e_step, C2_ROWS, 10, Xh.numel() * 4, d2h,
                                 "UnsupervisedDBN.fit(X_host), one epoch per layer per call; H2D of X and D2H of W,b,c per call")
    return rec


C3_TC_TRAFFIC = 122.5e6   # (algorithmic: 99.8 MB of staged input rows + 2.4 MB of targets + ~13 MB of weights in and out)  -- ncu dram__bytes_read.sum + dram__bytes_write.sum of one mlp_tc_epoch_kernel launch (profiles/r02_c3_tc_epoch_full.txt)


def bench_c3(c, steps, warmup, with_e2e=True):
    import ctypes as C
    import torch
    from dbn_b200 import _lib as L, engine as E
    args, dev, lib = c.args, c.dev, c.lib
    precision = pick_precision(args, "c3")
    rng = np.random.default_rng(0)
    np.random.seed(0)
    Xh = torch.empty(C2_ROWS, 784, dtype=torch.float32).pin_memory()
    Xh.numpy()[:] = (rng.random((C2_ROWS, 784)) < 0.13)
    y = rng.integers(0, 10, C2_ROWS)
    Yh = np.eye(10, dtype=np.float32)[y]
    Xd, Yd = Xh.to(dev), torch.from_numpy(Yh).to(dev)
    layers = []
    for v, h in zip(C2_DIMS[:-1], C2_DIMS[1:]):
        layers.append(E.DeviceRbm(np.random.randn(h, v) / np.sqrt(v), np.zeros(v), np.random.randn(h) / np.sqrt(v),
                                  "sigmoid", precision, dev))
    mlp = E.DeviceMlp(layers, np.random.randn(10, 2000) / np.sqrt(2000), np.zeros(10), "sigmoid", "softmax",
                      precision, dev)
    runner = E.MlpIterRunner(mlp, Xd, Yd, C2_BS, 0.1, 1.0, 0.0, 77)
    ms, clocks, eager = timed(c, runner.run, steps, warmup)
    P = [v * h for v, h in zip(C2_DIMS[:-1], C2_DIMS[1:])] + [2000 * 10]
    flops_per_step = (2.0 * sum(P) + 2.0 * sum(P[1:]) + 2.0 * sum(P)) * C2_ROWS
    launches = eager + steps * runner.launches_per_iter * (1 if (E._use_graphs() and not runner.tc_epoch) else 0)
    if runner.tc_epoch:
        # dominant kernel: mlp_tc_epoch_kernel, ONE cooperative launch per iteration (csrc/mlp_tc_epoch.cuh), timed alone
        def one():
            L.check(lib.dbn_mlp_fit_epoch(E.stream_ptr(), C.byref(runner.args), runner.n, runner.bs))
        dom_ms = time_alone(one, 5)
        dom_tf = flops_per_step / (dom_ms * 1e-3) / 1e12
        n_mb = -(-C2_ROWS // C2_BS)
        roof = {"bound": "tensor", "kernel": "mlp_tc_epoch_kernel: persistent tcgen05 back-propagation iteration (784-[500,500,2000]-10, %d mini-batches "
                                             "of %d rows in one cooperative launch; 7 phases per mini-batch, grid barrier between phases)" % (n_mb, C2_BS),
                "achieved": dom_tf, "peak": c.pk["bf16_burst"], "unit": "TFLOP/s", "frac": dom_tf / c.pk["bf16_burst"],
                "kernel_ms": dom_ms, "us_per_minibatch": 1e3 * dom_ms / n_mb, "peak_source": c.pk["source"] + " bf16 burst (kernel timed alone)",
                "algorithmic_flops_per_launch": flops_per_step, "traffic": C3_TC_TRAFFIC,
                "traffic_source": "ncu dram__bytes_read+write of this launch, profiles/r02_c3_tc_epoch_full.txt (algorithmic: the iteration's input "
                                  "rows once + every weight matrix in and out; activations and deltas stay in L2)",
                "note": "latency-bound by construction at batch 256: seven dependent phases per mini-batch separated by grid barriers"}
    else:
        # dominant kernel alone: the weight-gradient contraction + decayed SGD update of the widest layer
        # (2000 x 501 accumulator, K = 256 rows), timed back to back inside a CUDA graph
        l3 = layers[2]
        Ad = torch.zeros(C2_BS, E.bf16_pitch(2000), dtype=E.OP_DTYPE[E.PREC_IDS[precision]], device=dev)
        Bd = torch.zeros(C2_BS, E.bf16_pitch(500), dtype=E.OP_DTYPE[E.PREC_IDS[precision]], device=dev)
        Ad[:, :2000] = 0.01
        Bd[:, :501] = 1.0
        ops = L.Operands()
        ops.A, ops.B, ops.transA, ops.transB, ops.K = E.mat(Ad), E.mat(Bd), 1, 1, C2_BS
        u = L.UpdateEpilogue()
        scratchW = torch.zeros_like(l3.W)
        scratchc = torch.zeros_like(l3.c)
        scratchWb = torch.zeros_like(l3.Wb) if l3.Wb is not None else None
        u.W, u.ldw, u.Wb = scratchW.data_ptr(), 500, E.mat(scratchWb)
        u.vecM, u.decay, u.scale = scratchc.data_ptr(), 0.999, -1e-4
        prec = E.PREC_IDS[precision]

        def dom():
            L.check(lib.dbn_gemm_update(E.stream_ptr(), prec, 2000, 500, 0, 1, C.byref(ops), 1, C.byref(u)))
        dom_ms = time_alone(dom, 200, graph_batch=50 if E._use_graphs() else 0)
        dom_tf = 2.0 * C2_BS * 2000 * 501 / (dom_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "weight-gradient contraction + decayed SGD update of layer 3 (2000x501, K=256)",
                "achieved": dom_tf, "peak": c.pk["bf16_burst"], "unit": "TFLOP/s", "frac": dom_tf / c.pk["bf16_burst"],
                "kernel_ms": dom_ms, "peak_source": c.pk["source"] + " bf16 burst", "traffic": None,
                "note": "latency-bound at batch 256 (0.5 GFLOP per launch); step_roofline is the figure of merit"}
    rec = {"metric": "backprop samples/sec", "value": C2_ROWS * steps / (ms * 1e-3), "unit": "samples/s", "steps": steps,
           "warmup": warmup, "ms_per_step": ms / steps, "config": workload_config("c3", c.world), "clocks": clocks,
           "gpu_launches": int(launches), "launches_per_minibatch": runner.launches_per_iter / max(1, -(-C2_ROWS // C2_BS)),
           "roofline": roof, "step_roofline": step_roofline(c, flops_per_step, steps, ms)}
    if with_e2e:
        def e2e_step():
            from dbn_b200 import SupervisedDBNClassification
            m = SupervisedDBNClassification(hidden_layers_structure=C2_DIMS[1:], learning_rate=0.1, n_epochs_rbm=0,
                                            n_iter_backprop=1, batch_size=C2_BS, verbose=False,
                                            precision=args.precision)
            m.fit(Xh.numpy(), y)
            return m
        d2h = sum((v * h + h) * 4 for v, h in zip(C2_DIMS[:-1], C2_DIMS[1:])) + 2000 * 10 * 4
        rec["e2e"] = e2e_measure(c, e2e_step, C2_ROWS, 10, Xh.numel() * 4 + C2_ROWS * 8, d2h,
                                 "SupervisedDBNClassification.fit(X_host, y) with n_epochs_rbm=0, one back-prop iteration per call")
    return rec


def bench_c5(c, steps, warmup, with_e2e=True):
    import ctypes as C
    import torch
    from dbn_b200 import _lib as L, dp_peer, engine as E
    args, dev, lib, world, rank = c.args, c.dev, c.lib, c.world, c.rank
    precision = pick_precision(args, "c5")
    n_rows = C5_BS * C5_STEPS_PER_EPOCH
    gen = torch.Generator(device="cpu").manual_seed(0)
    W = (torch.randn(C5_H, C5_V, generator=gen) / np.sqrt(C5_V)).numpy()
    layer = E.DeviceRbm(W, np.zeros(C5_V), np.zeros(C5_H), "sigmoid", precision, dev)
    order = np.arange(n_rows)
    group, exchange = None, "none (one GPU)"
    nvlink = None
    if world > 1:
        nl = n_rows // world
        Xl = (torch.rand(nl, C5_V, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + rank)) < 0.5).float()
        use_peer = dp_peer.peer_path_wanted(layer, world)
        if use_peer:
            ok = torch.ones(1, dtype=torch.int32, device=dev)
            try:
                group = dp_peer.RealPeerGroup(rank, world, dp_peer.ArenaLayout(world, C5_V, C5_H, nl, E.OP_DTYPE[layer.prec]), dev)
            except Exception:
                group = None
                ok.zero_()
            c.dist.all_reduce(ok, op=c.dist.ReduceOp.MIN)
            use_peer = int(ok.item()) == 1
            if not use_peer and group is not None:
                group.close()
                group = None
        if use_peer:
            runner = dp_peer.DpPeerRunner(layer, Xl, n_rows, C5_BS, 1, 0.01, 99, group.member)
            runner.gather(order)
            exchange = ("dW reduce-scattered from the contraction's epilogue into the owners' receive slots over NVLink peer "
                        "memory, sharded update, bf16 rows stored into every rank's shadow (no NCCL on the data path)")
            hs, pitch = C5_H // world, E.bf16_pitch(C5_V)
            nvlink = {"tx_bytes_per_gpu_per_step": int((world - 1) * (hs * C5_V * 4 + hs * pitch * 2 + (C5_H + C5_V) * 8)),
                      "algorithmic": "fp32 dW rows of the %d other owners + this owner's bf16 rows to %d peers + statistics; "
                                     "every element is stored exactly once" % (world - 1, world - 1)}
        else:
            # NCCL fallback: every rank needs the whole dataset for its device-side gather
            Xd = torch.cat([(torch.rand(nl, C5_V, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + r)) < 0.5).float()
                            for r in range(world)])
            runner = E.DpRbmEpochRunner(layer, Xd, C5_BS, 1, 0.01, 99, (rank, world))
            local_idx = order.reshape(runner.steps, world, runner.bl)[:, rank, :].reshape(-1)
            runner.feeder.push(local_idx)
            layer.stage(Xd, idx=runner.feeder.dev, out=runner.Xp)
            exchange = "NCCL reduce-scatter of the fp32 delta, sharded update, NCCL all-gather of the bf16 shadow (CUDA IPC unavailable)"
        state = {"s": 0}

        def step():
            runner.step(state["s"] % runner.steps)
            state["s"] += 1
        drain = runner.drain
        Xop = runner.Xp
    else:
        Xd = (torch.rand(n_rows, C5_V, device=dev, generator=torch.Generator(device=dev).manual_seed(100)) < 0.5).float()
        runner = E.RbmEpochRunner(layer, Xd, C5_BS, 1, 0.01, 99)
        layer.stage(Xd, out=runner.Xp)
        a = runner.args
        state = {"s": 0}
        drain = None
        Xop = runner.Xp

        def step():
            a.B = C5_BS
            a.row0 = (state["s"] % C5_STEPS_PER_EPOCH) * C5_BS
            a.rng.row0 = a.row0
            L.check(lib.dbn_rbm_cd_step(E.stream_ptr(), C.byref(a)))
            state["s"] += 1
    ms, clocks, launches = timed(c, step, steps, warmup, finish=drain)
    flops_per_step = 10.0 * C5_V * C5_H * C5_BS
    # dominant kernel alone: v -> h contraction (+ bias, sigmoid, store) on this rank's rows
    bl = C5_BS // world
    out = torch.empty(bl, E.bf16_pitch(C5_H), dtype=E.OP_DTYPE[layer.prec], device=dev) if layer.prec in E.TC_PRECS else \
        torch.empty(bl, C5_H, dtype=torch.float32, device=dev)

    def dom():
        layer.hidden_means(Xop, 0, bl, out)
    dom_ms = time_alone(dom, 10)
    dom_tf = 2.0 * bl * C5_V * C5_H / (dom_ms * 1e-3) / 1e12
    roof = {"bound": "tensor", "kernel": "umma_gemm2_kernel<256, act> v->h (%dx4096, K=4096), cta_group::2" % bl,
            "achieved": dom_tf, "peak": c.pk["bf16_burst"], "unit": "TFLOP/s", "frac": dom_tf / c.pk["bf16_burst"],
            "kernel_ms": dom_ms, "peak_source": c.pk["source"] + " bf16 burst (kernel timed alone)",
            "traffic": 901e6 * bl / 32768.0,
            "traffic_source": "ncu dram__bytes_read+write per 32768-row launch, profiles/r01_c5_umma2cta_fat_full.txt "
                              "(algorithmic 570 MB: A once, W, bf16 output)"}
    rec = {"metric": METRIC, "value": C5_BS * steps / (ms * 1e-3), "unit": "samples/s", "steps": steps, "warmup": warmup,
           "ms_per_step": ms / steps, "config": workload_config("c5", world), "exchange": exchange, "clocks": clocks,
           "gpu_launches": int(launches), "roofline": roof, "step_roofline": step_roofline(c, flops_per_step, steps, ms)}
    if nvlink is not None:
        nvlink["tx_gbs_per_gpu"] = nvlink["tx_bytes_per_gpu_per_step"] / (ms / steps * 1e-3) / 1e9
        rec["nvlink"] = nvlink
    if group is not None:
        layer.Wb = layer.Wb.clone()
        del runner
        group.close()
    if with_e2e:
        del Xop
        torch.cuda.empty_cache()
        gen = torch.Generator(device="cpu").manual_seed(0)       # the same host dataset on every rank
        Xh = torch.empty(n_rows, C5_V, dtype=torch.float32).pin_memory()
        for s in range(0, n_rows, 16384):
            Xh[s:s + 16384] = (torch.rand(16384, C5_V, generator=gen) < 0.5).float()

        def e2e_step():
            from dbn_b200 import BinaryRBM
            os.environ["DBN_B200_DATA_PARALLEL"] = "1"
            np.random.seed(5)
            m = BinaryRBM(n_hidden_units=C5_H, learning_rate=0.01, n_epochs=1, batch_size=C5_BS, verbose=False,
                          precision=args.precision)
            m.fit(Xh.numpy())
            return m
        h2d = Xh.numel() * 4 // world
        rec["e2e"] = e2e_measure(c, e2e_step, n_rows, 5, h2d, (C5_V * C5_H + C5_V + C5_H) * 4,
                                 "BinaryRBM.fit(X_host), one epoch of %d rows per call; each rank uploads its 1/%d of X, "
                                 "every rank exports W,b,c" % (n_rows, world))
    return rec


def bench_small(c, name):
    """Configs 1 and 4 (BASELINE.json): too small for any roofline -- the whole fit is a few hundred microsecond-scale
    launches (strict FP32, on-chip persistent RBM kernel) -- reported through the estimator API: device-side epoch
    rate of the first RBM layer, and end-to-end fit rates, beside the reference on the host."""
    import torch
    from dbn_b200 import engine as E
    import dbn_b200
    rng = np.random.default_rng(0)
    if name == "c1":
        N, V, dims, bs, k, act, lr_rbm, lr, drop, n_out = 1437, 64, [256, 256], 32, 1, "relu", 0.05, 0.1, 0.2, 10
        X = (rng.integers(0, 17, (N, V)) / 16.0).astype(np.float32)
        y = rng.integers(0, 10, N); y[:10] = np.arange(10)
        desc = "c1: README example 64-[256,256]-10 on digits-shaped data (1437 rows in [0,1]), relu, CD-1, batch 32, dropout 0.2"
    else:
        N, V, dims, bs, k, act, lr_rbm, lr, drop, n_out = 4096, 13, [100], 16, 10, "relu", 0.01, 0.01, 0.0, 1
        X = rng.random((N, V)).astype(np.float32)
        y = (X @ rng.standard_normal(V) + 0.1 * rng.standard_normal(N)).astype(np.float64)
        desc = "c4: SupervisedDBNRegression 13-[100]-1 on synthetic min-max-scaled rows (4096), relu, CD-10, batch 16"
    dev = c.dev
    # device-timed: epochs of the first RBM layer, back to back
    layer = E.DeviceRbm(rng.uniform(-0.2, 0.2, (dims[0], V)) / np.sqrt(V), np.zeros(V), np.zeros(dims[0]), act, "fp32", dev)
    r = E.RbmEpochRunner(layer, torch.from_numpy(X).to(dev), bs, k, lr_rbm, 1)
    reps = 20
    n0 = c.lib.dbn_launch_count()
    ms = time_alone(r.run, reps) * 1.0
    launches = int(c.lib.dbn_launch_count() - n0)
    flops = (2 * k + 3) * 2.0 * V * dims[0]
    rec = {"config": {"workload": desc}, "rbm": {"metric": "RBM CD-%d samples/sec (layer 1, device-timed epochs)" % k,
                                                 "value": N / (ms * 1e-3), "unit": "samples/s", "ms_per_epoch": ms,
                                                 "kernel": "rbm_persistent_kernel (one cluster launch per epoch, W resident in shared memory)" if r.persistent else "tiled path",
                                                 "gpu_launches": launches,
                                                 "achieved_gflops": flops * N / (ms * 1e-3) / 1e9}}
    # end to end through the estimators (host buffers in, host parameters out)
    cls = dbn_b200.SupervisedDBNClassification if name == "c1" else dbn_b200.SupervisedDBNRegression

    def fit(ep_rbm, it_bp):
        np.random.seed(1)
        m = cls(hidden_layers_structure=dims, learning_rate_rbm=lr_rbm, learning_rate=lr, n_epochs_rbm=ep_rbm,
                n_iter_backprop=it_bp, batch_size=bs, activation_function=act, dropout_p=drop,
                contrastive_divergence_iter=k, verbose=False)
        t0 = time.perf_counter()
        m.fit(X, y)
        torch.cuda.synchronize()
        return time.perf_counter() - t0
    fit(1, 1)
    t_pre = min(fit(10, 0) for _ in range(2))
    t_ft = min(fit(0, 10) for _ in range(2))
    rec["e2e"] = {"pretrain_samples_per_s": N * 10 * len(dims) / t_pre, "backprop_samples_per_s": N * 10 / t_ft,
                  "unit": "samples/s (rows x epochs x layers / wall time of estimator.fit with host arrays)",
                  "h2d_bytes_per_step": int(X.nbytes), "d2h_bytes_per_step": int(sum(a * b for a, b in zip([V] + dims, dims + [n_out])) * 4)}
    # the reference on the host: one RBM epoch of layer 1, one back-prop iteration
    np.random.seed(0)
    rows = min(N, 1437 if name == "c1" else 2048)
    t_rbm = _rbm_pass(V, dims[0], X[:rows], bs, k, act, lr_rbm)
    t_bp = _finetune_pass([V] + dims, n_out, X[:rows], y[:rows], bs, act, lr, drop, name == "c1")
    rec["cpu_baseline"] = {"rbm_value": rows / t_rbm, "backprop_value": rows / t_bp, "unit": "samples/s", "cores": cpu_threads(),
                           "kind": cpu_kind(), "sample": "%d rows: one CD-%d epoch of layer 1, one back-prop iteration" % (rows, k)}
    return rec


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default=None, choices=[None, "c2", "c3", "c5"])
    ap.add_argument("--precision", default="auto", choices=["auto", "bf16", "fp16", "fp32"],
                    help="auto = what the estimators pick: fp16 operands for the sigmoid RBM workloads (c2, c5), bf16 for the fine-tune (c3)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="headline workload only (no c1-c4 sub-records at N = 1)")
    ap.add_argument("--global-batch", type=int, default=None, help="override the c5 global batch (experiments)")
    args = ap.parse_args()
    if args.global_batch:
        global C5_BS
        C5_BS = args.global_batch
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    c = make_ctx(args)
    wl = args.workload or "c5"
    if c.world > 1 and wl != "c5":
        raise SystemExit("only c5 shards over GPUs (c1-c4 are 'replicas only': SURVEY 8e)")
    fn = {"c2": bench_c2, "c3": bench_c3, "c5": bench_c5}[wl]
    rec = fn(c, args.steps, args.warmup, with_e2e=not args.no_e2e)
    out = {"metric": rec.pop("metric"), "value": rec.pop("value"), "unit": rec.pop("unit"), "n_gpus": c.world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec.pop("ms_per_step"), "higher_is_better": True,
           "scaling": "strong" if wl == "c5" else "weak", "vs_baseline": None,
           "dtype": {"fp32": "f32"}.get(pick_precision(args, wl), pick_precision(args, wl)), "data": "synthetic",
           "config": rec.pop("config"),
           "arithmetic": ("%s operands on tcgen05, fp32 accumulation in TMEM, fp32 master weights" % pick_precision(args, wl)
                          if pick_precision(args, wl) != "fp32" else "strict fp32 FFMA")}
    rec.pop("steps", None); rec.pop("warmup", None)
    out.update(rec)
    if c.rank == 0:
        if not args.no_cpu and c.world == 1:
            out["cpu_baseline"] = cpu_baseline(wl)
        if c.world == 1 and not args.no_sub and wl == "c5":
            sub = {}
            for name, f, st in (("c2", bench_c2, 10), ("c3", bench_c3, 10)):
                r = f(c, st, 3, with_e2e=not args.no_e2e)
                if not args.no_cpu:
                    r["cpu_baseline"] = cpu_baseline(name)
                sub[name] = r
            for name in ("c1", "c4"):
                sub[name] = bench_small(c, name)
            out["sub"] = sub
        print(json.dumps(out))
    if c.world > 1:
        c.dist.destroy_process_group()


if __name__ == "__main__":
    main()
