"""Where the wall time of one end-to-end data-parallel `BinaryRBM.fit(X_host)` (bench.py's e2e leg of config 5 at
N > 1) goes on rank 0: host-side time inside each stage, measured without adding synchronisations (debug tool).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/e2e_dp_breakdown.py
"""
import collections
import functools
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import dbn_b200  # noqa: E402
from dbn_b200 import dp_peer, engine as E, models as M  # noqa: E402

acc = collections.OrderedDict()


def timed(owner, name, label=None):
    fn = getattr(owner, name)
    label = label or "%s.%s" % (getattr(owner, "__name__", str(owner)), name)

    @functools.wraps(fn)
    def wrap(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            acc[label] = acc.get(label, 0.0) + time.perf_counter() - t
    setattr(owner, name, wrap)


rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
os.environ["DBN_B200_DATA_PARALLEL"] = "1"

timed(M._DrawAhead, "get", "_DrawAhead.get (waiting for the helper thread's draws)")
timed(E, "upload_f32")
timed(E.DeviceRbm, "__init__", "DeviceRbm.__init__")
timed(E.DeviceRbm, "export", "DeviceRbm.export")
timed(E, "fit_rbm_data_parallel")
timed(dp_peer.RealPeerGroup, "__init__", "RealPeerGroup.__init__ (arena, IPC handles)")
timed(dp_peer.RealPeerGroup, "close", "RealPeerGroup.close")
timed(dp_peer.DpPeerRunner, "__init__", "DpPeerRunner.__init__ (stage slice, workspace, barrier)")
timed(dp_peer.DpPeerRunner, "run", "DpPeerRunner.run (gather + steps + finalize, host side)")
timed(dist, "broadcast", "dist.broadcast")
timed(dist, "barrier", "dist.barrier")

V = H = 4096
bs, n = 32768, 32768 * 4
gen = torch.Generator(device="cpu").manual_seed(0)
Xh = torch.empty(n, V, dtype=torch.float32).pin_memory()
for s in range(0, n, 16384):
    Xh[s:s + 16384] = (torch.rand(16384, V, generator=gen) < 0.5).float()


def fit():
    np.random.seed(5)
    m = dbn_b200.BinaryRBM(n_hidden_units=H, learning_rate=0.01, n_epochs=1, batch_size=bs, verbose=False, precision="bf16")
    m.fit(Xh.numpy())
    return m


fit()
acc.clear()
reps = 3
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
for _ in range(reps):
    fit()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / reps
if rank == 0:
    print("world %d  wall per fit: %.1f ms  (%.0f samples/s)" % (world, wall * 1e3, n / wall))
    for k, v in acc.items():
        print("  %-70s %8.1f ms" % (k, v / reps * 1e3))
if world > 1:
    dist.destroy_process_group()
