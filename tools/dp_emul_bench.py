"""The kernels of the peer-memory data-parallel step timed on ONE GPU with 8 virtual ranks (local "peers": no NVLink),
at the per-rank shape of config 5 on 8 GPUs (4096 rows of the 32768-row batch, 512 owned rows of W).  Separates
what the kernels cost by themselves from what the wire adds in tools/dp_trace.py."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dbn_b200 import _lib as L, dp_peer, engine as E  # noqa: E402

dev = torch.device("cuda")
world, V, H = 8, 4096, 4096
bs = 32768
n = bs * 2
nl = n // world
lay = dp_peer.ArenaLayout(world, V, H, nl, torch.float16)
group = dp_peer.VirtualPeerGroup(world, lay, dev)
W = (torch.randn(H, V, generator=torch.Generator().manual_seed(0)) / 64).numpy()
runners = []
for m in group.members():
    layer = E.DeviceRbm(W, np.zeros(V), np.zeros(H), "sigmoid", "fp16", dev)
    Xl = (torch.rand(nl, V, device=dev) < 0.5).float()
    runners.append(dp_peer.DpPeerRunner(layer, Xl, n, bs, 1, 0.01, 99, m))
for r in runners:
    r.gather(np.arange(n))
# all flags far in the future: no kernel ever waits
for m in group.members():
    dp_peer.device_view(m.base + lay.off_flags, (2 * L.MAX_PEERS,), torch.int64, dev).fill_(1 << 60)
r = runners[0]
lib = r.lib


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


a = r.args
a.row0, a.rng.row0 = 0, 0
t_chain = timeit(lambda: L.check(lib.dbn_rbm_cd_step(E.stream_ptr(), C.byref(a))))
t_scatter = timeit(lambda: L.check(lib.dbn_rbm_delta_scatter(E.stream_ptr(), C.byref(a), C.byref(r.x))))
raw = torch.empty(H + 1, lay.ldr, device=dev)
a2 = r.layer._step_args(r.Xp, 1, r.scale, r.ws, 99, r.epoch_ctr)
a2.B, a2.flags, a2.raw_delta = r.bl, L.STEP_SKIP_UPDATE | L.STEP_BIAS_STATS, E.mat(raw)
t_raw = timeit(lambda: L.check(lib.dbn_rbm_delta_rows(E.stream_ptr(), C.byref(a2), 0, H)))
t_signal = timeit(lambda: L.check(lib.dbn_dp_signal(E.stream_ptr(), C.byref(r.x), r.stats, 1)))
t_apply = timeit(lambda: L.check(lib.dbn_dp_apply(E.stream_ptr(), C.byref(r.x), r.scale, E.ptr(r.layer.W), V, E.ptr(r.layer.b),
                                                  E.ptr(r.layer.c), r.stats, E.ptr(r.counter), 1)))
t_wait = timeit(lambda: L.check(lib.dbn_dp_wait(E.stream_ptr(), C.byref(r.x), 1, 1)))
print("one GPU, 8 virtual ranks, 4096 rows per rank (us):  chain %.1f | dW scattered %.1f (dW to one local buffer %.1f) | "
      "signal %.1f | apply %.1f | wait %.1f" % (t_chain, t_scatter, t_raw, t_signal, t_apply, t_wait))
group.close()
