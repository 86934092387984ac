"""Per-phase device time of the peer-memory data-parallel step of config 5 (CUDA events after every phase; the events
serialise nothing that is not already serial -- one stream).  Run under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/dp_trace.py
Prints, per rank: wait (weights in) | Gibbs chain (3 contractions) | dW contraction + scatter | signal | apply."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dbn_b200 import dp_peer, engine as E  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
V = H = 4096
bs = int(os.environ.get("GLOBAL_BATCH", 32768))
steps_per_epoch = 4
n = bs * steps_per_epoch
nl = n // world
W = (torch.randn(H, V, generator=torch.Generator().manual_seed(0)) / 64).numpy()
layer = E.DeviceRbm(W, np.zeros(V), np.zeros(H), "sigmoid", "fp16", dev)
Xl = (torch.rand(nl, V, device=dev) < 0.5).float()
group = dp_peer.RealPeerGroup(rank, world, dp_peer.ArenaLayout(world, V, H, nl, torch.float16), dev)
r = dp_peer.DpPeerRunner(layer, Xl, n, bs, 1, 0.01, 99, group.member)
r.gather(np.arange(n))
for s in range(8):
    r.step(s % r.steps)
r.drain()
torch.cuda.synchronize()
dist.barrier()
reps = 40
evs = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(reps)]
for i in range(reps):
    r.step_traced(i % r.steps, evs[i])
r.drain()
torch.cuda.synchronize()
acc = np.zeros(5)
for e in evs[4:]:
    acc += [e[j].elapsed_time(e[j + 1]) for j in range(5)]
acc /= (reps - 4)
tot = evs[4][0].elapsed_time(evs[-1][0]) / (reps - 5)
out = "rank %d  wait %.1f us | chain %.1f | dW+scatter %.1f | signal %.1f | apply %.1f   sum %.1f   step-to-step %.1f us" % (
    rank, acc[0] * 1e3, acc[1] * 1e3, acc[2] * 1e3, acc[3] * 1e3, acc[4] * 1e3, acc.sum() * 1e3, tot * 1e3)
gathered = [None] * world
dist.all_gather_object(gathered, out)
if rank == 0:
    print("\n".join(gathered), flush=True)
layer.Wb = layer.Wb.clone()
group.close()
dist.destroy_process_group()
