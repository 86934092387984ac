"""Multi-GPU parity check (run under torchrun on N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py

Every rank trains the same RBM data-parallel (DpRbmEpochRunner: row-sharded mini-batches, chunked NCCL
all-reduce of dW|db|dc); rank 0 also trains it alone (RbmEpochRunner) with the same seed and shuffle.
Philox counters are keyed by the global row index, so the sampled states are identical and the trained
weights may differ only by fp32 summation order.  Also checks that all replicas hold identical weights."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dbn_b200 import engine as E  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for (V, H, bs, steps, prec) in [(640, 512, 256 * world, 3, "bf16"), (200, 136, 64 * world, 2, "fp32"),
                                    (1024, 1280, 512 * world, 2, "bf16")]:
        rng = np.random.default_rng(1)
        n = bs * steps
        X = (rng.random((n, V)) < 0.3).astype(np.float32)
        W0 = rng.standard_normal((H, V)) / np.sqrt(V)
        b0, c0 = rng.standard_normal(V) * 0.1, rng.standard_normal(H) * 0.1
        Xd = torch.from_numpy(X).to(dev)
        orders = [np.random.default_rng(100 + e).permutation(n) for e in range(2)]
        layer = E.DeviceRbm(W0, b0, c0, "sigmoid", prec, dev)
        runner = E.DpRbmEpochRunner(layer, Xd, bs, 1, 0.05, 4242, (rank, world), n_chunks=4)
        for o in orders:
            runner.run(order=o)
        torch.cuda.synchronize()
        Wd, bd, cd = layer.export()
        # replicas identical?
        t = torch.from_numpy(Wd).to(dev)
        tmax, tmin = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        same = bool((tmax == tmin).all().item())
        if rank == 0:
            single = E.DeviceRbm(W0, b0, c0, "sigmoid", prec, dev)
            r1 = E.RbmEpochRunner(single, Xd, bs, 1, 0.05, 4242)
            for o in orders:
                r1.run(order=o)
            torch.cuda.synchronize()
            Ws, bs_, cs = single.export()
            upd = np.linalg.norm(Ws - W0)
            err = np.linalg.norm(Wd - Ws)
            errb = np.max(np.abs(bd - bs_))
            errc = np.max(np.abs(cd - cs))
            tol = 1e-5 if prec == "fp32" else 1e-4   # identical samples; fp32 partial sums in a different order
            good = same and err <= tol * max(upd, 1e-9) + 1e-7 and errb < 1e-5 and errc < 1e-5
            ok = ok and good
            print("dp_check V=%d H=%d bs=%d %s world=%d: replicas identical=%s  |W_dp-W_1|/|W_1-W_0|=%.2e  db=%.1e dc=%.1e  %s"
                  % (V, H, bs, prec, world, same, err / max(upd, 1e-30), errb, errc, "OK" if good else "FAIL"), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
