"""Multi-GPU parity check of every data-parallel exchange (run under torchrun on N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py

For each case every rank trains the same RBM data-parallel three ways --

    peer     dp_peer.DpPeerRunner: dW scattered into the owners' receive slots from the contraction's epilogue over
             NVLink peer memory, sharded update, bf16 rows written into every shadow (the default for BF16 layers
             with H % world == 0: what bench.py times at N > 1)
    sharded  engine.DpRbmEpochRunner, NCCL reduce-scatter + sharded update + all-gather of the shadow
    chunked  engine.DpRbmEpochRunner with n_chunks = 4: NCCL all-reduce in row blocks, replicated update

-- and rank 0 also trains it alone (RbmEpochRunner) with the same seed and shuffles.  Philox counters are keyed by
the global row index, so the sampled states are identical everywhere and the trained weights may differ only by fp32
summation order: every path must agree with the single-GPU run within 1e-4 * |W - W0| (1e-5 in FP32 mode), and W, b,
c must be BIT-IDENTICAL across the replicas of a path."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dbn_b200 import dp_peer, engine as E  # noqa: E402


def replicas_identical(arrs, dev):
    same = True
    for a in arrs:
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        tmax, tmin = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        same = same and bool((tmax == tmin).all().item())
    return same


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    cases = [(640, 512, 256 * world, 3, "bf16"), (200, 136, 64 * world, 2, "fp32"), (1024, 1280, 512 * world, 2, "fp16"),
             (4096, 4096, 512 * world, 2, "fp16")]
    for (V, H, bs, steps, prec) in cases:
        rng = np.random.default_rng(1)
        n = bs * steps
        X = (rng.random((n, V)) < 0.3).astype(np.float32)
        W0 = rng.standard_normal((H, V)) / np.sqrt(V)
        b0, c0 = rng.standard_normal(V) * 0.1, rng.standard_normal(H) * 0.1
        Xd = torch.from_numpy(X).to(dev)
        orders = [np.random.default_rng(100 + e).permutation(n) for e in range(2)]
        results = {}
        modes = ["chunked"]
        if prec in ("bf16", "fp16") and H % world == 0:
            modes = ["peer", "sharded", "chunked"]
        for mode in modes:
            layer = E.DeviceRbm(W0, b0, c0, "sigmoid", prec, dev)
            group = None
            if mode == "peer":
                nl = n // world
                group = dp_peer.RealPeerGroup(rank, world, dp_peer.ArenaLayout(world, V, H, nl, E.OP_DTYPE[layer.prec]), dev)
                runner = dp_peer.DpPeerRunner(layer, Xd[rank * nl:(rank + 1) * nl].contiguous(), n, bs, 1, 0.05, 4242,
                                              group.member)
            else:
                runner = E.DpRbmEpochRunner(layer, Xd, bs, 1, 0.05, 4242, (rank, world), n_chunks=4 if mode == "chunked" else 1)
                assert runner.sharded == (mode == "sharded"), (mode, runner.sharded)
            for o in orders:
                runner.run(o)
            torch.cuda.synchronize()
            Wd, bd, cd = layer.export()
            shadow = layer.Wb[:, :V].float().cpu().numpy() if layer.Wb is not None else np.zeros(1)
            if group is not None:
                layer.Wb = layer.Wb.clone()
                group.close()
            results[mode] = (Wd, bd, cd, replicas_identical((Wd, bd, cd, shadow), dev))
        if rank == 0:
            # the tiled kernels on both sides: this check is about the exchange (the persistent epoch kernel of small
            # layers sums in another order -- its parity tests are tests/test_gpu_tc_epoch.py, tolerance 2e-3)
            single = E.DeviceRbm(W0, b0, c0, "sigmoid", prec, dev)
            prev = layer.lib.dbn_set_tc_epoch(0)
            r1 = E.RbmEpochRunner(single, Xd, bs, 1, 0.05, 4242)
            for o in orders:
                r1.run(order=o)
            torch.cuda.synchronize()
            layer.lib.dbn_set_tc_epoch(prev)
            Ws, bs_, cs = single.export()
            upd = np.linalg.norm(Ws - W0)
            tol = 1e-5 if prec == "fp32" else 1e-4   # identical samples; fp32 partial sums in a different order
            for mode, (Wd, bd, cd, same) in results.items():
                err = np.linalg.norm(Wd - Ws)
                errb, errc = np.max(np.abs(bd - bs_)), np.max(np.abs(cd - cs))
                good = same and err <= tol * max(upd, 1e-9) + 1e-7 and errb < 2e-5 and errc < 2e-5
                ok = ok and good
                print("dp_check V=%d H=%d bs=%d %s world=%d %-7s: replicas bit-identical=%s  |W_dp-W_1|/|W_1-W_0|=%.2e  db=%.1e dc=%.1e  %s"
                      % (V, H, bs, prec, world, mode, same, err / max(upd, 1e-30), errb, errc, "OK" if good else "FAIL"), flush=True)
            if "peer" in results and "sharded" in results:
                d = np.linalg.norm(results["peer"][0] - results["sharded"][0]) / max(upd, 1e-30)
                print("dp_check V=%d H=%d          peer vs sharded: |dW|/|W_1-W_0|=%.2e" % (V, H, d), flush=True)
                ok = ok and d <= 1e-4
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
