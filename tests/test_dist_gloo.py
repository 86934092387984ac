"""World-size-2 `gloo` test (CPU) of the host-side data-parallel logic: the row sharding of each
global mini-batch, the global-row Philox keys and the sum all-reduce of the raw deltas.  The compute
on each rank is the float64 oracle (no GPU here); the property checked is the one the CUDA
data-parallel runner relies on: summed per-rank deltas == full-batch deltas, and samples
independent of the number of ranks."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gpu_harness as G
    from dbn_b200 import engine as E
    from oracle import dbn_oracle as orc
    rng = np.random.default_rng(0)
    V, H, bs, steps, seed, epoch = 24, 16, 32, 3, 1234, 2
    n = bs * steps
    X = (rng.random((n, V)) < 0.4).astype(np.float64)
    W = rng.standard_normal((H, V)) / np.sqrt(V)
    b, c = rng.standard_normal(V) * 0.1, rng.standard_normal(H) * 0.1
    order = np.random.default_rng(7).permutation(n)            # same shuffle on every rank
    bl = bs // world
    local = E.dp_local_rows(order, steps, world, bl, rank)
    ok = True
    for s in range(steps):
        rows = local[s * bl:(s + 1) * bl]
        row0 = E.dp_global_row0(s, bs, rank, bl)
        U = G.philox_uniforms(seed, 0, epoch, row0, bl, H).astype(np.float64)[:, None, :]
        dW, db, dc, ch = orc.cd_deltas(W, b, c, X[rows], U, 1, "sigmoid")
        raw = np.zeros((H + 1, V + 1))
        raw[:H, :V], raw[:H, V], raw[H, :V] = dW, dc, db
        t = torch.from_numpy(raw)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        # full-batch reference with the global-row uniforms
        Uf = G.philox_uniforms(seed, 0, epoch, s * bs, bs, H).astype(np.float64)[:, None, :]
        fW, fb, fc, fch = orc.cd_deltas(W, b, c, X[order[s * bs:(s + 1) * bs]], Uf, 1, "sigmoid")
        got = t.numpy()
        ok &= np.allclose(got[:H, :V], fW, atol=1e-10) and np.allclose(got[:H, V], fc, atol=1e-10)
        ok &= np.allclose(got[H, :V], fb, atol=1e-10)
        ok &= np.array_equal(ch["samples"][0], fch["samples"][0][rank * bl:(rank + 1) * bl])
        # identical update on every replica
        W = W + 0.05 * got[:H, :V] / bs
        b = b + 0.05 * got[H, :V] / bs
        c = c + 0.05 * got[:H, V] / bs
    gathered = [torch.zeros(H, V, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(W))
    ok &= all(torch.equal(gathered[0], g) for g in gathered)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_data_parallel_sharding_and_allreduce_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r, False) for r in range(world)), dict(ret)


def test_dp_partition_helpers():
    sys.path.insert(0, ROOT)
    from dbn_b200 import engine as E
    order = np.arange(24)[::-1].copy()
    parts = [E.dp_local_rows(order, 3, 2, 4, r) for r in range(2)]
    # every global mini-batch is the concatenation of the ranks' slices, in rank order
    for s in range(3):
        assert np.array_equal(np.concatenate([p[s * 4:(s + 1) * 4] for p in parts]), order[s * 8:(s + 1) * 8])
    assert E.dp_global_row0(2, 8, 1, 4) == 20
    assert E.dp_chunks(4096, 2) == [(0, 2048), (2048, 4096)]
    assert E.dp_chunks(500, 8) == [(0, 256), (256, 500)]
    assert E.dp_chunks(500, 1) == [(0, 500)] and E.dp_chunks(4096, 1) == [(0, 4096)]
    with pytest.raises(ValueError):
        E.dp_check_shapes(100, 32, 3)
