"""The peer-memory data-parallel exchange on ONE GPU: `world` virtual ranks in one process (VirtualPeerGroup) run the
same kernels as the multi-GPU path -- dW scattered from the contraction's epilogue into the owners' receive slots,
fixed-point bias statistics, ordered slot sum + sharded update + bf16 rows written into every shadow, epoch gather
from the rank that staged a row -- and must train the same weights as the single-GPU epoch runner.  (The multi-GPU
run of the same comparison is tools/dp_check.py under torchrun; logs in profiles/.)"""
import numpy as np
import pytest
import torch

import gpu_harness as G
from dbn_b200 import dp_peer, engine as E

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,V,H,bl,steps,prec", [(2, 640, 512, 128, 3, "bf16"),      # statistics path on both sides
                                                     (4, 1024, 1280, 64, 2, "fp16"),
                                                     (2, 600, 500, 128, 2, "bf16"),      # ragged: single GPU uses the ones row / column
                                                     (8, 4096, 4096, 256, 2, "fp16")])   # config-5 widths: CTA-pair kernel, smem-transposed scatter
def test_virtual_ranks_match_the_single_gpu_runner(world, V, H, bl, steps, prec):
    dev = torch.device("cuda")
    rng = np.random.default_rng(world + V)
    bs = bl * world
    n = bs * steps
    X = (rng.random((n, V)) < 0.3).astype(np.float32)
    W0 = rng.standard_normal((H, V)) / np.sqrt(V)
    b0, c0 = rng.standard_normal(V) * 0.1, rng.standard_normal(H) * 0.1
    orders = [np.random.default_rng(100 + e).permutation(n) for e in range(2)]
    lr, seed = 0.05, 4242
    nl = n // world
    group = dp_peer.VirtualPeerGroup(world, dp_peer.ArenaLayout(world, V, H, nl, E.OP_DTYPE[E.PREC_IDS[prec]]), dev)
    try:
        runners = []
        for m in group.members():
            layer = E.DeviceRbm(W0, b0, c0, "sigmoid", prec, dev)
            Xl = torch.from_numpy(X[m.rank * nl:(m.rank + 1) * nl]).to(dev)
            runners.append(dp_peer.DpPeerRunner(layer, Xl, n, bs, 1, lr, seed, m))
        for o in orders:
            dp_peer.run_epoch_lockstep(runners, o)
        torch.cuda.synchronize()
        res = [r.layer.export() for r in runners]
        shadows = [r.layer.Wb[:, :V].float().cpu().numpy() for r in runners]
    finally:
        group.close()
    # replicas: bit-identical parameters and shadows
    for (Wr, br, cr), sh in zip(res[1:], shadows[1:]):
        assert np.array_equal(Wr, res[0][0]) and np.array_equal(br, res[0][1]) and np.array_equal(cr, res[0][2])
        assert np.array_equal(sh, shadows[0])
    assert np.array_equal(shadows[0], G.shadow_round(res[0][0], prec).astype(np.float32))
    # single-GPU runner, same seed and shuffles: identical samples (Philox keyed by the global row), fp32 sums in another order
    # (the tiled kernels on both sides: this test is about the exchange.  The persistent epoch kernel sums a dot product
    #  in another order, which moves a handful of 16-bit roundings of P0 / V1 -- its own parity tests are in
    #  test_gpu_tc_epoch.py, tolerance 2e-3)
    single = E.DeviceRbm(W0, b0, c0, "sigmoid", prec, dev)
    prev = G.lib().dbn_set_tc_epoch(0)
    try:
        r1 = E.RbmEpochRunner(single, torch.from_numpy(X).to(dev), bs, 1, lr, seed)
        for o in orders:
            r1.run(order=o)
        torch.cuda.synchronize()
    finally:
        G.lib().dbn_set_tc_epoch(prev)
    Ws, bs_, cs = single.export()
    upd = np.linalg.norm(Ws - W0)
    assert upd > 0
    assert np.linalg.norm(res[0][0] - Ws) <= 1e-4 * upd + 1e-7, (np.linalg.norm(res[0][0] - Ws), upd)
    # biases: the two paths round their statistics differently (fp32 values here, bf16 operands through the ones column)
    assert np.max(np.abs(res[0][1] - bs_)) < 2e-5 and np.max(np.abs(res[0][2] - cs)) < 2e-5


@pytest.mark.parametrize("B,V,H,k", [(256, 784, 500, 1), (200, 130, 70, 3), (512, 1024, 768, 1)])
def test_bias_statistics_path_matches_the_ones_column_path_and_the_oracle(B, V, H, k):
    """dc / db taken as fixed-point column sums in the chain's epilogues (DBN_STEP_BIAS_STATS) against the ones row /
    column of the dW contraction, and both against the float64 oracle; the sums are order-independent, so two runs
    agree bit for bit."""
    from dbn_b200 import _lib as L
    from oracle import dbn_oracle as orc
    rng = np.random.default_rng(B + V + H)
    W = torch.as_tensor(rng.standard_normal((H, V)) / np.sqrt(V)).to(torch.bfloat16).double().numpy()
    b, c = rng.standard_normal(V) * 0.1, rng.standard_normal(H) * 0.1
    V0 = (rng.random((B, V)) < 0.2).astype(np.float64)
    U = rng.random((B, k, H), dtype=np.float32)
    for t in range(k):
        ch = orc.cd_chain(W, b, c, V0, U.astype(np.float64), k, "sigmoid")
        Vt = V0 if t == 0 else orc.visible_means(W, b, ch["samples"][t - 1], "sigmoid")
        U[:, t, :] = G.guard_band_uniforms(U[:, t, :], orc.hidden_means(W, c, Vt, "sigmoid"), 2e-2, rng)
    st = G.run_cd_step(W, b, c, V0, k, "sigmoid", "bf16", U=U, raw=True, flags=L.STEP_BIAS_STATS)
    st2 = G.run_cd_step(W, b, c, V0, k, "sigmoid", "bf16", U=U, raw=True, flags=L.STEP_BIAS_STATS)
    dW, db, dc, ch = orc.cd_deltas(W, b, c, V0, U.astype(np.float64), k, "sigmoid")
    assert np.array_equal(st["dc"], st2["dc"]) and np.array_equal(st["db"], st2["db"])
    assert np.array_equal(st["H0s"], ch["samples"][0])
    tol = 2e-3 * (1 if k == 1 else 10)
    assert np.max(np.abs(st["db"] - db)) < tol * B and np.max(np.abs(st["dc"] - dc)) < tol * B
    assert np.linalg.norm(st["dW"] - dW) < tol * np.linalg.norm(ch["P0"].T @ V0)
    # in-place update through the statistics (the update epilogue's owner threads fold them into b and c)
    lr_over_bs = 0.01 / 256
    up = G.run_cd_step(W, b, c, V0, k, "sigmoid", "bf16", U=U, lr_over_bs=lr_over_bs, flags=L.STEP_BIAS_STATS)
    assert np.max(np.abs(up["b"] - (b + lr_over_bs * db))) < 1e-6 + 3 * tol * lr_over_bs * B
    assert np.max(np.abs(up["c"] - (c + lr_over_bs * dc))) < 1e-6 + 3 * tol * lr_over_bs * B
    assert np.max(np.abs(up["W"] - (W + lr_over_bs * dW))) < 1e-6 + 3 * tol * lr_over_bs * B
    # and a second update starts from clean accumulators: same result as the first
    up2 = G.run_cd_step(W, b, c, V0, k, "sigmoid", "bf16", U=U, lr_over_bs=lr_over_bs, flags=L.STEP_BIAS_STATS)
    assert np.array_equal(up["b"], up2["b"]) and np.array_equal(up["c"], up2["c"])
