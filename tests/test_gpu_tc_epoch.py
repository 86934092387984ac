"""Parity of the persistent tcgen05 epoch kernel (csrc/rbm_tc_epoch.cuh: one cooperative launch per layer-epoch,
reference loop dbn/models.py:105-119) against the float64 oracle and against the tiled four-launch path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from gpu_harness import (guard_band_uniforms, lib, relF, run_rbm_epoch, shadow_round, tc_epoch_plan)  # noqa: E402
from oracle import dbn_oracle as orc  # noqa: E402


def _layer(rng, V, H, scale=1.0):
    W = rng.standard_normal((H, V)) * scale / np.sqrt(V)
    return W, rng.standard_normal(V) * 0.1, rng.standard_normal(H) * 0.1


def _oracle_epoch(W, b, c, X, U, bs, k, act, lr, precision):
    """dbn/models.py:108-119 in float64 on the same uniforms.  The device contracts the 16-bit shadow of W, so the
    oracle sees the weights rounded the same way at every step (the fp32 master keeps the un-rounded sum)."""
    W, b, c = W.copy(), b.copy(), c.copy()
    n = X.shape[0]
    for s in range(0, n, bs):
        Wq = shadow_round(W, precision)
        dW, db, dc, _ = orc.cd_deltas(Wq, b, c, X[s:s + bs], U[s:s + bs], k, act)
        W += lr / bs * dW
        b += lr / bs * db
        c += lr / bs * dc
    return W, b, c


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("V,H,bs,n,k,act", [
    (784, 500, 256, 768, 1, "sigmoid"),      # config 2 layer 1
    (500, 500, 256, 512 + 96, 1, "sigmoid"), # config 2 layer 2, short last mini-batch
    (500, 2000, 256, 512, 1, "sigmoid"),     # config 2 layer 3: split-K visible phase
    (256, 256, 32, 96, 1, "relu"),           # config 1 layer 2 shape in a tensor-core mode
    (320, 200, 64, 192, 3, "sigmoid"),       # CD-3, ragged widths
    (333, 257, 48, 48 * 3 + 7, 1, "sigmoid"),   # odd widths (not multiples of 8), odd batch, short last mini-batch
    (72, 136, 64, 200, 2, "relu"),           # narrow layer, CD-2
    (784, 500, 100, 250, 1, "sigmoid"),      # batch 100 (the reference's examples use 32 / 100-row batches)
    (64, 128, 16, 40, 1, "sigmoid"),         # smallest shapes of the plan
    (256, 256, 250, 250 + 13, 1, "sigmoid"), # mini-batch just under the 256-row tile
    (1000, 1100, 128, 256, 1, "sigmoid"),    # wider than config 2 on both sides
])
def test_epoch_against_oracle(precision, V, H, bs, n, k, act):
    assert tc_epoch_plan(V, H, min(bs, n)), "shape expected to fit the persistent kernel"
    rng = np.random.default_rng(V + H + k)
    W, b, c = _layer(rng, V, H)
    X = (rng.random((n, V)) < 0.3).astype(np.float64)
    lr = 0.05
    # uniforms guard-banded against the float64 chain so that every u < p is decided the same way on both sides
    U = rng.random((n, k, H), dtype=np.float32)
    Wt, bt, ct = W.copy(), b.copy(), c.copy()
    for s in range(0, n, bs):
        Wq = shadow_round(Wt, precision)
        Vt = X[s:s + bs]
        for t in range(k):
            Pt = orc.hidden_means(Wq, ct, Vt, act)
            U[s:s + bs, t, :] = guard_band_uniforms(U[s:s + bs, t, :], Pt, 2e-2, rng)
            Vt = orc.visible_means(Wq, bt, (U[s:s + bs, t, :] < Pt).astype(np.float64), act)
        dW, db, dc, _ = orc.cd_deltas(Wq, bt, ct, X[s:s + bs], U[s:s + bs].astype(np.float64), k, act)
        Wt += lr / bs * dW; bt += lr / bs * db; ct += lr / bs * dc
    Wo, bo, co = _oracle_epoch(W, b, c, X, U.astype(np.float64), bs, k, act, lr, precision)
    res = run_rbm_epoch(W, b, c, X, bs, k, act, precision, lr=lr, U=U.reshape(n, k * H), tc_epoch=True)
    assert res["launches"] <= 2, "the epoch must be one persistent launch"
    tol = 2e-3   # 16-bit tensor-core modes (north_star); measured on the parameter CHANGE of the epoch
    assert relF(res["W"] - W, Wo - W) < tol * (3 if k > 1 else 1.5), relF(res["W"] - W, Wo - W)
    assert relF(res["b"] - b, bo - b) < tol * 3, relF(res["b"] - b, bo - b)
    assert relF(res["c"] - c, co - c) < tol * 3, relF(res["c"] - c, co - c)
    assert np.array_equal(res["Wb"], shadow_round(res["W"], precision)), "shadow must be the rounded master"


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("V,H,bs,n", [(784, 500, 256, 2048 + 96), (500, 2000, 256, 1024), (500, 500, 256, 1024)])
def test_epoch_matches_tiled_path(precision, V, H, bs, n):
    """Philox mode, two epochs: the persistent kernel draws the same counters as the tiled path, so the trained
    weights agree up to the (rare) samples whose probability sits within fp32 round-off of its uniform."""
    rng = np.random.default_rng(11)
    W, b, c = _layer(rng, V, H)
    X = (rng.random((n, V)) < 0.2).astype(np.float64)
    r1 = run_rbm_epoch(W, b, c, X, bs, 1, "sigmoid", precision, seed=5, lr=0.1, tc_epoch=True, epochs=2)
    r0 = run_rbm_epoch(W, b, c, X, bs, 1, "sigmoid", precision, seed=5, lr=0.1, tc_epoch=False, epochs=2)
    assert r1["launches"] <= 4 and r0["launches"] >= 2 * 4 * (n // bs)
    assert relF(r1["W"] - W, r0["W"] - W) < 2e-3, relF(r1["W"] - W, r0["W"] - W)
    assert relF(r1["b"] - b, r0["b"] - b) < 2e-3 and relF(r1["c"] - c, r0["c"] - c) < 2e-3
    # and it is deterministic
    r2 = run_rbm_epoch(W, b, c, X, bs, 1, "sigmoid", precision, seed=5, lr=0.1, tc_epoch=True, epochs=2)
    assert np.array_equal(r1["W"], r2["W"]) and np.array_equal(r1["b"], r2["b"]) and np.array_equal(r1["c"], r2["c"])


def test_plan_falls_back():
    l = lib()
    prev = l.dbn_set_tc_epoch(1)      # (the suite may run with DBN_B200_TC_EPOCH=0: the switch is what is tested here)
    try:
        assert l.dbn_rbm_tc_epoch_plan(1, 4096, 4096, 256) == 0      # config 5: CTA-pair kernels
        assert l.dbn_rbm_tc_epoch_plan(0, 784, 500, 256) == 0        # strict FP32: SIMT / cluster kernel
        assert l.dbn_rbm_tc_epoch_plan(1, 784, 500, 256) > 0
        l.dbn_set_tc_epoch(0)
        assert l.dbn_rbm_tc_epoch_plan(1, 784, 500, 256) == 0
    finally:
        l.dbn_set_tc_epoch(prev)


@pytest.mark.parametrize("V,H", [(784, 500), (500, 2000)])
def test_full_size_epochs_are_reproducible(V, H):
    """BASELINE config 2 at its full size (60000 rows = 235 mini-batches per launch, 940 grid barriers): two runs from
    the same state give bit-identical parameters -- every hand-off of the kernel (L2 barrier, TMA refills of activations
    other CTAs wrote, split-K scratch, fixed-point statistics) is ordered, none is a race that happens to work."""
    rng = np.random.default_rng(V)
    W, b, c = _layer(rng, V, H)
    X = (rng.random((60000, V)) < 0.13).astype(np.float32)
    runs = [run_rbm_epoch(W, b, c, X, 256, 1, "sigmoid", "fp16", seed=9, lr=0.01, tc_epoch=True, epochs=2) for _ in range(2)]
    assert runs[0]["launches"] <= 4
    for k in ("W", "b", "c", "Wb"):
        assert np.array_equal(runs[0][k], runs[1][k]), k
    assert np.all(np.isfinite(runs[0]["W"])) and np.linalg.norm(runs[0]["W"] - W) > 0
