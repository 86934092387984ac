"""Parity of the persistent fine-tune kernel (csrc/mlp_tc_epoch.cuh: one cooperative launch per back-propagation
iteration, reference loop dbn/models.py:439-465) against the float64 oracle and against the tiled path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from gpu_harness import lib, relF, run_mlp_epoch, shadow_round  # noqa: E402
from oracle import dbn_oracle as orc  # noqa: E402


def _net(rng, n_in, hs, n_out):
    layers, prev = [], n_in
    for h in hs:
        layers.append((rng.standard_normal((h, prev)) / np.sqrt(prev), rng.standard_normal(h) * 0.1))
        prev = h
    return layers, rng.standard_normal((n_out, prev)) / np.sqrt(prev), rng.standard_normal(n_out) * 0.1


def _oracle_epoch(layers, Wo, bo, X, Y, act, head, lr, l2, bs, precision, masks=None):
    """dbn/models.py:439-465 in float64.  The device contracts the 16-bit shadows of the hidden weights, so the oracle
    takes its gradients at the weights rounded the same way and applies them to the un-rounded masters.  Also returns
    the yardsticks of tests/test_gpu_parity.py::test_finetune_step_against_oracle summed over the mini-batches: the size
    of the terms each weight / bias change is a signed sum of (lr/bs x || |delta|^T |a| ||, lr/bs x || sum |delta| ||)."""
    lay = [[W.copy(), c.copy()] for (W, c) in layers]
    Wo, bo = Wo.copy(), bo.copy()
    n = X.shape[0]
    losses = np.zeros(n)
    nl = len(lay) + 1
    scale_w, scale_b = np.zeros(nl), np.zeros(nl)
    for s in range(0, n, bs):
        q = [(shadow_round(W, precision), c) for (W, c) in lay]
        mk = [m[s:s + bs] for m in masks] if masks is not None else None
        gW, gb, out, acts, deltas = orc.finetune_grads(q, Wo, bo, X[s:s + bs], Y[s:s + bs], mk, act, head)
        losses[s:s + bs] = orc.sample_loss(out, Y[s:s + bs], head).sum(axis=1) / (Y.shape[1] if head == "softmax" else 1)
        decay = 1.0 - lr * l2 / n
        for l in range(nl):
            scale_w[l] += lr / bs * np.linalg.norm(np.abs(deltas[l]).T @ np.abs(acts[l]))
            scale_b[l] += lr / bs * np.linalg.norm(np.abs(deltas[l]).sum(axis=0))
        for l in range(len(lay)):
            lay[l][0] = decay * lay[l][0] - lr / bs * gW[l]
            lay[l][1] = lay[l][1] - lr / bs * gb[l]
        Wo = decay * Wo - lr / bs * gW[-1]
        bo = bo - lr / bs * gb[-1]
    return lay, Wo, bo, losses, scale_w, scale_b


CASES = [
    (784, [500, 500, 2000], 10, 256, 512 + 96, "sigmoid", "softmax"),   # config 3 (short last mini-batch, split-K backward)
    (64, [256, 256], 10, 32, 96, "relu", "softmax"),                    # config 1 widths
    (96, [128], 3, 64, 128, "sigmoid", "linear"),                       # one hidden layer, regression head
    (200, [136, 72, 104, 64], 5, 128, 256, "sigmoid", "softmax"),       # four hidden layers, ragged widths
]
EDGE_CASES = [
    (784, [500, 500, 2000], 10, 100, 250, "sigmoid", "softmax"),        # config 3 at batch 100 (short last mini-batch)
    (64, [64], 1, 16, 40, "sigmoid", "linear"),                         # smallest network of the plan, one output
    (32, [32, 32, 32, 32], 16, 16, 48, "relu", "softmax"),              # four minimal layers, widest head
    (100, [200, 200], 3, 250, 250 + 9, "sigmoid", "softmax"),           # mini-batch just under the 256-row tile
    (784, [500, 500], 10, 256, 512 + 40, "sigmoid", "softmax"),         # last hidden layer 500 wide: 500 % 8 == 4, the head's last chunk is half
    (784, [300], 10, 32, 96, "relu", "softmax"),                        # one 300-wide layer at the reference's default batch
    (200, [100, 100], 2, 64, 128, "sigmoid", "linear"),                 # the reference's default structure [100, 100]
]


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("n_in,hs,n_out,bs,n,act,head", CASES)
def test_iteration_against_oracle(precision, n_in, hs, n_out, bs, n, act, head):
    import ctypes as C
    dims = (C.c_int32 * len(hs))(*hs)
    prev = lib().dbn_set_tc_finetune(1)      # (the suite may run with DBN_B200_TC_FINETUNE=0)
    try:
        assert lib().dbn_mlp_tc_epoch_plan(1, n_in, dims, len(hs), n_out, min(bs, n)) > 0, "network expected to fit the plan"
    finally:
        lib().dbn_set_tc_finetune(prev)
    rng = np.random.default_rng(n_in + n_out + len(hs))
    layers, Wo, bo = _net(rng, n_in, hs, n_out)
    X = rng.integers(0, 17, (n, n_in)) / 16.0
    Y = np.eye(n_out)[rng.integers(0, n_out, n)] if head == "softmax" else rng.standard_normal((n, n_out))
    lr, l2 = 0.1, 1.0
    lay, Wo2, bo2, losses, scale_w, scale_b = _oracle_epoch(layers, Wo, bo, X, Y, act, head, lr, l2, bs, precision)
    res = run_mlp_epoch(layers, Wo, bo, X, Y, act, head, precision, bs, lr=lr, l2=l2, tc=True)
    assert res["launches"] <= 2, "the iteration must be one persistent launch"
    # same bound as the single-step test of the tiled path: 2e-3 per contraction, compounding through the chain
    # (depth), four times the margin for relu's gate flips; measured against the size of the summed terms
    depth = len(hs) + 1
    gate = 4.0 if act == "relu" else 1.0
    tol = 2e-3 * depth * gate
    for l in range(len(hs)):
        assert np.linalg.norm(res["layers"][l][0] - lay[l][0]) < tol * scale_w[l] + 2e-6 * np.linalg.norm(lay[l][0]), "layer %d weights" % l
        assert np.linalg.norm(res["layers"][l][1] - lay[l][1]) < tol * scale_b[l] + 1e-6, "layer %d bias" % l
        assert np.array_equal(res["shadows"][l], shadow_round(res["layers"][l][0], precision)), "shadow must be the rounded master"
    assert np.linalg.norm(res["Wo"] - Wo2) < tol * scale_w[-1] + 1e-6 and np.linalg.norm(res["bo"] - bo2) < tol * scale_b[-1] + 1e-6
    assert np.max(np.abs(res["loss"] - losses)) < tol * 10 * max(1.0, np.max(np.abs(losses)))


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("n_in,hs,n_out,bs,n,act,head", CASES[:2])
def test_iteration_matches_tiled_path(precision, n_in, hs, n_out, bs, n, act, head):
    rng = np.random.default_rng(5)
    layers, Wo, bo = _net(rng, n_in, hs, n_out)
    X = rng.integers(0, 17, (n, n_in)) / 16.0
    Y = np.eye(n_out)[rng.integers(0, n_out, n)]
    r1 = run_mlp_epoch(layers, Wo, bo, X, Y, act, head, precision, bs, tc=True, iters=2)
    r0 = run_mlp_epoch(layers, Wo, bo, X, Y, act, head, precision, bs, tc=False, iters=2)
    assert r1["launches"] <= 4 and r0["launches"] > 8
    gate = 4.0 if act == "relu" else 1.0
    for l in range(len(hs)):
        assert relF(r1["layers"][l][0] - layers[l][0], r0["layers"][l][0] - layers[l][0]) < 2e-3 * gate
        assert relF(r1["layers"][l][1] - layers[l][1], r0["layers"][l][1] - layers[l][1]) < 2e-3 * gate
    assert relF(r1["Wo"] - Wo, r0["Wo"] - Wo) < 2e-3 * gate
    r2 = run_mlp_epoch(layers, Wo, bo, X, Y, act, head, precision, bs, tc=True, iters=2)
    for l in range(len(hs)):
        assert np.array_equal(r1["layers"][l][0], r2["layers"][l][0]), "deterministic"


def test_full_size_iteration_is_reproducible_and_equals_the_tiled_path():
    """BASELINE config 3 at its full size (60000 rows, 235 mini-batches x 7 phases in one launch): bit-identical from run
    to run (every hand-off of the kernel is ordered), and the same training trajectory as the tiled path.  The two paths
    differ in the order of their fp32 sums (split-K), which a 16-bit shadow rounding amplifies step after step, so the
    comparison uses a learning rate at which the trajectory is stable (at lr 0.1 on random labels the softmax saturates
    within ten steps and BOTH paths leave the float64 oracle by the same few percent: tools/ft_drift_probe.py), and a
    bound that only a wrong operand or a missed hand-off would break (those give O(1) differences)."""
    rng = np.random.default_rng(3)
    layers, Wo, bo = _net(rng, 784, [500, 500, 2000], 10)
    X = (rng.random((60000, 784)) < 0.13).astype(np.float64)
    Y = np.eye(10)[rng.integers(0, 10, 60000)]
    r1 = run_mlp_epoch(layers, Wo, bo, X, Y, "sigmoid", "softmax", "bf16", 256, lr=0.01, tc=True)
    r2 = run_mlp_epoch(layers, Wo, bo, X, Y, "sigmoid", "softmax", "bf16", 256, lr=0.01, tc=True)
    r0 = run_mlp_epoch(layers, Wo, bo, X, Y, "sigmoid", "softmax", "bf16", 256, lr=0.01, tc=False)
    assert r1["launches"] <= 2
    for l in range(3):
        assert np.array_equal(r1["layers"][l][0], r2["layers"][l][0]) and np.array_equal(r1["layers"][l][1], r2["layers"][l][1])
        assert relF(r1["layers"][l][0] - layers[l][0], r0["layers"][l][0] - layers[l][0]) < 5e-2
    assert np.array_equal(r1["Wo"], r2["Wo"]) and np.array_equal(r1["loss"], r2["loss"])
    assert relF(r1["Wo"] - Wo, r0["Wo"] - Wo) < 5e-2 and np.max(np.abs(r1["loss"] - r0["loss"])) < 0.1


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("n_in,hs,n_out,bs,n,act,head", [CASES[0], CASES[1], CASES[3]])
def test_dropout_iteration_against_oracle(precision, n_in, hs, n_out, bs, n, act, head):
    """Dropout inside the persistent kernel (dbn/models.py:401-411: the input and every hidden activation are multiplied
    by Binomial(1, p) masks, no rescale; the derivative sees the masked activation).  With INJECTED masks -- the form the
    numpy-replay mode uses -- the iteration is checked against the float64 oracle on the same masks: the masked input is
    produced one phase ahead into a double buffer by all CTAs, the hidden masks are applied in the forward epilogue."""
    rng = np.random.default_rng(n_in + 7 * len(hs))
    layers, Wo, bo = _net(rng, n_in, hs, n_out)
    X = rng.integers(0, 17, (n, n_in)) / 16.0
    Y = np.eye(n_out)[rng.integers(0, n_out, n)] if head == "softmax" else rng.standard_normal((n, n_out))
    masks = [(rng.random((n, w)) < 0.8).astype(np.float64) for w in [n_in] + hs]
    lr, l2 = 0.1, 1.0
    lay, Wo2, bo2, losses, scale_w, scale_b = _oracle_epoch(layers, Wo, bo, X, Y, act, head, lr, l2, bs, precision, masks)
    res = run_mlp_epoch(layers, Wo, bo, X, Y, act, head, precision, bs, lr=lr, l2=l2, tc=True, keep_p=0.8, masks=masks)
    assert res["launches"] <= 2, "the iteration must be one persistent launch"
    depth = len(hs) + 1
    gate = 4.0 if act == "relu" else 1.0
    tol = 2e-3 * depth * gate
    for l in range(len(hs)):
        assert np.linalg.norm(res["layers"][l][0] - lay[l][0]) < tol * scale_w[l] + 2e-6 * np.linalg.norm(lay[l][0]), "layer %d weights" % l
        assert np.linalg.norm(res["layers"][l][1] - lay[l][1]) < tol * scale_b[l] + 1e-6, "layer %d bias" % l
    assert np.linalg.norm(res["Wo"] - Wo2) < tol * scale_w[-1] + 1e-6 and np.linalg.norm(res["bo"] - bo2) < tol * scale_b[-1] + 1e-6
    assert np.max(np.abs(res["loss"] - losses)) < tol * 10 * max(1.0, np.max(np.abs(losses)))
    # and the masks matter: without them the result is somewhere else
    plain = run_mlp_epoch(layers, Wo, bo, X, Y, act, head, precision, bs, lr=lr, l2=l2, tc=True)
    assert np.linalg.norm(plain["Wo"] - res["Wo"]) > 10 * tol * scale_w[-1] * 0.01


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("n_in,hs,n_out,bs,n,act,head", CASES[:2])
def test_philox_dropout_matches_the_tiled_path(precision, n_in, hs, n_out, bs, n, act, head):
    """Philox dropout: the persistent kernel draws the SAME masks as the tiled path (stream = layer, counter = global row
    and column, epoch = iteration), so two iterations agree with the tiled launches to rounding, short last batch
    included, and a different seed gives a different model."""
    rng = np.random.default_rng(11)
    layers, Wo, bo = _net(rng, n_in, hs, n_out)
    X = rng.integers(0, 17, (n, n_in)) / 16.0
    Y = np.eye(n_out)[rng.integers(0, n_out, n)]
    kw = dict(lr=0.05, iters=2, keep_p=0.75, seed=0x1234567890ABCDEF)
    r1 = run_mlp_epoch(layers, Wo, bo, X, Y, act, head, precision, bs, tc=True, **kw)
    r0 = run_mlp_epoch(layers, Wo, bo, X, Y, act, head, precision, bs, tc=False, **kw)
    assert r1["launches"] <= 4 and r0["launches"] > 8
    gate = 4.0 if act == "relu" else 1.0
    for l in range(len(hs)):
        assert relF(r1["layers"][l][0] - layers[l][0], r0["layers"][l][0] - layers[l][0]) < 2e-3 * gate
        assert relF(r1["layers"][l][1] - layers[l][1], r0["layers"][l][1] - layers[l][1]) < 2e-3 * gate
    assert relF(r1["Wo"] - Wo, r0["Wo"] - Wo) < 2e-3 * gate
    assert np.max(np.abs(r1["loss"] - r0["loss"])) < 2e-2
    r2 = run_mlp_epoch(layers, Wo, bo, X, Y, act, head, precision, bs, tc=True, **kw)
    assert all(np.array_equal(r1["layers"][l][0], r2["layers"][l][0]) for l in range(len(hs))), "deterministic"
    kw["seed"] = 99
    r3 = run_mlp_epoch(layers, Wo, bo, X, Y, act, head, precision, bs, tc=True, **kw)
    assert relF(r3["Wo"] - Wo, r1["Wo"] - Wo) > 1e-2


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
@pytest.mark.parametrize("n_in,hs,n_out,bs,n,act,head", EDGE_CASES)
def test_edge_shapes_against_oracle(precision, n_in, hs, n_out, bs, n, act, head):
    """The corners of the plan (odd batch sizes, minimal widths, four layers, one output, 16 outputs)."""
    test_iteration_against_oracle(precision, n_in, hs, n_out, bs, n, act, head)
