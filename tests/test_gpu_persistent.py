"""The on-chip persistent CD-k kernel (one cluster launch per epoch, W resident in distributed shared
memory) against the tiled multi-launch path and the float64 oracle."""
import os

import numpy as np
import pytest
import torch

import gpu_harness as G
from dbn_b200 import engine as E
from oracle import dbn_oracle as orc

pytestmark = pytest.mark.gpu


def train(X, W, b, c, act, bs, k, lr, epochs, persistent, seed=77, orders=None):
    os.environ["DBN_B200_PERSISTENT"] = "1" if persistent else "0"
    try:
        dev = torch.device("cuda")
        layer = E.DeviceRbm(W, b, c, act, "fp32", dev)
        Xd = torch.from_numpy(X).to(dev)
        r = E.RbmEpochRunner(layer, Xd, bs, k, lr, seed)
        assert r.persistent == persistent
        for e in range(epochs):
            r.run(order=orders[e])
        torch.cuda.synchronize()
        return layer.export(), r
    finally:
        os.environ.pop("DBN_B200_PERSISTENT", None)


@pytest.mark.parametrize("N,V,H,bs,k,act", [(1437, 64, 256, 32, 1, "relu"),      # config 1, layer 1  (cluster of 8)
                                            (300, 256, 256, 32, 1, "relu"),      # config 1, layer 2
                                            (404, 13, 100, 16, 10, "relu"),      # config 4, CD-10     (cluster of 2)
                                            (101, 40, 24, 16, 3, "sigmoid"),     # single CTA, ragged last batch
                                            (64, 130, 70, 64, 2, "sigmoid"),
                                            (48, 13, 5000, 16, 1, "sigmoid")])   # 628 hidden units per CTA > 512 threads
def test_persistent_epoch_equals_tiled_path_and_oracle(N, V, H, bs, k, act):
    rng = np.random.default_rng(N + V)
    X = (rng.integers(0, 17, (N, V)) / 16.0).astype(np.float32)
    W = rng.uniform(-0.2, 0.2, (H, V)) / np.sqrt(V) if act == "relu" else rng.standard_normal((H, V)) / np.sqrt(V)
    b, c = np.full(V, 0.1) / np.sqrt(V), rng.standard_normal(H) * 0.05
    orders = [np.random.default_rng(5 + e).permutation(N) for e in range(2)]
    assert G.lib().dbn_rbm_persistent_plan(0, V, H, min(bs, N)) > 0
    (Wp, bp, cp), rp = train(X, W, b, c, act, bs, k, 0.05, 2, True, orders=orders)
    (Wt, bt, ct), rt = train(X, W, b, c, act, bs, k, 0.05, 2, False, orders=orders)
    assert rp.launches_per_epoch == 1 and rt.launches_per_epoch > 1
    # same Philox counters -> same samples; only the fp32 summation order differs between the two paths
    upd = np.linalg.norm(Wt - W)
    assert np.linalg.norm(Wp - Wt) < 2e-5 * upd + 1e-7, (np.linalg.norm(Wp - Wt), upd)
    assert np.max(np.abs(bp - bt)) < 2e-5 and np.max(np.abs(cp - ct)) < 2e-5
    # oracle on the regenerated Philox stream (float64), epoch by epoch
    Wo, bo, co = W.copy(), b.copy(), c.copy()
    for e in range(2):
        data = X[orders[e]].astype(np.float64)
        U = np.stack([G.philox_uniforms(77, t, e, 0, N, H) for t in range(k)], axis=1).astype(np.float64)
        for s in range(0, N, bs):
            orc.cd_step(Wo, bo, co, data[s:s + bs], U[s:s + bs], k, act, 0.05, bs)
    # trajectories may differ by a sample flip where |u - p| is at fp32 rounding; bound the drift
    assert np.linalg.norm(Wp - Wo) < 2e-3 * np.linalg.norm(Wo - W) + 1e-6


def test_persistent_plan_limits():
    lib = G.lib()
    assert lib.dbn_rbm_persistent_plan(0, 64, 256, 32) == 8
    assert lib.dbn_rbm_persistent_plan(0, 13, 100, 16) == 2
    assert lib.dbn_rbm_persistent_plan(0, 40, 24, 16) == 1
    assert lib.dbn_rbm_persistent_plan(0, 784, 500, 256) == 0      # config 2 does not fit on chip
    assert lib.dbn_rbm_persistent_plan(1, 64, 256, 32) == 0        # BF16 mode uses the tensor-core path
