"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): one CD-1 step in both precisions, one
fine-tune step, one persistent-kernel epoch.  Keep it tiny: the sanitizer slows kernels 10-100x."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import gpu_harness as G
from dbn_b200 import engine as E

rng = np.random.default_rng(0)
for prec in ("fp32", "bf16"):
    B, V, H = 160, 200, 136
    W = rng.standard_normal((H, V)) / np.sqrt(V)
    r = G.run_cd_step(W, np.zeros(V), np.zeros(H), (rng.random((B, V)) < 0.3).astype(np.float64), 2, "sigmoid", prec,
                      seed=3, raw=False, lr_over_bs=0.001)
    assert np.all(np.isfinite(r["W"]))
    layers = [(rng.standard_normal((130, V)) / np.sqrt(V), np.zeros(130)), (rng.standard_normal((70, 130)) / 11.0, np.zeros(70))]
    Y = np.eye(5)[rng.integers(0, 5, B)]
    m = G.run_mlp_step(layers, rng.standard_normal((5, 70)) / 8.0, np.zeros(5), rng.random((B, V)), Y, "sigmoid", "softmax",
                       prec, keep_p=0.8, lr_over_bs=0.001, decay=0.999)
    assert np.all(np.isfinite(m["Wo"]))
dev = torch.device("cuda")
X = rng.random((101, 40)).astype(np.float32)
layer = E.DeviceRbm(rng.standard_normal((24, 40)) / 6.0, np.zeros(40), np.zeros(24), "sigmoid", "fp32", dev)
r = E.RbmEpochRunner(layer, torch.from_numpy(X).to(dev), 16, 3, 0.05, 5)
assert r.persistent
r.run()
X2 = rng.random((96, 64)).astype(np.float32)
layer2 = E.DeviceRbm(rng.uniform(-0.2, 0.2, (256, 64)) / 8.0, np.zeros(64), np.zeros(256), "relu", "fp32", dev)
r2 = E.RbmEpochRunner(layer2, torch.from_numpy(X2).to(dev), 32, 1, 0.05, 5)   # cluster of 8, DSMEM exchange
r2.run()
torch.cuda.synchronize()
print("sanitize_case ok")
