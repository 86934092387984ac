"""Where the wall time of one end-to-end `UnsupervisedDBN.fit(X_host)` (bench.py's e2e leg, config 2) goes:
host-side time spent inside each stage of the fit, measured without adding synchronisations (debug tool)."""
import collections
import functools
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import dbn_b200
from dbn_b200 import engine as E, models as M

acc = collections.OrderedDict()


def timed(owner, name, label=None):
    fn = getattr(owner, name)
    label = label or "%s.%s" % (getattr(owner, "__name__", str(owner)), name)

    @functools.wraps(fn)
    def wrap(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            acc[label] = acc.get(label, 0.0) + time.perf_counter() - t
    setattr(owner, name, wrap)


timed(M, "_to_device_f32")
timed(M.BinaryRBM, "_init_params")
timed(M.BinaryRBM, "_device_layer")
timed(M, "_draw_seed")
timed(E.RbmEpochRunner, "__init__", "RbmEpochRunner.__init__")
timed(E, "epoch_order")
timed(E._IndexFeeder, "__init__", "_IndexFeeder.__init__")
timed(E._IndexFeeder, "push", "_IndexFeeder.push")
timed(E.RbmEpochRunner, "_enqueue", "RbmEpochRunner._enqueue (stage + fit_epoch launches)")
timed(E.DeviceRbm, "stage", "DeviceRbm.stage")
timed(E.DeviceRbm, "fit", "DeviceRbm.fit (total incl. final sync)")
timed(E.DeviceRbm, "export", "DeviceRbm.export")
timed(E.DeviceRbm, "transform", "DeviceRbm.transform")
timed(E.DeviceRbm, "make_workspace", "DeviceRbm.make_workspace")
timed(E.DeviceRbm, "alloc_staged", "DeviceRbm.alloc_staged")

rng = np.random.default_rng(0)
Xh = torch.empty(60000, 784, dtype=torch.float32).pin_memory()
Xh.numpy()[:] = (rng.random((60000, 784)) < 0.13)


def fit():
    m = dbn_b200.UnsupervisedDBN(hidden_layers_structure=[500, 500, 2000], learning_rate_rbm=0.01, n_epochs_rbm=1,
                                 batch_size=256, verbose=False, precision="bf16")
    m.fit(Xh.numpy())
    return m


fit()
fit()
acc.clear()
reps = 5
t0 = time.perf_counter()
for _ in range(reps):
    fit()
wall = (time.perf_counter() - t0) / reps
print("wall per fit: %.2f ms  (%.0f samples/s)" % (wall * 1e3, 60000 / wall))
for k, v in acc.items():
    print("  %-60s %8.2f ms" % (k, v / reps * 1e3))
