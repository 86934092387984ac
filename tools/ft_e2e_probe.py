"""Debug: wall time of SupervisedDBNClassification.fit (one back-prop iteration) with / without the persistent kernel."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dbn_b200
from dbn_b200 import _lib as L
rng = np.random.default_rng(0)
Xh = torch.empty(60000, 784, dtype=torch.float32).pin_memory()
Xh.numpy()[:] = (rng.random((60000, 784)) < 0.13)
y = rng.integers(0, 10, 60000)
lib = L.load()
for tc in (1, 0, 1):
    lib.dbn_set_tc_finetune(tc)
    ts = []
    for i in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m = dbn_b200.SupervisedDBNClassification(hidden_layers_structure=[500, 500, 2000], learning_rate=0.1, n_epochs_rbm=0,
                                                  n_iter_backprop=1, batch_size=256, verbose=False, precision="auto")
        m.fit(Xh.numpy(), y)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print("tc_finetune=%d  ms per fit: %s" % (tc, " ".join("%.1f" % t for t in ts)))
