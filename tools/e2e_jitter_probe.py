"""Wall time of 30 consecutive end-to-end fits of config 2 and config 3 (bench.py's e2e legs), one number per fit: is the
mean made by a few slow fits?  Also prints the pinned pool's size before and after."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import dbn_b200
from dbn_b200 import engine as E

rng = np.random.default_rng(0)
big = torch.empty(131072, 4096, dtype=torch.float32).pin_memory() if os.environ.get("BIG", "1") == "1" else None   # bench.py holds config 5's dataset too
Xh = torch.empty(60000, 784, dtype=torch.float32).pin_memory()
Xh.numpy()[:] = (rng.random((60000, 784)) < 0.13)
y = rng.integers(0, 10, 60000)


def c2():
    dbn_b200.UnsupervisedDBN(hidden_layers_structure=[500, 500, 2000], learning_rate_rbm=0.01, n_epochs_rbm=1, batch_size=256,
                             verbose=False).fit(Xh.numpy())


def c3():
    dbn_b200.SupervisedDBNClassification(hidden_layers_structure=[500, 500, 2000], learning_rate=0.1, n_epochs_rbm=0,
                                         n_iter_backprop=1, batch_size=256, verbose=False).fit(Xh.numpy(), y)


for name, fit in (("c2", c2), ("c3", c3)):
    ts = []
    print(name, "pool before:", len(E._PinnedPool._free))
    for i in range(30):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fit()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t) * 1e3)
    print(name, "ms per fit:", " ".join("%.0f" % t for t in ts))
    print(name, "mean %.1f  median %.1f  min %.1f  max %.1f   pool after: %d" % (np.mean(ts), np.median(ts), min(ts), max(ts), len(E._PinnedPool._free)))
