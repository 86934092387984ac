"""Fine-tune with dropout (config-3 widths, dropout_p 0.2, Philox): device time per iteration through the persistent
kernel and through the tiled path (MlpIterRunner, CUDA events around 5 iterations after 2 warm-up iterations)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import dbn_b200
from dbn_b200 import engine as E, _lib as L

rng = np.random.default_rng(0)
n, dims = 60000, [784, 500, 500, 2000]
dev = torch.device("cuda")
X = torch.from_numpy((rng.random((n, 784)) < 0.13).astype(np.float32)).to(dev)
Y = torch.from_numpy(np.eye(10, dtype=np.float32)[rng.integers(0, 10, n)]).to(dev)
lib = L.load()
for drop in (0.0, 0.2):
    for tc in (1, 0):
        lib.dbn_set_tc_finetune(tc)
        layers = [E.DeviceRbm(rng.standard_normal((dims[i + 1], dims[i])) / np.sqrt(dims[i]), np.zeros(dims[i]), np.zeros(dims[i + 1]),
                              "sigmoid", "bf16", dev) for i in range(3)]
        mlp = E.DeviceMlp(layers, rng.standard_normal((10, 2000)) / 45.0, np.zeros(10), "sigmoid", "softmax", "bf16", dev)
        r = E.MlpIterRunner(mlp, X, Y, 256, 0.01, 1.0, drop, 1234)
        for _ in range(2):
            r.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            r.run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("dropout %.1f  persistent kernel %d: %.2f ms per iteration = %.2f M samples/s (launches per iteration: %d)"
              % (drop, tc, ms, n / ms / 1e3, r.launches_per_iter), flush=True)
lib.dbn_set_tc_finetune(1)
