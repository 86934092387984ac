"""Throughput of the small configurations (BASELINE.json configs 1 and 4) with the on-chip persistent
kernel on and off (debug / notes; not the headline bench)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dbn_b200 import engine as E  # noqa: E402

dev = torch.device("cuda")
rng = np.random.default_rng(0)
for name, N, V, H, bs, k, act in [("c1 layer1 64->256 bs32 CD-1", 1437, 64, 256, 32, 1, "relu"),
                                  ("c1 layer2 256->256 bs32 CD-1", 1437, 256, 256, 32, 1, "relu"),
                                  ("c4 13->100 bs16 CD-10", 4096, 13, 100, 16, 10, "relu")]:
    X = rng.random((N, V)).astype(np.float32)
    for persistent in (True, False):
        os.environ["DBN_B200_PERSISTENT"] = "1" if persistent else "0"
        layer = E.DeviceRbm(rng.uniform(-0.2, 0.2, (H, V)) / np.sqrt(V), np.zeros(V), np.zeros(H), act, "fp32", dev)
        r = E.RbmEpochRunner(layer, torch.from_numpy(X).to(dev), bs, k, 0.01, 1)
        for _ in range(3):
            r.run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            r.run()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("%-30s persistent=%-5s %10.0f samples/s  (%.2f ms/epoch, %.1f us/step)"
              % (name, persistent, N * reps / dt, dt / reps * 1e3, dt / reps / np.ceil(N / bs) * 1e6), flush=True)
