"""Persistent fine-tune kernel and tiled path against the float64 oracle as the number of mini-batches grows
(config-3 widths).  Prints ||device - oracle|| / yardstick for each path, and the distance between the two paths."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from gpu_harness import run_mlp_epoch
from test_gpu_tc_finetune import _oracle_epoch, _net

def relF(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))

rng = np.random.default_rng(3)
layers, Wo, bo = _net(rng, 784, [500, 500, 2000], 10)
Xf = (rng.random((60000, 784)) < 0.13).astype(np.float64)
Yf = np.eye(10)[rng.integers(0, 10, 60000)]
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
for nb in (1, 4, 6, 8, 12, 16, 32):
    n = nb * 256
    X, Y = Xf[:n], Yf[:n]
    lay, Wo2, bo2, losses, sw, sb = _oracle_epoch(layers, Wo, bo, X, Y, "sigmoid", "softmax", 0.1, 1.0, 256, prec)
    out = []
    rs = []
    for tc in (True, False):
        r = run_mlp_epoch(layers, Wo, bo, X, Y, "sigmoid", "softmax", prec, 256, tc=tc)
        rs.append(r)
        out.append([f"{np.linalg.norm(r['layers'][l][0] - lay[l][0]) / sw[l]:.1e}" for l in range(3)] + [f"{np.linalg.norm(r['Wo'] - Wo2) / sw[-1]:.1e}",
                   f"loss {np.max(np.abs(r['loss'] - losses)):.1e}"])
    print(nb, "tc", out[0], "tiled", out[1], "Wo upd", f"{np.linalg.norm(Wo2 - Wo):.3f}", "max loss", f"{losses.max():.2f}", flush=True)
