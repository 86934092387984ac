"""Debug: one fine-tune iteration through the persistent kernel vs the tiled path, error per quantity."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_harness as G

def net(rng, n_in, hs, n_out):
    layers, prev = [], n_in
    for h in hs:
        layers.append((rng.standard_normal((h, prev)) / np.sqrt(prev), rng.standard_normal(h) * 0.1)); prev = h
    return layers, rng.standard_normal((n_out, prev)) / np.sqrt(prev), rng.standard_normal(n_out) * 0.1

cases = [(96, [128], 3, 64, 64, "sigmoid", "linear"), (96, [128], 3, 64, 64, "sigmoid", "softmax"), (96, [128], 10, 64, 64, "sigmoid", "linear"),
         (96, [128], 3, 32, 32, "sigmoid", "linear"), (96, [256], 3, 64, 64, "sigmoid", "linear"),
         (64, [256, 256], 10, 32, 32, "sigmoid", "softmax"), (64, [128, 128], 10, 64, 64, "sigmoid", "softmax"),
         (64, [128, 128], 10, 64, 64, "relu", "softmax")]
for (n_in, hs, n_out, bs, n, act, head) in cases:
    rng = np.random.default_rng(1)
    layers, Wo, bo = net(rng, n_in, hs, n_out)
    X = rng.integers(0, 17, (n, n_in)) / 16.0
    Y = np.eye(n_out)[rng.integers(0, n_out, n)] if head == "softmax" else rng.standard_normal((n, n_out))
    r1 = G.run_mlp_epoch(layers, Wo, bo, X, Y, act, head, "bf16", bs, tc=True)
    r0 = G.run_mlp_epoch(layers, Wo, bo, X, Y, act, head, "bf16", bs, tc=False)
    print("case", n_in, hs, n_out, "bs", bs, "n", n, "launches", r1["launches"], r0["launches"])
    print("   loss   ", np.max(np.abs(r1["loss"] - r0["loss"])), "max", np.max(np.abs(r0["loss"])))
    print("   dWo    ", G.relF(r1["Wo"] - Wo, r0["Wo"] - Wo), " dbo ", G.relF(r1["bo"] - bo, r0["bo"] - bo), " |dWo_tc| %.3e |dWo_tiled| %.3e  |dbo_tc| %.3e |dbo_tiled| %.3e" % (
        np.linalg.norm(r1["Wo"] - Wo), np.linalg.norm(r0["Wo"] - Wo), np.linalg.norm(r1["bo"] - bo), np.linalg.norm(r0["bo"] - bo)))
    for l in range(len(hs) - 1, -1, -1):
        print("   dW[%d]  " % l, G.relF(r1["layers"][l][0] - layers[l][0], r0["layers"][l][0] - layers[l][0]),
              " dc ", G.relF(r1["layers"][l][1] - layers[l][1], r0["layers"][l][1] - layers[l][1]),
              " |dW_tc| %.3e |dW_tiled| %.3e" % (np.linalg.norm(r1["layers"][l][0] - layers[l][0]), np.linalg.norm(r0["layers"][l][0] - layers[l][0])))
