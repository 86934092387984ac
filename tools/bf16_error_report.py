"""Honest error of the two tensor-core modes (bf16 / fp16 operands) against the float64 oracle on IDENTICAL (float64, un-rounded) inputs,
term by term, for the shapes of BASELINE.json -- the numbers behind the 'bf16-float64-weights' parity cases
(profiles/r02_bf16_error_report.txt).  Tolerance of the north star for this mode: 2e-3."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_harness as G  # noqa: E402
from oracle import dbn_oracle as orc  # noqa: E402


def report(B, V, H, density, round_w, mode="bf16"):
    dt16 = torch.bfloat16 if mode == "bf16" else torch.float16
    rng = np.random.default_rng(B + V + H)
    W = rng.standard_normal((H, V)) / np.sqrt(V)
    if round_w:
        W = torch.as_tensor(W).to(dt16).double().numpy()
    b, c = rng.standard_normal(V) * 0.1, rng.standard_normal(H) * 0.1
    V0 = (rng.random((B, V)) < density).astype(np.float64)
    U = rng.random((B, 1, H), dtype=np.float32)
    U[:, 0, :] = G.guard_band_uniforms(U[:, 0, :], orc.hidden_means(W, c, V0, "sigmoid"), 2e-2, rng)
    res = G.run_cd_step(W, b, c, V0, 1, "sigmoid", mode, U=U, raw=True)
    dW, db, dc, ch = orc.cd_deltas(W, b, c, V0, U.astype(np.float64), 1, "sigmoid")
    pos, neg = ch["P0"].T @ V0, ch["Pk"].T @ ch["Vk"]
    # the device's own operands: what the dW contraction would give in exact arithmetic on the device's P0 / Vk / Pk
    q = lambda a: torch.as_tensor(a).float().to(dt16).double().numpy()
    pos_dev, neg_dev = q(res["P0"]).T @ V0, q(res["Pk"]).T @ q(res["Vk"])
    n = np.linalg.norm
    print("mode=%s B=%d V=%d H=%d density=%.2f weights=%s" % (mode, B, V, H, density, "representable in the operand type" if round_w else "float64"))
    print("   samples bit-exact: %s" % np.array_equal(res["H0s"], ch["samples"][0]))
    print("   max|P0-ref| %.2e   max|Vk-ref| %.2e   max|Pk-ref| %.2e" %
          (np.max(np.abs(res["P0"] - ch["P0"])), np.max(np.abs(res["Vk"] - ch["Vk"])), np.max(np.abs(res["Pk"] - ch["Pk"]))))
    print("   |dW-ref|/|P0^T V0| %.2e   of which: positive half %.2e, negative half %.2e (exact products of the device's"
          " bf16 operands); contraction arithmetic itself %.2e" %
          (n(res["dW"] - dW) / n(pos), n(pos_dev - pos) / n(pos), n(neg_dev - neg) / n(pos),
           n(res["dW"] - (pos_dev - neg_dev)) / n(pos)))
    print("   max|db-ref|/B %.2e   max|dc-ref|/B %.2e" % (np.max(np.abs(res["db"] - db)) / B, np.max(np.abs(res["dc"] - dc)) / B))


if __name__ == "__main__":
    for shape in [(256, 784, 500, 0.13), (96, 500, 2000, 0.13), (512, 2300, 2100, 0.13), (512, 4096, 4096, 0.5)]:
        for mode in ("bf16", "fp16"):
            for round_w in (True, False):
                report(*shape, round_w, mode)
