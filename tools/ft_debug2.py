"""Debug: one mini-batch through the persistent fine-tune kernel vs the tiled path, workspace buffer by buffer."""
import sys, os, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_harness as G
from dbn_b200 import _lib as L, engine as E

def align(x, a=1024): return (x + a - 1) // a * a
def pitch(c): return (c + 1 + 63) // 64 * 64

def run(layers, Wo, bo, X, Y, act, head, bs, tc):
    l = G.lib(); dev = torch.device("cuda"); n = X.shape[0]
    rbms = [E.DeviceRbm(W, np.zeros(W.shape[1]), c, act, "bf16", dev) for (W, c) in layers]
    mlp = E.DeviceMlp(rbms, Wo, bo, act, head, "bf16", dev)
    Xop = G.stage_operand(X, "bf16"); Yd = G._t(Y, torch.float32)
    ws = mlp.make_workspace(bs); loss = torch.zeros(n, device="cuda")
    a = mlp.step_args(Xop, Yd, 0.1 / bs, 1.0, 1.0, ws, 0, None, loss, None)
    prev = l.dbn_set_tc_finetune(1 if tc else 0)
    L.check(l.dbn_mlp_fit_epoch(E.stream_ptr(), C.byref(a), n, bs)); torch.cuda.synchronize()
    l.dbn_set_tc_finetune(prev)
    # workspace layout (mlp_ws_layout): a0, act[l], delta[l], Z, dO, dOb
    n_in = X.shape[1]; dims = [W.shape[0] for (W, _) in layers]; n_out = Wo.shape[0]
    off = 0; bufs = {}
    def carve(name, cols, b16):
        nonlocal off
        ld = pitch(cols) if b16 else cols
        nbytes = bs * ld * (2 if b16 else 4)
        t = ws[off:off + nbytes].view(torch.bfloat16 if b16 else torch.float32).view(bs, ld)
        bufs[name] = t[:, :cols + (1 if b16 else 0)].float().cpu().numpy()
        off = align(off + nbytes)
    carve("a0", n_in, True)
    for i, d in enumerate(dims): carve("act%d" % i, d, True)
    for i, d in enumerate(dims): carve("delta%d" % i, d, True)
    carve("Z", n_out, False); carve("dO", n_out, False); carve("dOb", n_out, True)
    return bufs

for (n_in, hs, n_out, bs, act, head) in [(96, [128], 3, 64, "sigmoid", "linear"), (64, [128, 128], 10, 64, "sigmoid", "softmax")]:
    rng = np.random.default_rng(1)
    layers, prev = [], n_in
    for h in hs:
        layers.append((rng.standard_normal((h, prev)) / np.sqrt(prev), rng.standard_normal(h) * 0.1)); prev = h
    Wo = rng.standard_normal((n_out, prev)) / np.sqrt(prev); bo = rng.standard_normal(n_out) * 0.1
    X = rng.integers(0, 17, (bs, n_in)) / 16.0
    Y = np.eye(n_out)[rng.integers(0, n_out, bs)] if head == "softmax" else rng.standard_normal((bs, n_out))
    b1 = run(layers, Wo, bo, X, Y, act, head, bs, True)
    b0 = run(layers, Wo, bo, X, Y, act, head, bs, False)
    print("case", n_in, hs, n_out, bs, head)
    for k in b1:
        d = np.abs(b1[k] - b0[k])
        print("   %-8s max|tc - tiled| %.3e   max|tiled| %.3e   rows differing: %s   last col tc %s tiled %s" % (
            k, d.max(), np.abs(b0[k]).max(), np.nonzero(d.max(axis=1) > 1e-3)[0][:8], b1[k][:2, -1], b0[k][:2, -1]))

# which formula did the persistent kernel's delta0 follow?
n_in, hs, n_out, bs, act, head = 64, [128, 128], 10, 64, "sigmoid", "softmax"
rng = np.random.default_rng(1)
layers, prev = [], n_in
for h in hs:
    layers.append((rng.standard_normal((h, prev)) / np.sqrt(prev), rng.standard_normal(h) * 0.1)); prev = h
Wo = rng.standard_normal((n_out, prev)) / np.sqrt(prev); bo = rng.standard_normal(n_out) * 0.1
X = rng.integers(0, 17, (bs, n_in)) / 16.0
Y = np.eye(n_out)[rng.integers(0, n_out, bs)]
b1 = run(layers, Wo, bo, X, Y, act, head, bs, True)
W1 = G.shadow_round(layers[1][0], "bf16")
d1, a0, a1 = b1["delta1"][:, :128].astype(np.float64), b1["act0"][:, :128].astype(np.float64), b1["act1"][:, :128].astype(np.float64)
got = b1["delta0"][:, :128].astype(np.float64)
cands = {"(d1 W1) * a0(1-a0)": (d1 @ W1) * a0 * (1 - a0), "(d1 W1)": d1 @ W1, "(d1 W1^T)*prime": (d1 @ W1.T) * a0 * (1 - a0),
         "(d1[:, :64] W1[:64]) * prime": (d1[:, :64] @ W1[:64]) * a0 * (1 - a0), "(d1[:, 64:] W1[64:]) * prime": (d1[:, 64:] @ W1[64:]) * a0 * (1 - a0),
         "(a1 W1) * prime": (a1 @ W1) * a0 * (1 - a0)}
for k, v in cands.items():
    print("   %-34s relF %.3e" % (k, np.linalg.norm(got - v) / np.linalg.norm(v)))
print("   got[0,:8]", got[0, :8]); print("   exp[0,:8]", cands["(d1 W1) * a0(1-a0)"][0, :8])
print("   got[40,:8]", got[40, :8]); print("   exp[40,:8]", cands["(d1 W1) * a0(1-a0)"][40, :8])
print("rows 32..63 only:")
cands["(d1 W1) no prime half0"] = d1[:, :64] @ W1[:64]
cands["(d1 W1) no prime half1"] = d1[:, 64:] @ W1[64:]
for k, v in cands.items():
    print("   %-34s relF %.3e" % (k, np.linalg.norm(got[32:] - v[32:]) / np.linalg.norm(v[32:])))
z1 = np.log(a1 / (1 - a1))
print("   corr(got[0], col sums)?  got[0,:4]", got[0, :4], " sum_rows dOb*a1 cols:", (b1["dOb"][:, :10].astype(np.float64).T @ a1)[:3, :4])
