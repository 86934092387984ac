"""End-to-end BinaryRBM.fit(X_host) on config 5 (131072 x 4096 rows, one epoch, batch 32768) with the dataset going up
as 16-bit operand rows rounded on the host (engine.upload_operand) and as fp32 rows; plus the pieces: host rounding
alone for several thread counts, and the H2D of each form."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import dbn_b200
from dbn_b200 import BinaryRBM, engine as E, _hoststage as hs

n, V, H, bs = 131072, 4096, 4096, 32768
rng = np.random.default_rng(0)
Xp = torch.empty(n, V, dtype=torch.float32).pin_memory()
Xp.numpy()[:] = (rng.random((n, V), dtype=np.float32) < 0.1)
X = Xp.numpy()
dev = torch.device("cuda")
torch.cuda.synchronize()

def fit():
    np.random.seed(1)
    t = time.perf_counter()
    BinaryRBM(n_hidden_units=H, learning_rate=0.01, n_epochs=1, batch_size=bs, verbose=False).fit(X)
    torch.cuda.synchronize()
    return (time.perf_counter() - t) * 1e3

for mode in ("1", "0", "1", "0"):
    os.environ["DBN_B200_OPERAND_UPLOAD"] = mode
    ts = [fit() for _ in range(4)]
    print("operand upload", mode, "fit ms:", " ".join("%.1f" % t for t in ts), " -> %.2f M samples/s" % (n / min(ts) / 1e3), flush=True)

pitch = E.bf16_pitch(V)
out = torch.empty(n, pitch, dtype=torch.float16).pin_memory()
for th in (1, 4, 8, 12, 15):
    ts = []
    for _ in range(3):
        t = time.perf_counter(); hs.stage_rows(X, None, out.data_ptr(), pitch, hs.F16, True, threads=th); ts.append((time.perf_counter() - t) * 1e3)
    print("host rounding, %2d threads: %.1f ms (%.1f GB/s read)" % (th, min(ts), n * V * 4 / min(ts) / 1e6), flush=True)
d16 = torch.empty(n, pitch, dtype=torch.float16, device=dev); d32 = torch.empty(n, V, dtype=torch.float32, device=dev)
for name, src, dst in (("16-bit rows", out, d16), ("fp32 rows", Xp, d32)):
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t = time.perf_counter(); dst.copy_(src, non_blocking=True); torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
    print("H2D %s: %.1f ms (%.1f GB/s)" % (name, min(ts), src.numel() * src.element_size() / min(ts) / 1e6), flush=True)
for piece in (4, 16, 64):
    E._OPERAND_PIECE = piece << 20
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t = time.perf_counter(); E.upload_operand(X, dev, "fp16"); torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
    print("upload_operand, %d MB pieces: %.1f ms" % (piece, min(ts)), flush=True)
