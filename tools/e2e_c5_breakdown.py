"""Where the wall time of one end-to-end `BinaryRBM.fit(X_host)` of config 5 on ONE GPU goes (bench.py's e2e leg at
N = 1): host-side time inside each stage, first without extra synchronisation (what the host waits for), then with a
device synchronise after each stage (what each stage costs on its own).  Debug tool."""
import collections, functools, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import dbn_b200  # noqa: E402
from dbn_b200 import engine as E, models as M  # noqa: E402

acc = collections.OrderedDict()
SYNC = [False]


def timed(owner, name, label=None):
    fn = getattr(owner, name)
    label = label or "%s.%s" % (getattr(owner, "__name__", str(owner)), name)

    @functools.wraps(fn)
    def wrap(*a, **k):
        t = time.perf_counter()
        try:
            r = fn(*a, **k)
            if SYNC[0]:
                torch.cuda.synchronize()
            return r
        finally:
            acc[label] = acc.get(label, 0.0) + time.perf_counter() - t
    setattr(owner, name, wrap)


timed(M._DrawAhead, "get", "_DrawAhead.get (waiting for the helper thread)")
timed(M, "_to_device_rows")
_real_upload = E.upload_f32
def _upload_logged(arr, device):
    t = time.perf_counter()
    r = _real_upload(arr, device)
    key = "upload_f32 of %s %s" % (np.asarray(arr).dtype, "x".join(map(str, np.shape(arr))))
    acc[key] = acc.get(key, 0.0) + time.perf_counter() - t
    return r
E.upload_f32 = _upload_logged
timed(E._PinnedPool, "take", "_PinnedPool.take")
timed(E, "download_f64")
timed(E.DeviceRbm, "__init__", "DeviceRbm.__init__ (upload W,b,c + shadow)")
timed(E.DeviceRbm, "fit", "DeviceRbm.fit (epoch: gather + steps, + final sync)")
timed(E.DeviceRbm, "export", "DeviceRbm.export")
timed(M.BinaryRBM, "_init_params")
timed(M.BinaryRBM, "_adopt")

V = H = 4096
bs, n = 32768, 32768 * 4
gen = torch.Generator(device="cpu").manual_seed(0)
Xh = torch.empty(n, V, dtype=torch.float32).pin_memory()
for s in range(0, n, 16384):
    Xh[s:s + 16384] = (torch.rand(16384, V, generator=gen) < 0.5).float()


def fit():
    np.random.seed(5)
    m = dbn_b200.BinaryRBM(n_hidden_units=H, learning_rate=0.01, n_epochs=1, batch_size=bs, verbose=False)
    m.fit(Xh.numpy())
    return m


for mode in os.environ.get("MODES", "0,1").split(","):
    os.environ["DBN_B200_OPERAND_UPLOAD"] = mode
    for sync in (False, True):
        SYNC[0] = sync
        fit()
        acc.clear()
        reps = 3
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fit()
        torch.cuda.synchronize()
        tot = (time.perf_counter() - t0) / reps * 1e3
        print("operand upload %s, sync after each stage %s: %.1f ms per fit (%.2f M samples/s)" % (mode, sync, tot, n / tot / 1e3))
        for k, v in acc.items():
            print("   %-60s %7.1f ms" % (k, v / reps * 1e3))
