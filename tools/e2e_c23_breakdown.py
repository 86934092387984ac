"""Where the wall time of bench.py's end-to-end legs of config 2 (`UnsupervisedDBN.fit(X_host)`, one epoch per layer) and
config 3 (`SupervisedDBNClassification.fit(X_host, y)`, one back-propagation iteration, no pre-training epochs) goes:
host-side time inside each stage, without extra synchronisation and with a device synchronise after each stage."""
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


timed(M._DrawAhead, "get", "_DrawAhead.get")
timed(M, "_to_device_f32")
timed(M, "_to_device_rows")
timed(E, "upload_f32")
timed(E, "download_f64")
timed(E.DeviceRbm, "__init__", "DeviceRbm.__init__")
timed(E.DeviceRbm, "fit", "DeviceRbm.fit")
timed(E.DeviceRbm, "transform", "DeviceRbm.transform")
timed(E.DeviceRbm, "export", "DeviceRbm.export")
timed(E.DeviceMlp, "__init__", "DeviceMlp.__init__")
timed(E.DeviceMlp, "fit", "DeviceMlp.fit")
timed(E.DeviceMlp, "export_head", "DeviceMlp.export_head")
timed(M.BinaryRBM, "_adopt")
timed(M.BinaryRBM, "_draw_params")
timed(M.UnsupervisedDBN, "_fit_device", "UnsupervisedDBN._fit_device")
timed(M.AbstractSupervisedDBN, "_fine_tuning", "_fine_tuning")
timed(M.AbstractSupervisedDBN, "_device_mlp", "_device_mlp")
timed(M.SupervisedDBNClassification, "_transform_labels_to_network_format", "labels -> one-hot")
timed(M.SupervisedDBNClassification, "_determine_num_output_neurons", "count classes")

rng = np.random.default_rng(0)
Xh = torch.empty(60000, 784, dtype=torch.float32).pin_memory()
Xh.numpy()[:] = (rng.random((60000, 784)) < 0.13)
y = rng.integers(0, 10, 60000)


def fit_c2():
    np.random.seed(5)
    dbn_b200.UnsupervisedDBN(hidden_layers_structure=[500, 500, 2000], learning_rate_rbm=0.01, n_epochs_rbm=1, batch_size=256,
                             verbose=False).fit(Xh.numpy())


def fit_c3():
    np.random.seed(5)
    dbn_b200.SupervisedDBNClassification(hidden_layers_structure=[500, 500, 2000], learning_rate=0.1, n_epochs_rbm=0,
                                         n_iter_backprop=1, batch_size=256, verbose=False).fit(Xh.numpy(), y)


for name, fit in (("c2", fit_c2), ("c3", fit_c3)):
    for sync in (False, True):
        SYNC[0] = sync
        fit(); fit()
        acc.clear()
        reps = 5
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fit()
        torch.cuda.synchronize()
        tot = (time.perf_counter() - t0) / reps * 1e3
        print("%s, sync after each stage %s: %.1f ms per fit (%.2f M samples/s)" % (name, sync, tot, 60000 / tot / 1e3))
        for k, v in acc.items():
            print("   %-50s %7.1f ms" % (k, v / reps * 1e3))
