"""scikit-learn style estimators of the B200 backend -- the host-side mirror of the reference's
backend interface (`dbn/models.py`; the TF backend `dbn/tensorflow/models.py` is the precedent for
what a backend overrides).  Same four classes, constructor arguments, defaults, fitted attributes,
method signatures and error behaviour; the training loops run on the device through
libdbn_b200.so (see engine.py) instead of per-sample NumPy.

Fitted parameters are exposed exactly like the reference's NumPy backend: `W (H,V)`, `b (V,)`,
`c (H,)` float64 ndarrays on the estimator (so pickles are plain NumPy and interoperate), head
`W (out,H_last)`, `b (out,)` on the supervised estimators.

Two extra, optional constructor arguments exist on every class (defaults keep the reference's
behaviour): `precision` in {'auto','fp32','bf16','fp16'} selects the strict-FP32 FFMA kernels or the
tcgen05 kernels with bf16 / fp16 operands ('auto': fp16 for sigmoid RBM pre-training, bf16 for relu layers and
the fine-tune, fp32 for layers too small for a tensor-core tile), and `rng` in {'philox','numpy'} selects in-kernel Philox draws or a replay of
the reference's `np.random` uniform / binomial stream (parity runs).
"""
from __future__ import annotations

import os
import pickle
import queue
import threading

import numpy as np
from scipy.stats import truncnorm
from sklearn.base import BaseEstimator, ClassifierMixin, RegressorMixin, TransformerMixin

from . import engine as E
from ._hostrng import legacy_randn, legacy_randn_f32
from .activations import ReLUActivationFunction, SigmoidActivationFunction


def _device():
    import torch
    E.require_device()
    return torch.device("cuda", torch.cuda.current_device())


def _to_device_f32(X):
    """Host array -> fp32 device tensor.  A pinned source is copied asynchronously (the caller goes on preparing
    the fit while the rows land); a pageable one goes through the driver's staging and blocks."""
    import torch
    X = np.ascontiguousarray(np.asarray(X), dtype=np.float32)
    t = torch.from_numpy(X)
    if t.is_pinned():
        return t.to(_device(), non_blocking=True)
    return t.to(_device())


def _to_device_rows(X, precision_of_first_layer, quiet):
    """The dataset a pre-training run reads, on the device.  When its only consumer is a tensor-core layer (which reads
    the rows through a 16-bit rounding anyway) and nothing else needs the un-rounded values (the verbose
    reconstruction error does), the rows are rounded on the host and go up as 16-bit operands: same numbers, half the
    PCIe bytes (engine.upload_operand), and a pageable source needs no driver staging.  Otherwise fp32 rows; a caller who
    page-locked a float32 dataset gets the single DMA straight from it (at 55 GB/s that beats rounding 2 GB on eight
    host threads: tools/e2e_c5_probe.py)."""
    X = np.asarray(X)
    if quiet and X.ndim == 2 and X.shape[0] > 0 and not E.is_pinned_f32(X):
        prec = precision_of_first_layer(int(X.shape[1]))
        if E.operand_upload_wanted(prec, int(X.shape[0]), int(X.shape[1])):
            return E.upload_operand(X, _device(), prec)
    return _to_device_f32(X)


def _resident(n, v, h):
    """Does a dataset of n rows fit the device for a layer v -> h: fp32 rows, their 16-bit epoch copy, and the fp32
    features handed to the next layer (engine.resident_budget_bytes; DBN_B200_RESIDENT_BYTES overrides)."""
    need = n * v * 4 + n * E.bf16_pitch(v) * 2 + n * h * 4
    return need <= E.resident_budget_bytes(_device())


def _draw_seed():
    """Philox key drawn from the global np.random stream, so `np.random.seed(...)` stays the
    user-facing reproducibility knob (reference: example_classification.py:3)."""
    return int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 32)


class _DrawAhead:
    """Everything a greedy pre-training run takes from the global `np.random` stream -- per layer the
    parameter initialisation (dbn/models.py:55-69), the Philox key, and one double shuffle per epoch
    (:106 + utils.py:12) -- is data independent.  A helper thread draws it, in EXACTLY the order the
    sequential code would (so a seeded fit gives the same model either way), while the caller uploads the
    dataset and keeps the GPU busy with the previous layer: legacy `randn` costs ~16 ns per weight, which
    at 784-[500,500,2000] is as long as the whole epoch takes on the device.  NumPy releases the GIL while
    it fills; the queue is bounded so a many-epoch fit never holds more than a few permutations."""

    def __init__(self, jobs, depth=4):
        self._q = queue.Queue(maxsize=depth)
        self._stop = False
        self._t = threading.Thread(target=self._run, args=(jobs,), daemon=True)
        self._t.start()

    def _put(self, item):
        while not self._stop:
            try:
                self._q.put(item, timeout=0.05)
                return True
            except queue.Full:
                pass
        return False

    def _run(self, jobs):
        try:
            for item in jobs:
                if not self._put(("ok", item)):
                    return
        except BaseException as e:   # surfaces in the caller, at the draw that failed
            self._put(("err", e))

    def get(self):
        kind, v = self._q.get()
        if kind == "err":
            raise v
        return v

    def close(self):
        """Stop drawing (after an error in the caller) and wait for the helper: nobody else touches the global
        stream while it runs."""
        self._stop = True
        while self._t.is_alive():
            try:
                self._q.get_nowait()
            except queue.Empty:
                pass
            self._t.join(timeout=0.01)


def _signature(*arrays):
    """Cheap fingerprint of host parameter arrays: identity, shape and a strided checksum.  Lets an estimator keep its
    fitted layers on the device between fit / transform / predict calls and still notice when a caller has replaced
    or edited `W`, `b`, `c` (they are plain ndarrays, as in the reference)."""
    sig = []
    for a in arrays:
        a = np.asarray(a)
        flat = a.reshape(-1)
        step = max(1, flat.size // 4096)
        sig.append((id(a), a.shape, a.dtype.str, float(flat[::step].sum()), float(flat[-1]) if flat.size else 0.0))
    return tuple(sig)


class BaseModel(object):
    """save / load by pickling the estimator (reference: dbn/models.py:11-23)."""

    _DEVICE_STATE = ("_dev_cache",)      # device-resident copies of the fitted parameters: never pickled / cloned

    def __getstate__(self):
        state = dict(self.__dict__)
        for k in self._DEVICE_STATE:
            state.pop(k, None)
        return state

    def save(self, save_path):
        with open(save_path, "wb") as fp:
            pickle.dump(self, fp)

    @classmethod
    def load(cls, load_path):
        with open(load_path, "rb") as fp:
            return pickle.load(fp)


class BinaryRBM(BaseEstimator, TransformerMixin, BaseModel):
    """Binary restricted Boltzmann machine trained with CD-k (reference: dbn/models.py:26-220)."""

    def __init__(self, n_hidden_units=100, activation_function='sigmoid', optimization_algorithm='sgd',
                 learning_rate=1e-3, n_epochs=10, contrastive_divergence_iter=1, batch_size=32, verbose=True,
                 precision='auto', rng='philox'):
        self.n_hidden_units = n_hidden_units
        self.activation_function = activation_function
        self.optimization_algorithm = optimization_algorithm
        self.learning_rate = learning_rate
        self.n_epochs = n_epochs
        self.contrastive_divergence_iter = contrastive_divergence_iter
        self.batch_size = batch_size
        self.verbose = verbose
        self.precision = precision
        self.rng = rng

    # ---- parameter initialisation: same draws, same order as dbn/models.py:55-69 ----
    def _draw_params(self, n_visible, for_upload=False):
        """Initial parameters (dbn/models.py:55-69).  for_upload: the weights are about to be uploaded and trained, so
        nobody will read them on the host -- a large sigmoid layer is then drawn straight as the fp32 the device stores,
        into pooled pinned memory (same np.random draws, each value rounded as the upload would round it): no float64
        array of the layer's size, no divide pass, no narrowing pass on the critical path of the fit."""
        h, s = self.n_hidden_units, np.sqrt(n_visible)
        if self.activation_function == 'sigmoid':   # legacy_randn == np.random.randn, bit for bit (see _hostrng.py)
            W = E.pooled_f32((h, n_visible)) if (for_upload and h * n_visible >= (1 << 17)) else None
            if W is not None:
                legacy_randn_f32(W, s)
            else:
                W = legacy_randn(h, n_visible) / s
            c = legacy_randn(h) / s
            b = legacy_randn(n_visible) / s
            return W, c, b, SigmoidActivationFunction
        if self.activation_function == 'relu':
            W = truncnorm.rvs(-0.2, 0.2, size=[h, n_visible]) / s
            return W, np.full(h, 0.1) / s, np.full(n_visible, 0.1) / s, ReLUActivationFunction
        raise ValueError("Invalid activation function.")

    def _init_params(self, n_visible, drawn=None):
        self.n_visible_units = n_visible
        self.W, self.c, self.b, self._activation_function_class = drawn if drawn is not None else self._draw_params(n_visible)

    def _upload_initial(self):
        """The freshly initialised layer on the device.  Weights that were drawn into pooled pinned memory
        (_draw_params(for_upload=True)) are handed back once their upload is queued; `W` is then unset until the fitted
        parameters are adopted."""
        layer = self._device_layer()
        if E.release_pooled(self.W):
            self.W = None
        return layer

    def _resolved_precision(self):
        return self._precision_for(self.n_visible_units)

    def _precision_for(self, n_visible):
        return E.resolve_precision(self.precision, n_visible, self.n_hidden_units, self.batch_size,
                                   activation=self.activation_function, phase="pretrain")

    def _device_layer(self, precision=None):
        """The layer on the device.  The fitted DeviceRbm is kept between calls (no re-upload, no re-cast of the
        shadow) for as long as W, b, c are the arrays it was built from."""
        precision = precision or self._resolved_precision()
        sig = _signature(self.W, self.b, self.c)
        cache = self.__dict__.setdefault("_dev_cache", {})
        hit = cache.get(precision)
        if hit is not None and hit[0] == sig:
            return hit[1]
        layer = E.DeviceRbm(self.W, self.b, self.c, self.activation_function, precision, _device())
        cache[precision] = (sig, layer)
        return layer

    def _adopt(self, layer):
        """Fitted parameters from the device; the device copy stays cached for transform / predict."""
        self.W, self.b, self.c = layer.export()
        self._dev_cache = {layer.prec_name: (_signature(self.W, self.b, self.c), layer)}

    def _fit_device(self, Xd, draws=None, defer_export=False):
        """Initialise and train on a device-resident fp32 dataset; returns the DeviceRbm so a
        stacked DBN can keep transforming on the device.  `draws` (a _DrawAhead) supplies, in order, the
        initial parameters, the Philox key and one shuffle per epoch instead of drawing them here;
        `defer_export` leaves the layer's work queued on the stream and the fitted parameters on the device
        (the caller assigns `W, b, c = layer.export()` later), so the host can go on to the next layer."""
        self._init_params(int(Xd.shape[1]), draws.get() if draws is not None else None)
        if self.optimization_algorithm != 'sgd':
            raise ValueError("Invalid optimization algorithm.")
        layer = self._upload_initial()
        if draws is not None:
            seed = draws.get()
        else:
            seed = _draw_seed() if self.rng == 'philox' else 0   # 'numpy' replays the reference stream untouched
        cb = None
        if self.verbose:
            cb = lambda it, err: print(">> Epoch %d finished \tRBM Reconstruction error %f" % (it, err))
        layer.fit(Xd, self.n_epochs, self.batch_size, self.contrastive_divergence_iter, self.learning_rate, seed,
                  rng_mode=self.rng, verbose=cb, orders=draws, sync=not defer_export)
        if not defer_export:
            self._adopt(layer)
        return layer

    def _draw_jobs(self, n, n_features):
        """The np.random draws of one fit, in sequential order (generator for _DrawAhead): parameters, Philox key,
        one double shuffle per epoch."""
        yield self._draw_params(n_features, for_upload=True)   # read back from the device even when nothing trains
        yield _draw_seed()
        for _ in range(self.n_epochs if n > 0 else 0):
            yield E.epoch_order(n)

    def _fit_data_parallel(self, X, dp):
        """One process per GPU (torchrun, DBN_B200_DATA_PARALLEL=1), same X on every rank.  Rank 0 alone consumes
        np.random -- parameters, Philox key and every epoch's shuffle are BROADCAST, so replicas start identical and
        agree on the rows of every mini-batch whatever the other ranks' generator state is."""
        import torch
        import torch.distributed as dist
        rank, world = dp
        if self.rng != 'philox':
            raise ValueError("rng='numpy' replays one np.random stream sample by sample and cannot be data-parallel.")
        if self.activation_function not in ('sigmoid', 'relu'):
            raise ValueError("Invalid activation function.")
        if self.optimization_algorithm != 'sgd':
            raise ValueError("Invalid optimization algorithm.")
        X = np.ascontiguousarray(np.asarray(X), dtype=np.float32)
        n, v = int(X.shape[0]), int(X.shape[1])
        E.dp_check_shapes(n, self.batch_size, world)
        dev = _device()
        self.n_visible_units = v
        self._activation_function_class = (SigmoidActivationFunction if self.activation_function == 'sigmoid'
                                           else ReLUActivationFunction)
        precision = self._resolved_precision()
        draws = None
        seed_t = torch.zeros(1, dtype=torch.int64, device=dev)
        # this rank's rows start crossing PCIe now, under rank 0's parameter draws and the broadcast everyone waits for
        Xl = E.dp_begin_upload(X, precision, self.n_hidden_units, dp, dev, quiet=not self.verbose)
        try:
            if rank == 0:
                draws = _DrawAhead(self._draw_jobs(n, v))      # drawn while the rows upload
                W, c, b, _cls = draws.get()
                layer = E.DeviceRbm(W, b, c, self.activation_function, precision, dev)
                E.release_pooled(W, dev)
                seed = draws.get()
                seed_t.fill_(seed - (1 << 64) if seed >= (1 << 63) else seed)
            else:
                layer = E.DeviceRbm.blank(self.n_hidden_units, v, self.activation_function, precision, dev)
            for t in (layer.W, layer.b, layer.c, seed_t):
                dist.broadcast(t, 0)
            if rank != 0:
                layer.refresh_shadow()
            seed = int(seed_t.item()) & 0xFFFFFFFFFFFFFFFF
            cb = None
            if self.verbose and rank == 0:
                cb = lambda it, err: print(">> Epoch %d finished \tRBM Reconstruction error %f" % (it, err))
            elif self.verbose:
                cb = lambda it, err: None
            E.fit_rbm_data_parallel(layer, X, self.n_epochs, self.batch_size, self.contrastive_divergence_iter,
                                    self.learning_rate, seed, dp, (draws.get if draws is not None else None), verbose=cb,
                                    Xl=Xl)
        finally:
            if draws is not None:
                draws.close()
        self._adopt(layer)
        return layer

    def _fit_host(self, X, draws=None):
        """Train on a dataset that stays on the HOST (engine.HostRbmEpochRunner: permuted chunks are gathered into
        pinned memory and uploaded while the previous chunk trains).  Same draws, same Philox counters as the resident
        path, hence bit-identical parameters.  Returns (DeviceRbm, runner)."""
        if self.rng != 'philox':
            raise ValueError("rng='numpy' (replay of the reference's uniform stream) needs the dataset resident on the device.")
        X = np.ascontiguousarray(np.asarray(X), dtype=np.float32)
        self._init_params(int(X.shape[1]), draws.get() if draws is not None else None)
        if self.optimization_algorithm != 'sgd':
            raise ValueError("Invalid optimization algorithm.")
        layer = self._upload_initial()
        seed = draws.get() if draws is not None else _draw_seed()
        runner = E.HostRbmEpochRunner(layer, X, self.batch_size, self.contrastive_divergence_iter, self.learning_rate, seed)
        import torch
        for epoch in range(1, self.n_epochs + 1):
            runner.run(draws.get() if draws is not None else None)
            if self.verbose:
                print(">> Epoch %d finished \tRBM Reconstruction error %f" % (epoch, runner.reconstruction_error()))
        torch.cuda.current_stream().synchronize()
        self._adopt(layer)
        return layer, runner

    def fit(self, X):
        """Fit the RBM to X (n_samples, n_features).  Returns self (dbn/models.py:49-75)."""
        dp = E.dp_context()
        if dp is not None:
            self._fit_data_parallel(X, dp)
            return self
        shape = np.shape(X)
        if len(shape) == 2 and shape[0] > 0 and self.rng == 'philox' and not _resident(shape[0], shape[1], self.n_hidden_units):
            self._fit_host(X)       # out of core: the dataset is streamed from the host every epoch
            return self
        draws = None
        if self.rng == 'philox' and os.environ.get("DBN_B200_DRAW_AHEAD", "1") != "0":
            X = np.asarray(X)   # parameters and shuffles are drawn by a helper thread while the rows upload
            if X.ndim == 2 and X.shape[0] * self.n_hidden_units >= (1 << 16):
                draws = _DrawAhead(self._draw_jobs(int(X.shape[0]), int(X.shape[1])))
        try:
            self._fit_device(_to_device_rows(X, self._precision_for, not self.verbose), draws=draws)
        finally:
            if draws is not None:
                draws.close()
        return self

    def transform(self, X):
        """Hidden-unit means of X; a 1-D sample gives a 1-D result (dbn/models.py:77-86)."""
        X = np.asarray(X)
        single = X.ndim == 1
        Xd = _to_device_f32(X[None, :] if single else X)
        out = self._device_layer().transform(Xd).double().cpu().numpy()
        return out[0] if single else out

    def _reconstruct(self, transformed_data):
        """Visible means given hidden values (dbn/models.py:88-94)."""
        import torch
        layer = self._device_layer('fp32')
        Hd = _to_device_f32(transformed_data)
        out = torch.empty(Hd.shape[0], layer.V, dtype=torch.float32, device=Hd.device)
        layer.visible_means(Hd, int(Hd.shape[0]), out)
        return out.double().cpu().numpy()

    def _compute_reconstruction_error(self, data):
        """mean over samples of the squared reconstruction distance (dbn/models.py:212-220)."""
        return self._device_layer().reconstruction_error(_to_device_f32(data))


class UnsupervisedDBN(BaseEstimator, TransformerMixin, BaseModel):
    """Greedy layer-wise stack of RBMs (reference: dbn/models.py:223-287)."""

    def __init__(self, hidden_layers_structure=[100, 100], activation_function='sigmoid',
                 optimization_algorithm='sgd', learning_rate_rbm=1e-3, n_epochs_rbm=10,
                 contrastive_divergence_iter=1, batch_size=32, verbose=True, precision='auto', rng='philox'):
        self.hidden_layers_structure = hidden_layers_structure
        self.activation_function = activation_function
        self.optimization_algorithm = optimization_algorithm
        self.learning_rate_rbm = learning_rate_rbm
        self.n_epochs_rbm = n_epochs_rbm
        self.contrastive_divergence_iter = contrastive_divergence_iter
        self.batch_size = batch_size
        self.rbm_layers = None
        self.verbose = verbose
        self.precision = precision
        self.rng = rng
        self.rbm_class = BinaryRBM

    def _fit_device(self, Xd):
        self.rbm_layers = [
            self.rbm_class(n_hidden_units=h, activation_function=self.activation_function,
                           optimization_algorithm=self.optimization_algorithm, learning_rate=self.learning_rate_rbm,
                           n_epochs=self.n_epochs_rbm, contrastive_divergence_iter=self.contrastive_divergence_iter,
                           batch_size=self.batch_size, verbose=self.verbose, precision=self.precision, rng=self.rng)
            for h in self.hidden_layers_structure]
        if self.verbose:
            print("[START] Pre-training step:")
        data = Xd
        ahead = (self.rbm_class is BinaryRBM and self.rng == 'philox' and
                 os.environ.get("DBN_B200_DRAW_AHEAD", "1") != "0")
        if not ahead:
            for i, rbm in enumerate(self.rbm_layers):      # strictly sequential: layer l+1 sees layer l's final features
                layer = rbm._fit_device(data)
                if i + 1 < len(self.rbm_layers):           # (the reference also transforms after the last layer, for nobody)
                    data = layer.transform(data)           # stays in HBM between layers
        else:
            # same draws in the same order, made by a helper thread that runs ahead of the device; every
            # layer's launches are queued without waiting for the previous layer, parameters come back at the end
            draws = _DrawAhead(self._draw_jobs(int(Xd.shape[0]), int(Xd.shape[1])))
            try:
                fitted = []
                for i, rbm in enumerate(self.rbm_layers):
                    layer = rbm._fit_device(data, draws=draws, defer_export=True)
                    fitted.append((rbm, layer))
                    if i + 1 < len(self.rbm_layers):
                        data = layer.transform(data)
                for rbm, layer in fitted:
                    rbm._adopt(layer)
            finally:
                draws.close()
        if self.verbose:
            print("[END] Pre-training step")
        return self

    def _draw_jobs(self, n, n_features):
        """The np.random draws of one greedy pre-training run, in sequential order (generator for _DrawAhead)."""
        v = n_features
        for rbm in self.rbm_layers:
            yield rbm._draw_params(v, for_upload=True)
            yield _draw_seed()
            for _ in range(rbm.n_epochs if n > 0 else 0):
                yield E.epoch_order(n)
            v = rbm.n_hidden_units

    def _fit_host(self, X):
        """Greedy pre-training with every layer's input on the host (out of core): layer l trains from host chunks,
        its features are computed chunk-wise back into host memory and become layer l+1's dataset."""
        self.rbm_layers = [
            self.rbm_class(n_hidden_units=h, activation_function=self.activation_function,
                           optimization_algorithm=self.optimization_algorithm, learning_rate=self.learning_rate_rbm,
                           n_epochs=self.n_epochs_rbm, contrastive_divergence_iter=self.contrastive_divergence_iter,
                           batch_size=self.batch_size, verbose=self.verbose, precision=self.precision, rng=self.rng)
            for h in self.hidden_layers_structure]
        if self.verbose:
            print("[START] Pre-training step:")
        data = np.ascontiguousarray(np.asarray(X), dtype=np.float32)
        for i, rbm in enumerate(self.rbm_layers):
            _layer, runner = rbm._fit_host(data)
            if i + 1 < len(self.rbm_layers):
                data = runner.transform_to_host()
        if self.verbose:
            print("[END] Pre-training step")
        return self

    def _out_of_core(self, shape):
        if len(shape) != 2 or shape[0] == 0 or self.rng != 'philox' or self.rbm_class is not BinaryRBM:
            return False
        widths = [shape[1]] + list(self.hidden_layers_structure)
        return not all(_resident(shape[0], v, h) for v, h in zip(widths[:-1], widths[1:]))

    def fit(self, X, y=None):
        """Pre-train every layer on the previous layer's features (dbn/models.py:248-276)."""
        if self._out_of_core(np.shape(X)):
            return self._fit_host(X)
        first = lambda v: E.resolve_precision(self.precision, v, self.hidden_layers_structure[0], self.batch_size,
                                              activation=self.activation_function, phase="pretrain")
        quiet = not self.verbose and self.rbm_class is BinaryRBM and len(self.hidden_layers_structure) > 0
        return self._fit_device(_to_device_rows(X, first, quiet))

    def _transform_device(self, Xd):
        for rbm in self.rbm_layers:
            Xd = rbm._device_layer().transform(Xd)
        return Xd

    def transform(self, X):
        """Stacked transform (dbn/models.py:278-287)."""
        X = np.asarray(X)
        single = X.ndim == 1
        out = self._transform_device(_to_device_f32(X[None, :] if single else X)).double().cpu().numpy()
        return out[0] if single else out


class AbstractSupervisedDBN(BaseEstimator, BaseModel):
    """Pre-train + fine-tune driver shared by the classifier and the regressor
    (reference: dbn/models.py:290-382 and the NumPy specialisation :385-559)."""

    _head = None  # 'softmax' | 'linear'

    def __init__(self, hidden_layers_structure=[100, 100], activation_function='sigmoid',
                 optimization_algorithm='sgd', learning_rate=1e-3, learning_rate_rbm=1e-3, n_iter_backprop=100,
                 l2_regularization=1.0, n_epochs_rbm=10, contrastive_divergence_iter=1, batch_size=32, dropout_p=0,
                 verbose=True, precision='auto', rng='philox'):
        self.hidden_layers_structure = hidden_layers_structure
        self.activation_function = activation_function
        self.optimization_algorithm = optimization_algorithm
        self.learning_rate = learning_rate
        self.learning_rate_rbm = learning_rate_rbm
        self.n_iter_backprop = n_iter_backprop
        self.l2_regularization = l2_regularization
        self.n_epochs_rbm = n_epochs_rbm
        self.contrastive_divergence_iter = contrastive_divergence_iter
        self.batch_size = batch_size
        self.dropout_p = dropout_p
        self.verbose = verbose
        self.precision = precision
        self.rng = rng
        self.unsupervised_dbn_class = UnsupervisedDBN
        self.unsupervised_dbn = UnsupervisedDBN(
            hidden_layers_structure=hidden_layers_structure, activation_function=activation_function,
            optimization_algorithm=optimization_algorithm, learning_rate_rbm=learning_rate_rbm,
            n_epochs_rbm=n_epochs_rbm, contrastive_divergence_iter=contrastive_divergence_iter,
            batch_size=batch_size, verbose=verbose, precision=precision, rng=rng)
        self.p = 1 - self.dropout_p

    # ---- public API (dbn/models.py:327-362) ----
    def fit(self, X, y=None, pre_train=True):
        if self.unsupervised_dbn._out_of_core(np.shape(X)):
            # out of core: both phases stream the dataset from the host, chunk by chunk
            Xh = np.ascontiguousarray(np.asarray(X), dtype=np.float32)
            if pre_train:
                self.unsupervised_dbn._fit_host(Xh)
            self._fine_tuning(Xh, y)
            return self
        Xd = _to_device_f32(X)
        if pre_train:
            self.unsupervised_dbn._fit_device(Xd)
        self._fine_tuning(Xd, y)   # labels as given: the label map's keys are the caller's elements (dbn/models.py:575-584)
        return self

    def pre_train(self, X):
        self.unsupervised_dbn.fit(X)
        return self

    def transform(self, *args):
        return self.unsupervised_dbn.transform(*args)

    def predict(self, X):
        X = np.asarray(X)
        if X.ndim == 1:
            X = X[None, :]
        return self._forward(_to_device_f32(X)).double().cpu().numpy()

    def _compute_output_units_matrix(self, matrix_visible_units):
        """Head outputs for already-transformed features (dbn/models.py:607-615, :705-711), on the device."""
        A = _to_device_f32(np.atleast_2d(np.asarray(matrix_visible_units)))
        return self._cached_mlp().head_outputs(A).double().cpu().numpy()

    # ---- device plumbing ----
    def _ft_precision(self):
        rbms = self.unsupervised_dbn.rbm_layers
        widths = [rbms[0].n_visible_units] + [r.n_hidden_units for r in rbms]
        return E.resolve_precision(self.precision, min(widths), min(widths), self.batch_size,
                                   activation=self.unsupervised_dbn.activation_function, phase="finetune")

    def _device_mlp(self, precision):
        dev = _device()
        layers = []
        for r in self.unsupervised_dbn.rbm_layers:
            # a layer that was just pre-trained is still on the device (BinaryRBM._adopt): copy it there instead of
            # narrowing and uploading its float64 host arrays again
            hit = next((l for (sig, l) in r.__dict__.get("_dev_cache", {}).values()
                        if sig == _signature(r.W, r.b, r.c)), None)
            layers.append(hit.clone_as(precision) if hit is not None else
                          E.DeviceRbm(r.W, r.b, r.c, r.activation_function, precision, dev))
        return E.DeviceMlp(layers, self.W, self.b, self.unsupervised_dbn.activation_function, self._head,
                           precision, dev)

    def _cached_mlp(self):
        """The fitted network on the device, kept between predict calls while the host parameters are unchanged."""
        rbms = self.unsupervised_dbn.rbm_layers
        precision = self._ft_precision()
        arrays = [self.W, self.b]
        for r in rbms:
            arrays += [r.W, r.c]
        sig = (precision,) + _signature(*arrays)
        cache = getattr(self, "_dev_cache", None)
        if cache is not None and cache[0] == sig:
            return cache[1]
        mlp = self._device_mlp(precision)
        self._dev_cache = (sig, mlp)
        return mlp

    def _forward(self, Xd):
        return self._cached_mlp().forward(Xd)

    def _fine_tuning(self, Xd, _labels):
        """Head init, 1/p pre-scale, SGD, p post-scale (dbn/models.py:517-551)."""
        rbms = self.unsupervised_dbn.rbm_layers
        self.num_classes = self._determine_num_output_neurons(_labels)
        h_last = rbms[-1].n_hidden_units
        self.W = np.random.randn(self.num_classes, h_last) / np.sqrt(h_last)
        self.b = np.random.randn(self.num_classes) / np.sqrt(h_last)
        labels = np.asarray(self._transform_labels_to_network_format(_labels), dtype=np.float64)
        if labels.ndim == 1:
            labels = labels[:, None]
        if self.unsupervised_dbn.optimization_algorithm != 'sgd':
            raise ValueError("Invalid optimization algorithm.")
        mlp = self._device_mlp(self._ft_precision())
        for layer in mlp.layers:
            layer.scale_hidden(1.0 / self.p)
        if self.verbose:
            print("[START] Fine tuning step:")
        cb = None
        if self.verbose:
            cb = lambda it, err: print(">> Epoch %d finished \tANN training loss %f" % (it, err))
        targets = np.ascontiguousarray(labels, dtype=np.float32) if isinstance(Xd, np.ndarray) else _to_device_f32(labels)
        mlp.fit(Xd, targets, self.n_iter_backprop, self.batch_size, self.learning_rate,
                self.l2_regularization, self.dropout_p, _draw_seed() if self.rng == 'philox' else 0,
                rng_mode=self.rng, verbose=cb)
        for layer in mlp.layers:
            layer.scale_hidden(self.p)
        for rbm, layer in zip(rbms, mlp.layers):
            rbm.W, _b, rbm.c = layer.export()   # rbm.b is not touched by backprop (dbn/models.py:454-460)
        self.W, self.b = mlp.export_head()
        arrays = [self.W, self.b]
        for r in rbms:
            arrays += [r.W, r.c]
        self._dev_cache = ((mlp.prec_name,) + _signature(*arrays), mlp)     # predict reuses the trained device copy
        if self.verbose:
            print("[END] Fine tuning step")


class SupervisedDBNClassification(AbstractSupervisedDBN, ClassifierMixin):
    """DBN + softmax classifier (reference: dbn/models.py:562-680)."""

    _head = 'softmax'

    def _transform_labels_to_network_format(self, labels):
        """One-hot rows with a first-appearance label->index map (dbn/models.py:568-584).  Same map, same
        rows as the reference's per-sample loop, built from one sort of the labels (the loop costs ~40 ms per
        60000 labels, more than the fine-tune iteration it precedes)."""
        self.label_to_idx_map, self.idx_to_label_map = {}, {}
        raw = labels
        labels = np.asarray(labels)
        onehot = np.zeros((len(labels), self.num_classes))
        # np.asarray coerces a mixed list such as [1, 'a'] to strings and an object array has no total order:
        # those take the reference's own loop, whose keys are the elements as given
        mixed = labels.dtype == object or labels.ndim != 1 or (
            labels.dtype.kind in "US" and not isinstance(raw, np.ndarray) and
            not all(isinstance(x, (str, bytes, np.str_, np.bytes_)) for x in raw))
        uniq = None
        if not mixed:
            try:
                uniq, first, inv = np.unique(labels, return_index=True, return_inverse=True)
            except TypeError:
                uniq = None
        if uniq is None:
            for i, label in enumerate(raw if mixed else labels):
                if label not in self.label_to_idx_map:
                    j = len(self.label_to_idx_map)
                    self.label_to_idx_map[label] = j
                    self.idx_to_label_map[j] = label
                onehot[i, self.label_to_idx_map[label]] = 1
            return onehot
        order = np.argsort(first, kind="stable")          # classes in order of first appearance
        rank = np.empty(len(order), dtype=np.intp)
        rank[order] = np.arange(len(order))
        src = raw if isinstance(raw, (list, tuple)) else labels   # keys are the caller's own elements, as in the loop
        for j, k in enumerate(order):
            key = src[first[k]]
            self.label_to_idx_map[key] = j
            self.idx_to_label_map[j] = key
        onehot[np.arange(len(labels)), rank[np.asarray(inv).reshape(-1)]] = 1
        return onehot

    def _transform_network_format_to_labels(self, indexes):
        return [self.idx_to_label_map[i] for i in indexes]

    def _determine_num_output_neurons(self, labels):
        return len(np.unique(labels))

    def predict_proba(self, X):
        """Class probabilities (n_samples, n_classes) (dbn/models.py:628-634)."""
        return super(SupervisedDBNClassification, self).predict(X)

    def predict_proba_dict(self, X):
        """One {label: probability} dict per sample (dbn/models.py:636-658)."""
        probs = self.predict_proba(X)
        return [{self.idx_to_label_map[j]: row[j] for j in range(probs.shape[1])} for row in probs]

    def predict(self, X):
        """List of original labels (dbn/models.py:660-663)."""
        return self._transform_network_format_to_labels(np.argmax(self.predict_proba(X), axis=1))


class SupervisedDBNRegression(AbstractSupervisedDBN, RegressorMixin):
    """DBN + linear regression head (reference: dbn/models.py:683-741)."""

    _head = 'linear'

    def _transform_labels_to_network_format(self, labels):
        return labels

    def _determine_num_output_neurons(self, labels):
        return 1 if np.ndim(labels) == 1 else np.shape(labels)[1]
