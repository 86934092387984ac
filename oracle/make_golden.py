#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (`/root/reference/dbn`, NumPy
backend) on seeded inputs.  Run in the build container only (the reference does not travel to the
GPU box); the fixtures it writes are committed.

    python oracle/make_golden.py            # rewrites tests/golden/

The reference has no tests / golden vectors of its own, so these outputs of
the reference itself are what pins the oracle (oracle/dbn_oracle.py) and, through it, the CUDA path.
Randomness is injected in two ways:
  * seed-and-replay: `np.random.seed(s)` then the reference's own `fit`;
  * patching: `np.random.random_sample` / `np.random.binomial` are temporarily replaced by functions
    popping from pre-generated arrays, so single steps see known uniforms / dropout masks.
"""
import contextlib
import os
import sys

import numpy as np

REF = os.environ.get("DBN_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _import_reference():
    if not os.path.isdir(os.path.join(REF, "dbn")):
        raise SystemExit("reference not found at %s (this script only runs in the build container)" % REF)
    sys.path.insert(0, REF)
    import dbn.models as ref_models  # noqa
    return ref_models


@contextlib.contextmanager
def injected(uniform_rows=None, mask_rows=None):
    """Patch the two RNG entry points the hot path uses (dbn/models.py:155, :402, :409)."""
    old_rs, old_bin = np.random.random_sample, np.random.binomial
    u_iter = iter(uniform_rows) if uniform_rows is not None else None
    m_iter = iter(mask_rows) if mask_rows is not None else None

    def fake_rs(size=None):
        row = next(u_iter)
        assert size == row.shape[0], (size, row.shape)
        return row

    def fake_bin(n, p, size=None):
        row = next(m_iter)
        assert size == row.shape[0], (size, row.shape)
        return row.astype(np.int64)

    if u_iter is not None:
        np.random.random_sample = fake_rs
    if m_iter is not None:
        np.random.binomial = fake_bin
    try:
        yield
    finally:
        np.random.random_sample, np.random.binomial = old_rs, old_bin


def make_rbm_fit(ref, name, act, k, seed, n=37, v=12, h=7, bs=8, epochs=2, lr=0.05):
    rng = np.random.default_rng(seed)
    X = (rng.random((n, v)) < 0.4).astype(np.float32) if act == "sigmoid" else rng.random((n, v)).astype(np.float32)
    np.random.seed(seed)
    m = ref.BinaryRBM(n_hidden_units=h, activation_function=act, learning_rate=lr, n_epochs=epochs,
                      contrastive_divergence_iter=k, batch_size=bs, verbose=False)
    m.fit(X)
    T = m.transform(X)
    err = m._compute_reconstruction_error(X)
    np.savez(os.path.join(OUT, name + ".npz"), X=X, W=m.W, b=m.b, c=m.c, T=T, recon=err,
             meta=np.array([n, v, h, bs, epochs, k, seed]), lr=lr, act=act)


def make_cd_step(ref, name, act, k, seed, B=9, v=11, h=6, lr=0.1, bs=16):
    rng = np.random.default_rng(seed)
    m = ref.BinaryRBM(n_hidden_units=h, activation_function=act, learning_rate=lr,
                      contrastive_divergence_iter=k, batch_size=bs, verbose=False)
    m.n_visible_units = v
    m.W = rng.standard_normal((h, v)) / np.sqrt(v)
    m.b = rng.standard_normal(v) * 0.1
    m.c = rng.standard_normal(h) * 0.1
    m._activation_function_class = (ref.SigmoidActivationFunction if act == "sigmoid"
                                    else ref.ReLUActivationFunction)
    V0 = rng.random((B, v))
    U = rng.random((B, k, h), dtype=np.float32).astype(np.float64)
    dW = np.zeros((h, v)); db = np.zeros(v); dc = np.zeros(h)
    with injected(uniform_rows=[U[i, t] for i in range(B) for t in range(k)]):
        for i in range(B):
            a, bb, cc = m._contrastive_divergence(V0[i])
            dW += a; db += bb; dc += cc
    P0 = m._compute_hidden_units_matrix(V0)
    np.savez(os.path.join(OUT, name + ".npz"), W=m.W, b=m.b, c=m.c, V0=V0, U=U, dW=dW, db=db, dc=dc,
             P0=P0, k=k, act=act, lr=lr, bs=bs)


def make_supervised_fit(ref, name, kind, act, dropout, seed, n=41, v=10, hs=(8, 6), bs=8,
                        ep_rbm=2, it=3, k=1):
    rng = np.random.default_rng(seed)
    X = rng.random((n, v)).astype(np.float32)
    if kind == "clf":
        y = np.array([3, 7, 5])[rng.integers(0, 3, n)]   # non-sorted first-appearance order matters
        cls = ref.SupervisedDBNClassification
    else:
        y = (X @ rng.standard_normal((v, 2)) + 0.05 * rng.standard_normal((n, 2)))
        cls = ref.SupervisedDBNRegression
    np.random.seed(seed)
    m = cls(hidden_layers_structure=list(hs), learning_rate_rbm=0.05, learning_rate=0.1,
            n_epochs_rbm=ep_rbm, n_iter_backprop=it, batch_size=bs, activation_function=act,
            dropout_p=dropout, contrastive_divergence_iter=k, verbose=False)
    m.fit(X, y)
    out = {"X": X, "y": y, "Wo": m.W, "bo": m.b, "dropout": dropout, "act": act, "kind": kind,
           "meta": np.array([n, v, bs, ep_rbm, it, k, seed]), "hs": np.array(hs)}
    for i, r in enumerate(m.unsupervised_dbn.rbm_layers):
        out["W%d" % i], out["b%d" % i], out["c%d" % i] = r.W, r.b, r.c
    P = m.predict_proba(X) if kind == "clf" else m.predict(X)
    out["pred"] = np.asarray(P)
    if kind == "clf":
        out["pred_labels"] = np.asarray(m.predict(X))
    np.savez(os.path.join(OUT, name + ".npz"), **out)


def make_finetune_step(ref, name, kind, act, dropout, seed, B=7, v=9, hs=(8, 5), n_out=3):
    """Summed gradients of one mini-batch through the reference's `_backpropagation`
    (dbn/models.py:471-515) with injected dropout masks."""
    rng = np.random.default_rng(seed)
    cls = ref.SupervisedDBNClassification if kind == "clf" else ref.SupervisedDBNRegression
    m = cls(hidden_layers_structure=list(hs), activation_function=act, dropout_p=dropout,
            verbose=False)
    actcls = ref.SigmoidActivationFunction if act == "sigmoid" else ref.ReLUActivationFunction
    layers = []
    prev = v
    for h in hs:
        r = ref.BinaryRBM(n_hidden_units=h, activation_function=act, verbose=False)
        r.n_visible_units = prev
        r.W = rng.standard_normal((h, prev)) / np.sqrt(prev)
        r.c = rng.standard_normal(h) * 0.1
        r.b = np.zeros(prev)
        r._activation_function_class = actcls
        layers.append(r)
        prev = h
    m.unsupervised_dbn.rbm_layers = layers
    m.num_classes = n_out
    m.W = rng.standard_normal((n_out, prev)) / np.sqrt(prev)
    m.b = rng.standard_normal(n_out) * 0.1
    X = rng.random((B, v))
    if kind == "clf":
        Y = np.eye(n_out)[rng.integers(0, n_out, B)]
    else:
        Y = rng.standard_normal((B, n_out))
    widths = [v] + list(hs)
    masks = [(rng.random((B, w)) < (1 - dropout)).astype(np.float64) for w in widths]
    gW = [np.zeros_like(r.W) for r in layers] + [np.zeros_like(m.W)]
    gb = [np.zeros_like(r.c) for r in layers] + [np.zeros_like(m.b)]
    outs = np.zeros((B, n_out))
    rows = [masks[j][i] for i in range(B) for j in range(len(widths))]
    with injected(mask_rows=rows if dropout > 0 else None):
        for i in range(B):
            a, bb, pred = m._backpropagation(np.array(X[i]), Y[i])
            for l in range(len(gW)):
                gW[l] += a[l]; gb[l] += bb[l]
            outs[i] = pred
    out = {"X": X, "Y": Y, "Wo": m.W, "bo": m.b, "out": outs, "dropout": dropout, "act": act,
           "kind": kind, "hs": np.array(hs)}
    for l, r in enumerate(layers):
        out["W%d" % l], out["c%d" % l] = r.W, r.c
    for l in range(len(gW)):
        out["gW%d" % l], out["gb%d" % l] = gW[l], gb[l]
    for j, mk in enumerate(masks):
        out["M%d" % j] = mk
    np.savez(os.path.join(OUT, name + ".npz"), **out)


def make_digits_anchor(ref, name, seed=1337):
    """Reduced README config (README.md:21-38 of the reference): digits 64-[256,256]-10, relu,
    dropout 0.2, batch 32 -- 2 RBM epochs / 3 backprop iterations so the fixture generates in seconds.
    Stores the statistical anchors (accuracy, reconstruction errors), not the weights."""
    from sklearn.datasets import load_digits
    from sklearn.model_selection import train_test_split
    d = load_digits()
    X = (d.data / 16).astype(np.float32)
    Xtr, Xte, ytr, yte = train_test_split(X, d.target, test_size=0.2, random_state=0)
    np.random.seed(seed)
    m = ref.SupervisedDBNClassification(hidden_layers_structure=[256, 256], learning_rate_rbm=0.05,
                                        learning_rate=0.1, n_epochs_rbm=2, n_iter_backprop=3,
                                        batch_size=32, activation_function='relu', dropout_p=0.2,
                                        verbose=False)
    m.pre_train(Xtr)
    inp = Xtr
    recon = []
    for r in m.unsupervised_dbn.rbm_layers:
        recon.append(r._compute_reconstruction_error(inp))
        inp = r.transform(inp)
    m.fit(Xtr, ytr, pre_train=False)
    acc = float(np.mean(np.asarray(m.predict(Xte)) == yte))
    np.savez(os.path.join(OUT, name + ".npz"), acc=acc, recon=np.array(recon), seed=seed,
             n_train=len(Xtr), n_test=len(Xte))
    print("digits anchor: acc=%.4f recon=%s" % (acc, recon))


def make_verbose_prints(ref, name, kind, act, dropout, seed, n=41, v=10, hs=(8, 6), bs=8, ep_rbm=3, it=4):
    """What the reference PRINTS with verbose=True (the default): the per-epoch reconstruction error of every RBM
    (dbn/models.py:120-122, :212-220) and the per-iteration 'ANN training loss' (:426-427, :448-451, :467-469) --
    for a classifier that is num_classes x the mean cross-entropy, because the scalar loss of a sample is broadcast
    into every column of the [N, num_classes] error matrix before its row sums are averaged (:450 with :680)."""
    import io
    import re
    rng = np.random.default_rng(seed)
    X = rng.random((n, v)).astype(np.float32)
    if kind == "clf":
        y = np.array([3, 7, 5, 9])[rng.integers(0, 4, n)]
        cls = ref.SupervisedDBNClassification
    else:
        y = (X @ rng.standard_normal((v, 2)) + 0.05 * rng.standard_normal((n, 2)))
        cls = ref.SupervisedDBNRegression
    np.random.seed(seed)
    m = cls(hidden_layers_structure=list(hs), learning_rate_rbm=0.05, learning_rate=0.1, n_epochs_rbm=ep_rbm,
            n_iter_backprop=it, batch_size=bs, activation_function=act, dropout_p=dropout, verbose=True)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        m.fit(X, y)
    text = buf.getvalue()
    rbm = [float(x) for x in re.findall(r"RBM Reconstruction error ([-0-9.eE+]+)", text)]
    ann = [float(x) for x in re.findall(r"ANN training loss ([-0-9.eE+]+)", text)]
    assert len(rbm) == ep_rbm * len(hs) and len(ann) == it, (len(rbm), len(ann))
    np.savez(os.path.join(OUT, name + ".npz"), X=X, y=y, rbm_errors=np.array(rbm), ann_losses=np.array(ann),
             dropout=dropout, act=act, kind=kind, meta=np.array([n, v, bs, ep_rbm, it, 1, seed]), hs=np.array(hs),
             num_classes=(len(np.unique(y)) if kind == "clf" else y.shape[1]), Wo=m.W)
    print(name, "rbm", rbm, "ann", ann)


def main():
    ref = _import_reference()
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])
    if only:      # regenerate selected fixtures only (np.savez archives carry timestamps: avoid churning the rest)
        if "verbose_clf_sigmoid" in only:
            make_verbose_prints(ref, "verbose_clf_sigmoid", "clf", "sigmoid", 0.0, 51)
        if "verbose_reg_relu" in only:
            make_verbose_prints(ref, "verbose_reg_relu", "reg", "relu", 0.0, 52, hs=(9,))
        return
    make_verbose_prints(ref, "verbose_clf_sigmoid", "clf", "sigmoid", 0.0, 51)
    make_verbose_prints(ref, "verbose_reg_relu", "reg", "relu", 0.0, 52, hs=(9,))
    make_rbm_fit(ref, "rbm_fit_sigmoid_k1", "sigmoid", 1, 11)
    make_rbm_fit(ref, "rbm_fit_sigmoid_k3", "sigmoid", 3, 12)
    make_rbm_fit(ref, "rbm_fit_relu_k1", "relu", 1, 13)
    make_rbm_fit(ref, "rbm_fit_relu_k2", "relu", 2, 14, n=16, bs=16)      # exact multiple of bs
    make_cd_step(ref, "cd_step_sigmoid_k1", "sigmoid", 1, 21)
    make_cd_step(ref, "cd_step_sigmoid_k4", "sigmoid", 4, 22)
    make_cd_step(ref, "cd_step_relu_k1", "relu", 1, 23)
    make_cd_step(ref, "cd_step_relu_k10", "relu", 10, 24, B=5, v=13, h=10)
    make_supervised_fit(ref, "sup_clf_sigmoid", "clf", "sigmoid", 0.0, 31)
    make_supervised_fit(ref, "sup_clf_relu_dropout", "clf", "relu", 0.25, 32)
    make_supervised_fit(ref, "sup_clf_sigmoid_dropout_hi", "clf", "sigmoid", 0.6, 35, hs=(7,))
    make_supervised_fit(ref, "sup_reg_relu", "reg", "relu", 0.0, 33, hs=(9,))
    make_supervised_fit(ref, "sup_reg_sigmoid_dropout", "reg", "sigmoid", 0.2, 34)
    make_finetune_step(ref, "ft_step_clf_sigmoid", "clf", "sigmoid", 0.0, 41)
    make_finetune_step(ref, "ft_step_clf_relu_dropout", "clf", "relu", 0.3, 42)
    make_finetune_step(ref, "ft_step_reg_relu", "reg", "relu", 0.0, 43, n_out=2)
    make_finetune_step(ref, "ft_step_reg_sigmoid_dropout", "reg", "sigmoid", 0.4, 44, n_out=1)
    make_digits_anchor(ref, "digits_anchor")
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
