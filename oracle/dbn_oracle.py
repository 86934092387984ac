"""CPU oracle for the deep-belief-network training hot path (TEST INFRASTRUCTURE ONLY).

This module is a float64 NumPy restatement of the reference's CD-k / backprop hot path
(albertbup/deep-belief-network 1.0.5, `dbn/models.py`, `dbn/activations.py`, `dbn/utils.py`).
It is the *checker*: only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it.  Nothing under `deep-belief-network_b200/`
(the product) imports it, and the product has no CPU fallback.

Parity status: PINNED.  The reference has no tests or golden vectors of its own
(SURVEY.md §4), so the oracle is pinned against outputs of the reference itself: the script
`oracle/make_golden.py` imports the unmodified reference from `/root/reference`, runs it on seeded
inputs (whole `fit` runs and single steps with injected uniforms / dropout masks) and commits the
results under `tests/golden/`; `tests/test_oracle_golden.py` checks every function here against
those fixtures (max-abs <= 1e-12; observed ~1e-16).

Two restatements live here:
  * batched (matrix) form  -- `cd_step`, `finetune_step`, `rbm_fit_replay`, `finetune_fit_replay`:
    the parity oracle for shapes where the reference's per-sample Python loop is too slow;
  * per-sample form        -- `cd_sample`, `rbm_epoch_per_sample`, `backprop_sample`,
    `finetune_epoch_per_sample`: has the reference's cost profile (one Python iteration, 4 mat-vecs
    and 2 outer products per row) and is what `bench.py` times as the CPU baseline ("port").

Every function cites the reference lines it follows (paths relative to the reference root).
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------------------------
# activations  (dbn/activations.py:23-38 sigmoid, :43-58 relu)
# ----------------------------------------------------------------------------------------------

SIGMOID = "sigmoid"
RELU = "relu"


def act_fn(name):
    """Return (f, fprime_of_activation).  fprime takes the *activation*, not the pre-activation
    (dbn/activations.py:38 `x*(1-x)`, :58 `(x>0)`)."""
    if name == SIGMOID:
        return (lambda x: 1.0 / (1.0 + np.exp(-x))), (lambda a: a * (1.0 - a))
    if name == RELU:
        return (lambda x: np.maximum(x, 0.0)), (lambda a: (a > 0).astype(np.float64))
    raise ValueError("Invalid activation function.")  # dbn/models.py:69


# ----------------------------------------------------------------------------------------------
# parameter init  (dbn/models.py:55-69, :524-528)
# ----------------------------------------------------------------------------------------------

def rbm_init(n_visible, n_hidden, activation):
    """Draw W (H,V), c (H,), b (V,) from the *global* legacy np.random stream in the reference's
    order: sigmoid -> randn W, randn c, randn b, all / sqrt(V) (dbn/models.py:58-60); relu ->
    truncnorm(-0.2, 0.2) / sqrt(V) for W and the constant 0.1/sqrt(V) for c, b (:63-66)."""
    s = np.sqrt(n_visible)
    if activation == SIGMOID:
        W = np.random.randn(n_hidden, n_visible) / s
        c = np.random.randn(n_hidden) / s
        b = np.random.randn(n_visible) / s
    elif activation == RELU:
        from scipy.stats import truncnorm
        W = truncnorm.rvs(-0.2, 0.2, size=[n_hidden, n_visible]) / s
        c = np.full(n_hidden, 0.1) / s
        b = np.full(n_visible, 0.1) / s
    else:
        raise ValueError("Invalid activation function.")
    return W, b, c


def head_init(n_out, n_last_hidden):
    """Head W (out,H_last) then b (out,), randn / sqrt(H_last) (dbn/models.py:524-528)."""
    s = np.sqrt(n_last_hidden)
    W = np.random.randn(n_out, n_last_hidden) / s
    b = np.random.randn(n_out) / s
    return W, b


def epoch_order(n):
    """The reference shuffles twice per epoch: `permutation` in the SGD loop
    (dbn/models.py:106, :435) and again inside batch_generator (dbn/utils.py:12).  Returns the
    composed index array: row i of the epoch is original row order[i]."""
    first = np.random.permutation(n)
    second = np.random.permutation(n)
    return first[second]


def label_maps(labels):
    """First-appearance label->index map (dbn/models.py:575-584)."""
    l2i, i2l = {}, {}
    for lab in labels:
        if lab not in l2i:
            i2l[len(l2i)] = lab
            l2i[lab] = len(l2i)
    return l2i, i2l


def one_hot(labels, l2i):
    Y = np.zeros((len(labels), len(l2i)))
    for i, lab in enumerate(labels):
        Y[i, l2i[lab]] = 1.0
    return Y


# ----------------------------------------------------------------------------------------------
# batched restatement: RBM
# ----------------------------------------------------------------------------------------------

def hidden_means(W, c, V, activation):
    """act(V W^T + c)  (dbn/models.py:176-183)."""
    f, _ = act_fn(activation)
    return f(V @ W.T + c[None, :])


def visible_means(W, b, Hs, activation):
    """act(H W + b)  (dbn/models.py:195-201)."""
    f, _ = act_fn(activation)
    return f(Hs @ W + b[None, :])


def cd_chain(W, b, c, V0, U, k, activation):
    """Gibbs chain of CD-k for a whole mini-batch (dbn/models.py:124-141, :148-155).

    V0 (B,V) float64; U (B,k,H) uniforms consumed in the reference's order (row-major per sample,
    k draws of H each).  Only hidden units are sampled (strict `<`, :155); reconstructions are
    means (:136).  Returns dict with P0 (t=0 hidden means == h_0 of :140), the list of hidden
    samples, Vk and Pk (:141)."""
    Vt = V0
    samples = []
    P0 = None
    for t in range(k):
        Pt = hidden_means(W, c, Vt, activation)
        if t == 0:
            P0 = Pt
        Ht = (U[:, t, :] < Pt).astype(np.float64)
        samples.append(Ht)
        Vt = visible_means(W, b, Ht, activation)
    if P0 is None:  # k == 0: no sampling, v_k == v_0
        P0 = hidden_means(W, c, V0, activation)
    Pk = hidden_means(W, c, Vt, activation)
    return {"P0": P0, "samples": samples, "Vk": Vt, "Pk": Pk}


def cd_deltas(W, b, c, V0, U, k, activation):
    """Summed CD-k statistics of a mini-batch (dbn/models.py:142-144 summed by :112-116)."""
    ch = cd_chain(W, b, c, V0, U, k, activation)
    dW = ch["P0"].T @ V0 - ch["Pk"].T @ ch["Vk"]
    db = (V0 - ch["Vk"]).sum(axis=0)
    dc = (ch["P0"] - ch["Pk"]).sum(axis=0)
    return dW, db, dc, ch


def cd_step(W, b, c, V0, U, k, activation, lr, batch_size):
    """One mini-batch update, in place.  Divides by the *configured* batch_size even on the short
    last batch (dbn/models.py:117-119)."""
    dW, db, dc, ch = cd_deltas(W, b, c, V0, U, k, activation)
    W += lr * (dW / batch_size)
    b += lr * (db / batch_size)
    c += lr * (dc / batch_size)
    return dW, db, dc, ch


def reconstruction_error(W, b, c, X, activation):
    """mean_n sum_j (recon - x)^2  (dbn/models.py:212-220)."""
    Hm = hidden_means(W, c, X, activation)
    R = visible_means(W, b, Hm, activation)
    return float(np.mean(np.sum((R - X) ** 2, axis=1)))


def rbm_fit_replay(X, n_hidden, activation="sigmoid", lr=1e-3, n_epochs=10, k=1, batch_size=32, epoch_errors=None):
    """Whole `BinaryRBM.fit` replayed on the global np.random stream in the reference's order
    (SURVEY.md A.3): init, then per epoch two permutations and `random_sample((N,k,H))`
    (== k sequential `random_sample(H)` per row, dbn/models.py:134-135,155).  `epoch_errors` (a list) receives what
    the reference prints after every epoch with verbose=True: the reconstruction error of the whole data set under
    the epoch's final weights (dbn/models.py:120-122; it consumes no random numbers)."""
    X = np.asarray(X)
    n, v = X.shape
    W, b, c = rbm_init(v, n_hidden, activation)
    for _ in range(n_epochs):
        data = X[np.random.permutation(n)]            # dbn/models.py:106-107
        data = data[np.random.permutation(n)]         # dbn/utils.py:12-13
        U = np.random.random_sample((n, k, n_hidden))
        for s in range(0, n, batch_size):
            cd_step(W, b, c, data[s:s + batch_size].astype(np.float64), U[s:s + batch_size],
                    k, activation, lr, batch_size)
        if epoch_errors is not None:
            epoch_errors.append(reconstruction_error(W, b, c, X.astype(np.float64), activation))
    return W, b, c


def dbn_transform(layers, X, activation):
    """Stacked transform (dbn/models.py:278-287); layers = [(W,b,c), ...]."""
    A = np.asarray(X, dtype=np.float64)
    for (W, _b, c) in layers:
        A = hidden_means(W, c, A, activation)
    return A


# ----------------------------------------------------------------------------------------------
# batched restatement: supervised fine-tune
# ----------------------------------------------------------------------------------------------

def head_outputs(Wo, bo, A, head):
    """softmax(A Wo^T + bo) without max-shift (dbn/models.py:607-615) or linear (:705-711)."""
    Z = A @ Wo.T + bo[None, :]
    if head == "softmax":
        E = np.exp(Z)
        return E / E.sum(axis=1, keepdims=True)
    return Z


def finetune_forward(layers, Wo, bo, X, masks, activation, head):
    """Dropout forward for a mini-batch (dbn/models.py:394-417).  `masks` = [M0 (B,V), M1 (B,H1),
    ...] of {0,1} or None (dropout_p == 0).  Masks are NOT rescaled by 1/p (:403, :411)."""
    f, _ = act_fn(activation)
    a = np.array(X, dtype=np.float64)
    if masks is not None:
        a = a * masks[0]
    acts = [a]
    for li, (W, c) in enumerate(layers):
        a = f(a @ W.T + c[None, :])
        if masks is not None:
            a = a * masks[li + 1]
        acts.append(a)
    out = head_outputs(Wo, bo, a, head)
    return acts, out


def finetune_grads(layers, Wo, bo, X, Y, masks, activation, head):
    """Summed per-layer gradients of a mini-batch (dbn/models.py:471-515 summed by :443-447).
    delta_out = out - Y for both heads (:617-626, :713-720).  prime() sees the masked
    activations (:496-499).  Returns (gW list, gbias list, out, acts); the last entry is the head."""
    _, fp = act_fn(activation)
    acts, out = finetune_forward(layers, Wo, bo, X, masks, activation, head)
    Ws = [W for (W, _c) in layers] + [Wo]
    L = len(layers)
    deltas = [None] * (L + 1)
    deltas[L] = out - Y
    for l in range(L - 1, -1, -1):
        deltas[l] = (deltas[l + 1] @ Ws[l + 1]) * fp(acts[l + 1])
    gW = [deltas[l].T @ acts[l] for l in range(L + 1)]
    gb = [deltas[l].sum(axis=0) for l in range(L + 1)]
    return gW, gb, out, acts, deltas


def finetune_step(layers, Wo, bo, X, Y, masks, activation, head, lr, l2, n_samples, batch_size):
    """One fine-tune mini-batch update (dbn/models.py:454-465): W <- (1 - lr*l2/N) W - lr*g/bs for
    every layer and the head; biases un-decayed.  `layers` is a list of [W, c] lists mutated /
    rebound in place; returns (Wo, bo, out)."""
    gW, gb, out, _acts, _d = finetune_grads([(l[0], l[1]) for l in layers], Wo, bo, X, Y, masks,
                                            activation, head)
    decay = 1.0 - (lr * l2) / n_samples
    for l, layer in enumerate(layers):
        layer[0] = decay * layer[0] - lr * (gW[l] / batch_size)
        layer[1] = layer[1] - lr * (gb[l] / batch_size)
    Wo = decay * Wo - lr * (gW[-1] / batch_size)
    bo = bo - lr * (gb[-1] / batch_size)
    return Wo, bo, out


def sample_loss(out, Y, head):
    """Per-sample loss row as stored in the verbose loss matrix (dbn/models.py:448-451):
    classification: -log p[label] broadcast to all num_classes columns (:673-680);
    regression: squared error per target (:733-741).  Returns the (B, C) block."""
    if head == "softmax":
        ce = -np.log(np.sum(out * Y, axis=1))
        return np.repeat(ce[:, None], out.shape[1], axis=1)
    return (out - Y) ** 2


def finetune_fit_replay(layers, X, Ynet, head, activation, lr, l2, n_iter, batch_size, dropout_p):
    """Whole `_fine_tuning` replayed on the global np.random stream (dbn/models.py:517-551).
    `layers` = list of [W, c] from pre-training (copied).  Ynet = labels already in network
    format (one-hot (N,C) or regression targets (N,T)).  Returns (layers, Wo, bo, epoch_losses)."""
    X = np.asarray(X)
    n = X.shape[0]
    p = 1.0 - dropout_p
    layers = [[np.array(W, dtype=np.float64), np.array(c, dtype=np.float64)] for (W, c) in layers]
    n_out = Ynet.shape[1]
    Wo, bo = head_init(n_out, layers[-1][0].shape[0])
    for l in layers:                                   # :533-535
        l[0] = l[0] / p
        l[1] = l[1] / p
    losses = []
    for _ in range(n_iter):
        idx = np.random.permutation(n)                 # :435-437
        data, labels = X[idx], Ynet[idx]
        idx2 = np.random.permutation(n)                # utils.py:12
        data, labels = data[idx2], labels[idx2]
        masks_all = None
        if dropout_p > 0:                              # per row: V then H_1.. (:402, :409)
            widths = [X.shape[1]] + [l[0].shape[0] for l in layers]
            masks_all = [np.empty((n, w)) for w in widths]
            for r in range(n):
                for mi, w in enumerate(widths):
                    masks_all[mi][r] = np.random.binomial(1, p, w)
        ep_loss = np.zeros((n, n_out))
        for s in range(0, n, batch_size):
            sl = slice(s, s + batch_size)
            masks = None if masks_all is None else [m[sl] for m in masks_all]
            Wo, bo, out = finetune_step(layers, Wo, bo, data[sl].astype(np.float64), labels[sl],
                                        masks, activation, head, lr, l2, n, batch_size)
            ep_loss[sl] = sample_loss(out, labels[sl], head)
        losses.append(float(np.mean(np.sum(ep_loss, axis=1))))  # :467-469
    for l in layers:                                   # :546-548 (head NOT rescaled)
        l[0] = l[0] * p
        l[1] = l[1] * p
    return layers, Wo, bo, losses


# ----------------------------------------------------------------------------------------------
# per-sample restatement (the reference's cost profile) -- CPU baseline for bench.py
# ----------------------------------------------------------------------------------------------

def cd_sample(W, b, c, v0, k, f):
    """CD-k statistics of one row, drawing from np.random like the reference
    (dbn/models.py:124-146): k x (sample h, mean v); then h_0, h_k recomputed; two outer products."""
    vt = v0
    for _ in range(k):
        ph = f(W @ vt + c)
        ht = (np.random.random_sample(ph.shape[0]) < ph).astype(np.int64)
        vt = f(ht @ W + b)
    h0 = f(W @ v0 + c)
    hk = f(W @ vt + c)
    return np.outer(h0, v0) - np.outer(hk, vt), v0 - vt, h0 - hk


def rbm_epoch_per_sample(W, b, c, X, k, activation, lr, batch_size):
    """One epoch of the reference's RBM SGD loop (dbn/models.py:105-119), row by row."""
    f, _ = act_fn(activation)
    data = X[epoch_order(len(X))]
    aW, ab, ac = np.zeros(W.shape), np.zeros(b.shape), np.zeros(c.shape)
    for s in range(0, len(data), batch_size):
        aW[:] = 0.0
        ab[:] = 0.0
        ac[:] = 0.0
        for row in data[s:s + batch_size]:
            dW, db, dc = cd_sample(W, b, c, row, k, f)
            aW += dW
            ab += db
            ac += dc
        W += lr * (aW / batch_size)
        b += lr * (ab / batch_size)
        c += lr * (ac / batch_size)


def backprop_sample(Ws, cs, Wo, bo, x, y, f, fp, head, p, dropout_p):
    """Forward + backward of one row (dbn/models.py:394-417, :471-515)."""
    a = np.array(x, dtype=np.float64)
    if dropout_p > 0:
        a *= np.random.binomial(1, p, a.shape[0])
    acts = [a]
    for W, c in zip(Ws, cs):
        a = f(W @ a + c)
        if dropout_p > 0:
            a = a * np.random.binomial(1, p, a.shape[0])
        acts.append(a)
    z = Wo @ a + bo
    if head == "softmax":
        e = np.exp(z)
        out = e / e.sum()
    else:
        out = z
    delta = out - y
    allW = list(Ws) + [Wo]
    gW, gb = [None] * len(allW), [None] * len(allW)
    for l in range(len(allW) - 1, -1, -1):
        gW[l] = np.outer(delta, acts[l])
        gb[l] = delta
        if l > 0:
            delta = (delta @ allW[l]) * fp(acts[l])
    return gW, gb, out


def finetune_epoch_per_sample(Ws, cs, Wo, bo, X, Ynet, activation, head, lr, l2, batch_size,
                              dropout_p):
    """One iteration of the reference's fine-tune SGD loop (dbn/models.py:434-465), row by row.
    Ws/cs are lists mutated in place; returns the rebound (Wo, bo)."""
    f, fp = act_fn(activation)
    p = 1.0 - dropout_p
    n = len(X)
    order = epoch_order(n)
    data, labels = X[order], Ynet[order]
    decay = 1.0 - (lr * l2) / n
    accW = [np.zeros(W.shape) for W in Ws] + [np.zeros(Wo.shape)]
    accb = [np.zeros(c.shape) for c in cs] + [np.zeros(bo.shape)]
    for s in range(0, n, batch_size):
        for a1, a2 in zip(accW, accb):
            a1[:] = 0.0
            a2[:] = 0.0
        for x, y in zip(data[s:s + batch_size], labels[s:s + batch_size]):
            gW, gb, _ = backprop_sample(Ws, cs, Wo, bo, x, y, f, fp, head, p, dropout_p)
            for l in range(len(accW)):
                accW[l] += gW[l]
                accb[l] += gb[l]
        for l in range(len(Ws)):
            Ws[l] = decay * Ws[l] - lr * (accW[l] / batch_size)
            cs[l] -= lr * (accb[l] / batch_size)
        Wo = decay * Wo - lr * (accW[-1] / batch_size)
        bo = bo - lr * (accb[-1] / batch_size)
    return Wo, bo
