"""Host-side logic of the estimators that needs no GPU: constructor defaults mirror the reference,
parameter init consumes the np.random stream exactly like the reference (checked against golden
weights via n_epochs=0 semantics), label maps, error behaviour, sklearn plumbing, pickle."""
import inspect
import os
import pickle

import numpy as np
import pytest

import dbn_b200
from dbn_b200 import engine as E
from dbn_b200.models import BinaryRBM, SupervisedDBNClassification, SupervisedDBNRegression, UnsupervisedDBN
from oracle import dbn_oracle as orc


def test_constructor_defaults_match_reference():
    # reference: dbn/models.py:31-39, :228-236, :296-309
    d = {k: v.default for k, v in inspect.signature(BinaryRBM.__init__).parameters.items() if k != "self"}
    assert d["n_hidden_units"] == 100 and d["activation_function"] == "sigmoid"
    assert d["optimization_algorithm"] == "sgd" and d["learning_rate"] == 1e-3 and d["n_epochs"] == 10
    assert d["contrastive_divergence_iter"] == 1 and d["batch_size"] == 32 and d["verbose"] is True
    d = {k: v.default for k, v in inspect.signature(UnsupervisedDBN.__init__).parameters.items() if k != "self"}
    assert d["hidden_layers_structure"] == [100, 100] and d["learning_rate_rbm"] == 1e-3 and d["n_epochs_rbm"] == 10
    d = {k: v.default for k, v in inspect.signature(SupervisedDBNClassification.__init__).parameters.items()
         if k != "self"}
    assert d["learning_rate"] == 1e-3 and d["n_iter_backprop"] == 100 and d["l2_regularization"] == 1.0
    assert d["dropout_p"] == 0 and d["batch_size"] == 32


def test_param_init_consumes_rng_like_reference():
    for act in ("sigmoid", "relu"):
        np.random.seed(5)
        m = BinaryRBM(n_hidden_units=7, activation_function=act)
        m._init_params(12)
        np.random.seed(5)
        W, b, c = orc.rbm_init(12, 7, act)   # oracle is pinned to the reference (test_oracle_golden)
        assert np.array_equal(m.W, W) and np.array_equal(m.b, b) and np.array_equal(m.c, c)
        assert m.n_visible_units == 12


def test_invalid_activation_and_optimizer_raise_value_error():
    with pytest.raises(ValueError, match="Invalid activation function."):
        BinaryRBM(activation_function="tanh")._init_params(4)
    with pytest.raises(ValueError, match="Invalid precision."):
        E.resolve_precision("fp8", 10, 10, 10)


def test_epoch_order_is_the_double_shuffle():
    np.random.seed(3)
    o = E.epoch_order(10)
    np.random.seed(3)
    a = np.random.permutation(10)
    b = np.random.permutation(10)
    X = np.arange(10)
    assert np.array_equal(X[o], X[a][b])


def test_label_maps_first_appearance_order():
    clf = SupervisedDBNClassification(hidden_layers_structure=[4])
    y = np.array([7, 3, 7, 5, 3])
    clf.num_classes = clf._determine_num_output_neurons(y)
    Y = clf._transform_labels_to_network_format(y)
    assert clf.label_to_idx_map == {7: 0, 3: 1, 5: 2} and clf.idx_to_label_map == {0: 7, 1: 3, 2: 5}
    assert Y.tolist() == [[1, 0, 0], [0, 1, 0], [1, 0, 0], [0, 0, 1], [0, 1, 0]]
    assert clf._transform_network_format_to_labels([2, 0]) == [5, 7]
    l2i, _ = orc.label_maps(y)
    assert l2i == clf.label_to_idx_map


def test_label_maps_of_mixed_and_string_labels_follow_the_reference_loop():
    """A mixed-type list keeps its elements as keys (np.asarray would coerce [1, 'a'] to strings), exactly like
    the reference's per-sample loop (dbn/models.py:575-584); string and object labels work too."""
    for y in ([1, 'a', 1, 'b', 'a'], ['x', 'y', 'x'], np.array(['u', 'v', 'u']), np.array([(1, 2), 'k', 'k'], dtype=object)):
        clf = SupervisedDBNClassification(hidden_layers_structure=[4])
        clf.num_classes = len(set(list(y)))
        Y = clf._transform_labels_to_network_format(y)
        ref_map = {}
        for label in y:                       # the reference loop
            ref_map.setdefault(label, len(ref_map))
        assert clf.label_to_idx_map == ref_map
        assert [type(k) for k in clf.label_to_idx_map] == [type(k) for k in ref_map]
        assert Y.argmax(axis=1).tolist() == [ref_map[l] for l in y] and Y.sum() == len(y)


def test_regression_output_width():
    r = SupervisedDBNRegression(hidden_layers_structure=[4])
    assert r._determine_num_output_neurons(np.zeros(5)) == 1
    assert r._determine_num_output_neurons(np.zeros((5, 3))) == 3


def test_sklearn_params_and_clone():
    from sklearn.base import clone
    m = SupervisedDBNClassification(hidden_layers_structure=[8, 4], dropout_p=0.2, batch_size=16)
    p = m.get_params()
    assert p["hidden_layers_structure"] == [8, 4] and p["dropout_p"] == 0.2
    c = clone(m)
    assert c.batch_size == 16 and c.p == pytest.approx(0.8)
    assert c.unsupervised_dbn.hidden_layers_structure == [8, 4]


def test_pickle_roundtrip_of_fitted_attributes(tmp_path):
    m = BinaryRBM(n_hidden_units=3, verbose=False)
    np.random.seed(0)
    m._init_params(5)
    path = os.path.join(tmp_path, "rbm.pkl")
    m.save(path)
    m2 = BinaryRBM.load(path)
    assert np.array_equal(m.W, m2.W) and m2._activation_function_class.name == "sigmoid"
    assert isinstance(pickle.loads(pickle.dumps(m)), BinaryRBM)


def test_precision_auto_policy():
    assert E.resolve_precision("auto", 784, 500, 256) == "bf16"
    assert E.resolve_precision("auto", 64, 256, 32) == "fp32"
    assert E.resolve_precision("auto", 13, 100, 16) == "fp32"
    assert E.resolve_precision("fp32", 4096, 4096, 32768) == "fp32"
    assert E.bf16_pitch(784) == 832 and E.bf16_pitch(500) == 512 and E.bf16_pitch(4096) == 4160


def test_public_names():
    assert set(dbn_b200.__all__) == {"BinaryRBM", "UnsupervisedDBN", "SupervisedDBNClassification",
                                     "SupervisedDBNRegression"}


def test_draw_ahead_thread_takes_the_same_draws_in_the_same_order():
    """The helper thread of the stacked pre-training consumes np.random exactly like the sequential code:
    per layer W, c, b (dbn/models.py:55-69), the Philox key, one double shuffle per epoch."""
    from dbn_b200 import UnsupervisedDBN, BinaryRBM
    from dbn_b200 import models as M, engine as E

    def make():
        m = UnsupervisedDBN(hidden_layers_structure=[7, 5], n_epochs_rbm=3, batch_size=4, verbose=False)
        m.rbm_layers = [BinaryRBM(n_hidden_units=h, n_epochs=3, verbose=False) for h in (7, 5)]
        return m

    np.random.seed(3)
    seq, v = [], 9
    for h in (7, 5):
        r = BinaryRBM(n_hidden_units=h)
        seq.append(r._draw_params(v))
        seq.append(M._draw_seed())
        seq += [E.epoch_order(11) for _ in range(3)]
        v = h
    tail_seq = np.random.random_sample()

    np.random.seed(3)
    d = M._DrawAhead(make()._draw_jobs(11, 9), depth=2)
    got = [d.get() for _ in range(len(seq))]
    d.close()
    tail_thr = np.random.random_sample()
    assert tail_seq == tail_thr
    for a, b in zip(seq, got):
        if isinstance(a, tuple):
            assert all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3])) and a[3] is b[3]
        else:
            assert np.array_equal(a, b)


def test_draw_ahead_reports_errors_at_the_failing_draw_and_stops():
    from dbn_b200 import models as M

    def jobs():
        yield 1
        raise ValueError("Invalid activation function.")

    d = M._DrawAhead(jobs())
    assert d.get() == 1
    with pytest.raises(ValueError):
        d.get()
    d.close()

    def endless():
        while True:
            yield np.random.random_sample()
    d = M._DrawAhead(endless(), depth=2)
    d.get()
    d.close()          # must not hang although the generator never ends
    assert not d._t.is_alive()


def test_legacy_randn_is_np_random_randn_bit_for_bit():
    """libdbn_hostrng.so restates NumPy's legacy Gaussian (MT19937 + polar method): same values, same state of the
    GLOBAL generator afterwards, whatever the length parity, cached Gaussian or block position."""
    from dbn_b200 import _hostrng as R
    assert R.available(), "libdbn_hostrng.so missing or failed its self-check (python __graft_entry__.py build)"
    for seed in (0, 7, 2024):
        np.random.seed(seed)
        ref = [np.random.randn(*s) for s in [(5000,), (3,), (129, 65), (4097,), (500, 784), (1,), (2000, 5)]]
        tail_ref = (np.random.random_sample(), np.random.permutation(17), np.random.randn(3))
        np.random.seed(seed)
        got = [R.legacy_randn(*s) for s in [(5000,), (3,), (129, 65), (4097,), (500, 784), (1,), (2000, 5)]]
        tail_got = (np.random.random_sample(), np.random.permutation(17), np.random.randn(3))
        for a, b in zip(ref, got):
            assert a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))
        assert tail_ref[0] == tail_got[0] and np.array_equal(tail_ref[1], tail_got[1]) and np.array_equal(tail_ref[2], tail_got[2])


def test_legacy_randn_parallel_path_is_bit_exact_too():
    """Half a million draws and more take the fully threaded path of csrc/host_rng.c (rejection pass on worker threads,
    only the MT19937 recurrence sequential): same values and same generator state as np.random.randn, for even / odd
    lengths, with and without a cached Gaussian, ending inside and at the edge of a chunk."""
    from dbn_b200 import _hostrng as R
    assert R.available()
    for seed, pre, shapes in ((3, 0, [(1 << 19,), (1025, 1031)]), (11, 1, [((1 << 19) + 1,), (700, 1500), (5,)]),
                              (42, 2, [(2048, 2049)])):
        np.random.seed(seed)
        if pre:
            np.random.randn(pre)
        ref = [np.random.randn(*s) for s in shapes]
        tail_ref = (np.random.random_sample(), np.random.randn(3), np.random.get_state()[2])
        np.random.seed(seed)
        if pre:
            np.random.randn(pre)
        got = [R.legacy_randn(*s) for s in shapes]
        tail_got = (np.random.random_sample(), np.random.randn(3), np.random.get_state()[2])
        for a, b in zip(ref, got):
            assert a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))
        assert tail_ref[0] == tail_got[0] and np.array_equal(tail_ref[1], tail_got[1]) and tail_ref[2] == tail_got[2]


def test_rbm_parameter_init_is_the_reference_stream():
    """BinaryRBM._draw_params consumes np.random exactly like dbn/models.py:55-69 (W, then c, then b)."""
    np.random.seed(11)
    h, v = 96, 200                     # large enough for the fast path (>= 4096 draws)
    W = np.random.randn(h, v) / np.sqrt(v)
    c = np.random.randn(h) / np.sqrt(v)
    b = np.random.randn(v) / np.sqrt(v)
    np.random.seed(11)
    m = BinaryRBM(n_hidden_units=h)
    m._init_params(v)
    assert np.array_equal(m.W, W) and np.array_equal(m.c, c) and np.array_equal(m.b, b)


@pytest.mark.parametrize("scalar", [0, 256])
def test_host_stage_rows_rounds_like_the_device_cast(scalar):
    """csrc/host_stage.c: gathered rows in the operand layout ([cast(X[rows[i]]) | 1 | 0...]) with the device's rounding
    -- bf16 nearest-even, fp16 nearest-even saturating at +-65504 (dbn_common.cuh mat_store) -- checked bit for bit
    against torch's CPU casts on ties, subnormals, overflow, infinities and random bit patterns, for the vector
    (AVX2 / F16C) and the portable scalar conversions, on one thread and several."""
    import torch
    from dbn_b200 import _hoststage as hs
    if not hs.available():
        pytest.skip("libdbn_hostrng.so not built")
    rng = np.random.default_rng(0)
    n, cols = 1200, 203
    X = rng.standard_normal((n, cols)).astype(np.float32)
    special = np.array([0.0, -0.0, 1.0, 65504.0, 65519.0, 65520.0, 70000.0, -70000.0, 1e38, -1e38, np.inf, -np.inf, 2.0 ** -14,
                        2.0 ** -24, 2.0 ** -25, 2.0 ** -25 * 1.0001, 2.0 ** -26, 6.1e-5, 5.96e-8, 3e-8, 1 + 2.0 ** -11,
                        1 + 2.0 ** -10 + 2.0 ** -11, 1 + 2.0 ** -8, 1 + 2.0 ** -7 + 2.0 ** -8, 3.3895313892515355e38], dtype=np.float32)
    X[0, :len(special)] = special
    bits = (rng.integers(0, 1 << 32, (40, cols), dtype=np.uint64).astype(np.uint32)).view(np.float32)
    bits[np.isnan(bits)] = 1.0
    X[1:41] = bits
    X[41:81] = bits * np.float32(1e-6)
    rows = rng.permutation(n)[:1000]
    ldo = (cols + 1 + 63) // 64 * 64
    for kind, dt, it in ((hs.BF16, torch.bfloat16, torch.int16), (hs.F16, torch.float16, torch.int16), (hs.F32, torch.float32, torch.int32)):
        src = torch.from_numpy(X[rows])
        ref = (src.clamp(-65504.0, 65504.0) if dt == torch.float16 else src).to(dt)
        for threads in (1, 5):
            out = torch.full((len(rows), ldo), 7, dtype=dt)
            hs.stage_rows(X, rows, out.data_ptr(), ldo, kind | scalar, True, threads=threads)
            assert torch.equal(out[:, :cols].contiguous().view(it), ref.contiguous().view(it)), (kind, threads)
            assert bool((out[:, cols] == 1).all()) and bool((out[:, cols + 1:] == 0).all())
        plain = torch.full((n, cols), 7, dtype=dt)      # no gather, no ones column, no padding
        hs.stage_rows(X, None, plain.data_ptr(), cols, kind | scalar, False)
        full = torch.from_numpy(X)
        assert torch.equal(plain.view(it), (full.clamp(-65504.0, 65504.0) if dt == torch.float16 else full).to(dt).view(it))


def test_legacy_randn_f32_is_the_narrowed_scaled_draw():
    """csrc/host_rng.c dbn_host_legacy_randn_f32: `(np.random.randn(h, v) / sqrt(v)).astype(float32)` (the initial
    weights, dbn/models.py:55-57, as the device stores them) without the float64 array -- values and generator state bit
    for bit, on the threaded path (>= 2^19 draws) and the sequential one, with and without a cached Gaussian."""
    from dbn_b200 import _hostrng as hr
    if not hr.available():
        pytest.skip("libdbn_hostrng.so not built")
    for shape, pre in (((700, 801), 0), ((1031, 517), 3), ((50, 100), 1), ((3,), 0)):
        s = np.sqrt(shape[-1])
        np.random.seed(77)
        np.random.randn(pre)                      # odd `pre` leaves a cached Gaussian behind
        ref = (np.random.randn(*shape) / s).astype(np.float32)
        tail_ref = np.random.randn(5)
        np.random.seed(77)
        np.random.randn(pre)
        out = np.full(shape, 7, dtype=np.float32)
        hr.legacy_randn_f32(out, s)
        tail_got = np.random.randn(5)
        assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), shape
        assert np.array_equal(tail_ref, tail_got), shape
