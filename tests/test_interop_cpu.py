"""Model interchange with the reference NumPy backend.  Needs the reference sources,
which exist only in the build container: skipped elsewhere (nothing on the GPU box reads them)."""
import os
import sys

import numpy as np
import pytest

REF = os.environ.get("DBN_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "dbn")), reason="reference sources not present")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, REF)
    import dbn.models as R
    yield R
    sys.path.remove(REF)


def test_reference_classifier_roundtrip(ref, tmp_path):
    from dbn_b200 import SupervisedDBNClassification, interop
    rng = np.random.default_rng(0)
    X = rng.random((60, 12)).astype(np.float32)
    y = np.array(["b", "a", "c"])[rng.integers(0, 3, 60)]
    np.random.seed(3)
    r = ref.SupervisedDBNClassification(hidden_layers_structure=[8, 6], n_epochs_rbm=2, n_iter_backprop=3,
                                        batch_size=16, dropout_p=0.1, activation_function="relu", verbose=False)
    r.fit(X, y)
    path = os.path.join(tmp_path, "ref.pkl")
    r.save(path)
    g = interop.load_reference_pickle(path)
    assert isinstance(g, SupervisedDBNClassification)
    assert g.label_to_idx_map == r.label_to_idx_map and g.num_classes == 3 and g.p == pytest.approx(0.9)
    for a, b in zip(g.unsupervised_dbn.rbm_layers, r.unsupervised_dbn.rbm_layers):
        assert np.array_equal(a.W, b.W) and np.array_equal(a.c, b.c) and np.array_equal(a.b, b.b)
        assert a._activation_function_class.__name__ == "ReLUActivationFunction"
        assert a._activation_function_class.__module__.startswith("dbn_b200")
    assert np.array_equal(g.W, r.W) and np.array_equal(g.b, r.b)
    # (serving the converted model on the GPU -- predict, predict_proba, _compute_output_units_matrix -- is
    #  tests/test_gpu_estimators.py::test_converted_reference_model_predicts_on_the_gpu)
    # and back: the converted-back model predicts exactly like the original on the CPU
    back = interop.to_reference(g)
    assert type(back).__module__ == "dbn.models"
    assert back.predict(X) == r.predict(X)
    assert np.array_equal(back.predict_proba(X), r.predict_proba(X))


def test_reference_rbm_and_regressor_conversion(ref):
    from dbn_b200 import interop
    rng = np.random.default_rng(1)
    X = rng.random((40, 7)).astype(np.float32)
    np.random.seed(1)
    rb = ref.BinaryRBM(n_hidden_units=5, n_epochs=1, batch_size=8, verbose=False).fit(X)
    g = interop.from_reference(rb)
    assert g.get_params()["n_hidden_units"] == 5 and np.array_equal(g.W, rb.W) and g.n_visible_units == 7
    assert np.array_equal(interop.to_reference(g).transform(X), rb.transform(X))
    np.random.seed(2)
    rr = ref.SupervisedDBNRegression(hidden_layers_structure=[6], n_epochs_rbm=1, n_iter_backprop=2, batch_size=8,
                                     verbose=False)
    rr.fit(X, X @ rng.standard_normal(7))
    gr = interop.from_reference(rr)
    assert gr.W.shape == (1, 6) and np.array_equal(gr.W, rr.W)
    assert np.array_equal(interop.to_reference(gr).predict(X), rr.predict(X))
    with pytest.raises(TypeError):
        interop.from_reference(object())
