"""Model interchange with the reference's NumPy backend.

The fitted state of every estimator is the same set of attributes in both packages (`W`, `b`, `c`,
`n_visible_units`, `rbm_layers`, `unsupervised_dbn`, head `W`/`b`, `num_classes`, label maps, `p`;
reference: dbn/models.py:55-66, :255-265, :296-325, :524-528, :575-584), so a model trained by one
backend can be served or fine-tuned by the other:

    ref_model = dbn.SupervisedDBNClassification.load('model.pkl')      # needs `dbn` importable
    gpu_model = dbn_b200.interop.from_reference(ref_model)             # same predictions, B200 kernels
    cpu_model = dbn_b200.interop.to_reference(gpu_model)               # back to the NumPy backend
"""
from __future__ import annotations

import numpy as np

from . import activations as A
from . import models as M

_RBM_PARAMS = ("n_hidden_units", "activation_function", "optimization_algorithm", "learning_rate", "n_epochs",
               "contrastive_divergence_iter", "batch_size", "verbose")
_DBN_PARAMS = ("hidden_layers_structure", "activation_function", "optimization_algorithm", "learning_rate_rbm",
               "n_epochs_rbm", "contrastive_divergence_iter", "batch_size", "verbose")
_SUP_PARAMS = ("n_iter_backprop", "l2_regularization", "learning_rate", "batch_size", "dropout_p", "verbose")
_RBM_FITTED = ("W", "b", "c", "n_visible_units")
_SUP_FITTED = ("W", "b", "num_classes", "label_to_idx_map", "idx_to_label_map")


def _copy(v):
    return np.array(v, dtype=np.float64) if isinstance(v, np.ndarray) else v


def _act_class(name, module):
    return {"sigmoid": module.SigmoidActivationFunction, "relu": module.ReLUActivationFunction}[name]


def _rbm(src, cls, act_module):
    dst = cls(**{k: getattr(src, k) for k in _RBM_PARAMS})
    for k in _RBM_FITTED:
        if hasattr(src, k):
            setattr(dst, k, _copy(getattr(src, k)))
    if hasattr(src, "W"):
        dst._activation_function_class = _act_class(src.activation_function, act_module)
    return dst


def _dbn(src, cls, rbm_cls, act_module):
    dst = cls(**{k: getattr(src, k) for k in _DBN_PARAMS})
    if src.rbm_layers is not None:
        dst.rbm_layers = [_rbm(r, rbm_cls, act_module) for r in src.rbm_layers]
    return dst


def _sup(src, cls, dbn_cls, rbm_cls, act_module):
    u = src.unsupervised_dbn
    kw = {k: getattr(u, k) for k in _DBN_PARAMS}
    kw.update({k: getattr(src, k) for k in _SUP_PARAMS})
    dst = cls(**kw)
    dst.unsupervised_dbn = _dbn(u, dbn_cls, rbm_cls, act_module)
    for k in _SUP_FITTED:
        if hasattr(src, k):
            setattr(dst, k, _copy(getattr(src, k)) if isinstance(getattr(src, k), np.ndarray) else getattr(src, k))
    return dst


def from_reference(est):
    """Reference (`dbn`) estimator -> equivalent `dbn_b200` estimator (deep copy of the fitted state)."""
    name = type(est).__name__
    if name == "BinaryRBM":
        return _rbm(est, M.BinaryRBM, A)
    if name == "UnsupervisedDBN":
        return _dbn(est, M.UnsupervisedDBN, M.BinaryRBM, A)
    if name == "SupervisedDBNClassification":
        return _sup(est, M.SupervisedDBNClassification, M.UnsupervisedDBN, M.BinaryRBM, A)
    if name == "SupervisedDBNRegression":
        return _sup(est, M.SupervisedDBNRegression, M.UnsupervisedDBN, M.BinaryRBM, A)
    raise TypeError("not a deep-belief-network estimator: %r" % (type(est),))


def to_reference(est):
    """`dbn_b200` estimator -> reference NumPy-backend estimator.  Requires `dbn` to be importable."""
    import dbn.activations as RA
    import dbn.models as R
    name = type(est).__name__
    if name == "BinaryRBM":
        return _rbm(est, R.BinaryRBM, RA)
    if name == "UnsupervisedDBN":
        return _dbn(est, R.UnsupervisedDBN, R.BinaryRBM, RA)
    if name in ("SupervisedDBNClassification", "SupervisedDBNRegression"):
        return _sup(est, getattr(R, name), R.UnsupervisedDBN, R.BinaryRBM, RA)
    raise TypeError("not a dbn_b200 estimator: %r" % (type(est),))


def load_reference_pickle(path):
    """Unpickle a model saved by the reference (`BaseModel.save`, dbn/models.py:11-16; needs `dbn`
    importable for the pickle's class references) and convert it."""
    import pickle
    with open(path, "rb") as fp:
        return from_reference(pickle.load(fp))
