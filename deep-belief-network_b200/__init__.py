"""dbn_b200 -- B200-native backend for the training hot path of albertbup/deep-belief-network.

Import it next to the reference's backends:

    from dbn import SupervisedDBNClassification              # NumPy (reference)
    from dbn.tensorflow import SupervisedDBNClassification   # TensorFlow (reference)
    from dbn_b200 import SupervisedDBNClassification         # this package (hand-written sm_100a CUDA)

The directory is named `deep-belief-network_b200` (not an importable identifier); the `dbn_b200`
shim package at the repository root points its `__path__` here.
"""
from .models import (BinaryRBM, UnsupervisedDBN, SupervisedDBNClassification,  # noqa: F401
                     SupervisedDBNRegression)

__all__ = ["BinaryRBM", "UnsupervisedDBN", "SupervisedDBNClassification", "SupervisedDBNRegression"]
