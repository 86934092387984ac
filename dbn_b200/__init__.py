"""Import shim: `dbn_b200` is the importable name of the package whose sources live in
`deep-belief-network_b200/` (a directory name that is not a Python identifier)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "deep-belief-network_b200")
__path__ = [_real]

from .models import (BinaryRBM, UnsupervisedDBN, SupervisedDBNClassification,  # noqa: E402,F401
                     SupervisedDBNRegression)

__all__ = ["BinaryRBM", "UnsupervisedDBN", "SupervisedDBNClassification", "SupervisedDBNRegression"]
