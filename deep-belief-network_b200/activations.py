"""Activation-function classes kept for attribute / pickle compatibility with the reference
(`rbm._activation_function_class`, dbn/activations.py).  The kernels implement sigmoid / relu and
their derivatives themselves (csrc/epilogue.cuh); nothing on the training path calls these."""
import numpy as np


class ActivationFunction(object):
    name = None

    @classmethod
    def function(cls, x):
        raise NotImplementedError

    @classmethod
    def prime(cls, x):
        raise NotImplementedError


class SigmoidActivationFunction(ActivationFunction):
    name = "sigmoid"

    @classmethod
    def function(cls, x):
        return 1.0 / (1.0 + np.exp(-np.asarray(x, dtype=np.float64)))

    @classmethod
    def prime(cls, a):
        """Derivative expressed in the activation a = sigmoid(x)."""
        a = np.asarray(a, dtype=np.float64)
        return a * (1.0 - a)


class ReLUActivationFunction(ActivationFunction):
    name = "relu"

    @classmethod
    def function(cls, x):
        return np.maximum(np.asarray(x, dtype=np.float64), 0.0)

    @classmethod
    def prime(cls, a):
        return (np.asarray(a) > 0).astype(int)
