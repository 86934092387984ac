"""ctypes binding of libdbn_b200.so (the C-ABI declared in include/dbn_b200.h).

The structures below mirror the header field for field.  The library is REQUIRED: importing the
estimators on a box where it has not been built raises, and every compute entry point returns
DBN_ERR_NO_DEVICE when no sm_100 device is present -- there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdbn_b200.so")

# enums (include/dbn_b200.h)
OK, ERR_INVALID_ARG, ERR_CUDA, ERR_NO_DEVICE, ERR_UNSUPPORTED = 0, 1, 2, 3, 4
ACT_IDENTITY, ACT_SIGMOID, ACT_RELU = 0, 1, 2
F32, BF16, F16 = 0, 1, 2
PREC_FP32, PREC_BF16, PREC_F16 = 0, 1, 2
MUL_NONE, MUL_PRIME_SIGMOID, MUL_PRIME_RELU = 0, 1, 2
RNG_NONE, RNG_SAMPLE, RNG_MASK = 0, 1, 2
HEAD_SOFTMAX, HEAD_LINEAR = 0, 1
STEP_SKIP_UPDATE = 1
STEP_SKIP_FIRST_HIDDEN = 2
STEP_BIAS_STATS = 4
MAX_LAYERS = 16
MAX_PEERS = 8
PEER_HANDLE_BYTES = 64
ABI_VERSION = 2


class Mat(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ld", C.c_int32), ("dtype", C.c_int32)]


class Rng(C.Structure):
    _fields_ = [("mode", C.c_int32), ("keep_p", C.c_float), ("seed", C.c_uint64), ("stream", C.c_uint32),
                ("epoch", C.c_uint32), ("epoch_ptr", C.c_void_p), ("row0", C.c_uint32), ("col0", C.c_uint32),
                ("injected", Mat)]


class Operands(C.Structure):
    _fields_ = [("A", Mat), ("B", Mat), ("transA", C.c_int32), ("transB", C.c_int32), ("K", C.c_int32),
                ("row0_A", C.c_int32), ("row0_B", C.c_int32), ("_pad", C.c_int32)]


class ColSum(C.Structure):
    _fields_ = [("acc", C.c_void_p), ("sign", C.c_float), ("ref_row0", C.c_int32), ("ref", Mat), ("square", C.c_int32),
                ("_pad", C.c_int32)]


class ActEpilogue(C.Structure):
    _fields_ = [("bias", C.c_void_p), ("act", C.c_int32), ("mul", C.c_int32), ("aux", Mat), ("rng", Rng),
                ("out", Mat), ("out2", Mat), ("sample", Mat), ("colsum", ColSum)]


class UpdateEpilogue(C.Structure):
    _fields_ = [("W", C.c_void_p), ("ldw", C.c_int32), ("_pad", C.c_int32), ("Wb", Mat), ("vecM", C.c_void_p),
                ("vecN", C.c_void_p), ("decay", C.c_float), ("scale", C.c_float), ("raw", Mat),
                ("statM", C.c_void_p), ("statN", C.c_void_p), ("raw_peers", C.c_void_p * MAX_PEERS),
                ("rows_per_peer", C.c_int32), ("raw_slot", C.c_int32)]


class DpExchange(C.Structure):
    _fields_ = [("n_peers", C.c_int32), ("rank", C.c_int32), ("V", C.c_int32), ("H", C.c_int32),
                ("rows_per_peer", C.c_int32), ("ld_recv", C.c_int32), ("recv", C.c_void_p * MAX_PEERS),
                ("stats", C.c_void_p * MAX_PEERS), ("flags", C.c_void_p * MAX_PEERS), ("Wb", Mat * MAX_PEERS)]


class RbmStepArgs(C.Structure):
    _fields_ = [("precision", C.c_int32), ("act", C.c_int32), ("B", C.c_int32), ("V", C.c_int32),
                ("H", C.c_int32), ("k", C.c_int32), ("lr_over_bs", C.c_float), ("row0", C.c_int32),
                ("X", Mat), ("W", C.c_void_p), ("ldw", C.c_int32), ("flags", C.c_int32), ("b", C.c_void_p),
                ("c", C.c_void_p), ("Wb", Mat), ("rng", Rng), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_size_t), ("ws_rows", C.c_int32), ("_pad2", C.c_int32),
                ("P0", Mat), ("Vk", Mat), ("Pk", Mat), ("H0s", Mat),
                ("raw_delta", Mat)]


class MlpStepArgs(C.Structure):
    _fields_ = [("precision", C.c_int32), ("act", C.c_int32), ("head", C.c_int32), ("n_layers", C.c_int32),
                ("B", C.c_int32), ("n_in", C.c_int32), ("n_out", C.c_int32),
                ("dims", C.c_int32 * MAX_LAYERS), ("W", C.c_void_p * MAX_LAYERS),
                ("ldw", C.c_int32 * MAX_LAYERS), ("c", C.c_void_p * MAX_LAYERS), ("Wb", Mat * MAX_LAYERS),
                ("Wo", C.c_void_p), ("ldwo", C.c_int32), ("_pad0", C.c_int32), ("bo", C.c_void_p),
                ("X", Mat), ("row0", C.c_int32), ("_pad1", C.c_int32),
                ("Y", C.c_void_p), ("ldy", C.c_int32), ("_pad2", C.c_int32),
                ("lr_over_bs", C.c_float), ("decay", C.c_float), ("keep_p", C.c_float), ("_pad3", C.c_int32),
                ("rng", Rng), ("masks", Mat * (MAX_LAYERS + 1)), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_size_t), ("ws_rows", C.c_int32), ("_pad4", C.c_int32),
                ("loss_rows", C.c_void_p), ("out", Mat),
                ("raw_grads", Mat * (MAX_LAYERS + 1)), ("aux_stream", C.c_void_p)]


# every symbol include/dbn_b200.h declares: name -> (restype, argtypes)
_P = C.POINTER
SYMBOLS = {
    "dbn_abi_version": (C.c_int, []),
    "dbn_last_error": (C.c_char_p, []),
    "dbn_device_count": (C.c_int, []),
    "dbn_launch_count": (C.c_uint64, []),
    "dbn_set_tail_split": (C.c_int, [C.c_int]),
    "dbn_tail_split_count": (C.c_uint64, []),
    "dbn_struct_size": (C.c_size_t, [C.c_int]),
    "dbn_debug_set_trace": (C.c_int, [C.c_void_p]),
    "dbn_set_sm_reserve": (C.c_int, [C.c_int]),
    "dbn_philox_uniforms_host": (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                           C.c_void_p]),
    "dbn_gemm_act": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _P(Operands), C.c_int, _P(ActEpilogue)]),
    "dbn_gemm_update": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P(Operands), C.c_int,
                                  _P(UpdateEpilogue)]),
    "dbn_rbm_hidden_means": (C.c_int, [C.c_void_p, C.c_int, _P(Mat), C.c_int, C.c_int, C.c_int, C.c_int, _P(Mat),
                                       C.c_void_p, C.c_int, _P(Mat)]),
    "dbn_rbm_visible_means": (C.c_int, [C.c_void_p, C.c_int, _P(Mat), C.c_int, C.c_int, C.c_int, _P(Mat),
                                        C.c_void_p, C.c_int, _P(Mat)]),
    "dbn_rbm_reconstruction_sq": (C.c_int, [C.c_void_p, C.c_int, _P(Mat), C.c_int, C.c_int, C.c_int, _P(Mat), C.c_void_p,
                                            C.c_int, _P(Mat), C.c_int, C.c_void_p]),
    "dbn_rbm_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "dbn_rbm_workspace_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "dbn_rbm_cd_step": (C.c_int, [C.c_void_p, _P(RbmStepArgs)]),
    "dbn_rbm_fit_epoch": (C.c_int, [C.c_void_p, _P(RbmStepArgs), C.c_int, C.c_int]),
    "dbn_rbm_tc_epoch_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "dbn_set_tc_epoch": (C.c_int, [C.c_int]),
    "dbn_set_tc_finetune": (C.c_int, [C.c_int]),
    "dbn_mlp_tc_epoch_plan": (C.c_int, [C.c_int, C.c_int, _P(C.c_int32), C.c_int, C.c_int, C.c_int]),
    "dbn_rbm_tc_epoch_plan_info": (C.c_int, [C.c_int, C.c_int, C.c_int, _P(C.c_int32), C.c_int]),
    "dbn_rbm_persistent_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "dbn_rbm_fit_epoch_persistent": (C.c_int, [C.c_void_p, _P(RbmStepArgs), C.c_void_p, C.c_int, C.c_int]),
    "dbn_rbm_first_hidden": (C.c_int, [C.c_void_p, _P(RbmStepArgs), C.c_int, C.c_int]),
    "dbn_rbm_delta_rows": (C.c_int, [C.c_void_p, _P(RbmStepArgs), C.c_int, C.c_int]),
    "dbn_rbm_apply_delta": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p,
                                      C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, _P(Mat)]),
    "dbn_rbm_workspace_stats": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "dbn_peer_alloc": (C.c_int, [C.c_size_t, _P(C.c_void_p), C.c_void_p]),
    "dbn_peer_free": (C.c_int, [C.c_void_p]),
    "dbn_peer_open": (C.c_int, [C.c_void_p, _P(C.c_void_p)]),
    "dbn_peer_close": (C.c_int, [C.c_void_p]),
    "dbn_rbm_delta_scatter": (C.c_int, [C.c_void_p, _P(RbmStepArgs), _P(DpExchange)]),
    "dbn_dp_signal": (C.c_int, [C.c_void_p, _P(DpExchange), C.c_void_p, C.c_uint64]),
    "dbn_dp_apply": (C.c_int, [C.c_void_p, _P(DpExchange), C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_uint64]),
    "dbn_dp_wait": (C.c_int, [C.c_void_p, _P(DpExchange), C.c_int, C.c_uint64]),
    "dbn_gather_rows_peer": (C.c_int, [C.c_void_p, _P(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                       _P(Mat)]),
    "dbn_mlp_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, _P(C.c_int32), C.c_int, C.c_int]),
    "dbn_mlp_workspace_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _P(C.c_int32), C.c_int, C.c_int,
                                         C.c_void_p]),
    "dbn_mlp_train_step": (C.c_int, [C.c_void_p, _P(MlpStepArgs)]),
    "dbn_mlp_fit_epoch": (C.c_int, [C.c_void_p, _P(MlpStepArgs), C.c_int, C.c_int]),
    "dbn_gather_rows": (C.c_int, [C.c_void_p, _P(Mat), C.c_void_p, C.c_int, C.c_int, _P(Mat), C.c_int, _P(Rng)]),
    "dbn_scale_cast": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, _P(Mat)]),
    "dbn_fill_ones_column": (C.c_int, [C.c_void_p, _P(Mat), C.c_int, C.c_int]),
    "dbn_sum_sq_diff": (C.c_int, [C.c_void_p, _P(Mat), _P(Mat), C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "dbn_head_delta": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_int, C.c_void_p]),
}

_lib = None


class DbnB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DbnB200Error(
            "libdbn_b200.so not found at %s -- build it with `python __graft_entry__.py build` "
            "(nvcc, sm_100a).  This backend has no CPU / PyTorch fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    if lib.dbn_abi_version() != ABI_VERSION:
        raise DbnB200Error("libdbn_b200.so ABI version mismatch; rebuild")
    for i, st in enumerate((Mat, Rng, Operands, ActEpilogue, UpdateEpilogue, RbmStepArgs, MlpStepArgs, ColSum, DpExchange)):
        if lib.dbn_struct_size(i) != C.sizeof(st):
            raise DbnB200Error("struct layout mismatch for %s: C %d vs ctypes %d"
                               % (st.__name__, lib.dbn_struct_size(i), C.sizeof(st)))
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        msg = load().dbn_last_error()
        raise DbnB200Error("dbn_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))
