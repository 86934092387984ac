"""Helpers for the GPU parity tests, smoke() and bench.py: drive the C-ABI (ctypes) with torch
tensors as device memory and return NumPy results.  Test infrastructure, not product code."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from dbn_b200 import _lib as L  # noqa: E402
from dbn_b200 import engine as E  # noqa: E402


def lib():
    return L.load()


# ---- Philox4x32-10 in NumPy: the exact generator of csrc/dbn_common.cuh -----------------------------
def _philox4x32_10(c0, c1, c2, c3, k0, k1):
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
    c0, c1, c2, c3 = [np.asarray(x, dtype=np.uint32) for x in (c0, c1, c2, c3)]
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def philox_uniforms(seed, stream, epoch, row0, rows, cols):
    """Uniforms the kernels draw for rows [row0, row0+rows) x cols: counter = (row, col/4, epoch,
    stream), key = seed; u = (x >> 8) * 2^-24 (float32-exact)."""
    r = (np.arange(rows, dtype=np.uint64) + np.uint64(row0)).astype(np.uint32)
    nq = (cols + 3) // 4
    q = np.arange(nq, dtype=np.uint32)
    R, Q = np.meshgrid(r, q, indexing="ij")
    o = _philox4x32_10(R, Q, np.full_like(R, epoch), np.full_like(R, stream), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    bits = np.stack(o, axis=-1).reshape(rows, nq * 4)[:, :cols]
    return ((bits >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0))


def count_flips(S_gpu, S_ref, U, P_ref, band):
    """Sample mismatches whose uniform is NOT within `band` of the reference probability (those
    inside the band may legitimately flip under fp32 / bf16 rounding of p)."""
    safe = np.abs(U - P_ref) >= band
    return int(np.sum((S_gpu != S_ref) & safe))


def guard_band_uniforms(U, P_ref, band, rng):
    """Re-draw uniforms that fall within `band` of the reference probability so that the
    comparison u < p is decided identically in fp32 and float64 (hard parts)."""
    U = U.copy()
    for _ in range(100):
        bad = np.abs(U - P_ref) < band
        if not bad.any():
            break
        U[bad] = rng.random(int(bad.sum()), dtype=np.float32)
    return U


# ---- device helpers ---------------------------------------------------------------------------------
def _t(a, dtype=None):
    import torch
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def stage_operand(X, precision):
    """fp32 [n, cols] -> operand layout of the precision (bf16: pitch with ones column)."""
    import torch
    Xd = _t(X, torch.float32)
    if precision == "fp32":
        return Xd
    n, cols = Xd.shape
    out = torch.zeros(n, E.bf16_pitch(cols), dtype=E.OP_DTYPE[E.PREC_IDS[precision]], device="cuda")
    L.check(lib().dbn_gather_rows(E.stream_ptr(), E.mat(Xd), None, n, cols, E.mat(out), cols, None))
    return out


def run_cd_step(W, b, c, V0, k, act, precision="fp32", seed=0, raw=False, U=None, lr_over_bs=0.0, epoch=0,
                row0=0, flags=0):
    """One dbn_rbm_cd_step.  Returns P0, H0s, Vk, Pk and either the raw deltas (raw=True) or the
    updated parameters."""
    import torch
    l = lib()
    H, V = W.shape
    B = V0.shape[0]
    layer = E.DeviceRbm(W, b, c, act, precision, torch.device("cuda"))
    Xop = stage_operand(V0, precision)
    ws = layer.make_workspace(B)
    a = layer._step_args(Xop, k, lr_over_bs, ws, seed, None, None)
    a.B, a.row0 = B, 0
    a.flags = flags
    a.rng.row0, a.rng.epoch = row0, epoch
    keep = []
    if U is not None:
        Ud = _t(np.asarray(U, dtype=np.float32).reshape(B, k * H), torch.float32)
        a.rng.injected = E.mat(Ud)
        keep.append(Ud)
    P0 = torch.zeros(B, H, device="cuda"); Vk = torch.zeros(B, V, device="cuda")
    Pk = torch.zeros(B, H, device="cuda"); H0s = torch.zeros(B, H, device="cuda")
    a.P0, a.Vk, a.Pk, a.H0s = E.mat(P0), E.mat(Vk), E.mat(Pk), E.mat(H0s)
    out = {}
    if raw:
        ld = (V + 1 + 3) // 4 * 4
        R = torch.zeros(H + 1, ld, device="cuda")
        a.raw_delta = E.mat(R)
    L.check(l.dbn_rbm_cd_step(E.stream_ptr(), C.byref(a)))
    torch.cuda.synchronize()
    out.update(P0=P0.double().cpu().numpy(), Vk=Vk.double().cpu().numpy(), Pk=Pk.double().cpu().numpy(),
               H0s=H0s.double().cpu().numpy())
    if raw:
        Rn = R.double().cpu().numpy()
        out.update(dW=Rn[:H, :V], dc=Rn[:H, V], db=Rn[H, :V])
    Wn, bn, cn = layer.export()
    out.update(W=Wn, b=bn, c=cn)
    if layer.Wb is not None:
        out["Wb"] = layer.Wb[:, :V].float().double().cpu().numpy()
    return out


def run_mlp_step(layers, Wo, bo, X, Y, act, head, precision="fp32", masks=None, keep_p=1.0, seed=0, raw=False,
                 lr_over_bs=0.0, decay=1.0, row0=0, epoch=0):
    """One dbn_mlp_train_step.  layers = [(W, c), ...].  Returns head outputs, loss rows and either
    raw gradients (list of (gW, gb)) or the updated parameters."""
    import torch
    dev = torch.device("cuda")
    B = X.shape[0]
    rbms = [E.DeviceRbm(W, np.zeros(W.shape[1]), c, act, precision, dev) for (W, c) in layers]
    mlp = E.DeviceMlp(rbms, Wo, bo, act, head, precision, dev)
    Xop = stage_operand(X, precision)
    Yd = _t(Y, torch.float32)
    ws = mlp.make_workspace(B)
    loss = torch.zeros(B, device="cuda")
    md = None
    if masks is not None:
        md = [_t(m, torch.float32) for m in masks]
    a = mlp.step_args(Xop, Yd, lr_over_bs, decay, keep_p, ws, seed, None, loss, md)
    a.B, a.row0 = B, 0
    a.rng.row0, a.rng.epoch = row0, epoch
    out_t = torch.zeros(B, mlp.n_out, device="cuda")
    a.out = E.mat(out_t)
    raws = []
    if raw:
        widths = [X.shape[1]] + [W.shape[0] for (W, _c) in layers]
        outs = [W.shape[0] for (W, _c) in layers] + [Wo.shape[0]]
        for l in range(len(outs)):
            ld = (widths[l] + 1 + 3) // 4 * 4
            R = torch.zeros(outs[l], ld, device="cuda")
            a.raw_grads[l] = E.mat(R)
            raws.append((R, widths[l]))
    L.check(lib().dbn_mlp_train_step(E.stream_ptr(), C.byref(a)))
    torch.cuda.synchronize()
    res = {"out": out_t.double().cpu().numpy(), "loss": loss.double().cpu().numpy()}
    if raw:
        res["grads"] = [(R.double().cpu().numpy()[:, :w], R.double().cpu().numpy()[:, w]) for (R, w) in raws]
    res["layers"] = [(r.export()[0], r.export()[2]) for r in rbms]
    res["Wo"], res["bo"] = mlp.export_head()
    return res


def shadow_round(W, precision="bf16"):
    """float64 array rounded the way the device rounds the 16-bit weight shadow of a tensor-core mode."""
    import torch
    return torch.as_tensor(np.asarray(W)).float().to(E.OP_DTYPE[E.PREC_IDS[precision]]).double().numpy()


def relF(a, b):
    """Norm-wise relative error ||a-b||_F / ||b||_F."""
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0))


def run_rbm_epoch(W, b, c, X, batch_size, k, act, precision, seed=0, lr=0.1, U=None, tc_epoch=True, epochs=1):
    """`epochs` passes of dbn_rbm_fit_epoch over the rows of X in their stored order (no shuffle), either as ONE
    persistent launch per epoch (tc_epoch=True, csrc/rbm_tc_epoch.cuh) or as the tiled four-launch steps.  Returns the
    trained parameters, the 16-bit shadow and the number of kernels the library launched."""
    import torch
    l = lib()
    H, V = W.shape
    n = X.shape[0]
    layer = E.DeviceRbm(W, b, c, act, precision, torch.device("cuda"))
    Xop = stage_operand(X, precision)
    ws = layer.make_workspace(min(batch_size, n))
    ctr = torch.zeros(1, dtype=torch.int32, device="cuda")
    Ud = None
    if U is not None:
        Ud = _t(np.asarray(U, dtype=np.float32).reshape(n, k * H), torch.float32)
    a = layer._step_args(Xop, k, lr / batch_size, ws, seed, ctr, Ud)
    prev = l.dbn_set_tc_epoch(1 if tc_epoch else 0)
    try:
        before = l.dbn_launch_count()
        for _ in range(epochs):
            L.check(l.dbn_rbm_fit_epoch(E.stream_ptr(), C.byref(a), n, batch_size))
            ctr.add_(1)
        torch.cuda.synchronize()
        launches = int(l.dbn_launch_count() - before)
    finally:
        l.dbn_set_tc_epoch(prev)
    Wn, bn, cn = layer.export()
    return {"W": Wn, "b": bn, "c": cn, "Wb": layer.Wb[:, :V].float().double().cpu().numpy(), "launches": launches}


def tc_epoch_plan(V, H, B):
    """Tile plan of the persistent epoch kernel as a dict (empty: the shape does not fit)."""
    out = (C.c_int32 * 16)()
    m = lib().dbn_rbm_tc_epoch_plan_info(V, H, B, out, 16)
    names = ["ctas", "hid_rows", "hid_units", "hid_tiles", "vis_rows", "vis_units", "ksplit", "kc_slabs", "vis_tiles",
             "dw_width", "dw_tiles", "smem", "vk_early", "p1_in_act", "scratch", "kv_slabs"]
    return {names[i]: int(out[i]) for i in range(min(m, 16))}


def run_mlp_epoch(layers, Wo, bo, X, Y, act, head, precision, batch_size, lr=0.1, l2=1.0, tc=True, iters=1, keep_p=1.0,
                  masks=None, seed=0):
    """`iters` passes of dbn_mlp_fit_epoch over the rows of X in their stored order (no shuffle; dropout when keep_p < 1:
    Philox keyed by `seed`, or the injected `masks` = [n x n_in, n x h_1, ...] of zeros / ones), either as
    ONE persistent launch per pass (tc=True, csrc/mlp_tc_epoch.cuh) or as the tiled launches per mini-batch.  Returns
    the trained parameters, the per-row losses of the last pass and the number of kernels the library launched."""
    import torch
    l = lib()
    dev = torch.device("cuda")
    n = X.shape[0]
    rbms = [E.DeviceRbm(W, np.zeros(W.shape[1]), c, act, precision, dev) for (W, c) in layers]
    mlp = E.DeviceMlp(rbms, Wo, bo, act, head, precision, dev)
    Xop = stage_operand(X, precision)
    Yd = _t(Y, torch.float32)
    ws = mlp.make_workspace(min(batch_size, n))
    loss = torch.zeros(n, device="cuda")
    iter_ctr = torch.zeros(1, dtype=torch.int32, device="cuda")
    masks_d = [_t(m, torch.float32) for m in masks] if masks is not None else None   # injected dropout masks, rows of X
    a = mlp.step_args(Xop, Yd, lr / batch_size, 1.0 - lr * l2 / n, keep_p, ws, seed, iter_ctr, loss, masks_d)
    prev = l.dbn_set_tc_finetune(1 if tc else 0)
    try:
        before = l.dbn_launch_count()
        for _ in range(iters):
            L.check(l.dbn_mlp_fit_epoch(E.stream_ptr(), C.byref(a), n, batch_size))
            iter_ctr.add_(1)
        torch.cuda.synchronize()
        launches = int(l.dbn_launch_count() - before)
    finally:
        l.dbn_set_tc_finetune(prev)
    res = {"layers": [(r.export()[0], r.export()[2]) for r in rbms], "loss": loss.double().cpu().numpy(), "launches": launches}
    res["Wo"], res["bo"] = mlp.export_head()
    res["shadows"] = [r.Wb[:, :r.V].float().double().cpu().numpy() for r in rbms]
    return res
