"""Device-side drivers of the hot path: the host loops of the reference
(`BinaryRBM._stochastic_gradient_descent` dbn/models.py:96-122, `UnsupervisedDBN.fit` :248-276,
`NumPyAbstractSupervisedDBN._stochastic_gradient_descent` :419-469, `_fine_tuning` :517-551)
re-expressed as: parameters resident in HBM, one CUDA graph per epoch made of C-ABI launches,
permutation indices streamed from pinned host memory.

PyTorch is used for device memory, streams, CUDA-graph capture and torch.distributed only; every
arithmetic kernel is in libdbn_b200.so.  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib as L

ACT_IDS = {"sigmoid": L.ACT_SIGMOID, "relu": L.ACT_RELU}
PREC_IDS = {"fp32": L.PREC_FP32, "bf16": L.PREC_BF16}
_CHUNK = 16384  # rows per inference / metric chunk


def _use_graphs():
    return os.environ.get("DBN_B200_NO_GRAPH", "0") != "1"


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def mat(t, ld=None):
    """dbn_mat view of a 2-D (or 1-D) torch tensor."""
    if t is None:
        return L.Mat(None, 0, L.F32)
    dt = L.BF16 if t.dtype == torch.bfloat16 else L.F32
    assert t.dtype in (torch.bfloat16, torch.float32), t.dtype
    if ld is None:
        ld = t.stride(0) if t.dim() == 2 else t.numel()
    return L.Mat(t.data_ptr(), int(ld), dt)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)


def bf16_pitch(cols):
    """Row pitch of a bf16 operand: room for the ones column, multiple of 8 elements (TMA: 16 B)."""
    return (cols + 1 + 7) // 8 * 8


def resolve_precision(precision, n_in, n_out, batch_size):
    """'auto' -> tcgen05 BF16 where the tiles are worth it, strict FP32 otherwise."""
    if precision not in ("auto", "fp32", "bf16"):
        raise ValueError("Invalid precision.")
    if precision != "auto":
        return precision
    return "bf16" if (n_in >= 128 and n_out >= 128 and batch_size >= 64) else "fp32"


def require_device():
    lib = L.load()
    if not torch.cuda.is_available() or lib.dbn_device_count() <= 0:
        raise L.DbnB200Error("dbn_b200 needs an sm_100 (B200) CUDA device; there is no CPU fallback")
    return lib


def epoch_order(n):
    """Composed double shuffle of the reference (dbn/models.py:106 then dbn/utils.py:12), drawn
    from the global np.random stream in the same order."""
    first = np.random.permutation(n)
    second = np.random.permutation(n)
    return first[second]


class _IndexFeeder:
    """Double-buffered pinned staging of per-epoch permutation indices."""

    def __init__(self, n, device):
        self.host = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(2)]
        self.events = [torch.cuda.Event() for _ in range(2)]
        self.used = [False, False]
        self.dev = torch.empty(n, dtype=torch.int32, device=device)
        self.slot = 0

    def push(self, order):
        s = self.slot
        if self.used[s]:
            self.events[s].synchronize()
        self.host[s].numpy()[:] = order
        self.dev.copy_(self.host[s], non_blocking=True)
        self.events[s].record()
        self.used[s] = True
        self.slot ^= 1
        return self.dev


class DeviceRbm:
    """One RBM layer resident on the device: fp32 masters W (H,V), b (V,), c (H,) and, in BF16
    mode, the bf16 shadow of W that the tcgen05 kernels read."""

    def __init__(self, W, b, c, activation, precision, device):
        self.lib = require_device()
        self.device = device
        self.H, self.V = int(W.shape[0]), int(W.shape[1])
        self.act = ACT_IDS[activation]
        self.prec_name = precision
        self.prec = PREC_IDS[precision]
        self.W = torch.as_tensor(np.ascontiguousarray(W), dtype=torch.float32).to(device).contiguous()
        self.b = torch.as_tensor(np.asarray(b), dtype=torch.float32).to(device).contiguous()
        self.c = torch.as_tensor(np.asarray(c), dtype=torch.float32).to(device).contiguous()
        self.Wb = None
        if self.prec == L.PREC_BF16:
            self.Wb = torch.zeros(self.H, bf16_pitch(self.V), dtype=torch.bfloat16, device=device)
            self.refresh_shadow()

    # -- parameters -------------------------------------------------------------------------------
    def refresh_shadow(self):
        if self.Wb is not None:
            L.check(self.lib.dbn_scale_cast(stream_ptr(), ptr(self.W), self.V, self.H, self.V, 1.0, mat(self.Wb)))

    def scale_hidden(self, s):
        """rbm.W, rbm.c *= s (dbn/models.py:533-535, :546-548)."""
        L.check(self.lib.dbn_scale_cast(stream_ptr(), ptr(self.W), self.V, self.H, self.V, float(s), mat(self.W)))
        L.check(self.lib.dbn_scale_cast(stream_ptr(), ptr(self.c), self.H, 1, self.H, float(s), mat(self.c, self.H)))
        self.refresh_shadow()

    def export(self):
        torch.cuda.synchronize()
        return (self.W.double().cpu().numpy(), self.b.double().cpu().numpy(), self.c.double().cpu().numpy())

    def w_operand(self):
        return mat(self.Wb) if self.prec == L.PREC_BF16 else mat(self.W)

    # -- operand staging ---------------------------------------------------------------------------
    def stage(self, Xd, idx=None, out=None, dropout=None):
        """Rows of Xd (fp32 or bf16) -> operand layout of this precision (bf16 + ones column in
        BF16 mode), optionally permuted by idx (device int32)."""
        n = int(idx.numel()) if idx is not None else int(Xd.shape[0])
        cols = int(Xd.shape[1])
        if out is None:
            if self.prec == L.PREC_BF16:
                out = torch.empty(n, bf16_pitch(cols), dtype=torch.bfloat16, device=self.device)
            else:
                out = torch.empty(n, cols, dtype=torch.float32, device=self.device)
        ones = cols if self.prec == L.PREC_BF16 else -1
        L.check(self.lib.dbn_gather_rows(stream_ptr(), mat(Xd), ptr(idx), n, cols, mat(out), ones,
                                         C.byref(dropout) if dropout is not None else None))
        return out

    # -- inference ----------------------------------------------------------------------------------
    def hidden_means(self, Xop, row0, n, out):
        """_compute_hidden_units_matrix (dbn/models.py:176-183) on staged rows."""
        L.check(self.lib.dbn_rbm_hidden_means(stream_ptr(), self.prec, mat(Xop), row0, n, self.V, self.H,
                                              self.w_operand(), ptr(self.c), self.act, mat(out)))

    def visible_means(self, Hop, n, out):
        """_compute_visible_units_matrix (dbn/models.py:195-201)."""
        L.check(self.lib.dbn_rbm_visible_means(stream_ptr(), self.prec, mat(Hop), n, self.V, self.H,
                                               self.w_operand(), ptr(self.b), self.act, mat(out)))

    def transform(self, Xd):
        """fp32 [N,V] device tensor -> fp32 [N,H] hidden means (BinaryRBM.transform, :77-86)."""
        n = int(Xd.shape[0])
        out = torch.empty(n, self.H, dtype=torch.float32, device=self.device)
        for s in range(0, n, _CHUNK):
            m = min(_CHUNK, n - s)
            xs = Xd[s:s + m]
            xop = self.stage(xs) if (self.prec == L.PREC_BF16 or xs.dtype != torch.float32) else xs
            self.hidden_means(xop, 0, m, out[s:s + m])
        return out

    def reconstruction_error(self, Xd):
        """mean_n sum_j (reconstruct(transform(x)) - x)^2 (dbn/models.py:212-220)."""
        n = int(Xd.shape[0])
        acc = torch.zeros(1, dtype=torch.float32, device=self.device)
        hdt = torch.bfloat16 if self.prec == L.PREC_BF16 else torch.float32
        hld = bf16_pitch(self.H) if self.prec == L.PREC_BF16 else self.H
        for s in range(0, n, _CHUNK):
            m = min(_CHUNK, n - s)
            xs = Xd[s:s + m]
            xop = self.stage(xs) if self.prec == L.PREC_BF16 else xs
            hm = torch.empty(m, hld, dtype=hdt, device=self.device)
            self.hidden_means(xop, 0, m, hm)
            rec = torch.empty(m, self.V, dtype=torch.float32, device=self.device)
            self.visible_means(hm, m, rec)
            L.check(self.lib.dbn_sum_sq_diff(stream_ptr(), mat(rec), mat(xs), 0, m, self.V, 1.0 / n, ptr(acc)))
        return float(acc.item())

    # -- training -------------------------------------------------------------------------------------
    def _step_args(self, Xop, k, lr_over_bs, ws, seed, epoch_ctr, U=None):
        a = L.RbmStepArgs()
        a.precision, a.act = self.prec, self.act
        a.V, a.H, a.k = self.V, self.H, int(k)
        a.lr_over_bs = float(lr_over_bs)
        a.X = mat(Xop)
        a.W, a.ldw, a.b, a.c = self.W.data_ptr(), self.V, self.b.data_ptr(), self.c.data_ptr()
        a.Wb = mat(self.Wb)
        a.rng.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        a.rng.epoch_ptr = epoch_ctr.data_ptr() if epoch_ctr is not None else None
        a.rng.keep_p = 1.0
        if U is not None:
            a.rng.injected = mat(U)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        return a

    def make_workspace(self, bmax):
        nbytes = self.lib.dbn_rbm_workspace_bytes(self.prec, bmax, self.V, self.H)
        ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=self.device)
        L.check(self.lib.dbn_rbm_workspace_init(stream_ptr(), self.prec, bmax, self.V, self.H, ptr(ws)))
        return ws

    def fit(self, Xd, n_epochs, batch_size, k, lr, seed, rng_mode="philox", verbose=None, dp=None):
        """BinaryRBM._stochastic_gradient_descent (dbn/models.py:96-122) on a device-resident fp32
        dataset Xd [N,V].  Each epoch = one CUDA graph: gather (double shuffle) + all mini-batch
        CD-k steps.  `verbose(epoch, err)` is called with the reconstruction error if given.
        dp = (rank, world): data-parallel over mini-batch rows with a sum all-reduce of the deltas."""
        n = int(Xd.shape[0])
        if n == 0 or n_epochs <= 0:
            return
        if dp is not None and dp[1] > 1:
            return self._fit_dp(Xd, n_epochs, batch_size, k, lr, seed, verbose, dp)
        feeder = _IndexFeeder(n, self.device)
        Xp = self.stage(Xd)  # allocates the epoch array; contents rewritten every epoch
        ws = self.make_workspace(min(batch_size, n))
        epoch_ctr = torch.zeros(1, dtype=torch.int32, device=self.device)
        U = None
        if rng_mode == "numpy":
            U = torch.empty(n, k * self.H, dtype=torch.float32, device=self.device)
        args = self._step_args(Xp, k, lr / batch_size, ws, seed, epoch_ctr, U)
        # warm-up launch of every kernel of the step outside graph capture (lr = 0: parameters unchanged)
        warm = self._step_args(Xp, k, 0.0, ws, seed, epoch_ctr, U)
        warm.B, warm.row0 = min(batch_size, n), 0
        self.stage(Xd[:warm.B], out=Xp[:warm.B])
        if U is not None:
            U[:warm.B].fill_(0.5)
        L.check(self.lib.dbn_rbm_cd_step(stream_ptr(), C.byref(warm)))
        torch.cuda.current_stream().synchronize()

        graph = None

        def run_epoch():
            self.stage(Xd, idx=feeder.dev, out=Xp)
            L.check(self.lib.dbn_rbm_fit_epoch(stream_ptr(), C.byref(args), n, batch_size))
            epoch_ctr.add_(1)

        for epoch in range(1, n_epochs + 1):
            feeder.push(epoch_order(n))
            if U is not None:  # replay the reference's uniform stream
                U.copy_(torch.from_numpy(np.random.random_sample((n, k * self.H)).astype(np.float32)))
            if _use_graphs():
                if graph is None:
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        run_epoch()
                graph.replay()
            else:
                run_epoch()
            if verbose is not None:
                verbose(epoch, self.reconstruction_error(Xd))
        torch.cuda.current_stream().synchronize()

    def _fit_dp(self, Xd, n_epochs, batch_size, k, lr, seed, verbose, dp):
        """Data-parallel epoch loop: every rank holds the full dataset and the same shuffle; rank r
        computes rows [r*bl, (r+1)*bl) of each global mini-batch (bl = batch_size / world), writes
        raw deltas, sum-all-reduces them (NCCL) and applies the identical update.  Philox counters use
        the global row index, so samples are independent of the GPU count."""
        import torch.distributed as dist
        rank, world = dp
        n = int(Xd.shape[0])
        if batch_size % world != 0 or n % batch_size != 0:
            raise ValueError("data-parallel fit needs batch_size % world == 0 and N % batch_size == 0")
        bl = batch_size // world
        steps = n // batch_size
        n_local = steps * bl
        feeder = _IndexFeeder(n_local, self.device)
        Xp = self.stage(Xd[:n_local])
        ws = self.make_workspace(bl)
        epoch_ctr = torch.zeros(1, dtype=torch.int32, device=self.device)
        ldr = (self.V + 1 + 3) // 4 * 4
        raw = torch.zeros(self.H + 1, ldr, dtype=torch.float32, device=self.device)
        args = self._step_args(Xp, k, lr / batch_size, ws, seed, epoch_ctr)
        args.raw_delta = mat(raw)
        args.B = bl
        for epoch in range(1, n_epochs + 1):
            order = epoch_order(n).reshape(steps, world, bl)[:, rank, :].reshape(-1)
            feeder.push(order)
            self.stage(Xd, idx=feeder.dev, out=Xp)
            for s in range(steps):
                args.row0 = s * bl
                args.rng.row0 = s * batch_size + rank * bl
                L.check(self.lib.dbn_rbm_cd_step(stream_ptr(), C.byref(args)))
                dist.all_reduce(raw, op=dist.ReduceOp.SUM)
                L.check(self.lib.dbn_rbm_apply_delta(stream_ptr(), self.V, self.H, lr / batch_size, ptr(raw), ldr,
                                                     ptr(self.W), self.V, ptr(self.b), ptr(self.c),
                                                     C.byref(mat(self.Wb))))
            epoch_ctr.add_(1)
            if verbose is not None:
                verbose(epoch, self.reconstruction_error(Xd))
        torch.cuda.current_stream().synchronize()


class DeviceMlp:
    """Stack of pre-trained layers + output head resident on the device for the supervised
    fine-tune (dbn/models.py:419-551)."""

    def __init__(self, layers, Wo, bo, activation, head, precision, device):
        self.lib = require_device()
        self.device = device
        self.layers = layers  # list[DeviceRbm]
        self.act = ACT_IDS[activation]
        self.head = L.HEAD_SOFTMAX if head == "softmax" else L.HEAD_LINEAR
        self.prec_name = precision
        self.prec = PREC_IDS[precision]
        for l in layers:
            assert l.prec == self.prec
        self.n_in = layers[0].V
        self.dims = [l.H for l in layers]
        self.n_out = int(Wo.shape[0])
        self.Wo = torch.as_tensor(np.ascontiguousarray(Wo), dtype=torch.float32).to(device).contiguous()
        self.bo = torch.as_tensor(np.asarray(bo), dtype=torch.float32).to(device).contiguous()

    def export_head(self):
        torch.cuda.synchronize()
        return self.Wo.double().cpu().numpy(), self.bo.double().cpu().numpy()

    def _dims_array(self):
        return (C.c_int32 * len(self.dims))(*self.dims)

    def make_workspace(self, bmax):
        nbytes = self.lib.dbn_mlp_workspace_bytes(self.prec, bmax, self.n_in, self._dims_array(), len(self.dims),
                                                  self.n_out)
        ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=self.device)
        L.check(self.lib.dbn_mlp_workspace_init(stream_ptr(), self.prec, bmax, self.n_in, self._dims_array(),
                                                len(self.dims), self.n_out, ptr(ws)))
        return ws

    def step_args(self, Xop, Yd, lr_over_bs, decay, keep_p, ws, seed, iter_ctr, loss_rows=None, masks=None):
        a = L.MlpStepArgs()
        a.precision, a.act, a.head = self.prec, self.act, self.head
        a.n_layers, a.n_in, a.n_out = len(self.layers), self.n_in, self.n_out
        for i, l in enumerate(self.layers):
            a.dims[i] = l.H
            a.W[i], a.ldw[i], a.c[i] = l.W.data_ptr(), l.V, l.c.data_ptr()
            a.Wb[i] = mat(l.Wb)
        a.Wo, a.ldwo, a.bo = self.Wo.data_ptr(), self.Wo.stride(0), self.bo.data_ptr()
        a.X = mat(Xop)
        a.Y, a.ldy = Yd.data_ptr(), Yd.stride(0)
        a.lr_over_bs, a.decay, a.keep_p = float(lr_over_bs), float(decay), float(keep_p)
        a.rng.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        a.rng.epoch_ptr = iter_ctr.data_ptr() if iter_ctr is not None else None
        a.rng.keep_p = float(keep_p)
        if masks is not None:
            for j, m in enumerate(masks):
                a.masks[j] = mat(m)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        a.loss_rows = loss_rows.data_ptr() if loss_rows is not None else None
        return a

    def fit(self, Xd, Yd, n_iter, batch_size, lr, l2, dropout_p, seed, rng_mode="philox", verbose=None):
        """NumPyAbstractSupervisedDBN._stochastic_gradient_descent (dbn/models.py:419-469).  The
        1/p pre-scale and p post-scale of _fine_tuning (:533-548) are done by the caller."""
        n = int(Xd.shape[0])
        if n == 0 or n_iter <= 0:
            return
        keep_p = 1.0 - dropout_p
        first = self.layers[0]
        feeder = _IndexFeeder(n, self.device)
        Xp = first.stage(Xd)
        Yp = torch.empty(n, self.n_out, dtype=torch.float32, device=self.device)
        ws = self.make_workspace(min(batch_size, n))
        iter_ctr = torch.zeros(1, dtype=torch.int32, device=self.device)
        loss_rows = torch.zeros(n, dtype=torch.float32, device=self.device)
        masks = None
        if rng_mode == "numpy" and dropout_p > 0:
            masks = [torch.empty(n, w, dtype=torch.float32, device=self.device) for w in [self.n_in] + self.dims]
        decay = 1.0 - (lr * l2) / n
        args = self.step_args(Xp, Yp, lr / batch_size, decay, keep_p, ws, seed, iter_ctr, loss_rows, masks)
        warm = self.step_args(Xp, Yp, 0.0, 1.0, keep_p, ws, seed, iter_ctr, loss_rows, masks)
        warm.B, warm.row0 = min(batch_size, n), 0
        first.stage(Xd[:warm.B], out=Xp[:warm.B])
        Yp[:warm.B].copy_(Yd[:warm.B])
        if masks is not None:
            for m in masks:
                m[:warm.B].fill_(1.0)
        L.check(self.lib.dbn_mlp_train_step(stream_ptr(), C.byref(warm)))
        torch.cuda.current_stream().synchronize()
        graph = None

        def run_iter():
            first.stage(Xd, idx=feeder.dev, out=Xp)
            L.check(self.lib.dbn_gather_rows(stream_ptr(), mat(Yd), ptr(feeder.dev), n, self.n_out, mat(Yp), -1, None))
            L.check(self.lib.dbn_mlp_fit_epoch(stream_ptr(), C.byref(args), n, batch_size))
            iter_ctr.add_(1)

        widths = [self.n_in] + self.dims
        for it in range(1, n_iter + 1):
            feeder.push(epoch_order(n))
            if masks is not None:  # reference draw order: per row, V then H_1.. (dbn/models.py:402, :409)
                flat = np.random.binomial(1, keep_p, (n, sum(widths))).astype(np.float32)
                o = 0
                for m, w in zip(masks, widths):
                    m.copy_(torch.from_numpy(np.ascontiguousarray(flat[:, o:o + w])))
                    o += w
            if _use_graphs():
                if graph is None:
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        run_iter()
                graph.replay()
            else:
                run_iter()
            if verbose is not None:
                scale = self.n_out if self.head == L.HEAD_SOFTMAX else 1.0  # dbn/models.py:450,468 quirk
                verbose(it, float(loss_rows.sum().item()) / n * scale)
        torch.cuda.current_stream().synchronize()

    def forward(self, Xd):
        """predict path: stacked transform (dbn/models.py:278-287) + output units (:607-615 / :705-711)."""
        n = int(Xd.shape[0])
        out = torch.empty(n, self.n_out, dtype=torch.float32, device=self.device)
        A = Xd
        for l in self.layers:
            A = l.transform(A)
        ops = L.Operands()
        ops.A, ops.B = mat(A), mat(self.Wo)
        ops.K = self.dims[-1]
        epi = L.ActEpilogue()
        epi.bias, epi.act = self.bo.data_ptr(), L.ACT_IDENTITY
        epi.rng.keep_p = 1.0
        epi.out = mat(out)
        L.check(self.lib.dbn_gemm_act(stream_ptr(), L.PREC_FP32, n, self.n_out, C.byref(ops), 1, C.byref(epi)))
        if self.head == L.HEAD_SOFTMAX:
            L.check(self.lib.dbn_head_delta(stream_ptr(), self.head, ptr(out), self.n_out, None, 0, n, self.n_out,
                                            None, 0, None))
        return out
