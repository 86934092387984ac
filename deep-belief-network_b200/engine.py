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
import threading

import numpy as np
import torch

from . import _lib as L

ACT_IDS = {"sigmoid": L.ACT_SIGMOID, "relu": L.ACT_RELU}
PREC_IDS = {"fp32": L.PREC_FP32, "bf16": L.PREC_BF16, "fp16": L.PREC_F16}
TC_PRECS = (L.PREC_BF16, L.PREC_F16)     # the tensor-core modes: every operand buffer in one 16-bit type
OP_DTYPE = {L.PREC_BF16: torch.bfloat16, L.PREC_F16: torch.float16, L.PREC_FP32: torch.float32}
_CHUNK = 16384  # rows per inference / metric chunk


def _use_graphs():
    return os.environ.get("DBN_B200_NO_GRAPH", "0") != "1"


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def mat(t, ld=None):
    """dbn_mat view of a 2-D (or 1-D) torch tensor."""
    if t is None:
        return L.Mat(None, 0, L.F32)
    assert t.dtype in (torch.bfloat16, torch.float16, torch.float32), t.dtype
    dt = {torch.bfloat16: L.BF16, torch.float16: L.F16, torch.float32: L.F32}[t.dtype]
    if ld is None:
        ld = t.stride(0) if t.dim() == 2 else t.numel()
    return L.Mat(t.data_ptr(), int(ld), dt)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)


def bf16_pitch(cols):
    """Row pitch of a bf16 operand: room for the ones column, multiple of 64 elements so every row starts
    on a 128-byte line (TMA itself only needs 16 B)."""
    return (cols + 1 + 63) // 64 * 64


def resolve_precision(precision, n_in, n_out, batch_size, activation="relu", phase="finetune"):
    """'auto' -> a tensor-core mode where the tiles are worth it, strict FP32 otherwise.  Of the two 16-bit modes,
    fp16 (11 significant bits) serves sigmoid RBM pre-training, where every operand is a probability, a binary state
    or a weight; bf16 (fp32's range) serves relu layers and the back-propagated deltas of the fine-tune."""
    if precision not in ("auto", "fp32", "bf16", "fp16"):
        raise ValueError("Invalid precision.")
    if precision != "auto":
        return precision
    if not (n_in >= 128 and n_out >= 128 and batch_size >= 64):
        return "fp32"
    return "fp16" if (phase == "pretrain" and activation == "sigmoid") else "bf16"


def require_device():
    lib = L.load()
    if not torch.cuda.is_available() or lib.dbn_device_count() <= 0:
        raise L.DbnB200Error("dbn_b200 needs an sm_100 (B200) CUDA device; there is no CPU fallback")
    return lib


def dp_context():
    """(rank, world) when the process runs under torch.distributed with DBN_B200_DATA_PARALLEL=1,
    else None.  One process per GPU (torchrun); replicas-only configs simply leave the flag unset."""
    if os.environ.get("DBN_B200_DATA_PARALLEL", "0") != "1":
        return None
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    return dist.get_rank(), dist.get_world_size()


def dp_check_shapes(n, batch_size, world):
    if batch_size % world != 0 or n % batch_size != 0:
        raise ValueError("data-parallel fit needs batch_size % world == 0 and N % batch_size == 0")


def dp_local_rows(order, steps, world, bl, rank):
    """Rows of the (globally shuffled) epoch that `rank` computes: slice [rank*bl, (rank+1)*bl) of every
    global mini-batch, concatenated in step order."""
    return np.asarray(order).reshape(steps, world, bl)[:, rank, :].reshape(-1)


def dp_global_row0(step, batch_size, rank, bl):
    """Global epoch-row index of the first local row of `step` (the Philox counter base)."""
    return step * batch_size + rank * bl


def dp_chunks(H, n_chunks):
    """Row blocks (multiples of 256 rows = one CTA-pair tile) in which the raw delta is produced,
    all-reduced and applied."""
    if n_chunks <= 1:
        return [(0, H)]
    rows = max(256, (H + n_chunks - 1) // n_chunks // 256 * 256)
    return [(h0, min(H, h0 + rows)) for h0 in range(0, H, rows)]


def epoch_order(n):
    """Composed double shuffle of the reference (dbn/models.py:106 then dbn/utils.py:12), drawn
    from the global np.random stream in the same order."""
    first = np.random.permutation(n)
    second = np.random.permutation(n)
    return first[second]


class _PinnedPool:
    """Process-wide cache of pinned staging buffers.  cudaHostAlloc / cudaFreeHost cost milliseconds and
    synchronise the device, and a fit needs the same few sizes every time: buffers are handed out together
    with the event of their last asynchronous copy and come back when their user is done."""
    _free = []
    _lock = threading.Lock()

    @classmethod
    def take(cls, nbytes):
        nbytes = max(int(nbytes), 256)
        with cls._lock:
            pick = None
            for i, (buf, ev) in enumerate(cls._free):
                if nbytes <= buf.numel() <= 4 * nbytes + 65536:
                    idle = ev is None or ev.query()
                    if pick is None or (idle and not pick[1]) or (idle == pick[1] and buf.numel() < cls._free[pick[0]][0].numel()):
                        pick = (i, idle)
            if pick is not None and not pick[1] and nbytes <= (1 << 20):
                pick = None     # only busy candidates: a small buffer is cheaper to allocate than to wait for (its last
                                # copy may sit behind a multi-gigabyte upload in the copy engine's queue)
            if pick is not None:
                buf, ev = cls._free.pop(pick[0])
            else:
                buf, ev = None, None
        if buf is None:
            return torch.empty(nbytes, dtype=torch.uint8).pin_memory(), torch.cuda.Event()
        if ev is not None:
            ev.synchronize()      # normally complete long ago
        return buf, ev if ev is not None else torch.cuda.Event()

    @classmethod
    def give(cls, buf, ev):
        with cls._lock:
            if len(cls._free) < 64:
                cls._free.append((buf, ev))


_POOLED = {}     # id(ndarray) -> (ndarray, pinned buffer): host arrays living in pooled pinned memory (pooled_f32)


def pooled_f32(shape):
    """Float32 ndarray of `shape` in pooled pinned memory, or None when too many are outstanding.  For host data that is
    produced only to be uploaded (initial weights drawn straight in the device's form): the upload is one DMA from
    where the values were written, and `release_pooled` hands the memory back once that copy has been queued."""
    if len(_POOLED) >= 8:
        return None
    n = int(np.prod(shape))
    buf, _ev = _PinnedPool.take(n * 4)
    a = buf[:n * 4].view(torch.float32).view(*shape).numpy()
    _POOLED[id(a)] = (a, buf)
    return a


def release_pooled(a, device=None):
    """Give a pooled_f32 array back (after the copies that read it were queued on the upload stream); the array must not
    be touched afterwards.  False if `a` is not one."""
    ent = _POOLED.get(id(a)) if isinstance(a, np.ndarray) else None
    if ent is None or ent[0] is not a:
        return False
    del _POOLED[id(a)]
    ev = torch.cuda.Event()
    ev.record(_upload_stream(device if device is not None else torch.device("cuda", torch.cuda.current_device())))
    _PinnedPool.give(ent[1], ev)
    return True


_UPLOAD_STREAMS = {}


_STAGE_CHUNK = 32 << 20     # bytes of pinned staging per chunk of a large pageable upload


def _upload_stream(device):
    key = (device.type, device.index)
    up = _UPLOAD_STREAMS.get(key)
    if up is None:
        up = _UPLOAD_STREAMS[key] = torch.cuda.Stream(device=device)
    return up


def upload_f32(arr, device):
    """Host array -> new fp32 device tensor WITHOUT waiting for the work already queued on the compute stream:
    copied on a dedicated upload stream, and the compute stream is made to wait for that copy only.  (A pageable
    `.to(device)` blocks the host until everything queued before it has run, which serialises the layers of a
    stacked fit.)
      * float32, C-contiguous, page-locked source (a caller that pinned its dataset): one DMA straight from it;
      * anything else: converted chunk by chunk into two pooled pinned buffers -- the conversion of chunk i+1
        overlaps the DMA of chunk i, and no staging buffer of the array's size is ever pinned."""
    arr = np.asarray(arr)
    up = _upload_stream(device)
    cur = torch.cuda.current_stream(device)
    with torch.cuda.stream(up):
        dev = torch.empty(arr.shape, dtype=torch.float32, device=device)
    direct = None
    if arr.dtype == np.float32 and arr.flags.c_contiguous and arr.size > 0:
        t = torch.from_numpy(arr)
        if t.is_pinned():
            direct = t
    if direct is not None:
        with torch.cuda.stream(up):
            dev.copy_(direct, non_blocking=True)
    elif arr.size > 0:
        flat_src = arr.reshape(-1) if arr.flags.c_contiguous else np.ascontiguousarray(arr).reshape(-1)
        flat_dst = dev.view(-1)
        per = max(1, _STAGE_CHUNK // 4)
        n = flat_src.size
        bufs = [_PinnedPool.take(min(n, per) * 4) for _ in range(1 if n <= per else 2)]
        for i, s in enumerate(range(0, n, per)):
            buf, ev = bufs[i % len(bufs)]
            if i >= len(bufs):
                ev.synchronize()                      # the DMA that last read this buffer has finished
            m = min(per, n - s)
            stage = buf[:m * 4].view(torch.float32)
            np.copyto(stage.numpy(), flat_src[s:s + m], casting="unsafe")     # converts float64 -> float32 on the way
            with torch.cuda.stream(up):
                flat_dst[s:s + m].copy_(stage, non_blocking=True)
                ev.record(up)
        for buf, ev in bufs:
            _PinnedPool.give(buf, ev)
    cur.wait_stream(up)
    dev.record_stream(cur)        # allocated in the upload stream's pool, used on the compute stream
    return dev


def is_pinned_f32(arr):
    """Is this host array something the copy engine can read in place (float32, C-contiguous, page-locked)."""
    return (isinstance(arr, np.ndarray) and arr.dtype == np.float32 and arr.flags.c_contiguous and arr.size > 0 and
            torch.from_numpy(arr).is_pinned())


def operand_upload_wanted(precision, n, cols):
    """Should a dataset of n x cols host rows go up as 16-bit operands (upload_operand) rather than as fp32 rows: a
    tensor-core precision, the host helper present, and enough bytes for the PCIe time to matter."""
    from . import _hoststage
    precision = PREC_IDS.get(precision, precision)
    return (precision in TC_PRECS and n * cols >= (1 << 20) and os.environ.get("DBN_B200_OPERAND_UPLOAD", "1") != "0" and
            _hoststage.available())


def upload_operand(arr, device, precision):
    """Host rows -> new device tensor in the OPERAND layout of a tensor-core precision ([n, bf16_pitch(cols)] 16-bit:
    rounded row | ones column | zero padding), returned as its [n, cols] view.  Host threads round piece i+1 into pooled
    pinned memory exactly as the device's staging kernel would (csrc/host_stage.c) while piece i is in flight on the
    upload stream: half the PCIe bytes of upload_f32 -- the bound of an end-to-end fit of a large dataset -- and
    bit-identical operands, since the tensor-core path reads the rows through this rounding anyway.  The source needs
    no pinning (the threads read it with plain loads); like upload_f32 it does not wait for the compute stream."""
    from . import _hoststage
    arr = np.asarray(arr)
    n, cols = int(arr.shape[0]), int(arr.shape[1])
    dt = OP_DTYPE[PREC_IDS.get(precision, precision)]
    kind = _hoststage.kind_of(dt)
    pitch = bf16_pitch(cols)
    up = _upload_stream(device)
    cur = torch.cuda.current_stream(device)
    with torch.cuda.stream(up):
        dev = torch.empty(n, pitch, dtype=dt, device=device)
    fast = arr.dtype == np.float32 and arr.flags.c_contiguous
    rows_per = max(1, min(n, _OPERAND_PIECE // (2 * pitch)))
    bufs = [_PinnedPool.take(rows_per * pitch * 2) for _ in range(min(3, (n + rows_per - 1) // rows_per))]
    for i, s0 in enumerate(range(0, n, rows_per)):
        buf, ev = bufs[i % len(bufs)]
        if i >= len(bufs):
            ev.synchronize()                          # the DMA that last read this buffer has finished
        m = min(rows_per, n - s0)
        stage = buf[:m * pitch * 2].view(dt).view(m, pitch)
        src = arr[s0:s0 + m] if fast else np.ascontiguousarray(arr[s0:s0 + m], dtype=np.float32)
        _hoststage.stage_rows(src, None, stage.data_ptr(), pitch, kind, True)
        with torch.cuda.stream(up):
            dev[s0:s0 + m].copy_(stage, non_blocking=True)
            ev.record(up)
    for buf, ev in bufs:
        _PinnedPool.give(buf, ev)
    cur.wait_stream(up)
    dev.record_stream(cur)
    return dev[:, :cols]


_OPERAND_PIECE = 64 << 20   # bytes of 16-bit rows per piece of upload_operand (thread start per piece: 42 ms per 2 GB at 64 MB, 51 at 16)

_CAST_POOL = None


def download_f64(t):
    """fp32 device tensor -> float64 ndarray (the dtype of the reference's parameters): one DMA into pooled pinned
    memory, then the widening cast on a few host threads (NumPy releases the GIL inside the cast loop).  The
    obvious `t.double().cpu()` converts on the device and then moves twice the bytes through a pageable copy."""
    global _CAST_POOL
    n = t.numel()
    if n < (1 << 20):
        return t.detach().double().cpu().numpy()
    buf, ev = _PinnedPool.take(n * 4)
    stage = buf[:n * 4].view(torch.float32)
    stage.copy_(t.detach().reshape(-1), non_blocking=True)
    ev.record()
    out = np.empty(n, dtype=np.float64)
    ev.synchronize()
    src = stage.numpy()
    parts = 8
    if _CAST_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _CAST_POOL = ThreadPoolExecutor(max_workers=parts)
    step = (n + parts - 1) // parts

    def cast(i):
        out[i * step:(i + 1) * step] = src[i * step:(i + 1) * step]
    list(_CAST_POOL.map(cast, range(parts)))
    _PinnedPool.give(buf, ev)
    return out.reshape(tuple(t.shape))


class _IndexFeeder:
    """Double-buffered pinned staging of per-epoch permutation indices (the host runs at most two epochs
    ahead of the device)."""

    def __init__(self, n, device):
        pairs = [_PinnedPool.take(4 * n) for _ in range(2)]
        self._bufs = [p[0] for p in pairs]
        self.host = [b[:4 * n].view(torch.int32) for b in self._bufs]
        self.events = [p[1] for p in pairs]
        self.used = [False, False]
        self.dev = torch.empty(n, dtype=torch.int32, device=device)
        self.slot = 0

    def __del__(self):
        try:
            for b, e in zip(self._bufs, self.events):
                _PinnedPool.give(b, e)
        except Exception:      # interpreter shutdown
            pass

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
        self.W = upload_f32(W, device)
        self.b = upload_f32(b, device)
        self.c = upload_f32(c, device)
        self.Wb = None
        if self.prec in TC_PRECS:
            self.Wb = torch.zeros(self.H, bf16_pitch(self.V), dtype=OP_DTYPE[self.prec], device=device)
            self.refresh_shadow()

    @classmethod
    def blank(cls, H, V, activation, precision, device):
        """Layer with uninitialised parameters allocated on the device (a data-parallel rank that receives its
        initial weights by broadcast instead of drawing and uploading them)."""
        self = cls.__new__(cls)
        self.lib = require_device()
        self.device = device
        self.H, self.V = int(H), int(V)
        self.act = ACT_IDS[activation]
        self.prec_name = precision
        self.prec = PREC_IDS[precision]
        self.W = torch.empty(self.H, self.V, dtype=torch.float32, device=device)
        self.b = torch.empty(self.V, dtype=torch.float32, device=device)
        self.c = torch.empty(self.H, dtype=torch.float32, device=device)
        self.Wb = None
        if self.prec in TC_PRECS:
            self.Wb = torch.zeros(self.H, bf16_pitch(self.V), dtype=OP_DTYPE[self.prec], device=device)
        return self

    def clone_as(self, precision):
        """Independent copy of this layer on the device, in `precision` (fp32 masters copied device to device, the 16-bit
        shadow rounded afresh): how a pre-trained layer enters the fine-tune without a round trip through the host."""
        act = next(k for k, v in ACT_IDS.items() if v == self.act)
        o = DeviceRbm.blank(self.H, self.V, act, precision, self.device)
        o.W.copy_(self.W)
        o.b.copy_(self.b)
        o.c.copy_(self.c)
        o.refresh_shadow()
        return o

    def move_shadow(self, Wb):
        """Keep the bf16 shadow in caller-provided memory (a peer-visible arena) from now on."""
        assert self.prec in TC_PRECS and tuple(Wb.shape) == (self.H, bf16_pitch(self.V)) and Wb.dtype == OP_DTYPE[self.prec]
        Wb.zero_()
        self.Wb = Wb
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
        torch.cuda.current_stream().synchronize()
        return (download_f64(self.W), download_f64(self.b), download_f64(self.c))

    def w_operand(self):
        return mat(self.Wb) if self.prec in TC_PRECS else mat(self.W)

    # -- operand staging ---------------------------------------------------------------------------
    def alloc_staged(self, n, cols):
        """Operand buffer for n rows of `cols` features in this precision's layout (BF16: zero-filled, since
        the padding columns are read by the tensor-core path; FP32: uninitialised)."""
        if self.prec in TC_PRECS:   # zero padding: the tensor-core path reads K tails up to the next 64 columns
            return torch.zeros(n, bf16_pitch(cols), dtype=OP_DTYPE[self.prec], device=self.device)
        return torch.empty(n, cols, dtype=torch.float32, device=self.device)

    def stage(self, Xd, idx=None, out=None, dropout=None):
        """Rows of Xd (fp32 or bf16) -> operand layout of this precision (bf16 + ones column in
        BF16 mode), optionally permuted by idx (device int32)."""
        n = int(idx.numel()) if idx is not None else int(Xd.shape[0])
        cols = int(Xd.shape[1])
        if out is None:
            out = self.alloc_staged(n, cols)
        ones = cols if self.prec in TC_PRECS else -1
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

    def transform(self, Xd, out=None):
        """fp32 [N,V] device tensor -> fp32 [N,H] hidden means (BinaryRBM.transform, :77-86)."""
        n = int(Xd.shape[0])
        if out is None:
            out = torch.empty(n, self.H, dtype=torch.float32, device=self.device)
        for s in range(0, n, _CHUNK):
            m = min(_CHUNK, n - s)
            xs = Xd[s:s + m]
            xop = self.stage(xs) if (self.prec in TC_PRECS or xs.dtype != torch.float32) else xs
            self.hidden_means(xop, 0, m, out[s:s + m])
        return out

    def reconstruction_sq(self, Hop, n, Xd, row0, acc):
        """acc[j] += sum_rows (x - v'(h))^2 in fixed point, taken in the epilogue that produces v' (no reconstruction
        buffer, no reduction kernel)."""
        L.check(self.lib.dbn_rbm_reconstruction_sq(stream_ptr(), self.prec, mat(Hop), n, self.V, self.H, self.w_operand(),
                                                   ptr(self.b), self.act, mat(Xd), row0, ptr(acc)))

    def reconstruction_error(self, Xd):
        """mean_n sum_j (reconstruct(transform(x)) - x)^2 (dbn/models.py:212-220)."""
        n = int(Xd.shape[0])
        acc = torch.zeros(self.V, dtype=torch.int64, device=self.device)
        hdt = OP_DTYPE[self.prec]
        hld = bf16_pitch(self.H) if self.prec in TC_PRECS else self.H
        for s in range(0, n, _CHUNK):
            m = min(_CHUNK, n - s)
            xs = Xd[s:s + m]
            xop = self.stage(xs) if self.prec in TC_PRECS else xs
            hm = torch.zeros(m, hld, dtype=hdt, device=self.device)   # operand of the next contraction: finite padding
            self.hidden_means(xop, 0, m, hm)
            self.reconstruction_sq(hm, m, xs, 0, acc)
        return float(acc.sum().item()) / 4294967296.0 / n

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
        a.ws_rows = getattr(ws, "dbn_rows", 0)
        return a

    def make_workspace(self, bmax):
        nbytes = self.lib.dbn_rbm_workspace_bytes(self.prec, bmax, self.V, self.H)
        ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=self.device)
        L.check(self.lib.dbn_rbm_workspace_init(stream_ptr(), self.prec, bmax, self.V, self.H, ptr(ws)))
        ws.dbn_rows = int(bmax)   # the layout (and its ones columns) is fixed by this row count
        return ws

    def fit(self, Xd, n_epochs, batch_size, k, lr, seed, rng_mode="philox", verbose=None, orders=None, sync=True):
        """BinaryRBM._stochastic_gradient_descent (dbn/models.py:96-122) on a device-resident fp32
        dataset Xd [N,V].  Each epoch = one CUDA graph: gather (double shuffle) + all mini-batch
        CD-k steps.  `verbose(epoch, err)` is called with the reconstruction error if given.
        orders: object whose get() returns the next epoch's row order (drawn ahead by the caller) instead of
        drawing it here; sync=False returns with the work queued on the current stream.
        (Data-parallel training: fit_rbm_data_parallel below.)"""
        n = int(Xd.shape[0])
        if n == 0 or n_epochs <= 0:
            return
        runner = RbmEpochRunner(self, Xd, batch_size, k, lr, seed, rng_mode)
        for epoch in range(1, n_epochs + 1):
            runner.run(orders.get() if orders is not None else None)
            if verbose is not None:
                verbose(epoch, self.reconstruction_error(Xd))
        if sync:
            torch.cuda.current_stream().synchronize()


def dp_begin_upload(X_host, precision, H, dp, device, quiet=True):
    """Start the upload of this rank's share of the dataset BEFORE the initial parameters exist: rank 0 spends ~20 ms
    drawing a 4096 x 4096 layer and the other ranks wait for its broadcast -- the rows cross PCIe meanwhile.  Returns
    the device tensor fit_rbm_data_parallel takes as `Xl` (the rank's 1/world of the rows when the peer-memory path is
    the one that will run), or None."""
    from . import dp_peer
    rank, world = dp
    n, V = int(X_host.shape[0]), int(X_host.shape[1])
    if n == 0 or n % world or not dp_peer.peer_path_wanted_for(precision, H, world):
        return None
    nl = n // world
    mine = X_host[rank * nl:(rank + 1) * nl]
    if quiet and not is_pinned_f32(mine) and operand_upload_wanted(precision, nl, V):
        return upload_operand(mine, device, precision)     # rounded on the host: half the PCIe bytes
    return upload_f32(mine, device)


def fit_rbm_data_parallel(layer, X_host, n_epochs, batch_size, k, lr, seed, dp, next_order, verbose=None, Xl=None):
    """Data-parallel BinaryRBM._stochastic_gradient_descent (dbn/models.py:96-122) of one layer whose parameters are
    already identical on every rank.  X_host: the whole dataset on the host (float32 [n, V], the same on every rank);
    next_order(): on rank 0 returns the next epoch's row order (the reference's double shuffle), ignored elsewhere --
    the order is broadcast, so the ranks cannot disagree about which rows a mini-batch holds.

    BF16 layers with H % world == 0 take the peer-memory path (dp_peer.DpPeerRunner): a rank uploads only its
    1/world of the rows and no NCCL call sits on the data path.  Everything else takes DpRbmEpochRunner (NCCL
    reduce-scatter / all-reduce), for which every rank uploads the whole dataset."""
    import torch.distributed as dist
    from . import dp_peer
    rank, world = dp
    n, V = int(X_host.shape[0]), int(X_host.shape[1])
    dp_check_shapes(n, batch_size, world)
    dev = layer.device
    runner, group = None, None
    if dp_peer.peer_path_wanted(layer, world):
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            nl = n // world
            group = dp_peer.acquire_group(rank, world, dp_peer.ArenaLayout(world, V, layer.H, nl, OP_DTYPE[layer.prec]), dev)
        except Exception as e:            # no CUDA IPC in this container, peers not reachable, ...
            group = None
            ok.zero_()
            reason = str(e)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)   # all ranks take the same path
        if int(ok.item()) == 1:
            if Xl is None or int(Xl.shape[0]) != nl:
                Xl = dp_begin_upload(X_host, layer.prec, layer.H, dp, dev, quiet=verbose is None)
            runner = dp_peer.DpPeerRunner(layer, Xl, n, batch_size, k, lr, seed, group.member)
        else:
            if group is not None:
                dp_peer.release_groups()
                group = None
            if rank == 0:
                import warnings
                warnings.warn("dbn_b200: peer-memory exchange unavailable, falling back to NCCL collectives")
    Xd = None
    if runner is None:
        Xd = upload_f32(X_host, dev)
        runner = DpRbmEpochRunner(layer, Xd, batch_size, k, lr, seed, dp)
    order_dev = torch.empty(n, dtype=torch.int32, device=dev)
    try:
        for epoch in range(1, n_epochs + 1):
            if rank == 0:
                order_dev.copy_(torch.from_numpy(np.ascontiguousarray(next_order(), dtype=np.int32)))
            dist.broadcast(order_dev, 0)
            runner.run(order_dev.cpu().numpy())
            if verbose is not None:
                if Xd is None:
                    Xd = upload_f32(X_host, dev)
                verbose(epoch, layer.reconstruction_error(Xd))
        torch.cuda.current_stream().synchronize()
    finally:
        if group is not None and layer.Wb is not None:
            # the arena stays cached for the next fit (dp_peer.acquire_group): the layer takes a private copy of its shadow
            torch.cuda.current_stream().synchronize()
            layer.Wb = layer.Wb.clone()
    return runner


class RbmEpochRunner:
    """One RBM layer + one dataset prepared for epoch-at-a-time training.  `run()` = one epoch of
    dbn/models.py:105-119, executed as
      * ONE launch of the persistent on-chip cluster kernel when the layer fits (strict FP32), or
      * the tiled path: gather (double shuffle) + four contractions per mini-batch, launched eagerly for the
        first epoch (which doubles as warm-up) and as a captured CUDA graph from the second epoch on."""

    def __init__(self, layer, Xd, batch_size, k, lr, seed, rng_mode="philox"):
        self.layer, self.Xd = layer, Xd
        self.n, self.bs, self.k = int(Xd.shape[0]), int(batch_size), int(k)
        n, lib = self.n, layer.lib
        self.feeder = _IndexFeeder(n, layer.device)
        self.epoch_ctr = torch.zeros(1, dtype=torch.int32, device=layer.device)
        self.U = None
        if rng_mode == "numpy":
            self.U = torch.empty(n, k * layer.H, dtype=torch.float32, device=layer.device)
        # layers that fit on chip: one persistent cluster kernel per epoch, W resident in shared memory
        self.persistent = (os.environ.get("DBN_B200_PERSISTENT", "1") != "0" and Xd.dtype == torch.float32 and
                           Xd.is_contiguous() and
                           lib.dbn_rbm_persistent_plan(layer.prec, layer.V, layer.H, min(self.bs, n)) > 0)
        # tensor-core modes at batch <= 256: dbn_rbm_fit_epoch is ONE cooperative launch of the persistent tcgen05 kernel
        # (csrc/rbm_tc_epoch.cuh) -- nothing to capture, and a cooperative launch is not captured anyway
        self.tc_epoch = (not self.persistent and layer.prec in TC_PRECS and
                         lib.dbn_rbm_tc_epoch_plan(layer.prec, layer.V, layer.H, min(self.bs, n)) > 0)
        if self.persistent:   # reads the un-permuted dataset through the index array: no epoch array, no workspace
            self.Xp = Xd
            self.ws = torch.empty(16, dtype=torch.uint8, device=layer.device)
        else:
            self.Xp = layer.alloc_staged(n, int(Xd.shape[1]))  # epoch array; (re)written by the gather every epoch
            self.ws = layer.make_workspace(min(self.bs, n))
        self.args = layer._step_args(self.Xp, k, lr / batch_size, self.ws, seed, self.epoch_ctr, self.U)
        self.graph = None
        self.launches_per_epoch = 0
        self.epochs_run = 0

    def _enqueue(self):
        lib = self.layer.lib
        if self.persistent:
            L.check(lib.dbn_rbm_fit_epoch_persistent(stream_ptr(), C.byref(self.args), ptr(self.feeder.dev), self.n,
                                                     min(self.bs, self.n)))
        else:
            self.layer.stage(self.Xd, idx=self.feeder.dev, out=self.Xp)
            L.check(lib.dbn_rbm_fit_epoch(stream_ptr(), C.byref(self.args), self.n, self.bs))
        self.epoch_ctr.add_(1)

    def run(self, order=None):
        """One epoch.  `order` defaults to the reference's double shuffle drawn from np.random."""
        layer = self.layer
        self.feeder.push(epoch_order(self.n) if order is None else order)
        if self.U is not None:  # replay the reference's uniform stream
            self.U.copy_(torch.from_numpy(np.random.random_sample((self.n, self.k * layer.H)).astype(np.float32)))
        # Epoch 1 runs eagerly (it doubles as the warm-up that loads every kernel before capture, and a
        # one-epoch fit never pays for capture + instantiation); epoch 2 is captured; later epochs replay.
        if _use_graphs() and self.epochs_run >= 1 and not self.persistent and not self.tc_epoch:
            if self.graph is None:
                torch.cuda.current_stream().synchronize()
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._enqueue()
            self.graph.replay()
        else:
            before = layer.lib.dbn_launch_count()
            self._enqueue()
            self.launches_per_epoch = int(layer.lib.dbn_launch_count() - before)
        self.epochs_run += 1


def resident_budget_bytes(device):
    """How many bytes of dataset (fp32 rows + their staged 16-bit epoch copy) a fit may keep in HBM: the
    DBN_B200_RESIDENT_BYTES override, else 60 % of the memory currently free on the device."""
    env = os.environ.get("DBN_B200_RESIDENT_BYTES")
    if env:
        return int(float(env))
    free, _total = torch.cuda.mem_get_info(device)
    return int(0.6 * free)


def parallel_take(X, rows, out):
    """out[i] = X[rows[i]] on a few host threads (the reference's `data = _data[idx]`, dbn/models.py:106-107 and
    dbn/utils.py:12-13, for one chunk of the epoch; NumPy releases the GIL inside take)."""
    global _CAST_POOL
    m = len(rows)
    if m * X.shape[1] < (1 << 18):
        np.take(X, rows, axis=0, out=out[:m], mode="clip")
        return
    if _CAST_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _CAST_POOL = ThreadPoolExecutor(max_workers=8)
    parts = 8
    step = (m + parts - 1) // parts

    def job(i):
        a, b = i * step, min(m, (i + 1) * step)
        if a < b:
            np.take(X, rows[a:b], axis=0, out=out[a:b], mode="clip")
    list(_CAST_POOL.map(job, range(parts)))


class HostChunkFeeder:
    """Software pipeline that brings a HOST dataset (and optional row-aligned companions, e.g. the targets) to the
    device chunk by chunk: host threads gather chunk c+1 of the given row order into pinned buffers, the upload stream
    copies it into one of two device buffers, and the compute stream works on chunk c meanwhile."""

    def __init__(self, arrays, chunk_rows, device):
        self.arrays = [np.ascontiguousarray(a, dtype=np.float32) for a in arrays]
        self.n = int(self.arrays[0].shape[0])
        self.R = int(chunk_rows)
        self.device = device
        self.pinned, self.host, self.dev = [], [], []
        for a in self.arrays:
            w = int(a.shape[1])
            bufs = [_PinnedPool.take(self.R * w * 4)[0] for _ in range(2)]
            self.pinned.append(bufs)
            self.host.append([b[:self.R * w * 4].view(torch.float32).view(self.R, w) for b in bufs])
            self.dev.append([torch.empty(self.R, w, dtype=torch.float32, device=device) for _ in range(2)])
        self.ev_up = [torch.cuda.Event() for _ in range(2)]
        self.ev_done = [torch.cuda.Event() for _ in range(2)]
        self.used = [False, False]

    def __del__(self):
        try:
            for bufs in self.pinned:
                for b in bufs:
                    _PinnedPool.give(b, None)
        except Exception:
            pass

    def iterate(self, order, visit):
        """visit(slot, s0, m) is enqueued once rows order[s0:s0+m] (or s0..s0+m if order is None) of every array are
        in dev[i][slot][:m]."""
        up = _upload_stream(self.device)
        cur = torch.cuda.current_stream()
        for c, s0 in enumerate(range(0, self.n, self.R)):
            m = min(self.R, self.n - s0)
            slot = c & 1
            if self.used[slot]:
                self.ev_up[slot].synchronize()       # the H2D that last read these pinned buffers
            for a, host in zip(self.arrays, self.host):
                if order is None:
                    host[slot].numpy()[:m] = a[s0:s0 + m]
                else:
                    parallel_take(a, order[s0:s0 + m], host[slot].numpy())
            with torch.cuda.stream(up):
                if self.used[slot]:
                    up.wait_event(self.ev_done[slot])   # the kernels that last read these device buffers
                for host, dev in zip(self.host, self.dev):
                    dev[slot][:m].copy_(host[slot][:m], non_blocking=True)
                self.ev_up[slot].record(up)
            cur.wait_event(self.ev_up[slot])
            visit(slot, s0, m)
            self.ev_done[slot].record(cur)
            self.used[slot] = True


def host_chunk_rows(batch_size, n, row_bytes, chunk_rows=None):
    """Rows per chunk of a host-streamed epoch: ~64 MB of fp32 rows, a whole number of mini-batches (at least 4)."""
    if chunk_rows is None:
        chunk_rows = max(4 * batch_size, (64 << 20) // max(1, row_bytes))
    env = os.environ.get("DBN_B200_CHUNK_ROWS")
    if env:
        chunk_rows = int(env)
    R = max(batch_size, chunk_rows // batch_size * batch_size)
    return min(R, (n + batch_size - 1) // batch_size * batch_size)


class HostRbmEpochRunner:
    """Epoch-at-a-time training of one RBM layer on a dataset that STAYS ON THE HOST (it does not fit in HBM, or the
    caller caps the resident bytes): the staging pipeline behind SURVEY 8f-4.  Per epoch the host permutation (the
    reference's double shuffle, dbn/models.py:106-107 + dbn/utils.py:12-13) is applied chunk by chunk:

        host threads   gather chunk c+1 of the shuffled epoch into a pinned buffer        (parallel_take)
        upload stream  H2D of chunk c+1 into one of two device buffers
        compute stream stage chunk c (cast + ones column), then its mini-batch CD steps   (dbn_rbm_fit_epoch)

    so the kernels consume chunk c while chunk c+1 lands.  Chunks are whole numbers of mini-batches and the Philox
    counters carry the GLOBAL epoch row, so the trained parameters are bit-identical to the resident path's."""

    def __init__(self, layer, X_host, batch_size, k, lr, seed, chunk_rows=None):
        self.layer = layer
        self.n, self.V = int(X_host.shape[0]), int(X_host.shape[1])
        self.bs, self.k = int(batch_size), int(k)
        dev = layer.device
        self.R = host_chunk_rows(self.bs, self.n, 4 * self.V, chunk_rows)
        self.feed = HostChunkFeeder([X_host], self.R, dev)
        self.dev32 = self.feed.dev[0]
        self.Xp = [layer.alloc_staged(self.R, self.V) for _ in range(2)]
        self.ws = layer.make_workspace(min(self.bs, self.n))
        self.epoch_ctr = torch.zeros(1, dtype=torch.int32, device=dev)
        self.args = [layer._step_args(xp, k, lr / batch_size, self.ws, seed, self.epoch_ctr) for xp in self.Xp]
        self.launches_per_epoch = 0

    def run(self, order=None):
        """One epoch (dbn/models.py:105-119)."""
        layer, lib = self.layer, self.layer.lib
        if order is None:
            order = epoch_order(self.n)
        order = np.asarray(order)
        before = lib.dbn_launch_count()

        def visit(slot, s0, m):
            layer.stage(self.dev32[slot][:m], out=self.Xp[slot][:m])
            a = self.args[slot]
            a.row0 = 0
            a.rng.row0 = s0                           # global epoch row of the chunk's first row
            L.check(lib.dbn_rbm_fit_epoch(stream_ptr(), C.byref(a), m, self.bs))
        self.feed.iterate(order, visit)
        self.epoch_ctr.add_(1)
        self.launches_per_epoch = int(lib.dbn_launch_count() - before)

    def reconstruction_error(self):
        """_compute_reconstruction_error (dbn/models.py:212-220) over the host dataset, chunk by chunk."""
        layer = self.layer
        acc = torch.zeros(layer.V, dtype=torch.int64, device=layer.device)
        tc = layer.prec in TC_PRECS
        hdt, hld = OP_DTYPE[layer.prec], (bf16_pitch(layer.H) if tc else layer.H)

        def visit(slot, s0, m):
            xs = self.dev32[slot][:m]
            xop = layer.stage(xs, out=self.Xp[slot][:m]) if tc else xs
            hm = torch.zeros(m, hld, dtype=hdt, device=layer.device)
            layer.hidden_means(xop, 0, m, hm)
            layer.reconstruction_sq(hm, m, xs, 0, acc)
        self.feed.iterate(None, visit)
        return float(acc.sum().item()) / 4294967296.0 / self.n

    def transform_to_host(self):
        """transform (dbn/models.py:77-86) of the whole host dataset into a new host array [n, H] float32: the next
        layer of a stack trains on it, again from the host."""
        layer = self.layer
        out = np.empty((self.n, layer.H), dtype=np.float32)
        tc = layer.prec in TC_PRECS
        pending = []

        def visit(slot, s0, m):
            xs = self.dev32[slot][:m]
            xop = layer.stage(xs, out=self.Xp[slot][:m]) if tc else xs
            hm = torch.empty(m, layer.H, dtype=torch.float32, device=layer.device)
            layer.hidden_means(xop, 0, m, hm)
            buf, ev = _PinnedPool.take(m * layer.H * 4)
            stage = buf[:m * layer.H * 4].view(torch.float32).view(m, layer.H)
            stage.copy_(hm, non_blocking=True)
            ev.record()
            pending.append((s0, m, stage, buf, ev))
            while len(pending) > 2:
                self._drain_one(pending, out)
        self.feed.iterate(None, visit)
        while pending:
            self._drain_one(pending, out)
        return out

    @staticmethod
    def _drain_one(pending, out):
        s0, m, stage, buf, ev = pending.pop(0)
        ev.synchronize()
        out[s0:s0 + m] = stage.numpy()
        _PinnedPool.give(buf, ev)


class HostMlpIterRunner:
    """The supervised fine-tune (dbn/models.py:434-465) on inputs and targets that stay on the HOST: the same chunk
    pipeline as HostRbmEpochRunner, with the targets gathered alongside the rows.  Philox dropout masks are keyed by
    the global row of the iteration and the L2 decay uses the full dataset size, so the trained parameters are
    bit-identical to the resident MlpIterRunner's."""

    def __init__(self, mlp, X_host, Y_host, batch_size, lr, l2, dropout_p, seed, chunk_rows=None):
        self.mlp = mlp
        self.n, self.bs = int(X_host.shape[0]), int(batch_size)
        dev = mlp.device
        self.first = mlp.layers[0]
        self.R = host_chunk_rows(self.bs, self.n, 4 * int(X_host.shape[1]), chunk_rows)
        self.feed = HostChunkFeeder([X_host, Y_host], self.R, dev)
        self.Xp = [self.first.alloc_staged(self.R, mlp.n_in) for _ in range(2)]
        self.ws = mlp.make_workspace(min(self.bs, self.n))
        self.iter_ctr = torch.zeros(1, dtype=torch.int32, device=dev)
        self.loss_rows = torch.zeros(self.n, dtype=torch.float32, device=dev)
        decay = 1.0 - (lr * l2) / self.n
        self.args = [mlp.step_args(xp, yd, lr / batch_size, decay, 1.0 - dropout_p, self.ws, seed, self.iter_ctr,
                                   self.loss_rows, None) for xp, yd in zip(self.Xp, self.feed.dev[1])]
        self.launches_per_iter = 0

    def run(self, order=None):
        lib = self.mlp.lib
        if order is None:
            order = epoch_order(self.n)
        order = np.asarray(order)
        before = lib.dbn_launch_count()

        def visit(slot, s0, m):
            self.first.stage(self.feed.dev[0][slot][:m], out=self.Xp[slot][:m])
            a = self.args[slot]
            a.row0 = 0
            a.rng.row0 = s0
            a.loss_rows = self.loss_rows.data_ptr() + 4 * s0
            L.check(lib.dbn_mlp_fit_epoch(stream_ptr(), C.byref(a), m, self.bs))
        self.feed.iterate(order, visit)
        self.iter_ctr.add_(1)
        self.launches_per_iter = int(lib.dbn_launch_count() - before)


class DpRbmEpochRunner:
    """Data-parallel epoch of one RBM layer (BASELINE.json config 5).

    Every rank holds the dataset and draws the same shuffle; rank r computes rows
    [r*bl, (r+1)*bl) of each global mini-batch (bl = batch_size / world).  Philox counters are keyed
    by the GLOBAL row index, so the sampled states do not depend on the GPU count.  Per step the
    Gibbs chain runs locally and leaves the raw delta [(H+1) x (V+1)] = dW | dc ; db; then

      * default in BF16 mode (`step_sharded`): reduce-scatter of the delta, every rank updates its H/world
        rows of the fp32 master, all-gather of the refreshed bf16 shadow;
      * otherwise: sum all-reduce of the delta and the identical update on every rank, optionally
        pipelined in `n_chunks` row blocks (block i exchanged and applied on a side stream while block
        i+1 is produced, the next step's first hidden phase started per block).  Measured on 8 x B200
        (profiles/r01_notes.md) the pipeline's extra launches and smaller contractions cost more than
        the overlap returns, so the default is ONE block; DBN_B200_DP_CHUNKS selects more."""

    def __init__(self, layer, Xd, batch_size, k, lr, seed, dp, n_chunks=1):
        import torch.distributed as dist
        self.dist = dist
        self.layer, self.Xd = layer, Xd
        self.rank, self.world = dp
        self.n, self.bs, self.k = int(Xd.shape[0]), int(batch_size), int(k)
        dp_check_shapes(self.n, self.bs, self.world)
        # optional: leave SMs for the NCCL kernels (measured: no gain on 8 x B200, default 0)
        L.check(layer.lib.dbn_set_sm_reserve(int(os.environ.get("DBN_B200_SM_RESERVE", "0"))))
        n_chunks = int(os.environ.get("DBN_B200_DP_CHUNKS", n_chunks))
        self.nocomm = os.environ.get("DBN_B200_DP_NOCOMM", "0") == "1"   # debug: time the step without the exchange
        self.bl = self.bs // self.world
        self.steps = self.n // self.bs
        self.n_local = self.steps * self.bl
        self.scale = lr / batch_size
        H, V = layer.H, layer.V
        self.feeder = _IndexFeeder(self.n_local, layer.device)
        self.Xp = layer.alloc_staged(self.n_local, int(Xd.shape[1]))
        self.ws = layer.make_workspace(self.bl)
        self.epoch_ctr = torch.zeros(1, dtype=torch.int32, device=layer.device)
        self.ldr = (V + 1 + 3) // 4 * 4
        self.raw = torch.zeros(H + 1, self.ldr, dtype=torch.float32, device=layer.device)
        self.args = layer._step_args(self.Xp, k, self.scale, self.ws, seed, self.epoch_ctr)
        self.args.raw_delta = mat(self.raw)
        self.args.B = self.bl
        self.chunks = dp_chunks(H, n_chunks)
        self.comm = torch.cuda.Stream(device=layer.device)
        self.ev_ready = [torch.cuda.Event() for _ in self.chunks]
        self.ev_applied = [torch.cuda.Event() for _ in self.chunks]
        self.pending = False
        self.launches_per_epoch = 0
        # sharded update (reduce-scatter -> row-block update -> all-gather of the bf16 shadow): BF16 mode only,
        # because every contraction then reads the shadow and the fp32 master may stay sharded between steps
        self.sharded = (layer.prec in TC_PRECS and H % self.world == 0 and len(self.chunks) == 1 and
                        os.environ.get("DBN_B200_DP_SHARDED", "1") != "0")
        self.master_sharded = False
        # BF16 mode: dc / db as fixed-point column statistics of the chain's epilogues (the single-GPU step does the
        # same for these widths), written into the raw delta's last column / row by dbn_rbm_delta_rows
        self.stats_flag = L.STEP_BIAS_STATS if layer.prec in TC_PRECS else 0
        if self.sharded:
            self.hs = H // self.world
            self.shard = torch.zeros(self.hs, self.ldr, dtype=torch.float32, device=layer.device)

    def step(self, s):
        """One data-parallel mini-batch.  Row block i of W is final for this step as soon as block i
        of the previous step's delta has been all-reduced and applied, and the t = 0 hidden phase on
        hidden units of block i needs only those rows -- so the tail of the exchange hides behind the
        head of the next step."""
        if self.sharded:
            return self.step_sharded(s)
        layer, lib, a = self.layer, self.layer.lib, self.args
        a.row0 = s * self.bl
        a.rng.row0 = dp_global_row0(s, self.bs, self.rank, self.bl)
        main = torch.cuda.current_stream()
        a.flags = L.STEP_SKIP_UPDATE | L.STEP_SKIP_FIRST_HIDDEN | self.stats_flag
        for i, (h0, h1) in enumerate(self.chunks):
            if self.pending:
                main.wait_event(self.ev_applied[i])
            L.check(lib.dbn_rbm_first_hidden(stream_ptr(), C.byref(a), h0, h1))
        L.check(lib.dbn_rbm_cd_step(stream_ptr(), C.byref(a)))          # rest of the Gibbs chain (needs all of W)
        for i, (h0, h1) in enumerate(self.chunks):
            L.check(lib.dbn_rbm_delta_rows(stream_ptr(), C.byref(a), h0, h1))
            self.ev_ready[i].record(main)
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(self.ev_ready[i])
                blk = self.raw[h0:(h1 + 1 if h1 == layer.H else h1)]
                if not self.nocomm:
                    self.dist.all_reduce(blk, op=self.dist.ReduceOp.SUM)
                L.check(lib.dbn_rbm_apply_delta(stream_ptr(), layer.V, layer.H, h0, h1, 1 if h1 == layer.H else 0,
                                                self.scale, ptr(self.raw), self.ldr, 0, ptr(layer.W), layer.V,
                                                ptr(layer.b), ptr(layer.c), C.byref(mat(layer.Wb))))
                self.ev_applied[i].record(self.comm)
        self.pending = True

    def step_sharded(self, s):
        """One data-parallel mini-batch with a sharded update: reduce-scatter the fp32 delta so that rank r
        receives the summed rows [r*H/w, (r+1)*H/w), update that row block of the fp32 master (and of c),
        then all-gather the refreshed bf16 shadow that every contraction reads.  Against a plain all-reduce
        this moves 3/4 of the bytes (fp32 delta one way, bf16 weights back) and divides the update pass by
        the number of ranks; the fp32 master is gathered once at the end (`finalize`)."""
        layer, lib, a, dist = self.layer, self.layer.lib, self.args, self.dist
        H, V = layer.H, layer.V
        a.row0 = s * self.bl
        a.rng.row0 = dp_global_row0(s, self.bs, self.rank, self.bl)
        a.flags = L.STEP_SKIP_UPDATE | self.stats_flag
        L.check(lib.dbn_rbm_cd_step(stream_ptr(), C.byref(a)))                     # whole Gibbs chain, local rows
        L.check(lib.dbn_rbm_delta_rows(stream_ptr(), C.byref(a), 0, H))           # raw = dW | dc ; db  (partial sums)
        h0, h1 = self.rank * self.hs, (self.rank + 1) * self.hs
        if not self.nocomm:
            dist.reduce_scatter_tensor(self.shard, self.raw[:H], op=dist.ReduceOp.SUM)
            dist.all_reduce(self.raw[H:H + 1], op=dist.ReduceOp.SUM)                 # visible-bias row (tiny)
        else:
            self.shard.copy_(self.raw[h0:h1])
        L.check(lib.dbn_rbm_apply_delta(stream_ptr(), V, H, h0, h1, 0, self.scale, ptr(self.shard), self.ldr, h0,
                                        ptr(layer.W), V, ptr(layer.b), ptr(layer.c), C.byref(mat(layer.Wb))))
        L.check(lib.dbn_rbm_apply_delta(stream_ptr(), V, H, H, H, 1, self.scale, ptr(self.raw), self.ldr, 0,
                                        ptr(layer.W), V, ptr(layer.b), ptr(layer.c), C.byref(mat(layer.Wb))))
        if not self.nocomm:
            dist.all_gather_into_tensor(layer.Wb, layer.Wb[h0:h1])
            dist.all_gather_into_tensor(layer.c, layer.c[h0:h1])
        self.master_sharded = True

    def finalize(self):
        """Bring the fp32 master back to every rank after sharded steps (before export / metrics in fp32)."""
        self.drain()
        if self.sharded and self.master_sharded and not self.nocomm:
            h0, h1 = self.rank * self.hs, (self.rank + 1) * self.hs
            self.dist.all_gather_into_tensor(self.layer.W, self.layer.W[h0:h1])
            self.master_sharded = False

    def drain(self):
        """Make the main stream wait for every outstanding all-reduce + apply."""
        if self.pending:
            main = torch.cuda.current_stream()
            for ev in self.ev_applied:
                main.wait_event(ev)
            self.pending = False

    def run(self, order=None):
        before = self.layer.lib.dbn_launch_count()
        if order is None:
            order = epoch_order(self.n)
        local = dp_local_rows(order, self.steps, self.world, self.bl, self.rank)
        self.feeder.push(local)
        self.layer.stage(self.Xd, idx=self.feeder.dev, out=self.Xp)
        for s in range(self.steps):
            self.step(s)
        self.finalize()
        self.epoch_ctr.add_(1)
        self.launches_per_epoch = int(self.layer.lib.dbn_launch_count() - before)


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
        ws.dbn_rows = int(bmax)
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
        a.ws_rows = getattr(ws, "dbn_rows", 0)
        a.loss_rows = loss_rows.data_ptr() if loss_rows is not None else None
        return a

    def fit(self, Xd, Yd, n_iter, batch_size, lr, l2, dropout_p, seed, rng_mode="philox", verbose=None):
        """NumPyAbstractSupervisedDBN._stochastic_gradient_descent (dbn/models.py:419-469).  The
        1/p pre-scale and p post-scale of _fine_tuning (:533-548) are done by the caller.  Xd / Yd are device tensors
        (resident fine-tune) or host ndarrays (streamed from the host chunk by chunk: HostMlpIterRunner)."""
        n = int(Xd.shape[0])
        if n == 0 or n_iter <= 0:
            return
        if isinstance(Xd, np.ndarray):
            if rng_mode != "philox":
                raise ValueError("rng='numpy' (replay of the reference's dropout stream) needs the dataset resident on the device.")
            runner = HostMlpIterRunner(self, Xd, Yd, batch_size, lr, l2, dropout_p, seed)
        else:
            runner = MlpIterRunner(self, Xd, Yd, batch_size, lr, l2, dropout_p, seed, rng_mode)
        for it in range(1, n_iter + 1):
            runner.run()
            if verbose is not None:
                scale = self.n_out if self.head == L.HEAD_SOFTMAX else 1.0  # dbn/models.py:450,468 quirk
                verbose(it, float(runner.loss_rows.sum().item()) / n * scale)
        torch.cuda.current_stream().synchronize()

    def head_outputs(self, A):
        """_compute_output_units_matrix (dbn/models.py:607-615 / :705-711) on device-resident features [n, H_last]."""
        n = int(A.shape[0])
        out = torch.empty(n, self.n_out, dtype=torch.float32, device=self.device)
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

    def forward(self, Xd):
        """predict path: stacked transform (dbn/models.py:278-287) + output units (:607-615 / :705-711)."""
        A = Xd
        for l in self.layers:
            A = l.transform(A)
        return self.head_outputs(A)


class MlpIterRunner:
    """One fine-tune problem prepared for iteration-at-a-time training: gathered inputs / targets,
    workspace, warm-up launch and the captured CUDA graph of one iteration (dbn/models.py:434-465)."""

    def __init__(self, mlp, Xd, Yd, batch_size, lr, l2, dropout_p, seed, rng_mode="philox"):
        self.mlp, self.Xd, self.Yd = mlp, Xd, Yd
        self.n, self.bs = int(Xd.shape[0]), int(batch_size)
        n, dev, lib = self.n, mlp.device, mlp.lib
        self.keep_p = 1.0 - dropout_p
        self.first = mlp.layers[0]
        self.feeder = _IndexFeeder(n, dev)
        self.Xp = self.first.alloc_staged(n, int(Xd.shape[1]))
        self.Yp = torch.empty_like(Yd)
        self.ws = mlp.make_workspace(min(self.bs, n))
        self.iter_ctr = torch.zeros(1, dtype=torch.int32, device=dev)
        self.loss_rows = torch.zeros(n, dtype=torch.float32, device=dev)
        self.widths = [mlp.n_in] + mlp.dims
        self.masks = None
        if rng_mode == "numpy" and dropout_p > 0:
            self.masks = [torch.ones(n, w, dtype=torch.float32, device=dev) for w in self.widths]
        decay = 1.0 - (lr * l2) / n
        self.args = mlp.step_args(self.Xp, self.Yp, lr / batch_size, decay, self.keep_p, self.ws, seed, self.iter_ctr,
                                  self.loss_rows, self.masks)
        # second stream: the weight updates run next to the delta contractions below them (two branches of the
        # captured graph) instead of after all of them
        self.aux = None
        if os.environ.get("DBN_B200_MLP_FORK", "1") != "0":
            self.aux = torch.cuda.Stream(device=dev)
            self.args.aux_stream = self.aux.cuda_stream
        self.graph = None
        self.launches_per_iter = 0
        self.iters_run = 0
        # tensor-core modes: dbn_mlp_fit_epoch is ONE cooperative launch of the persistent fine-tune kernel
        # (csrc/mlp_tc_epoch.cuh; dropout by Philox or injected masks included) -- nothing to capture
        dims = (C.c_int32 * len(mlp.dims))(*mlp.dims)
        self.tc_epoch = (mlp.prec in TC_PRECS and
                         lib.dbn_mlp_tc_epoch_plan(mlp.prec, mlp.n_in, dims, len(mlp.dims), mlp.n_out, min(self.bs, n)) > 0)

    def _enqueue(self):
        lib = self.mlp.lib
        self.first.stage(self.Xd, idx=self.feeder.dev, out=self.Xp)
        L.check(lib.dbn_gather_rows(stream_ptr(), mat(self.Yd), ptr(self.feeder.dev), self.n, self.mlp.n_out,
                                    mat(self.Yp), -1, None))
        L.check(lib.dbn_mlp_fit_epoch(stream_ptr(), C.byref(self.args), self.n, self.bs))
        self.iter_ctr.add_(1)

    def run(self, order=None):
        lib = self.mlp.lib
        self.feeder.push(epoch_order(self.n) if order is None else order)
        if self.masks is not None:  # reference draw order: per row, V then H_1.. (dbn/models.py:402, :409)
            flat = np.random.binomial(1, self.keep_p, (self.n, sum(self.widths))).astype(np.float32)
            o = 0
            for m, w in zip(self.masks, self.widths):
                m.copy_(torch.from_numpy(np.ascontiguousarray(flat[:, o:o + w])))
                o += w
        if _use_graphs() and self.iters_run >= 1 and not self.tc_epoch:   # iteration 1 eager (warm-up), iteration 2 captured, then replays
            if self.graph is None:
                torch.cuda.current_stream().synchronize()
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._enqueue()
            self.graph.replay()
        else:
            before = lib.dbn_launch_count()
            self._enqueue()
            self.launches_per_iter = int(lib.dbn_launch_count() - before)
        self.iters_run += 1
