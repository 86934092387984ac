"""Data-parallel CD training over NVLink peer memory (BASELINE.json config 5).

The reference sums the deltas of a mini-batch sample by sample and updates once (dbn/models.py:112-119).  Across
GPUs that sum is a reduce-scatter and the refreshed weights an all-gather.  Both are fused into this library's
own kernels instead of being NCCL calls issued after them (see include/dbn_b200.h, "data-parallel exchange"):

  * every rank owns one peer-visible ARENA (cudaMalloc + CUDA IPC handle, mapped by the other ranks) holding, at the
    same offsets on every rank: the receive slots of the dW rows it owns, the statistics slots, the flag words, its
    bf16 weight shadow, and its slice of the staged dataset;
  * `dbn_rbm_delta_scatter` stores the dW tiles straight into the owners' receive slots from the contraction's epilogue;
  * `dbn_dp_apply` sums the slots in rank order, updates the owner's fp32 rows and writes their bf16 rounding into
    EVERY rank's shadow; b and c are updated on every rank from the exact integer sum of the per-rank statistics;
  * the epoch shuffle reads rows from the rank that staged them (`dbn_gather_rows_peer`), so a rank uploads only
    1/world of the dataset.

torch.distributed is used for plumbing only: exchanging the 64-byte handles, broadcasting the initial parameters and
the epoch orders from rank 0, and the final all-gather of the sharded fp32 master.

`VirtualPeerGroup` runs the same kernels for `world` virtual ranks inside ONE process on ONE GPU (arenas are plain
allocations, the "peer" pointers local): the single-GPU test of the protocol, as the profiling guide asks for.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib as L
from . import engine as E


class _DevView:
    """Minimal __cuda_array_interface__ carrier: lets torch wrap memory this library allocated."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2}


def device_view(ptr, shape, dtype, device):
    """torch tensor over raw device memory [ptr, ptr + prod(shape) * itemsize) (no copy, no ownership)."""
    if dtype == torch.bfloat16:
        return torch.as_tensor(_DevView(ptr, shape, "<i2"), device=device).view(torch.bfloat16)
    typestr = {torch.float32: "<f4", torch.float16: "<f2", torch.int64: "<i8", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
    return torch.as_tensor(_DevView(ptr, shape, typestr), device=device)


def _align(x, a=1024):
    return (x + a - 1) // a * a


class ArenaLayout:
    """Offsets inside a rank's arena; identical on every rank.  `op_dtype`: the 16-bit operand type of the layer's
    tensor-core mode (weight shadow and staged rows)."""

    def __init__(self, world, V, H, rows_local, op_dtype=torch.bfloat16):
        assert H % world == 0
        self.world, self.V, self.H = world, V, H
        self.op_dtype = op_dtype
        self.hs = H // world
        self.ldr = (V + 3) // 4 * 4
        self.pitch = E.bf16_pitch(V)
        self.rows_local = rows_local
        off = 0
        self.off_recv = off
        off = _align(off + world * self.hs * self.ldr * 4)
        self.off_stats = off
        off = _align(off + world * (H + V) * 8)
        self.off_flags = off
        off = _align(off + 2 * L.MAX_PEERS * 8)
        self.off_wb = off
        off = _align(off + H * self.pitch * 2)
        self.off_xs = off
        off = _align(off + max(rows_local, 1) * self.pitch * 2)
        self.nbytes = off


class PeerMember:
    """One rank's view of the group: its own arena and the addresses of everybody's, as mapped on this device."""

    def __init__(self, rank, world, bases, layout, device, group=None):
        self.rank, self.world, self.bases, self.layout, self.device, self.group = rank, world, list(bases), layout, device, group

    @property
    def base(self):
        return self.bases[self.rank]

    def exchange(self):
        lay = self.layout
        x = L.DpExchange()
        x.n_peers, x.rank, x.V, x.H = self.world, self.rank, lay.V, lay.H
        x.rows_per_peer, x.ld_recv = lay.hs, lay.ldr
        for j, b in enumerate(self.bases):
            x.recv[j] = b + lay.off_recv
            x.stats[j] = b + lay.off_stats
            x.flags[j] = b + lay.off_flags
            x.Wb[j] = L.Mat(b + lay.off_wb, lay.pitch, L.F16 if lay.op_dtype == torch.float16 else L.BF16)
        return x

    def shadow(self):
        lay = self.layout
        return device_view(self.base + lay.off_wb, (lay.H, lay.pitch), lay.op_dtype, self.device)

    def staged_slice(self):
        lay = self.layout
        return device_view(self.base + lay.off_xs, (max(lay.rows_local, 1), lay.pitch), lay.op_dtype, self.device)

    def staged_sources(self):
        arr = (C.c_void_p * L.MAX_PEERS)()
        for j, b in enumerate(self.bases):
            arr[j] = b + self.layout.off_xs
        return arr

    def barrier(self):
        torch.cuda.current_stream().synchronize()
        if self.group is not None:
            self.group.dist.barrier()


class RealPeerGroup:
    """One process per GPU: allocate the local arena, swap IPC handles over torch.distributed, map the peers."""

    def __init__(self, rank, world, layout, device):
        import torch.distributed as dist
        self.dist = dist
        self.lib = E.require_device()
        self.rank, self.world = rank, world
        ptr = C.c_void_p()
        handle = (C.c_ubyte * L.PEER_HANDLE_BYTES)()
        self.own, self.opened = None, []
        if self.lib.dbn_peer_alloc(layout.nbytes, C.byref(ptr), handle) == L.OK:
            self.own = ptr.value
        mine = bytes(handle) if self.own is not None else bytes(L.PEER_HANDLE_BYTES)
        handles = [None] * world
        dist.all_gather_object(handles, mine)       # every rank takes part, also one whose allocation failed
        if any(h == bytes(L.PEER_HANDLE_BYTES) for h in handles):
            if self.own is not None:
                self.lib.dbn_peer_free(C.c_void_p(self.own))
                self.own = None
            raise L.DbnB200Error("a rank could not allocate or export its peer arena (CUDA IPC)")
        bases = []
        try:
            for j in range(world):
                if j == rank:
                    bases.append(self.own)
                    continue
                p = C.c_void_p()
                buf = (C.c_ubyte * L.PEER_HANDLE_BYTES).from_buffer_copy(handles[j])
                L.check(self.lib.dbn_peer_open(buf, C.byref(p)))
                self.opened.append(p.value)
                bases.append(p.value)
        except L.DbnB200Error:
            self._release()
            raise
        self.member = PeerMember(rank, world, bases, layout, device, group=self)

    def members(self):
        return [self.member]

    def _release(self):
        for p in self.opened:
            self.lib.dbn_peer_close(C.c_void_p(p))
        if self.own is not None:
            self.lib.dbn_peer_free(C.c_void_p(self.own))
        self.own, self.opened = None, []

    def close(self):
        """Every rank is done with everybody's memory before anything is unmapped or freed."""
        if self.own is None:
            return
        torch.cuda.synchronize()
        self.dist.barrier()
        self._release()


# A fit needs the same arena every time (same layer shape, same slice of the dataset), and setting one up costs
# far more than training on it for an epoch: cudaMalloc + IPC export / import ~15 ms, unmapping and freeing it
# ~110 ms (measured on 2 x B200, tools/e2e_dp_breakdown.py).  Groups are therefore kept per process and reused by
# the next fit with the same layout; every rank takes the same decision because the key is the layout itself.
_GROUPS = {}


def acquire_group(rank, world, layout, device):
    """A RealPeerGroup for this layout: the cached one (flags cleared, all ranks synchronised) or a new one."""
    key = (rank, world, layout.V, layout.H, layout.rows_local, str(layout.op_dtype), device.index)
    g = _GROUPS.get(key)
    if g is not None and g.own is not None:
        flags = device_view(g.member.base + layout.off_flags, (2 * L.MAX_PEERS,), torch.int64, device)
        flags.zero_()
        g.member.barrier()
        return g
    for k in [k for k in _GROUPS if k[0] == rank and k[1] == world and k != key]:
        _GROUPS.pop(k).close()          # one cached arena per process: a different shape replaces it
    g = RealPeerGroup(rank, world, layout, device)
    _GROUPS[key] = g
    return g


def release_groups():
    """Unmap and free every cached arena (collective: every rank calls it)."""
    for k in list(_GROUPS):
        _GROUPS.pop(k).close()


class VirtualPeerGroup:
    """`world` virtual ranks in ONE process on ONE GPU: the arenas are ordinary allocations and the peer pointers are
    local.  The caller drives the ranks phase by phase in lockstep (see run_epoch), so every flag a kernel waits for
    has been raised by an earlier launch on the same stream."""

    def __init__(self, world, layout, device):
        self.lib = E.require_device()
        self.world = world
        self.ptrs = []
        for _ in range(world):
            ptr = C.c_void_p()
            handle = (C.c_ubyte * L.PEER_HANDLE_BYTES)()
            L.check(self.lib.dbn_peer_alloc(layout.nbytes, C.byref(ptr), handle))
            self.ptrs.append(ptr.value)
        self._members = [PeerMember(r, world, self.ptrs, layout, device) for r in range(world)]

    def members(self):
        return self._members

    def close(self):
        torch.cuda.synchronize()
        for p in self.ptrs:
            self.lib.dbn_peer_free(C.c_void_p(p))
        self.ptrs = []


class DpPeerRunner:
    """One rank of the peer-memory data-parallel epoch (BF16 mode, H % world == 0).

    `X_local`: this rank's slice of the dataset, fp32 [n / world, V] on the device (rows [rank * n/world, ...) of the
    global dataset).  It is staged once (bf16 + ones column) into the arena; every epoch the rank gathers the rows it
    computes -- slice [rank * bl, (rank+1) * bl) of every global mini-batch of the shuffled epoch -- from whichever
    rank staged them."""

    def __init__(self, layer, X_local, n_total, batch_size, k, lr, seed, member):
        self.layer, self.member = layer, member
        self.lib = layer.lib
        self.rank, self.world = member.rank, member.world
        self.n, self.bs, self.k = int(n_total), int(batch_size), int(k)
        E.dp_check_shapes(self.n, self.bs, self.world)
        lay = member.layout
        assert layer.prec in E.TC_PRECS and layer.H % self.world == 0 and lay.op_dtype == E.OP_DTYPE[layer.prec]
        assert lay.rows_local * self.world == self.n and int(X_local.shape[0]) == lay.rows_local
        self.bl = self.bs // self.world
        self.steps = self.n // self.bs
        self.n_local = self.steps * self.bl
        self.scale = lr / batch_size
        dev = layer.device
        layer.move_shadow(member.shadow())
        self.Xs = member.staged_slice()
        self.Xs.zero_()
        layer.stage(X_local, out=self.Xs)                      # bf16 + ones column, once per fit
        self.srcs = member.staged_sources()
        self.feeder = E._IndexFeeder(self.n_local, dev)
        self.Xp = layer.alloc_staged(self.n_local, layer.V)
        self.ws = layer.make_workspace(self.bl)
        self.epoch_ctr = torch.zeros(1, dtype=torch.int32, device=dev)
        self.args = layer._step_args(self.Xp, k, self.scale, self.ws, seed, self.epoch_ctr)
        self.args.B = self.bl
        self.args.flags = L.STEP_SKIP_UPDATE | L.STEP_BIAS_STATS
        self.x = member.exchange()
        self.stats = C.c_void_p(self.lib.dbn_rbm_workspace_stats(layer.prec, self.bl, layer.V, layer.H, E.ptr(self.ws)))
        self.counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self.value = 0                 # flag value of the last step issued
        self.master_sharded = False
        self.launches_per_epoch = 0
        member.barrier()               # every rank's slice is staged (and its flags are zero) before anyone reads it

    # -- one mini-batch, in the three phases a lockstep driver needs ------------------------------------------
    def step_compute(self, s):
        """Gibbs chain on this rank's rows, dW scattered to the owners, statistics + 'deltas out' flags."""
        lib, a = self.lib, self.args
        if self.value > 0:             # the previous step's weight rows have arrived from every owner
            L.check(lib.dbn_dp_wait(E.stream_ptr(), C.byref(self.x), 1, self.value))
        self.value += 1
        a.row0 = s * self.bl
        a.rng.row0 = E.dp_global_row0(s, self.bs, self.rank, self.bl)
        L.check(lib.dbn_rbm_cd_step(E.stream_ptr(), C.byref(a)))
        L.check(lib.dbn_rbm_delta_scatter(E.stream_ptr(), C.byref(a), C.byref(self.x)))
        L.check(lib.dbn_dp_signal(E.stream_ptr(), C.byref(self.x), self.stats, self.value))

    def step_apply(self):
        """Owner side: ordered sum of the slots, sharded update, bf16 rows into every shadow, 'weights in' flags."""
        layer = self.layer
        L.check(self.lib.dbn_dp_apply(E.stream_ptr(), C.byref(self.x), self.scale, E.ptr(layer.W), layer.V,
                                      E.ptr(layer.b), E.ptr(layer.c), self.stats, E.ptr(self.counter), self.value))
        self.master_sharded = True

    def step(self, s):
        self.step_compute(s)
        self.step_apply()

    def step_traced(self, s, ev):
        """One step with a CUDA event after every phase (debug: tools/dp_trace.py); ev = list of 7 timing events."""
        lib, a = self.lib, self.args
        st = torch.cuda.current_stream()
        ev[0].record(st)
        if self.value > 0:
            L.check(lib.dbn_dp_wait(E.stream_ptr(), C.byref(self.x), 1, self.value))
        ev[1].record(st)
        self.value += 1
        a.row0 = s * self.bl
        a.rng.row0 = E.dp_global_row0(s, self.bs, self.rank, self.bl)
        L.check(lib.dbn_rbm_cd_step(E.stream_ptr(), C.byref(a)))
        ev[2].record(st)
        L.check(lib.dbn_rbm_delta_scatter(E.stream_ptr(), C.byref(a), C.byref(self.x)))
        ev[3].record(st)
        L.check(lib.dbn_dp_signal(E.stream_ptr(), C.byref(self.x), self.stats, self.value))
        ev[4].record(st)
        self.step_apply()
        ev[5].record(st)

    def drain(self):
        if self.value > 0:
            L.check(self.lib.dbn_dp_wait(E.stream_ptr(), C.byref(self.x), 1, self.value))

    def gather(self, order):
        """This rank's rows of the shuffled epoch, read from the ranks that staged them."""
        local = E.dp_local_rows(order, self.steps, self.world, self.bl, self.rank)
        idx = self.feeder.push(local)
        L.check(self.lib.dbn_gather_rows_peer(E.stream_ptr(), self.srcs, self.world, self.member.layout.rows_local,
                                              self.member.layout.pitch, E.ptr(idx), self.n_local, C.byref(E.mat(self.Xp))))

    def finalize(self):
        """All steps done everywhere; bring the sharded fp32 master back to every rank."""
        self.drain()
        if self.master_sharded and self.member.group is not None:
            hs = self.member.layout.hs
            h0 = self.rank * hs
            self.member.group.dist.all_gather_into_tensor(self.layer.W, self.layer.W[h0:h0 + hs])
        self.master_sharded = False

    def run(self, order):
        before = self.lib.dbn_launch_count()
        self.gather(order)
        for s in range(self.steps):
            self.step(s)
        self.finalize()
        self.epoch_ctr.add_(1)
        self.launches_per_epoch = int(self.lib.dbn_launch_count() - before)


def run_epoch_lockstep(runners, order):
    """Drive the virtual ranks of a VirtualPeerGroup through one epoch on one stream: all ranks' compute phases, then
    all ranks' apply phases, per step -- the order in which the flags are raised before they are waited for."""
    for r in runners:
        r.gather(order)
    for s in range(runners[0].steps):
        for r in runners:
            r.step_compute(s)
        for r in runners:
            r.step_apply()
    for r in runners:
        r.drain()
    # the fp32 master is sharded by rows: assemble it on every virtual rank
    hs = runners[0].member.layout.hs
    for r in runners:
        for o in runners:
            if o is not r:
                r.layer.W[o.rank * hs:(o.rank + 1) * hs].copy_(o.layer.W[o.rank * hs:(o.rank + 1) * hs])
    for r in runners:
        r.master_sharded = False
        r.epoch_ctr.add_(1)


def peer_path_wanted(layer, world):
    return peer_path_wanted_for(layer.prec, layer.H, world)


def peer_path_wanted_for(prec, H, world):
    return (E.PREC_IDS.get(prec, prec) in E.TC_PRECS and H % world == 0 and world <= L.MAX_PEERS and
            os.environ.get("DBN_B200_DP_P2P", "1") != "0")
