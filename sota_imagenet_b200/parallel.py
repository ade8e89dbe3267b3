"""Single-node data parallelism: one process per GPU, NCCL over NVLink through
torch.distributed.  Replaces `DistributedDataParallel(model, device_ids=[local_rank])`
(reference train.py:113-114) and adds SyncBN (north_star; oracle = one process on the global
batch, torch/nn/modules/_functions.py:39-106).

  * parameters + buffers are broadcast from rank 0 once (DDP constructor semantics);
  * BatchNorm statistics are all-reduced as one [2, C] fp32 tensor per layer and pass;
  * gradients live in one flat arena; as the backward pass walks the residual blocks in reverse,
    finished slices of that arena are averaged with asynchronous all-reduces (reverse-order
    buckets, like DDP's 25 MB buckets) that overlap the remaining backward kernels.
"""
import torch
import torch.distributed as dist
import torch.nn as nn

from .modules import BatchNorm2d, SibModule


def plan_buckets(block_starts, total, bucket_elems):
    """Reverse-order gradient buckets.  `block_starts[i]` = arena offset of residual block i
    (ascending); backward finishes blocks from the last to the first.  Returns {block_index:
    (lo, hi)} for the blocks whose completion closes a bucket of >= bucket_elems elements, plus
    the key -1 for the final bucket (stem and whatever is left) issued when backward ends."""
    plan, hi = {}, total
    for i in range(len(block_starts) - 1, -1, -1):
        lo = block_starts[i]
        if hi - lo >= bucket_elems:
            plan[i] = (lo, hi)
            hi = lo
    plan[-1] = (0, hi)
    return plan


def allreduce_mean_(flat, lo, hi, group=None, async_op=False):
    """In-place average of flat[lo:hi] over the ranks (SUM then scale: works on gloo and nccl)."""
    if hi <= lo:
        return None
    buf = flat[lo:hi]
    world = dist.get_world_size(group)
    if dist.get_backend(group) == "nccl":
        return dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
    work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group, async_op=False)
    buf.div_(world)
    return work if async_op else None


class PeerAllReduce:
    """One-shot all-reduce of small fp32 vectors over NVLink peer memory (csrc/peer.cu): the 106
    sequentially dependent SyncBN reductions of a ResNet-50 step no longer pay a collective-library
    latency each.  Single node, one process per GPU, <= 8 ranks; every rank must issue the same
    sequence of calls per step (it does: same model).  Call slots restart at every forward pass."""
    SLOTS = 512
    CAPACITY = 1 << 21            # 8-byte words per parity (16 MB): sum over calls of world * n

    def __init__(self, group=None):
        import ctypes
        from . import _lib, ops
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        assert 1 <= self.world <= 8
        self.device = torch.device("cuda", torch.cuda.current_device())
        lib = _lib.load()
        self._lib, self._check = lib, _lib.check
        flag_bytes = 2 * self.SLOTS * self.world * 4
        mine, err = [], None
        try:
            for nbytes in (2 * self.CAPACITY * 8, flag_bytes):
                ptr, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
                _lib.check(lib.sib_ipc_alloc(nbytes, ctypes.byref(ptr), handle))
                mine.append((ptr.value, bytes(handle)))
            torch.cuda.synchronize()
        except Exception as e:
            err = e
        gathered = [None] * self.world
        dist.all_gather_object(gathered, None if err else [h for _, h in mine], group=group)
        self._local = [p for p, _ in mine]
        self._opened = []
        table = []
        try:
            if err or any(g is None for g in gathered):
                raise RuntimeError("IPC allocation failed on some rank: %s" % err)
            for kind in range(2):
                for r in range(self.world):
                    if r == self.rank:
                        table.append(mine[kind][0])
                    else:
                        ptr = ctypes.c_void_p()
                        buf = (ctypes.c_ubyte * 64).from_buffer_copy(gathered[r][kind])
                        _lib.check(lib.sib_ipc_open(buf, ctypes.byref(ptr)))
                        self._opened.append(ptr.value)
                        table.append(ptr.value)
                table += [0] * (8 - self.world)
        except Exception as e:
            err = e
        ok = torch.tensor([0.0 if err else 1.0], device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)     # also: all mailboxes mapped
        if ok.item() != 1.0:
            raise RuntimeError("peer mailboxes could not be mapped on every rank (%s)" % err)
        self.table = torch.tensor(table, dtype=torch.int64).to(self.device)     # PeerTable
        self.epochs = torch.zeros(self.SLOTS, dtype=torch.int32, device=self.device)
        self.slot = 0
        self.offset = 0
        self._layout = {}          # slot -> (offset, n): fixed after the first step

    def begin_forward(self):
        self.slot, self.offset = 0, 0

    def reset_layout(self):
        """Forget the per-slot sizes (a different model is about to use the mailboxes).  Collective:
        call on every rank, followed by a barrier, before the next reduction."""
        self._layout = {}
        self.slot, self.offset = 0, 0

    def allreduce_(self, t):
        """In-place sum of a contiguous fp32 tensor (numel % 4 == 0) over the ranks."""
        from . import ops
        n = t.numel()
        s = self.slot
        lay = self._layout.get(s)
        if lay is None:
            lay = self._layout[s] = (self.offset, n)
        off, n0 = lay
        if n0 != n or s >= self.SLOTS or off + self.world * n > self.CAPACITY or not t.is_contiguous():
            raise RuntimeError("PeerAllReduce: call sequence changed (slot %d: %d vs %d floats)" % (s, n0, n))
        self.slot += 1
        self.offset = off + self.world * n
        ops.call("sib_peer_allreduce", ops._p(t), n, ops._p(self.table), off, self.CAPACITY, s, self.SLOTS,
                 ops._p(self.epochs), self.rank, self.world, ops._stream())
        return t

    def close(self):
        for p in self._opened:
            self._lib.sib_ipc_close(ctypes_void(p))
        for p in self._local:
            self._lib.sib_ipc_free(ctypes_void(p))
        self._opened, self._local = [], []


def ctypes_void(v):
    import ctypes
    return ctypes.c_void_p(v)


def convert_sync_batchnorm(module, process_group=None):
    for m in module.modules():
        if isinstance(m, BatchNorm2d):
            m.sync = True
            m.process_group = process_group
    return module


class DataParallel(nn.Module):
    def __init__(self, module, sync_bn=True, bucket_mb=25.0, process_group=None,
                 broadcast_buffers=True):
        super().__init__()
        if not isinstance(module, SibModule):
            raise TypeError("DataParallel wraps sota_imagenet_b200 models")
        self.module = module
        self.process_group = process_group
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.bucket_elems = int(bucket_mb * 1024 * 1024 / 4)
        self._handles = []
        self._plan = None
        if sync_bn:
            convert_sync_batchnorm(module, process_group)
            # SyncBN statistics over NVLink peer memory instead of one NCCL all-reduce per layer
            import os
            from . import ops
            if (self.world_size > 1 and self.world_size <= 8 and dist.get_backend(process_group) == "nccl"
                    and os.environ.get("SIB_PEER_ALLREDUCE", "1") != "0"):
                if ops.PEER is None:
                    # every rank must take the same path: agree on success before using it
                    try:      # (the constructor agrees on success across ranks before raising)
                        ops.PEER = PeerAllReduce(process_group)
                    except RuntimeError as e:    # no peer access / IPC unavailable: NCCL path
                        print("sota_imagenet_b200: peer-memory all-reduce unavailable (%s); using NCCL" % e)
                else:
                    torch.cuda.synchronize()
                    ops.PEER.reset_layout()
                    dist.barrier(group=process_group)
        arena = module.ensure_arena()
        if self.world_size > 1:
            dist.broadcast(arena.flat, 0, group=process_group)
            if broadcast_buffers:
                for b in module.buffers():
                    dist.broadcast(b, 0, group=process_group)
            arena.refresh_shadow(force=True)
        # arena offsets at which each residual block starts (arena order == registration order)
        self._block_starts = []
        if hasattr(module, "blocks"):
            for blk in module.blocks():
                first = next(blk.parameters())
                self._block_starts.append(arena.offset_of[id(first)])
        module._block_bwd_cb = self._on_block_done
        module._bwd_hooks = [h for h in module._bwd_hooks if getattr(h, "__self__", None) is not self]
        module._bwd_hooks.append(self._on_backward_done)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    # ---- bucketed, overlapped gradient averaging --------------------------------------
    def _reduce(self, lo, hi):
        if self.world_size == 1 or hi <= lo:
            return
        h = allreduce_mean_(self.module._arena.grad, lo, hi, self.process_group, async_op=True)
        if h is not None:
            self._handles.append(h)

    def _on_block_done(self, block_index):
        if self._plan is None:
            self._plan = plan_buckets(self._block_starts, self.module._arena.total, self.bucket_elems)
        rng = self._plan.get(block_index)
        if rng is not None:
            self._reduce(*rng)

    def _on_backward_done(self, module):
        if self._plan is None:
            self._plan = plan_buckets(self._block_starts, module._arena.total, self.bucket_elems)
        self._reduce(*self._plan[-1])
        for h in self._handles:
            h.wait()
        self._handles = []

    # nn.Module plumbing so optimizers / checkpoints see the wrapped model's names
    def state_dict(self, *args, **kwargs):
        return self.module.state_dict(*args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        return self.module.load_state_dict(*args, **kwargs)
