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
