"""Drop-in nn.Modules whose forward/backward are explicit sequences of libsib200 kernel launches.

Every module implements
    fwd(x, train) -> (y, saved)      plain tensors in / out, no autograd graph
    bwd(dy, saved) -> dx             writes parameter gradients into the arena as a side effect
and `forward()` wraps that pair into ONE autograd node, so a whole ResNet is a single node in
the autograd graph (no per-op dispatch, no AccumulateGrad kernels).  Activations are bf16
channels_last ([N,C,H,W] logical, NHWC in memory).

Module / parameter names follow torchvision's ResNet so checkpoints interchange
(reference train.py:98-101 loads with strict=False).
"""
import math
import os

import torch
import torch.nn as nn

from . import _lib, ops
from .arena import ParamArena


# Fuse the BatchNorm-backward reduction (and the activation mask) into the epilogue of the dgrad
# that produces the gradient (csrc/conv.cu igemm_epilogue); SIB_FUSE_BN_BWD=0 restores the separate
# bn_bwd_reduce passes (kept for A/B measurements and as the parity cross-check).
FUSE_BN_BWD = os.environ.get("SIB_FUSE_BN_BWD", "1") != "0" and not ops.DETERMINISTIC
# 3x3 / stride-2 dgrad by row parity (csrc/conv.cu dgrad_s2_impl); 0 = zero-inserted dy
FUSE_BN_BWD_ALL = os.environ.get("SIB_FUSE_BN_BWD_ALL", "0") == "1"   # A/B: fuse even where it loses
DGRAD_S2 = os.environ.get("SIB_DGRAD_S2", "1") != "0"
# BatchNorm + activation applied inside the CONSUMER conv's operand prologue (csrc/conv.cu
# BnPrologue): the normalised activation is not stored in forward; backward re-materialises it
# inside bn_bwd_apply for the weight gradient.  0 = off (separate bn_finalize_apply pass),
# 1 = where it wins (measured per layer on B200, profiles/r02_prologue_fusion_layers.log):
# conv3 <- bn2 (1x1) up to 256 input channels and the stride-2 conv2 <- bn1 (the separate pass
# would run over the 4x larger input) up to 256; 2 = 1 + every conv3, 3 = everything that fits
# (3x3 stride-1 loses: its transform runs once per filter tap on the im2col kernel).
FUSE_BN_FWD = 0 if ops.DETERMINISTIC else int(os.environ.get("SIB_FUSE_BN_FWD", "1"))


def _fuse_fwd(kind, planes, stride):
    if FUSE_BN_FWD <= 0 or not ops.fprop_bnact_ok(planes):
        return False
    if FUSE_BN_FWD >= 3:
        return True
    if kind == "conv3":
        return planes <= 256 or FUSE_BN_FWD >= 2
    return stride == 2 and planes <= 256


# --------------------------------------------------------------------------- autograd glue
class _SibFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dummy, module):
        y, saved = module.fwd(x, True)
        ctx.module = module
        ctx.saved = saved
        ctx.x_needs_grad = ctx.needs_input_grad[0]
        return y

    @staticmethod
    def backward(ctx, dy):
        module = ctx.module
        module._begin_backward()
        dx = module.bwd(dy, ctx.saved, need_dx=ctx.x_needs_grad)
        ctx.saved = None
        module._end_backward()
        return dx, None, None


class SibModule(nn.Module):
    """Base class: owns the arena when used as the root of a forward call."""

    def __init__(self):
        super().__init__()
        self._arena = None
        self._bwd_hooks = []   # callables(module) run when backward of this root finishes

    # -- arena management -----------------------------------------------------------
    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._arena = None
        return out

    def ensure_arena(self):
        """(Re)build the flat arena for all parameters below this module if needed."""
        a = self._arena
        if a is not None and a.intact():
            return a
        params = [(n, p) for n, p in self.named_parameters()]
        if not params:
            return None
        dev = params[0][1].device
        if dev.type != "cuda":
            raise _lib.SibError("sota_imagenet_b200 modules run on CUDA sm_100a only; call .cuda()")
        for _, p in params:
            if p.dtype != torch.float32:
                raise _lib.SibError("parameters must stay float32 (bf16 shadows are kept internally)")
        existing = getattr(params[0][1], "_sib_arena", None)
        # reuse only an arena laid out in REGISTRATION order: the data-parallel bucket plan
        # (parallel.plan_buckets) assumes arena order == registration order; an arena built by an
        # optimizer in param-group order (step()/load_state_dict() before the first forward) is
        # rebuilt here and the optimizer migrates its state (optimizers._collect_arenas)
        if existing is not None and existing.intact() and len(existing.entries) == len(params) and all(
                e[1] is p for e, (_, p) in zip(existing.entries, params)):
            self._arena = existing
        else:
            self._arena = ParamArena(params, dev)
        for m in self.modules():
            if isinstance(m, SibModule) and m is not self:
                m._arena = self._arena
        # one counter tensor for every BatchNorm below this root: `num_batches_tracked += 1` was 53
        # tiny launches per step; the buffers become views and the root bumps them all at once
        bns = [m for m in self.modules() if isinstance(m, BatchNorm2d)]
        self._nbt = None
        if bns:
            self._nbt = torch.stack([m.num_batches_tracked.to(dev) for m in bns]).contiguous()
            for i, m in enumerate(bns):
                m._buffers["num_batches_tracked"] = self._nbt[i]
                m._nbt_batched = True
        return self._arena

    def _begin_backward(self):
        a = self.ensure_arena()
        if a is not None:
            a.prepare_grads()
            ops.begin_pass(a.device)

    def _end_backward(self):
        ops.side_join()
        for h in self._bwd_hooks:
            h(self)

    # -- nn.Module API ----------------------------------------------------------------
    def forward(self, x):
        _lib.require_device()
        a = self.ensure_arena()
        if a is not None:
            a.refresh_shadow()
        ops.begin_pass(x.device, forward=True)
        x = self._prepare_input(x)
        if self.training and getattr(self, "_nbt", None) is not None:
            self._nbt += 1
        if torch.is_grad_enabled() and (self.training or x.requires_grad):
            # A fresh leaf per call (no kernel: torch.empty): it only makes the node's output require
            # grad.  A cached leaf keeps ONE AccumulateGrad node, bound to the stream of its first
            # use, alive for as long as any earlier loss is still referenced; a later CUDA-graph
            # capture then fails with "dependency created on uncaptured work in another stream"
            # (scripts/probes/graph_capture_probe.py, variant 64).
            dummy = torch.empty((), device=x.device, requires_grad=True)
            return _SibFn.apply(x, dummy, self)
        y, _ = self.fwd(x, self.training)
        return y

    def _prepare_input(self, x):
        if x.dim() == 4 and (x.dtype != torch.bfloat16 or not x.permute(0, 2, 3, 1).is_contiguous()):
            if x.requires_grad:
                raise _lib.SibError("inputs that require grad must already be bf16 channels_last")
            return ops.to_nhwc_bf16(x)
        return x

    # helpers for parameters living in the arena
    def _w16(self, p):
        return self._arena.shadow_view(p)

    def _wd16(self, p):
        return self._arena.dgrad_view(p)

    def _grad(self, p):
        return self._arena.grad_view(p)


# --------------------------------------------------------------------------- layers
class Conv2d(SibModule):
    """bias-free NHWC bf16 convolution on tcgen05 (csrc/conv.cu)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=False,
                 needs_dgrad=True):
        super().__init__()
        assert not bias, "conv bias is not used by the ResNet family (BN follows every conv)"
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = (kernel_size, kernel_size)
        self.stride, self.padding = stride, padding
        w = torch.empty(out_channels, in_channels, kernel_size, kernel_size)
        nn.init.kaiming_normal_(w, mode="fan_out", nonlinearity="relu")
        self.weight = nn.Parameter(w)
        self.weight._sib_layout = "krsc"
        self.weight._sib_needs_dgrad = needs_dgrad
        self.weight._sib_stride = stride

    def extra_repr(self):
        return "%d, %d, kernel_size=%s, stride=%d, padding=%d" % (
            self.in_channels, self.out_channels, self.kernel_size, self.stride, self.padding)

    def run(self, x, stats=None):
        return ops.conv2d_fprop(x, self._w16(self.weight), self.stride, self.padding, stats=stats)

    def run_dgrad(self, dy, x_shape, out=None, residual=None, bn_bwd=None):
        k = self.kernel_size[0]
        w_s2 = self._arena.dgrad_s2_view(self.weight) if self.stride == 2 and k == 3 and DGRAD_S2 else None
        return ops.conv2d_dgrad(dy, self._wd16(self.weight), x_shape, k, k, self.stride,
                                self.padding, out=out, residual=residual, bn_bwd=bn_bwd, w_s2=w_s2)

    def run_wgrad(self, x, dy):
        dw = self._grad(self.weight)
        ops.side_launch(lambda: ops.conv2d_wgrad(x, dy, dw, self.stride, self.padding), x, dy)

    def fwd(self, x, train):
        return self.run(x), (x,)

    def bwd(self, dy, saved, need_dx=True):
        (x,) = saved
        dy = _as_act(dy)
        self.run_wgrad(x, dy)
        return self.run_dgrad(dy, tuple(x.shape)) if need_dx else None


class StemConv(SibModule):
    """KxK stride-2 convolution over the 3-channel image, run as a (K+1)/2 x 1 stride-1
    tensor-core convolution over a row-pair-packed 64-channel copy (csrc/augment.cu)."""

    def __init__(self, out_channels=64, kernel_size=7, padding=3):
        super().__init__()
        assert kernel_size % 2 == 1 and padding == kernel_size // 2
        self.in_channels, self.out_channels = 3, out_channels
        self.kernel_size = (kernel_size, kernel_size)
        self.stride, self.padding = 2, padding
        # rows 2p + r - pad = 2(p + a - a0) + b  ->  r = 2a + b - off
        self.off = padding % 2 if padding % 2 == 1 else 0
        self.a0 = (padding + self.off) // 2
        self.na = (kernel_size - 1 + self.off) // 2 + 1
        w = torch.empty(out_channels, 3, kernel_size, kernel_size)
        nn.init.kaiming_normal_(w, mode="fan_out", nonlinearity="relu")
        self.weight = nn.Parameter(w)   # plain OIHW layout

    def _packed_weight(self):
        wq = ops.stem_pack_weight(self.weight.data, self.na, self.off)
        return wq.permute(0, 3, 1, 2)   # logical [K, 64, NA, 1], KRSC memory

    def _prepare_input(self, x):
        return x

    def run(self, x, stats=None):
        n, c, h, w = x.shape
        xq = ops.stem_pack(x, self.kernel_size[1], self.padding)
        y = ops.conv2d_fprop(xq, self._packed_weight(), 1, 0, stats=stats, pad_hw=(self.a0, 0),
                             out_hw=(h // 2, w // 2))
        return y, xq

    def run_wgrad(self, xq, dy):
        k = self.out_channels
        dwq = torch.zeros((k, self.na, 1, 64), dtype=torch.float32, device=dy.device).permute(0, 3, 1, 2)
        ops.conv2d_wgrad(xq, dy, dwq, 1, 0, pad_hw=(self.a0, 0))
        ops.stem_unpack_wgrad(dwq, self._grad(self.weight), self.na, self.off, accumulate=True)

    def fwd(self, x, train):
        y, xq = self.run(x)
        return y, (xq,)

    def bwd(self, dy, saved, need_dx=False):
        assert not need_dx, "the stem has no data gradient (image input)"
        self.run_wgrad(saved[0], _as_act(dy))
        return None


class BatchNorm2d(SibModule):
    """BatchNorm (+ fused activation), nn.BatchNorm2d-compatible state.  `sync` turns the batch
    statistics into cross-rank statistics (SyncBN) with one 2C-float all-reduce per pass."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, activation="identity", slope=0.01):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.activation, self.slope = activation, slope
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self.process_group = None
        self.sync = False

    def extra_repr(self):
        return "%d, eps=%g, momentum=%g, activation=%s" % (self.num_features, self.eps,
                                                             self.momentum, self.activation)

    @property
    def act(self):
        return ops.ACT_CODES[self.activation]

    def _world(self):
        if self.sync and torch.distributed.is_available() and torch.distributed.is_initialized():
            return torch.distributed.get_world_size(self.process_group)
        return 1

    def finalize(self, stats, count, train):
        """stats [2,C] (sum, sumsq of this rank) -> (mean_invstd, scale_shift, total count)."""
        if not train:
            return None, ops.bn_eval_scale(self.weight.data, self.bias.data, self.running_mean,
                                           self.running_var, self.eps), count
        world = self._world()
        if world > 1:
            ops.small_allreduce_(stats, self.process_group)
            count = count * world
        mi, ss = ops.bn_finalize(stats, self.weight.data, self.bias.data, self.running_mean,
                                 self.running_var, count, self.eps, self.momentum)
        self._count_batch()
        return mi, ss, count

    def stats_args(self, stats):
        """(stats, gamma, beta, running_mean, running_var) for the fused finalize+apply kernel;
        all-reduces the statistics first under SyncBN.  Returns (args, world)."""
        world = self._world()
        if world > 1:
            ops.small_allreduce_(stats, self.process_group)
        self._count_batch()
        return (stats, self.weight.data, self.bias.data, self.running_mean, self.running_var), world

    def _count_batch(self):
        if not getattr(self, "_nbt_batched", False):     # else the root module counts for everyone
            self.num_batches_tracked += 1

    def reduce_sums(self, sums):
        if self._world() > 1:
            ops.small_allreduce_(sums, self.process_group)
        return sums

    def grad_ptrs(self):
        """(dgamma, dbeta, scale) for bn_bwd_apply, which accumulates the parameter gradients itself.
        Under SyncBN the backward sums are totals over ALL ranks while DataParallel averages the
        parameter gradients over the ranks afterwards, hence scale = 1 / world (torch's SyncBatchNorm
        takes the rank-local sums for the same reason)."""
        return self._grad(self.weight), self._grad(self.bias), 1.0 / self._world()

    # standalone use: stats pass + apply
    def fwd(self, x, train, res=None):
        n, c, h, w = x.shape
        stats = ops.bn_stats(x) if train else None
        mi, ss, count = self.finalize(stats, n * h * w, train)
        y = ops.bn_apply(x, ss, self.act, self.slope, res=res)
        return y, (x, y if res is not None else None, mi, count, ss)

    def bwd(self, dy, saved, need_dx=True):
        x, y, mi, count, ss = saved
        dy = _as_act(dy)
        sums = self.reduce_sums(ops.bn_bwd_reduce(dy, y, x, mi, self.act, self.slope, mask_ss=ss))
        dx, _, _ = ops.bn_bwd_apply(dy, y, x, mi, self.weight.data, sums, count, self.act, self.slope,
                                    mask_ss=ss, param_grads=self.grad_ptrs())
        return dx


class MaxPool3x3s2(SibModule):
    def fwd(self, x, train):
        y, idx = ops.maxpool3x3s2_fwd(x, want_idx=train)
        return y, (idx, tuple(x.shape))

    def bwd(self, dy, saved, need_dx=True):
        idx, shape = saved
        return ops.maxpool3x3s2_bwd(_as_act(dy), idx, shape)


class Linear(SibModule):
    """Fully connected head run as a 1x1 convolution on the same tcgen05 kernel; bias fused in
    the epilogue.  weight [out, in] fp32 like nn.Linear."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        assert in_features % 8 == 0 and out_features % 8 == 0
        w = torch.empty(out_features, in_features)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        self.weight = nn.Parameter(w.view(out_features, in_features, 1, 1))
        self.weight._sib_layout = "krsc"
        self.weight._sib_needs_dgrad = True
        self._register_state_dict_hook(_linear_sd_hook)
        self._register_load_state_dict_pre_hook(_linear_load_hook)
        if bias:
            bound = 1 / math.sqrt(in_features)
            self.bias = nn.Parameter(torch.empty(out_features).uniform_(-bound, bound))
        else:
            self.register_parameter("bias", None)

    def fwd(self, x, train):
        """x [N, in, 1, 1] bf16 -> logits [N, out] bf16"""
        x4 = x if x.dim() == 4 else x.view(x.shape[0], x.shape[1], 1, 1)
        y = ops.conv2d_fprop(x4, self._w16(self.weight), 1, 0,
                             bias=self.bias.data if self.bias is not None else None)
        return y.reshape(y.shape[0], y.shape[1]), (x4,)

    def bwd(self, dy, saved, need_dx=True):
        (x4,) = saved
        n = x4.shape[0]
        dy4 = _as_act(dy.reshape(n, self.out_features, 1, 1))
        ops.conv2d_wgrad(x4, dy4, self._grad(self.weight), 1, 0)
        if self.bias is not None:
            col = ops.bn_stats(dy4)       # row 0 = per-column sum of dlogits
            self._grad(self.bias).add_(col[0])
        if not need_dx:
            return None
        return ops.conv2d_dgrad(dy4, self._wd16(self.weight), tuple(x4.shape), 1, 1, 1, 0)


def _linear_sd_hook(module, state_dict, prefix, local_metadata):
    # expose the nn.Linear shape [out, in] in checkpoints
    key = prefix + "weight"
    if key in state_dict:
        state_dict[key] = state_dict[key].reshape(module.out_features, module.in_features)


def _linear_load_hook(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                      error_msgs):
    key = prefix + "weight"
    if key in state_dict and state_dict[key].dim() == 2:
        w = state_dict[key]
        state_dict[key] = w.reshape(w.shape[0], w.shape[1], 1, 1)


def _as_act(t):
    """Gradient tensors arriving from autograd may be fp32 / NCHW-contiguous: normalise."""
    if t.dtype == torch.bfloat16 and t.dim() == 4 and t.permute(0, 2, 3, 1).is_contiguous():
        return t
    return ops.to_nhwc_bf16(t)


# --------------------------------------------------------------------------- residual block
class Bottleneck(SibModule):
    """ResNet v1.5 bottleneck (stride on the 3x3), torchvision naming.  Forward:
         conv1 -> bn1+act -> conv2 -> bn2+act -> conv3 -> bn3 (+ bn_ds(conv_ds(x)) | + x) -> act
    BN statistics come out of the conv epilogues; bn3 + shortcut BN + add + act is one kernel.
    """
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=False, activation="relu", slope=0.01):
        super().__init__()
        out = planes * self.expansion
        self.conv1 = Conv2d(inplanes, planes, 1)
        self.bn1 = BatchNorm2d(planes, activation=activation, slope=slope)
        self.conv2 = Conv2d(planes, planes, 3, stride=stride, padding=1)
        self.bn2 = BatchNorm2d(planes, activation=activation, slope=slope)
        self.conv3 = Conv2d(planes, out, 1)
        self.bn3 = BatchNorm2d(out, activation=activation, slope=slope)
        self.stride = stride
        if downsample:
            ds_bn = BatchNorm2d(out, activation="identity")
            self.downsample = nn.Sequential(Conv2d(inplanes, out, 1, stride=stride), ds_bn)
        else:
            self.downsample = None

    @staticmethod
    def _conv(conv, x, train):
        if train and ops.DETERMINISTIC:       # statistics by the fixed-order reduction, not the epilogue atomics
            y = conv.run(x, None)
            return y, ops.bn_stats(y)
        stats = ops.new_acc(2, conv.out_channels, x.device) if train else None
        return conv.run(x, stats), stats

    @staticmethod
    def _bn_act(bn, c, stats, train, res=None, bn2=None, stats2=None):
        """act(bn(c) [+ res | + bn2(res)]) -> y, mi, ss, count[, mi2]."""
        n, _, h, w = c.shape
        count = n * h * w
        if not train:
            ss = ops.bn_eval_scale(bn.weight.data, bn.bias.data, bn.running_mean, bn.running_var, bn.eps)
            ss2 = None
            if bn2 is not None:
                ss2 = ops.bn_eval_scale(bn2.weight.data, bn2.bias.data, bn2.running_mean, bn2.running_var, bn2.eps)
            return ops.bn_apply(c, ss, bn.act, bn.slope, res=res, scale_shift2=ss2), None, ss, count, None
        a1, world = bn.stats_args(stats)
        a2 = bn2.stats_args(stats2)[0] if bn2 is not None else None
        y, (mi, ss), (mi2, _) = ops.bn_finalize_apply(c, a1, res=res, bn2=a2, act=bn.act, slope=bn.slope,
                                                      count=count * world, eps=bn.eps, momentum=bn.momentum)
        return y, mi, ss, count * world, mi2

    @staticmethod
    def _conv_bnact(conv, bn, c, stats_c, train):
        """conv(act(bn(c))) with the BatchNorm applied in the conv's operand prologue.
        -> y, stats(y), mean_invstd, scale_shift, count (of bn)."""
        n, _, h, w = c.shape
        count = n * h * w
        w16 = conv._w16(conv.weight)
        if not train:
            ss = ops.bn_eval_scale(bn.weight.data, bn.bias.data, bn.running_mean, bn.running_var, bn.eps)
            y, _, _ = ops.conv2d_fprop_bnact(c, w16, None, conv.stride, conv.padding, act=bn.act,
                                             slope=bn.slope, scale_shift=ss)
            return y, None, None, ss, count
        args, world = bn.stats_args(stats_c)
        st = ops.new_acc(2, conv.out_channels, c.device)
        y, mi, ss = ops.conv2d_fprop_bnact(c, w16, args, conv.stride, conv.padding, stats=st, act=bn.act,
                                           slope=bn.slope, count=count * world, eps=bn.eps,
                                           momentum=bn.momentum)
        return y, st, mi, ss, count * world

    def fwd(self, x, train):
        c1, st1 = self._conv(self.conv1, x, train)
        planes = self.conv2.in_channels
        if _fuse_fwd("conv2", planes, self.stride):
            a1 = None
            c2, st2, mi1, ss1, cnt1 = self._conv_bnact(self.conv2, self.bn1, c1, st1, train)
        else:
            a1, mi1, ss1, cnt1, _ = self._bn_act(self.bn1, c1, st1, train)
            c2, st2 = self._conv(self.conv2, a1, train)
        if _fuse_fwd("conv3", planes, self.stride):
            a2 = None
            c3, st3, mi2, ss2, cnt2 = self._conv_bnact(self.conv3, self.bn2, c2, st2, train)
        else:
            a2, mi2, ss2, cnt2, _ = self._bn_act(self.bn2, c2, st2, train)
            c3, st3 = self._conv(self.conv3, a2, train)
        if self.downsample is not None:
            cd, std = self._conv(self.downsample[0], x, train)
            out, mi3, _, cnt3, mid = self._bn_act(self.bn3, c3, st3, train, res=cd, bn2=self.downsample[1],
                                                  stats2=std)
        else:
            cd = mid = None
            out, mi3, _, cnt3, _ = self._bn_act(self.bn3, c3, st3, train, res=x)
        if not train:
            return out, None
        return out, (x, c1, mi1, a1, c2, mi2, a2, c3, mi3, cd, mid, out, cnt1, cnt2, cnt3, ss1, ss2)

    def fuse_info(self, saved):
        """What the NEXT block's final dgrad needs to fuse this block's bn3 + add + act backward
        reduction into its epilogue (identity-shortcut blocks only): the mask comes from the stored
        block output, xhat from conv3's output."""
        if not FUSE_BN_BWD or self.downsample is not None or saved is None:
            return None
        c3, mi3, out = saved[7], saved[8], saved[11]
        return dict(mask_src=out, mask_ss=None, xhat_src=c3, mean_invstd=mi3, act=self.bn3.act,
                    slope=self.bn3.slope)

    def bwd(self, dout, saved, need_dx=True, dout_sums=None, prev=None):
        """dout: gradient w.r.t. the block output.  With `dout_sums` it is already masked by this
        block's final activation and the bn3 reduction (sum g, sum g*xhat) is `dout_sums` (both
        produced by the next block's dgrad epilogue).  `prev` = fuse_info() of the preceding block:
        when this block can honour it, returns (dx_masked, sums_for_prev) instead of dx."""
        x, c1, mi1, a1, c2, mi2, a2, c3, mi3, cd, mid, out, cnt1, cnt2, cnt3, ss1, ss2 = saved
        dout = _as_act(dout)
        bn1, bn2, bn3 = self.bn1, self.bn2, self.bn3
        none = ops.ACT_CODES["identity"]
        # ---- bn3 (+ shortcut bn) + add + act ----
        if self.downsample is not None:
            assert dout_sums is None
            bnd = self.downsample[1]
            sums = bn3.reduce_sums(ops.bn_bwd_reduce(dout, out, c3, mi3, bn3.act, bn3.slope, x2=cd,
                                                     mean_invstd2=mid))
            dc3, dcd, _ = ops.bn_bwd_apply(dout, out, c3, mi3, bn3.weight.data, sums, cnt3, bn3.act,
                                           bn3.slope, x2=cd, mean_invstd2=mid,
                                           gamma2=bnd.weight.data, param_grads=bn3.grad_ptrs(),
                                           param_grads2=bnd.grad_ptrs())
            g = None
        elif dout_sums is not None:
            # dout is g (masked): it is also the identity-shortcut gradient
            sums = bn3.reduce_sums(dout_sums)
            dc3, dcd, _ = ops.bn_bwd_apply(dout, None, c3, mi3, bn3.weight.data, sums, cnt3, none, 0.0,
                                           param_grads=bn3.grad_ptrs())
            g = dout
        else:
            sums = bn3.reduce_sums(ops.bn_bwd_reduce(dout, out, c3, mi3, bn3.act, bn3.slope))
            dc3, dcd, g = ops.bn_bwd_apply(dout, out, c3, mi3, bn3.weight.data, sums, cnt3, bn3.act,
                                           bn3.slope, want_g=need_dx, param_grads=bn3.grad_ptrs())
        # ---- conv3, bn2 + act ----
        # (activation mask recomputed from c2 and the forward scale/shift: a2 is not re-read)
        # a2 / a1 are None when the forward applied the BatchNorm inside the consumer conv's
        # prologue: bn_bwd_apply then re-materialises them and the weight gradient follows it
        if a2 is not None:
            self.conv3.run_wgrad(a2, dc3)
        if FUSE_BN_BWD:
            da2, sums = self.conv3.run_dgrad(dc3, tuple(c2.shape), bn_bwd=dict(
                mask_src=c2, mask_ss=ss2, mean_invstd=mi2, act=bn2.act, slope=bn2.slope))
            sums, mask = bn2.reduce_sums(sums), (none, 0.0, None)
        else:
            da2 = self.conv3.run_dgrad(dc3, tuple(c2.shape))
            sums = bn2.reduce_sums(ops.bn_bwd_reduce(da2, None, c2, mi2, bn2.act, bn2.slope, mask_ss=ss2))
            mask = (bn2.act, bn2.slope, ss2)
        if a2 is not None:
            dc2, _, _ = ops.bn_bwd_apply(da2, None, c2, mi2, bn2.weight.data, sums, cnt2, mask[0], mask[1],
                                         mask_ss=mask[2], param_grads=bn2.grad_ptrs())
        else:
            dc2, a2 = ops.bn_bwd_apply_remat(da2, c2, mi2, bn2.weight.data, sums, cnt2, ss2, bn2.act, bn2.slope,
                                             act=mask[0], slope=mask[1], mask_ss=mask[2],
                                             param_grads=bn2.grad_ptrs())
            self.conv3.run_wgrad(a2, dc3)
        # ---- conv2, bn1 + act ----
        if a1 is not None:
            self.conv2.run_wgrad(a1, dc2)
        # Measured (B200, batch 256): the fused reduction LOSES where conv2's dgrad runs on the
        # halo-reuse kernel (64 -> 64 at 56x56: 0.172 ms fused vs 0.062 + 0.045 ms separate; its
        # epilogue reads the BN input with exposed global loads) and on the row-parity stride-2
        # path (0.195 vs 0.097 + 0.078 ms); everywhere else it wins or is neutral.
        fuse1 = FUSE_BN_BWD and (FUSE_BN_BWD_ALL or self.stride == 1) and not (not FUSE_BN_BWD_ALL and ops.halo_applies(c1.shape[1], dc2.shape[1], 3, 1,
                                                                                                    c1.shape[3]))
        if fuse1:
            da1, sums = self.conv2.run_dgrad(dc2, tuple(c1.shape), bn_bwd=dict(
                mask_src=c1, mask_ss=ss1, mean_invstd=mi1, act=bn1.act, slope=bn1.slope))
            sums, mask = bn1.reduce_sums(sums), (none, 0.0, None)
        else:
            da1 = self.conv2.run_dgrad(dc2, tuple(c1.shape))
            sums = bn1.reduce_sums(ops.bn_bwd_reduce(da1, None, c1, mi1, bn1.act, bn1.slope, mask_ss=ss1))
            mask = (bn1.act, bn1.slope, ss1)
        if a1 is not None:
            dc1, _, _ = ops.bn_bwd_apply(da1, None, c1, mi1, bn1.weight.data, sums, cnt1, mask[0], mask[1],
                                         mask_ss=mask[2], param_grads=bn1.grad_ptrs())
        else:
            dc1, a1 = ops.bn_bwd_apply_remat(da1, c1, mi1, bn1.weight.data, sums, cnt1, ss1, bn1.act, bn1.slope,
                                             act=mask[0], slope=mask[1], mask_ss=mask[2],
                                             param_grads=bn1.grad_ptrs())
            self.conv2.run_wgrad(a1, dc2)
        # ---- conv1 (+ shortcut) ----
        self.conv1.run_wgrad(x, dc1)
        if self.downsample is not None:
            self.downsample[0].run_wgrad(x, dcd)
        if not need_dx:
            return None
        if self.downsample is None:
            # identity shortcut: dx = dgrad(conv1) + g, the add fused into the conv epilogue
            return self.conv1.run_dgrad(dc1, tuple(x.shape), residual=g, bn_bwd=prev)
        dx = self.conv1.run_dgrad(dc1, tuple(x.shape))
        if self.stride == 1:
            return self.downsample[0].run_dgrad(dcd, tuple(x.shape), out=dx, residual=dx, bn_bwd=prev)
        assert prev is None, "a strided shortcut ends in a scatter-add and cannot fuse the reduction"
        return self.downsample[0].run_dgrad(dcd, tuple(x.shape), out=dx)   # strided scatter-add

    def can_fuse_prev(self):
        """True when this block's input gradient comes out of ONE stride-1 dgrad epilogue."""
        return FUSE_BN_BWD and (self.downsample is None or self.stride == 1)


class GlobalAvgPool(SibModule):
    """FastGlobalAvgPool2d(flatten=True) of pytorch_tools: [N,C,H,W] -> [N,C,1,1] bf16."""

    def fwd(self, x, train):
        return ops.gap_fwd(x), (tuple(x.shape),)

    def bwd(self, dy, saved, need_dx=True):
        dy = dy.reshape(dy.shape[0], dy.shape[1], 1, 1)
        return ops.gap_bwd(_as_act(dy), saved[0])


class Concat(nn.Module):
    """channel concatenation of tagged inputs (reference model.py:1109-1111)"""

    def forward(self, *args):
        return torch.cat(args, dim=1).contiguous(memory_format=torch.channels_last)
