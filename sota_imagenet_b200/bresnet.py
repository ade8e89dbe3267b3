"""BResNet-50: `pytorch_tools.models.resnet50(**model_params)` with the kwargs of the reference's
BResNet configs (configs/_old_configs/_first_attempts/BResNet50_encoder.yaml:44-60, BResNet50.yaml:7-10):
deep stem, anti-aliased down-sampling (BlurPool, AvgPool shortcut), ECA attention, in-place ABN with
leaky ReLU, dropout, drop-connect, weight standardisation, EMA (optimizer side).

pytorch_tools is absent and unpinned, so the module graph below is this repo's restatement of those
switches; oracle/bresnet_ref.py is its fp32 PyTorch twin used for parity.  Convs
run on the tcgen05 kernels; the extra operators are the memory-bound kernels of csrc/extra.cu.  The
block is composed from per-operator fwd/bwd pairs (not yet fused like the ResNet-50 bottleneck)."""
import os

import torch
import torch.nn as nn

from . import _lib, ops
from .modules import BatchNorm2d, Conv2d, Linear, SibModule, StemConv, _as_act


class PaddedConv2d(SibModule):
    """3x3 convolution whose logical channel counts (32 in the deep stem) are zero-padded to the 64
    the implicit-GEMM k-block needs.  Tiny filters: padding / packing / slicing use torch ops."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, phys_in=64, phys_out=64):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding = (kernel_size, kernel_size), stride, padding
        self.phys_in, self.phys_out = phys_in, phys_out
        w = torch.empty(out_channels, in_channels, kernel_size, kernel_size)
        nn.init.kaiming_normal_(w, mode="fan_out", nonlinearity="relu")
        self.weight = nn.Parameter(w)
        self.ws = False
        self.ws_eps = 1e-7

    def _effective_weight(self):
        w = self.weight.data
        if self.ws:
            var, mean = torch.var_mean(w, dim=(1, 2, 3), keepdim=True, unbiased=False)
            self._ws_stats = (mean, torch.rsqrt(var + self.ws_eps))
            w = (w - mean) * self._ws_stats[1]
        return w

    def _packed(self):
        k = self.kernel_size[0]
        w = self._effective_weight()
        full = torch.zeros((self.phys_out, k, k, self.phys_in), dtype=torch.bfloat16, device=w.device)
        full[:self.out_channels, :, :, :self.in_channels] = w.permute(0, 2, 3, 1)
        return full

    def run(self, x, stats=None):
        full = self._packed()
        self._wd = full.flip(1, 2).permute(3, 1, 2, 0).contiguous()
        return ops.conv2d_fprop(x, full.permute(0, 3, 1, 2), self.stride, self.padding, stats=stats)

    def run_dgrad(self, dy, x_shape):
        k = self.kernel_size[0]
        return ops.conv2d_dgrad(dy, self._wd, x_shape, k, k, self.stride, self.padding)

    def run_wgrad(self, x, dy):
        k = self.kernel_size[0]
        dw = torch.zeros((self.phys_out, k, k, self.phys_in), dtype=torch.float32, device=dy.device).permute(0, 3, 1, 2)
        ops.conv2d_wgrad(x, dy, dw, self.stride, self.padding)
        g = dw[:self.out_channels, :self.in_channels]
        if self.ws:
            mean, invstd = self._ws_stats
            what = (self.weight.data - mean) * invstd
            g = invstd * (g - g.mean(dim=(1, 2, 3), keepdim=True) - what * (g * what).mean(dim=(1, 2, 3), keepdim=True))
        self._grad(self.weight).add_(g)

    def fwd(self, x, train):
        return self.run(x), (x,)

    def bwd(self, dy, saved, need_dx=True):
        dy = _as_act(dy)
        self.run_wgrad(saved[0], dy)
        return self.run_dgrad(dy, tuple(saved[0].shape)) if need_dx else None


class DeepStemFirst(StemConv):
    """3x3/2 conv 3 -> 32 through the row-pair packing; emits 64 physical channels (32 zero)."""

    def __init__(self, out_channels=32, phys_out=64):
        super().__init__(out_channels, 3, 1)
        self.phys_out = phys_out
        self.ws = False
        self.ws_eps = 1e-7

    def _effective_weight(self):
        return PaddedConv2d._effective_weight(self)

    def _packed_weight(self):
        w = self._effective_weight().contiguous()
        wq = ops.stem_pack_weight(w, self.na, self.off)                      # [K, NA, 1, 64]
        full = torch.zeros((self.phys_out, self.na, 1, 64), dtype=torch.bfloat16, device=w.device)
        full[:self.out_channels] = wq
        return full.permute(0, 3, 1, 2)

    def run_wgrad(self, xq, dy):
        dwq = torch.zeros((self.phys_out, self.na, 1, 64), dtype=torch.float32, device=dy.device).permute(0, 3, 1, 2)
        ops.conv2d_wgrad(xq, dy, dwq, 1, 0, pad_hw=(self.a0, 0))
        g = torch.zeros_like(self.weight.data)
        ops.stem_unpack_wgrad(dwq[:self.out_channels].contiguous(memory_format=torch.channels_last), g,
                              self.na, self.off, accumulate=False)
        if self.ws:
            mean, invstd = self._ws_stats
            what = (self.weight.data - mean) * invstd
            g = invstd * (g - g.mean(dim=(1, 2, 3), keepdim=True) - what * (g * what).mean(dim=(1, 2, 3), keepdim=True))
        self._grad(self.weight).add_(g)


class PaddedBatchNorm2d(BatchNorm2d):
    """BN over `num_features` logical channels living in a wider physical tensor: the pad channels
    are all-zero, stay zero (gamma = 1, beta = 0 there) and carry no parameters."""

    def __init__(self, num_features, phys, **kw):
        super().__init__(num_features, **kw)
        self.phys = phys

    def _padded(self, t, fill):
        out = torch.full((self.phys,), fill, dtype=torch.float32, device=t.device)
        out[:self.num_features] = t
        return out

    def fwd(self, x, train, stats=None):
        n, c, h, w = x.shape
        if train:
            stats = stats if stats is not None else ops.bn_stats(x)
            world = self._world()
            count = n * h * w
            if world > 1:
                ops.small_allreduce_(stats, self.process_group)
                count *= world
            rm, rv = self._padded(self.running_mean, 0.0), self._padded(self.running_var, 1.0)
            mi, ss = ops.bn_finalize(stats, self._padded(self.weight.data, 1.0), self._padded(self.bias.data, 0.0),
                                     rm, rv, count, self.eps, self.momentum)
            self.running_mean.copy_(rm[:self.num_features])
            self.running_var.copy_(rv[:self.num_features])
            self._count_batch()
        else:
            mi, count = None, n * h * w
            ss = ops.bn_eval_scale(self._padded(self.weight.data, 1.0), self._padded(self.bias.data, 0.0),
                                   self._padded(self.running_mean, 0.0), self._padded(self.running_var, 1.0), self.eps)
        y = ops.bn_apply(x, ss, self.act, self.slope)
        return y, (x, mi, count, ss)

    def bwd(self, dy, saved, need_dx=True):
        x, mi, count, ss = saved
        dy = _as_act(dy)
        sums = self.reduce_sums(ops.bn_bwd_reduce(dy, None, x, mi, self.act, self.slope, mask_ss=ss))
        gamma = self._padded(self.weight.data, 1.0)
        dx, _, _ = ops.bn_bwd_apply(dy, None, x, mi, gamma, sums, count, self.act, self.slope, mask_ss=ss)
        pg = 1.0 / self._world()          # sums are totals over all ranks under SyncBN (see grad_ptrs)
        self._grad(self.bias).add_(sums[0, :self.num_features], alpha=pg)
        self._grad(self.weight).add_(sums[1, :self.num_features], alpha=pg)
        return dx


class BNAct(BatchNorm2d):
    """ABN / InplaceABN: BN + activation, statistics from the producing conv's epilogue when given."""

    def fwd(self, x, train, stats=None):
        n, c, h, w = x.shape
        if not train:
            mi, ss, count = self.finalize(None, n * h * w, False)
            return ops.bn_apply(x, ss, self.act, self.slope), (x, mi, count, ss)
        if stats is None:
            stats = ops.bn_stats(x)
        # finalize (+ running statistics) and apply in ONE launch (was bn_finalize + bn_apply)
        args, world = self.stats_args(stats)
        count = n * h * w * world
        y, (mi, ss), _ = ops.bn_finalize_apply(x, args, act=self.act, slope=self.slope, count=count,
                                               eps=self.eps, momentum=self.momentum)
        return y, (x, mi, count, ss)

    def fuse_info(self, saved):
        """What a dgrad needs to fuse this BatchNorm's backward reduction (and activation mask) into
        its epilogue (ops.conv2d_dgrad bn_bwd=...)."""
        x, mi, count, ss = saved
        return dict(mask_src=x, mask_ss=ss, mean_invstd=mi, act=self.act, slope=self.slope)

    def bwd(self, dy, saved, need_dx=True, sums=None):
        """`sums` given: dy is already masked by this BN's activation and (sum g, sum g*xhat) come
        from the producing dgrad's epilogue -- the separate reduction pass is skipped."""
        x, mi, count, ss = saved
        dy = _as_act(dy)
        if sums is not None:
            dx, _, _ = ops.bn_bwd_apply(dy, None, x, mi, self.weight.data, self.reduce_sums(sums), count,
                                        ops.ACT_NONE, 0.0, param_grads=self.grad_ptrs())
            return dx
        sums = self.reduce_sums(ops.bn_bwd_reduce(dy, None, x, mi, self.act, self.slope, mask_ss=ss))
        dx, _, _ = ops.bn_bwd_apply(dy, None, x, mi, self.weight.data, sums, count, self.act, self.slope,
                                    mask_ss=ss, param_grads=self.grad_ptrs())
        return dx


class BlurPool(SibModule):
    def __init__(self, channels=0):
        super().__init__()
        self.channels = channels

    def fwd(self, x, train):
        return ops.blurpool_fwd(x), (tuple(x.shape),)

    def bwd(self, dy, saved, need_dx=True):
        return ops.blurpool_bwd(_as_act(dy), saved[0])


# Fold the drop-connect keep mask into the ECA gate (one scale pass instead of two, forward and
# backward); SIB_FUSE_DROP_CONNECT=0 keeps the two-pass sequence (A/B and parity cross-check).
FUSE_DROP_CONNECT = os.environ.get("SIB_FUSE_DROP_CONNECT", "1") != "0"
# Apply the ECA gate inside the shortcut-add + activation pass (act(x*gate + r) in one kernel: the
# gated tensor is never written); SIB_FUSE_ECA_TAIL=0 keeps scale_nc + add_act.
FUSE_ECA_TAIL = os.environ.get("SIB_FUSE_ECA_TAIL", "1") != "0"
FUSE_BN_BWD = os.environ.get("SIB_FUSE_BN_BWD", "1") != "0"
# Block tail without the normalised bn3 output: the ECA gate pools the RAW conv3 output (BatchNorm is
# affine per channel, so mean_hw(bn3(c3)) = scale * mean_hw(c3) + shift), the output pass computes
# act(c3 * (scale * gate) + shift * gate + shortcut) straight from c3, and in backward ONE pass yields
# the masked gradient plus the per-(sample, channel) sums from which both the gate's gradient and the
# BatchNorm-backward sums follow algebraically; the gate scale rides inside bn_bwd_apply.  12 passes
# over the widest tensor of the block become 7.  SIB_FUSE_BN3_TAIL=0 restores the operator sequence.
FUSE_BN3_TAIL = os.environ.get("SIB_FUSE_BN3_TAIL", "1") != "0"


class ECA(SibModule):
    """x * sigmoid(conv1d_k3(GAP(x))) over the channel axis (attn_type: eca)."""

    def __init__(self, channels=None, kernel_size=3):
        super().__init__()
        assert kernel_size == 3
        self.weight = nn.Parameter(torch.empty(1, 1, 3).uniform_(-0.5, 0.5))

    def fwd(self, x, train, extra=None, defer=False):
        """`extra` [N,C]: a further per-(sample, channel) factor applied in the same pass (the
        drop-connect keep mask of the block: y = x * gate * extra), saving one full read + write of
        the activation in forward and one in backward."""
        n, c, h, w = x.shape
        p = ops.chan_reduce(x, scale=1.0 / (h * w))
        s = ops.eca_gate_fwd(p, self.weight.data.view(3).contiguous())
        saved = (x, p, s) if extra is None else (x, p, s, extra)
        se = s if extra is None else s * extra
        if defer:                        # the caller applies the scale inside its own fused pass
            return se, saved
        return ops.scale_nc(x, se), saved

    def bwd(self, dy, saved, need_dx=True):
        x, p, s = saved[:3]
        extra = saved[3] if len(saved) > 3 else None
        dy = _as_act(dy)
        n, c, h, w = x.shape
        ds = ops.chan_reduce(dy, x)
        if extra is not None:            # d/d(gate) of x*gate*extra; extra is constant over (h, w)
            ds = ds * extra
        dw = torch.zeros(3, dtype=torch.float32, device=x.device)
        dp = ops.eca_gate_bwd(ds, s, p, self.weight.data.view(3).contiguous(), dw)
        self._grad(self.weight).view(3).add_(dw)
        return ops.scale_nc(dy, s if extra is None else s * extra, add=dp / (h * w))


class BBottleneck(SibModule):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=False, norm_act="leaky_relu", antialias=True,
                 attn=True, keep_prob=1.0):
        super().__init__()
        out = planes * self.expansion
        self.stride, self.antialias, self.keep_prob = stride, antialias, keep_prob
        conv_stride = 1 if (antialias and stride > 1) else stride
        self.conv1 = Conv2d(inplanes, planes, 1)
        self.bn1 = BNAct(planes, activation=norm_act)
        self.conv2 = Conv2d(planes, planes, 3, stride=conv_stride, padding=1)
        self.bn2 = BNAct(planes, activation=norm_act)
        self.blur = BlurPool(planes) if (antialias and stride > 1) else None
        self.conv3 = Conv2d(planes, out, 1)
        self.bn3 = BNAct(out, activation="identity")
        self.eca = ECA(out) if attn else None
        self.act = ops.ACT_CODES[norm_act]
        if downsample:
            ds_stride = 1 if (antialias and stride > 1) else stride
            self.downsample = nn.Sequential(Conv2d(inplanes, out, 1, stride=ds_stride), BNAct(out, activation="identity"))
            self.pool_shortcut = antialias and stride > 1
        else:
            self.downsample = None
            self.pool_shortcut = False

    @staticmethod
    def _conv_bn(conv, bn, x, train):
        stats = ops.new_acc(2, conv.out_channels, x.device) if train else None
        c = conv.run(x, stats)
        y, s = bn.fwd(c, train, stats=stats)
        return y, s

    def fwd(self, x, train):
        a1, s1 = self._conv_bn(self.conv1, self.bn1, x, train)
        a2, s2 = self._conv_bn(self.conv2, self.bn2, a1, train)
        sb = None
        a2b = a2
        if self.blur is not None:
            a2b, sb = self.blur.fwd(a2, train)
        fused_tail = train and self.eca is not None and FUSE_BN3_TAIL
        if fused_tail:
            return self._fwd_fused_tail(x, a1, s1, a2, s2, a2b, sb)
        y3, s3 = self._conv_bn(self.conv3, self.bn3, a2b, train)
        se = None
        mask = None
        if train and self.keep_prob < 1.0:
            n, c = y3.shape[0], y3.shape[1]
            keep = (torch.rand(n, 1, device=y3.device) < self.keep_prob).float() / self.keep_prob
            mask = keep.expand(n, c).contiguous()
        fused = mask is not None and self.eca is not None and FUSE_DROP_CONNECT
        tail_scale = None                # ECA gate (* keep mask) applied inside the add + act pass
        if self.eca is not None:
            if FUSE_ECA_TAIL and (mask is None or fused):
                tail_scale, se = self.eca.fwd(y3, train, extra=mask if fused else None, defer=True)
            else:
                y3, se = self.eca.fwd(y3, train, extra=mask if fused else None)
        if mask is not None and not fused:
            y3 = ops.scale_nc(y3, mask)
        xs, sp, sd = x, None, None
        if self.downsample is not None:
            if self.pool_shortcut:
                xs = ops.avgpool2_fwd(x)
            r, sd = self._conv_bn(self.downsample[0], self.downsample[1], xs, train)
        else:
            r = x
        if tail_scale is not None:
            out = ops.scale_add_act(y3, tail_scale, r, self.act, self.bn1.slope)
        else:
            out = ops.add_act(y3, r, self.act, self.bn1.slope)
        if not train:
            return out, None
        return out, (x, a1, s1, a2, s2, a2b, sb, s3, se, mask, xs, sd, out)

    def _shortcut(self, x, train):
        xs, sd = x, None
        if self.downsample is not None:
            if self.pool_shortcut:
                xs = ops.avgpool2_fwd(x)
            r, sd = self._conv_bn(self.downsample[0], self.downsample[1], xs, train)
        else:
            r = x
        return r, xs, sd

    def _fwd_fused_tail(self, x, a1, s1, a2, s2, a2b, sb):
        """Training forward of conv3 .. block output without materialising bn3's output (see
        FUSE_BN3_TAIL)."""
        bn3 = self.bn3
        st3 = ops.new_acc(2, self.conv3.out_channels, x.device)
        c3 = self.conv3.run(a2b, st3)
        n, c, h, w = c3.shape
        args, world = bn3.stats_args(st3)
        cnt3 = n * h * w * world
        mi3, ss3 = ops.bn_finalize(*args, cnt3, bn3.eps, bn3.momentum)
        pc = ops.chan_reduce(c3, scale=1.0 / (h * w))              # mean_hw of the raw conv output
        keep = None
        if self.keep_prob < 1.0:
            keep = (torch.rand(n, 1, device=c3.device) < self.keep_prob).float() / self.keep_prob
        # p = mean_hw(bn3(c3)) (what ECA pools), the gate (* keep) and the output pass's coefficients
        p, gate, gate_k, mul, add = (torch.empty_like(pc) for _ in range(5))
        ops.call("sib_eca_tail_fwd", ops._p(pc), ops._p(ss3), ops._p(self.eca.weight.data.view(3)), ops._p(keep),
                 ops._p(p), ops._p(gate), ops._p(gate_k), ops._p(mul), ops._p(add), n, c, ops._stream())
        r, xs, sd = self._shortcut(x, True)
        out = ops.scale_add_act(c3, mul, r, self.act, self.bn1.slope, add=add)
        tail = (c3, mi3, ss3, cnt3, pc, p, gate, gate_k, keep)
        return out, (x, a1, s1, a2, s2, a2b, sb, ("fused", tail), None, None, xs, sd, out)

    def _bwd_fused_tail(self, dout, out, tail):
        """-> g (masked block-output gradient = shortcut gradient), dc3."""
        c3, mi3, ss3, cnt3, pc, p, gate, gate_k, keep = tail
        bn3 = self.bn3
        n, c, h, w = c3.shape
        hw = float(h * w)
        g, s1, s2 = ops.act_bwd_reduce(_as_act(dout), out, c3, self.act, self.bn1.slope)
        # one small kernel: ECA filter gradient (straight into the gradient arena), the pooled-path term
        # add_nc = dp / HW, and the BatchNorm-backward sums of d = g * gate_k + add_nc derived from the
        # per-(sample, channel) sums (sum_hw g * xhat = invstd * (s2 - mean * s1), sum_hw xhat = invstd * HW *
        # (pc - mean))
        add_nc = torch.empty_like(s1)
        sums = ops.new_acc(2, c, c3.device)
        if not ops._ACC_POOL.owns(sums):
            sums.zero_()
        ops.call("sib_eca_tail_bwd", ops._p(s1), ops._p(s2), ops._p(ss3), ops._p(mi3), ops._p(pc), ops._p(p),
                 ops._p(gate), ops._p(gate_k), ops._p(keep), ops._p(self.eca.weight.data.view(3)), ops._p(add_nc),
                 ops._p(sums), ops._p(self.eca._grad(self.eca.weight).view(3)), n, c, hw, ops._stream())
        sums = bn3.reduce_sums(sums)
        dc3 = ops.bn_bwd_apply_scaled(g, gate_k, add_nc, c3, mi3, bn3.weight.data, sums,
                                      cnt3, param_grads=bn3.grad_ptrs())
        return g, dc3

    def bwd(self, dout, saved, need_dx=True):
        x, a1, s1, a2, s2, a2b, sb, s3, se, mask, xs, sd, out = saved
        if isinstance(s3, tuple) and len(s3) == 2 and s3[0] == "fused":
            g, dc3 = self._bwd_fused_tail(dout, out, s3[1])
        else:
            g = ops.act_bwd(_as_act(dout), out, self.act, self.bn1.slope)
            d = g
            if mask is not None and not (se is not None and len(se) > 3):   # else folded into the ECA gate
                d = ops.scale_nc(d, mask)
            if self.eca is not None:
                d = self.eca.bwd(d, se)
            dc3 = self.bn3.bwd(d, s3)
        self.conv3.run_wgrad(a2b, dc3)
        # BN-backward reduction + leaky mask fused into the dgrad epilogue that produces the gradient
        # (same machinery as modules.Bottleneck; not across the blur-pool, not where it loses: the
        # halo-reuse 3x3 kernel)
        if FUSE_BN_BWD and self.blur is None:
            d, sums = self.conv3.run_dgrad(dc3, tuple(a2b.shape), bn_bwd=self.bn2.fuse_info(s2))
            dc2 = self.bn2.bwd(d, s2, sums=sums)
        else:
            d = self.conv3.run_dgrad(dc3, tuple(a2b.shape))
            if self.blur is not None:
                d = self.blur.bwd(d, sb)
            dc2 = self.bn2.bwd(d, s2)
        self.conv2.run_wgrad(a1, dc2)
        if FUSE_BN_BWD and not ops.halo_applies(a1.shape[1], dc2.shape[1], 3, self.conv2.stride, a1.shape[3]):
            d, sums = self.conv2.run_dgrad(dc2, tuple(a1.shape), bn_bwd=self.bn1.fuse_info(s1))
            dc1 = self.bn1.bwd(d, s1, sums=sums)
        else:
            d = self.conv2.run_dgrad(dc2, tuple(a1.shape))
            dc1 = self.bn1.bwd(d, s1)
        self.conv1.run_wgrad(x, dc1)
        if self.downsample is None:
            return self.conv1.run_dgrad(dc1, tuple(x.shape), residual=g) if need_dx else None
        dcd = self.downsample[1].bwd(g, sd)
        self.downsample[0].run_wgrad(xs, dcd)
        if not need_dx:
            return None
        dx = self.conv1.run_dgrad(dc1, tuple(x.shape))
        if self.pool_shortcut:
            dxs = self.downsample[0].run_dgrad(dcd, tuple(xs.shape))
            return ops.add_act(dx, ops.avgpool2_bwd(dxs, tuple(x.shape)), ops.ACT_NONE)
        if self.downsample[0].stride == 1:
            return self.downsample[0].run_dgrad(dcd, tuple(x.shape), out=dx, residual=dx)
        return self.downsample[0].run_dgrad(dcd, tuple(x.shape), out=dx)


class BResNet(SibModule):
    def __init__(self, layers=(3, 4, 6, 3), num_classes=1000, stem_type="deep", antialias=True, attn_type="eca",
                 norm_layer="inplaceabn", norm_act="leaky_relu", drop_rate=0.2, drop_connect_rate=0.2,
                 weight_standardization=False, **unused):
        super().__init__()
        if stem_type != "deep":
            raise _lib.SibError("BResNet: stem_type must be 'deep' (space2depth is not built)")
        if norm_layer not in ("abn", "inplaceabn"):
            raise _lib.SibError("BResNet: norm_layer must be abn / inplaceabn")
        self.drop_rate, self.antialias, self.norm_act = drop_rate, antialias, norm_act
        self.weight_standardization = weight_standardization
        # deep stem: conv3x3/2(3->32)-BN-act, conv3x3(32->32)-BN-act, conv3x3(32->64), then bn1-act
        self.conv1 = nn.Sequential(
            DeepStemFirst(32, 64), PaddedBatchNorm2d(32, 64, activation=norm_act), nn.Identity(),
            PaddedConv2d(32, 32, phys_in=64, phys_out=64), PaddedBatchNorm2d(32, 64, activation=norm_act), nn.Identity(),
            PaddedConv2d(32, 64, phys_in=64, phys_out=64))
        self.bn1 = BNAct(64, activation=norm_act)
        self.blurpool = BlurPool(64) if antialias else None
        inplanes, nblocks, bi = 64, sum(layers), 0
        for i, (planes, n) in enumerate(zip((64, 128, 256, 512), layers)):
            blocks = []
            for j in range(n):
                keep = 1.0 - drop_connect_rate * bi / nblocks
                blocks.append(BBottleneck(inplanes, planes, stride=(1 if i == 0 or j > 0 else 2), downsample=(j == 0),
                                          norm_act=norm_act, antialias=antialias, attn=(attn_type == "eca"),
                                          keep_prob=keep))
                inplanes = planes * BBottleneck.expansion
                bi += 1
            setattr(self, "layer%d" % (i + 1), nn.Sequential(*blocks))
        self.fc = Linear(inplanes, (num_classes + 7) // 8 * 8)
        self._out_features = num_classes

    def _prepare_input(self, x):
        return x

    def blocks(self):
        for i in range(1, 5):
            for blk in getattr(self, "layer%d" % i):
                yield blk

    def enable_weight_standardization(self, eps=1e-7):
        self.weight_standardization = True
        self._ws_eps = eps
        for m in self.modules():
            if isinstance(m, (PaddedConv2d, DeepStemFirst)):
                m.ws, m.ws_eps = True, eps
        self._arena = None
        return self

    def ensure_arena(self):
        fresh = self._arena is None or not self._arena.intact()
        a = super().ensure_arena()
        if fresh and self.weight_standardization:
            convs = [m.weight for m in self.modules() if isinstance(m, Conv2d)]
            a.enable_weight_standardization(convs, getattr(self, "_ws_eps", 1e-7))
            for m in self.modules():
                if isinstance(m, (PaddedConv2d, DeepStemFirst)):
                    m.ws = True
        return a

    def _end_backward(self):
        if self.weight_standardization:
            self._arena.standardize_grads()
        super()._end_backward()

    def fwd(self, x, train):
        c = self.conv1
        y, xq = c[0].run(x)
        y, sb0 = c[1].fwd(y, train)
        y2 = c[3].run(y)
        a, sb1 = c[4].fwd(y2, train)
        stats = ops.new_acc(2, 64, x.device) if train else None
        y3 = c[6].run(a, stats)
        a0, sbn = self.bn1.fwd(y3, train, stats=stats)
        if self.antialias:
            p, pidx = ops.maxpool3x3s1_fwd(a0)
            p2, sblur = self.blurpool.fwd(p, train)
        else:
            p2, pidx = ops.maxpool3x3s2_fwd(a0)
            sblur = None
        h = p2
        saved = []
        for blk in self.blocks():
            h, s = blk.fwd(h, train)
            saved.append(s)
        feat = ops.gap_fwd(h)
        dmask = None
        if train and self.drop_rate > 0:
            dmask = ((torch.rand(feat.shape, device=feat.device) >= self.drop_rate).to(torch.bfloat16) / (1 - self.drop_rate))
            feat = feat * dmask
        logits, fc_saved = self.fc.fwd(feat, train)
        logits = logits[:, :self._out_features] if logits.shape[1] != self._out_features else logits
        if not train:
            return logits, None
        return logits, (xq, sb0, y, sb1, a, sbn, tuple(a0.shape), pidx, sblur, saved, tuple(h.shape), dmask, fc_saved)

    def bwd(self, dlogits, saved_all, need_dx=False):
        xq, sb0, y, sb1, a, sbn, a0_shape, pidx, sblur, saved, h_shape, dmask, fc_saved = saved_all
        n = dlogits.shape[0]
        if dlogits.shape[1] != self.fc.out_features:
            full = torch.zeros((n, self.fc.out_features), dtype=torch.bfloat16, device=dlogits.device)
            full[:, :dlogits.shape[1]] = dlogits
            dlogits = full
        dfeat = self.fc.bwd(dlogits.to(torch.bfloat16).contiguous(), fc_saved)
        if dmask is not None:
            dfeat = dfeat * dmask
        d = ops.gap_bwd(_as_act(dfeat), h_shape)
        blocks = list(self.blocks())
        for i in range(len(blocks) - 1, -1, -1):
            d = blocks[i].bwd(d, saved[i], need_dx=True)
            saved[i] = None
            self._after_block_backward(i)
        if self.antialias:
            d = self.blurpool.bwd(d, sblur)
            d = ops.maxpool3x3s1_bwd(d, pidx)
        else:
            d = ops.maxpool3x3s2_bwd(d, pidx, a0_shape)
        c = self.conv1
        dy3 = self.bn1.bwd(d, sbn)
        c[6].run_wgrad(a, dy3)
        d = c[6].run_dgrad(dy3, tuple(a.shape))
        d = c[4].bwd(d, sb1)
        c[3].run_wgrad(y, d)
        d = c[3].run_dgrad(d, tuple(y.shape))
        d = c[1].bwd(d, sb0)
        c[0].run_wgrad(xq, d)
        return None

    def _after_block_backward(self, block_index):
        cb = getattr(self, "_block_bwd_cb", None)
        if cb is not None:
            ops.side_join()      # the block's weight gradients (side stream) must be complete
            cb(block_index)


def bresnet50(num_classes=1000, **kwargs):
    """BResNet-50 encoder config (BResNet50_encoder.yaml:44-51)."""
    return BResNet((3, 4, 6, 3), num_classes=num_classes, **kwargs)
