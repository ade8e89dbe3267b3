"""Model constructors with the surface of `pytorch_tools.models.resnet50`
(reference train.py:64 `hydra.utils.call(cfg.model)`, configs/hydra_exp/1.r50_baseline.yaml:22-23).

ResNet-50 == torchvision ResNet-50 v1.5 module graph and state_dict keys; the forward and
backward passes are kernel sequences from libsib200 (see modules.py).
"""
import os

import torch
import torch.nn as nn

from . import _lib, ops
from .modules import BatchNorm2d, Bottleneck, Linear, MaxPool3x3s2, SibModule, StemConv


FUSE_STEM_POOL = os.environ.get("SIB_FUSE_STEM_POOL", "1") != "0"


class ResNet(SibModule):
    def __init__(self, layers=(3, 4, 6, 3), num_classes=1000, in_channels=3, norm_act="relu",
                 drop_rate=0.0, zero_init_residual=False, embedding_size=None):
        super().__init__()
        assert in_channels == 3
        self.num_classes = num_classes
        self.drop_rate = drop_rate
        self.conv1 = StemConv(64, 7, 3)
        self.bn1 = BatchNorm2d(64, activation=norm_act)
        self.maxpool = MaxPool3x3s2()
        inplanes = 64
        for i, (planes, n) in enumerate(zip((64, 128, 256, 512), layers)):
            stride = 1 if i == 0 else 2
            blocks = [Bottleneck(inplanes, planes, stride, downsample=True, activation=norm_act)]
            inplanes = planes * Bottleneck.expansion
            blocks += [Bottleneck(inplanes, planes, activation=norm_act) for _ in range(1, n)]
            setattr(self, "layer%d" % (i + 1), nn.Sequential(*blocks))
        out_features = embedding_size if embedding_size else num_classes
        # pad the class dimension to a multiple of 8 for 16-byte rows; extra logits are sliced off
        self._out_features = out_features
        self.fc = Linear(inplanes, (out_features + 7) // 8 * 8)
        if zero_init_residual:
            for m in self.modules():
                if isinstance(m, Bottleneck):
                    nn.init.zeros_(m.bn3.weight)

    # SibModule.forward() normalises 4-D inputs to bf16 NHWC; the stem takes the image itself
    def _prepare_input(self, x):
        return x

    def blocks(self):
        for i in range(1, 5):
            for blk in getattr(self, "layer%d" % i):
                yield blk

    def fwd(self, x, train):
        saved = []
        stats = ops.new_acc(2, 64, x.device) if train and not ops.DETERMINISTIC else None
        c0, xq = self.conv1.run(x, stats)
        if train and ops.DETERMINISTIC:
            stats = ops.bn_stats(c0)
        n, _, h, w = c0.shape
        if train and FUSE_STEM_POOL:
            # bn1 + act + max pool in one pass: the 112x112 normalised activation is never written
            # (its backward mask is recomputed from c0, the pooling backward needs only idx)
            bn = self.bn1
            a1, world = bn.stats_args(stats)
            cnt0 = n * h * w * world
            p0, idx, mi0, ss0 = ops.bn_act_maxpool3x3s2_fwd(c0, a1, bn.act, bn.slope, cnt0, bn.eps,
                                                            bn.momentum)
            pool_saved = (idx, tuple(c0.shape))
        else:
            a0, mi0, ss0, cnt0, _ = Bottleneck._bn_act(self.bn1, c0, stats, train)
            p0, pool_saved = self.maxpool.fwd(a0, train)
        y = p0
        for blk in self.blocks():
            y, s = blk.fwd(y, train)
            saved.append(s)
        feat = ops.gap_fwd(y)
        logits, fc_saved = self.fc.fwd(feat, train)
        if logits.shape[1] != self._out_features:
            logits = logits[:, :self._out_features]
        if not train:
            return logits, None
        return logits, (xq, c0, mi0, cnt0, ss0, pool_saved, saved, tuple(y.shape), fc_saved)

    def bwd(self, dlogits, saved_all, need_dx=False):
        xq, c0, mi0, cnt0, ss0, pool_saved, saved, y_shape, fc_saved = saved_all
        n = dlogits.shape[0]
        padded = self.fc.out_features
        if dlogits.shape[1] != padded:
            full = torch.zeros((n, padded), dtype=torch.bfloat16, device=dlogits.device)
            full[:, :dlogits.shape[1]] = dlogits
            dlogits = full
        dfeat = self.fc.bwd(dlogits.to(torch.bfloat16).contiguous(), fc_saved)
        dy = ops.gap_bwd(dfeat, y_shape)
        blocks = list(self.blocks())
        dy_sums = None
        for i in range(len(blocks) - 1, -1, -1):
            # fuse the previous block's bn3 + add + act backward reduction into this block's last
            # dgrad epilogue when both sides allow it (modules.Bottleneck.fuse_info)
            prev = blocks[i - 1].fuse_info(saved[i - 1]) if i > 0 and blocks[i].can_fuse_prev() else None
            r = blocks[i].bwd(dy, saved[i], need_dx=True, dout_sums=dy_sums, prev=prev)
            dy, dy_sums = r if prev is not None else (r, None)
            saved[i] = None
            self._after_block_backward(i)
        da0 = self.maxpool.bwd(dy, pool_saved)
        bn1 = self.bn1
        sums = bn1.reduce_sums(ops.bn_bwd_reduce(da0, None, c0, mi0, bn1.act, bn1.slope, mask_ss=ss0))
        dc0, _, _ = ops.bn_bwd_apply(da0, None, c0, mi0, bn1.weight.data, sums, cnt0, bn1.act,
                                     bn1.slope, mask_ss=ss0, param_grads=bn1.grad_ptrs())
        self.conv1.run_wgrad(xq, dc0)
        return None

    def _after_block_backward(self, block_index):
        """Hook point for the data-parallel wrapper (bucketed gradient all-reduce)."""
        cb = getattr(self, "_block_bwd_cb", None)
        if cb is not None:
            ops.side_join()      # the block's weight gradients (side stream) must be complete
            cb(block_index)

    # checkpoints written by torchvision have fc.weight [1000, 2048]; ours may be padded
    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        for key in ("fc.weight", "fc.bias"):
            k = prefix + key
            if k in state_dict:
                t = state_dict[k]
                pad = self.fc.out_features - t.shape[0]
                if pad > 0:
                    state_dict[k] = torch.cat([t, t.new_zeros((pad,) + tuple(t.shape[1:]))], 0)
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


_BRESNET_KEYS = ("stem_type", "antialias", "attn_type", "norm_layer", "drop_connect_rate", "deep_stem")


def resnet50(num_classes=1000, pretrained=None, **kwargs):
    """`pytorch_tools.models.resnet50` / `torchvision.models.resnet50` replacement.  The BResNet
    switches of the reference's configs (stem_type, antialias, attn_type, norm_layer, norm_act,
    drop_rate, drop_connect_rate) select the BResNet-50 graph (bresnet.py)."""
    if pretrained:
        raise _lib.SibError("no pretrained weights are available offline")
    if any(k in kwargs for k in _BRESNET_KEYS):
        from .bresnet import bresnet50
        return bresnet50(num_classes=num_classes, **kwargs)
    return ResNet((3, 4, 6, 3), num_classes=num_classes, **kwargs)


def resnet26(num_classes=1000, **kwargs):
    """Small member of the family (bottleneck [2,2,2,2]); used by the tests."""
    return ResNet((2, 2, 2, 2), num_classes=num_classes, **kwargs)


def resnet101(num_classes=1000, **kwargs):
    return ResNet((3, 4, 23, 3), num_classes=num_classes, **kwargs)


class ResNetEmbedding(torch.nn.Module):
    """ResNet-50 trunk -> `embedding_size`-d embedding -> SphereLinearLayer(embedding, classes):
    the model side of the angular-margin configs (reference angular_losses.py:202-214 layer in
    front of AdditiveAngularMarginLoss / AdaCos criteria).  Returns cosine logits [N, classes]."""

    def __init__(self, embedding_size=512, num_classes=1000, **kwargs):
        super().__init__()
        from .losses import SphereLinearLayer
        self.encoder = ResNet((3, 4, 6, 3), num_classes=num_classes, embedding_size=embedding_size, **kwargs)
        self.head = SphereLinearLayer(embedding_size, num_classes)

    def forward(self, x):
        return self.head(self.encoder(x))


def resnet50_embedding(embedding_size=512, num_classes=1000, **kwargs):
    return ResNetEmbedding(embedding_size, num_classes, **kwargs)
