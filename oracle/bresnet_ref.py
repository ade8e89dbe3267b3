"""fp32 PyTorch restatement of BResNet-50 (oracle; test infrastructure only).

The reference builds this model through the absent, unpinned pytorch_tools package
(`pytorch_tools.models.resnet50(**model_params)`, kwargs at configs/_old_configs/_first_attempts/
BResNet50_encoder.yaml:44-51, weight standardisation train.py:66-67), so its arithmetic is NOT pinned
by anything in /root/reference: this file restates SURVEY.md App. C.2 switch by switch and is the twin
of sota_imagenet_b200/bresnet.py (same module names => interchangeable state_dicts).  Parity unpinned.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class WSConv2d(nn.Conv2d):
    """conv_to_ws_conv: per-output-channel (w - mean) / sqrt(var_biased + eps)."""
    ws = False
    eps = 1e-7

    def forward(self, x):
        w = self.weight
        if self.ws:
            var, mean = torch.var_mean(w, dim=(1, 2, 3), keepdim=True, unbiased=False)
            w = (w - mean) * torch.rsqrt(var + self.eps)
        return F.conv2d(x, w, None, self.stride, self.padding)


class BlurPool(nn.Module):
    """[1,2,1] x [1,2,1] / 16 depthwise, stride 2, zero padding 1."""

    def forward(self, x):
        f = torch.tensor([1.0, 2.0, 1.0], device=x.device, dtype=x.dtype)
        k = (f[:, None] * f[None, :] / 16.0)[None, None].repeat(x.shape[1], 1, 1, 1)
        return F.conv2d(x, k, stride=2, padding=1, groups=x.shape[1])


class ECA(nn.Module):
    def __init__(self, channels=None):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(1, 1, 3).uniform_(-0.5, 0.5))

    def forward(self, x):
        p = x.mean(dim=(2, 3))                                   # [N, C]
        s = torch.sigmoid(F.conv1d(p[:, None, :], self.weight, padding=1))[:, 0]
        return x * s[:, :, None, None]


def _act(name):
    return nn.LeakyReLU(0.01) if name == "leaky_relu" else (nn.ReLU() if name == "relu" else nn.Identity())


class BNAct(nn.BatchNorm2d):
    def __init__(self, c, activation="leaky_relu"):
        super().__init__(c)
        self.act_fn = _act(activation)

    def forward(self, x):
        return self.act_fn(super().forward(x))


class BBottleneck(nn.Module):
    def __init__(self, inplanes, planes, stride=1, downsample=False, norm_act="leaky_relu", antialias=True,
                 attn=True, keep_prob=1.0):
        super().__init__()
        out = planes * 4
        cs = 1 if (antialias and stride > 1) else stride
        self.conv1 = WSConv2d(inplanes, planes, 1, bias=False)
        self.bn1 = BNAct(planes, norm_act)
        self.conv2 = WSConv2d(planes, planes, 3, cs, 1, bias=False)
        self.bn2 = BNAct(planes, norm_act)
        self.blur = BlurPool() if (antialias and stride > 1) else None
        self.conv3 = WSConv2d(planes, out, 1, bias=False)
        self.bn3 = BNAct(out, "identity")
        self.eca = ECA(out) if attn else None
        self.act_fn = _act(norm_act)
        self.keep_prob = keep_prob
        self.pool_shortcut = downsample and antialias and stride > 1
        if downsample:
            self.downsample = nn.Sequential(WSConv2d(inplanes, out, 1, cs, bias=False), BNAct(out, "identity"))
        else:
            self.downsample = None

    def forward(self, x):
        o = self.bn1(self.conv1(x))
        o = self.bn2(self.conv2(o))
        if self.blur is not None:
            o = self.blur(o)
        o = self.bn3(self.conv3(o))
        if self.eca is not None:
            o = self.eca(o)
        if self.training and self.keep_prob < 1.0:
            keep = (torch.rand(o.shape[0], 1, 1, 1, device=o.device) < self.keep_prob).float() / self.keep_prob
            o = o * keep
        r = x
        if self.downsample is not None:
            if self.pool_shortcut:
                r = F.avg_pool2d(r, 2, 2)
            r = self.downsample(r)
        return self.act_fn(o + r)


class BResNet(nn.Module):
    def __init__(self, layers=(3, 4, 6, 3), num_classes=1000, antialias=True, attn_type="eca", norm_act="leaky_relu",
                 drop_rate=0.2, drop_connect_rate=0.2, weight_standardization=False):
        super().__init__()
        self.conv1 = nn.Sequential(
            WSConv2d(3, 32, 3, 2, 1, bias=False), BNAct(32, norm_act), nn.Identity(),
            WSConv2d(32, 32, 3, 1, 1, bias=False), BNAct(32, norm_act), nn.Identity(),
            WSConv2d(32, 64, 3, 1, 1, bias=False))
        self.bn1 = BNAct(64, norm_act)
        self.antialias = antialias
        self.blurpool = BlurPool() if antialias else None
        inplanes, nblocks, bi = 64, sum(layers), 0
        for i, (planes, n) in enumerate(zip((64, 128, 256, 512), layers)):
            blocks = []
            for j in range(n):
                keep = 1.0 - drop_connect_rate * bi / nblocks
                blocks.append(BBottleneck(inplanes, planes, 1 if i == 0 or j > 0 else 2, j == 0, norm_act, antialias,
                                          attn_type == "eca", keep))
                inplanes = planes * 4
                bi += 1
            setattr(self, "layer%d" % (i + 1), nn.Sequential(*blocks))
        self.drop = nn.Dropout(drop_rate)
        self.fc = nn.Linear(inplanes, num_classes)
        if weight_standardization:
            for m in self.modules():
                if isinstance(m, WSConv2d):
                    m.ws = True

    def forward(self, x):
        x = self.bn1(self.conv1(x))
        x = F.max_pool2d(x, 3, 1 if self.antialias else 2, 1)
        if self.antialias:
            x = self.blurpool(x)
        for i in range(1, 5):
            x = getattr(self, "layer%d" % i)(x)
        x = x.mean(dim=(2, 3))
        return self.fc(self.drop(x))


def bresnet50(seed=0, **kw):
    torch.manual_seed(seed)
    return BResNet(**kw)


def bbottleneck_bf16_faithful(blk, x, round_bn3=False):
    """BBottleneck.forward with values rounded to bfloat16 exactly where the B200 pipeline stores
    bf16 tensors (conv outputs, bn1 / bn2 + activation outputs, blur / avg-pool outputs, the block
    output; the ECA gate and the pooled means stay fp32, the gate is applied inside the final add +
    activation pass; `round_bn3` = the operator-sequence tail that also stores bn3's output): the per-block strict oracle of tests/test_gpu_bresnet.py."""
    from .torch_ref import _RoundFwd, _q

    def conv(m, t):
        w = m.weight
        if m.ws:
            var, mean = torch.var_mean(w, dim=(1, 2, 3), keepdim=True, unbiased=False)
            w = (w - mean) * torch.rsqrt(var + m.eps)
        return _q(F.conv2d(t, _RoundFwd.apply(w), None, m.stride, m.padding))

    def bn(m, t):
        return _q(m(t))

    o = bn(blk.bn1, conv(blk.conv1, x))
    o = bn(blk.bn2, conv(blk.conv2, o))
    if blk.blur is not None:
        o = _q(blk.blur(o))
    # the fused block tail (bresnet.FUSE_BN3_TAIL, the default) never stores bn3's output: the gate
    # and the shortcut add are applied to fp32 values and the block output is rounded once
    o = blk.bn3(conv(blk.conv3, o))
    if round_bn3:
        o = _q(o)
    if blk.eca is not None:
        p = o.mean(dim=(2, 3))
        s = torch.sigmoid(F.conv1d(p[:, None, :], blk.eca.weight, padding=1))[:, 0]
        o = o * s[:, :, None, None]
    r = x
    if blk.downsample is not None:
        if blk.pool_shortcut:
            r = _q(F.avg_pool2d(r, 2, 2))
        r = bn(blk.downsample[1], conv(blk.downsample[0], r))
    return _q(blk.act_fn(o + r))
