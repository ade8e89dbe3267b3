"""fp32 PyTorch restatement of the training step (oracle; see oracle/__init__.py).

  * ResNet-50            = torchvision.models.resnet50(weights=None): BASELINE.json config #1 and
                           reference README.md:42 ("default torchvision version of Resnet50");
                           the reference builds it through pytorch_tools.models.resnet50
                           (train.py:64, configs/hydra_exp/1.r50_baseline.yaml:22-23) — absent.
  * smooth_cross_entropy = pytorch_tools.losses.smooth.CrossEntropyLoss restated from SURVEY.md
                           App. C.1 (call sites: arg_parser.py:140-142, angular_losses.py:572-576).
  * sgd                  = torch.optim.SGD(foreach=True) == torch.optim._multi_tensor.SGD
                           (arg_parser.py:136-138).
  * angular heads        = restated line by line from reference angular_losses.py and pinned to
                           the reference file itself by oracle/make_golden.py.
"""
import math

import torch
import torch.nn.functional as F


def smooth_cross_entropy(y_pred, y_true, smoothing=0.0, temperature=1.0):
    """App. C.1: one-hot via scatter for index targets, dense targets used as is."""
    logp = F.log_softmax(y_pred.float() / temperature, dim=1)
    if y_true.dim() == 1:
        onehot = torch.zeros_like(logp).scatter_(1, y_true[:, None].long(), 1.0)
    else:
        onehot = y_true.float()
    nll = -(logp * onehot).sum(1)
    smooth = -logp.mean(1)
    return ((1 - smoothing) * nll + smoothing * smooth).mean()


def sphere_linear(x, weight):
    """angular_losses.py:212-214"""
    return F.linear(F.normalize(x), F.normalize(weight))


def arcface_logits(cosine, y_true, s=10.0, m=0.2):
    """angular_losses.py:118-143 (AdditiveAngularMarginLoss up to final_criterion)."""
    cos_m, sin_m = math.cos(m), math.sin(m)
    th, mm = math.cos(math.pi - m), math.sin(math.pi - m) * m
    sine = torch.sqrt(1.0 - torch.pow(cosine, 2))
    phi = cosine * cos_m - sine * sin_m
    phi = torch.where(cosine > th, phi, cosine - mm)
    one_hot = torch.zeros_like(cosine).scatter_(1, y_true[..., None].long(), 1.0)
    return ((one_hot * phi) + ((1.0 - one_hot) * cosine)) * s


def cosface_logits(cosine, y_true, s=30.0, m=0.4):
    """angular_losses.py:189-196 (LargeMarginCosineLoss) == AdaCos fixed_s path :332-333."""
    one_hot = torch.zeros_like(cosine).scatter_(1, y_true.view(-1, 1).long(), 1.0)
    return ((one_hot * (cosine - m)) + ((1.0 - one_hot) * cosine)) * s


def arccos_logits(cosine, y_true=None, s=1.0, m=0.0):
    """reference angular_losses.py:572-576 (ArcCosSoftmax: s=1, m=0) and :323-330 (AdaCos
    arc_logits): -(acos(clamp(cos, -1+1e-7, 1-1e-7)) + m on columns with target mass) * s."""
    eps = 1e-7
    theta = torch.acos(cosine.clamp(-1 + eps, 1 - eps))
    if m:
        onehot = y_true if y_true.dim() == 2 else torch.zeros_like(cosine).scatter_(1, y_true[:, None], 1.0)
        theta = theta.where(onehot.eq(0), theta + m)
    return theta.neg() * s


def novograd_step(params, grads, state, lr, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-2,
                  ema_norm_init=1e-3, unitwise=False):
    """One step of the reference's MyNovograd written out (sota_imagenet/optimizers.py:85-160).
    `state` is a list of dicts (ema_grad tensor, ema_norm tensor of the norm-group shape)."""
    b1, b2 = betas
    for p, g, st in zip(params, grads, state):
        if not st:
            st["ema_grad"] = torch.zeros_like(p)
            st["ema_norm"] = None
        if unitwise:                                             # :18-22, :133-134
            nrm = p.norm(2) if p.ndim <= 1 else p.norm(2, dim=tuple(range(1, p.ndim)), keepdim=True)
        else:
            nrm = p.pow(2).sum()                                 # :136 (weights, not gradients)
        if st["ema_norm"] is None:
            st["ema_norm"] = torch.full_like(nrm, ema_norm_init)
        st["ema_norm"] = st["ema_norm"] * b2 + (1 - b2) * nrm    # :139-140
        st["ema_grad"] = st["ema_grad"] * b1 + (1 - b1) * g      # :143,147
        denom = st["ema_norm"].sqrt() + eps                      # :144-145
        p.copy_((p - lr * st["ema_grad"] / denom) * (1 - lr * weight_decay))   # :155,158


def resnet50(num_classes=1000, seed=0):
    import torchvision
    torch.manual_seed(seed)
    return torchvision.models.resnet50(weights=None, num_classes=num_classes)


def make_sgd(params, lr, momentum=0.9, weight_decay=3e-5, nesterov=False):
    return torch.optim.SGD(params, lr=lr, momentum=momentum, weight_decay=weight_decay,
                           nesterov=nesterov, foreach=True)


def train_step(model, opt, x, y, smoothing=0.1):
    """One fwd + bwd + SGD step; returns the loss (float)."""
    opt.zero_grad(set_to_none=True)
    loss = smooth_cross_entropy(model(x), y, smoothing)
    loss.backward()
    opt.step()
    return float(loss.detach())


def synthetic_batch(batch, size, num_classes=1000, seed=0):
    """x ~ randn (DALI output range), y ~ randint."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, 3, size, size, generator=g)
    y = torch.randint(0, num_classes, (batch,), generator=g)
    return x, y


# ---------------------------------------------------------------------------------------------
# bf16-faithful variant: the same fp32 arithmetic with values rounded to bfloat16 at exactly the
# points where the B200 pipeline stores bf16 tensors (conv outputs, activation outputs, their
# gradients, the packed filters).  At random init a 50-layer ReLU network is chaotic: even stock
# torch.autocast(bfloat16) reaches only ~0.0-0.1 per-parameter gradient cosine against pure fp32
# at batch 16 (see DESIGN.md "Parity"), so whole-network gradient parity is asserted against
# this oracle, and kernel-level parity against pure fp32 per layer.
# ---------------------------------------------------------------------------------------------
class _RoundBoth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


class _RoundFwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


def _q(x):
    return _RoundBoth.apply(x)


def _conv_q(m, x):
    return _q(F.conv2d(x, _RoundFwd.apply(m.weight), None, m.stride, m.padding))


def _bn(m, x):
    if m.training:
        m.num_batches_tracked += 1
    return F.batch_norm(x, m.running_mean, m.running_var, m.weight, m.bias, m.training,
                        m.momentum, m.eps)


def bf16_faithful_forward(model, x, act=F.relu):
    """torchvision ResNet forward with bf16 storage rounding (see comment above); `act` replaces
    every ReLU (the leaky-ReLU regime of tests/test_gpu_model.py)."""
    x = _RoundFwd.apply(x)
    y = _q(act(_bn(model.bn1, _conv_q(model.conv1, x))))
    y = F.max_pool2d(y, 3, 2, 1)
    for layer in (model.layer1, model.layer2, model.layer3, model.layer4):
        for blk in layer:
            idt = y
            o = _q(act(_bn(blk.bn1, _conv_q(blk.conv1, y))))
            o = _q(act(_bn(blk.bn2, _conv_q(blk.conv2, o))))
            o = _bn(blk.bn3, _conv_q(blk.conv3, o))
            if blk.downsample is not None:
                idt = _bn(blk.downsample[1], _conv_q(blk.downsample[0], y))
            y = _q(act(o + idt))
    feat = _q(y.mean(dim=(2, 3)))
    return _q(F.linear(feat, _RoundFwd.apply(model.fc.weight), model.fc.bias))
