"""Criteria with the reference's constructor / call signatures, run by the fused head kernels.

  CrossEntropyLoss           pytorch_tools.losses.smooth.CrossEntropyLoss (reference
                             arg_parser.py:140-142; smoothing 0.1 in 1.r50_baseline.yaml:34-35)
  SphereLinearLayer          reference angular_losses.py:202-214
  AdditiveAngularMarginLoss  reference angular_losses.py:98-146 (ArcFace, easy-margin fallback)
  LargeMarginCosineLoss      reference angular_losses.py:149-199 (CosFace, owns W)
  AdaCos                     reference angular_losses.py:248-334 (fixed_s = CosFace on cosines;
                             adaptive scale statistics kept with the reference's update rule)
  AngularPenaltySMLoss       reference angular_losses.py:13-95 (arcface / cosface variants)
  ArcCosSoftmax              reference angular_losses.py:572-576 (CE over -acos(cos))
"""
import torch
import torch.nn as nn

from . import _lib, ops


class Loss(nn.Module):
    """Stand-in for pytorch_tools.losses.Loss (base class used by the reference's criteria)."""


class _FusedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, smoothing, temperature, margin_kind, s, m):
        want_grad = ctx.needs_input_grad[0]
        loss, rows, dlogits = ops.ce_fwd_bwd(logits, target, smoothing, temperature, margin_kind, s,
                                             m, want_grad=want_grad)
        ctx.dlogits = dlogits
        return loss

    @staticmethod
    def backward(ctx, g):
        d = ctx.dlogits
        ctx.dlogits = None
        # d already holds d(mean loss)/d(logits); scale by the incoming scalar gradient
        d = d * g.to(d.dtype)
        return d, None, None, None, None, None, None


def _fused_ce(logits, target, smoothing=0.0, temperature=1.0, margin_kind=ops.MARGIN_NONE, s=1.0, m=0.0):
    if logits.dtype not in (torch.float32, torch.bfloat16):
        logits = logits.float()
    if logits.stride(-1) != 1:
        logits = logits.contiguous()
    return _FusedCE.apply(logits, target, float(smoothing), float(temperature), margin_kind,
                          float(s), float(m))


class CrossEntropyLoss(Loss):
    """CE over index or dense (one-hot / soft) targets with label smoothing and temperature:
    loss = mean_i[(1-s) * -sum_c t_ic log p_ic + s * -mean_c log p_ic], p = softmax(x / T)."""

    def __init__(self, mode="multiclass", smoothing=0.0, weight=1.0, reduction="mean",
                 temperature=1.0, normalize=False):
        super().__init__()
        if mode != "multiclass" or reduction != "mean" or normalize:
            raise _lib.SibError("fused CrossEntropyLoss supports mode='multiclass', reduction='mean', "
                                "normalize=False")
        self.smoothing, self.temperature, self.loss_weight = smoothing, temperature, weight

    def forward(self, y_pred, y_true):
        loss = _fused_ce(y_pred, y_true, self.smoothing, self.temperature)
        return loss * self.loss_weight if self.loss_weight != 1.0 else loss


class _SphereLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, normalize_x):
        cosv, saved = ops.sphere_linear_fwd(x, w, normalize_x)
        ctx.saved = saved
        ctx.normalize_x = normalize_x
        return cosv

    @staticmethod
    def backward(ctx, dcos):
        dx, dw = ops.sphere_linear_bwd(dcos.float().contiguous(), ctx.saved,
                                       need_dx=ctx.needs_input_grad[0],
                                       need_dw=ctx.needs_input_grad[1],
                                       normalize_x=ctx.normalize_x)
        ctx.saved = None
        return dx, dw, None


def sphere_linear(x, w, normalize_x=True):
    in_dtype = x.dtype
    out = _SphereLinear.apply(x.float().contiguous(), w.float().contiguous(), normalize_x)
    return out if in_dtype == torch.float32 else out


class SphereLinearLayer(nn.Module):
    """cos = normalize(x) . normalize(W)^T"""

    def __init__(self, embedding_size, num_classes):
        super().__init__()
        self.register_parameter("weight", nn.Parameter(torch.zeros(num_classes, embedding_size)))
        nn.init.xavier_uniform_(self.weight)

    def forward(self, x):
        return sphere_linear(x, self.weight)


class SphereMLPLayer(nn.Module):
    """Reference angular_losses.py:217-245: in training mode (or with `val_projector`) the
    embedding first goes through a projector FC(no bias) - BatchNorm1d - act - FC, then
    cos = normalize(.) . normalize(W)^T; in validation only the cosine layer.  Same parameter names
    (`weight`, `projector.{0,1,3}.*`), so checkpoints interchange.  The two FCs run on the tcgen05
    1x1-conv kernel, BatchNorm + ReLU on the fused NHWC kernels (statistics over the batch dimension
    = BatchNorm1d), the cosine layer on the sphere-linear kernels; `hswish` runs the activation as
    a torch op between two fused modules."""

    def __init__(self, embedding_size, num_classes, hidden_size=4096, act="relu", val_projector=False):
        super().__init__()
        from .modules import BatchNorm2d, Linear
        if act not in ("relu", "hswish"):
            raise _lib.SibError("SphereMLPLayer: act must be 'relu' or 'hswish'")
        self.register_parameter("weight", nn.Parameter(torch.zeros(num_classes, embedding_size)))
        nn.init.xavier_uniform_(self.weight)
        self.projector = nn.Sequential(
            Linear(embedding_size, hidden_size, bias=False),
            BatchNorm2d(hidden_size, activation="relu" if act == "relu" else "identity"),
            nn.Identity() if act == "relu" else nn.Hardswish(),
            Linear(hidden_size, embedding_size),
        )
        self.val_projector = val_projector

    def forward(self, x):
        if self.training or self.val_projector:
            fc1, bn, act, fc2 = self.projector
            h = fc1(x.to(torch.bfloat16).reshape(x.shape[0], x.shape[1], 1, 1))
            h = bn(h.reshape(h.shape[0], h.shape[1], 1, 1))
            if not isinstance(act, nn.Identity):
                h = act(h.float()).to(torch.bfloat16)
            x = fc2(h.reshape(h.shape[0], h.shape[1], 1, 1))
        return sphere_linear(x, self.weight)


def _smoothing_of(criterion):
    """Extract (smoothing, temperature) from a `final_criterion` the reference would pass."""
    if criterion is None:
        return 0.0, 1.0
    if isinstance(criterion, CrossEntropyLoss):
        return criterion.smoothing, criterion.temperature
    if isinstance(criterion, nn.CrossEntropyLoss):
        return criterion.label_smoothing, 1.0
    raise _lib.SibError("final_criterion must be a (label-smoothing) cross entropy for the fused head")


class AdditiveAngularMarginLoss(nn.Module):
    """ArcFace on cosine logits; needs index labels (the reference scatters them, :140)."""

    def __init__(self, final_criterion=None, s=10.0, m=0.2):
        super().__init__()
        self.s, self.m = s, m
        self.final_criterion = final_criterion if final_criterion is not None else nn.CrossEntropyLoss()
        self.smoothing, self.temperature = _smoothing_of(self.final_criterion)

    def forward(self, cosine, y_true):
        if y_true.dim() != 1:
            raise _lib.SibError("AdditiveAngularMarginLoss expects class indices, not one-hot targets")
        return _fused_ce(cosine.float(), y_true, self.smoothing, self.temperature, ops.MARGIN_ARC,
                         self.s, self.m)


class LargeMarginCosineLoss(nn.Module):
    """CosFace; owns W and normalises only W (features are expected L2-normalised, :183-187)."""

    def __init__(self, in_features, out_features, s=30.0, m=0.40, criterion="cross_entropy"):
        super().__init__()
        if criterion != "cross_entropy":
            raise _lib.SibError("LargeMarginCosineLoss: only criterion='cross_entropy' is fused")
        self.in_features, self.out_features, self.s, self.m = in_features, out_features, s, m
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        nn.init.xavier_uniform_(self.weight)

    def forward(self, features, y_true):
        cosine = sphere_linear(features, self.weight, normalize_x=False)
        return _fused_ce(cosine, y_true.view(-1), 0.0, 1.0, ops.MARGIN_COS, self.s, self.m)


class AngularPenaltySMLoss(nn.Module):
    """arcface / cosface over owned W.  The reference's exclude-target-from-denominator form
    (:92-94) is algebraically softmax-CE over the margin-modified logits."""
    _default_values = {"arcface": (64.0, 0.5), "sphereface": (64.0, 1.35), "cosface": (30.0, 0.4)}

    def __init__(self, in_features=512, out_features=3088, loss_type="arcface", s=None, m=None,
                 criterion=None):
        super().__init__()
        if loss_type not in ("arcface", "cosface"):
            raise _lib.SibError("AngularPenaltySMLoss: fused types are 'arcface' and 'cosface'")
        self.s, self.m = self._default_values[loss_type]
        self.s = self.s if not s else s
        self.m = self.m if not m else m
        self.loss_type = loss_type
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        nn.init.xavier_uniform_(self.weight)

    def forward(self, features, y_true):
        cosine = sphere_linear(features, self.weight)
        if self.loss_type == "cosface":
            return _fused_ce(cosine, y_true, 0.0, 1.0, ops.MARGIN_COS, self.s, self.m)
        # cos(acos(clamp(c)) + m) on the target column, no easy-margin fallback (reference :78-83)
        eps = 1e-7
        return _fused_ce(cosine.clamp(-1 + eps, 1 - eps), y_true, 0.0, 1.0, ops.MARGIN_ARC_PURE,
                         self.s, self.m)


class AdaCos(nn.Module):
    """AdaCos on cosine logits.  With `fixed_s` this is CosFace (margin on the target column,
    constant scale); `arc_logits` feeds -(theta + margin) * s instead.  The adaptive scale follows reference :301-314 (running B, running median
    cosine, s = log B / (max(med, 0.7) - margin) capped at max_s); those no-grad statistics are
    a handful of tiny tensor ops, the margin + scale + CE forward/backward is the fused kernel."""

    def __init__(self, final_criterion=None, margin=0, max_s=20, fixed_s=None, momentum=0.95,
                 arc_logits=False, arc_margin=False):
        super().__init__()
        assert (not arc_logits) or arc_margin, "arc_logits=True and arc_margin=False are not supported!"
        self.arc_logits, self.arc_margin = arc_logits, arc_margin
        self.final_criterion = final_criterion
        self.smoothing, self.temperature = _smoothing_of(final_criterion)
        self.margin, self.momentum, self.max_s, self.fixed_s = margin, momentum, max_s, fixed_s
        self.prev_s = max_s
        self.running_B = 1000
        self.running_cos = 0.7
        self.idx = 0

    def forward(self, cosine, y_true):
        cosine = cosine.float()
        with torch.no_grad():
            if y_true.dim() == 1:
                idx = y_true.long()
                onehot_zero = torch.ones_like(cosine, dtype=torch.bool)
                onehot_zero.scatter_(1, idx[:, None], False)
            else:
                idx = y_true.argmax(-1).long()
                onehot_zero = y_true.eq(0)
            if self.fixed_s is None:
                b_batch = cosine[onehot_zero].mul(self.prev_s).exp().sum().div(cosine.size(0))
                med_cos = cosine.gather(1, idx[:, None]).median()
                self.running_B = self.running_B * self.momentum + b_batch * (1 - self.momentum)
                self.running_cos = self.running_cos * self.momentum + med_cos * (1 - self.momentum)
                s = self.running_B.log() / (self.running_cos.clamp_min(0.7) - self.margin)
                self.prev_s = min(float(s), self.max_s)
        self.idx += 1
        scale = self.fixed_s if self.fixed_s is not None else self.prev_s
        # arc_logits (:323-327): logits = -(acos(clamp(cos)) + margin[target]) * s; otherwise the
        # margin is subtracted from the target cosine (:329) whatever `arc_margin` says
        kind = ops.MARGIN_ARCCOS if self.arc_logits else ops.MARGIN_COS
        return _fused_ce(cosine, y_true, self.smoothing, self.temperature, kind, scale, self.margin)


class ArcCosSoftmax(CrossEntropyLoss):
    """Smooth CE over -acos(clamp(cos, -1+1e-7, 1-1e-7)) (reference angular_losses.py:572-576)."""

    def forward(self, y_pred, y_true):
        loss = _fused_ce(y_pred.float(), y_true, self.smoothing, self.temperature,
                         ops.MARGIN_ARCCOS, 1.0, 0.0)
        return loss * self.loss_weight if self.loss_weight != 1.0 else loss


LOSS_FROM_NAME = {"arcface": AdditiveAngularMarginLoss, "cross_entropy": CrossEntropyLoss}
