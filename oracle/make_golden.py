"""Generate tests/golden/*.pt by running the REFERENCE's own code (angular_losses.py,
torch.optim SGD, torch CrossEntropyLoss) and the oracle restatements on seeded inputs.
Run in the build container (needs /root/reference):  python -m oracle.make_golden
"""
import os

import numpy as np
import torch

from oracle import augment_ref, torch_ref
from oracle.reference_stub import load_reference_modules

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def heads():
    ang, _, RefCE = load_reference_modules()
    torch.manual_seed(0)
    b, d, c = 16, 64, 40
    x = torch.randn(b, d)
    w = torch.randn(c, d)
    y = torch.randint(0, c, (b,))
    out = {"x": x, "w": w, "y": y}
    # SphereLinearLayer + AdditiveAngularMarginLoss (ArcFace), reference classes
    layer = ang.SphereLinearLayer(d, c)
    layer.weight.data.copy_(w)
    for name, crit in (
        ("arc", ang.AdditiveAngularMarginLoss(final_criterion=RefCE(smoothing=0.1), s=10.0, m=0.2)),
        ("arc_s64", ang.AdditiveAngularMarginLoss(final_criterion=RefCE(), s=64.0, m=0.5)),
        ("adacos_fixed", ang.AdaCos(final_criterion=RefCE(smoothing=0.1), margin=0.2, fixed_s=10)),
    ):
        xr = x.clone().requires_grad_(True)
        layer.weight.grad = None
        cos = layer(xr)
        loss = crit(cos, y)
        loss.backward()
        out[name] = {"cos": cos.detach(), "loss": loss.detach(), "dx": xr.grad.clone(),
                     "dw": layer.weight.grad.clone()}
    # LargeMarginCosineLoss owns W, normalises only W
    lm = ang.LargeMarginCosineLoss(d, c, s=30.0, m=0.4)
    lm.weight.data.copy_(w)
    xn = torch.nn.functional.normalize(x).requires_grad_(True)
    loss = lm(xn, y)
    loss.backward()
    out["cosface_lm"] = {"xn": xn.detach().clone(), "loss": loss.detach(), "dx": xn.grad.clone(),
                         "dw": lm.weight.grad.clone()}
    # AngularPenaltySMLoss arcface / cosface
    for lt in ("arcface", "cosface"):
        ap = ang.AngularPenaltySMLoss(d, c, loss_type=lt)
        ap.weight.data.copy_(w)
        xr = x.clone().requires_grad_(True)
        loss = ap(xr, y)
        loss.backward()
        out["aps_" + lt] = {"loss": loss.detach(), "dx": xr.grad.clone(), "dw": ap.weight.grad.clone()}
    # adaptive AdaCos: three steps of the running statistics
    ada = ang.AdaCos(final_criterion=RefCE(smoothing=0.1), margin=0.1)
    trace = []
    for step in range(3):
        xr = (x * (1 + 0.1 * step)).requires_grad_(True)
        loss = ada(layer(xr), y)
        trace.append({"loss": loss.detach(), "s": float(ada.prev_s), "B": float(ada.running_B),
                      "cos": float(ada.running_cos)})
    out["adacos_adaptive"] = trace
    # restatements must agree with the reference classes
    cos = torch_ref.sphere_linear(x, w)
    l_arc = torch_ref.smooth_cross_entropy(torch_ref.arcface_logits(cos, y, 10.0, 0.2), y, 0.1)
    assert torch.allclose(l_arc, out["arc"]["loss"], atol=1e-6), (l_arc, out["arc"]["loss"])
    l_cos = torch_ref.smooth_cross_entropy(torch_ref.cosface_logits(cos, y, 10.0, 0.2), y, 0.1)
    assert torch.allclose(l_cos, out["adacos_fixed"]["loss"], atol=1e-6)
    torch.save(out, os.path.join(OUT, "heads.pt"))


def sphere_mlp():
    """Reference SphereMLPLayer (angular_losses.py:217-245): train-mode forward / backward and the
    validation forward of the class itself, bf16-representable inputs and FC weights."""
    ang, _, RefCE = load_reference_modules()
    torch.manual_seed(3)
    b, d, c, hid = 32, 64, 40, 128
    out = {}
    for act in ("relu", "hswish"):
        layer = ang.SphereMLPLayer(d, c, hidden_size=hid, act=act)
        with torch.no_grad():
            for p in layer.projector.parameters():
                if p.dim() == 2:
                    p.copy_(p.bfloat16().float())
            layer.projector[1].weight.uniform_(0.5, 1.5)
            layer.projector[1].bias.normal_(0, 0.2)
        x = torch.randn(b, d).bfloat16().float()
        g = torch.randn(b, c)
        sd = {k: v.clone() for k, v in layer.state_dict().items()}
        layer.train()
        xr = x.clone().requires_grad_(True)
        cos = layer(xr)
        cos.backward(g)
        rec = {"state_dict": sd, "x": x, "g": g, "cos_train": cos.detach(), "dx": xr.grad.clone(),
               "grads": {n: p.grad.clone() for n, p in layer.named_parameters()},
               "running_mean": layer.projector[1].running_mean.clone(),
               "running_var": layer.projector[1].running_var.clone()}
        layer.eval()
        rec["cos_eval"] = layer(x).detach()
        out[act] = rec
    torch.save(out, os.path.join(OUT, "sphere_mlp.pt"))


def cross_entropy():
    torch.manual_seed(1)
    b, c = 12, 50
    logits = torch.randn(b, c) * 3
    y = torch.randint(0, c, (b,))
    soft = torch.softmax(torch.randn(b, c), 1)
    out = {"logits": logits, "y": y, "soft": soft, "cases": []}
    for sm, temp in ((0.0, 1.0), (0.1, 1.0), (0.1, 0.15)):
        lr = logits.clone().requires_grad_(True)
        loss = torch_ref.smooth_cross_entropy(lr, y, sm, temp)
        loss.backward()
        if temp == 1.0:   # anchor the restatement on torch.nn.CrossEntropyLoss
            ref = torch.nn.functional.cross_entropy(logits, y, label_smoothing=sm)
            assert torch.allclose(loss, ref, atol=1e-6)
        ls = logits.clone().requires_grad_(True)
        loss_soft = torch_ref.smooth_cross_entropy(ls, soft, sm, temp)
        loss_soft.backward()
        out["cases"].append({"smoothing": sm, "temperature": temp, "loss": loss.detach(),
                             "grad": lr.grad.clone(), "loss_soft": loss_soft.detach(),
                             "grad_soft": ls.grad.clone()})
    torch.save(out, os.path.join(OUT, "cross_entropy.pt"))


def sgd():
    torch.manual_seed(2)
    n = 1000
    p0 = torch.randn(n)
    grads = [torch.randn(n) for _ in range(5)]
    lrs = [0.1, 0.2, 0.05, 0.4, 0.01]
    out = {"p0": p0, "grads": grads, "lrs": lrs, "runs": {}}
    for nesterov in (False, True):
        p = p0.clone().requires_grad_(True)
        from torch.optim._multi_tensor import SGD as MultiTensorSGD  # the reference optimizer class
        opt = MultiTensorSGD([p], lr=0.0, momentum=0.9, weight_decay=3e-5,
                             nesterov=nesterov)
        traj = []
        for g, lr in zip(grads, lrs):
            opt.param_groups[0]["lr"] = lr
            p.grad = g.clone()
            opt.step()
            traj.append(p.detach().clone())
        out["runs"]["nesterov" if nesterov else "plain"] = traj
    torch.save(out, os.path.join(OUT, "sgd.pt"))


def augment():
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, size=(4, 48, 64, 3), dtype=np.uint8)
    boxes = [augment_ref.rrc_box(48, 64, 0.08, 1.0, 1234, i) for i in range(4)]
    outs = np.stack([augment_ref.augment_image(img[i], boxes[i], 32) for i in range(4)])
    many = np.array([augment_ref.rrc_box(256, 256, 0.08, 1.0, 42, i) for i in range(256)], np.int32)
    wide = np.array([augment_ref.rrc_box(100, 400, 0.9, 1.0, 7, i) for i in range(64)], np.int32)
    torch.save({"img": torch.from_numpy(img), "boxes": torch.tensor(boxes, dtype=torch.int32),
                "out": torch.from_numpy(outs), "boxes_256_seed42": torch.from_numpy(many),
                "boxes_100x400_seed7_minarea09": torch.from_numpy(wide)},
               os.path.join(OUT, "augment.pt"))


def heads_arccos():
    """ArcCosSoftmax and AdaCos(arc_logits) of the reference (angular_losses.py:572-576,
    :323-330), index and soft targets; includes cosines at and beyond the clamp range."""
    ang, _, RefCE = load_reference_modules()
    torch.manual_seed(3)
    b, c = 16, 40
    cos = torch.rand(b, c) * 2 - 1
    cos[0, 0], cos[0, 1], cos[1, 2], cos[2, 3] = 1.0, -1.0, 1.0 - 1e-7, 0.99999
    y = torch.randint(0, c, (b,))
    y[0], y[1] = 0, 2
    soft = torch.zeros(b, c).scatter_(1, y[:, None], 0.7)
    soft.scatter_(1, ((y + 5) % c)[:, None], 0.3)
    out = {"cos": cos, "y": y, "soft": soft}
    for name, crit, tgt in (
        ("arccos", ang.ArcCosSoftmax(smoothing=0.1), y),
        ("arccos_t015", ang.ArcCosSoftmax(smoothing=0.1, temperature=0.15), y),
        ("arccos_soft", ang.ArcCosSoftmax(smoothing=0.1), soft),
        ("adacos_arc", ang.AdaCos(final_criterion=RefCE(smoothing=0.1), margin=0.2, fixed_s=10,
                                  arc_logits=True, arc_margin=True), y),
        ("adacos_arc_soft", ang.AdaCos(final_criterion=RefCE(smoothing=0.1), margin=0.2, fixed_s=10,
                                       arc_logits=True, arc_margin=True), soft),
    ):
        cr = cos.clone().requires_grad_(True)
        loss = crit(cr, tgt)
        loss.backward()
        out[name] = {"loss": loss.detach(), "dcos": cr.grad.clone()}
    l = torch_ref.smooth_cross_entropy(torch_ref.arccos_logits(cos), y, 0.1)
    assert torch.allclose(l, out["arccos"]["loss"], atol=1e-6)
    l = torch_ref.smooth_cross_entropy(torch_ref.arccos_logits(cos, soft, 10.0, 0.2), soft, 0.1)
    assert torch.allclose(l, out["adacos_arc_soft"]["loss"], atol=1e-5)
    torch.save(out, os.path.join(OUT, "heads_arccos.pt"))


def novograd():
    """The reference's own MyNovograd (sota_imagenet/optimizers.py:35-161) on a small mixed bag
    of tensors (conv with odd fan 7*7*3, 3x3 conv, linear, 1-D, a gradient-less tensor), five
    steps with a changing learning rate, whole-tensor and unitwise norms."""
    _, ropt, _ = load_reference_modules()
    torch.manual_seed(4)
    shapes = [(8, 3, 7, 7), (16, 8, 3, 3), (10, 12), (24,), (6, 4, 1, 1), (5,)]
    p0 = [torch.randn(s) * 0.5 for s in shapes]
    steps = 5
    grads = [[torch.randn(s) for s in shapes] for _ in range(steps)]
    lrs = [1e-2, 2e-2, 5e-3, 4e-2, 1e-3]
    nograd_index = 4            # this tensor never receives a gradient (reference :104-110 skips it)
    out = {"shapes": shapes, "p0": p0, "grads": grads, "lrs": lrs, "nograd_index": nograd_index,
           "runs": {}}
    for unitwise in (False, True):
        ps = [p.clone().requires_grad_(True) for p in p0]
        opt = ropt.MyNovograd(ps, lr=0.0, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-2,
                              unitwise_norm=unitwise)
        traj = []
        for gs, lr in zip(grads, lrs):
            opt.param_groups[0]["lr"] = lr
            for i, (p, g) in enumerate(zip(ps, gs)):
                p.grad = None if i == nograd_index else g.clone()
            opt.step()
            traj.append([p.detach().clone() for p in ps])
        out["runs"]["unitwise" if unitwise else "tensor"] = {
            "traj": traj,
            "ema_grad": [opt.state[p]["ema_grad"].clone() if p in opt.state else None for p in ps],
            "ema_norm": [opt.state[p]["ema_norm"].clone() if p in opt.state else None for p in ps]}
        # the written-out restatement must reproduce the reference class
        qs = [p.clone() for p in p0]
        state = [{} for _ in qs]
        for gs, lr, want in zip(grads, lrs, traj):
            idx = [i for i in range(len(qs)) if i != nograd_index]
            torch_ref.novograd_step([qs[i] for i in idx], [gs[i] for i in idx],
                                    [state[i] for i in idx], lr, unitwise=unitwise)
            for q, w in zip(qs, want):
                assert torch.allclose(q, w, atol=1e-6, rtol=1e-5)
    torch.save(out, os.path.join(OUT, "novograd.pt"))


def resnet_step():
    """Tiny pin of the whole-step oracle: torchvision ResNet-50, B=2, 64x64, seed 0."""
    model = torch_ref.resnet50(seed=0)
    opt = torch_ref.make_sgd(model.parameters(), lr=0.1)
    x, y = torch_ref.synthetic_batch(2, 64, seed=0)
    losses = [torch_ref.train_step(model, opt, x, y) for _ in range(2)]
    torch.save({"losses": losses, "fc_bias_after": model.fc.bias.detach().clone(),
                "bn1_running_mean": model.bn1.running_mean.clone()},
               os.path.join(OUT, "resnet50_step.pt"))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    heads()
    cross_entropy()
    sgd()
    augment()
    resnet_step()
    heads_arccos()
    novograd()
    sphere_mlp()
    print("golden vectors written to", os.path.abspath(OUT))
