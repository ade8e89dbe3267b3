"""CPU: the oracle restatements against the committed golden vectors, which were produced by
the REFERENCE's own code (oracle/make_golden.py imports /root/reference/sota_imagenet/
angular_losses.py and torch.optim._multi_tensor.SGD) in the build container."""
import os

import numpy as np
import torch

from oracle import augment_ref, torch_ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def test_heads_restatement_matches_reference_classes():
    g = _load("heads.pt")
    x, w, y = g["x"], g["w"], g["y"]
    for name, (s, m, sm, fn) in {
        "arc": (10.0, 0.2, 0.1, torch_ref.arcface_logits),
        "arc_s64": (64.0, 0.5, 0.0, torch_ref.arcface_logits),
        "adacos_fixed": (10.0, 0.2, 0.1, torch_ref.cosface_logits),
    }.items():
        xr = x.clone().requires_grad_(True)
        wr = w.clone().requires_grad_(True)
        cos = torch_ref.sphere_linear(xr, wr)
        loss = torch_ref.smooth_cross_entropy(fn(cos, y, s, m), y, sm)
        loss.backward()
        assert torch.allclose(cos, g[name]["cos"], atol=1e-6)
        assert torch.allclose(loss, g[name]["loss"], atol=1e-5), name
        assert torch.allclose(xr.grad, g[name]["dx"], atol=1e-5, rtol=1e-4), name
        assert torch.allclose(wr.grad, g[name]["dw"], atol=1e-5, rtol=1e-4), name
    # LargeMarginCosineLoss: only W is normalised
    xn = g["cosface_lm"]["xn"]
    cos = torch.nn.functional.linear(xn, torch.nn.functional.normalize(w))
    loss = torch_ref.smooth_cross_entropy(torch_ref.cosface_logits(cos, y, 30.0, 0.4), y, 0.0)
    assert torch.allclose(loss, g["cosface_lm"]["loss"], atol=1e-5)


def test_cross_entropy_restatement():
    g = _load("cross_entropy.pt")
    for case in g["cases"]:
        lr = g["logits"].clone().requires_grad_(True)
        loss = torch_ref.smooth_cross_entropy(lr, g["y"], case["smoothing"], case["temperature"])
        loss.backward()
        assert torch.allclose(loss, case["loss"], atol=1e-6)
        assert torch.allclose(lr.grad, case["grad"], atol=1e-7)
        # closed-form gradient of SURVEY App. E.2 (what the CUDA kernel implements)
        t = torch.zeros_like(g["logits"]).scatter_(1, g["y"][:, None], 1.0)
        s, T = case["smoothing"], case["temperature"]
        q = (1 - s) * t + s / t.shape[1]
        p = torch.softmax(g["logits"] / T, 1)
        grad = (p * q.sum(1, keepdim=True) - q) / (t.shape[0] * T)
        assert torch.allclose(grad, case["grad"], atol=1e-6)


def test_sgd_arithmetic_restatement():
    """torch/optim/sgd.py semantics written out == reference optimizer runs."""
    g = _load("sgd.pt")
    for nesterov, key in ((False, "plain"), (True, "nesterov")):
        p = g["p0"].clone()
        buf = None
        for grad, lr, want in zip(g["grads"], g["lrs"], g["runs"][key]):
            d = grad + 3e-5 * p
            buf = d.clone() if buf is None else 0.9 * buf + d
            d = d + 0.9 * buf if nesterov else buf
            p = p - lr * d
            assert torch.allclose(p, want, atol=1e-6, rtol=1e-6)


def test_augment_oracle_vs_golden_and_host_twin():
    g = _load("augment.pt")
    boxes = [augment_ref.rrc_box(256, 256, 0.08, 1.0, 42, i) for i in range(256)]
    assert np.array_equal(np.array(boxes, np.int32), g["boxes_256_seed42"].numpy())
    from sota_imagenet_b200 import ops
    host = [ops.rrc_box_host(256, 256, 0.08, 1.0, 42, i) for i in range(256)]
    assert host == boxes                         # C twin of the CUDA generator, bit-exact
    wide = [ops.rrc_box_host(100, 400, 0.9, 1.0, 7, i) for i in range(64)]
    assert np.array_equal(np.array(wide, np.int32), g["boxes_100x400_seed7_minarea09"].numpy())
    # invariants of dali_dataloader.py:65-72: box inside the image, aspect in [0.75,1.25] (+rounding)
    for x0, y0, w, h, flip in boxes:
        assert 0 <= x0 and 0 <= y0 and x0 + w <= 256 and y0 + h <= 256 and flip in (0, 1)
        assert 0.70 <= w / h <= 1.32
    out = augment_ref.augment_image(g["img"][0].numpy(), g["boxes"][0].tolist(), 32)
    assert np.allclose(out, g["out"][0].numpy(), atol=1e-6)
    assert out.min() >= -2.5001 and out.max() <= 2.5001


def test_resnet_oracle_pinned():
    g = _load("resnet50_step.pt")
    model = torch_ref.resnet50(seed=0)
    opt = torch_ref.make_sgd(model.parameters(), lr=0.1)
    x, y = torch_ref.synthetic_batch(2, 64, seed=0)
    losses = [torch_ref.train_step(model, opt, x, y) for _ in range(2)]
    assert np.allclose(losses, g["losses"], rtol=1e-4)
    assert torch.allclose(model.bn1.running_mean, g["bn1_running_mean"], atol=1e-5)


def test_arccos_restatement_matches_reference_classes():
    """oracle arccos_logits + smooth CE == the reference's ArcCosSoftmax / AdaCos(arc_logits)
    (angular_losses.py:572-576, :323-330), losses and gradients, index and soft targets."""
    g = _load("heads_arccos.pt")
    cos, y, soft = g["cos"], g["y"], g["soft"]
    for name, tgt, s, m, temp in (("arccos", y, 1.0, 0.0, 1.0), ("arccos_t015", y, 1.0, 0.0, 0.15),
                                  ("arccos_soft", soft, 1.0, 0.0, 1.0), ("adacos_arc", y, 10.0, 0.2, 1.0),
                                  ("adacos_arc_soft", soft, 10.0, 0.2, 1.0)):
        cr = cos.clone().requires_grad_(True)
        loss = torch_ref.smooth_cross_entropy(torch_ref.arccos_logits(cr, tgt, s, m), tgt, 0.1, temp)
        loss.backward()
        assert torch.allclose(loss, g[name]["loss"], atol=1e-5, rtol=1e-5), name
        assert torch.allclose(cr.grad, g[name]["dcos"], atol=1e-5, rtol=1e-4), name
    # the clamp kills the gradient at exactly +-1 and keeps it finite next to it
    assert float(g["arccos"]["dcos"][0, 0]) == 0.0 and float(g["arccos"]["dcos"][0, 1]) == 0.0
    assert torch.isfinite(g["arccos"]["dcos"]).all()


def test_novograd_restatement_matches_reference_optimizer():
    """oracle novograd_step == the reference's own MyNovograd class (optimizers.py:35-161)."""
    g = _load("novograd.pt")
    skip = g["nograd_index"]
    for unitwise, key in ((False, "tensor"), (True, "unitwise")):
        run = g["runs"][key]
        ps = [p.clone() for p in g["p0"]]
        state = [{} for _ in ps]
        idx = [i for i in range(len(ps)) if i != skip]
        for grads, lr, want in zip(g["grads"], g["lrs"], run["traj"]):
            torch_ref.novograd_step([ps[i] for i in idx], [grads[i] for i in idx],
                                    [state[i] for i in idx], lr, unitwise=unitwise)
            for p, w in zip(ps, want):
                assert torch.allclose(p, w, atol=1e-6, rtol=1e-5)
        assert torch.equal(ps[skip], g["p0"][skip])                 # gradient-less tensor untouched
        for i in idx:
            assert torch.allclose(state[i]["ema_grad"], run["ema_grad"][i], atol=1e-6)
            assert torch.allclose(state[i]["ema_norm"].expand_as(ps[i]), run["ema_norm"][i], rtol=1e-5)


def test_val_transform_oracle_properties_and_host_twin():
    """Validation transform (dali_dataloader.py:146-160) restatement: geometry == the C host twin
    of the kernel's, constant images stay constant, identity when nothing is resized, and the
    centre of a down-scaled gradient image keeps the gradient."""
    from sota_imagenet_b200 import ops
    for sh, sw, s, rs in ((256, 256, 224, 256), (320, 427, 224, 256), (500, 375, 128, 144),
                          (64, 48, 32, 32), (333, 500, 224, 224)):
        assert augment_ref.val_geometry(sh, sw, s, rs) == ops.val_geometry_host(sh, sw, s, rs)
    rh, rw, oy0, ox0 = augment_ref.val_geometry(320, 427, 224, 256)
    assert (rh, rw) == (256, 342) and (oy0, ox0) == (16, 59)
    const = np.full((40, 56, 3), 200, np.uint8)
    out = augment_ref.val_transform_image(const, 16, 20)
    assert np.allclose(out, (200 - 127.5) / 51.0, atol=1e-5)
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, size=(24, 24, 3), dtype=np.uint8)
    same = augment_ref.val_transform_image(img, 24, 24)          # no resize, no crop
    assert np.allclose(same, (img.astype(np.float32) - 127.5) / 51.0, atol=1e-5)
    crop = augment_ref.val_transform_image(img, 16, 24)          # no resize, centre crop
    assert np.allclose(crop, (img[4:20, 4:20].astype(np.float32) - 127.5) / 51.0, atol=1e-5)
    ramp = np.tile(np.arange(64, dtype=np.uint8)[None, :, None] * 4, (64, 1, 3))
    out = augment_ref.val_transform_image(ramp, 16, 32)          # 2x down-scale, centre 16 of 32
    d = np.diff(out[8, :, 0])
    assert np.allclose(d, 8.0 / 51.0, atol=1e-4)                 # 2 source pixels (x4) per output pixel
