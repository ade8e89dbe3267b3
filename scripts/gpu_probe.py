"""On-GPU diagnostic battery (not a test): runs kernel groups in isolated subprocesses so one
CUDA fault cannot hide the other results.  Usage:
    python scripts/gpu_probe.py            # all groups, each in its own process
    python scripts/gpu_probe.py conv_fprop # one group in this process
Checker = torch fp32 ops on the GPU (cuDNN) over the same bf16-rounded inputs; that is test
infrastructure only.
"""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

GROUPS = ["conv_fprop", "conv_fprop_big", "conv_dgrad", "conv_wgrad", "norm", "head", "optim_data", "timing"]


def rel_err(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def report(name, err, tol, extra=""):
    print("%-58s rel_err=%.3e %s %s" % (name, err, "OK " if err <= tol else "FAIL", extra), flush=True)


def describe_mismatch(y, ref):
    """Print where a [N,C,H,W] result differs (helps decode layout / descriptor mistakes)."""
    import torch
    d = (y.float() - ref.float()).abs()
    tol = 0.05 * ref.float().abs().max().item() + 1e-3
    bad = d > tol
    print("    mismatching elements: %d / %d" % (bad.sum().item(), bad.numel()))
    if bad.any():
        n, c, h, w = y.shape
        print("    bad per image    :", bad.sum(dim=(1, 2, 3)).tolist()[:8])
        print("    bad per channel  :", bad.sum(dim=(0, 2, 3)).tolist()[:64])
        print("    bad per row (h)  :", bad.sum(dim=(0, 1, 3)).tolist()[:64])
        print("    bad per col (w)  :", bad.sum(dim=(0, 1, 2)).tolist()[:64])
        idx = bad.nonzero()[:5].tolist()
        for i in idx:
            print("    at", i, "got", y[tuple(i)].item(), "want", ref[tuple(i)].item())


def run_conv_fprop(big=False):
    import torch
    import torch.nn.functional as F
    from sota_imagenet_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    cases = [
        # name, N, H, W, C, K, R, stride, pad, flags
        ("1x1 tiled 64->64", 2, 16, 16, 64, 64, 1, 1, 0, 0),
        ("1x1 tiled 128->256", 2, 16, 16, 128, 256, 1, 1, 0, 0),
        ("1x1 im2col 64->64", 2, 16, 16, 64, 64, 1, 1, 0, ops.FLAG_FORCE_IM2COL),
        ("1x1 im2col 256->128 M-tail", 3, 7, 7, 256, 128, 1, 1, 0, ops.FLAG_FORCE_IM2COL),
        ("3x3 s1 64->64", 2, 16, 16, 64, 64, 3, 1, 1, 0),
        ("3x3 s1 128->128 14x14", 4, 14, 14, 128, 128, 3, 1, 1, 0),
        ("3x3 s2 128->128", 2, 28, 28, 128, 128, 3, 2, 1, 0),
        ("1x1 s2 256->512", 2, 28, 28, 256, 512, 1, 2, 0, 0),
        ("3x3 s1 7x7 512->512", 3, 7, 7, 512, 512, 3, 1, 1, 0),
        ("FC-like 2048->1000", 32, 1, 1, 2048, 1000, 1, 1, 0, 0),
    ]
    if big:
        cases = [
            ("1x1 tiled 64->256 56x56 B64 (persistent)", 64, 56, 56, 64, 256, 1, 1, 0, 0),
            ("1x1 tiled 64->256 56x56 B64 (streaming)", 64, 56, 56, 64, 256, 1, 1, 0, 0),
            ("1x1 256->1024 14x14 B256 (ws, K=256)", 256, 14, 14, 256, 1024, 1, 1, 0, 0),
            ("1x1 512->128 28x28 B64 (ws, K=512)", 64, 28, 28, 512, 128, 1, 1, 0, 0),
            ("1x1 s2 256->512 56x56 B32 (ws im2col)", 32, 56, 56, 256, 512, 1, 2, 0, 0),
            ("1x1 256->64 56x56 B64 (ws 64-wide)", 64, 56, 56, 256, 64, 1, 1, 0, 0),
            ("3x3 s1 512->512 7x7 B256 N-tiles", 256, 7, 7, 512, 512, 3, 1, 1, 0),
            ("1x1 tiled 256->64 56x56 B32", 32, 56, 56, 256, 64, 1, 1, 0, 0),
            ("3x3 s1 64->64 56x56 B32", 32, 56, 56, 64, 64, 3, 1, 1, 0),
            ("3x3 s2 256->256 28x28 B32", 32, 28, 28, 256, 256, 3, 2, 1, 0),
            ("1x1 1024->2048 s2 14x14 B32", 32, 14, 14, 1024, 2048, 1, 2, 0, 0),
        ]
    for name, n, h, w, c, k, r, stride, pad, flags in cases:
        x = torch.randn(n, c, h, w, device=dev)
        wt = torch.randn(k, c, r, r, device=dev) / (c * r * r) ** 0.5
        xb = ops.to_nhwc_bf16(x)
        wb = wt.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        ref = F.conv2d(xb.float(), wb.float(), stride=stride, padding=pad)
        stats = torch.empty(2, k, device=dev)
        try:
            y = ops.conv2d_fprop(xb, wb, stride=stride, pad=pad, stats=stats, flags=flags)
            torch.cuda.synchronize()
        except Exception as e:  # noqa
            print("%-58s EXCEPTION %s" % (name, e), flush=True)
            raise
        e = rel_err(y, ref)
        report("fprop " + name, e, 1e-2)
        if e > 1e-2:
            describe_mismatch(y, ref)
        yf = y.float()
        sref = torch.stack([yf.sum(dim=(0, 2, 3)), (yf * yf).sum(dim=(0, 2, 3))])
        report("  stats " + name, rel_err(stats, sref), 1e-3)
    # stem-like: 4x1 filter, asymmetric padding (2 top / 1 bottom), Cin = 64
    n, h, w, c, k = 2, 16, 16, 64, 64
    x = torch.randn(n, c, h, w, device=dev)
    wt = torch.randn(k, c, 4, 1, device=dev) / 16
    xb = ops.to_nhwc_bf16(x)
    wb = wt.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ref = F.conv2d(F.pad(xb.float(), (0, 0, 2, 1)), wb.float())
    y = ops.conv2d_fprop(xb, wb, stride=1, pad_hw=(2, 0), out_hw=(h, w))
    torch.cuda.synchronize()
    e = rel_err(y, ref)
    report("fprop 4x1 asym pad (stem form)", e, 1e-2)
    if e > 1e-2:
        describe_mismatch(y, ref)


def run_conv_dgrad():
    import torch
    import torch.nn.functional as F
    from sota_imagenet_b200 import ops
    torch.manual_seed(1)
    dev = "cuda"
    cases = [
        ("1x1 s1 64->256", 2, 16, 16, 64, 256, 1, 1, 0),
        ("3x3 s1 64->64", 2, 16, 16, 64, 64, 3, 1, 1),
        ("3x3 s1 128->128 7x7", 3, 7, 7, 128, 128, 3, 1, 1),
        ("1x1 s2 256->512", 2, 28, 28, 256, 512, 1, 2, 0),
        ("3x3 s2 128->128", 2, 28, 28, 128, 128, 3, 2, 1),
        ("FC-like 2048->1000", 32, 1, 1, 2048, 1000, 1, 1, 0),
    ]
    for name, n, h, w, c, k, r, stride, pad in cases:
        x = torch.randn(n, c, h, w, device=dev, requires_grad=True)
        wt = (torch.randn(k, c, r, r, device=dev) / (c * r * r) ** 0.5).to(torch.bfloat16).float()
        y = F.conv2d(x, wt, stride=stride, padding=pad)
        dy = torch.randn_like(y)
        dyb = ops.to_nhwc_bf16(dy)
        (ref,) = torch.autograd.grad(y, x, dyb.float())
        wb = wt.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        wd = ops.pack_dgrad_weight(wb)
        # check the pack itself
        wd_ref = wb.float().flip(2, 3).permute(1, 2, 3, 0).contiguous()
        report("pack  " + name, rel_err(wd, wd_ref), 0.0)
        dx = ops.conv2d_dgrad(dyb, wd, (n, c, h, w), r, r, stride=stride, pad=pad)
        torch.cuda.synchronize()
        e = rel_err(dx, ref)
        report("dgrad " + name, e, 1e-2)
        if e > 1e-2:
            describe_mismatch(dx, ref)
        if stride == 1:
            base = ops.to_nhwc_bf16(torch.randn(n, c, h, w, device=dev))
            acc = ops.conv2d_dgrad(dyb, wd, (n, c, h, w), r, r, stride=stride, pad=pad, residual=base)
            torch.cuda.synchronize()
            report("dgrad+residual " + name, rel_err(acc, ref + base.float()), 1e-2)
            inpl = base.clone()
            ops.conv2d_dgrad(dyb, wd, (n, c, h, w), r, r, stride=stride, pad=pad, out=inpl, residual=inpl)
            torch.cuda.synchronize()
            report("dgrad+residual in place " + name, rel_err(inpl, ref + base.float()), 1e-2)
        elif r == 1:
            base = ops.to_nhwc_bf16(torch.randn(n, c, h, w, device=dev))
            acc = base.clone()
            ops.conv2d_dgrad(dyb, wd, (n, c, h, w), r, r, stride=stride, pad=pad, out=acc)
            torch.cuda.synchronize()
            report("dgrad strided scatter-add " + name, rel_err(acc, ref + base.float()), 1e-2)


def run_conv_wgrad():
    import torch
    import torch.nn.functional as F
    from sota_imagenet_b200 import ops
    torch.manual_seed(2)
    dev = "cuda"
    cases = [
        ("1x1 s1 64->64", 2, 16, 16, 64, 64, 1, 1, 0),
        ("1x1 s1 128->256", 2, 16, 16, 128, 256, 1, 1, 0),
        ("3x3 s1 64->64", 2, 16, 16, 64, 64, 3, 1, 1),
        ("3x3 s1 128->128 7x7 (pixel tail)", 3, 7, 7, 128, 128, 3, 1, 1),
        ("3x3 s2 128->128", 2, 28, 28, 128, 128, 3, 2, 1),
        ("1x1 s2 256->512", 2, 28, 28, 256, 512, 1, 2, 0),
        ("FC-like 2048->1000", 32, 1, 1, 2048, 1000, 1, 1, 0),
        ("3x3 s1 256->256 14x14 B16", 16, 14, 14, 256, 256, 3, 1, 1),
    ]
    for name, n, h, w, c, k, r, stride, pad in cases:
        xb = ops.to_nhwc_bf16(torch.randn(n, c, h, w, device=dev))
        wt = torch.zeros(k, c, r, r, device=dev, requires_grad=True)
        y = F.conv2d(xb.float(), wt, stride=stride, padding=pad)
        dyb = ops.to_nhwc_bf16(torch.randn_like(y))
        (ref,) = torch.autograd.grad(y, wt, dyb.float())
        dw = torch.zeros(k, r, r, c, device=dev).permute(0, 3, 1, 2)
        ops.conv2d_wgrad(xb, dyb, dw, stride=stride, pad=pad)
        torch.cuda.synchronize()
        e = rel_err(dw, ref)
        report("wgrad " + name, e, 1e-2)
        if e > 1e-2:
            describe_mismatch(dw, ref)


def run_norm():
    import torch
    import torch.nn.functional as F
    from sota_imagenet_b200 import ops
    torch.manual_seed(3)
    dev = "cuda"
    for (n, c, h, w) in [(4, 64, 16, 16), (3, 256, 7, 7), (2, 2048, 7, 7), (8, 128, 28, 28)]:
        x = ops.to_nhwc_bf16(torch.randn(n, c, h, w, device=dev) * 2 + 0.5)
        xf = x.float()
        stats = ops.bn_stats(x)
        sref = torch.stack([xf.sum(dim=(0, 2, 3)), (xf * xf).sum(dim=(0, 2, 3))])
        report("bn_stats %s" % ((n, c, h, w),), rel_err(stats, sref), 1e-5)
        gamma = torch.rand(c, device=dev) + 0.5
        beta = torch.randn(c, device=dev)
        rm = torch.zeros(c, device=dev)
        rv = torch.ones(c, device=dev)
        m = n * h * w
        mi, ss = ops.bn_finalize(stats, gamma, beta, rm, rv, m, 1e-5, 0.1)
        bn = torch.nn.BatchNorm2d(c).to(dev)
        bn.weight.data.copy_(gamma)
        bn.bias.data.copy_(beta)
        xr = xf.clone().requires_grad_(True)
        res = ops.to_nhwc_bf16(torch.randn(n, c, h, w, device=dev))
        yref = F.relu(bn(xr) + res.float())
        y = ops.bn_apply(x, ss, act=ops.ACT_RELU, res=res)
        report("bn_apply+res+relu", rel_err(y, yref), 1e-2)
        report("running_mean", rel_err(rm, bn.running_mean), 1e-4)
        report("running_var", rel_err(rv, bn.running_var), 1e-4)
        dy = ops.to_nhwc_bf16(torch.randn(n, c, h, w, device=dev))
        # reference gradient with the mask taken from the bf16 output actually stored
        mask = (y.float() > 0).float()
        g = dy.float() * mask
        xhat = (xf - mi[0].view(1, -1, 1, 1)) * mi[1].view(1, -1, 1, 1)
        sg = g.sum(dim=(0, 2, 3))
        sgx = (g * xhat).sum(dim=(0, 2, 3))
        dxref = (gamma * mi[1]).view(1, -1, 1, 1) * (g - sg.view(1, -1, 1, 1) / m - xhat * sgx.view(1, -1, 1, 1) / m)
        sums = ops.bn_bwd_reduce(dy, y, x, mi, ops.ACT_RELU)
        # plain BN+ReLU with the mask recomputed from x
        y0 = ops.bn_apply(x, ss, act=ops.ACT_RELU)
        s_out = ops.bn_bwd_reduce(dy, y0, x, mi, ops.ACT_RELU)
        s_re = ops.bn_bwd_reduce(dy, None, x, mi, ops.ACT_RELU, mask_ss=ss)
        report("bn_bwd_reduce remask == out-mask", rel_err(s_re, s_out), 1e-6)
        d_out, _, _ = ops.bn_bwd_apply(dy, y0, x, mi, gamma, s_out, m, ops.ACT_RELU)
        d_re, _, _ = ops.bn_bwd_apply(dy, None, x, mi, gamma, s_out, m, ops.ACT_RELU, mask_ss=ss)
        report("bn_bwd_apply remask == out-mask", rel_err(d_re, d_out), 0.0)
        report("bn_bwd_reduce", rel_err(sums, torch.stack([sg, sgx])), 1e-4)
        dx, _, gout = ops.bn_bwd_apply(dy, y, x, mi, gamma, sums, m, ops.ACT_RELU, want_g=True)
        report("bn_bwd_apply dx", rel_err(dx, dxref), 1e-2)
        report("bn_bwd_apply g", rel_err(gout, g), 1e-2)
    # max pool / gap
    x = ops.to_nhwc_bf16(torch.randn(4, 64, 32, 32, device=dev))
    y, idx = ops.maxpool3x3s2_fwd(x)
    xr = x.float().requires_grad_(True)
    yr = F.max_pool2d(xr, 3, 2, 1)
    report("maxpool fwd", rel_err(y, yr), 0.0)
    dy = ops.to_nhwc_bf16(torch.randn_like(yr))
    (dxr,) = torch.autograd.grad(yr, xr, dy.float())
    dx = ops.maxpool3x3s2_bwd(dy, idx, x.shape)
    report("maxpool bwd", rel_err(dx, dxr), 1e-2)
    x = ops.to_nhwc_bf16(torch.randn(4, 2048, 7, 7, device=dev))
    report("gap fwd", rel_err(ops.gap_fwd(x), x.float().mean(dim=(2, 3), keepdim=True)), 1e-2)
    dy = ops.to_nhwc_bf16(torch.randn(4, 2048, 1, 1, device=dev))
    report("gap bwd", rel_err(ops.gap_bwd(dy, x.shape), dy.float().expand(4, 2048, 7, 7) / 49), 1e-2)


def run_head():
    import torch
    import torch.nn.functional as F
    from sota_imagenet_b200 import ops
    torch.manual_seed(4)
    dev = "cuda"
    b, c = 64, 1000
    logits = torch.randn(b, c, device=dev) * 3
    y = torch.randint(0, c, (b,), device=dev)
    for sm in (0.0, 0.1):
        lr = logits.clone().requires_grad_(True)
        ref = F.cross_entropy(lr, y, label_smoothing=sm)
        (gref,) = torch.autograd.grad(ref, lr)
        loss, rows, dl = ops.ce_fwd_bwd(logits, y, smoothing=sm)
        report("ce fp32 smoothing=%.1f loss" % sm, abs(loss.item() - ref.item()) / abs(ref.item()), 1e-5)
        report("ce fp32 smoothing=%.1f grad" % sm, rel_err(dl, gref), 1e-4)
        onehot = ops.one_hot(y, c)
        loss2, _, dl2 = ops.ce_fwd_bwd(logits, onehot, smoothing=sm)
        report("ce dense-target loss", abs(loss2.item() - ref.item()) / abs(ref.item()), 1e-5)
        lb = logits.to(torch.bfloat16)
        loss3, _, dl3 = ops.ce_fwd_bwd(lb, y, smoothing=sm)
        lr3 = lb.float().requires_grad_(True)
        ref3 = F.cross_entropy(lr3, y, label_smoothing=sm)
        report("ce bf16 loss", abs(loss3.item() - ref3.item()) / abs(ref3.item()), 1e-5)
    # sphere linear + arc / cos margins
    d = 512
    x = torch.randn(b, d, device=dev)
    w = torch.randn(c, d, device=dev)
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    cos_ref = F.linear(F.normalize(xr, dim=1), F.normalize(wr, dim=1))
    cosv, saved = ops.sphere_linear_fwd(x, w)
    report("sphere_linear fwd", rel_err(cosv, cos_ref), 1e-5)
    import math
    for kind, s, m in ((ops.MARGIN_ARC, 10.0, 0.2), (ops.MARGIN_COS, 30.0, 0.4)):
        onehot = F.one_hot(y, c).float()
        if kind == ops.MARGIN_ARC:
            sine = torch.sqrt((1.0 - cos_ref * cos_ref).clamp_min(0))
            phi = cos_ref * math.cos(m) - sine * math.sin(m)
            phi = torch.where(cos_ref > math.cos(math.pi - m), phi, cos_ref - math.sin(math.pi - m) * m)
            out = s * (onehot * phi + (1 - onehot) * cos_ref)
        else:
            out = s * (cos_ref - m * onehot)
        ref = F.cross_entropy(out, y, label_smoothing=0.1)
        gx, gw = torch.autograd.grad(ref, (xr, wr), retain_graph=True)
        loss, _, dcos = ops.ce_fwd_bwd(cosv, y, smoothing=0.1, margin_kind=kind, s=s, m=m)
        dx, dw = ops.sphere_linear_bwd(dcos, saved)
        report("margin kind=%d loss" % kind, abs(loss.item() - ref.item()) / abs(ref.item()), 1e-4)
        report("margin kind=%d dx" % kind, rel_err(dx, gx), 1e-3)
        report("margin kind=%d dw" % kind, rel_err(dw, gw), 1e-3)


def run_optim_data():
    import torch
    from sota_imagenet_b200 import ops
    torch.manual_seed(5)
    dev = "cuda"
    n = 4096 * 3 + 64
    p = torch.randn(n, device=dev)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.SGD([pr], lr=0.1, momentum=0.9, weight_decay=3e-5, nesterov=True)
    buf = torch.zeros(n, device=dev)
    pb = torch.empty(n, dtype=torch.bfloat16, device=dev)
    for step in range(3):
        g = torch.randn(n, device=dev)
        pr.grad = g.clone()
        opt.step()
        segs = ops.sgd_segments([(n, 0.1, 3e-5, 0.9, 0.0, True)], dev)
        ops.sgd_step(p, g, buf, pb, segs, 1, step == 0)
        report("sgd nesterov step %d" % step, rel_err(p, pr.detach()), 1e-6)
    report("sgd bf16 copy", rel_err(pb, p), 4e-3)
    # augmentation boxes: device vs host twin (bit exact)
    boxes = ops.rrc_boxes(512, 256, 256, 0.08, 1.0, 1234, 0, True, dev).cpu().tolist()
    host = [ops.rrc_box_host(256, 256, 0.08, 1.0, 1234, i) for i in range(512)]
    print("rrc boxes device == host twin: %s" % (boxes == host), flush=True)
    src = torch.randint(0, 256, (8, 256, 256, 3), dtype=torch.uint8, device=dev)
    bx = ops.rrc_boxes(8, 256, 256, 0.08, 1.0, 1, 0, True, dev)
    out0 = ops.augment(src, bx, 224, out_mode=0)
    out1 = ops.augment(src, bx, 224, out_mode=1)
    report("augment NHWC4 vs NCHW", rel_err(out0[:, :3], out1), 4e-3)
    print("augment range: min %.3f max %.3f mean %.3f" % (out1.min().item(), out1.max().item(), out1.mean().item()))
    # stem packing + 4x1 conv == 7x7/2 conv
    import torch.nn.functional as F
    x = torch.randn(4, 3, 64, 64, device=dev)
    w = torch.randn(64, 3, 7, 7, device=dev) / 12
    xq = ops.stem_pack(x, 7, 3)
    wq = ops.stem_pack_weight(w, 4, 1)
    y = ops.conv2d_fprop(xq, wq.permute(0, 3, 1, 2), stride=1, pad_hw=(2, 0), out_hw=(32, 32))
    ref = F.conv2d(x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), stride=2, padding=3)
    report("stem 7x7/2 via packed 4x1 conv", rel_err(y, ref), 1e-2)
    xr = x.to(torch.bfloat16).float()
    wr = w.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, stride=2, padding=3)
    dy = ops.to_nhwc_bf16(torch.randn_like(yr))
    (gref,) = torch.autograd.grad(yr, wr, dy.float())
    dwq = torch.zeros(64, 4, 1, 64, device=dev).permute(0, 3, 1, 2)
    ops.conv2d_wgrad(xq, dy, dwq, stride=1, pad_hw=(2, 0))
    dw = torch.zeros_like(w)
    ops.stem_unpack_wgrad(dwq, dw, 4, 1, accumulate=False)
    report("stem wgrad via packed form", rel_err(dw, gref), 1e-2)


def run_timing():
    import torch
    import torch.nn.functional as F
    from sota_imagenet_b200 import ops
    dev = "cuda"
    B = 256

    def timeit(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    shapes = [(64, 64, 56, 1, 1, 0), (64, 64, 56, 3, 1, 1), (64, 256, 56, 1, 1, 0), (256, 64, 56, 1, 1, 0),
              (128, 128, 28, 3, 1, 1), (256, 256, 14, 3, 1, 1), (512, 512, 7, 3, 1, 1),
              (1024, 256, 14, 1, 1, 0), (256, 1024, 14, 1, 1, 0), (512, 2048, 7, 1, 1, 0),
              (128, 128, 56, 3, 2, 1), (256, 512, 56, 1, 2, 0)]
    print("%-28s %10s %10s %10s | cudnn fwd" % ("shape Cin,Cout,H,k,s", "fprop ms", "dgrad ms", "wgrad ms"))
    for (c, k, h, r, stride, pad) in shapes:
        x = ops.to_nhwc_bf16(torch.randn(B, c, h, h, device=dev))
        w = (torch.randn(k, c, r, r, device=dev) / (c * r * r) ** 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        wd = ops.pack_dgrad_weight(w)
        stats = torch.empty(2, k, device=dev)
        y = ops.conv2d_fprop(x, w, stride=stride, pad=pad, stats=stats)
        dy = torch.randn_like(y)
        dw = torch.zeros(k, r, r, c, device=dev).permute(0, 3, 1, 2)
        t_f = timeit(lambda: ops.conv2d_fprop(x, w, stride=stride, pad=pad, stats=stats))
        dxbuf = torch.zeros_like(x)
        t_d = timeit(lambda: ops.conv2d_dgrad(dy, wd, tuple(x.shape), r, r, stride=stride, pad=pad, out=dxbuf))
        t_w = timeit(lambda: ops.conv2d_wgrad(x, dy, dw, stride=stride, pad=pad))
        t_c = timeit(lambda: F.conv2d(x, w, stride=stride, padding=pad))
        flops = 2.0 * B * y.shape[2] * y.shape[3] * k * c * r * r
        print("%-28s %7.3f(%4.0fT) %7.3f(%4.0fT) %7.3f(%4.0fT) | %7.3f(%4.0fT)" % (
            str((c, k, h, r, stride)), t_f, flops / t_f / 1e9, t_d, flops / t_d / 1e9, t_w,
            flops / t_w / 1e9, t_c, flops / t_c / 1e9), flush=True)
    # memory-bound kernels
    x = ops.to_nhwc_bf16(torch.randn(B, 256, 56, 56, device=dev))
    gb = x.numel() * 2 / 1e9
    t = timeit(lambda: ops.bn_stats(x))
    print("bn_stats   256x56x56: %.3f ms  %.0f GB/s" % (t, gb / t * 1e3))
    ss = torch.randn(2, 256, device=dev)
    out = torch.empty_like(x)
    t = timeit(lambda: ops.bn_apply(x, ss, act=ops.ACT_RELU, out=out))
    print("bn_apply   256x56x56: %.3f ms  %.0f GB/s" % (t, 2 * gb / t * 1e3))
    t = timeit(lambda: ops.bn_apply(x, ss, act=ops.ACT_RELU, res=x, out=out))
    print("bn_add_relu 256x56x56: %.3f ms  %.0f GB/s" % (t, 3 * gb / t * 1e3))


RUNNERS = {
    "conv_fprop": lambda: run_conv_fprop(False),
    "conv_fprop_big": lambda: run_conv_fprop(True),
    "conv_dgrad": run_conv_dgrad,
    "conv_wgrad": run_conv_wgrad,
    "norm": run_norm,
    "head": run_head,
    "optim_data": run_optim_data,
    "timing": run_timing,
}

if __name__ == "__main__":
    if len(sys.argv) > 1:
        RUNNERS[sys.argv[1]]()
    else:
        for g in GROUPS:
            print("===== %s =====" % g, flush=True)
            t0 = time.time()
            r = subprocess.run([sys.executable, os.path.abspath(__file__), g], timeout=600)
            print("===== %s exit=%d (%.1fs) =====" % (g, r.returncode, time.time() - t0), flush=True)
