"""Per-layer conv timing (CUDA events) for every distinct ResNet-50 conv shape at batch 256, with the
roofline time of each (max of tensor time at the measured bf16 peak and HBM time at the measured copy
bandwidth).  `--profile` runs every (shape, pass) once inside a cudaProfiler range for ncu
(`ncu --profile-from-start off ...`)."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from sota_imagenet_b200 import ops
B = 256
peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))
TF, BW = peaks["bf16_tflops_sustained"] * 1e12, peaks["hbm_gbs"] * 1e9
# (Cin, Cout, Hin, k, stride, pad, count)
SHAPES = [(64,64,56,1,1,0,1),(64,64,56,3,1,1,3),(64,256,56,1,1,0,4),(256,64,56,1,1,0,2),(256,128,56,1,1,0,1),
          (128,128,56,3,2,1,1),(128,512,28,1,1,0,4),(256,512,56,1,2,0,1),(512,128,28,1,1,0,3),(128,128,28,3,1,1,3),
          (512,256,28,1,1,0,1),(256,256,28,3,2,1,1),(256,1024,14,1,1,0,6),(512,1024,28,1,2,0,1),(1024,256,14,1,1,0,5),
          (256,256,14,3,1,1,5),(1024,512,14,1,1,0,1),(512,512,14,3,2,1,1),(512,2048,7,1,1,0,3),(1024,2048,14,1,2,0,1),
          (2048,512,7,1,1,0,2),(512,512,7,3,1,1,2)]
FLAGS = int(os.environ.get("SIB_FLAGS", "0"))      # e.g. 2 = N tile capped at 128, 4 = no 2-CTA kernel
profile = "--profile" in sys.argv
only = [a for a in sys.argv[1:] if not a.startswith("--")]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn, iters=8):
    for _ in range(2): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

tot = {"fprop": 0.0, "dgrad": 0.0, "wgrad": 0.0}
ideal_tot = 0.0
print("%-24s %5s | %8s %8s %8s | %8s (ms)  TF/s f/d/w" % ("Cin,Cout,H,k,s", "cnt", "fprop", "dgrad", "wgrad", "ideal"))
for (c, k, h, r, stride, pad, cnt) in SHAPES:
    if only and ("%d,%d,%d,%d,%d" % (c, k, h, r, stride)) not in only:
        continue
    oh = (h + 2 * pad - r) // stride + 1
    x = ops.to_nhwc_bf16(torch.randn(B, c, h, h, device="cuda"))
    w = (torch.randn(k, c, r, r, device="cuda") / (c * r * r) ** 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wd = ops.pack_dgrad_weight(w)
    ws2 = ops.pack_dgrad_s2(wd) if ops.dgrad_s2_ok((B, c, h, h), r, r, stride, pad) and not os.environ.get("NO_S2") else None
    dx_buf = ops.new_act(B, c, h, h, "cuda") if stride > 1 and r == 1 else None
    stats = torch.empty(2, k, device="cuda")
    y = ops.conv2d_fprop(x, w, stride=stride, pad=pad, stats=stats)
    dy = torch.randn_like(y)
    dw = torch.zeros(k, r, r, c, device="cuda").permute(0, 3, 1, 2)
    fns = {"fprop": lambda: ops.conv2d_fprop(x, w, stride=stride, pad=pad, stats=stats, flags=FLAGS),
           "dgrad": lambda: ops.conv2d_dgrad(dy, wd, tuple(x.shape), r, r, stride=stride, pad=pad, flags=FLAGS, w_s2=ws2, out=dx_buf),
           "wgrad": lambda: ops.conv2d_wgrad(x, dy, dw, stride=stride, pad=pad)}
    if "--fused" in sys.argv:
        # dgrad with the fused BN-backward reduction (mask recomputed from the BN input)
        mi = torch.stack([x.float().mean(dim=(0, 2, 3)), torch.ones(c, device="cuda")]).contiguous()
        ss = torch.stack([torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")]).contiguous()
        fuse = dict(mask_src=x, mask_ss=ss, mean_invstd=mi, act=ops.ACT_RELU, slope=0.0)
        fns["dgrad+bn"] = lambda: ops.conv2d_dgrad(dy, wd, tuple(x.shape), r, r, stride=stride, pad=pad, flags=FLAGS, w_s2=ws2, bn_bwd=fuse)
        fns["reduce"] = lambda: ops.bn_bwd_reduce(x, None, x, mi, ops.ACT_RELU, 0.0, mask_ss=ss)
    if "--prologue" in sys.argv:
        # fprop with the producer BatchNorm + ReLU fused into the A-operand prologue vs the separate
        # bn_finalize_apply pass followed by the plain conv (conv2 <- bn1, conv3 <- bn2 shapes)
        if not (ops.fprop_bnact_ok(c) and ((r == 3) or (r == 1 and k == 4 * c and stride == 1))):
            continue
        g1, b1 = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
        rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
        st = ops.bn_stats(x)
        bn = (st, g1, b1, rm, rv)
        fns = {"fprop": fns["fprop"],
               "bn_apply": lambda: ops.bn_finalize_apply(x, bn, act=ops.ACT_RELU, count=B * h * h),
               "fused": lambda: ops.conv2d_fprop_bnact(x, w, bn, stride=stride, pad=pad, stats=stats, count=B * h * h)}
        t = {n: timeit(f) for n, f in fns.items()}
        for n in t: tot[n] = tot.get(n, 0.0) + t[n] * cnt
        print("%-24s %3d | fprop %.3f + bn_apply %.3f = %.3f | fused %.3f ms  (%+.3f)" % (
            str((c, k, h, r, stride)), cnt, t["fprop"], t["bn_apply"], t["fprop"] + t["bn_apply"], t["fused"],
            t["fused"] - t["fprop"] - t["bn_apply"]))
        continue
    flops = 2.0 * B * oh * oh * k * c * r * r
    byt = 2.0 * B * (h * h * c + oh * oh * k) + 2.0 * k * c * r * r
    ideal = max(flops / TF, byt / BW) * 1e3
    if profile:
        for f in fns.values(): f()
        torch.cuda.synchronize()
        flush.zero_()
        torch.cuda.profiler.start()
        for f in fns.values(): f()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        continue
    t = {n: timeit(f) for n, f in fns.items()}
    for n in t: tot[n] = tot.get(n, 0.0) + t[n] * cnt
    ideal_tot += ideal * cnt
    if "--fused" in sys.argv:
        print("%-24s dgrad %.3f  dgrad+bn %.3f  separate reduce %.3f ms" % (str((c, k, h, r, stride)), t["dgrad"], t["dgrad+bn"], t["reduce"]))
        continue
    print("%-24s %5d | %8.3f %8.3f %8.3f | %8.3f   %4.0f %4.0f %4.0f" % (
        str((c, k, h, r, stride)), cnt, t["fprop"], t["dgrad"], t["wgrad"], ideal,
        flops / t["fprop"] / 1e9, flops / t["dgrad"] / 1e9, flops / t["wgrad"] / 1e9))
    del x, w, wd, y, dy, dw
if "--prologue" in sys.argv:
    print("weighted: fprop %.3f + bn_apply %.3f = %.3f | fused %.3f ms" % (
        tot["fprop"], tot["bn_apply"], tot["fprop"] + tot["bn_apply"], tot["fused"]))
elif not profile:
    print("weighted totals (no stem): fprop %.3f dgrad %.3f wgrad %.3f | ideal per pass %.3f ms" % (
        tot["fprop"], tot["dgrad"], tot["wgrad"], ideal_tot))
