"""Achieved HBM bandwidth of the BatchNorm-family kernels on the large ResNet-50 tensors."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from sota_imagenet_b200 import ops
B = 256
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
relu = ops.ACT_CODES["relu"]
for (c, h) in [(64, 112), (64, 56), (256, 56), (128, 28), (512, 28), (1024, 14), (2048, 7)]:
    x = ops.to_nhwc_bf16(torch.randn(B, c, h, h, device="cuda"))
    dy = ops.to_nhwc_bf16(torch.randn(B, c, h, h, device="cuda"))
    out = ops.to_nhwc_bf16(torch.relu(torch.randn(B, c, h, h, device="cuda")))
    e = B * c * h * h
    stats = ops.bn_stats(x)
    g, b = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    y, (mi, ss), _ = ops.bn_finalize_apply(x, (stats.clone(), g, b, rm, rv), act=relu, slope=0.0, count=B * h * h)
    rows = []
    t = timeit(lambda: ops.bn_finalize_apply(x, (stats, g, b, rm, rv), act=relu, slope=0.0, count=B * h * h))
    rows.append(("finalize_apply", t, 4 * e))
    t = timeit(lambda: ops.bn_finalize_apply(x, (stats, g, b, rm, rv), res=out, act=relu, slope=0.0, count=B * h * h))
    rows.append(("finalize_apply+res", t, 6 * e))
    t = timeit(lambda: ops.bn_bwd_reduce(dy, None, x, mi, relu, 0.0, mask_ss=ss))
    rows.append(("bwd_reduce(mask_ss)", t, 4 * e))
    t = timeit(lambda: ops.bn_bwd_reduce(dy, out, x, mi, relu, 0.0))
    rows.append(("bwd_reduce(out)", t, 6 * e))
    sums = ops.bn_bwd_reduce(dy, None, x, mi, relu, 0.0, mask_ss=ss)
    t = timeit(lambda: ops.bn_bwd_apply(dy, None, x, mi, g, sums, B * h * h, ops.ACT_NONE, 0.0))
    rows.append(("bwd_apply(premasked)", t, 6 * e))
    t = timeit(lambda: ops.bn_bwd_apply(dy, out, x, mi, g, sums, B * h * h, relu, 0.0, want_g=True))
    rows.append(("bwd_apply(out,+g)", t, 10 * e))
    print("C=%4d H=%3d  " % (c, h) + "  ".join("%s %.3f ms %4.0f GB/s" % (n, t, by / t / 1e6) for n, t, by in rows))
