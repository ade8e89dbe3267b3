"""CUDA-event time of the sphere-linear head (BASELINE config 5: 256 x 512 -> 1000 classes) through the
C-ABI: forward (row norms + tensor-core cosine GEMM), margin + smooth CE, backward (d(xn), d(wn) GEMMs
in one launch + the normalisation Jacobians).   gpurun -- python scripts/gpu_head_timing.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sota_imagenet_b200 import ops  # noqa: E402


def timed(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    b, c, d = 256, 1000, 512
    x = torch.randn(b, d, device="cuda")
    w = torch.randn(c, d, device="cuda") * 0.1
    y = torch.randint(0, c, (b,), device="cuda")
    cosv, saved = ops.sphere_linear_fwd(x, w)
    _, _, dcos = ops.ce_fwd_bwd(cosv, y, smoothing=0.1, margin_kind=ops.MARGIN_ARC, s=10.0, m=0.2)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cosv2, saved2 = ops.sphere_linear_fwd(x, w)
        _, _, dcos2 = ops.ce_fwd_bwd(cosv2, y, smoothing=0.1, margin_kind=ops.MARGIN_ARC, s=10.0, m=0.2)
        ops.sphere_linear_bwd(dcos2, saved2)
    print("sphere_linear_fwd       %.4f ms (eager, incl. C-ABI call overhead)" % timed(lambda: ops.sphere_linear_fwd(x, w)))
    print("sphere_linear_bwd       %.4f ms (eager)" % timed(lambda: ops.sphere_linear_bwd(dcos, saved)))
    print("fwd + arcface CE + bwd  %.4f ms (one CUDA-graph replay: device time of the whole head)" % timed(g.replay))
    ref = torch.nn.functional.normalize(x.double(), dim=1) @ torch.nn.functional.normalize(w.double(), dim=1).t()
    print("max |cos - float64|     %.3e" % float((cosv.double() - ref).abs().max()))


if __name__ == "__main__":
    main()
