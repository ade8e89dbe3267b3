"""Run one conv configuration a few times (for ncu captures).
usage: gpu_one_conv.py <fprop|dgrad|wgrad> C K H R stride pad [B]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from sota_imagenet_b200 import ops
mode, c, k, h, r, stride, pad = sys.argv[1], *[int(v) for v in sys.argv[2:8]]
B = int(sys.argv[8]) if len(sys.argv) > 8 else 256
x = ops.to_nhwc_bf16(torch.randn(B, c, h, h, device="cuda"))
w = (torch.randn(k, c, r, r, device="cuda") / (c * r * r) ** 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
wd = ops.pack_dgrad_weight(w)
stats = torch.empty(2, k, device="cuda")
y = ops.conv2d_fprop(x, w, stride=stride, pad=pad, stats=stats)
dy = torch.randn_like(y)
dw = torch.zeros(k, r, r, c, device="cuda").permute(0, 3, 1, 2)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(4):
    flush.zero_()
    if mode == "fprop":
        ops.conv2d_fprop(x, w, stride=stride, pad=pad, stats=stats)
    elif mode == "dgrad":
        ops.conv2d_dgrad(dy, wd, tuple(x.shape), r, r, stride=stride, pad=pad)
    else:
        ops.conv2d_wgrad(x, dy, dw, stride=stride, pad=pad)
torch.cuda.synchronize()
print("done")
