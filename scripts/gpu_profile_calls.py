"""Per-C-ABI-call CUDA-event time of one training step (wgrad serialised on the main stream), for
any model config:  python scripts/gpu_profile_calls.py bresnet|r50 [size]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from sota_imagenet_b200 import _lib, losses, models, ops, optimizers

which = sys.argv[1] if len(sys.argv) > 1 else "bresnet"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 224
B = 256
if which == "bresnet":
    net = models.resnet50(stem_type="deep", antialias=True, attn_type="eca", norm_layer="inplaceabn",
                          norm_act="leaky_relu", drop_rate=0.2, drop_connect_rate=0.2,
                          weight_standardization=True)
else:
    net = models.resnet50()
net = net.cuda().train()
crit = losses.CrossEntropyLoss(smoothing=0.1)
opt = optimizers.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=3e-5, nesterov=True)
x = torch.zeros(B, size, size, 4, device="cuda", dtype=torch.bfloat16)
x[..., :3] = torch.randn(B, size, size, 3, device="cuda")
x = x.permute(0, 3, 1, 2)
y = torch.randint(0, 1000, (B,), device="cuda")


def step():
    opt.zero_grad()
    loss = crit(net(x), y)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
ops._SideStream.enabled = False
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_lib.PROFILE = []
e0.record()
step()
e1.record()
torch.cuda.synchronize()
prof, _lib.PROFILE = _lib.PROFILE, None
groups = {}
for name, a, s0, s1 in prof:
    d = groups.setdefault(name, [0.0, 0])
    d[0] += s0.elapsed_time(s1)
    d[1] += 1
tot = sum(v[0] for v in groups.values())
print("%s %dx%d batch %d: step %.3f ms (events around the whole step), C-ABI calls %.3f ms in %d calls"
      % (which, size, size, B, e0.elapsed_time(e1), tot, len(prof)))
for name, (ms, n) in sorted(groups.items(), key=lambda kv: -kv[1][0]):
    print("%-34s n=%4d %8.3f ms %5.1f%%" % (name, n, ms, 100 * ms / tot))
if "--shapes" in sys.argv:
    # per (entry point, shape) table of the convolution calls: the small integer arguments identify the layer
    shapes = {}
    for name, a, s0, s1 in prof:
        if "conv2d" not in name:
            continue
        key = (name,) + tuple(v for v in a if isinstance(v, int) and 0 <= v < 100000)[:12]
        d = shapes.setdefault(key, [0.0, 0])
        d[0] += s0.elapsed_time(s1)
        d[1] += 1
    print("conv calls by shape (N, H, W, C, K, R, S, stride, pad_h, pad_w, OH, OW ...):")
    for key, (ms, n) in sorted(shapes.items(), key=lambda kv: -kv[1][0])[:40]:
        print("%-26s %-52s n=%3d %7.3f ms  %.3f ms/call" % (key[0], key[1:], n, ms, ms / n))
