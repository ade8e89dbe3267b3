"""Which host-side detail decides whether the whole training step captures into a CUDA graph."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from sota_imagenet_b200 import losses, models, optimizers

variant = int(sys.argv[1])
if variant & 16:
    import numpy as np
    from sota_imagenet_b200 import runner
if variant & 2:
    torch.cuda.set_device(0)
if not variant & 32:
    torch.manual_seed(0)
net = models.resnet50().cuda().train()
crit = losses.CrossEntropyLoss(smoothing=0.1)
opt = optimizers.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=3e-5, nesterov=True)
B, S = int(os.environ.get("PB", 64)), int(os.environ.get("PS", 128))
if variant & 32:
    torch.manual_seed(0)
x = torch.zeros(B, S, S, 4, device="cuda", dtype=torch.bfloat16)
x[..., :3] = torch.randn(B, S, S, 3, device="cuda")
x = x.permute(0, 3, 1, 2)
y = torch.randint(0, 1000, (B,), device="cuda")


def step():
    opt.zero_grad()
    loss = crit(net(x), y)
    loss.backward()
    opt.step()
    return loss


if variant & 4:
    def timed():
        return step()
else:
    timed = step
for _ in range(2):
    if variant & 64:
        held = step()          # keeps the previous step's loss (and its autograd nodes) alive
    else:
        step()
torch.cuda.synchronize()
if variant & 8:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    step()
torch.cuda.current_stream().wait_stream(side)
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        if variant & 1:
            keep = step()
        else:
            step()
    g.replay()
    torch.cuda.synchronize()
    print("variant", variant, "capture OK")
except Exception as e:
    print("variant", variant, "capture FAILED:", repr(e)[:120])
