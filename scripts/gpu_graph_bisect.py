"""Find which part of the step breaks CUDA-graph capture."""
import os, sys, traceback
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from sota_imagenet_b200 import models, losses, optimizers

B, S = 32, 64
net = models.resnet50().cuda().train()
crit = losses.CrossEntropyLoss(smoothing=0.1)
opt = optimizers.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=3e-5, nesterov=True)
x = torch.randn(B, 3, S, S, device="cuda")
y = torch.randint(0, 1000, (B,), device="cuda")

def fwd():
    with torch.no_grad():
        net.ensure_arena(); return net.fwd(x, True)[0]
def fwd_loss():
    return crit(net(x), y)
def fwd_bwd():
    opt.zero_grad()
    l = crit(net(x), y); l.backward(); return l
def full():
    opt.zero_grad()
    l = crit(net(x), y); l.backward(); opt.step(); return l

for name, fn in (("fwd", fwd), ("fwd_loss", fwd_loss), ("fwd_bwd", fwd_bwd), ("full", full)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    for mode in ("global", "thread_local"):
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g, capture_error_mode=mode):
                out = fn()
            g.replay(); torch.cuda.synchronize()
            print(name, mode, "capture OK", float(out.float().sum()) if out is not None else "")
        except Exception:
            print(name, mode, "capture FAILED")
            traceback.print_exc(limit=6)
            torch.cuda.synchronize()
