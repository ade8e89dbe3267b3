"""Quick end-to-end probe: build ResNet-50, run steps, print timing (eager, no graph)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from sota_imagenet_b200 import models, losses, optimizers

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = int(sys.argv[2]) if len(sys.argv) > 2 else 224
net = models.resnet50().cuda()
crit = losses.CrossEntropyLoss(smoothing=0.1)
opt = optimizers.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=3e-5, nesterov=True)
x = torch.randn(B, 3, S, S, device="cuda")
y = torch.randint(0, 1000, (B,), device="cuda")
def step():
    opt.zero_grad()
    loss = crit(net(x), y)
    loss.backward()
    opt.step()
    return loss
for i in range(3):
    l = step()
torch.cuda.synchronize()
print("warm loss", l.item(), "mem GB", torch.cuda.max_memory_allocated() / 1e9)
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
t0 = time.time(); e0.record()
for i in range(10):
    l = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("eager: %.2f ms/step  %.0f img/s (wall %.2f ms) loss %.4f" % (ms, B / ms * 1e3, (time.time() - t0) * 100, l.item()))
# fwd only / bwd only split
torch.cuda.synchronize(); e0.record()
for i in range(10):
    with torch.no_grad():
        net.train(); out, saved = net.fwd(x, True)
e1.record(); torch.cuda.synchronize()
print("fwd only: %.2f ms" % (e0.elapsed_time(e1) / 10))
# CUDA graph of the whole step
g = torch.cuda.CUDAGraph()
opt.zero_grad()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for i in range(2):
        step()
torch.cuda.current_stream().wait_stream(s)
try:
    with torch.cuda.graph(g):
        lg = step()
    for i in range(3):
        g.replay()
    torch.cuda.synchronize(); e0.record()
    for i in range(10):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("graph: %.2f ms/step  %.0f img/s loss %.4f" % (ms, B / ms * 1e3, lg.item()))
except Exception as e:
    print("graph capture failed:", repr(e)[:400])
