"""Dense-timeline determinism probe: capture forward(+loss) of ResNet-50 in a CUDA graph, replay it
many times on identical inputs/weights and report the spread of the loss and of per-layer checksums.
A missing dependency between consecutive kernels shows up as outliers far beyond atomics-order noise."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from sota_imagenet_b200 import models, losses, ops
B, S, N = int(os.environ.get("B", 32)), int(os.environ.get("S", 128)), int(os.environ.get("N", 40))
torch.manual_seed(0)
net = models.resnet50().cuda().train()
crit = losses.CrossEntropyLoss(smoothing=0.1)
x = torch.randn(B, 3, S, S, device="cuda")
y = torch.randint(0, 1000, (B,), device="cuda")
mode = sys.argv[1] if len(sys.argv) > 1 else "graph"
NAMES = ["x", "c1", "mi1", "a1", "c2", "mi2", "a2", "c3", "mi3", "cd", "mid", "out"]
def fwd():
    """returns [(name, tensor)] of every intermediate in forward order, logits last"""
    with torch.no_grad():
        logits, sv = net.fwd(x, True)
    xq, c0, mi0, cnt0, ss0, pool_saved, saved, _, _ = sv
    outs = [("stem.c0", c0), ("stem.mi0", mi0)]
    for bi, blk in enumerate(saved):
        for nm, t in zip(NAMES, blk[:12]):
            if t is not None and nm != "x":
                outs.append(("b%d.%s" % (bi, nm), t))
    outs.append(("logits", logits))
    return outs
def snap(outs):
    return torch.stack([t.float().abs().sum() for _, t in outs] + [t.float().sum() for _, t in outs])
net(x)  # builds the arena, warms up
torch.cuda.synchronize()
res = []
if mode == "graph":
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            ops.begin_pass(x.device); outs = fwd()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ops.begin_pass(x.device)
        outs = fwd()
    for i in range(N):
        g.replay()
        res.append(snap(outs))
else:
    for i in range(N):
        ops.begin_pass(x.device)
        outs = fwd()
        res.append(snap(outs))
torch.cuda.synchronize()
R = torch.stack(res).double().cpu()          # [N, 2T]
T = len(outs)
scale = R[:, :T].mean(0)                      # abs-sum as the scale of each tensor
dev = ((R - R[0:1]).abs()[:, :T] / scale).max(0).values.tolist()
dev2 = ((R - R[0:1]).abs()[:, T:] / scale).max(0).values.tolist()
print(mode, "pool" if ops._USE_POOL else "nopool", "fused_stem" if models.FUSE_STEM_POOL else "plain_stem")
line = []
for (nm, _), d, d2 in zip(outs, dev, dev2):
    line.append("%s:%.0e/%.0e" % (nm, d, d2))
print(" ".join(line))
