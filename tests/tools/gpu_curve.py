"""200-step loss curves: fp32 CPU oracle vs stock autocast vs ours (prints every 10th step)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np, torch
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from oracle import torch_ref
from sota_imagenet_b200 import models, losses, optimizers
lr = float(sys.argv[1]) if len(sys.argv) > 1 else 0.01
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pool_x, pool_y = torch_ref.synthetic_batch(64, 64, seed=3)
px, py = pool_x.cuda(), pool_y.cuda()
ref = torch_ref.resnet50(seed=0).train()
opt_ref = torch_ref.make_sgd(ref.parameters(), lr=lr, nesterov=True)
c32 = []
for step in range(200):
    lo = (step * 16) % 64
    c32.append(torch_ref.train_step(ref, opt_ref, pool_x[lo:lo + 16], pool_y[lo:lo + 16]))
curves = {"fp32": c32}
for r in range(runs):
    net = models.resnet50(); net.load_state_dict(torch_ref.resnet50(seed=0).state_dict()); net = net.cuda().train()
    opt = optimizers.SGD(net.parameters(), lr=lr, momentum=0.9, weight_decay=3e-5, nesterov=True)
    crit = losses.CrossEntropyLoss(smoothing=0.1)
    c = []
    for step in range(200):
        lo = (step * 16) % 64
        opt.zero_grad(); loss = crit(net(px[lo:lo + 16]), py[lo:lo + 16]); loss.backward(); opt.step()
        c.append(loss.item())
    curves["ours%d" % r] = c
amp = torch_ref.resnet50(seed=0).cuda().train()
opt_amp = torch_ref.make_sgd(amp.parameters(), lr=lr, nesterov=True)
c = []
for step in range(200):
    lo = (step * 16) % 64
    opt_amp.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = amp(px[lo:lo + 16])
    l = torch_ref.smooth_cross_entropy(out, py[lo:lo + 16], 0.1); l.backward(); opt_amp.step(); c.append(float(l.detach()))
curves["autocast"] = c
names = list(curves)
print("step " + " ".join("%9s" % n for n in names))
for s in range(0, 200, 8):
    print("%4d " % s + " ".join("%9.4f" % curves[n][s] for n in names))
