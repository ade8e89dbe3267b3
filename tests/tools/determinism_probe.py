"""Two identical training steps of ResNet-26 from the same initial state; prints one SHA-1 over the
loss, every parameter gradient and the BatchNorm running statistics per run.  Under
SIB_DETERMINISTIC=1 the two digests must be equal (tests/test_gpu_model.py)."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch

from sota_imagenet_b200 import losses, models, optimizers

torch.manual_seed(0)
sd = models.resnet26(num_classes=16).state_dict()
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(16, 3, 96, 96, device="cuda", generator=g)
y = torch.randint(0, 16, (16,), device="cuda", generator=g)
for run in range(2):
    net = models.resnet26(num_classes=16)
    net.load_state_dict(sd)
    net = net.cuda().train()
    opt = optimizers.SGD(net.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4, nesterov=True)
    crit = losses.CrossEntropyLoss(smoothing=0.1)
    h = hashlib.sha1()
    for step in range(2):
        opt.zero_grad()
        loss = crit(net(x), y)
        loss.backward()
        h.update(loss.detach().cpu().numpy().tobytes())
        for p in net.parameters():
            h.update(p.grad.detach().float().cpu().numpy().tobytes())
        opt.step()
    for b in net.buffers():
        h.update(b.detach().float().cpu().numpy().tobytes())
    for p in net.parameters():
        h.update(p.detach().float().cpu().numpy().tobytes())
    print("DIGEST", run, h.hexdigest(), "loss %.6f" % float(loss))
