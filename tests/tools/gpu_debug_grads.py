"""Layer-by-layer comparison of block outputs and block-output gradients: ours vs the
bf16-faithful oracle (and vs pure fp32)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from oracle import torch_ref
from oracle.torch_ref import _q, _conv_q, _bn, _RoundFwd
from sota_imagenet_b200 import models, losses, ops

def cos(a, b):
    a = a.double().flatten().cpu(); b = b.double().flatten().cpu()
    return float(a @ b / (a.norm() * b.norm() + 1e-30))
def rel(a, b):
    a = a.double().flatten().cpu(); b = b.double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))

B, S = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 128
ref = torch_ref.resnet50(seed=0)
net = models.resnet50(); net.load_state_dict(ref.state_dict()); net = net.cuda().train()
faithful = torch_ref.resnet50(seed=0).cuda().train()
x, y = torch_ref.synthetic_batch(B, S, seed=0)
x, y = x.cuda(), y.cuda()

# oracle forward with retained block outputs
acts = []
model = faithful
xx = _RoundFwd.apply(x)
t = _q(F.relu(_bn(model.bn1, _conv_q(model.conv1, xx)))); t.retain_grad(); acts.append(("stem", t))
t = F.max_pool2d(t, 3, 2, 1); t.retain_grad(); acts.append(("pool", t))
for li, layer in enumerate((model.layer1, model.layer2, model.layer3, model.layer4)):
    for bi, blk in enumerate(layer):
        idt = t
        o = _q(F.relu(_bn(blk.bn1, _conv_q(blk.conv1, t))))
        o = _q(F.relu(_bn(blk.bn2, _conv_q(blk.conv2, o))))
        o = _bn(blk.bn3, _conv_q(blk.conv3, o))
        if blk.downsample is not None:
            idt = _bn(blk.downsample[1], _conv_q(blk.downsample[0], t))
        t = _q(F.relu(o + idt)); t.retain_grad(); acts.append(("layer%d.%d" % (li + 1, bi), t))
feat = _q(t.mean(dim=(2, 3)))
logits = _q(F.linear(feat, _RoundFwd.apply(model.fc.weight), model.fc.bias)); logits.retain_grad()
loss_ref = torch_ref.smooth_cross_entropy(logits, y, 0.1); loss_ref.backward()

# ours, with hooks
mine_acts, mine_grads = {}, {}
blocks = list(net.blocks())
names = ["layer%d.%d" % (li + 1, bi) for li in range(4) for bi in range(len(getattr(net, "layer%d" % (li + 1))))]
for nm, blk in zip(names, blocks):
    def wrap(blk=blk, nm=nm):
        f0, b0 = blk.fwd, blk.bwd
        def fwd(xi, train):
            out, s = f0(xi, train); mine_acts[nm] = out; return out, s
        def bwd(dy, s, need_dx=True):
            mine_grads[nm] = dy; return b0(dy, s, need_dx=need_dx)
        blk.fwd, blk.bwd = fwd, bwd
    wrap()
crit = losses.CrossEntropyLoss(smoothing=0.1)
out = net(x); out.retain_grad()
loss = crit(out, y); loss.backward(); torch.cuda.synchronize()
print("loss ours %.5f faithful %.5f" % (loss.item(), loss_ref.item()))
print("logits rel %.2e  dlogits cos %.6f" % (rel(out, logits), cos(out.grad, logits.grad)))
for nm, t in acts[::-1]:
    if nm in mine_acts:
        print("%-10s act rel %.2e   grad cos %.6f rel %.2e" % (nm, rel(mine_acts[nm], t), cos(mine_grads[nm], t.grad), rel(mine_grads[nm], t.grad)))
rp = dict(faithful.named_parameters())
bad = sorted(((cos(p.grad.reshape(rp[n].shape), rp[n].grad), n) for n, p in net.named_parameters()))
print("worst param cosines:", bad[:6])
print("fc.weight cos", cos(net.fc.weight.grad.reshape(rp["fc.weight"].shape), rp["fc.weight"].grad))
