"""Run under torchrun with 2 ranks: DP(2 x B) with SyncBN + bucketed all-reduce must equal one
process on the concatenated 2B batch (gradients after one step, BN running statistics)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, torch.distributed as dist
from oracle import torch_ref
from sota_imagenet_b200 import losses, models, parallel

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
import datetime
dist.init_process_group("nccl", timeout=datetime.timedelta(seconds=120))
B, S = 8, 128
x, y = torch_ref.synthetic_batch(B * world, S, seed=5)
sd = torch_ref.resnet50(seed=0).state_dict()
crit = losses.CrossEntropyLoss(smoothing=0.1)

net = models.resnet50(); net.load_state_dict(sd); net = net.cuda().train()
dp = parallel.DataParallel(net, sync_bn=True, bucket_mb=8.0)
xs, ys = x[rank * B:(rank + 1) * B].cuda(), y[rank * B:(rank + 1) * B].cuda()
loss = crit(dp(xs), ys); loss.backward(); torch.cuda.synchronize()
g_dp = {n: p.grad.detach().float().clone() for n, p in net.named_parameters()}
bufs_dp = {n: b.clone() for n, b in net.named_buffers() if "running" in n}
ltot = loss.detach().clone(); dist.all_reduce(ltot); ltot /= world

ok = True
if rank == 0:
    big = models.resnet50(); big.load_state_dict(sd); big = big.cuda().train()
    lb = crit(big(x.cuda()), y.cuda()); lb.backward(); torch.cuda.synchronize()
    worst = 1.0
    for n, p in big.named_parameters():
        a, b = g_dp[n].double().flatten(), p.grad.double().flatten()
        c = float(a @ b / (a.norm() * b.norm() + 1e-30))
        worst = min(worst, c)
    def buf_err(prefixes):
        return max(float((bufs_dp[n] - b).norm() / (b.norm() + 1e-12)) for n, b in big.named_buffers()
                   if "running" in n and n.startswith(prefixes))
    early, late = buf_err(("bn1.", "layer1.0.")), buf_err(("layer4.",))
    fc = float((g_dp["fc.weight"].flatten() @ big.fc.weight.grad.float().flatten()) /
               (g_dp["fc.weight"].norm() * big.fc.weight.grad.float().norm()))
    print("loss dp %.5f big %.5f | worst grad cosine %.4f fc.weight cosine %.5f | running-stat err early %.2e late %.2e"
          % (ltot.item(), lb.item(), worst, fc, early, late))
    # SyncBN statistics of the first layers match tightly; deeper quantities inherit the bf16
    # chaos described in DESIGN.md (the two runs differ only in fp32 summation order, which is
    # enough to flip bf16 roundings), so they get loose gates.
    ok = abs(ltot.item() - lb.item()) / lb.item() < 1e-2 and early < 2e-3 and late < 0.15 and fc > 0.95
# every rank holds identical averaged gradients
chk = torch.stack([g.sum() for g in g_dp.values()]).sum()
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
ok = ok and bool(lo == hi)
flag = torch.tensor([1.0 if ok else 0.0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0 and flag.item() == 1.0:
    print("DP_EQUIVALENCE_OK")
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)
