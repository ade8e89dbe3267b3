"""Run under torchrun with 2 ranks: DP(2 x B) with SyncBN + bucketed all-reduce must equal one
process on the concatenated 2B batch (gradients after one step, BN running statistics)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
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
# ---- the peer-memory one-shot all-reduce itself: == NCCL, bitwise identical across ranks, eager and
#      replayed from a CUDA graph (epochs / parity buffers keep working across replays)
from sota_imagenet_b200 import ops
if ops.PEER is not None:
    torch.cuda.synchronize(); ops.PEER.reset_layout(); dist.barrier()
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    sizes = [128, 4096, 8192, 512, 16]
    vecs = [torch.randn(n, device="cuda", generator=g) for n in sizes]
    def run_all(out):
        ops.PEER.begin_forward()
        for v, o in zip(vecs, out):
            o.copy_(v); ops.PEER.allreduce_(o)
    outs = [torch.empty_like(v) for v in vecs]
    for it in range(5):
        run_all(outs)
    refs = [v.clone() for v in vecs]
    for r_ in refs: dist.all_reduce(r_)
    torch.cuda.synchronize()
    peer_ok = all(float((o - r_).abs().max()) <= 1e-5 * float(r_.abs().max()) for o, r_ in zip(outs, refs))
    s_ = torch.cuda.Stream(); s_.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s_):
        run_all(outs)
    torch.cuda.current_stream().wait_stream(s_)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        run_all(outs)
    for it in range(7):
        graph.replay()
    torch.cuda.synchronize()
    peer_ok = peer_ok and all(float((o - r_).abs().max()) <= 1e-5 * float(r_.abs().max()) for o, r_ in zip(outs, refs))
    same = torch.stack([o.double().sum() for o in outs])
    lo_, hi_ = same.clone(), same.clone()
    dist.all_reduce(lo_, op=dist.ReduceOp.MIN); dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
    peer_ok = peer_ok and bool((lo_ == hi_).all())
    if rank == 0:
        print("peer all-reduce: %s (%d calls eager + graph)" % ("OK" if peer_ok else "MISMATCH", 13 * len(sizes)))
    ok = ok and peer_ok
elif rank == 0:
    print("peer all-reduce: not active (NCCL path)")
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
