"""Run under torchrun with N ranks (2, 4 or 8): DP(N x B) with SyncBN + bucketed all-reduce must equal
one process on the concatenated N*B batch -- per parameter: gradient cosine >= 0.999 AND gradient-norm
ratio within 1 % (a SUM-instead-of-AVG reduction, a missed bucket or unsynchronised statistics all
fail it), BN running statistics, the loss; plus one Bottleneck under SyncBN against the same block on
the global batch.  The network runs in the damped regime of tests/test_gpu_model.py
(bn3.weight = 0.1) so that bf16 summation-order noise is not chaotically amplified and the gates can
be strict."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch, torch.distributed as dist
from oracle import torch_ref
from sota_imagenet_b200 import losses, models, modules, ops, parallel

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
import datetime
dist.init_process_group("nccl", timeout=datetime.timedelta(seconds=120))
B, S = 8, 128
x, y = torch_ref.synthetic_batch(B * world, S, seed=5)
sd = torch_ref.resnet50(seed=0).state_dict()
g0 = torch.Generator().manual_seed(9)
for k, v in sd.items():
    if k.endswith("bn3.weight"):
        v.fill_(0.1)
    elif k.endswith(".weight") and v.dim() == 1:
        v.copy_(torch.rand(v.shape, generator=g0) * 0.6 + 0.7)
crit = losses.CrossEntropyLoss(smoothing=0.1)


def compare(g_dp, named_big, only=None):
    """-> (worst cosine, worst |norm ratio - 1|, name of the worst) over the parameters
    (only = "w": conv / fc weight tensors, "v": BatchNorm / bias vectors)."""
    worst_c, worst_n, who = 1.0, 0.0, ""
    for n, p in named_big:
        if only is not None and (p.dim() > 1) != (only == "w"):
            continue
        a, b = g_dp[n].double().flatten(), p.grad.double().flatten()
        c = float(a @ b / (a.norm() * b.norm() + 1e-30))
        r = abs(float(a.norm() / (b.norm() + 1e-30)) - 1.0)
        if c < worst_c or r > worst_n:
            who = n
        worst_c, worst_n = min(worst_c, c), max(worst_n, r)
    return worst_c, worst_n, who


ok = True
# ---- one Bottleneck: SyncBN over N ranks == BatchNorm over the global batch -------------------
torch.manual_seed(3)
blk_sd = modules.Bottleneck(256, 64).state_dict()
xb_all = torch.relu(torch.randn(4 * world, 256, 28, 28)).bfloat16()
dy_all = torch.randn(4 * world, 256, 28, 28).bfloat16()
blk = modules.Bottleneck(256, 64); blk.load_state_dict(blk_sd); blk = blk.cuda().train()
dpb = parallel.DataParallel(blk, sync_bn=True)
xl = ops.to_nhwc_bf16(xb_all[rank * 4:(rank + 1) * 4].cuda()).requires_grad_(True)
out = dpb(xl)
# DataParallel averages parameter gradients over ranks: feed dy * world so that the average equals the
# global-batch gradient of sum(out * dy)
out.backward(ops.to_nhwc_bf16((dy_all[rank * 4:(rank + 1) * 4].float() * world).cuda()))
torch.cuda.synchronize()
gb = {n: p.grad.detach().float().clone() for n, p in blk.named_parameters()}
rb = {n: b.clone() for n, b in blk.named_buffers() if "running" in n}
if rank == 0:
    big = modules.Bottleneck(256, 64); big.load_state_dict(blk_sd); big = big.cuda().train()
    xa = ops.to_nhwc_bf16(xb_all.cuda()).requires_grad_(True)
    oa = big(xa)
    oa.backward(ops.to_nhwc_bf16(dy_all.cuda()))
    torch.cuda.synchronize()
    c, r, who = compare(gb, list(big.named_parameters()))
    o_err = float((out.float() - oa[:4].float()).norm() / oa[:4].float().norm())
    dx_err = float((xl.grad.float() / world - xa.grad[:4].float()).norm() / xa.grad[:4].float().norm())
    s_err = max(float((rb[n] - b).norm() / (b.norm() + 1e-12)) for n, b in big.named_buffers() if "running" in n)
    print("bottleneck SyncBN x%d vs global batch: out err %.2e dx err %.2e | worst grad cosine %.6f, norm dev %.2e (%s) | "
          "running-stat err %.2e" % (world, o_err, dx_err, c, r, who, s_err))
    # (measured over N = 2, 4, 8 and several boxes: out 1e-5..3e-4, dx 6e-4..8e-3 -- bf16 rounding flips caused by the
    #  different summation order of the statistics; a wrong reduction moves the NORMS by a factor, not by a percent)
    ok = ok and o_err < 5e-3 and dx_err < 2e-2 and c >= 0.999 and r < 1e-2 and s_err < 1e-4
if ops.PEER is not None:
    torch.cuda.synchronize(); ops.PEER.reset_layout()
dist.barrier()

# ---- whole ResNet-50, two regimes (see tests/test_gpu_model.py) ------------------------------------
#   leaky 0.8: strict gates (weights cosine >= 0.999, norm 1 %; BN / bias vectors >= 0.97, 5 %: the stem's
#              bn1.bias, a sum over all 3.2 M pixels with heavy cancellation, measured 0.9797 .. 0.9897 over
#              N = 2, 4, 8 and several boxes, i.e. the former 0.98 sat inside the run-to-run spread);
#   relu     : the production activation; a flipped mask is a 100 % change of that element, so the
#              cosine gates are looser, the NORM gates (what a SUM-for-AVG would break) stay tight.
for slope, gates in ((0.8, (0.999, 1e-2, 0.97, 5e-2)), (0.0, (0.95, 2e-2, 0.93, 1e-1))):
    def make():
        if slope > 0:
            m = models.resnet50(norm_act="leaky_relu")
            for q in m.modules():
                if isinstance(q, modules.BatchNorm2d) and q.activation == "leaky_relu":
                    q.slope = slope
        else:
            m = models.resnet50()
        m.load_state_dict(sd)
        return m.cuda().train()
    net = make()
    dp = parallel.DataParallel(net, sync_bn=True, bucket_mb=8.0)
    xs, ys = x[rank * B:(rank + 1) * B].cuda(), y[rank * B:(rank + 1) * B].cuda()
    loss = crit(dp(xs), ys); loss.backward(); torch.cuda.synchronize()
    g_dp = {n: p.grad.detach().float().clone() for n, p in net.named_parameters()}
    bufs_dp = {n: b.clone() for n, b in net.named_buffers() if "running" in n}
    ltot = loss.detach().clone(); dist.all_reduce(ltot); ltot /= world
    if rank == 0:
        big = make()
        lb = crit(big(x.cuda()), y.cuda()); lb.backward(); torch.cuda.synchronize()
        named = list(big.named_parameters())
        wc, wn, wwho = compare(g_dp, named, "w")
        vc, vn, vwho = compare(g_dp, named, "v")
        def buf_err(prefixes):
            return max(float((bufs_dp[n] - b).norm() / (b.norm() + 1e-12)) for n, b in big.named_buffers()
                       if "running" in n and n.startswith(prefixes))
        early, late = buf_err(("bn1.", "layer1.0.")), buf_err(("layer4.",))
        print("resnet50 DP x%d vs global batch, %s: loss dp %.5f big %.5f | 54 weight tensors: worst cosine %.6f, worst norm "
              "deviation %.2e (%s) | 107 BN/bias vectors: %.6f, %.2e (%s) | running-stat err early %.2e late %.2e"
              % (world, "leaky_relu(%.1f)" % slope if slope else "relu", ltot.item(), lb.item(), wc, wn, wwho, vc, vn, vwho,
                 early, late))
        ok = ok and abs(ltot.item() - lb.item()) / lb.item() < 2e-3 and early < 2e-3 and late < 2e-2 \
            and wc >= gates[0] and wn < gates[1] and vc >= gates[2] and vn < gates[3]
    if ops.PEER is not None:
        torch.cuda.synchronize(); ops.PEER.reset_layout()
    dist.barrier()
# ---- the peer-memory one-shot all-reduce itself: == NCCL, bitwise identical across ranks, eager and
#      replayed from a CUDA graph (epochs / parity buffers keep working across replays)
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
