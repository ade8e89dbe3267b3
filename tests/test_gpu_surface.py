"""Drop-in surface on the GPU: CModel graphs of fused modules, the angular-margin model +
criterion chain (BASELINE config #5), EMA, gradient accumulation, checkpoint / resume, weight-decay
filtering — the pieces reference train.py wires together (train.py:64-152)."""
import os

import pytest
import torch

from oracle import torch_ref

pytestmark = pytest.mark.gpu


def test_cmodel_of_fused_modules_trains():
    from sota_imagenet_b200 import cmodel, losses, optimizers
    net = cmodel.CModel([
        dict(module="StemConv", args=[64, 7, 3]),
        dict(module="BatchNorm2d", args=[64], kwargs=dict(activation="'relu'")),
        dict(module="MaxPool3x3s2", tag="pool"),
        dict(module="Bottleneck", args=[64, 64], kwargs=dict(downsample=True)),
        dict(module="Bottleneck", args=[256, 64], repeat=2, tag="b"),
        dict(module="Conv2d", args=[256, 64, 1]),
        dict(module="BatchNorm2d", args=[64], kwargs=dict(activation="'relu'"), tag="side"),
        dict(module="Concat", inputs=["side", "pool"]),
        dict(module="Conv2d", args=[128, 256, 3], kwargs=dict(padding=1)),
        dict(module="GlobalAvgPool"),
        dict(module="Linear", args=[256, 16]),
    ]).cuda()
    opt = optimizers.SGD(net.parameters(), lr=0.005, momentum=0.9)
    crit = losses.CrossEntropyLoss(smoothing=0.1)
    x = torch.randn(8, 3, 64, 64, device="cuda")
    y = torch.randint(0, 16, (8,), device="cuda")
    first = None
    for _ in range(30):
        opt.zero_grad()
        loss = crit(net(x), y)
        loss.backward()
        opt.step()
        first = first if first is not None else loss.item()
    assert torch.isfinite(loss) and loss.item() < 0.8 * first, (first, loss.item())


@pytest.mark.parametrize("cfg_name", ["r50_arcface.yaml", "r50_cosface.yaml"])
def test_angular_margin_config_step(cfg_name):
    """ResNet-50 -> 512-d embedding -> SphereLinearLayer -> ArcFace / CosFace + smoothing, against
    the torch restatement on the same embedding network (torchvision trunk with a 512-wide fc)."""
    from sota_imagenet_b200 import config, optimizers
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = config.load_config(os.path.join(root, "configs", cfg_name))
    model = config.call(cfg.model)
    ref = torch_ref.resnet50(num_classes=512, seed=0)
    model.encoder.load_state_dict(ref.state_dict())
    model = model.cuda().train()
    crit = config.call(cfg.criterion).cuda()
    w = model.head.weight.detach().cpu().clone()
    x, y = torch_ref.synthetic_batch(8, 128, seed=4)
    ref.train()
    cos = torch_ref.sphere_linear(ref(x), w)
    logits = torch_ref.arcface_logits(cos, y, 10.0, 0.2) if "arcface" in cfg_name else torch_ref.cosface_logits(cos, y, 10.0, 0.2)
    loss_ref = torch_ref.smooth_cross_entropy(logits, y, 0.1)
    params = [{"params": list(model.parameters()) + list(crit.parameters())}]
    opt = config.call(cfg.optim, params)
    assert isinstance(opt, optimizers.SGD)
    opt.zero_grad()
    loss = crit(model(x.cuda()), y.cuda())
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) / loss_ref.item() < 1e-2, (loss.item(), loss_ref.item())
    assert model.head.weight.grad is not None and torch.isfinite(model.head.weight.grad).all()
    before = model.head.weight.detach().clone()
    opt.param_groups[0]["lr"] = 0.1
    opt.step()
    assert not torch.equal(before, model.head.weight.detach())          # head weight is optimised too


def test_ema_accumulation_checkpoint_resume(tmp_path):
    from sota_imagenet_b200 import losses, models, optimizers, runner
    torch.manual_seed(0)
    net = models.resnet26(num_classes=16).cuda().train()
    groups = runner.filter_from_weight_decay(net, skip_list=("bn", "bias"))
    assert len(groups) == 2 and groups[1]["weight_decay"] == 0.0
    opt = optimizers.SGD(groups, lr=0.02, momentum=0.9, weight_decay=1e-4, ema_decay=0.5)
    crit = losses.CrossEntropyLoss()
    x = torch.randn(8, 3, 64, 64, device="cuda")
    y = torch.randint(0, 16, (8,), device="cuda")
    # gradient accumulation: two half-batches (scaled 1/2) == ... at least accumulates, not overwrites
    opt.zero_grad()
    (crit(net(x[:4]), y[:4]) / 2).backward()
    g1 = net.fc.weight.grad.detach().clone()
    (crit(net(x[4:]), y[4:]) / 2).backward()
    assert not torch.equal(g1, net.fc.weight.grad)
    opt.step()
    p1 = net.fc.weight.detach().clone()
    opt.zero_grad()
    crit(net(x), y).backward()
    opt.step()
    ema = opt.ema_state_dict(net)["fc.weight"]
    # ema after 2 steps with decay .5 starting from the initial weights: lies between p0 and p2
    assert torch.isfinite(ema).all() and not torch.equal(ema, net.fc.weight.detach())
    # checkpoint {state_dict, epoch, optimizer} -> resume (reference train.py:98-109)
    path = os.path.join(str(tmp_path), "model.chpn")
    torch.save({"state_dict": net.state_dict(), "epoch": 3, "optimizer": opt.state_dict()}, path)
    ck = torch.load(path, map_location="cuda", weights_only=False)
    net2 = models.resnet26(num_classes=16).cuda().train()
    net2.load_state_dict(ck["state_dict"], strict=False)
    opt2 = optimizers.SGD(runner.filter_from_weight_decay(net2, skip_list=("bn", "bias")), lr=0.02, momentum=0.9,
                          weight_decay=1e-4)
    opt2.load_state_dict(ck["optimizer"])
    for o, n in ((opt, net), (opt2, net2)):
        o.zero_grad()
        crit(n(x), y).backward()
        o.step()
    err = (net.fc.weight - net2.fc.weight).abs().max().item()
    assert err < 5e-3, err          # same weights + same momentum -> same next step (up to bf16 chaos)
    assert not torch.equal(p1, net.fc.weight.detach())


def test_model_ema_callback_swaps_weights_for_validation_and_checkpoint(tmp_path):
    """Reference train.py:112,138: ModelEma placed after CheckpointSaver -- validation and the
    epoch's checkpoint see the EMA weights, training continues on the raw ones.  Both the fused
    (optimizer-maintained) and the stand-alone average are exercised."""
    from sota_imagenet_b200 import losses, models, optimizers, runner

    class Loader:
        batch_size = 8

        def __init__(self, n):
            g = torch.Generator(device="cuda").manual_seed(0)
            self.b = [(torch.randn(8, 3, 64, 64, device="cuda", generator=g),
                       torch.randint(0, 16, (8,), device="cuda", generator=g)) for _ in range(n)]

        def __len__(self):
            return len(self.b)

        def __iter__(self):
            return iter(self.b)

    for fused in (True, False):
        torch.manual_seed(0)
        net = models.resnet26(num_classes=16).cuda().train()
        net.ensure_arena()
        w0 = net.fc.weight.detach().clone()
        opt = optimizers.SGD(net.parameters(), lr=0.05, momentum=0.9, ema_decay=0.5 if fused else 0.0)
        seen = {}

        class Spy(runner.Callback):
            def on_loader_end(self):
                seen["val" if not self.state.is_train else "train"] = net.fc.weight.detach().clone()

        ema = runner.ModelEma(net, 0.5, opt if fused else None)
        saver = runner.CheckpointSaver(str(tmp_path), "m%d.chpn" % fused)
        run = runner.Runner(net, opt, losses.CrossEntropyLoss(), callbacks=[saver, ema, Spy()])
        run.fit(Loader(3), val_loader=Loader(1), epochs=1)
        raw_after = net.fc.weight.detach().clone()
        assert torch.equal(raw_after, seen["train"])                 # swapped back after the epoch
        assert not torch.equal(seen["val"], seen["train"])           # validation ran on other weights
        # the EMA after 3 steps with decay 0.5 lies strictly between the initial and the final weights
        d_total = (raw_after - w0).norm()
        assert 0 < (seen["val"] - w0).norm() < d_total
        ck = torch.load(os.path.join(str(tmp_path), "m%d.chpn" % fused), map_location="cuda", weights_only=False)
        assert torch.equal(ck["state_dict"]["fc.weight"].reshape(seen["val"].shape[:2]),
                           seen["val"].reshape(seen["val"].shape[:2]))    # the checkpoint holds the EMA weights


def test_optimizer_state_survives_an_arena_rebuild():
    """An optimizer stepped (or loaded) BEFORE the model's first forward builds a loose arena in
    param-group order; the model's first forward rebuilds it in registration order (what the
    data-parallel bucket plan needs) and the momentum accumulated so far must migrate, not vanish."""
    from sota_imagenet_b200 import models, optimizers, runner
    torch.manual_seed(0)
    net = models.resnet26(num_classes=16).cuda().train()
    groups = runner.filter_from_weight_decay(net, skip_list=("bn", "bias"))   # two groups: order != registration
    opt = optimizers.SGD(groups, lr=0.0, momentum=0.9)
    for p in net.parameters():
        p.grad = torch.ones_like(p)
    opt.step()                                   # momentum buffers = 1 everywhere, loose arena
    loose = opt._arenas[0]
    assert getattr(loose, "_loose", False)
    net(torch.randn(2, 3, 64, 64, device="cuda"))   # first forward: arena rebuilt in registration order
    arena = net._arena
    assert arena is not loose and [e[1] for e in arena.entries] == [p for _, p in net.named_parameters()]
    for p in net.parameters():
        p.grad.zero_()
    opt.step()                                   # re-collects: buf = 0.9 * 1 + 0
    for p in net.parameters():
        assert torch.allclose(opt.state[p]["momentum_buffer"], torch.full_like(p, 0.9)), p.shape


def test_graph_step_replays_match_eager_and_follow_the_lr_schedule():
    """runner.GraphStep: the whole training step captured once per input signature and replayed;
    a per-batch LR change (PhasesScheduler) must take effect across replays (the optimizer reads
    its hyper-parameters from a persistent device table, optimizers.SGD.sync_hyperparams)."""
    from sota_imagenet_b200 import losses, models, optimizers, runner
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(32, 3, 64, 64, device="cuda", generator=g)     # (batch 8 is too noisy: see test_gpu_model.py)
    y = torch.randint(0, 16, (32,), device="cuda", generator=g)
    crit = losses.CrossEntropyLoss(smoothing=0.1)
    nets, opts = [], []
    for _ in range(2):
        torch.manual_seed(0)
        net = models.resnet26(num_classes=16).cuda().train()
        nets.append(net)
        opts.append(optimizers.SGD(net.parameters(), lr=0.02, momentum=0.9, weight_decay=1e-4, nesterov=True))
    gs = runner.GraphStep(nets[0], crit, opts[0])
    eager = runner.GraphStep(nets[1], crit, opts[1], enabled=False)
    lrs = [0.002, 0.002, 0.002, 0.001, 0.0, 0.003]     # small: the comparison must not turn chaotic
    for i, lr in enumerate(lrs):
        before = nets[0].fc.weight.detach().clone()
        for o in opts:
            o.param_groups[0]["lr"] = lr
        l0 = gs(x, y)[0].item()
        l1 = eager(x, y)[0].item()
        assert abs(l0 - l1) / abs(l1) < 0.1, (i, l0, l1)        # two trajectories of a tiny noisy problem
        changed = not torch.equal(before, nets[0].fc.weight.detach())
        assert changed == (lr != 0.0), (i, lr)          # lr = 0 inside a replay: weights stay put
    assert gs.replays == len(lrs) - runner.GraphStep.WARMUP and eager.replays == 0
    a, b = nets[0].fc.weight.detach().float(), nets[1].fc.weight.detach().float()
    assert float((a - b).norm() / b.norm()) < 0.1
    # a new input signature (progressive resize) captures its own graph
    x2 = torch.randn(32, 3, 96, 96, device="cuda", generator=g)
    for _ in range(4):
        gs(x2, y)
    assert len(gs.graphs) == 2 and gs.replays == len(lrs) - 2 + 2
    # Runner drives it and the meters get loss / accuracy from the replayed step
    run = runner.Runner(nets[0], opts[0], crit, callbacks=[runner.PhasesScheduler([dict(ep=(0, 1), lr=(0.002, 0.0), mode="cos")])])
    loader = [(x, y)] * 6
    loss, metrics = run._run_loader(loader, train=True)
    assert run.graph_step.replays >= 4 and 0 < loss < 10 and 0 <= metrics["Acc@1"] <= 100


def test_cuda_graph_capture_while_previous_loss_is_alive():
    """A training loop that keeps `loss` around (logging) must still capture: the autograd glue may
    not cache a leaf whose AccumulateGrad node stays bound to the eager steps' stream
    (scripts/probes/graph_capture_probe.py)."""
    from sota_imagenet_b200 import losses, models, optimizers
    torch.manual_seed(0)
    net = models.resnet26(num_classes=16).cuda().train()
    opt = optimizers.SGD(net.parameters(), lr=0.01, momentum=0.9, nesterov=True)
    crit = losses.CrossEntropyLoss(smoothing=0.1)
    x = torch.randn(8, 3, 64, 64, device="cuda")
    y = torch.randint(0, 16, (8,), device="cuda")

    def step():
        opt.zero_grad()
        loss = crit(net(x), y)
        loss.backward()
        opt.step()
        return loss

    held = [step() for _ in range(2)]          # losses (and their autograd nodes) stay referenced
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    out = torch.zeros((), device="cuda")
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        out.copy_(step())
    before = float(out)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert torch.isfinite(out) and float(out) != before and len(held) == 2
