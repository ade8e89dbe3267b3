"""CPU: host-side logic — config schema / `_target_` remap, LR phases, data stages, CModel
constructor contract, gradient bucket planning, optimizer segment merging, and a world-size-2
gloo run of the bucketed gradient averaging."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sota_imagenet_b200 import cmodel, config, parallel, runner

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config_loads_reference_schema_and_remaps_targets():
    cfg = config.load_config(os.path.join(ROOT, "configs", "r50_baseline.yaml"), ["loader.batch_size=128", "debug=true"])
    assert cfg.loader.batch_size == 128 and cfg.debug and cfg.loader.image_size == 224
    assert cfg.optim["momentum"] == 0.9 and cfg.optim["weight_decay"] == 3e-5
    assert cfg.criterion["smoothing"] == 0.1 and cfg.bn_momentum == 0.1
    assert [s.lr for s in cfg.run.stages] == [[0.001, 1.0], [1.0, 0]]
    from sota_imagenet_b200 import losses, models, optimizers
    assert config.resolve_target("pytorch_tools.models.resnet50") is models.resnet50
    assert config.resolve_target("torch.optim._multi_tensor.SGD") is optimizers.SGD
    assert config.resolve_target("sota_imagenet.angular_losses.AdaCos") is losses.AdaCos
    assert config.resolve_target("src.model.CModel") is cmodel.CModel
    crit = config.call(cfg.criterion)
    assert isinstance(crit, losses.CrossEntropyLoss) and crit.smoothing == 0.1
    arc = config.load_config(os.path.join(ROOT, "configs", "r50_arcface.yaml"))
    c = config.call(arc.criterion)
    assert isinstance(c, losses.AdditiveAngularMarginLoss) and c.smoothing == 0.1 and c.s == 10.0
    with pytest.raises(KeyError):
        config.load_config(None, ["no_such_key=1"])


def test_reference_yaml_runs_unchanged():
    ref = "/root/reference/configs/hydra_exp/1.r50_baseline.yaml"
    if not os.path.exists(ref):
        pytest.skip("reference tree not mounted")
    cfg = config.load_config(ref)
    assert cfg.model["_target_"] == "pytorch_tools.models.resnet50"
    assert cfg.run.stages[1].lr_mode == "cos" and cfg.optim["weight_decay"] == 3e-5


def test_phases_scheduler():
    stages = [dict(ep=(0, 8), lr=(0.001, 1.0), mode="linear"), dict(ep=(8, 90), lr=(1.0, 0), mode="cos")]
    lr = runner.PhasesScheduler.lr_at
    assert abs(lr(stages, 0.0) - 0.001) < 1e-9
    assert abs(lr(stages, 4.0) - 0.5005) < 1e-9
    assert abs(lr(stages, 8.0) - 1.0) < 1e-9
    assert abs(lr(stages, 49.0) - 0.5) < 1e-9
    assert abs(lr(stages, 90.0)) < 1e-9


def test_data_manager_stage_semantics():
    from sota_imagenet_b200 import data
    cfg = config.load_config(os.path.join(ROOT, "configs", "r50_progressive.yaml"))
    made = []

    class FakeLoader:
        def __init__(self, c, *a, **k):
            made.append((c.image_size, k.get("train")))
    orig, data.SyntheticLoader = data.SyntheticLoader, FakeLoader
    try:
        dm = data.DataManager(cfg)
        assert len(dm) == 3 and dm.tot_epochs == 9
        dm.set_stage(0)
        assert (dm.start_epoch, dm.end_epoch) == (0, 3) and made[-2:] == [(128, True), (128, False)]
        dm.set_stage(1)
        assert made[-2:] == [(192, True), (192, False)]      # val size follows train size
        cfg.run.stages[2].extra_args = None
        n = len(made)
        dm.set_stage(2)
        assert len(made) == n and dm.end_epoch == 9           # LR-only stage keeps the loaders
        cfg.run.stages[1].start = 4
        with pytest.raises(AssertionError):
            data.DataManager(cfg)
    finally:
        data.SyntheticLoader = orig


def test_cmodel_constructor_contract():
    """Mirror of the reference's inline self-test (model.py:1270-1376) on plain torch modules."""
    assert cmodel._update_dict({"foo": {"a": 10, "b": 20}}, {"foo": {"a": 12, "c": 30}}) == {"foo": {"a": 12, "b": 20, "c": 30}}
    layers = [
        dict(module="nn.Conv2d", args=[3, 8, 3], kwargs=dict(padding=1), tag="c1"),
        dict(module="nn.Conv2d", args=[8, 8, 3], kwargs=dict(padding="1"), repeat=2),
        dict(module="Concat", inputs=["_prev_", "c1"]),
        dict(module="nn.Conv2d", args=[16, 4, 1]),
    ]
    m = cmodel.CModel(layers)
    assert m(torch.zeros(1, 3, 8, 8)).shape == (1, 4, 8, 8)
    m2 = cmodel.CModel([dict(module="nn.Conv2d", args=[3, 8, 3])], extra_kwargs={"nn.Conv2d": {"bias": False, "padding": 1}})
    assert m2[0].bias is None and m2[0].padding == (1, 1)
    r50 = cmodel.CModel([
        dict(module="StemConv", args=[64, 7, 3]),
        dict(module="BatchNorm2d", args=[64], kwargs=dict(activation="'relu'")),
        dict(module="MaxPool3x3s2"),
        dict(module="Bottleneck", args=[64, 64], kwargs=dict(downsample=True)),
        dict(module="Bottleneck", args=[256, 64], repeat=2),
        dict(module="GlobalAvgPool"),
        dict(module="Linear", args=[256, 1000]),
    ])
    assert sum(p.numel() for p in r50.parameters()) > 2e5


def test_bucket_plan():
    starts = [100, 300, 600, 1000]
    plan = parallel.plan_buckets(starts, 1500, 450)
    assert plan == {3: (1000, 1500), 1: (300, 1000), -1: (0, 300)}
    covered = sorted(plan.values())
    assert covered[0][0] == 0 and covered[-1][1] == 1500
    assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))      # disjoint, complete


def test_sgd_segment_merging():
    from sota_imagenet_b200 import optimizers

    class FakeArena:
        total = 512
    p = [torch.nn.Parameter(torch.zeros(4)) for _ in range(4)]
    a = FakeArena()
    a.entries = [("a", p[0], 0, 4, "plain"), ("b", p[1], 128, 4, "plain"), ("c", p[2], 256, 4, "plain"), ("d", p[3], 384, 4, "plain")]
    opt = optimizers.SGD([{"params": [p[0], p[1]]}, {"params": [p[2]], "weight_decay": 0.0}], lr=0.1, momentum=0.9, weight_decay=1e-4)
    recs = opt._segments(a)
    assert recs == [(256, 0.1, 1e-4, 0.9, 0.0, False), (384, 0.1, 0.0, 0.9, 0.0, False), (512, 0.0, 0.0, 0.0, 0.0, False)]


def _gloo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat = torch.arange(1500, dtype=torch.float32) * (rank + 1)
    plan = parallel.plan_buckets([100, 300, 600, 1000], 1500, 450)
    for key in (3, 1, -1):                       # reverse order, as backward would issue them
        parallel.allreduce_mean_(flat, *plan[key])
    torch.save(flat, os.path.join(out, "r%d.pt" % rank))
    dist.destroy_process_group()


def test_bucketed_mean_allreduce_gloo_world2():
    with tempfile.TemporaryDirectory() as out:
        mp.spawn(_gloo_worker, args=(2, 29431, out), nprocs=2, join=True)
        r0, r1 = torch.load(os.path.join(out, "r0.pt")), torch.load(os.path.join(out, "r1.pt"))
    want = torch.arange(1500, dtype=torch.float32) * 1.5
    assert torch.equal(r0, want) and torch.equal(r1, want)


def _meter_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sota_imagenet_b200 import runner
    run = runner.Runner(torch.nn.Linear(2, 2), None, None)
    run.state.loss_meter.update(1.0 + rank, 10 * (rank + 1))           # rank 0: 10 samples of 1, rank 1: 20 of 2
    run.state.metric_meters["Acc@1"].update(50.0 * rank, 10 * (rank + 1))
    run._reduce_meters()
    torch.save((run.state.loss_meter.avg, run.state.loss_meter.n, run.state.metric_meters["Acc@1"].avg),
               os.path.join(out, "m%d.pt" % rank))
    dist.destroy_process_group()


def test_runner_meters_are_reduced_over_ranks_gloo_world2():
    """Each rank sees 1 / world of the data; rank 0 must log whole-dataset numbers."""
    with tempfile.TemporaryDirectory() as out:
        mp.spawn(_meter_worker, args=(2, 29433, out), nprocs=2, join=True)
        m0, m1 = torch.load(os.path.join(out, "m0.pt")), torch.load(os.path.join(out, "m1.pt"))
    assert m0 == m1
    assert abs(m0[0] - (10 * 1.0 + 20 * 2.0) / 30) < 1e-12 and m0[1] == 30 and abs(m0[2] - 20 * 50.0 / 30) < 1e-9


def test_initialize_gain_and_ema_only_for_accepting_optimizers():
    """pt.utils.misc.initialize(model, gamma) stand-in (reference train.py:70-71) and the ema_decay
    plumbing of train.py: only optimizers that accept the keyword receive it."""
    import inspect
    from sota_imagenet_b200 import models, optimizers, runner
    torch.manual_seed(0)
    net = models.resnet26(num_classes=16)
    runner.initialize(net, 1.72)
    w = net.layer2[0].conv2.weight
    fan_in = w[0].numel()
    assert abs(float(w.std()) * fan_in ** 0.5 / 1.72 - 1) < 0.05
    assert float(net.layer1[0].bn1.weight.min()) == 1.0 and float(net.layer1[0].bn1.bias.abs().max()) == 0.0
    assert "ema_decay" in inspect.signature(optimizers.SGD.__init__).parameters
    assert "ema_decay" in inspect.signature(optimizers.MyNovograd.__init__).parameters
    assert "ema_decay" not in inspect.signature(torch.optim.AdamW.__init__).parameters


def test_batch_pixel_aug_parameter_table():
    """Host side of the photometric augmentations (reference dali_dataloader.py:81-111): the colour
    twist as one affine map, its transport into normalised space, probabilities, box geometry."""
    import numpy as np
    from sota_imagenet_b200 import data
    a, t = data.color_twist_matrix()
    assert np.allclose(a, np.eye(3), atol=1e-12) and np.allclose(t, 0)
    a, t = data.color_twist_matrix(saturation=0.0)                 # saturation 0 = luma in every channel
    assert np.allclose(a, np.tile([0.299, 0.587, 0.114], (3, 1)), atol=2e-3)
    a, t = data.color_twist_matrix(contrast=0.5, brightness=2.0)
    assert np.allclose(a, np.eye(3)) and np.allclose(t, 128.0)     # 2 * (128 + 0.5 * (v - 128)) = v + 128
    a, _ = data.color_twist_matrix(hue_deg=120.0)
    assert np.allclose(a @ a @ a, np.eye(3), atol=1e-9)            # three 120-degree hue turns = identity

    class Cfg:
        blur_prob, color_twist_prob, gray_prob, re_prob, re_count = 0.5, 1.0, 0.25, 1.0, 2
        contrast_range, brightness_range = (0.7, 1.3), (0.7, 1.3)

    aug = data.BatchPixelAug(Cfg(), seed=1)
    assert aug.active
    sigma, p = aug.draw(4000, 64, batch_index=3)
    assert p.shape == (4000, 16 + 8) and 0.45 < (sigma > 0).mean() < 0.55
    assert sigma[sigma > 0].min() >= 0.5 and sigma.max() <= 1.1 and 0.2 < p[:, 14].mean() < 0.3
    assert np.all(p[:, 12] == -2.5) and np.all(p[:, 13] == 2.5)
    boxes = p[:, 16:].reshape(-1, 2, 4)
    ext_h, ext_w = boxes[..., 2] - boxes[..., 0], boxes[..., 3] - boxes[..., 1]
    assert boxes.min() >= 0 and boxes.max() <= 64 and ext_h.max() <= 17 and ext_w.max() <= 17 and ext_h.mean() > 6
    # the affine map in normalised space reproduces the 255-space map
    rng = np.random.RandomState(0)
    v = rng.uniform(0, 255, size=3)
    a, t = data.color_twist_matrix(1.2, 0.9, 15.0, 1.1)
    xn = (v - data.DATA_MEAN) / data.DATA_STD
    o = (a @ np.full(3, data.DATA_MEAN) + t - data.DATA_MEAN) / data.DATA_STD
    assert np.allclose(a @ xn + o, (a @ v + t - data.DATA_MEAN) / data.DATA_STD)
    sigma2, p2 = aug.draw(16, 64, batch_index=3)
    s3, p3 = data.BatchPixelAug(Cfg(), seed=1).draw(16, 64, batch_index=3)
    assert np.array_equal(p3, p2) and np.array_equal(s3, sigma2)   # deterministic


def test_novograd_norm_groups_and_table_layout():
    """Host side of MyNovograd: norm groups per tensor / per output unit (reference
    optimizers.py:18-22) and the 56-byte records csrc/optim.cu NovoTensor expects."""
    import numpy as np
    from sota_imagenet_b200 import ops, optimizers
    ng = optimizers.MyNovograd.norm_groups
    assert ng((64, 3, 7, 7), False) == (64 * 147, 1)
    assert ng((64, 3, 7, 7), True) == (147, 64)
    assert ng((1000, 2048), True) == (2048, 1000)
    assert ng((256,), True) == (256, 1) and ng((), False) == (1, 1)
    recs = [(0, 128, 147, 8, 0, 0.01, 1 - 1e-4, 0.9, 0.1, 0.99, 0.01),
            (128, 1280, 1152, 1, 8, 0.0, 1.0, 1.0, 0.0, 1.0, 0.0)]
    raw = ops.novograd_table(recs)
    assert raw.dtype == torch.uint8 and raw.numel() == 2 * 56
    b = raw.numpy().tobytes()
    assert np.frombuffer(b[:16], "<i8").tolist() == [0, 128]
    assert np.frombuffer(b[16:28], "<i4").tolist() == [147, 8, 0]
    assert np.allclose(np.frombuffer(b[28:52], "<f4"), [0.01, 1 - 1e-4, 0.9, 0.1, 0.99, 0.01])
    assert np.frombuffer(b[56:72], "<i8").tolist() == [128, 1280]
    with pytest.raises(ValueError):
        optimizers.MyNovograd([torch.nn.Parameter(torch.zeros(4))], betas=(1.0, 0.99))


def test_cutmix_mixup_host_logic_with_oracle_kernels(monkeypatch):
    """Host side of CutmixMixup (reference callbacks.py:232-247): coin flips, previous-batch
    memory, box arithmetic and target weights, with the two kernels replaced by the numpy oracle
    (the GPU test runs the same replay against the CUDA kernels)."""
    import numpy as np
    from oracle import augment_ref
    from sota_imagenet_b200 import ops

    def mix_batch(x, prev, perm, mode, lam=1.0, one_minus_lam=0.0, box=(0, 0, 0, 0)):
        xr, pr, pm = x.numpy(), prev.numpy(), perm.numpy()
        out = augment_ref.mixup_batch(xr, pr, pm, lam) if mode == 0 else augment_ref.cutmix_batch(xr, pr, pm, box)
        return torch.from_numpy(out)

    monkeypatch.setattr(ops, "mix_batch", mix_batch)
    monkeypatch.setattr(ops, "mix_targets", lambda t, p, perm, a, b: torch.from_numpy(
        augment_ref.mix_targets(t.numpy(), p.numpy(), perm.numpy(), a, b)))
    monkeypatch.setattr(ops, "one_hot", lambda l, n: torch.eye(n)[l])

    class State:
        is_train, input = True, None

    np.random.seed(3)
    torch.manual_seed(3)
    cb = runner.CutmixMixup(cutmix_alpha=1.0, mixup_alpha=0.2, prob=1.0, num_classes=10)
    cb.set_state(State())
    first = None
    for i in range(8):
        x = torch.randn(4, 3, 12, 12)
        y = torch.randint(0, 10, (4,))
        if first is None:
            first = x.clone()
        cb.state.input = (x, y)
        cb.on_batch_begin()
        md, mt = cb.state.input
        assert md.shape == x.shape and tuple(mt.shape) == (4, 10)
        assert torch.allclose(mt.sum(1), torch.ones(4), atol=1e-6)        # soft targets stay normalised
        assert torch.equal(cb.prev_input[0], x)                           # memory = the UNMIXED batch
        # every output pixel comes from this batch, the previous one, or a blend of the two
        assert float(md.abs().max()) <= max(float(x.abs().max()), float(cb_prev_max if i else x.abs().max())) + 1e-6
        cb_prev_max = float(x.abs().max())
    box, lam = augment_ref.cutmix_bbox(224, 224, 0.25, 0, 223)            # clipped at two borders
    assert box == (0, 167, 56, 224) and abs(lam - 56 * 57 / 224 ** 2) < 1e-12
    assert runner.Cutmix.rand_bbox.__doc__
    cb.state.is_train = False
    x = torch.randn(4, 3, 12, 12)
    cb.state.input = (x, torch.randint(0, 10, (4,)))
    cb.on_batch_begin()
    assert cb.state.input[0] is x


def test_bench_profile_merge_takes_per_call_minimum():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class Ev:
        def __init__(self, t):
            self.t = t

        def elapsed_time(self, other):
            return other.t - self.t

    def run(ts):
        return [(n, (i,), Ev(0.0), Ev(t)) for i, (n, t) in enumerate(zip(("a", "b", "a"), ts))]

    merged = bench.merge_profiles([run((1.0, 5.0, 2.0)), run((1.5, 2.0, 2.5)), run((0.9, 2.2, 9.0))])
    assert merged == [("a", (0,), 0.9), ("b", (1,), 2.0), ("a", (2,), 2.0)]
    odd = [run((1.0, 5.0, 2.0)), run((1.0, 5.0, 2.0))[:2]]
    assert bench.merge_profiles(odd) == [("a", (0,), 1.0), ("b", (1,), 5.0), ("a", (2,), 2.0)]
    assert bench.merge_profiles([run((3.0, 1.0, 2.0))])[0][2] == 3.0


def test_reference_targets_of_the_widened_rows_resolve():
    """`_target_` strings of the reference's configs for MyNovograd / CutmixMixup / ArcCosSoftmax
    (e.g. configs/hydra_exp/67.vgg-cmodel_s2d4_frn1.yaml:126) resolve to this package's classes."""
    from sota_imagenet_b200 import losses, optimizers
    opt = config.call({"_target_": "sota_imagenet.optimizers.MyNovograd", "weight_decay": 1e-3,
                       "unitwise_norm": True}, [torch.nn.Parameter(torch.zeros(4))])
    assert isinstance(opt, optimizers.MyNovograd) and opt.unitwise_norm and opt.defaults["betas"] == (0.9, 0.99)
    cb = config.call({"_target_": "sota_imagenet.callbacks.CutmixMixup", "cutmix_alpha": 1.0, "mixup_alpha": 0.2})
    assert isinstance(cb, runner.CutmixMixup) and cb.prob == 0.5
    assert isinstance(config.call({"_target_": "src.callbacks.CutmixMixup", "cutmix_alpha": 1.0, "mixup_alpha": 1.0}),
                      runner.CutmixMixup)
    crit = config.call({"_target_": "src.angular_losses.ArcCosSoftmax", "smoothing": 0.1})
    assert isinstance(crit, losses.ArcCosSoftmax) and crit.smoothing == 0.1
    with pytest.raises(AssertionError):
        losses.AdaCos(final_criterion=None, arc_logits=True)      # reference angular_losses.py:277


def test_cmodel_graph_executor_matches_manual_wiring_and_releases_outputs():
    """Tagged multi-input graph (the reference's U-Net / FPN style self-test, model.py:1304-1356):
    same numbers as wiring the layers by hand; retained outputs are dropped after their last reader."""
    torch.manual_seed(0)
    layers = [
        dict(module="nn.Conv2d", args=[3, 4, 3], kwargs=dict(padding=1), tag="a"),
        dict(module="nn.ReLU"),
        dict(module="nn.Conv2d", args=[4, 4, 3], kwargs=dict(padding=1), tag="b"),
        dict(module="Concat", inputs=["a", "b"], tag="ab"),
        dict(module="nn.Conv2d", args=[8, 2, 1]),
        dict(module="Concat", inputs=["_prev_", "ab", "a"]),
    ]
    m = cmodel.CModel(layers)
    assert m.saved_layers_idx == [0, 2, 3, 0] and m._last_reader == {0: 5, 2: 3, 3: 5}
    assert [l.input_indexes for l in m] == [[-1], [-1], [-1], [0, 2], [-1], [-1, 3, 0]]
    x = torch.randn(2, 3, 5, 5)
    a = m[0](x)
    b = m[2](torch.relu(a))
    ab = torch.cat([a, b], 1)
    want = torch.cat([m[4](ab), ab, a], 1)
    assert torch.allclose(m(x), want, atol=1e-6) and want.shape == (2, 14, 5, 5)   # Concat emits channels_last
    with pytest.raises(KeyError):
        cmodel.CModel([dict(module="nn.ReLU", inputs=["nope"])])
    # ModuleStructure instances are accepted as well as dicts; a chain keeps nn.Sequential.forward
    chain = cmodel.CModel([cmodel.ModuleStructure(module="nn.ReLU"), dict(module="nn.Identity", repeat=3)])
    assert chain.saved_layers_idx == [] and isinstance(chain[1], torch.nn.Sequential) and len(chain[1]) == 3
    assert "forward" not in chain.__dict__
