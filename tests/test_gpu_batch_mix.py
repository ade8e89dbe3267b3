"""CutMix / Mixup on the resident batch (reference callbacks.py:232-247 on top
of pytorch_tools' Cutmix / Mixup, restated in oracle/augment_ref.py) — kernels bit-exact against
the numpy oracle, callback semantics (coin flip, previous-batch memory, progressive resizing)."""
import numpy as np
import pytest
import torch

from oracle import augment_ref
from sota_imagenet_b200 import ops, runner

pytestmark = pytest.mark.gpu


def _bf16_batch(n, c, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(n, c, h, w, generator=g) * 5 - 2.5).bfloat16()
    return x.cuda().contiguous(memory_format=torch.channels_last), x.float().numpy()


def _bits(t):
    return t.detach().float().cpu().numpy()


@pytest.mark.parametrize("shape", [(8, 4, 32, 32), (5, 4, 24, 40), (3, 8, 7, 9)])
def test_mix_kernels_bit_exact_nhwc_bf16(shape):
    n, c, h, w = shape
    if (w * c) % 8:
        w += 1
    x, xr = _bf16_batch(n, c, h, w, 1)
    p, pr = _bf16_batch(n, c, h, w, 2)
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(3))
    perm_d = perm.to("cuda", torch.int32)
    lam = np.float32(0.3137)
    out = ops.mix_batch(x, p, perm_d, 0, lam, np.float32(1) - lam)
    want = torch.from_numpy(augment_ref.mixup_batch(xr, pr, perm.numpy(), lam)).bfloat16().float().numpy()
    assert np.array_equal(_bits(out), want)
    assert out.permute(0, 2, 3, 1).is_contiguous() and out.dtype == torch.bfloat16
    for box in ((0, 0, h, w), (2, 3, h - 1, w - 2), (0, 0, 0, 0), (h // 2, 1, h // 2 + 1, 2)):
        out = ops.mix_batch(x, p, perm_d, 1, box=box)
        assert np.array_equal(_bits(out), augment_ref.cutmix_batch(xr, pr, perm.numpy(), box))
    # previous batch == the batch itself (first call of the callbacks): out of place, no race
    out = ops.mix_batch(x, x, perm_d, 1, box=(1, 1, h - 1, w - 1))
    assert np.array_equal(_bits(out), augment_ref.cutmix_batch(xr, xr, perm.numpy(), (1, 1, h - 1, w - 1)))


def test_mix_kernels_fp32_nchw_and_targets():
    g = torch.Generator().manual_seed(5)
    x = torch.randn(6, 3, 16, 20, generator=g)
    p = torch.randn(6, 3, 24, 28, generator=g)           # previous stage of a progressive schedule
    perm = torch.randperm(6, generator=g)
    perm_d = perm.to("cuda", torch.int32)
    box = (3, 4, 11, 16)
    out = ops.mix_batch(x.cuda(), p.cuda(), perm_d, 1, box=box)
    assert np.array_equal(out.cpu().numpy(), augment_ref.cutmix_batch(x.numpy(), p.numpy(), perm.numpy(), box))
    q = torch.randn(6, 3, 16, 20, generator=g)
    lam = np.float32(0.77)
    out = ops.mix_batch(x.cuda(), q.cuda(), perm_d, 0, lam, np.float32(1) - lam)
    assert np.array_equal(out.cpu().numpy(), augment_ref.mixup_batch(x.numpy(), q.numpy(), perm.numpy(), lam))
    with pytest.raises(Exception):                       # mixup across sizes fails in the original too
        ops.mix_batch(x.cuda(), p.cuda(), perm_d, 0, lam, np.float32(1) - lam)
    with pytest.raises(Exception):                       # box outside the common extent
        ops.mix_batch(x.cuda(), p.cuda(), perm_d, 1, box=(0, 0, 17, 20))
    t = torch.softmax(torch.randn(6, 50, generator=g), 1)
    pt_ = torch.softmax(torch.randn(6, 50, generator=g), 1)
    out = ops.mix_targets(t.cuda(), pt_.cuda(), perm_d, np.float32(0.25), np.float32(0.75))
    assert np.array_equal(out.cpu().numpy(), augment_ref.mix_targets(t.numpy(), pt_.numpy(), perm.numpy(), 0.25, 0.75))
    assert torch.allclose(out.sum(1).cpu(), torch.ones(6), atol=1e-6)


class _State:
    is_train = True
    input = None


def test_cutmix_mixup_callback_semantics():
    """Replays the callback's host RNG (np.random + torch CPU generator, same call order) and
    checks every batch against the oracle, including the previous-batch memory."""
    np.random.seed(11)
    torch.manual_seed(11)
    cb = runner.CutmixMixup(cutmix_alpha=1.0, mixup_alpha=0.2, prob=0.7, num_classes=20)
    cb.set_state(_State())
    # shadow generators
    np_state, t_state = np.random.get_state(), torch.get_rng_state()
    batches = []
    for i in range(6):
        x, xr = _bf16_batch(4, 4, 16, 16, 100 + i)
        y = torch.randint(0, 20, (4,), generator=torch.Generator().manual_seed(200 + i))
        batches.append((x, xr, y))
    outs = []
    for x, xr, y in batches:
        cb.state.input = (x, y.cuda())
        cb.on_batch_begin()
        outs.append(cb.state.input)
    # replay
    np.random.set_state(np_state)
    torch.set_rng_state(t_state)
    prev = None
    kinds = set()
    for (x, xr, y), (md, mt) in zip(batches, outs):
        t = np.eye(20, dtype=np.float32)[y.numpy()]
        use_cutmix = np.random.rand() > 0.5
        alpha = 1.0 if use_cutmix else 0.2
        if np.random.rand() > 0.7:
            kinds.add("skip")
            assert np.array_equal(_bits(md), xr) and np.array_equal(mt.cpu().numpy(), t)
            continue
        pd, ptg = (xr, t) if prev is None else prev
        prev = (xr, t)
        perm = torch.randperm(4).numpy()
        lam = float(torch.distributions.Beta(alpha, alpha).sample())
        if use_cutmix:
            kinds.add("cutmix")
            lam = min(lam, 1 - lam)
            ch, cw = np.random.randint(16), np.random.randint(16)
            box, lam = augment_ref.cutmix_bbox(16, 16, lam, ch, cw)
            assert np.array_equal(_bits(md), augment_ref.cutmix_batch(xr, pd, perm, box))
            assert np.array_equal(mt.cpu().numpy(), augment_ref.mix_targets(t, ptg, perm, 1 - lam, lam))
        else:
            kinds.add("mixup")
            c = np.float32(lam)
            want = torch.from_numpy(augment_ref.mixup_batch(xr, pd, perm, c)).bfloat16().float().numpy()
            assert np.array_equal(_bits(md), want)
            assert np.array_equal(mt.cpu().numpy(), augment_ref.mix_targets(t, ptg, perm, c, np.float32(1) - c))
        assert abs(float(mt.sum()) - 4.0) < 1e-4
    assert {"cutmix", "mixup"} <= kinds
    # validation batches pass through untouched (targets still become one-hot)
    cb.state.is_train = False
    x, xr, y = batches[0]
    cb.state.input = (x, y.cuda())
    cb.on_batch_begin()
    assert cb.state.input[0] is x and tuple(cb.state.input[1].shape) == (4, 20)
