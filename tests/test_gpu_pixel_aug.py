"""Photometric augmentations of the resident batch (colour twist, grayscale, random erasing, Gaussian
blur: reference dali_dataloader.py:81-111) against the numpy oracle (oracle/augment_ref.py), both
model-input layouts, plus the loader wiring."""
import numpy as np
import pytest
import torch

from oracle import augment_ref
from sota_imagenet_b200 import data, ops

pytestmark = pytest.mark.gpu


class Cfg:
    image_size, batch_size, num_classes = 48, 8, 10
    min_area = 0.3
    blur_prob, color_twist_prob, gray_prob, re_prob, re_count = 0.5, 0.7, 0.3, 0.6, 3
    contrast_range, brightness_range = (0.7, 1.3), (0.7, 1.3)
    seed = 5


def _inputs(layout, n=8, size=48, seed=0):
    g = torch.Generator().manual_seed(seed)
    v = torch.randint(0, 256, (n, size, size, 3), generator=g).float()
    xn = (v - data.DATA_MEAN) / data.DATA_STD
    if layout == 0:
        x = torch.zeros(n, size, size, 4)
        x[..., :3] = xn
        x = x.cuda().to(torch.bfloat16).permute(0, 3, 1, 2)
        ref_in = x.permute(0, 2, 3, 1)[..., :3].float().cpu().numpy()
    else:
        x = xn.permute(0, 3, 1, 2).contiguous().cuda()
        ref_in = xn.numpy()
    return x, ref_in


def _as_nhwc3(x):
    return x.permute(0, 2, 3, 1)[..., :3].float().cpu().numpy()


@pytest.mark.parametrize("layout", [0, 1])
def test_pixel_ops_match_numpy_oracle(layout):
    aug = data.BatchPixelAug(Cfg(), seed=3)
    x, ref_in = _inputs(layout)
    n, size = 8, 48
    _, params = aug.draw(n, size, batch_index=11)
    assert params[:, 14].sum() > 0 and params[:, 16:].sum() > 0 and not np.allclose(params[:, 0:9], np.eye(3).reshape(-1))
    boxes = torch.zeros(n, 5, dtype=torch.int32)
    boxes[::2, 4] = 1                                     # every other sample was mirrored by the crop
    want = augment_ref.pixel_ops(ref_in, params, boxes[:, 4].numpy() != 0, aug.re_count)
    ops.pixel_ops_(x, torch.from_numpy(params).cuda(), boxes.cuda(), aug.re_count)
    got = _as_nhwc3(x)
    if layout == 0:
        want_q = torch.from_numpy(want).to(torch.bfloat16).float().numpy()
        # same fp32 arithmetic, one bf16 rounding: identical except where the fp32 value sits on a rounding tie
        assert np.mean(got == want_q) > 0.999 and np.abs(got - want).max() < 2e-2
        assert float(x[:, 3].float().abs().max()) == 0.0   # the padding channel stays zero
    else:
        assert np.abs(got - want).max() < 1e-5
    # erased pixels carry the fill value (DATA_MEAN -> 0 after normalisation); mirrored boxes moved
    h1, w1, h2, w2 = (int(v) for v in params[0, 16:20])
    if h2 > h1 and w2 > w1:
        assert np.all(got[0, h1:h2, size - w2:size - w1] == 0.0)


@pytest.mark.parametrize("layout", [0, 1])
def test_gaussian_blur_matches_numpy_oracle(layout):
    x, ref_in = _inputs(layout, n=4, size=40, seed=2)
    sigma = np.array([0.0, 0.5, 0.8, 1.1], dtype=np.float32)
    want = augment_ref.gaussian_blur11(ref_in, sigma)
    got = _as_nhwc3(ops.gaussian_blur(x, torch.from_numpy(sigma).cuda()))
    tol = 2e-2 if layout == 0 else 2e-5
    assert np.abs(got - want).max() < tol
    assert np.array_equal(got[0], ref_in[0].astype(np.float32))          # sigma 0: copied through
    assert got[3].std() < ref_in[3].std() * 0.75                         # and it actually blurs


def test_loader_applies_the_configured_augmentations():
    cfg = Cfg()
    src = data.SyntheticSource(pool=32, height=64, width=64, num_classes=10, seed=1)
    plain_cfg = Cfg()
    plain_cfg.blur_prob = plain_cfg.color_twist_prob = plain_cfg.gray_prob = plain_cfg.re_prob = 0
    a = next(iter(data.SyntheticLoader(cfg, src, epoch_size=32)))[0]
    b = next(iter(data.SyntheticLoader(plain_cfg, src, epoch_size=32)))[0]
    assert a.shape == b.shape and not torch.equal(a, b)
    assert torch.isfinite(a.float()).all() and float(a.float().abs().max()) <= 2.51
    # deterministic: same seed and batch index -> same batch
    a2 = next(iter(data.SyntheticLoader(cfg, src, epoch_size=32)))[0]
    assert torch.equal(a, a2)
