"""Ragged-batch kernels (sib_rrc_boxes_ragged / sib_augment_ragged / sib_val_transform_ragged) against
the numpy oracle and against the uniform kernels.

First run on a B200 in round 2 (profiles/r02_pytest_gpu_ragged.log): both tests pass, the opt-in gate
is gone."""
import os

import numpy as np
import pytest
import torch

from oracle import augment_ref
from sota_imagenet_b200 import ops, records

pytestmark = pytest.mark.gpu


def _batch(shapes, seed=0):
    rng = np.random.RandomState(seed)
    images = [rng.randint(0, 256, size=(h, w, 3), dtype=np.uint8) for h, w in shapes]
    buf, offsets, dims = records.pack_batch(images)
    return images, buf.cuda(), offsets.cuda(), dims.cuda()


def test_ragged_boxes_bit_exact():
    shapes = [(48, 64), (100, 400), (375, 500), (500, 333), (17, 19), (256, 256)]
    _, _, _, dims = _batch(shapes)
    boxes = ops.rrc_boxes_ragged(dims, 0.08, 1.0, 42, 1000, True).cpu().tolist()
    want = [augment_ref.rrc_box(h, w, 0.08, 1.0, 42, 1000 + i) for i, (h, w) in enumerate(shapes)]
    assert boxes == want
    nf = ops.rrc_boxes_ragged(dims, 0.9, 1.0, 7, 0, False).cpu()
    assert int(nf[:, 4].abs().sum()) == 0


def test_ragged_train_and_val_match_oracle_and_uniform_kernels():
    shapes = [(40, 56), (56, 40), (33, 47), (64, 64)]
    images, buf, offsets, dims = _batch(shapes, seed=1)
    boxes = ops.rrc_boxes_ragged(dims, 0.2, 1.0, 3, 0, True)
    out = ops.augment_ragged(buf, offsets, dims, boxes, 24, out_mode=1).cpu().numpy().transpose(0, 2, 3, 1)
    for i, im in enumerate(images):
        want = augment_ref.augment_image(im, boxes[i].tolist(), 24)
        assert np.allclose(out[i], want, atol=2e-5), i
    val = ops.val_transform_ragged(buf, offsets, dims, 24, 32, out_mode=1).cpu().numpy().transpose(0, 2, 3, 1)
    for i, im in enumerate(images):
        assert np.allclose(val[i], augment_ref.val_transform_image(im, 24, 32), atol=2e-5), i
    v16 = ops.val_transform_ragged(buf, offsets, dims, 24, 32, out_mode=0)
    assert tuple(v16.shape) == (4, 4, 24, 24) and float(v16[:, 3].abs().max()) == 0.0
    # equal-sized images: bit-identical to the uniform kernels
    same = [(48, 48)] * 5
    images, buf, offsets, dims = _batch(same, seed=2)
    stacked = torch.from_numpy(np.stack(images)).cuda()
    b_u = ops.rrc_boxes(5, 48, 48, 0.08, 1.0, 9, 20, True, "cuda")
    b_r = ops.rrc_boxes_ragged(dims, 0.08, 1.0, 9, 20, True)
    assert torch.equal(b_u, b_r)
    assert torch.equal(ops.augment(stacked, b_u, 32), ops.augment_ragged(buf, offsets, dims, b_r, 32))
    assert torch.equal(ops.val_transform(stacked, 32, 40), ops.val_transform_ragged(buf, offsets, dims, 32, 40))
