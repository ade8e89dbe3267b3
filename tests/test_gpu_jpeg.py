"""Hybrid JPEG decode on the device (jpeg.decode_batch -> sib_jpeg_idct_rgb) against PIL / libjpeg-turbo:
bit-identical pixels for every stream of a mixed batch (4:4:4 / 4:2:2 / 4:2:0 / grey, odd extents, restart
markers, an ImageNet-sized image, plus streams that take the host route: progressive JPEG, PNG), against
the committed fixture, and through the real-data loader (device decode == host decode, same batches)."""
import io
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from sota_imagenet_b200 import data, jpeg

pytestmark = pytest.mark.gpu
PIL = pytest.importorskip("PIL.Image")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jpeg_fixture.npz")


def synth(h, w, seed=0):
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 100 * np.sin(xx / 7.0 + yy / 13.0), 127 + 100 * np.cos(xx / 5.0 - yy / 9.0),
                    (xx * 3 + yy * 5) % 256], -1) + rng.randn(h, w, 3) * 20
    return PIL.fromarray(np.clip(img, 0, 255).astype(np.uint8))


def encode(im, fmt="JPEG", **kw):
    b = io.BytesIO()
    im.save(b, fmt, **kw)
    return b.getvalue()


def pil_rgb(d):
    return np.asarray(PIL.open(io.BytesIO(d)).convert("RGB"))


def unpack(buf, offsets, dims):
    flat = buf.cpu().numpy()
    return [flat[o:o + h * w * 3].reshape(h, w, 3) for o, (h, w) in zip(offsets.tolist(), dims.tolist())]


def test_mixed_batch_is_bit_identical_to_pil():
    datas = [
        encode(synth(37, 53, 1), quality=75, subsampling=2),
        encode(synth(64, 64, 2), quality=90, subsampling=0),
        encode(synth(48, 80, 3), quality=85, subsampling=1),
        encode(synth(50, 50, 4).convert("L"), quality=60),
        encode(synth(60, 90, 5), quality=80, subsampling=2, restart_marker_blocks=3),
        encode(synth(40, 40, 6), quality=80, progressive=True),          # host route
        encode(synth(30, 20, 7), "PNG"),                                  # host route
        encode(synth(375, 500, 8), quality=92, subsampling=2),
        encode(synth(9, 5, 9), quality=90, subsampling=2),
        encode(synth(333, 500, 10), quality=100, subsampling=1, optimize=True),
        encode(synth(17, 16, 11), quality=30, subsampling=2),
    ]
    statuses = [jpeg.parse(d).status for d in datas]
    assert statuses[5] == 3 and statuses[6] == 1 and sum(s != 0 for s in statuses) == 2
    for workers in (1, 4):
        buf, offsets, dims, labels = jpeg.decode_batch([(d, i) for i, d in enumerate(datas)], workers=workers)
        assert buf.is_cuda and labels.tolist() == list(range(len(datas)))
        for i, (got, d) in enumerate(zip(unpack(buf, offsets, dims), datas)):
            assert np.array_equal(got, pil_rgb(d)), (i, workers)


def test_row_limited_decode_on_device():
    """`crop_fn`: every stream is decoded only down to the MCU row its crop ends in; the rows the crop reads are
    bit-identical to PIL, the boxes come back unchanged."""
    datas = [encode(synth(100, 75, 1), quality=85, subsampling=2), encode(synth(64, 64, 2), quality=90, subsampling=0),
             encode(synth(120, 90, 3), quality=85, subsampling=1, restart_marker_blocks=5),
             encode(synth(80, 60, 4).convert("L"), quality=70), encode(synth(40, 40, 5), quality=80, progressive=True),
             encode(synth(375, 500, 6), quality=92, subsampling=2)]
    ends = [1, 30, 97, 80, 17, 200]

    def crop_fn(dims):
        assert dims.tolist() == [[100, 75], [64, 64], [120, 90], [80, 60], [40, 40], [375, 500]]
        return [[3, max(e - 5, 0), 20, min(e, 5), i & 1] for i, e in enumerate(ends)]       # y0 + h = e

    buf, offsets, dims, labels, boxes = jpeg.decode_batch([(d, 0) for d in datas], workers=2, crop_fn=crop_fn)
    assert boxes.tolist() == crop_fn(dims.numpy())
    for i, (got, d) in enumerate(zip(unpack(buf, offsets, dims), datas)):
        assert np.array_equal(got[:ends[i]], pil_rgb(d)[:ends[i]]), i


def test_committed_fixture_on_device():
    g = np.load(GOLD)
    names = sorted({k.split("/")[0] for k in g.files})
    buf, offsets, dims, _ = jpeg.decode_batch([(g[n + "/bytes"].tobytes(), 0) for n in names])
    for name, got in zip(names, unpack(buf, offsets, dims)):
        assert np.array_equal(got, g[name + "/rgb"]), name


def test_record_loader_device_decode_equals_host_decode(tmp_path):
    root = tmp_path / "imagenet"
    k = 0
    for split in ("train", "val"):
        for cname in ("n01", "n02"):
            (root / split / cname).mkdir(parents=True)
            for j in range(4):
                k += 1
                im = synth(40 + 7 * j, 64 - 5 * j, k)
                (root / split / cname / ("%d.JPEG" % j)).write_bytes(
                    encode(im, quality=70 + 5 * j, subsampling=(0, 1, 2, 2)[j], progressive=(j == 3 and cname == "n02")))
    cfg = SimpleNamespace(image_size=32, batch_size=4, num_classes=2, min_area=0.3, seed=5,
                          root_data_dir=str(root), use_tfrecords=False)
    for train in (False, True):
        split = "train" if train else "val"
        runs = {}
        for mode, prefetch in (("device", 2), ("host", 0), ("device", 0)):
            loader = data.RecordLoader(cfg, data.make_reader(cfg, str(root), split), train=train, decode=mode,
                                       prefetch=prefetch)
            assert loader.decode == mode and loader.prefetch == prefetch
            runs[(mode, prefetch)] = [(x.clone(), t.clone()) for x, t in loader]
        assert len(runs[("device", 2)]) == 2
        for other in (("host", 0), ("device", 0)):          # loader thread + side stream == synchronous == host decode
            for (xd, td), (xh, th) in zip(runs[("device", 2)], runs[other]):
                assert torch.equal(xd, xh) and torch.equal(td, th)
    # a consumer that stops early must not leave the loader thread blocked on its queue
    loader = data.RecordLoader(cfg, data.make_reader(cfg, str(root), "train"), train=True, prefetch=1)
    for _ in loader:
        break
    import threading
    assert not [t for t in threading.enumerate() if t.name == "sib-record-loader" and t.is_alive()]
