"""JPEG decode, host side (no GPU): the marker parser and the Huffman stage of `csrc/jpeg.cu` called
through the C ABI, checked three ways --
  * coefficients == an independent pure-Python entropy decoder (oracle/jpeg_ref.huffman_decode);
  * coefficients pushed through the numpy restatement of the DEVICE stages (islow IDCT, fancy
    upsampling, YCbCr -> RGB; oracle/jpeg_ref.decode_rgb) == PIL / libjpeg-turbo, bit for bit;
  * the same against the committed fixture (tests/golden/jpeg_fixture.npz, oracle/make_jpeg_golden.py).
Streams outside the device subset must be reported, not mis-decoded; truncated input must not crash."""
import io
import os

import numpy as np
import pytest

from oracle import jpeg_ref
from sota_imagenet_b200 import jpeg

PIL = pytest.importorskip("PIL.Image")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jpeg_fixture.npz")


def synth(h, w, seed=0):
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 100 * np.sin(xx / 7.0 + yy / 13.0), 127 + 100 * np.cos(xx / 5.0 - yy / 9.0),
                    (xx * 3 + yy * 5) % 256], -1) + rng.randn(h, w, 3) * 20
    return PIL.fromarray(np.clip(img, 0, 255).astype(np.uint8))


def encode(im, **kw):
    b = io.BytesIO()
    im.save(b, "JPEG", **kw)
    return b.getvalue()


def pil_rgb(data):
    return np.asarray(PIL.open(io.BytesIO(data)).convert("RGB"))


CASES = [
    (64, 64, dict(quality=90, subsampling=0), False),                 # 4:4:4
    (37, 53, dict(quality=75, subsampling=2), False),                 # 4:2:0, odd extents (partial MCUs)
    (48, 80, dict(quality=85, subsampling=1), False),                 # 4:2:2
    (33, 47, dict(quality=95, subsampling=2, optimize=True), False),  # optimised Huffman tables
    (50, 50, dict(quality=60), True),                                 # grey
    (100, 75, dict(quality=30, subsampling=2), False),                # coarse quantisation
    (9, 5, dict(quality=90, subsampling=2), False),                   # smallest width the fancy path takes
    (17, 16, dict(quality=100, subsampling=0), False),                # quality 100: large coefficients (slow path)
    (60, 90, dict(quality=80, subsampling=2, restart_marker_blocks=3), False),   # restart intervals
    (60, 90, dict(quality=80, subsampling=1, restart_marker_rows=1), False),
    (375, 500, dict(quality=92, subsampling=2), False),               # an ImageNet-sized image
]


@pytest.mark.parametrize("h,w,kw,grey", CASES)
def test_host_huffman_stage_and_device_arithmetic_match_pil(h, w, kw, grey):
    im = synth(h, w, seed=h * 131 + w)
    data = encode(im.convert("L") if grey else im, **kw)
    info = jpeg.parse(data)
    assert info.status == 0 and (info.height, info.width) == (h, w)
    assert info.ncomp == (1 if grey else 3)
    coef = jpeg.decode_coefficients(data, info=info)
    assert coef.dtype == np.int16 and coef.size == info.coef_count
    if h * w <= 96 * 96:
        _, ref_coef = jpeg_ref.huffman_decode(data)
        assert np.array_equal(coef, ref_coef)
    assert np.array_equal(jpeg_ref.decode_rgb(info.as_dict(), coef), pil_rgb(data))


def test_committed_fixture():
    g = np.load(GOLD)
    names = sorted({k.split("/")[0] for k in g.files})
    assert len(names) == 7
    for name in names:
        data = g[name + "/bytes"].tobytes()
        info = jpeg.parse(data)
        assert info.status == 0, name
        got = jpeg_ref.decode_rgb(info.as_dict(), jpeg.decode_coefficients(data, info=info))
        assert np.array_equal(got, g[name + "/rgb"]), name


def test_streams_outside_the_device_subset_are_reported():
    im = synth(40, 40)
    assert jpeg.parse(encode(im, progressive=True)).status == 3
    cmyk = PIL.fromarray(np.random.RandomState(0).randint(0, 255, (40, 40, 4), dtype=np.uint8), "CMYK")
    assert jpeg.parse(encode(cmyk)).status == 5
    b = io.BytesIO()
    im.save(b, "PNG")
    assert jpeg.parse(b.getvalue()).status == 1
    assert jpeg.parse(encode(synth(8, 3), subsampling=2)).status == 6      # chroma rows of 2 samples: not the fancy path
    assert jpeg.parse(b"").status == 1 and jpeg.parse(b"\xff\xd8\xff").status != 0
    big = bytearray(encode(im, quality=80))
    sof = big.index(b"\xff\xc0")
    big[sof + 5:sof + 9] = b"\xff\xff\xff\xff"            # a header announcing 65535 x 65535 pixels
    assert jpeg.parse(bytes(big)).status == 8
    with pytest.raises(Exception):
        jpeg.decode_coefficients(encode(im, progressive=True))


def test_truncated_and_corrupt_streams_do_not_crash():
    data = encode(synth(64, 64), quality=90)
    rng = np.random.RandomState(1)
    for cut in (len(data) // 2, len(data) - 3, 700):
        info = jpeg.parse(data[:cut])
        if info.status == 0:
            jpeg.decode_coefficients(data[:cut], info=info)          # zeros are fed past the end, like libjpeg
    for _ in range(20):                                              # flip bytes inside the entropy-coded segment
        bad = bytearray(data)
        for p in rng.randint(len(data) // 2, len(data) - 2, size=4):
            bad[p] = rng.randint(0, 256)
        info = jpeg.parse(bytes(bad))
        if info.status == 0:
            jpeg.decode_coefficients(bytes(bad), info=info)


def test_batch_plan_layout():
    datas = [encode(synth(37, 53), subsampling=2), encode(synth(24, 40), subsampling=0), encode(synth(40, 40), progressive=True)]
    infos = [jpeg.parse(d) for d in datas]
    plan = jpeg.plan_batch(infos, {2: (40, 40)})
    assert plan["dims"].tolist() == [[37, 53], [24, 40], [40, 40]]
    assert all(int(o) % 16 == 0 for o in plan["offsets"])
    assert plan["offsets"].tolist() == [0, 5888, 8768]                 # 37*53*3 = 5883 -> 5888, + 2880
    t = plan["table"]
    assert t["coef_off"][0].tolist() == [0, 6 * 8 * 64, 6 * 8 * 64 + 3 * 4 * 64]   # Y 6x8 blocks, Cb 3x4, Cr 3x4
    assert int(t["coef_off"][1][0]) == plan["table"]["coef_off"][0][2] + 3 * 4 * 64
    assert plan["coef_total"] == infos[0].coef_count + infos[1].coef_count
    assert plan["max_pixels"] == 37 * 53 and plan["max_blocks"] == 48 + 12 + 12
    assert jpeg.IMAGE_DTYPE.itemsize == 488


@pytest.mark.parametrize("sub", [0, 1, 2])
def test_row_limited_decode(sub):
    """ROI decoding: stopping after the MCU rows a crop needs yields the same coefficients for those rows, leaves
    the rest of the buffer untouched, and the rows the consumer reads (two rows of margin for the chroma
    interpolation) come out bit-identical to the full decode."""
    data = encode(synth(100, 75, seed=sub), quality=85, subsampling=sub, restart_marker_blocks=(4 if sub == 2 else 0))
    info = jpeg.parse(data)
    full = jpeg.decode_coefficients(data, info=info)
    ref = pil_rgb(data)
    for luma_rows in (1, 13, 14, 15, 16, 47, 98, 99, 100):
        mr = jpeg.mcu_rows_for(info, luma_rows)
        assert 0 <= mr < info.mcus_y and (mr == 0 or mr * 8 * info.vmax >= luma_rows + 2)
        part = np.full(info.coef_count, 12345, dtype=np.int16)
        jpeg.decode_coefficients(data, part, info, mcu_rows=mr)
        off = 0
        for c in range(info.ncomp):
            n = info.blocks_w[c] * info.blocks_h[c] * 64
            rows = (mr if mr else info.mcus_y) * info.vs[c]
            k = info.blocks_w[c] * rows * 64
            assert np.array_equal(part[off:off + k], full[off:off + k])
            assert (part[off + k:off + n] == 12345).all()
            off += n
        part[part == 12345] = 0
        got = jpeg_ref.decode_rgb(info.as_dict(), part)
        assert np.array_equal(got[:luma_rows], ref[:luma_rows]), (sub, luma_rows)
