"""Writes tests/golden/jpeg_fixture.npz: a handful of small JPEG streams (bytes) together with the RGB
pixels PIL / libjpeg-turbo decodes them to -- the pin of the JPEG decoder stages (oracle/jpeg_ref.py and
csrc/jpeg.cu) that does not depend on the PIL build present at test time.
    python oracle/make_jpeg_golden.py        (Pillow 12.2.0, libjpeg-turbo, IJG API 6.2)"""
import io
import os

import numpy as np
from PIL import Image


def synth(h, w, seed):
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 100 * np.sin(xx / 7.0 + yy / 13.0), 127 + 100 * np.cos(xx / 5.0 - yy / 9.0),
                    (xx * 3 + yy * 5) % 256], -1) + rng.randn(h, w, 3) * 20
    return Image.fromarray(np.clip(img, 0, 255).astype(np.uint8))


CASES = [
    ("444_q90", 24, 40, dict(quality=90, subsampling=0), False),
    ("420_q75_odd", 37, 53, dict(quality=75, subsampling=2), False),
    ("422_q85", 24, 48, dict(quality=85, subsampling=1), False),
    ("420_q95_optimized", 33, 47, dict(quality=95, subsampling=2, optimize=True), False),
    ("grey_q60", 30, 26, dict(quality=60), True),
    ("420_restart", 40, 56, dict(quality=80, subsampling=2, restart_marker_blocks=3), False),
    ("420_tiny", 9, 5, dict(quality=90, subsampling=2), False),
]


def main():
    out = {}
    for i, (name, h, w, kw, grey) in enumerate(CASES):
        im = synth(h, w, i)
        if grey:
            im = im.convert("L")
        b = io.BytesIO()
        im.save(b, "JPEG", **kw)
        data = b.getvalue()
        out[name + "/bytes"] = np.frombuffer(data, dtype=np.uint8)
        out[name + "/rgb"] = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "jpeg_fixture.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
