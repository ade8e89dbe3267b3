"""Hybrid JPEG decode of a 256-image batch of ImageNet-sized streams (500x375, 4:2:0, q90, ~100 KB each):
host Huffman stage (thread pool) and device stage (CUDA events) against PIL's full decode on the same
threads.   gpurun -- python scripts/gpu_jpeg_timing.py"""
import io
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from PIL import Image  # noqa: E402

from sota_imagenet_b200 import jpeg, ops, records  # noqa: E402


def synth(h, w, seed):
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 100 * np.sin(xx / 17.0 + yy / 23.0), 127 + 100 * np.cos(xx / 15.0 - yy / 19.0),
                    (xx + yy * 2) % 256], -1) + rng.randn(h, w, 3) * 12
    b = io.BytesIO()
    Image.fromarray(np.clip(img, 0, 255).astype(np.uint8)).save(b, "JPEG", quality=90, subsampling=2)
    return b.getvalue()


def main():
    B, workers = 256, min(16, os.cpu_count() or 8)
    pool = [synth(375, 500, i) for i in range(16)]
    samples = [(pool[i % 16], i % 1000) for i in range(B)]
    print("batch %d, %.1f KB per stream, %d host threads" % (B, sum(len(s[0]) for s in samples) / B / 1024, workers))
    jpeg.decode_batch(samples, workers=workers)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3):
        out = jpeg.decode_batch(samples, workers=workers)
    torch.cuda.synchronize()
    t_dev = (time.time() - t0) / 3
    t0 = time.time()
    for _ in range(3):
        records.decode_batch(samples, workers=workers, pinned=True)[0].cuda()
    torch.cuda.synchronize()
    t_host = (time.time() - t0) / 3
    # device stage alone
    infos = [jpeg.parse(d) for d, _ in samples]
    plan = jpeg.plan_batch(infos, {})
    coef = torch.zeros(plan["coef_total"], dtype=torch.int16, device="cuda")
    table = torch.from_numpy(plan["table"].view(np.uint8).reshape(B, -1).copy()).cuda()
    planes = torch.empty(plan["plane_total"], dtype=torch.uint8, device="cuda")
    outb = torch.empty(plan["out_total"], dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ops.jpeg_idct_rgb(coef, table, B, plan["max_blocks"], plan["max_pixels"], planes, outb)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.jpeg_idct_rgb(coef, table, B, plan["max_blocks"], plan["max_pixels"], planes, outb)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    traffic = plan["coef_total"] * 2 + 2 * plan["plane_total"] + plan["out_total"]
    print("hybrid decode_batch (host Huffman + H2D + device)  %.1f ms / batch = %.0f images/s" % (t_dev * 1e3, B / t_dev))
    print("host decode (PIL full decode + pack + H2D)          %.1f ms / batch = %.0f images/s" % (t_host * 1e3, B / t_host))
    print("device stage alone (IDCT + upsample + RGB)          %.3f ms / batch, %.0f GB/s of %.0f MB algorithmic traffic"
          % (ms, traffic / ms / 1e6, traffic / 1e6))
    assert out[0].is_cuda


if __name__ == "__main__":
    main()
