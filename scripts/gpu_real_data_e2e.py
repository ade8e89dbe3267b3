"""Real-data training rate end to end: class folders of JPEG files -> `records.FileReader` -> `data.RecordLoader`
(hybrid JPEG decode, crop boxes, resample, flip, normalise on the device; loader thread + side stream) ->
ResNet-50 step replayed from a CUDA graph (`runner.GraphStep`).  The files are synthetic ImageNet-sized JPEGs
written to a temporary directory (there is no data set on the box).
    gpurun -- python scripts/gpu_real_data_e2e.py [steps]"""
import io
import os
import sys
import tempfile
import time
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from PIL import Image  # noqa: E402

from sota_imagenet_b200 import data, losses, models, optimizers, runner  # noqa: E402


def write_tree(root, classes=16, per_class=64):
    rng = np.random.RandomState(0)
    yy, xx = np.mgrid[0:375, 0:500]
    for c in range(classes):
        d = os.path.join(root, "train", "n%08d" % c)
        os.makedirs(d)
        for k in range(per_class):
            img = np.stack([127 + 100 * np.sin(xx / (9.0 + c) + yy / 23.0), 127 + 100 * np.cos(xx / 15.0 - yy / (11.0 + k % 7)),
                            (xx + yy * 2 + 13 * k) % 256], -1) + rng.randn(375, 500, 3) * 12
            Image.fromarray(np.clip(img, 0, 255).astype(np.uint8)).save(os.path.join(d, "%d.JPEG" % k), quality=90)
    return classes * per_class


def run(cfg, root, steps, decode, prefetch, workers):
    net = models.resnet50().cuda().train()
    crit = losses.CrossEntropyLoss(smoothing=0.1)
    opt = optimizers.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=3e-5, nesterov=True)
    step = runner.GraphStep(net, crit, opt)
    loader = data.RecordLoader(cfg, data.make_reader(cfg, root, "train"), train=True, decode=decode,
                               prefetch=prefetch, decode_workers=workers)
    n, t0, loss = 0, None, None
    while n < steps + 3:
        for x, t in loader:
            if n == 3:
                torch.cuda.synchronize()
                t0 = time.time()
            loss = step(x, t)[0]
            n += 1
            if n >= steps + 3:
                break
    final = float(loss)
    torch.cuda.synchronize()
    dt = time.time() - t0
    return cfg.batch_size * steps / dt, final


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    workers = min(32, os.cpu_count() or 8)
    with tempfile.TemporaryDirectory() as root:
        n = write_tree(root)
        cfg = SimpleNamespace(image_size=224, batch_size=256, num_classes=1000, min_area=0.08, seed=0,
                              root_data_dir=root, use_tfrecords=False)
        print("%d JPEG files (500x375, q90, 4:2:0), batch 256, %d decode threads, %d host cores" % (n, workers, os.cpu_count()))
        run(cfg, root, 4, "device", 0, workers)          # page cache, allocator, pinned staging buffers
        for decode, prefetch in (("device", 0), ("device", 2), ("host", 0)):
            rate, loss = run(cfg, root, steps, decode, prefetch, workers)
            print("decode=%-6s prefetch=%d : %7.0f images/s end to end (files -> decode -> augment -> ResNet-50 step), loss %.3f"
                  % (decode, prefetch, rate, loss))


if __name__ == "__main__":
    main()
