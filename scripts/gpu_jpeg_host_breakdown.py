"""Where the host time of one real-data batch goes on the GPU box (256 ImageNet-sized JPEG files, warm page cache):
serial file reads, header parsing, batch plan, crop boxes, the Huffman pool, and `jpeg.decode_batch` as a whole with
and `jpeg.decode_batch` as a whole.   gpurun -- python scripts/gpu_jpeg_host_breakdown.py"""
import io
import os
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from PIL import Image  # noqa: E402

from sota_imagenet_b200 import jpeg, ops, records  # noqa: E402


def best(fn, n=4):
    ts = []
    for _ in range(n):
        t = time.time()
        fn()
        ts.append(time.time() - t)
    return min(ts) * 1e3


def main():
    workers = min(32, os.cpu_count() or 8)
    rng = np.random.RandomState(0)
    yy, xx = np.mgrid[0:375, 0:500]
    with tempfile.TemporaryDirectory() as root:
        for c in range(4):
            d = os.path.join(root, "train", "n%d" % c)
            os.makedirs(d)
            for k in range(64):
                img = np.stack([127 + 100 * np.sin(xx / (9.0 + c) + yy / 23.0), 127 + 100 * np.cos(xx / 15.0 - yy / (11.0 + k % 7)),
                                (xx + yy * 2 + 13 * k) % 256], -1) + rng.randn(375, 500, 3) * 12
                Image.fromarray(np.clip(img, 0, 255).astype(np.uint8)).save(os.path.join(d, "%d.JPEG" % k), quality=90)
        rd = records.FileReader(os.path.join(root, "train"), random_shuffle=True)
        eager = list(rd)
        idx = list(range(len(rd)))
        datas = [s[0] for s in eager]
        infos = [jpeg.parse(d) for d in datas]
        plan = jpeg.plan_batch(infos, {})
        boxes = [ops.rrc_box_host(375, 500, 0.08, 1.0, 0, i) for i in range(256)]
        out = np.zeros(plan["coef_total"], np.int16)
        out[:] = 1

        def huff(i):
            jpeg.decode_coefficients(datas[i], out[int(plan["table"][i]["coef_off"][0]):], infos[i],
                                     jpeg.mcu_rows_for(infos[i], boxes[i][1] + boxes[i][3]))

        print("%d host cores, %d threads, %.1f KB per file" % (os.cpu_count(), workers, sum(map(len, datas)) / 256 / 1024))
        print("serial file reads (FileReader iteration)   %6.1f ms" % best(lambda: list(rd)))
        with ThreadPoolExecutor(workers) as ex:
            print("the same reads from the decode pool        %6.1f ms" % best(lambda: list(ex.map(rd.sample, idx))))
            print("header parse, serial                       %6.1f ms" % best(lambda: [jpeg.parse(d) for d in datas]))
            print("batch plan                                 %6.1f ms" % best(lambda: jpeg.plan_batch(infos, {})))
            print("crop boxes on the host                     %6.1f ms" % best(lambda: [ops.rrc_box_host(375, 500, 0.08, 1.0, 0, i) for i in range(256)]))
            print("Huffman stage, %2d threads (ROI)            %6.1f ms" % (workers, best(lambda: list(ex.map(huff, range(256))))))
        print("Huffman stage, 1 thread (ROI)              %6.1f ms" % best(lambda: [huff(i) for i in range(256)], 2))

        def crop_fn(hw):
            return [ops.rrc_box_host(int(h), int(w), 0.08, 1.0, 0, i) for i, (h, w) in enumerate(hw.tolist())]

        def whole(samples):
            jpeg.decode_batch(samples, workers=workers, crop_fn=crop_fn)
            torch.cuda.synchronize()

        whole(eager)
        print("jpeg.decode_batch (incl. H2D + device stage, synchronised) %6.1f ms" % best(lambda: whole(eager)))


if __name__ == "__main__":
    main()
