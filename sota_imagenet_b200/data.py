"""Synthetic-input GPU data path standing in for sota_imagenet/dali_dataloader.py.

  SyntheticSource   deterministic pool of uint8 HWC "decoded images" + labels (replaces the
                    file / TFRecord readers + nvJPEG decode, dali_dataloader.py:47-72; the
                    north_star scopes real decoding out);
  GpuAugment        random-resized-crop (area [min_area,1], aspect [0.75,1.25], 100 attempts) ->
                    triangular resize to SxS -> mirror p=.5 -> (v-127.5)/51 (:65-74,113-122) as one
                    kernel; crop boxes from a counter-based Philox stream (bit-exact vs the oracle);
  HostPrefetcher    double-buffered pinned-host -> device staging on a copy stream (DALI's prefetch
                    queue), used by bench.py's end-to-end measurement;
  SyntheticLoader   DaliLoader surface (:163-186): iterable of (data, one-hot label), batch_size,
                    __len__, drop-last;
  DataManager       DaliDataManager surface (:189-239): stages with extra_args overrides
                    (progressive resize / batch change), loaders rebuilt only when data args change.
"""
import copy
import math
import os

import torch

from . import ops

DATA_MEAN, DATA_STD = 127.5, 51.0   # dali_dataloader.py:27-29 (0.5*255, 0.2*255)


class GpuAugment:
    def __init__(self, image_size=224, min_area=0.08, max_area=1.0, seed=0, flip=True,
                 output="nhwc4_bf16"):
        assert output in ("nhwc4_bf16", "nchw_f32")
        self.image_size, self.min_area, self.max_area = image_size, min_area, max_area
        self.seed, self.flip, self.output = seed, flip, output

    def boxes(self, batch, src_h, src_w, first_sample, device):
        return ops.rrc_boxes(batch, src_h, src_w, self.min_area, self.max_area, self.seed,
                             first_sample, self.flip, device)

    def __call__(self, src_u8, first_sample=0, return_boxes=False):
        """src_u8: [B, H, W, 3] uint8 CUDA.  Returns the model input (bf16 channels_last
        [B,4,S,S] with a zero 4th channel, or fp32 NCHW [B,3,S,S] like DALI emits)."""
        b, h, w, _ = src_u8.shape
        bx = self.boxes(b, h, w, first_sample, src_u8.device)
        out = ops.augment(src_u8, bx, self.image_size, DATA_MEAN, DATA_STD,
                          0 if self.output == "nhwc4_bf16" else 1)
        return (out, bx) if return_boxes else out


_RGB2YIQ = ((0.299, 0.587, 0.114), (0.596, -0.274, -0.321), (0.211, -0.523, 0.311))


def color_twist_matrix(contrast=1.0, brightness=1.0, hue_deg=0.0, saturation=1.0):
    """fn.color_twist as ONE affine map of the uint8 pixel: hue / saturation = a rotation + scaling
    of the chroma plane in YIQ, then out = brightness * (128 + contrast * (. - 128)).
    -> (A [3,3], t [3]) float64 with out = A @ rgb + t (before the [0, 255] clamp)."""
    import numpy as np
    m = np.array(_RGB2YIQ, dtype=np.float64)
    h = math.radians(hue_deg)
    rot = np.array([[1.0, 0.0, 0.0], [0.0, saturation * math.cos(h), -saturation * math.sin(h)],
                    [0.0, saturation * math.sin(h), saturation * math.cos(h)]])
    a = brightness * contrast * (np.linalg.inv(m) @ rot @ m)
    t = np.full(3, brightness * 128.0 * (1.0 - contrast))
    return a, t


class BatchPixelAug:
    """The photometric part of the reference's train_pipeline (dali_dataloader.py:81-111) on the
    resident, already normalised batch: gaussian blur (prob blur_prob, sigma ~ U(0.5, 1.1), window
    11), colour twist (prob color_twist_prob; contrast / brightness from their ranges, hue ~ U(-20,
    20) degrees, saturation ~ U(0.7, 1.3)), grayscale (prob gray_prob), random erasing (prob re_prob,
    re_count boxes, anchor ~ U(0, 1), extent ~ U(0.05, 0.25) of the image, fill = DATA_MEAN).
    Per-sample parameters come from a host numpy stream seeded by (seed, batch index) -- the numpy
    oracle (oracle/augment_ref.py) consumes the same table -- and travel as one small H2D copy; the
    pixels never leave the device.  DALI keeps uint8 between the operators; here the chain is
    evaluated in fp32 on the normalised values and rounded once (restated, unpinned: DALI is absent)."""

    def __init__(self, cfg, seed=0):
        self.blur_prob = float(getattr(cfg, "blur_prob", 0) or 0)
        self.twist_prob = float(getattr(cfg, "color_twist_prob", 0) or 0)
        self.gray_prob = float(getattr(cfg, "gray_prob", 0) or 0)
        self.re_prob = float(getattr(cfg, "re_prob", 0) or 0)
        self.re_count = int(getattr(cfg, "re_count", 3) or 0) if self.re_prob > 0 else 0
        self.contrast_range = tuple(getattr(cfg, "contrast_range", (0.7, 1.3)))
        self.brightness_range = tuple(getattr(cfg, "brightness_range", (0.7, 1.3)))
        self.seed = seed

    @property
    def active(self):
        return self.blur_prob > 0 or self.twist_prob > 0 or self.gray_prob > 0 or self.re_prob > 0

    def draw(self, batch, size, batch_index):
        """-> (sigma [B] float32, params [B, 16 + 4 * re_count] float32) numpy arrays."""
        import numpy as np
        rng = np.random.RandomState((self.seed * 1000003 + batch_index * 7919 + 17) & 0x7FFFFFFF)
        sigma = np.zeros(batch, dtype=np.float32)
        params = np.zeros((batch, 16 + 4 * self.re_count), dtype=np.float32)
        lo, hi = (0.0 - DATA_MEAN) / DATA_STD, (255.0 - DATA_MEAN) / DATA_STD
        for i in range(batch):
            if self.blur_prob > 0 and rng.rand() < self.blur_prob:
                sigma[i] = rng.uniform(0.5, 1.1)
            a, t = np.eye(3), np.zeros(3)
            if self.twist_prob > 0 and rng.rand() < self.twist_prob:
                a, t = color_twist_matrix(rng.uniform(*self.contrast_range), rng.uniform(*self.brightness_range),
                                          rng.uniform(-20.0, 20.0), rng.uniform(0.7, 1.3))
            # the same map on normalised values x = (v - mean) / std
            params[i, 0:9] = a.reshape(-1)
            params[i, 9:12] = (a @ np.full(3, DATA_MEAN) + t - DATA_MEAN) / DATA_STD
            params[i, 12], params[i, 13] = lo, hi
            params[i, 14] = 1.0 if (self.gray_prob > 0 and rng.rand() < self.gray_prob) else 0.0
            params[i, 15] = 0.0                               # erase fill = DATA_MEAN -> 0 after normalisation
            if self.re_count and rng.rand() < self.re_prob:
                anchor = rng.uniform(0.0, 1.0, size=2 * self.re_count)
                shape = rng.uniform(0.05, 0.25, size=2 * self.re_count)
                for b in range(self.re_count):
                    h1, w1 = int(anchor[2 * b] * size), int(anchor[2 * b + 1] * size)
                    h2 = min(size, int((anchor[2 * b] + shape[2 * b]) * size))
                    w2 = min(size, int((anchor[2 * b + 1] + shape[2 * b + 1]) * size))
                    params[i, 16 + 4 * b:20 + 4 * b] = (h1, w1, h2, w2)
        return sigma, params

    def __call__(self, x, crop_boxes, batch_index):
        if not self.active:
            return x
        sigma, params = self.draw(x.shape[0], x.shape[2], batch_index)
        if self.blur_prob > 0 and float(sigma.max()) > 0:
            x = ops.gaussian_blur(x, torch.from_numpy(sigma).to(x.device, non_blocking=True))
        if self.twist_prob > 0 or self.gray_prob > 0 or self.re_count:
            ops.pixel_ops_(x, torch.from_numpy(params).to(x.device, non_blocking=True), crop_boxes, self.re_count)
        return x


class GpuValTransform:
    """Validation transform of the reference (dali_dataloader.py:146-160): resize the shorter side
    to `crop_size` (= image_size when full_crop, else ceil((1.14*image_size + 8) // 16 * 16)),
    centre crop image_size, normalise."""

    def __init__(self, image_size=224, full_crop=False, output="nhwc4_bf16"):
        assert output in ("nhwc4_bf16", "nchw_f32")
        self.image_size, self.output = image_size, output
        self.crop_size = image_size if full_crop else math.ceil((image_size * 1.14 + 8) // 16 * 16)

    def __call__(self, src_u8, first_sample=0):
        return ops.val_transform(src_u8, self.image_size, self.crop_size, DATA_MEAN, DATA_STD,
                                 0 if self.output == "nhwc4_bf16" else 1)


class SyntheticSource:
    """Pool of `pool` random images (uint8 HWC) and labels, generated once with a fixed seed."""

    def __init__(self, pool=1024, height=256, width=256, num_classes=1000, seed=0, device="cuda",
                 pinned_host=False):
        g = torch.Generator().manual_seed(seed)
        imgs = torch.randint(0, 256, (pool, height, width, 3), dtype=torch.uint8, generator=g)
        labels = torch.randint(0, num_classes, (pool,), generator=g)
        self.pool, self.num_classes = pool, num_classes
        if pinned_host:
            self.images, self.labels = imgs.pin_memory(), labels.pin_memory()
        else:
            self.images, self.labels = imgs.to(device), labels.to(device)

    def batch(self, index, batch_size):
        lo = (index * batch_size) % self.pool
        if lo + batch_size <= self.pool:
            return self.images[lo:lo + batch_size], self.labels[lo:lo + batch_size]
        idx = (torch.arange(batch_size) + lo) % self.pool
        idx = idx.to(self.images.device)
        return self.images[idx], self.labels[idx]


class HostPrefetcher:
    """Double-buffered host -> device staging of (uint8 images, labels) batches, the role DALI's
    prefetch queue plays in the reference (dali_dataloader.py:163-186 `prefetch_queue_depth`): the
    copy of batch i+1 runs on its own stream while step i computes.  Every batch is still copied
    from pinned host memory exactly once; slots are reused only after the step that read them has
    been waited for by the caller (the per-step loss read-back in a training loop)."""

    def __init__(self, source, batch_size, device="cuda", first_index=0):
        self.source, self.batch_size, self.device = source, batch_size, torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        imgs, labels = source.batch(first_index, batch_size)
        self.slots = [(torch.empty_like(imgs, device=self.device), torch.empty_like(labels, device=self.device),
                       torch.cuda.Event()) for _ in range(2)]
        self.next_index = first_index
        self._issue()

    def _issue(self):
        imgs, labels = self.source.batch(self.next_index, self.batch_size)
        d_imgs, d_labels, ev = self.slots[self.next_index % 2]
        with torch.cuda.stream(self.stream):
            d_imgs.copy_(imgs, non_blocking=True)
            d_labels.copy_(labels, non_blocking=True)
            ev.record(self.stream)

    def next(self):
        """(images, labels, index) of the next batch on the device; starts the copy of the one after.
        The caller must have finished (synchronised on) the step before the previous one."""
        i = self.next_index
        d_imgs, d_labels, ev = self.slots[i % 2]
        torch.cuda.current_stream(self.device).wait_event(ev)
        self.next_index = i + 1
        self._issue()
        return d_imgs, d_labels, i


class SyntheticLoader:
    def __init__(self, cfg, source=None, epoch_size=None, rank=0, world_size=1, train=True,
                 output="nhwc4_bf16", one_hot=True, device="cuda"):
        self.cfg = cfg
        self.batch_size = cfg.batch_size
        self.num_classes = cfg.num_classes
        self.rank, self.world_size, self.train, self.one_hot = rank, world_size, train, one_hot
        self.device = device
        self.source = source or SyntheticSource(num_classes=cfg.num_classes, device=device)
        self.epoch_size = epoch_size or self.source.pool
        if train:
            self.augment = GpuAugment(cfg.image_size, getattr(cfg, "min_area", 0.08), 1.0,
                                      seed=getattr(cfg, "seed", 0), flip=True, output=output)
        else:
            self.augment = GpuValTransform(cfg.image_size, getattr(cfg, "full_crop", False), output)
        self.pixel_aug = BatchPixelAug(cfg, seed=getattr(cfg, "seed", 0)) if train else None
        self._epoch = 0

    def __len__(self):
        # drop-last, sharded by rank like the DALI readers (dali_dataloader.py:47,175)
        return self.epoch_size // (self.batch_size * self.world_size)

    def __iter__(self):
        n = len(self)
        for i in range(n):
            gi = (self._epoch * n + i) * self.world_size + self.rank
            imgs, labels = self.source.batch(gi, self.batch_size)
            if not imgs.is_cuda:
                imgs = imgs.to(self.device, non_blocking=True)
                labels = labels.to(self.device, non_blocking=True)
            if self.pixel_aug is not None and self.pixel_aug.active:
                data, bx = self.augment(imgs, first_sample=gi * self.batch_size, return_boxes=True)
                data = self.pixel_aug(data, bx, gi)
            else:
                data = self.augment(imgs, first_sample=gi * self.batch_size)
            target = ops.one_hot(labels, self.num_classes) if self.one_hot else labels
            yield data, target
        self._epoch += 1


class RecordLoader:
    """(data, target) batches from real images: a `records.TFRecordReader` / `records.FileReader`
    (sharded by rank like the DALI readers, dali_dataloader.py:47) -> host decode -> ONE pinned
    ragged buffer -> H2D -> crop boxes / resample on the device (`sib_*_ragged`) -> one-hot targets.
    Same interface as `SyntheticLoader` (`batch_size`, `__len__`, iteration; drop-last)."""

    def __init__(self, cfg, reader, train=True, output="nhwc4_bf16", one_hot=True, device="cuda",
                 decode_workers=8, decode=None, prefetch=None):
        from . import jpeg, records
        self._records, self._jpeg = records, jpeg
        # "device": hybrid JPEG decode like the reference's device="mixed" decoders (dali_dataloader.py:65-72,
        # 140-145) -- Huffman stage on host threads, IDCT / upsampling / colour conversion on the GPU
        # (jpeg.decode_batch); "host": the whole decoder on host threads (PIL).  Same pixels either way.
        # Default: "device" on a CUDA device.
        if decode is None:
            decode = "device" if str(device).startswith("cuda") else "host"
        if decode not in ("device", "host"):
            raise ValueError("decode must be 'device' or 'host'")
        self.decode = decode
        # prefetch > 0: batches prepared ahead of the consumer by a loader thread on its own CUDA stream (the
        # reference's DALI pipelines run with prefetch_queue_depth 2).  Default 0 = prepared in the caller:
        # with the step replayed from a CUDA graph the caller's thread is idle while the GPU runs, so the host
        # Huffman stage of batch i+1 already overlaps the step of batch i, and the extra thread only adds GIL
        # contention (measured, 16 cores: 10.3 k images/s synchronous, 9.2 k with the thread); worth it when
        # the consumer itself is host-bound (eager launches, heavy callbacks).
        self.prefetch = int(prefetch or 0)
        # (Measured on a 16-core box per 256 files: reads 7-9 ms, Huffman stage 12.6-13.8 ms on 16 threads, plan +
        #  boxes + parse 2 ms.  Moving the reads into the decode pool (10.1 ms) or into a reader thread running ahead
        #  (9.0 k images/s against 9.8 k without) did not help: the GIL serialises the many small system calls and
        #  the extra thread competes with the caller; not kept.)
        self._pool = None
        self.cfg, self.reader, self.train, self.one_hot = cfg, reader, train, one_hot
        self.batch_size, self.num_classes, self.device = cfg.batch_size, cfg.num_classes, device
        self.image_size = cfg.image_size
        self.out_mode = 0 if output == "nhwc4_bf16" else 1
        self.min_area, self.seed = getattr(cfg, "min_area", 0.08), getattr(cfg, "seed", 0)
        full_crop = getattr(cfg, "full_crop", False)
        self.crop_size = self.image_size if full_crop else math.ceil((self.image_size * 1.14 + 8) // 16 * 16)
        self.decode_workers = decode_workers
        self.pixel_aug = BatchPixelAug(cfg, seed=self.seed) if train else None
        self._batches = 0
        # global sample counter of the crop / flip Philox stream: epoch * dataset + position, offset by
        # the shard so that ranks draw different randoms; carried across loader rebuilds by DataManager
        self.rank = getattr(reader, "shard_id", 0)
        self.world = max(getattr(reader, "num_shards", 1), 1)
        self._seen = 0
        self._epoch = 0

    def __len__(self):
        return len(self.reader) // self.batch_size

    def _workers(self):
        if self._pool is None and self.decode_workers > 1:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=self.decode_workers, thread_name_prefix="sib-decode")
        return self._pool

    def _emit(self, samples):
        dev = self.device
        first = (self._epoch * len(self.reader) + self._seen) * self.world + self.rank * len(samples)
        boxes = None
        if self.decode == "device" and self.train:
            # the crop boxes only need the image sizes (known from the headers): computed on the host by the C
            # twin of the device routine (bit-identical) BEFORE decoding, so that every stream is entropy-decoded
            # only down to the row its crop ends in
            def crop_fn(hw):
                return [ops.rrc_box_host(int(h), int(w), self.min_area, 1.0, self.seed, first + i)
                        for i, (h, w) in enumerate(hw.tolist())]
            buf, offsets, dims, labels, boxes = self._jpeg.decode_batch(samples, workers=self.decode_workers,
                                                                        device=dev, crop_fn=crop_fn,
                                                                        pool=self._workers())
            boxes = boxes.to(dev, non_blocking=True)
        elif self.decode == "device":
            buf, offsets, dims, labels = self._jpeg.decode_batch(samples, workers=self.decode_workers, device=dev,
                                                                 pool=self._workers())
        else:
            buf, offsets, dims, labels = self._records.decode_batch(samples, workers=self.decode_workers,
                                                                    pinned=torch.cuda.is_available())
        buf, offsets, dims = (t.to(dev, non_blocking=True) for t in (buf, offsets, dims))
        labels = labels.to(dev, non_blocking=True)
        if self.train:
            if boxes is None:
                boxes = ops.rrc_boxes_ragged(dims, self.min_area, 1.0, self.seed, first, True)
            data = ops.augment_ragged(buf, offsets, dims, boxes, self.image_size, DATA_MEAN, DATA_STD,
                                      self.out_mode)
            if self.pixel_aug is not None and self.pixel_aug.active:
                data = self.pixel_aug(data, boxes, self._batches * self.world + self.rank)
            self._batches += 1
        else:
            data = ops.val_transform_ragged(buf, offsets, dims, self.image_size, self.crop_size,
                                            DATA_MEAN, DATA_STD, self.out_mode)
        self._seen += len(samples)
        target = ops.one_hot(labels, self.num_classes) if self.one_hot else labels
        return data, target

    def _sample_batches(self):
        batch = []
        for sample in self.reader:
            batch.append(sample)
            if len(batch) == self.batch_size:
                yield batch
                batch = []
        # the ragged tail is dropped (LastBatchPolicy.DROP, dali_dataloader.py:175)

    @staticmethod
    def _in_thread(make_iter, depth, name):
        """Run the iterator `make_iter()` in a daemon thread, `depth` items ahead of the consumer; exceptions
        surface in the consumer, a consumer that stops early releases the thread."""
        import queue
        import threading
        q = queue.Queue(maxsize=depth)
        stop = threading.Event()

        def put(item):
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.1)
                    return True
                except queue.Full:
                    pass
            return False

        def produce():
            try:
                for item in make_iter():
                    if stop.is_set() or not put((item,)):
                        return
                put(None)
            except BaseException as e:
                put(e)

        worker = threading.Thread(target=produce, name=name, daemon=True)
        worker.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                yield item[0]
        finally:
            stop.set()
            while worker.is_alive():            # unblock a producer waiting on a full queue
                try:
                    q.get_nowait()
                except queue.Empty:
                    pass
                worker.join(timeout=0.05)

    def __iter__(self):
        self._seen = 0
        if hasattr(self.reader, "epoch"):
            self.reader.epoch = self._epoch      # a steps_per_epoch / debug break must not replay the order
        try:
            if self.prefetch <= 0:
                for batch in self._sample_batches():
                    yield self._emit(batch)
            else:
                yield from self._iter_prefetched()
        finally:
            self._epoch += 1

    def _iter_prefetched(self):
        side = torch.cuda.Stream(device=self.device)

        def batches():
            with torch.cuda.device(side.device), torch.cuda.stream(side):
                for batch in self._sample_batches():
                    out = self._emit(batch)
                    done = torch.cuda.Event()
                    done.record(side)
                    yield out, done

        for (data, target), done in self._in_thread(batches, self.prefetch, "sib-record-loader"):
            cur = torch.cuda.current_stream()
            cur.wait_event(done)
            data.record_stream(cur)
            target.record_stream(cur)
            yield data, target


def real_data_root(cfg):
    """cfg.root_data_dir (default "${env:IMAGENET_DIR}", arg_parser.py:25) if it names a directory
    that holds the reference's layout, else None (-> synthetic data)."""
    root = getattr(cfg, "root_data_dir", None)
    if not root:
        return None
    if root.startswith("${env:") and root.endswith("}"):
        root = os.environ.get(root[6:-1], "")
    if not root or not os.path.isdir(root):
        return None
    sub = "train_records" if getattr(cfg, "use_tfrecords", False) else "train"
    return root if os.path.isdir(os.path.join(root, sub)) else None


def make_reader(cfg, root, split, rank=0, world_size=1):
    """TFRecord shards + DALI indexes (`use_tfrecords`) or class folders, sharded by rank; the
    train split is shuffled every epoch (dali_dataloader.py:47), validation is read in order (:130)."""
    from . import records
    kw = dict(shard_id=rank, num_shards=world_size, random_shuffle=(split == "train"),
              seed=getattr(cfg, "seed", 0))
    if getattr(cfg, "use_tfrecords", False):
        return records.TFRecordReader.from_root(root, split, **kw)
    return records.FileReader(os.path.join(root, split), **kw)


class DataManager:
    """Stage list semantics of DaliDataManager (dali_dataloader.py:189-239)."""

    def __init__(self, cfg, source=None, rank=0, world_size=1, **loader_kw):
        self.cfg = cfg
        self.stages = cfg.run.stages
        self.tot_epochs = max(stage.end for stage in self.stages)
        self._validate_stages()
        self.source, self.rank, self.world_size, self.loader_kw = source, rank, world_size, loader_kw
        self.loader = None
        self.val_loader = None
        self.start_epoch = None
        self.end_epoch = None

    def __len__(self):
        return len(self.stages)

    def _validate_stages(self):
        end = 0
        for stage in self.stages:
            assert stage.start == end, "error in data stages. start != end"
            assert stage.end > stage.start, "error in data stages, end <= start"
            end = stage.end

    def set_stage(self, idx):
        self.start_epoch = self.stages[idx].start
        self.end_epoch = self.stages[idx].end
        if self.stages[idx].extra_args is None and self.loader is not None:
            return   # only the learning rate changed
        train_cfg = copy.deepcopy(self.cfg.loader)
        val_cfg = copy.deepcopy(self.cfg.val_loader)
        if self.stages[idx].extra_args is not None:
            for key, value in self.stages[idx].extra_args.items():
                setattr(train_cfg, key, value)
        val_cfg.image_size = train_cfg.image_size
        unsupported = [k for k in ("random_interpolation",) if getattr(train_cfg, k, 0)]
        if unsupported:
            import warnings
            warnings.warn("sota_imagenet_b200.data: augmentation fields %s of the reference train_pipeline "
                          "(dali_dataloader.py:75-79: a coin flip between cubic and triangular resampling) are not implemented "
                          "and are ignored" % unsupported)
        root = real_data_root(train_cfg) if self.source is None else None
        prev_epoch = getattr(self.loader, "_epoch", 0) if self.loader is not None else 0
        if root is not None:
            # real images on disk, laid out like the reference expects (dali_dataloader.py:46-65)
            self.loader = RecordLoader(train_cfg, make_reader(train_cfg, root, "train", self.rank, self.world_size),
                                       train=True, **self.loader_kw)
            self.val_loader = RecordLoader(val_cfg, make_reader(val_cfg, root, "val", self.rank, self.world_size),
                                           train=False, **self.loader_kw)
            self.loader._epoch = prev_epoch          # progressive resizing must not replay epoch 0
            return
        self.loader = SyntheticLoader(train_cfg, self.source, rank=self.rank,
                                      world_size=self.world_size, train=True, **self.loader_kw)
        self.val_loader = SyntheticLoader(val_cfg, self.source, rank=self.rank,
                                          world_size=self.world_size, train=False, **self.loader_kw)
        self.loader._epoch = prev_epoch
