"""Hybrid JPEG decode of a batch: host Huffman stage -> device IDCT / upsampling / colour conversion.

What the reference reaches through `fn.decoders.image_random_crop(..., device="mixed")` /
`fn.decoders.image(device="mixed")` (dali_dataloader.py:65-72, 140-145): nvJPEG's hybrid back end
entropy-decodes on the host and runs the rest of the decoder on the GPU.  Same split here (`csrc/jpeg.cu`):

  parse(data)                      `sib_jpeg_parse`: geometry + quantisation tables, or the reason the
                                   stream is outside the device subset (progressive, CMYK, ...)
  decode_coefficients(data, out)   `sib_jpeg_decode_coefficients`: int16 DCT coefficients (thread pool;
                                   ctypes releases the GIL, the C routine keeps no global state)
  decode_batch(samples)            ONE pinned coefficient buffer + ONE descriptor table -> H2D ->
                                   `sib_jpeg_idct_rgb` writes every image as uint8 [H][W][3] into the packed
                                   ragged buffer (`records.pack_batch` layout) that `sib_rrc_boxes_ragged`
                                   / `sib_augment_ragged` / `sib_val_transform_ragged` consume.

Streams the device path does not take (`JpegInfo.status != 0`: progressive / arithmetic, 12-bit, CMYK,
unusual sampling, PNG files of the ImageNet tree) are decoded by `records.decode_image` on the host and
copied into the same packed buffer; everything downstream is identical.  The pixels are bit-identical
to PIL / libjpeg-turbo either way.
"""
import ctypes
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib


class JpegInfo(ctypes.Structure):
    """sib_jpeg_info (include/sib200.h)"""
    _fields_ = [("coef_count", ctypes.c_long), ("status", ctypes.c_int), ("width", ctypes.c_int),
                ("height", ctypes.c_int), ("ncomp", ctypes.c_int), ("hs", ctypes.c_int * 3),
                ("vs", ctypes.c_int * 3), ("hmax", ctypes.c_int), ("vmax", ctypes.c_int),
                ("mcus_x", ctypes.c_int), ("mcus_y", ctypes.c_int), ("blocks_w", ctypes.c_int * 3),
                ("blocks_h", ctypes.c_int * 3), ("restart_interval", ctypes.c_int),
                ("quant", (ctypes.c_ushort * 64) * 3)]

    def as_dict(self):
        n = self.ncomp
        return dict(width=self.width, height=self.height, ncomp=n, hmax=self.hmax, vmax=self.vmax,
                    blocks_w=list(self.blocks_w)[:n], blocks_h=list(self.blocks_h)[:n],
                    quant=[np.array(self.quant[c], dtype=np.int64) for c in range(n)])


# sib_jpeg_image (include/sib200.h): one record per image of a batch
IMAGE_DTYPE = np.dtype([("coef_off", "<i8", 3), ("plane_off", "<i8", 3), ("out_off", "<i8"),
                        ("width", "<i4"), ("height", "<i4"), ("ncomp", "<i4"), ("hmax", "<i4"), ("vmax", "<i4"),
                        ("blocks_w", "<i4", 3), ("blocks_h", "<i4", 3), ("mcu_rows", "<i4"),
                        ("quant", "<u2", (3, 64))], align=True)
assert IMAGE_DTYPE.itemsize == 488

STATUS = {0: "ok", 1: "not a JPEG", 2: "corrupt", 3: "progressive / arithmetic / lossless process",
          4: "not 8-bit", 5: "colour space (CMYK, Adobe RGB)", 6: "sampling factors", 7: "more than one scan",
          8: "larger than MAX_PIXELS"}
# a (corrupt) header may announce up to 65535 x 65535 pixels: such streams go to the host decoder, which has its
# own decompression-bomb guard, instead of sizing a multi-gigabyte pinned staging buffer here
MAX_PIXELS = 1 << 26


def _as_bytes(data):
    # ctypes passes a `bytes` object as a pointer to its buffer: no copy, no per-call array type
    return data if isinstance(data, bytes) else bytes(data)


def parse(data):
    """bytes -> JpegInfo; `status == 0` means the device path decodes this stream."""
    info = JpegInfo()
    data = _as_bytes(data)
    _lib.check(_lib.load().sib_jpeg_parse(data, len(data), ctypes.byref(info)))
    if info.status == 0 and info.width * info.height > MAX_PIXELS:
        info.status = 8
    return info


def mcu_rows_for(info, luma_rows):
    """MCU rows a consumer that reads the luma rows [0, luma_rows) needs: two rows of margin because the fancy
    chroma upsampling of a row interpolates against the next chroma row; 0 = the whole image."""
    need = -(-(luma_rows + 2) // (8 * info.vmax))
    return 0 if need >= info.mcus_y else need


def decode_coefficients(data, out=None, info=None, mcu_rows=0):
    """Host Huffman stage: bytes -> int16 coefficients (flat numpy array, or written into `out`); `mcu_rows` > 0
    stops after that many rows of MCUs (the rest of `out` is left untouched)."""
    info = info or parse(data)
    if info.status != 0:
        raise _lib.SibError("jpeg: stream not decodable on the device path (%s)" % STATUS.get(info.status, info.status))
    if out is None:
        out = np.empty(info.coef_count, dtype=np.int16)
    assert out.dtype == np.int16 and out.size >= info.coef_count and out.flags.c_contiguous
    data = _as_bytes(data)
    _lib.check(_lib.load().sib_jpeg_decode_coefficients_rows(data, len(data), ctypes.c_void_p(out.ctypes.data),
                                                             int(mcu_rows)))
    return out


def plan_batch(infos, fallback_dims):
    """Layout of one batch: per-image descriptor table (IMAGE_DTYPE), total coefficients, scratch bytes,
    packed-output offsets / dims (the `records.pack_batch` layout: every image 16-byte aligned)."""
    n = len(infos)
    table = np.zeros(n, dtype=IMAGE_DTYPE)
    coef_off, plane_off, out_off = table["coef_off"], table["plane_off"], table["out_off"]
    bws, bhs, quant = table["blocks_w"], table["blocks_h"], table["quant"]
    dims = np.zeros((n, 2), dtype=np.int32)
    offsets = np.zeros(n, dtype=np.int64)
    head = np.zeros((n, 5), dtype=np.int32)               # width, height, ncomp, hmax, vmax
    coef_total = plane_total = out_total = 0
    max_blocks = max_pixels = 0
    for i, info in enumerate(infos):
        if info is None or info.status != 0:
            h, w = fallback_dims[i]
        else:
            h, w, nc = info.height, info.width, info.ncomp
            head[i] = (w, h, nc, info.hmax, info.vmax)
            bw, bh = info.blocks_w[:nc], info.blocks_h[:nc]
            blocks = 0
            for c in range(nc):
                coef_off[i, c] = plane_off[i, c] = coef_total       # (one byte of plane per coefficient)
                coef_total += bw[c] * bh[c] * 64
                blocks += bw[c] * bh[c]
            bws[i, :nc], bhs[i, :nc] = bw, bh
            quant[i] = np.frombuffer(info.quant, dtype=np.uint16).reshape(3, 64)
            out_off[i] = out_total
            max_blocks, max_pixels = max(max_blocks, blocks), max(max_pixels, h * w)
        dims[i] = (h, w)
        offsets[i] = out_total
        size = h * w * 3
        out_total += size + (-size) % 16
    plane_total = coef_total
    for k, name in enumerate(("width", "height", "ncomp", "hmax", "vmax")):
        table[name] = head[:, k]
    return dict(table=table, dims=dims, offsets=offsets, coef_total=coef_total, plane_total=plane_total,
                out_total=out_total, max_blocks=max_blocks, max_pixels=max_pixels)


class _PinnedRing:
    """Two grow-only pinned int16 staging buffers used alternately: `cudaHostAlloc` of ~140 MB per batch costs
    more than the Huffman stage itself, and a buffer may only be rewritten once the H2D copy that read it has
    completed (its event is waited for on the host before reuse)."""

    def __init__(self):
        self.slots = [[None, None], [None, None]]      # (tensor, event of the last copy out of it)
        self.next = 0

    def take(self, n):
        import torch
        slot = self.slots[self.next]
        self.next ^= 1
        if slot[1] is not None:
            slot[1].synchronize()
        if slot[0] is None or slot[0].numel() < n:
            slot[0] = torch.empty(max(n, 1), dtype=torch.int16).pin_memory()
        return slot

    def copied(self, slot):
        import torch
        slot[1] = torch.cuda.Event()
        slot[1].record()


_RING = _PinnedRing()


def decode_batch(samples, workers=8, device="cuda", host_decode=None, crop_fn=None, pool=None):
    """samples: list of (encoded bytes, label) -> (packed uint8 device buffer, offsets, dims, labels), the
    tuple `records.decode_batch(..., canvas=None)` returns, with the buffer already resident on `device`.
    `host_decode(bytes) -> uint8 [H, W, 3]` takes the streams outside the device subset.

    `crop_fn(dims int32 [B, 2]) -> boxes int32 [B, 5] {x0, y0, w, h, flip}` (host): decode only what the crops
    read -- the image sizes come from the headers, so the boxes exist before any entropy decoding starts and
    every stream is decoded only down to the MCU row its crop ends in (ROI decoding of
    `fn.decoders.image_random_crop`, dali_dataloader.py:65-72); pixels below that row are left undefined.
    Returns the boxes as a fifth element.  `pool`: a persistent ThreadPoolExecutor (else one is made per call)."""
    import torch
    from . import ops, records
    _lib.require_device()
    host_decode = host_decode or records.decode_image
    own_pool = pool is None and workers > 1 and len(samples) > 1     # (a loader passes its persistent pool)
    if own_pool:
        pool = ThreadPoolExecutor(max_workers=workers)

    def fetch(s):
        data = _as_bytes(s[0])
        return data, s[1], parse(data)

    try:
        got = list(pool.map(fetch, samples)) if pool else [fetch(s) for s in samples]
    except BaseException:
        if own_pool:
            pool.shutdown()
        raise
    datas, infos = [g[0] for g in got], [g[2] for g in got]
    labels = torch.tensor([g[1] for g in got], dtype=torch.int64)
    on_device = [i for i, inf in enumerate(infos) if inf.status == 0]
    on_host = [i for i, inf in enumerate(infos) if inf.status != 0]
    host_images = {}
    try:
        if on_host:
            dec = list(pool.map(host_decode, [datas[i] for i in on_host])) if pool else \
                [host_decode(datas[i]) for i in on_host]
            host_images = dict(zip(on_host, dec))
        plan = plan_batch(infos, {i: im.shape[:2] for i, im in host_images.items()})
        boxes = None
        if crop_fn is not None:
            boxes = np.ascontiguousarray(crop_fn(plan["dims"]), dtype=np.int32).reshape(len(datas), 5)
            for i in on_device:
                plan["table"]["mcu_rows"][i] = mcu_rows_for(infos[i], int(boxes[i, 1] + boxes[i, 3]))
        slot = _RING.take(plan["coef_total"])
        coef = slot[0][:max(plan["coef_total"], 1)]
        coef_np = coef.numpy()

        def huff(i):
            t = plan["table"][i]
            decode_coefficients(datas[i], coef_np[int(t["coef_off"][0]):], infos[i], int(t["mcu_rows"]))

        if pool:
            list(pool.map(huff, on_device))
        else:
            for i in on_device:
                huff(i)
    finally:
        if own_pool:
            pool.shutdown()
    out = torch.empty(max(plan["out_total"], 1), dtype=torch.uint8, device=device)
    if on_device:
        # compact table of the device-decoded images only (the kernels index it by blockIdx.y)
        table = torch.from_numpy(plan["table"][on_device].view(np.uint8).reshape(len(on_device), -1).copy())
        table_dev = table.pin_memory().to(device, non_blocking=True)
        coef_dev = coef.to(device, non_blocking=True)
        _RING.copied(slot)
        planes = torch.empty(max(plan["plane_total"], 1), dtype=torch.uint8, device=device)
        ops.jpeg_idct_rgb(coef_dev, table_dev, len(on_device), plan["max_blocks"], plan["max_pixels"], planes, out)
    for i, im in host_images.items():
        flat = torch.from_numpy(np.array(im, dtype=np.uint8).reshape(-1))      # (a writable copy)
        off = int(plan["offsets"][i])
        out[off:off + flat.numel()].copy_(flat, non_blocking=False)
    res = (out, torch.from_numpy(plan["offsets"]), torch.from_numpy(plan["dims"]), labels)
    return res if crop_fn is None else res + (torch.from_numpy(boxes),)
