"""Host side of the real-data ingest: the on-disk formats in front of the GPU
augmentation kernels.

  * TFRecord shards as written by the reference's `sota_imagenet/create_records.py:84-106`
    (`tf.train.Example` with `image/encoded`, `image/class/label`, `image/filename`) plus the DALI
    index files `tfrecord2idx` produces (`create_records.py:106`; one "offset size" line per record),
    which is what `fn.readers.tfrecord(path=..., index_path=...)` consumes
    (`dali_dataloader.py:49-63`);
  * the class-per-folder layout read by `fn.readers.file(file_root=...)` (`dali_dataloader.py:65`);
  * sharding by (shard_id, num_shards) and the epoch shuffle of the DALI readers
    (`dali_dataloader.py:47`: `random_shuffle=True, shard_id=rank, num_shards=world`).

Pure Python + numpy (no tensorflow, no protobuf package): the wire formats are small enough to state
here.  JPEG decoding uses PIL (or OpenCV) on host threads and packs a batch into ONE pinned uint8
buffer with per-image offsets / sizes, ready for a single H2D copy; `data.RecordLoader` hands it
to the ragged kernels (`sib_rrc_boxes_ragged`, `sib_augment_ragged`, `sib_val_transform_ragged`).
`decode_batch(..., canvas=(H, W))` letterboxes onto a fixed canvas for the uniform-shape kernels.
"""
import io
import os
import struct
import threading

import numpy as np

# ------------------------------------------------------------------ CRC32C (Castagnoli), TFRecord framing
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        poly = 0x82F63B78
        tab = np.zeros(256, dtype=np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ poly if c & 1 else c >> 1
            tab[i] = c
        _CRC_TABLE = [int(v) for v in tab]
    return _CRC_TABLE


def _crc32c_py(data: bytes) -> int:
    """Per-byte table walk (the specification; used to pin the native routine in the tests)."""
    tab = _crc_table()
    c = 0xFFFFFFFF
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


_CRC_NATIVE = None


def crc32c(data: bytes) -> int:
    """CRC-32C through the library's host routine (slicing-by-8, sib_crc32c_host): verify=True at
    ImageNet scale is ~1 GB/s per core instead of ~2 MB/s for the Python loop."""
    global _CRC_NATIVE
    if _CRC_NATIVE is None:
        try:
            import ctypes
            from . import _lib
            fn = _lib.load().sib_crc32c_host

            def native(buf):
                b = bytes(buf)
                out = ctypes.c_uint()            # per call: readers verify records from several threads
                _lib.check(fn(b, len(b), ctypes.byref(out)))
                return out.value
            native(b"")
            _CRC_NATIVE = native
        except Exception:        # library not built (e.g. no nvcc on a data-preparation host)
            _CRC_NATIVE = _crc32c_py
    return _CRC_NATIVE(data)


def masked_crc(data: bytes) -> int:
    """TFRecord's masked CRC: rotate right by 15 and add a constant."""
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ------------------------------------------------------------------ protobuf wire format (the subset Example uses)
def _varint(n: int) -> bytes:
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf, pos):
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _ld(field: int, payload: bytes) -> bytes:
    """length-delimited field"""
    return _varint(field << 3 | 2) + _varint(len(payload)) + payload


def _fields(buf):
    """yield (field number, wire type, value) of one message; value is bytes for wire type 2."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _read_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _read_varint(buf, pos)
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            val = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 1:
            val = bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 5:
            val = bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield field, wt, val


def encode_example(features: dict) -> bytes:
    """tf.train.Example{features: Features{feature: map<string, Feature>}}; values are bytes
    (BytesList, Feature field 1), int or list of int (Int64List, field 3), float lists (field 2)."""
    entries = b""
    for key in sorted(features):
        v = features[key]
        if isinstance(v, (bytes, bytearray)):
            feat = _ld(1, _ld(1, bytes(v)))
        elif isinstance(v, (int, np.integer)) or (isinstance(v, (list, tuple)) and all(
                isinstance(x, (int, np.integer)) for x in v)):
            vals = [v] if isinstance(v, (int, np.integer)) else list(v)
            feat = _ld(3, _ld(1, b"".join(_varint(int(x)) for x in vals)))      # packed int64
        else:
            vals = [float(x) for x in (v if isinstance(v, (list, tuple)) else [v])]
            feat = _ld(2, _ld(1, struct.pack("<%df" % len(vals), *vals)))       # packed float
        entries += _ld(1, _ld(1, key.encode()) + _ld(2, feat))                 # map entry
    return _ld(1, entries)


def parse_example(buf: bytes) -> dict:
    """-> {name: bytes | [int] | [float]} (first bytes value of a BytesList, all ints / floats)."""
    out = {}
    for f, _, features in _fields(buf):
        if f != 1:
            continue
        for f2, _, entry in _fields(features):
            if f2 != 1:
                continue
            key, feat = None, b""
            for f3, _, v in _fields(entry):
                if f3 == 1:
                    key = v.decode()
                elif f3 == 2:
                    feat = v
            for kind, _, lst in _fields(feat):
                if kind == 1:       # BytesList
                    vals = [v for f4, _, v in _fields(lst) if f4 == 1]
                    out[key] = vals[0] if len(vals) == 1 else vals
                elif kind == 3:     # Int64List: packed or repeated varints
                    vals = []
                    for f4, wt, v in _fields(lst):
                        if f4 != 1:
                            continue
                        if wt == 2:
                            p = 0
                            while p < len(v):
                                x, p = _read_varint(v, p)
                                vals.append(x - (1 << 64) if x >> 63 else x)
                        else:
                            vals.append(v - (1 << 64) if v >> 63 else v)
                    out[key] = vals
                elif kind == 2:     # FloatList
                    vals = []
                    for f4, wt, v in _fields(lst):
                        if f4 == 1:
                            vals += list(struct.unpack("<%df" % (len(v) // 4), v))
                    out[key] = vals
    return out


# ------------------------------------------------------------------ TFRecord files + DALI index files
def write_tfrecord(path, records):
    """records: iterable of serialized byte strings.  Returns [(offset, size)] like tfrecord2idx."""
    index, off = [], 0
    with open(path, "wb") as f:
        for rec in records:
            head = struct.pack("<Q", len(rec))
            blob = head + struct.pack("<I", masked_crc(head)) + rec + struct.pack("<I", masked_crc(rec))
            f.write(blob)
            index.append((off, len(blob)))
            off += len(blob)
    return index


def build_index(path):
    """Scan a TFRecord file -> [(offset, size)] (what `tfrecord2idx` writes, create_records.py:106)."""
    index, off = [], 0
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        while off < size:
            f.seek(off)
            head = f.read(8)
            if len(head) < 8:
                raise ValueError("truncated TFRecord header at offset %d of %s" % (off, path))
            (ln,) = struct.unpack("<Q", head)
            total = 8 + 4 + ln + 4
            if off + total > size:
                raise ValueError("truncated TFRecord payload at offset %d of %s" % (off, path))
            index.append((off, total))
            off += total
    return index


def write_index(index, idx_path):
    with open(idx_path, "w") as f:
        for off, size in index:
            f.write("%d %d\n" % (off, size))


def read_index(idx_path):
    out = []
    with open(idx_path) as f:
        for line in f:
            parts = line.split()
            if parts:
                out.append((int(parts[0]), int(parts[1])))
    return out


def read_record(f, offset, size, verify=False):
    """One framed record at (offset, size) of an open binary file -> payload bytes."""
    if isinstance(f, int):
        blob = os.pread(f, size, offset)         # a file descriptor: positional read, safe from any thread
    else:
        f.seek(offset)
        blob = f.read(size)
    if len(blob) != size:
        raise ValueError("short read: record at %d wants %d bytes" % (offset, size))
    (ln,) = struct.unpack("<Q", blob[:8])
    if ln != size - 16:
        raise ValueError("index / file mismatch at offset %d: length field %d, index size %d" % (offset, ln, size))
    payload = blob[12:12 + ln]
    if verify:
        if struct.unpack("<I", blob[8:12])[0] != masked_crc(blob[:8]):
            raise ValueError("corrupt TFRecord length CRC at offset %d" % offset)
        if struct.unpack("<I", blob[12 + ln:16 + ln])[0] != masked_crc(payload):
            raise ValueError("corrupt TFRecord data CRC at offset %d" % offset)
    return payload


# ------------------------------------------------------------------ readers with DALI's sharding semantics
def shard_range(n, shard_id, num_shards):
    """DALI readers: shard i owns samples [floor(n*i/S), floor(n*(i+1)/S))."""
    if not 0 <= shard_id < num_shards:
        raise ValueError("shard_id %d outside [0, %d)" % (shard_id, num_shards))
    return n * shard_id // num_shards, n * (shard_id + 1) // num_shards


class _ShardedReader:
    """Common part: a global sample list, this rank's contiguous shard of it, optional per-epoch
    shuffle inside the shard (seeded: every epoch a new permutation, reproducible)."""

    def __init__(self, n, shard_id=0, num_shards=1, random_shuffle=False, seed=0):
        self.n_total = n
        self.shard_id, self.num_shards = shard_id, num_shards
        self.lo, self.hi = shard_range(n, shard_id, num_shards)
        self.random_shuffle, self.seed, self.epoch = random_shuffle, seed, 0

    def __len__(self):
        return self.hi - self.lo

    def order(self):
        idx = np.arange(self.lo, self.hi)
        if self.random_shuffle:
            np.random.RandomState((self.seed * 1000003 + self.epoch) & 0x7FFFFFFF).shuffle(idx)
        return idx

    def __iter__(self):
        for i in self.order():
            yield self.sample(int(i))
        self.epoch += 1


class TFRecordReader(_ShardedReader):
    """`fn.readers.tfrecord(path=records, index_path=indexes, features={image/encoded,
    image/class/label})` (dali_dataloader.py:49-63): yields (jpeg bytes, int label)."""

    def __init__(self, records, indexes=None, verify=False, **kw):
        self.paths = [str(p) for p in records]
        if indexes is None:
            per_file = [build_index(p) for p in self.paths]
        else:
            if len(indexes) != len(self.paths):
                raise ValueError("%d record files but %d index files" % (len(self.paths), len(indexes)))
            per_file = [read_index(str(p)) for p in indexes]
        self.table = [(fi, off, size) for fi, idx in enumerate(per_file) for off, size in idx]
        self.verify = verify
        self._files = {}                 # file index -> descriptor (os.pread: no shared file position)
        self._open_lock = threading.Lock()
        super().__init__(len(self.table), **kw)

    @classmethod
    def from_root(cls, root, split="train", **kw):
        """records / indexes directories as create_records.py:112-113 names them."""
        rec_dir, idx_dir = os.path.join(root, split + "_records"), os.path.join(root, split + "_indexes")
        recs = sorted(os.path.join(rec_dir, f) for f in os.listdir(rec_dir))
        idxs = sorted(os.path.join(idx_dir, f) for f in os.listdir(idx_dir))
        return cls(recs, idxs, **kw)

    def sample(self, i):
        fi, off, size = self.table[i]
        f = self._files.get(fi)
        if f is None:
            with self._open_lock:
                f = self._files.get(fi)
                if f is None:
                    f = self._files[fi] = os.open(self.paths[fi], os.O_RDONLY)
        ex = parse_example(read_record(f, off, size, self.verify))
        if "image/encoded" not in ex:
            raise ValueError("record %d of %s has no image/encoded feature" % (i, self.paths[fi]))
        label = ex.get("image/class/label", [-1])        # default -1 like the DALI feature spec (:55)
        return ex["image/encoded"], int(label[0]) if label else -1

    def close(self):
        for f in self._files.values():
            os.close(f)
        self._files = {}


class FileReader(_ShardedReader):
    """`fn.readers.file(file_root=root/train)`: one sub-directory per class, label = rank of the
    directory name in sorted order (create_records.py:147-148 uses the same map)."""
    EXTENSIONS = (".jpg", ".jpeg", ".png", ".bmp", ".JPEG", ".JPG", ".PNG")

    def __init__(self, file_root, **kw):
        classes = sorted(d for d in os.listdir(file_root) if os.path.isdir(os.path.join(file_root, d)))
        if not classes:
            raise ValueError("no class directories under %s" % file_root)
        self.classes = classes
        self.samples = []
        for label, c in enumerate(classes):
            d = os.path.join(file_root, c)
            for name in sorted(os.listdir(d)):
                if name.endswith(self.EXTENSIONS):
                    self.samples.append((os.path.join(d, name), label))
        if not self.samples:
            raise ValueError("no images in the class directories of %s" % file_root)
        super().__init__(len(self.samples), **kw)

    def sample(self, i):
        path, label = self.samples[i]
        with open(path, "rb") as f:
            return f.read(), label


# ------------------------------------------------------------------ decode + pack
def decode_image(data: bytes) -> np.ndarray:
    """JPEG / PNG bytes -> uint8 [H, W, 3] RGB (`output_type=types.RGB`, dali_dataloader.py:71,144;
    CMYK and grey images are converted like DALI's decoder does)."""
    from PIL import Image
    with Image.open(io.BytesIO(data)) as im:
        return np.asarray(im.convert("RGB"), dtype=np.uint8)


def pack_batch(images, pinned=False):
    """list of uint8 [H_i, W_i, 3] -> (flat uint8 buffer, int64 offsets [B], int32 dims [B, 2]):
    one contiguous (optionally pinned) host buffer for a single H2D copy of a ragged batch."""
    import torch
    dims = np.array([im.shape[:2] for im in images], dtype=np.int32).reshape(-1, 2)
    sizes = dims[:, 0].astype(np.int64) * dims[:, 1] * 3
    pad = (-sizes) % 16                                   # keep every image 16-byte aligned
    offsets = np.concatenate([[0], np.cumsum(sizes + pad)[:-1]]).astype(np.int64) if len(images) else \
        np.zeros(0, np.int64)
    total = int((sizes + pad).sum())
    buf = torch.empty(max(total, 1), dtype=torch.uint8)
    if pinned:
        buf = buf.pin_memory()
    flat = buf.numpy()
    for im, off, n in zip(images, offsets, sizes):
        if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3:
            raise ValueError("images must be uint8 [H, W, 3]")
        flat[off:off + n] = np.ascontiguousarray(im).reshape(-1)
    return buf, torch.from_numpy(offsets), torch.from_numpy(dims)


def letterbox(im, height, width, fill=0):
    """Centre an image on a fixed [height, width] canvas (cropping what does not fit): lets the
    uniform-shape kernels run on real images until the ragged entry points exist."""
    out = np.full((height, width, 3), fill, dtype=np.uint8)
    h, w = im.shape[:2]
    ch, cw = min(h, height), min(w, width)
    sy, sx = (h - ch) // 2, (w - cw) // 2
    dy, dx = (height - ch) // 2, (width - cw) // 2
    out[dy:dy + ch, dx:dx + cw] = im[sy:sy + ch, sx:sx + cw]
    return out


def decode_batch(samples, canvas=None, workers=4, pinned=False):
    """samples: list of (encoded bytes, label).  canvas=None -> ragged pack (buffer, offsets, dims,
    labels); canvas=(H, W) -> uniform uint8 [B, H, W, 3] tensor + labels for `GpuAugment`."""
    import torch
    from concurrent.futures import ThreadPoolExecutor

    def one(s):
        return decode_image(s[0]), s[1]

    if workers > 1 and len(samples) > 1:
        with ThreadPoolExecutor(max_workers=workers) as ex:      # PIL releases the GIL while decoding
            done = list(ex.map(one, samples))
    else:
        done = [one(s) for s in samples]
    images = [d[0] for d in done]
    labels = torch.tensor([d[1] for d in done], dtype=torch.int64)
    if canvas is None:
        return pack_batch(images, pinned) + (labels,)
    h, w = canvas
    out = torch.empty((len(images), h, w, 3), dtype=torch.uint8)
    if pinned:
        out = out.pin_memory()
    dst = out.numpy()
    for i, im in enumerate(images):
        dst[i] = letterbox(im, h, w)
    return out, labels
