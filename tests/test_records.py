"""CPU: host side of the real-data ingest (sota_imagenet_b200/records.py) — TFRecord framing and
`tf.train.Example` wire format as the reference's create_records.py:84-106 writes them, DALI index
files, reader sharding / shuffling semantics (dali_dataloader.py:47-65), decode + ragged packing."""
import io
import struct

import numpy as np
import pytest
import torch

from sota_imagenet_b200 import records


def _jpeg(rng, h, w, mode="RGB", fmt="JPEG"):
    from PIL import Image
    ch = {"RGB": 3, "L": 1, "CMYK": 4}[mode]
    arr = rng.randint(0, 256, size=(h, w, ch) if ch > 1 else (h, w), dtype=np.uint8)
    buf = io.BytesIO()
    Image.fromarray(arr, mode=mode).save(buf, format=fmt, quality=95)
    return buf.getvalue(), arr


def test_crc32c_and_example_known_answers():
    assert records.crc32c(b"123456789") == 0xE3069283            # CRC-32C check value
    assert records.crc32c(b"") == 0
    import numpy as np
    rng = np.random.RandomState(0)
    for n in (1, 7, 8, 9, 63, 64, 1000, 4099):                    # native slicing-by-8 == per-byte walk
        blob = rng.randint(0, 256, size=n, dtype=np.uint8).tobytes()
        assert records.crc32c(blob) == records._crc32c_py(blob), n
    # masked CRC as TFRecord defines it: rotr15(crc) + 0xa282ead8
    c = 0xE3069283
    assert records.masked_crc(b"123456789") == (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF
    # Example{features{feature{"a": int64_list{value: [1]}}}} written out by hand from the
    # protobuf wire format (Int64List is packed)
    want = bytes.fromhex("0a0c0a0a0a01611205" "1a030a0101")
    assert records.encode_example({"a": 1}) == want
    assert records.parse_example(want) == {"a": [1]}
    # unpacked (repeated varint) Int64List, as older writers emit it
    unpacked = bytes.fromhex("0a0b0a090a01611204" "1a020801")
    assert records.parse_example(unpacked) == {"a": [1]}
    ex = records.encode_example({"image/encoded": b"\xff\xd8jpeg", "image/class/label": 999,
                                 "image/filename": b"n01440764_1.JPEG", "neg": -1, "f": [0.5, 2.0]})
    got = records.parse_example(ex)
    assert got["image/encoded"] == b"\xff\xd8jpeg" and got["image/class/label"] == [999]
    assert got["image/filename"] == b"n01440764_1.JPEG" and got["neg"] == [-1] and got["f"] == [0.5, 2.0]


def test_tfrecord_roundtrip_index_and_corruption(tmp_path):
    recs = [records.encode_example({"image/encoded": bytes([i]) * (10 + 7 * i), "image/class/label": i})
            for i in range(9)]
    path = str(tmp_path / "train-0-1.tfrecord")
    index = records.write_tfrecord(path, recs)
    assert index[0] == (0, 16 + len(recs[0])) and index[1][0] == index[0][1]
    assert records.build_index(path) == index                    # what tfrecord2idx would emit
    idx_path = str(tmp_path / "train-0-1.idx")
    records.write_index(index, idx_path)
    assert open(idx_path).readline() == "0 %d\n" % (16 + len(recs[0]))
    assert records.read_index(idx_path) == index
    # framing: little-endian u64 length first
    raw = open(path, "rb").read()
    assert struct.unpack("<Q", raw[:8])[0] == len(recs[0])
    rd = records.TFRecordReader([path], [idx_path], verify=True)
    out = list(rd)
    assert [l for _, l in out] == list(range(9)) and out[3][0] == bytes([3]) * 31
    assert len(rd) == 9 and rd.epoch == 1
    rd.close()
    # corrupt one payload byte: the CRC check catches it, the unchecked reader does not care
    bad = bytearray(raw)
    bad[index[2][0] + 20] ^= 0xFF
    open(path, "wb").write(bytes(bad))
    with pytest.raises(ValueError, match="CRC"):
        list(records.TFRecordReader([path], [idx_path], verify=True))
    # truncated file
    open(path, "wb").write(raw[:-3])
    with pytest.raises(ValueError, match="truncated"):
        records.build_index(path)
    with pytest.raises(ValueError):
        records.TFRecordReader([path], [idx_path, idx_path])


def test_reader_sharding_and_shuffle(tmp_path):
    paths, idxs, n = [], [], 0
    for s in range(3):
        recs = [records.encode_example({"image/encoded": b"x%d" % (n + i), "image/class/label": n + i})
                for i in range(5 + s)]
        n += len(recs)
        p = str(tmp_path / ("val-%d-3.tfrecord" % s))
        records.write_index(records.write_tfrecord(p, recs), p + ".idx")
        paths.append(p)
        idxs.append(p + ".idx")
    assert n == 18
    # shards partition the sample space contiguously (no overlap, nothing dropped)
    seen = []
    for rank in range(4):
        rd = records.TFRecordReader(paths, idxs, shard_id=rank, num_shards=4)
        labels = [l for _, l in rd]
        assert labels == list(range(*records.shard_range(18, rank, 4)))
        seen += labels
    assert seen == list(range(18))
    assert [records.shard_range(10, i, 3) for i in range(3)] == [(0, 3), (3, 6), (6, 10)]
    with pytest.raises(ValueError):
        records.shard_range(10, 3, 3)
    # shuffle: a permutation of the shard, new every epoch, reproducible from the seed
    a = records.TFRecordReader(paths, idxs, shard_id=1, num_shards=2, random_shuffle=True, seed=7)
    e0 = [l for _, l in a]
    e1 = [l for _, l in a]
    b = records.TFRecordReader(paths, None, shard_id=1, num_shards=2, random_shuffle=True, seed=7)   # index rebuilt by scanning
    assert sorted(e0) == list(range(9, 18)) == sorted(e1) and e0 != e1 and e0 != sorted(e0)
    assert [l for _, l in b] == e0
    # records may be read from worker threads (positional reads: no shared file position; per-call CRC output)
    from concurrent.futures import ThreadPoolExecutor
    rd = records.TFRecordReader(paths, idxs, verify=True)
    eager = list(rd)
    with ThreadPoolExecutor(max_workers=4) as ex:
        for _ in range(5):
            assert list(ex.map(rd.sample, range(18))) == eager
    rd.close()


def test_file_reader_and_decode_pack(tmp_path):
    rng = np.random.RandomState(0)
    root = tmp_path / "train"
    want = {}
    for ci, cname in enumerate(["n02", "n01", "n03"]):           # labels follow SORTED directory names
        (root / cname).mkdir(parents=True)
        for k in range(2):
            data, arr = _jpeg(rng, 20 + 4 * ci, 30 + k, fmt="PNG")
            (root / cname / ("img%d.png" % k)).write_bytes(data)
            want[(cname, k)] = arr
    (root / "n01" / "notes.txt").write_text("ignored")
    rd = records.FileReader(str(root))
    assert rd.classes == ["n01", "n02", "n03"] and len(rd) == 6
    samples = list(rd)
    assert [l for _, l in samples] == [0, 0, 1, 1, 2, 2]
    assert np.array_equal(records.decode_image(samples[2][0]), want[("n02", 0)])      # lossless PNG
    # ragged pack: one buffer, 16-byte aligned offsets, exact bytes back
    buf, offsets, dims, labels = records.decode_batch(samples, workers=2)
    assert labels.tolist() == [0, 0, 1, 1, 2, 2] and dims.dtype == torch.int32 and offsets.dtype == torch.int64
    assert all(int(o) % 16 == 0 for o in offsets)
    for i, (cname, k) in enumerate([("n01", 0), ("n01", 1), ("n02", 0), ("n02", 1), ("n03", 0), ("n03", 1)]):
        h, w = dims[i].tolist()
        got = buf[int(offsets[i]):int(offsets[i]) + h * w * 3].numpy().reshape(h, w, 3)
        assert np.array_equal(got, want[(cname, k)])
    # fixed canvas for the uniform kernels: centred, cropped / zero padded
    imgs, labels = records.decode_batch(samples, canvas=(24, 28))
    assert tuple(imgs.shape) == (6, 24, 28, 3) and imgs.dtype == torch.uint8
    src = want[("n01", 0)]                                       # 24 x 30 -> all rows, cols 1..28
    assert src.shape == (24, 30, 3) and np.array_equal(imgs[0].numpy(), src[:, 1:29])
    tall = want[("n03", 1)]                                      # 28 x 31 -> rows 2..25, cols 1..28
    assert tall.shape == (28, 31, 3) and np.array_equal(imgs[5].numpy(), tall[2:26, 1:29])
    small = records.letterbox(np.full((4, 6, 3), 9, np.uint8), 8, 8)
    assert small.sum() == 9 * 4 * 6 * 3 and small[2:6, 1:7].min() == 9 and small[0].max() == 0
    # JPEG, grey and CMYK sources all come out as RGB (create_records.py:71-83 re-encodes the CMYK ones)
    for mode in ("RGB", "L", "CMYK"):
        data, arr = _jpeg(rng, 16, 24, mode=mode)
        im = records.decode_image(data)
        assert im.shape == (16, 24, 3) and im.dtype == np.uint8
    with pytest.raises(ValueError):
        records.FileReader(str(tmp_path))                        # one level too high: no images found


def test_record_loader_host_flow_with_oracle_kernels(tmp_path, monkeypatch):
    """RecordLoader + make_reader + real_data_root on a tiny class-folder data set of images of
    different sizes; the three ragged kernels are replaced by the numpy oracle (the GPU test
    tests/test_gpu_ragged.py runs the same comparison against the CUDA kernels)."""
    from types import SimpleNamespace
    from oracle import augment_ref
    from sota_imagenet_b200 import data, ops
    rng = np.random.RandomState(1)
    root = tmp_path / "imagenet"
    imgs = {}
    for split, per_class in (("train", 3), ("val", 2)):
        for ci, cname in enumerate(["n01", "n02"]):
            (root / split / cname).mkdir(parents=True)
            for k in range(per_class):
                d, arr = _jpeg(rng, 24 + 5 * k + ci, 40 - 3 * k, fmt="PNG")
                (root / split / cname / ("%d.png" % k)).write_bytes(d)
                imgs[(split, ci, k)] = arr

    def unpack(packed, offsets, dims):
        out = []
        for o, (h, w) in zip(offsets.tolist(), dims.tolist()):
            out.append(packed[o:o + h * w * 3].numpy().reshape(h, w, 3))
        return out

    def boxes_ragged(dims, min_area, max_area, seed, first, flip):
        return torch.tensor([augment_ref.rrc_box(h, w, min_area, max_area, seed, first + i)
                             for i, (h, w) in enumerate(dims.tolist())], dtype=torch.int32)

    def aug_ragged(packed, offsets, dims, boxes, size, mean, std, out_mode):
        ims = unpack(packed, offsets, dims)
        return torch.from_numpy(np.stack([augment_ref.augment_image(im, b, size, mean, std)
                                          for im, b in zip(ims, boxes.tolist())]))

    def val_ragged(packed, offsets, dims, size, rs, mean, std, out_mode):
        return torch.from_numpy(np.stack([augment_ref.val_transform_image(im, size, rs, mean, std)
                                          for im in unpack(packed, offsets, dims)]))

    monkeypatch.setattr(ops, "rrc_boxes_ragged", boxes_ragged)
    monkeypatch.setattr(ops, "augment_ragged", aug_ragged)
    monkeypatch.setattr(ops, "val_transform_ragged", val_ragged)
    monkeypatch.setattr(ops, "one_hot", lambda l, n: torch.eye(n)[l])
    cfg = SimpleNamespace(image_size=16, batch_size=2, num_classes=2, min_area=0.3, seed=5,
                          root_data_dir=str(root), use_tfrecords=False)
    assert data.real_data_root(cfg) == str(root)
    monkeypatch.setenv("SIB_TEST_ROOT", str(root))
    assert data.real_data_root(SimpleNamespace(root_data_dir="${env:SIB_TEST_ROOT}")) == str(root)
    assert data.real_data_root(SimpleNamespace(root_data_dir="${env:SIB_NOT_SET_ANYWHERE}")) is None
    assert data.real_data_root(SimpleNamespace(root_data_dir=str(root), use_tfrecords=True)) is None
    # validation: in order, resize-shorter (crop_size ceil((16*1.14+8)//16*16) = 16) + centre crop
    vl = data.RecordLoader(cfg, data.make_reader(cfg, str(root), "val"), train=False, device="cpu")
    assert vl.crop_size == 16 and len(vl) == 2 and vl.batch_size == 2
    batches = list(vl)
    assert len(batches) == 2
    x, t = batches[0]
    assert tuple(x.shape) == (2, 16, 16, 3) and t.tolist() == [[1.0, 0.0], [1.0, 0.0]]
    want = augment_ref.val_transform_image(imgs[("val", 0, 1)], 16, 16)
    assert np.allclose(x[1].numpy(), want)
    # training: sharded by rank, shuffled per epoch, crop RNG keyed by the running sample index
    seen = []
    for rank in range(2):
        tl = data.RecordLoader(cfg, data.make_reader(cfg, str(root), "train", rank, 2), train=True, device="cpu")
        assert len(tl) == 1                                   # 3 samples per shard, batch 2, drop last
        for x, t in tl:
            assert tuple(x.shape) == (2, 16, 16, 3) and t.sum().item() == 2.0
            seen.append(t.argmax(1).tolist())
        assert tl._seen == 2
    assert seen[0] == [0, 0] or seen[0] == [0, 0][::-1]      # shard 0 holds class n01 only
    assert set(seen[1]) == {1}
    # the stage manager picks the real-data loaders when the data set is there
    run = SimpleNamespace(stages=[SimpleNamespace(start=0, end=1, extra_args=None)])
    full = SimpleNamespace(loader=cfg, val_loader=SimpleNamespace(**dict(vars(cfg), batch_size=2, full_crop=True)), run=run)
    dm = data.DataManager(full, device="cpu")
    dm.set_stage(0)
    assert isinstance(dm.loader, data.RecordLoader) and isinstance(dm.val_loader, data.RecordLoader)
    assert dm.val_loader.crop_size == 16 and len(dm.loader) == 3


def test_loader_background_iterator():
    """`RecordLoader._in_thread` (the loader thread of `prefetch > 0`): items arrive in order, an exception in the
    producer surfaces in the consumer, a consumer that stops early releases the thread."""
    import threading
    import time
    from sota_imagenet_b200 import data
    assert list(data.RecordLoader._in_thread(lambda: iter(range(50)), 2, "sib-test-iter")) == list(range(50))

    def failing():
        yield 1
        raise KeyError("boom")

    it = data.RecordLoader._in_thread(failing, 2, "sib-test-iter")
    assert next(it) == 1
    with pytest.raises(KeyError):
        next(it)

    produced = []

    def endless():
        i = 0
        while True:
            produced.append(i)
            yield i
            i += 1

    it = data.RecordLoader._in_thread(endless, 3, "sib-test-iter")
    assert [next(it) for _ in range(4)] == [0, 1, 2, 3]
    it.close()                                   # the consumer stops: the producer must not stay blocked
    time.sleep(0.3)
    assert not [t for t in threading.enumerate() if t.name == "sib-test-iter" and t.is_alive()]
    assert len(produced) <= 4 + 3 + 2            # never ran more than the queue depth ahead
