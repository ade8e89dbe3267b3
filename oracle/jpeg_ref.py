"""TEST INFRASTRUCTURE ONLY (imported from tests/ and nowhere else).

numpy restatement of the JPEG decoder stages that run on the device in `csrc/jpeg.cu`, i.e. everything
after entropy decoding in the reference's `fn.decoders.image*(device="mixed")` (dali_dataloader.py:65-72,
140-145).  DALI / nvJPEG are absent (closed GPU libraries); the arithmetic restated here is the published
integer pipeline of the IJG / libjpeg-turbo decoder with its default settings --
  jidctint.c  jpeg_idct_islow            (CONST_BITS 13, PASS1_BITS 2)
  jdsample.c  h2v1 / h2v2 fancy upsample (triangle filter, edges replicated)
  jdcolor.c   ycc_rgb_convert            (16-bit fixed-point tables)
-- and it is PINNED against that decoder itself: tests/test_jpeg.py requires `decode_rgb` (fed with the
coefficients of the product's host Huffman stage) to equal PIL's output bit for bit on generated streams
(4:4:4 / 4:2:2 / 4:2:0 / grey, odd extents, restart markers, optimised tables) and on a committed
fixture.  A second, independent entropy decoder (`huffman_decode`, pure Python, small images only)
checks the product's C++ Huffman stage coefficient by coefficient.
"""
import numpy as np

NATURAL = np.array([
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
    28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61,
    54, 47, 55, 62, 63], dtype=np.int64)

F_0_298631336, F_0_390180644, F_0_541196100, F_0_765366865 = 2446, 3196, 4433, 6270
F_0_899976223, F_1_175875602, F_1_501321110, F_1_847759065 = 7373, 9633, 12299, 15137
F_1_961570560, F_2_053119869, F_2_562915447, F_3_072711026 = 16069, 16819, 20995, 25172


def _idct_1d(d, shift):
    """d: [..., 8] int64 -> [..., 8]; one pass of jpeg_idct_islow, descaled by `shift` with rounding."""
    d0, d1, d2, d3, d4, d5, d6, d7 = [d[..., i] for i in range(8)]
    z1 = (d2 + d6) * F_0_541196100
    tmp2 = z1 + d6 * (-F_1_847759065)
    tmp3 = z1 + d2 * F_0_765366865
    tmp0 = (d0 + d4) << 13
    tmp1 = (d0 - d4) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    t0, t1, t2, t3 = d7, d5, d3, d1
    z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
    z5 = (z3 + z4) * F_1_175875602
    t0, t1, t2, t3 = t0 * F_0_298631336, t1 * F_2_053119869, t2 * F_3_072711026, t3 * F_1_501321110
    z1, z2, z3, z4 = z1 * -F_0_899976223, z2 * -F_2_562915447, z3 * -F_1_961570560, z4 * -F_0_390180644
    z3, z4 = z3 + z5, z4 + z5
    t0, t1, t2, t3 = t0 + z1 + z3, t1 + z2 + z4, t2 + z2 + z3, t3 + z1 + z4
    rnd = 1 << (shift - 1)
    out = [tmp10 + t3, tmp11 + t2, tmp12 + t1, tmp13 + t0, tmp13 - t0, tmp12 - t1, tmp11 - t2, tmp10 - t3]
    return np.stack([(o + rnd) >> shift for o in out], axis=-1)


def idct_blocks(coef, quant):
    """coef: [nblocks, 64] int16 (natural order), quant: [64] -> uint8 [nblocks, 8, 8]."""
    x = coef.astype(np.int64).reshape(-1, 8, 8) * quant.astype(np.int64).reshape(1, 8, 8)
    ws = _idct_1d(x.transpose(0, 2, 1), 13 - 2).transpose(0, 2, 1)      # pass 1: columns
    out = _idct_1d(ws, 13 + 2 + 3) + 128                                # pass 2: rows
    return np.clip(out, 0, 255).astype(np.uint8)


def plane_from_blocks(px, blocks_w, blocks_h):
    return px.reshape(blocks_h, blocks_w, 8, 8).transpose(0, 2, 1, 3).reshape(blocks_h * 8, blocks_w * 8)


def upsample_h2v1(pl):
    """[h, w] (real extent) -> [h, 2w]"""
    p = pl.astype(np.int64)
    left = np.concatenate([p[:, :1], p[:, :-1]], axis=1)
    right = np.concatenate([p[:, 1:], p[:, -1:]], axis=1)
    even = (3 * p + left + 1) >> 2
    odd = (3 * p + right + 2) >> 2
    even[:, 0] = p[:, 0]
    odd[:, -1] = p[:, -1]
    out = np.empty((p.shape[0], 2 * p.shape[1]), np.int64)
    out[:, 0::2], out[:, 1::2] = even, odd
    return out.astype(np.uint8)


def upsample_h2v2(pl):
    """[h, w] (real extent) -> [2h, 2w]"""
    p = pl.astype(np.int64)
    above = np.concatenate([p[:1], p[:-1]], axis=0)
    below = np.concatenate([p[1:], p[-1:]], axis=0)
    rows = np.empty((2 * p.shape[0], p.shape[1]), np.int64)
    rows[0::2], rows[1::2] = 3 * p + above, 3 * p + below             # column sums per output row
    last = np.concatenate([rows[:, :1], rows[:, :-1]], axis=1)
    nxt = np.concatenate([rows[:, 1:], rows[:, -1:]], axis=1)
    even = (3 * rows + last + 8) >> 4
    odd = (3 * rows + nxt + 7) >> 4
    even[:, 0] = (rows[:, 0] * 4 + 8) >> 4
    odd[:, -1] = (rows[:, -1] * 4 + 7) >> 4
    out = np.empty((rows.shape[0], 2 * rows.shape[1]), np.int64)
    out[:, 0::2], out[:, 1::2] = even, odd
    return out.astype(np.uint8)


def ycc_to_rgb(y, cb, cr):
    y, cb, cr = y.astype(np.int64), cb.astype(np.int64) - 128, cr.astype(np.int64) - 128
    r = y + ((91881 * cr + 32768) >> 16)
    g = y + ((-22554 * cb + 32768 - 46802 * cr) >> 16)
    b = y + ((116130 * cb + 32768) >> 16)
    return np.clip(np.stack([r, g, b], axis=-1), 0, 255).astype(np.uint8)


def decode_rgb(info, coef):
    """info: dict(width, height, ncomp, hmax, vmax, blocks_w, blocks_h, quant [ncomp][64]); coef: flat int16
    (component after component, blocks in raster order) -> uint8 [H, W, 3]."""
    w, h, planes, off = info["width"], info["height"], [], 0
    for c in range(info["ncomp"]):
        bw, bh = info["blocks_w"][c], info["blocks_h"][c]
        n = bw * bh * 64
        px = idct_blocks(np.asarray(coef[off:off + n]).reshape(-1, 64), np.asarray(info["quant"][c]))
        planes.append(plane_from_blocks(px, bw, bh))
        off += n
    if info["ncomp"] == 1:
        return np.repeat(planes[0][:h, :w, None], 3, axis=2)
    hmax, vmax = info["hmax"], info["vmax"]
    dw, dh = -(-w // hmax), -(-h // vmax)
    chroma = []
    for pl in planes[1:]:
        real = pl[:dh, :dw]
        if hmax == 2 and vmax == 2:
            real = upsample_h2v2(real)
        elif hmax == 2:
            real = upsample_h2v1(real)
        chroma.append(real[:h, :w])
    return ycc_to_rgb(planes[0][:h, :w], chroma[0], chroma[1])


# ------------------------------------------------------------------ independent entropy decoder (small images)
def huffman_decode(data):
    """Pure-Python baseline decoder of a single interleaved scan (ITU T.81 F.2.2) -> (info dict, coef int16).
    Independent of csrc/jpeg.cu: bit-at-a-time canonical decoding, no look-ahead tables."""
    d, p = bytes(data), 2
    qt, huff, comps, dri = {}, {}, [], 0
    assert d[:2] == b"\xff\xd8"
    while True:
        assert d[p] == 0xFF
        m = d[p + 1]
        p += 2
        if m == 0xFF:
            p -= 1
            continue
        ln = (d[p] << 8) | d[p + 1]
        s = d[p + 2:p + ln]
        if m in (0xC0, 0xC1):
            height, width, nc = (s[1] << 8) | s[2], (s[3] << 8) | s[4], s[5]
            comps = [dict(id=s[6 + 3 * i], h=s[7 + 3 * i] >> 4, v=s[7 + 3 * i] & 15, tq=s[8 + 3 * i]) for i in range(nc)]
        elif m == 0xDB:
            q = 0
            while q < len(s):
                pq, tq = s[q] >> 4, s[q] & 15
                vals = [(s[q + 1 + 2 * i] << 8) | s[q + 2 + 2 * i] if pq else s[q + 1 + i] for i in range(64)]
                t = np.zeros(64, np.int64)
                t[NATURAL] = vals
                qt[tq] = t
                q += 1 + 64 * (pq + 1)
        elif m == 0xC4:
            q = 0
            while q < len(s):
                tc, th = s[q] >> 4, s[q] & 15
                bits = list(s[q + 1:q + 17])
                n = sum(bits)
                vals = list(s[q + 17:q + 17 + n])
                codes, code, k = {}, 0, 0
                for ln_ in range(1, 17):
                    for _ in range(bits[ln_ - 1]):
                        codes[(ln_, code)] = vals[k]
                        code += 1
                        k += 1
                    code <<= 1
                huff[(tc, th)] = codes
                q += 17 + n
        elif m == 0xDD:
            dri = (s[0] << 8) | s[1]
        elif m == 0xDA:
            ns = s[0]
            assert ns == len(comps)
            for i in range(ns):
                comps[i]["td"], comps[i]["ta"] = s[2 + 2 * i] >> 4, s[2 + 2 * i] & 15
            p += ln
            break
        p += ln
    if len(comps) == 1:
        comps[0]["h"] = comps[0]["v"] = 1
    hmax, vmax = max(c["h"] for c in comps), max(c["v"] for c in comps)
    mx, my = -(-width // (8 * hmax)), -(-height // (8 * vmax))
    # unstuffed bit stream, split at restart markers
    segs, cur = [], []
    while p < len(d):
        b = d[p]
        if b == 0xFF:
            b2 = d[p + 1]
            if b2 == 0:
                cur.append(0xFF)
                p += 2
                continue
            if 0xD0 <= b2 <= 0xD7:
                segs.append(cur)
                cur = []
                p += 2
                continue
            break
        cur.append(b)
        p += 1
    segs.append(cur)
    blocks = [np.zeros((my * c["v"], mx * c["h"], 64), np.int16) for c in comps]

    class Bits:
        def __init__(self, by):
            self.by, self.i = by, 0

        def bit(self):
            byte = self.by[self.i >> 3] if (self.i >> 3) < len(self.by) else 0
            v = (byte >> (7 - (self.i & 7))) & 1
            self.i += 1
            return v

        def get(self, n):
            v = 0
            for _ in range(n):
                v = (v << 1) | self.bit()
            return v

    def sym(br, table):
        code = 0
        for ln_ in range(1, 17):
            code = (code << 1) | br.bit()
            if (ln_, code) in table:
                return table[(ln_, code)]
        raise ValueError("bad Huffman code")

    def ext(v, s):
        return v - (1 << s) + 1 if v < (1 << (s - 1)) else v

    seg, br, pred, count = 0, Bits(segs[0]), [0] * len(comps), 0
    for yy in range(my):
        for xx in range(mx):
            if dri and count == dri:
                seg += 1
                br, pred, count = Bits(segs[seg]), [0] * len(comps), 0
            for ci, c in enumerate(comps):
                for v in range(c["v"]):
                    for h in range(c["h"]):
                        blk = blocks[ci][yy * c["v"] + v, xx * c["h"] + h]
                        s = sym(br, huff[(0, c["td"])])
                        if s:
                            pred[ci] += ext(br.get(s), s)
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = sym(br, huff[(1, c["ta"])])
                            r, s = rs >> 4, rs & 15
                            if s:
                                k += r
                                blk[NATURAL[k]] = ext(br.get(s), s)
                            elif r != 15:
                                break
                            else:
                                k += 15
                            k += 1
            count += 1
    info = dict(width=width, height=height, ncomp=len(comps), hmax=hmax, vmax=vmax,
                blocks_w=[mx * c["h"] for c in comps], blocks_h=[my * c["v"] for c in comps],
                quant=[qt[c["tq"]] for c in comps])
    return info, np.concatenate([b.reshape(-1) for b in blocks])
