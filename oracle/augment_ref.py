"""numpy restatement of the train-pipeline augmentation (reference dali_dataloader.py:65-74
image_random_crop with random_aspect_ratio=[0.75,1.25], random_area=[min_area,1], 100 attempts;
:74 resize to SxS INTERP_TRIANGULAR; :113-122 crop_mirror_normalize mean 127.5 std 51, coin
flip; :123 one_hot).  DALI itself is absent (closed GPU library) so its RNG stream cannot be
reproduced; the crop RNG is specified here as Philox4x32-10 keyed by (seed, sample index) and
the CUDA kernel must match it bit-exactly.  Parity unpinned by the reference (no DALI vectors).
"""
import math

import numpy as np

M32 = 0xFFFFFFFF
LOG_R_LO = -0.2876820724517809   # ln 0.75
LOG_R_HI = 0.22314355131420976   # ln 1.25


def philox4x32_10(counter, key):
    c0, c1, c2, c3 = [int(c) & M32 for c in counter]
    k0, k1 = [int(k) & M32 for k in key]
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        n0 = ((p1 >> 32) ^ c1 ^ k0) & M32
        n1 = p1 & M32
        n2 = ((p0 >> 32) ^ c3 ^ k1) & M32
        n3 = p0 & M32
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


def rrc_box(H, W, min_area, max_area, seed, sample):
    """-> [x0, y0, w, h, flip] (dali_dataloader.py:65-72 crop, :113-116 mirror coin)."""
    key = (seed & M32, (seed >> 32) & M32)
    s0, s1 = sample & M32, (sample >> 32) & M32
    box = None
    for attempt in range(100):
        r = philox4x32_10((s0, s1, attempt, 0), key)
        ua = r[0] * (1.0 / 4294967296.0)
        ur = r[1] * (1.0 / 4294967296.0)
        area = (min_area + (max_area - min_area) * ua) * float(H) * float(W)
        ratio = math.exp(LOG_R_LO + (LOG_R_HI - LOG_R_LO) * ur)
        w = int(math.floor(math.sqrt(area * ratio) + 0.5))
        h = int(math.floor(math.sqrt(area / ratio) + 0.5))
        if 0 < w <= W and 0 < h <= H:
            box = [r[2] % (W - w + 1), r[3] % (H - h + 1), w, h]
            break
    if box is None:
        in_ratio = W / H
        if in_ratio < 0.75:
            w, h = W, int(math.floor(W / 0.75 + 0.5))
        elif in_ratio > 1.25:
            h, w = H, int(math.floor(H * 1.25 + 0.5))
        else:
            w, h = W, H
        h, w = min(h, H), min(w, W)
        box = [(W - w) // 2, (H - h) // 2, w, h]
    r = philox4x32_10((s0, s1, 100, 1), key)
    return box + [r[0] & 1]


def _tri_weights(out_size, crop, flip):
    """Per output index: list of (source index inside crop, weight) for the triangular filter
    whose support scales with the down-scale factor (anti-aliased bilinear)."""
    f32 = np.float32
    scale = f32(crop) / f32(out_size)
    sup = max(scale, f32(1.0))
    taps = []
    for o in range(out_size):
        so = (out_size - 1 - o) if flip else o
        c = (f32(so) + f32(0.5)) * scale
        lo = int(math.floor(c - sup))
        hi = int(math.ceil(c + sup))
        row = []
        for i in range(lo, hi):
            w = max(f32(0.0), f32(1.0) - abs((f32(i) + f32(0.5) - c) / sup))
            if w > 0:
                row.append((min(max(i, 0), crop - 1), f32(w)))
        taps.append(row)
    return taps


def augment_image(img, box, size, mean=127.5, std=51.0):
    """img uint8 [H,W,3] -> float32 [size,size,3] normalised (before bf16 rounding)."""
    x0, y0, cw, ch, flip = box
    tx = _tri_weights(size, cw, flip)
    ty = _tri_weights(size, ch, False)
    crop = img[y0:y0 + ch, x0:x0 + cw].astype(np.float32)
    out = np.zeros((size, size, 3), np.float32)
    for oy in range(size):
        for ox in range(size):
            acc = np.zeros(3, np.float64)
            ws = 0.0
            for (sy, wy) in ty[oy]:
                for (sx, wx) in tx[ox]:
                    w = float(wx) * float(wy)
                    acc += w * crop[sy, sx]
                    ws += w
            out[oy, ox] = (acc / ws - mean) / std
    return out


def one_hot(labels, num_classes):
    out = np.zeros((len(labels), num_classes), np.float32)
    out[np.arange(len(labels)), labels] = 1.0
    return out


# ---------------------------------------------------------------------------------------------
# Batch-level mixing.  pt_clb.Mixup / pt_clb.Cutmix live in the absent
# pytorch_tools package (unpinned); the reference combines them in sota_imagenet/callbacks.py:
# 232-247.  Restated in NCHW, fp32 arithmetic with one rounding at the end:
#   mixup : out = c * x + (1 - c) * prev[perm]      targets likewise
#   cutmix: out[:, :, h1:h2, w1:w2] = prev[perm][:, :, h1:h2, w1:w2];  t = (1-lam) t + lam prev_t[perm]
def mixup_batch(x, prev, perm, c):
    c = np.float32(c)
    omc = np.float32(1.0) - c
    return (c * x.astype(np.float32) + omc * prev[perm].astype(np.float32)).astype(np.float32)


def cutmix_batch(x, prev, perm, box):
    h1, w1, h2, w2 = box
    out = x.copy()
    out[:, :, h1:h2, w1:w2] = prev[perm][:, :, h1:h2, w1:w2]
    return out


def mix_targets(t, prev_t, perm, w_self, w_prev):
    return (np.float32(w_self) * t.astype(np.float32) + np.float32(w_prev) * prev_t[perm].astype(np.float32))


def cutmix_bbox(H, W, lam, ch, cw):
    """box of area ~ lam*H*W centred on (ch, cw), clipped to the image; returns the box and the
    clipped box's true share of the image (the lambda the targets are mixed with)."""
    cut_rat = np.sqrt(lam)
    cut_h, cut_w = int(H * cut_rat), int(W * cut_rat)
    box = (int(np.clip(ch - cut_h // 2, 0, H)), int(np.clip(cw - cut_w // 2, 0, W)),
           int(np.clip(ch + cut_h // 2, 0, H)), int(np.clip(cw + cut_w // 2, 0, W)))
    return box, (box[2] - box[0]) * (box[3] - box[1]) / (H * W)


# ---------------------------------------------------------------------------------------------
# Validation transform (reference dali_dataloader.py:146-160): fn.resize(resize_shorter=crop_size,
# INTERP_TRIANGULAR) then crop_mirror_normalize(crop=(S, S)) (centre crop, no mirror).  DALI is
# absent: restated, unpinned.  Rounding conventions fixed here: the longer side is
# floor(long * RS / short + 0.5), the crop origin floor(0.5 * (R - S) + 0.5).
def val_geometry(sh, sw, size, resize_shorter):
    if sh <= sw:
        rh, rw = resize_shorter, int(math.floor(sw * resize_shorter / sh + 0.5))
    else:
        rw, rh = resize_shorter, int(math.floor(sh * resize_shorter / sw + 0.5))
    return [rh, rw, int(math.floor(0.5 * (rh - size) + 0.5)), int(math.floor(0.5 * (rw - size) + 0.5))]


def _tri_weights_window(out_size, origin, in_size, resized):
    """taps of output index o = pixel (o + origin) of the virtual `resized`-long axis, over an
    `in_size`-long source axis; taps clamp at the image border."""
    f32 = np.float32
    scale = f32(in_size) / f32(resized)
    sup = max(scale, f32(1.0))
    taps = []
    for o in range(out_size):
        c = (f32(o + origin) + f32(0.5)) * scale
        lo, hi = int(math.floor(c - sup)), int(math.ceil(c + sup))
        row = []
        for i in range(lo, hi):
            w = max(f32(0.0), f32(1.0) - abs((f32(i) + f32(0.5) - c) / sup))
            if w > 0:
                row.append((min(max(i, 0), in_size - 1), f32(w)))
        taps.append(row)
    return taps


def val_transform_image(img, size, resize_shorter, mean=127.5, std=51.0):
    """img uint8 [H,W,3] -> float32 [size,size,3]"""
    sh, sw = img.shape[:2]
    rh, rw, oy0, ox0 = val_geometry(sh, sw, size, resize_shorter)
    ty = _tri_weights_window(size, oy0, sh, rh)
    tx = _tri_weights_window(size, ox0, sw, rw)
    src = img.astype(np.float32)
    out = np.zeros((size, size, 3), np.float32)
    for oy in range(size):
        for ox in range(size):
            acc = np.zeros(3, np.float64)
            ws = 0.0
            for (sy, wy) in ty[oy]:
                for (sx, wx) in tx[ox]:
                    w = float(wx) * float(wy)
                    acc += w * src[sy, sx]
                    ws += w
            out[oy, ox] = (acc / ws - mean) / std
    return out


# ---------------------------------------------------------------------------------------------
# Photometric augmentations of the resident batch (reference dali_dataloader.py:81-111:
# fn.gaussian_blur / fn.color_twist / fn.hsv(saturation=0) / fn.erase).  DALI itself is absent
# (parity unpinned); these restate the operators as the affine maps their documentation gives,
# evaluated on the NORMALISED values like the CUDA kernels (csrc/batchaug.cu).
# ---------------------------------------------------------------------------------------------
def pixel_ops(x, params, flips, nboxes):
    """x: float array [N, H, W, 3] (normalised values); params: [N, 16 + 4 * nboxes] as produced
    by data.BatchPixelAug.draw; flips: [N] bool (crop was mirrored).  Returns the transformed
    array (float32 arithmetic in the kernel's operation order)."""
    x = np.asarray(x, dtype=np.float32)
    out = np.empty_like(x)
    n, h, w, _ = x.shape
    for i in range(n):
        p = params[i].astype(np.float32)
        m, o = p[0:9].reshape(3, 3), p[9:12]
        v = x[i]
        y = np.empty_like(v)
        for c in range(3):
            t = np.float32(m[c, 0]) * v[..., 0] + o[c]
            t = np.float32(m[c, 1]) * v[..., 1] + t
            t = np.float32(m[c, 2]) * v[..., 2] + t
            y[..., c] = np.minimum(np.maximum(t, p[12]), p[13])
        if p[14] != 0:
            g = np.float32(0.299) * y[..., 0]
            g = np.float32(0.587) * y[..., 1] + g
            g = np.float32(0.114) * y[..., 2] + g
            y[...] = g[..., None]
        for b in range(nboxes):
            h1, w1, h2, w2 = (int(q) for q in p[16 + 4 * b:20 + 4 * b])
            if flips[i]:
                w1, w2 = w - w2, w - w1
            y[h1:h2, w1:w2, :] = p[15]
        out[i] = y
    return out


def _reflect101(i, n):
    if n == 1:
        return 0
    while i < 0 or i >= n:
        i = -i if i < 0 else 2 * (n - 1) - i
    return i


def gaussian_blur11(x, sigma):
    """Separable 11-tap Gaussian with reflect-101 borders per sample; sigma <= 0 copies through.
    x: [N, H, W, 3] float."""
    x = np.asarray(x, dtype=np.float64)
    out = x.copy()
    n, h, w, _ = x.shape
    for i in range(n):
        if sigma[i] <= 0:
            continue
        k = np.arange(-5, 6)
        wt = np.exp(-(k.astype(np.float64) ** 2) / (2.0 * float(sigma[i]) ** 2))
        wt /= wt.sum()
        rows = np.array([[_reflect101(r + d, h) for d in k] for r in range(h)])
        cols = np.array([[_reflect101(c + d, w) for d in k] for c in range(w)])
        hor = np.einsum("hwkc,k->hwc", x[i][:, cols, :], wt)
        out[i] = np.einsum("hkwc,k->hwc", hor[rows, :, :], wt)
    return out
