"""Functional wrappers: torch tensors in, C-ABI kernel launches on the current stream.

Activations are torch tensors of logical shape [N, C, H, W], dtype bfloat16, in
torch.channels_last memory format (= NHWC in memory).  Conv filters are [K, C, R, S]
channels_last (= KRSC in memory).  PyTorch is only the allocator / stream provider here.
"""
import os

import torch

from . import _lib
from ._lib import c_void_p


def call(name, *args):
    _lib.call(name, *args)

# SIB_DETERMINISTIC=1: bitwise run-to-run reproducibility (goldens, debugging) at the price of speed.
# No floating-point atomics anywhere on the step: BatchNorm statistics come from a stand-alone
# fixed-order reduction instead of the conv epilogues, the BN-backward sums from bn_bwd_reduce
# (fixed order) instead of the dgrad epilogues, the weight gradient is not split over pixels, and
# the conv-prologue fusion (which consumes epilogue statistics) is off.  Read at import time; the
# C side reads the same variable (csrc/norm.cu det_scratch, csrc/conv.cu sib_conv2d_wgrad).
DETERMINISTIC = os.environ.get("SIB_DETERMINISTIC", "0") == "1"

ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2
MARGIN_NONE, MARGIN_ARC, MARGIN_COS, MARGIN_ARC_PURE, MARGIN_ARCCOS = 0, 1, 2, 3, 4
FLAG_FORCE_IM2COL = 1
FLAG_TILE_N128 = 2
FLAG_NO_2CTA = 4
FLAG_STATS_ZEROED = 8
FLAG_FORCE_HALO = 16
FLAG_NO_HALO = 32
ACT_FLAG_PREZEROED = 0x100
ACT_CODES = {None: ACT_NONE, "identity": ACT_NONE, "none": ACT_NONE, "relu": ACT_RELU,
             "leaky_relu": ACT_LEAKY}


def _p(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_act(t, name="activation"):
    if t.dtype != torch.bfloat16 or not t.is_cuda or t.dim() != 4:
        raise _lib.SibError("%s must be a 4-D CUDA bfloat16 tensor, got %s %s" % (name, t.dtype, tuple(t.shape)))
    if not t.permute(0, 2, 3, 1).is_contiguous():
        raise _lib.SibError("%s must be channels_last (NHWC) contiguous" % name)


def new_act(n, c, h, w, device, dtype=torch.bfloat16):
    return torch.empty((n, h, w, c), dtype=dtype, device=device).permute(0, 3, 1, 2)


def to_nhwc_bf16(x):
    """Any float NCHW tensor -> channels_last bf16 (a torch copy; test / glue use only)."""
    n, c, h, w = x.shape
    out = new_act(n, c, h, w, x.device)
    out.copy_(x)
    return out


class _AccPool:
    """fp32 accumulators that kernels add into with atomics (BN statistics from the conv epilogues,
    backward sums) are carved out of ONE buffer that a single fill zeroes at the start of each
    forward / backward pass, instead of one cudaMemsetAsync node per accumulator (~100 per step)."""
    SIZE = 1 << 19   # floats (2 MB)

    def __init__(self):
        self.buf, self.off = None, 0

    def reset(self, device):
        if self.buf is None or self.buf.device != device:
            self.buf = torch.zeros(self.SIZE, dtype=torch.float32, device=device)
        else:
            self.buf.zero_()
        self.off = 0

    def take(self, rows, c, device):
        n = rows * c
        if self.buf is None or self.buf.device != device or self.off + n > self.SIZE:
            return None
        t = self.buf[self.off:self.off + n].view(rows, c)
        self.off += (n + 3) // 4 * 4     # keep 16-byte alignment
        return t

    def owns(self, t):
        if t is None or self.buf is None or t.device != self.buf.device:
            return False
        base = self.buf.data_ptr()
        return base <= t.data_ptr() < base + 4 * self.SIZE


_ACC_POOL = _AccPool()
_DEBUG_POOL = os.environ.get("SIB_DEBUG_POOL", "0") == "1"
_DIRTY = []


def _report_dirty():
    bad = [(int(n), off, rows, c, tb) for n, off, rows, c, tb in _DIRTY if int(n) != 0]
    print("ACC POOL: %d slices checked, %d dirty" % (len(_DIRTY), len(bad)))
    for n, off, rows, c, tb in bad[:5]:
        print("dirty slice: %d non-zeros, end offset %d, shape [%d,%d]\n%s" % (n, off, rows, c, tb))


if _DEBUG_POOL:
    import atexit
    atexit.register(_report_dirty)
_USE_POOL = os.environ.get("SIB_ACC_POOL", "1") != "0"


PEER = None      # parallel.PeerAllReduce once a SyncBN data-parallel wrapper has set it up


def begin_pass(device, forward=False):
    """Called by the root module at the start of a forward or backward pass."""
    _ACC_POOL.reset(device)
    if forward and PEER is not None:
        PEER.begin_forward()


def small_allreduce_(t, group=None):
    """Sum a small fp32 statistics tensor over the ranks: peer-memory one-shot kernel when
    available (same group), else torch.distributed."""
    if PEER is not None and PEER.group is group and t.is_cuda:
        return PEER.allreduce_(t)
    torch.distributed.all_reduce(t, group=group)
    return t


def new_acc(rows, c, device):
    """[rows, c] fp32 accumulator: a pre-zeroed pool slice when available (the kernels then skip
    their own memset), else an uninitialised tensor the kernel zeroes itself."""
    t = _ACC_POOL.take(rows, c, device) if _USE_POOL else None
    if t is not None and _DEBUG_POOL and not torch.cuda.is_current_stream_capturing():
        # asynchronous check (keeps the GPU timeline dense): count non-zeros now, report at exit
        import traceback
        _DIRTY.append(((t != 0).sum(), _ACC_POOL.off, rows, c, "".join(traceback.format_stack(limit=4)[:-1])))
    return t if t is not None else torch.empty((rows, c), dtype=torch.float32, device=device)


class _SideStream:
    """Weight-gradient kernels have no consumer until the optimizer (or the gradient all-reduce of
    their block), so they run on a second stream: the tensor-core-bound wgrad then overlaps the
    HBM-bound BatchNorm-backward kernels and the head / tail of the dgrad that follow on the main
    stream.  Inputs are kept alive until the join (no allocator stream bookkeeping, graph-capture
    safe: the fork / join become parallel branches of the captured graph)."""
    enabled = os.environ.get("SIB_WGRAD_STREAM", "1") != "0"
    streams = {}
    keep = []
    forked = False


def side_launch(fn, *tensors):
    """Run fn() (kernel launches reading `tensors`) on the side stream, ordered after everything
    enqueued so far on the current stream."""
    if not _SideStream.enabled:
        fn()
        return
    main = torch.cuda.current_stream()
    side = _SideStream.streams.get(main.device)
    if side is None:
        side = _SideStream.streams[main.device] = torch.cuda.Stream(device=main.device)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        fn()
    _SideStream.keep.extend(tensors)
    _SideStream.forked = True


def side_join():
    """Make the current stream wait for the side stream; drops the keep-alive references."""
    if _SideStream.forked:
        main = torch.cuda.current_stream()
        side = _SideStream.streams.get(main.device)
        if side is not None:
            main.wait_stream(side)
        _SideStream.forked = False
    _SideStream.keep.clear()


def conv_out_hw(h, w, r, s, stride, pad):
    return (h + 2 * pad - r) // stride + 1, (w + 2 * pad - s) // stride + 1


# ------------------------------------------------------------------ convolution
def conv2d_fprop(x, w, stride=1, pad=0, stats=None, bias=None, flags=0, pad_hw=None, out_hw=None):
    _lib.require_device()
    _check_act(x, "x")
    n, c, h, wd = x.shape
    k, c2, r, s = w.shape
    assert c2 == c, "filter/input channel mismatch"
    ph, pw = pad_hw if pad_hw is not None else (pad, pad)
    oh, ow = out_hw if out_hw is not None else conv_out_hw(h, wd, r, s, stride, pad)
    y = new_act(n, k, oh, ow, x.device)
    if _ACC_POOL.owns(stats):
        flags |= FLAG_STATS_ZEROED
    call("sib_conv2d_fprop", _p(x), _p(w), _p(y), n, h, wd, c, k, r, s, stride, ph, pw, oh, ow,
         _p(bias), _p(stats), flags, _stream())
    return y


def conv2d_fprop_bnact(x, w, bn, stride=1, pad=0, stats=None, act=ACT_RELU, slope=0.01, count=None,
                       eps=1e-5, momentum=0.1, scale_shift=None, flags=0):
    """y = conv(act(BN(x)), w) with the BatchNorm (+ activation) of the producer layer applied to
    the conv's A operand inside the kernel: `x` is the RAW output of the previous conv, the
    normalised activation never reaches HBM (csrc/conv.cu BnPrologue).
    Training: bn = (stats, gamma, beta, running_mean, running_var); returns y, mean_invstd,
    scale_shift (the finalize step is folded into the conv).  Eval: bn = None, `scale_shift` given."""
    _lib.require_device()
    _check_act(x, "x")
    n, c, h, wd = x.shape
    k, c2, r, s = w.shape
    assert c2 == c, "filter/input channel mismatch"
    oh, ow = conv_out_hw(h, wd, r, s, stride, pad)
    y = new_act(n, k, oh, ow, x.device)
    if _ACC_POOL.owns(stats):
        flags |= FLAG_STATS_ZEROED
    mi = None
    if bn is not None:
        mi = torch.empty((2, c), dtype=torch.float32, device=x.device)
        scale_shift = torch.empty((2, c), dtype=torch.float32, device=x.device)
    b = bn if bn is not None else (None,) * 5
    call("sib_conv2d_fprop_bnact", _p(x), _p(w), _p(y), n, h, wd, c, k, r, s, stride, pad, pad, oh, ow,
         _p(stats), flags, _p(b[0]), _p(b[1]), _p(b[2]), _p(b[3]), _p(b[4]), _p(mi), _p(scale_shift),
         float(count if count is not None else n * h * wd), float(eps), float(momentum), act,
         float(slope), _stream())
    return y, mi, scale_shift


def fprop_bnact_ok(c_in):
    """Mirror of the C-side limits of the fused prologue (csrc/conv.cu kProMaxC)."""
    return c_in % 64 == 0 and c_in <= 512


def pack_dgrad_s2(w_dgrad, sub0=None, sub1=None):
    """[C][3][3][K] flipped pack -> row-parity sub-filters ([2C][1][2][K], [2C][2][2][K]) of the
    3x3 / stride-2 dgrad."""
    c, r, s, k = w_dgrad.shape
    assert r == 3 and s == 3
    if sub0 is None:
        sub0 = torch.empty((2 * c, 1, 2, k), dtype=torch.bfloat16, device=w_dgrad.device)
        sub1 = torch.empty((2 * c, 2, 2, k), dtype=torch.bfloat16, device=w_dgrad.device)
    call("sib_pack_dgrad_s2", _p(w_dgrad), _p(sub0), _p(sub1), c, k, _stream())
    return sub0, sub1


def halo_applies(c_out, c_in, r, stride, width):
    """Mirror of the automatic dispatch in csrc/conv.cu run_igemm: the halo-reuse kernel takes the
    3x3 / stride-1 / pad-1 convolutions with 64 -> 64 channels at >= 28 columns."""
    return r == 3 and stride == 1 and c_out == 64 and c_in == 64 and 28 <= width <= 61


def dgrad_s2_ok(x_shape, r, s, stride, pad):
    n, c, h, w = x_shape
    return r == 3 and s == 3 and stride == 2 and pad == 1 and h % 2 == 0 and w % 2 == 0 and w // 2 <= 32 \
        and (2 * c) % 64 == 0


def conv2d_dgrad(dy, w_dgrad, x_shape, r, s, stride=1, pad=0, out=None, residual=None, flags=0,
                 bn_bwd=None, w_s2=None):
    """dx for conv(x, w); `w_dgrad` is the [C][R][S][K] flipped pack of w.
    stride 1 (or strided RxS): dx = dgrad [+ residual].  strided 1x1: dx (= `out`, or zeros) +=
    dgrad at every stride-th pixel.
    bn_bwd = dict(mask_src, mask_ss, xhat_src, mean_invstd, act, slope): fuse the backward
    reduction of the BatchNorm (+activation) this gradient flows into; returns (dx_masked, sums)."""
    _lib.require_device()
    _check_act(dy, "dy")
    n, c, h, wd = x_shape
    k = dy.shape[1]
    oh, ow = dy.shape[2], dy.shape[3]
    if w_s2 is not None and residual is None and k % 64 == 0 and dgrad_s2_ok(x_shape, r, s, stride, pad) \
            and (bn_bwd is None or bn_bwd.get("xhat_src") is None):
        # row-parity decomposition: no zero-inserted copy of dy, 12 instead of 36 taps of work
        if out is None:
            out = new_act(n, c, h, wd, dy.device)
        sums = None
        fl = flags
        if bn_bwd is not None:
            sums = new_acc(2, c, dy.device)
            if _ACC_POOL.owns(sums):
                fl |= FLAG_STATS_ZEROED
        bb = bn_bwd or {}
        call("sib_conv2d_dgrad_s2", _p(dy), _p(w_s2[0]), _p(w_s2[1]), _p(out), n, h, wd, c, k, fl,
             _p(bb.get("mask_src")), _p(bb.get("mask_ss")), _p(bb.get("mean_invstd")),
             bb.get("act", ACT_NONE), float(bb.get("slope", 0.0)), _p(sums), _stream())
        return (out, sums) if bn_bwd is not None else out
    ws = None
    if stride > 1 and r == 1 and s == 1:
        assert residual is None
        if out is None:
            out = torch.zeros((n, h, wd, c), dtype=torch.bfloat16, device=dy.device).permute(0, 3, 1, 2)
        ws = torch.empty((n, oh, ow, c), dtype=torch.bfloat16, device=dy.device)
    elif stride > 1:
        uh, uw = (oh - 1) * stride + 1, (ow - 1) * stride + 1
        ws = torch.empty((n, uh, uw, k), dtype=torch.bfloat16, device=dy.device)
    if out is None:
        out = new_act(n, c, h, wd, dy.device)
    if bn_bwd is not None:
        sums = new_acc(2, c, dy.device)
        if _ACC_POOL.owns(sums):
            flags |= FLAG_STATS_ZEROED
        call("sib_conv2d_dgrad_bnbwd", _p(dy), _p(w_dgrad), _p(out), _p(residual), _p(ws), n, h, wd,
             c, k, r, s, stride, pad, flags, _p(bn_bwd["mask_src"]), _p(bn_bwd.get("mask_ss")),
             _p(bn_bwd.get("xhat_src")), _p(bn_bwd["mean_invstd"]), bn_bwd["act"],
             float(bn_bwd.get("slope", 0.0)), _p(sums), _stream())
        return out, sums
    call("sib_conv2d_dgrad", _p(dy), _p(w_dgrad), _p(out), _p(residual), _p(ws), n, h, wd, c, k, r,
         s, stride, pad, flags, _stream())
    return out


def conv2d_wgrad(x, dy, dw, stride=1, pad=0, flags=0, pad_hw=None):
    """dw[K][R][S][C] (fp32, channels_last view of [K,C,R,S]) += wgrad(x, dy)."""
    _lib.require_device()
    _check_act(x, "x")
    _check_act(dy, "dy")
    n, c, h, wd = x.shape
    k, c2, r, s = dw.shape
    assert c2 == c and dy.shape[1] == k
    assert dw.dtype == torch.float32 and dw.permute(0, 2, 3, 1).is_contiguous()
    ph, pw = pad_hw if pad_hw is not None else (pad, pad)
    call("sib_conv2d_wgrad", _p(x), _p(dy), _p(dw), n, h, wd, c, k, r, s, stride, ph, pw,
         dy.shape[2], dy.shape[3], flags, _stream())
    return dw


def pack_dgrad_weight(w):
    """Single-filter helper (tests / generic modules): KRSC bf16 -> flipped [C][R][S][K]."""
    k, c, r, s = w.shape
    src = w.permute(0, 2, 3, 1).contiguous()
    out = torch.empty((c, r, s, k), dtype=torch.bfloat16, device=w.device)
    table = pack_table([(0, 0, k, r * s, c)], w.device)
    call("sib_pack_dgrad_weights", _p(src), _p(out), _p(table[0]), 1, table[1], _stream())
    return out


def pack_table(entries, device):
    """entries: (src_off, dst_off, K, RS, C) -> (device table, total_blocks)."""
    import numpy as np
    rec = np.zeros(len(entries), dtype=np.dtype([("src", "<i8"), ("dst", "<i8"), ("K", "<i4"),
                                                 ("RS", "<i4"), ("C", "<i4"), ("bb", "<i4")]))
    blocks = 0
    for i, (so, do, k, rs, c) in enumerate(entries):
        rec[i] = (so, do, k, rs, c, blocks)
        blocks += rs * ((k + 31) // 32) * ((c + 31) // 32)
    t = torch.from_numpy(rec.view(np.uint8).copy()).to(device)
    return t, blocks


# ------------------------------------------------------------------ batch norm family
def bn_stats(x):
    _check_act(x)
    n, c, h, w = x.shape
    stats = torch.empty((2, c), dtype=torch.float32, device=x.device)
    call("sib_bn_stats", _p(x), n * h * w, c, _p(stats), _stream())
    return stats


def bn_finalize(stats, gamma, beta, running_mean, running_var, count, eps, momentum):
    c = stats.shape[1]
    mean_invstd = torch.empty((2, c), dtype=torch.float32, device=stats.device)
    scale_shift = torch.empty((2, c), dtype=torch.float32, device=stats.device)
    call("sib_bn_finalize", _p(stats), _p(gamma), _p(beta), _p(running_mean), _p(running_var),
         _p(mean_invstd), _p(scale_shift), c, float(count), float(eps), float(momentum), _stream())
    return mean_invstd, scale_shift


def bn_finalize_apply(x, bn1, res=None, bn2=None, act=ACT_NONE, slope=0.01, count=None, eps=1e-5,
                      momentum=0.1):
    """bn1 / bn2 = (stats, gamma, beta, running_mean, running_var).  Returns y, (mi, ss)[, (mi2, ss2)]."""
    _check_act(x)
    n, c, h, w = x.shape
    y = new_act(n, c, h, w, x.device)
    mk = lambda: (torch.empty((2, c), dtype=torch.float32, device=x.device),
                  torch.empty((2, c), dtype=torch.float32, device=x.device))
    mi, ss = mk()
    mi2, ss2 = mk() if bn2 is not None else (None, None)
    b2 = bn2 if bn2 is not None else (None,) * 5
    call("sib_bn_finalize_apply", _p(x), _p(bn1[0]), _p(bn1[1]), _p(bn1[2]), _p(bn1[3]), _p(bn1[4]),
         _p(mi), _p(ss), _p(res), _p(b2[0]), _p(b2[1]), _p(b2[2]), _p(b2[3]), _p(b2[4]), _p(mi2),
         _p(ss2), _p(y), n * h * w, c, float(count if count is not None else n * h * w), float(eps),
         float(momentum), act, float(slope), _stream())
    return y, (mi, ss), (mi2, ss2)


def bn_eval_scale(gamma, beta, running_mean, running_var, eps):
    c = running_mean.shape[0]
    scale_shift = torch.empty((2, c), dtype=torch.float32, device=running_mean.device)
    call("sib_bn_eval_scale", _p(gamma), _p(beta), _p(running_mean), _p(running_var),
         _p(scale_shift), c, float(eps), _stream())
    return scale_shift


def bn_apply(x, scale_shift, act=ACT_NONE, slope=0.01, res=None, scale_shift2=None, out=None):
    _check_act(x)
    n, c, h, w = x.shape
    y = out if out is not None else new_act(n, c, h, w, x.device)
    call("sib_bn_apply", _p(x), _p(scale_shift), _p(res), _p(scale_shift2), _p(y), n * h * w, c,
         act, float(slope), _stream())
    return y


def bn_bwd_reduce(dy, out, x, mean_invstd, act, slope=0.01, x2=None, mean_invstd2=None,
                  mask_ss=None):
    """`out` (stored forward output) or `mask_ss` (forward scale/shift, mask recomputed from x)."""
    n, c, h, w = x.shape
    sums = new_acc(4 if x2 is not None else 2, c, x.device)
    if _ACC_POOL.owns(sums):
        act |= ACT_FLAG_PREZEROED
    call("sib_bn_bwd_reduce", _p(dy), _p(out), _p(mask_ss), _p(x), _p(mean_invstd), _p(x2),
         _p(mean_invstd2), n * h * w, c, act, float(slope), _p(sums), _stream())
    return sums


def bn_bwd_apply(dy, out, x, mean_invstd, gamma, sums, count, act, slope=0.01, x2=None,
                 mean_invstd2=None, gamma2=None, want_g=False, dx_out=None, mask_ss=None,
                 param_grads=(None, None), param_grads2=(None, None), pgrad_scale=1.0):
    n, c, h, w = x.shape
    if len(param_grads) > 2:          # (dgamma, dbeta, scale) from BatchNorm2d.grad_ptrs()
        pgrad_scale = param_grads[2]
    dx = dx_out if dx_out is not None else new_act(n, c, h, w, x.device)
    dx2 = new_act(n, c, h, w, x.device) if x2 is not None else None
    g = new_act(n, c, h, w, x.device) if want_g else None
    call("sib_bn_bwd_apply", _p(dy), _p(out), _p(mask_ss), _p(x), _p(mean_invstd), _p(gamma),
         _p(sums), _p(x2), _p(mean_invstd2), _p(gamma2), _p(dx), _p(dx2), _p(g),
         _p(param_grads[0]), _p(param_grads[1]), _p(param_grads2[0]), _p(param_grads2[1]),
         n * h * w, c, float(count), act, float(slope), float(pgrad_scale), _stream())
    return dx, dx2, g


def bn_bwd_apply_remat(dy, x, mean_invstd, gamma, sums, count, act_ss, fwd_act, fwd_slope=0.01,
                       act=ACT_NONE, slope=0.01, mask_ss=None, param_grads=(None, None), pgrad_scale=1.0):
    """bn_bwd_apply of a plain BatchNorm (+ activation) that also re-materialises the forward
    activation a = fwd_act(fmaf(x, scale, shift)) (the fused conv prologue never stored it; the
    consumer conv's weight gradient reads it).  Returns dx, a."""
    n, c, h, w = x.shape
    if len(param_grads) > 2:
        pgrad_scale = param_grads[2]
    dx = new_act(n, c, h, w, x.device)
    a = new_act(n, c, h, w, x.device)
    call("sib_bn_bwd_apply_remat", _p(dy), _p(mask_ss), _p(x), _p(mean_invstd), _p(gamma), _p(sums),
         _p(dx), _p(param_grads[0]), _p(param_grads[1]), _p(act_ss), fwd_act, float(fwd_slope), _p(a),
         n * h * w, c, float(count), act, float(slope), float(pgrad_scale), _stream())
    return dx, a


def bn_param_grad(sums, dgamma, dbeta, accumulate=True):
    c = sums.shape[1]
    call("sib_bn_param_grad", _p(sums), _p(dgamma), _p(dbeta), c, int(accumulate), _stream())


# ------------------------------------------------------------------ pooling
def maxpool3x3s2_fwd(x, want_idx=True):
    _check_act(x)
    n, c, h, w = x.shape
    oh, ow = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
    y = new_act(n, c, oh, ow, x.device)
    idx = torch.empty((n, oh, ow, c), dtype=torch.uint8, device=x.device) if want_idx else None
    call("sib_maxpool3x3s2_fwd", _p(x), _p(y), _p(idx), n, h, w, c, _stream())
    return y, idx


def bn_act_maxpool3x3s2_fwd(x, bn, act=ACT_RELU, slope=0.01, count=None, eps=1e-5, momentum=0.1):
    """Stem: y = maxpool3x3s2(act(bn(x))) in one pass.  bn = (stats, gamma, beta, running_mean,
    running_var).  Returns y, idx, mean_invstd, scale_shift."""
    _check_act(x)
    n, c, h, w = x.shape
    oh, ow = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
    y = new_act(n, c, oh, ow, x.device)
    idx = torch.empty((n, oh, ow, c), dtype=torch.uint8, device=x.device)
    mi = torch.empty((2, c), dtype=torch.float32, device=x.device)
    ss = torch.empty((2, c), dtype=torch.float32, device=x.device)
    call("sib_bn_act_maxpool3x3s2_fwd", _p(x), _p(bn[0]), _p(bn[1]), _p(bn[2]), _p(bn[3]), _p(bn[4]),
         _p(mi), _p(ss), _p(y), _p(idx), n, h, w, c, float(count if count is not None else n * h * w),
         float(eps), float(momentum), act, float(slope), _stream())
    return y, idx, mi, ss


def maxpool3x3s2_bwd(dy, idx, x_shape):
    n, c, h, w = x_shape
    dx = new_act(n, c, h, w, dy.device)
    call("sib_maxpool3x3s2_bwd", _p(dy), _p(idx), _p(dx), n, h, w, c, _stream())
    return dx


def gap_fwd(x):
    _check_act(x)
    n, c, h, w = x.shape
    y = new_act(n, c, 1, 1, x.device)
    call("sib_gap_fwd", _p(x), _p(y), n, h * w, c, _stream())
    return y


def gap_bwd(dy, x_shape):
    n, c, h, w = x_shape
    dx = new_act(n, c, h, w, dy.device)
    call("sib_gap_bwd", _p(dy), _p(dx), n, h * w, c, _stream())
    return dx


# ------------------------------------------------------------------ BResNet extras
def blurpool_fwd(x):
    n, c, h, w = x.shape
    y = new_act(n, c, (h - 1) // 2 + 1, (w - 1) // 2 + 1, x.device)
    call("sib_blurpool_fwd", _p(x), _p(y), n, h, w, c, _stream())
    return y


def blurpool_bwd(dy, x_shape):
    n, c, h, w = x_shape
    dx = new_act(n, c, h, w, dy.device)
    call("sib_blurpool_bwd", _p(dy), _p(dx), n, h, w, c, _stream())
    return dx


def avgpool2_fwd(x):
    n, c, h, w = x.shape
    y = new_act(n, c, h // 2, w // 2, x.device)
    call("sib_avgpool2_fwd", _p(x), _p(y), n, h, w, c, _stream())
    return y


def avgpool2_bwd(dy, x_shape):
    n, c, h, w = x_shape
    dx = new_act(n, c, h, w, dy.device)
    call("sib_avgpool2_bwd", _p(dy), _p(dx), n, h, w, c, _stream())
    return dx


def maxpool3x3s1_fwd(x):
    n, c, h, w = x.shape
    y = new_act(n, c, h, w, x.device)
    idx = torch.empty((n, h, w, c), dtype=torch.uint8, device=x.device)
    call("sib_maxpool3x3s1_fwd", _p(x), _p(y), _p(idx), n, h, w, c, _stream())
    return y, idx


def maxpool3x3s1_bwd(dy, idx):
    n, c, h, w = dy.shape
    dx = new_act(n, c, h, w, dy.device)
    call("sib_maxpool3x3s1_bwd", _p(dy), _p(idx), _p(dx), n, h, w, c, _stream())
    return dx


def chan_reduce(a, b=None, scale=1.0):
    n, c, h, w = a.shape
    out = torch.empty((n, c), dtype=torch.float32, device=a.device)
    call("sib_chan_reduce", _p(a), _p(b), _p(out), n, h * w, c, float(scale), _stream())
    return out


def scale_nc(x, mul, add=None):
    n, c, h, w = x.shape
    y = new_act(n, c, h, w, x.device)
    call("sib_scale_nc", _p(x), _p(mul), _p(add), _p(y), n, h * w, c, _stream())
    return y


def eca_gate_fwd(p, w):
    s = torch.empty_like(p)
    call("sib_eca_gate_fwd", _p(p), _p(w), _p(s), p.shape[0], p.shape[1], _stream())
    return s


def eca_gate_bwd(ds, s, p, w, dw):
    dp = torch.empty_like(p)
    call("sib_eca_gate_bwd", _p(ds), _p(s), _p(p), _p(w), _p(dp), _p(dw), p.shape[0], p.shape[1], _stream())
    return dp


def scale_add_act(x, mul, res, act, slope=0.01, add=None):
    """act(x * mul[n, c] (+ add[n, c]) + res) in one pass (x, res: bf16 channels_last of equal shape)."""
    n, c, h, w = x.shape
    assert res.shape == x.shape
    y = new_act(n, c, h, w, x.device)
    call("sib_scale_add_act", _p(x), _p(mul), _p(add), _p(res), _p(y), n, h * w, c, act, float(slope), _stream())
    return y


def act_bwd_reduce(dy, y, x, act, slope=0.01):
    """g = dy * act'(y); s1[n, c] = sum_hw g, s2[n, c] = sum_hw g * x  -> g, s1, s2."""
    n, c, h, w = y.shape
    g = torch.empty_like(y)
    s1 = torch.empty((n, c), dtype=torch.float32, device=y.device)
    s2 = torch.empty((n, c), dtype=torch.float32, device=y.device)
    call("sib_act_bwd_reduce", _p(dy), _p(y), _p(x), _p(g), _p(s1), _p(s2), n, h * w, c, act, float(slope), _stream())
    return g, s1, s2


def bn_bwd_apply_scaled(dy, dy_mul, dy_add, x, mean_invstd, gamma, sums, count, param_grads=(None, None),
                        pgrad_scale=1.0):
    """BN backward (no activation) of the gradient dy * dy_mul[n, c] + dy_add[n, c] -> dx."""
    n, c, h, w = x.shape
    if len(param_grads) > 2:
        pgrad_scale = param_grads[2]
    dx = new_act(n, c, h, w, x.device)
    call("sib_bn_bwd_apply_scaled", _p(dy), _p(dy_mul), _p(dy_add), _p(x), _p(mean_invstd), _p(gamma), _p(sums),
         _p(dx), _p(param_grads[0]), _p(param_grads[1]), n, h * w, c, float(count), float(pgrad_scale), _stream())
    return dx


def add_act(a, b, act, slope=0.01):
    y = torch.empty_like(a)
    call("sib_add_act", _p(a), _p(b), _p(y), a.numel(), act, float(slope), _stream())
    return y


def act_bwd(dy, y, act, slope=0.01):
    g = torch.empty_like(y)
    call("sib_act_bwd", _p(dy), _p(y), _p(g), y.numel(), act, float(slope), _stream())
    return g


def weight_standardize(w_flat, out_bf16, mean_invstd, out_channels, fan, eps):
    call("sib_weight_standardize", _p(w_flat), c_void_p(0), _p(out_bf16), _p(mean_invstd), out_channels, fan,
         float(eps), _stream())


def weight_standardize_bwd(w_flat, mean_invstd, g_flat, out_channels, fan):
    call("sib_weight_standardize_bwd", _p(w_flat), c_void_p(0), _p(mean_invstd), _p(g_flat), _p(g_flat),
         out_channels, fan, _stream())


# ------------------------------------------------------------------ heads
def ce_fwd_bwd(logits, target, smoothing=0.0, temperature=1.0, margin_kind=MARGIN_NONE, s=1.0,
               m=0.0, want_grad=True, grad_scale=1.0):
    """Returns (mean loss [0-dim fp32], per-row loss, dlogits or None)."""
    _lib.require_device()
    assert logits.dim() == 2 and logits.is_cuda and logits.stride(1) == 1
    b, c = logits.shape
    fp32 = logits.dtype == torch.float32
    if not fp32 and logits.dtype != torch.bfloat16:
        raise _lib.SibError("logits must be float32 or bfloat16")
    labels = dense = None
    if target.dim() == 1:
        labels = target.to(torch.int64).contiguous()
    else:
        dense = target.to(torch.float32).contiguous()
    loss_rows = torch.empty((b,), dtype=torch.float32, device=logits.device)
    loss = torch.empty((), dtype=torch.float32, device=logits.device)
    dlogits = torch.empty_like(logits) if want_grad else None
    call("sib_ce_fwd_bwd", _p(logits), int(fp32), _p(labels), _p(dense), b, c, logits.stride(0),
         float(smoothing), float(temperature), margin_kind, float(s), float(m), _p(loss_rows),
         _p(loss), _p(dlogits), float(grad_scale), _stream())
    return loss, loss_rows, dlogits


def sphere_linear_fwd(x, w, normalize_x=True):
    b, d = x.shape
    c = w.shape[0]
    dev = x.device
    cosv = torch.empty((b, c), dtype=torch.float32, device=dev)
    xn = torch.empty_like(x) if normalize_x else None
    wn = torch.empty_like(w)
    xnorm = torch.empty((b,), dtype=torch.float32, device=dev) if normalize_x else None
    wnorm = torch.empty((c,), dtype=torch.float32, device=dev)
    call("sib_sphere_linear_fwd", _p(x), _p(w), _p(cosv), _p(xn), _p(wn), _p(xnorm), _p(wnorm), b,
         c, d, int(normalize_x), _stream())
    return cosv, (xn if normalize_x else x, wn, xnorm, wnorm)


def sphere_linear_bwd(dcos, saved, need_dx=True, need_dw=True, normalize_x=True):
    xn, wn, xnorm, wnorm = saved
    b, d = xn.shape
    c = wn.shape[0]
    dx = torch.empty_like(xn) if need_dx else None
    dw = torch.empty_like(wn) if need_dw else None
    scratch = torch.empty((b + c, d), dtype=torch.float32, device=xn.device)
    call("sib_sphere_linear_bwd", _p(dcos), _p(xn), _p(wn), _p(xnorm), _p(wnorm), _p(dx), _p(dw),
         _p(scratch), b, c, d, int(normalize_x), _stream())
    return dx, dw


# ------------------------------------------------------------------ optimizer helpers
def sgd_segments(records, device):
    """records: (end, lr, weight_decay, momentum, dampening, nesterov) -> device byte tensor."""
    import numpy as np
    rec = np.zeros(len(records), dtype=np.dtype([("end", "<i8"), ("lr", "<f4"), ("wd", "<f4"),
                                                 ("mom", "<f4"), ("damp", "<f4"), ("nest", "<i4"),
                                                 ("pad", "<i4")]))
    for i, r in enumerate(records):
        rec[i] = (r[0], r[1], r[2], r[3], r[4], int(r[5]), 0)
    return torch.from_numpy(rec.view(np.uint8).copy()).to(device, non_blocking=True)


def sgd_step(params, grads, buf, params_bf16, segs, nseg, first_step, ema=None, ema_decay=0.0):
    call("sib_sgd_step", _p(params), _p(grads), _p(buf), _p(params_bf16), _p(ema), float(ema_decay),
         _p(segs), nseg, params.numel(), int(first_step), _stream())


def novograd_table(records):
    """records: (begin, end, unit_len, ngroups, norm_base, lr, decay, beta1, 1-beta1, beta2,
    1-beta2) per tensor -> host byte tensor laid out like csrc/optim.cu NovoTensor."""
    import numpy as np
    rec = np.zeros(len(records), dtype=np.dtype([
        ("begin", "<i8"), ("end", "<i8"), ("unit", "<i4"), ("ngroups", "<i4"), ("base", "<i4"),
        ("lr", "<f4"), ("decay", "<f4"), ("b1", "<f4"), ("omb1", "<f4"), ("b2", "<f4"),
        ("omb2", "<f4"), ("pad", "<i4")]))
    for i, r in enumerate(records):
        rec[i] = tuple(r) + (0,)
    return torch.from_numpy(rec.view(np.uint8).copy())


def novograd_step(params, grads, ema_grad, params_bf16, table, ntensors, sumsq, ema_norm, denom,
                  eps, unitwise, ema=None, ema_decay=0.0):
    call("sib_novograd_step", _p(params), _p(grads), _p(ema_grad), _p(params_bf16), _p(ema),
         float(ema_decay), _p(table), ntensors, _p(sumsq), _p(ema_norm), _p(denom), params.numel(),
         float(eps), int(bool(unitwise)), _stream())


def cast_bf16(src, dst):
    call("sib_cast_bf16", _p(src), _p(dst), src.numel(), _stream())


# ------------------------------------------------------------------ data
def rrc_boxes(batch, h, w, min_area, max_area, seed, first_sample, do_flip, device):
    boxes = torch.empty((batch, 5), dtype=torch.int32, device=device)
    call("sib_rrc_boxes", _p(boxes), batch, h, w, float(min_area), float(max_area), int(seed),
         int(first_sample), int(do_flip), _stream())
    return boxes


def rrc_box_host(h, w, min_area, max_area, seed, sample):
    import ctypes
    box = (ctypes.c_int * 5)()
    _lib.load().sib_rrc_box_host(h, w, float(min_area), float(max_area), int(seed), int(sample), box)
    return list(box)


def augment(src_u8, boxes, size, mean=127.5, std=51.0, out_mode=0):
    b, sh, sw, ch = src_u8.shape
    assert ch == 3 and src_u8.dtype == torch.uint8 and src_u8.is_contiguous()
    if out_mode == 0:
        out = new_act(b, 4, size, size, src_u8.device)
    else:
        out = torch.empty((b, 3, size, size), dtype=torch.float32, device=src_u8.device)
    call("sib_augment", _p(src_u8), _p(boxes), _p(out), b, sh, sw, size, float(mean), float(std),
         out_mode, _stream())
    return out


def val_transform(src_u8, size, resize_shorter, mean=127.5, std=51.0, out_mode=0):
    """resize-shorter + centre crop + normalise (reference val_pipeline, dali_dataloader.py:146-160)"""
    b, sh, sw, ch = src_u8.shape
    assert ch == 3 and src_u8.dtype == torch.uint8 and src_u8.is_contiguous()
    if out_mode == 0:
        out = new_act(b, 4, size, size, src_u8.device)
    else:
        out = torch.empty((b, 3, size, size), dtype=torch.float32, device=src_u8.device)
    call("sib_val_transform", _p(src_u8), _p(out), b, sh, sw, size, int(resize_shorter), float(mean),
         float(std), out_mode, _stream())
    return out


def val_geometry_host(sh, sw, size, resize_shorter):
    import ctypes
    g = (ctypes.c_int * 4)()
    _lib.load().sib_val_geometry_host(sh, sw, size, int(resize_shorter), g)
    return list(g)


def _check_ragged(packed, offsets, dims):
    if packed.dtype != torch.uint8 or not packed.is_cuda or not packed.is_contiguous():
        raise _lib.SibError("ragged batch: packed buffer must be a contiguous uint8 CUDA tensor")
    if offsets.dtype != torch.int64 or dims.dtype != torch.int32 or dims.dim() != 2 or dims.shape[1] != 2 \
            or offsets.numel() != dims.shape[0] or not offsets.is_cuda or not dims.is_cuda:
        raise _lib.SibError("ragged batch: offsets int64 [B] and dims int32 [B, 2] on the device expected")
    return dims.shape[0]


def rrc_boxes_ragged(dims, min_area, max_area, seed, first_sample, do_flip):
    """Crop boxes for images of different sizes (dims int32 [B, 2] = {H, W} on the device)."""
    b = dims.shape[0]
    boxes = torch.empty((b, 5), dtype=torch.int32, device=dims.device)
    call("sib_rrc_boxes_ragged", _p(boxes), _p(dims.contiguous()), b, float(min_area), float(max_area),
         int(seed), int(first_sample), int(do_flip), _stream())
    return boxes


def augment_ragged(packed, offsets, dims, boxes, size, mean=127.5, std=51.0, out_mode=0):
    b = _check_ragged(packed, offsets, dims)
    if out_mode == 0:
        out = new_act(b, 4, size, size, packed.device)
    else:
        out = torch.empty((b, 3, size, size), dtype=torch.float32, device=packed.device)
    call("sib_augment_ragged", _p(packed), _p(offsets.contiguous()), _p(dims.contiguous()), _p(boxes), _p(out),
         b, size, float(mean), float(std), out_mode, _stream())
    return out


def val_transform_ragged(packed, offsets, dims, size, resize_shorter, mean=127.5, std=51.0, out_mode=0):
    b = _check_ragged(packed, offsets, dims)
    if out_mode == 0:
        out = new_act(b, 4, size, size, packed.device)
    else:
        out = torch.empty((b, 3, size, size), dtype=torch.float32, device=packed.device)
    call("sib_val_transform_ragged", _p(packed), _p(offsets.contiguous()), _p(dims.contiguous()), _p(out), b,
         size, int(resize_shorter), float(mean), float(std), out_mode, _stream())
    return out


def jpeg_idct_rgb(coef, table, n_images, max_blocks, max_pixels, planes, out):
    """Device half of the hybrid JPEG decoder (jpeg.decode_batch): int16 coefficients + per-image
    descriptor table (jpeg.IMAGE_DTYPE records as bytes) -> packed uint8 RGB images in `out`."""
    _lib.require_device()
    assert coef.dtype == torch.int16 and coef.is_cuda and table.dtype == torch.uint8 and table.is_cuda
    assert planes.dtype == torch.uint8 and out.dtype == torch.uint8 and planes.is_cuda and out.is_cuda
    call("sib_jpeg_idct_rgb", _p(coef), _p(table), int(n_images), int(max_blocks), int(max_pixels), _p(planes),
         _p(out), _stream())
    return out


def one_hot(labels, num_classes):
    out = torch.empty((labels.shape[0], num_classes), dtype=torch.float32, device=labels.device)
    call("sib_one_hot", _p(labels), _p(out), labels.shape[0], num_classes, _stream())
    return out


def mix_batch(x, prev, perm, mode, lam=1.0, one_minus_lam=0.0, box=(0, 0, 0, 0)):
    """Mixup (mode 0) / CutMix (mode 1) of a resident batch with a permuted previous batch.
    x, prev: bf16 channels_last [N,C,H,W] or fp32 NCHW; perm int32 [N] on the device;
    box = (h1, w1, h2, w2).  Out of place (prev may be x itself)."""
    _lib.require_device()
    n, c, h, w = x.shape
    ph, pw = prev.shape[2], prev.shape[3]
    if prev.shape[0] != n or prev.shape[1] != c or prev.dtype != x.dtype:
        raise _lib.SibError("mix_batch: previous batch must have the same batch size, channels and dtype")
    if x.dtype == torch.bfloat16 and x.permute(0, 2, 3, 1).is_contiguous() and \
            prev.permute(0, 2, 3, 1).is_contiguous():
        layout = 0
    elif x.dtype == torch.float32 and x.is_contiguous() and prev.is_contiguous():
        layout = 1
    else:
        raise _lib.SibError("mix_batch: batches must be bf16 channels_last or fp32 NCHW contiguous")
    if perm.dtype != torch.int32 or not perm.is_cuda or perm.numel() != n:
        raise _lib.SibError("mix_batch: perm must be an int32 device tensor of the batch size")
    out = torch.empty_like(x)
    h1, w1, h2, w2 = (int(v) for v in box)
    call("sib_mix_batch", _p(x), _p(prev), _p(perm), _p(out), n, h, w, c, ph, pw, layout, int(mode),
         float(lam), float(one_minus_lam), h1, w1, h2, w2, _stream())
    return out


def _aug_layout(x):
    if x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[1] == 4 and x.permute(0, 2, 3, 1).is_contiguous():
        return 0
    if x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3 and x.is_contiguous():
        return 1
    raise _lib.SibError("batch augmentation: bf16 channels_last [N,4,H,W] or fp32 NCHW [N,3,H,W] expected")


def pixel_ops_(x, params, crop_boxes=None, nboxes=0):
    """Colour twist / grayscale / random erasing of the resident batch, in place (csrc/batchaug.cu).
    params: fp32 [N, 16 + 4 * nboxes] on the device; crop_boxes: the int32 [N, 5] boxes of rrc_boxes."""
    _lib.require_device()
    n, _, h, w = x.shape
    if params.dtype != torch.float32 or not params.is_cuda or tuple(params.shape) != (n, 16 + 4 * nboxes):
        raise _lib.SibError("pixel_ops: params must be a CUDA fp32 tensor [N, 16 + 4 * nboxes]")
    call("sib_pixel_ops", _p(x), _p(params.contiguous()), _p(crop_boxes), n, h, w, _aug_layout(x), nboxes, _stream())
    return x


def gaussian_blur(x, sigma):
    """11-tap Gaussian blur per sample (sigma fp32 [N] on the device, <= 0: copied through)."""
    _lib.require_device()
    n, _, h, w = x.shape
    out = torch.empty_like(x)
    if x.dtype == torch.bfloat16:
        out.zero_()                  # the zero 4th channel
    call("sib_gaussian_blur", _p(x), _p(out), _p(sigma.contiguous()), n, h, w, _aug_layout(x), _stream())
    return out


def mix_targets(t, prev_t, perm, w_self, w_prev):
    _lib.require_device()
    if t.dtype != torch.float32 or prev_t.dtype != torch.float32 or t.shape != prev_t.shape:
        raise _lib.SibError("mix_targets: dense fp32 targets of equal shape expected")
    t, prev_t = t.contiguous(), prev_t.contiguous()
    out = torch.empty_like(t)
    call("sib_mix_targets", _p(t), _p(prev_t), _p(perm), _p(out), t.shape[0], t.shape[1],
         float(w_self), float(w_prev), _stream())
    return out


def stem_pack(x, kw, pad_w):
    """[N,3,H,W] fp32 NCHW or [N,4,H,W] bf16 channels_last -> [N,64,H/2,W/2] packed rows."""
    n, c, h, w = x.shape
    if x.dtype == torch.float32 and c == 3 and x.is_contiguous():
        mode = 1
    elif x.dtype == torch.bfloat16 and c == 4 and x.permute(0, 2, 3, 1).is_contiguous():
        mode = 0
    else:
        raise _lib.SibError("stem input must be fp32 NCHW [N,3,H,W] or bf16 channels_last [N,4,H,W]")
    xq = new_act(n, 64, h // 2, w // 2, x.device)
    call("sib_stem_pack", _p(x), _p(xq), n, h, w, kw, pad_w, mode, _stream())
    return xq


def stem_pack_weight(w, na, off, out=None):
    k, c, kh, kw = w.shape
    assert c == 3 and w.is_contiguous() and w.dtype == torch.float32
    wq = out if out is not None else torch.empty((k, na, 1, 64), dtype=torch.bfloat16, device=w.device)
    call("sib_stem_pack_weight", _p(w), _p(wq), k, kh, kw, na, off, _stream())
    return wq


def stem_unpack_wgrad(dwq, dw, na, off, accumulate=True):
    k, c, kh, kw = dw.shape
    call("sib_stem_unpack_wgrad", _p(dwq), _p(dw), k, kh, kw, na, off, int(accumulate), _stream())
