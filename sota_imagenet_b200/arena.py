"""Flat parameter arena: one fp32 buffer for all parameters of a model (plus gradient, bf16
shadow and dgrad-packed filter buffers), with the nn.Parameters re-pointed at views of it.

Why: the fused SGD kernel updates all 161 ResNet-50 tensors in one launch and refreshes the
bf16 filters the conv kernels read (reference path: torch.optim._multi_tensor.SGD foreach
launches, arg_parser.py:136-138); gradient buckets for the data-parallel all-reduce are
contiguous slices of the gradient buffer.
"""
import torch

from . import ops

ALIGN = 128  # elements; keeps every tensor 256B (bf16) / 512B (fp32) aligned for TMA


def _round_up(n, a=ALIGN):
    return (n + a - 1) // a * a


class ParamArena:
    def __init__(self, named_params, device):
        self.device = device
        self.entries = []   # (name, param, offset, numel, layout)
        off = 0
        for name, p in named_params:
            layout = getattr(p, "_sib_layout", "plain")
            self.entries.append((name, p, off, p.numel(), layout))
            off += _round_up(p.numel())
        self.total = max(off, ALIGN)
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.grad = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.shadow = torch.zeros(self.total, dtype=torch.bfloat16, device=device)
        self.momentum = None       # allocated by the optimizer
        self.ema = None
        self.offset_of = {}
        # dgrad-pack arena for conv filters that need a data gradient
        pack_entries, doff = [], 0
        self.dgrad_offset = {}
        for name, p, o, n, layout in self.entries:
            view = self.view_of(self.flat, o, p.shape, layout)
            view.copy_(p.data)
            p.data = view
            p.grad = self.view_of(self.grad, o, p.shape, layout)
            p._sib_arena = self
            p._sib_offset = o
            self.offset_of[id(p)] = o
            if layout == "krsc" and getattr(p, "_sib_needs_dgrad", False):
                k, c, r, s = p.shape
                pack_entries.append((o, doff, k, r * s, c))
                self.dgrad_offset[id(p)] = doff
                doff += _round_up(n)
        self.wdgrad = torch.zeros(max(doff, ALIGN), dtype=torch.bfloat16, device=device)
        # row-parity sub-filters for the 3x3 / stride-2 dgrads (ops.pack_dgrad_s2)
        self.s2 = {}
        for name, p, o, n, layout in self.entries:
            if id(p) in self.dgrad_offset and getattr(p, "_sib_stride", 1) == 2 and tuple(p.shape[2:]) == (3, 3) \
                    and p.shape[0] % 8 == 0:
                k, c = p.shape[0], p.shape[1]
                self.s2[id(p)] = (p, torch.zeros((2 * c, 1, 2, k), dtype=torch.bfloat16, device=device),
                                  torch.zeros((2 * c, 2, 2, k), dtype=torch.bfloat16, device=device))
        self.pack_table, self.pack_blocks = ops.pack_table(pack_entries, device)
        self.pack_count = len(pack_entries)
        self._sig = None
        # weight standardisation (reference train.py:66-67): (param, mean_invstd buffer) pairs whose
        # bf16 shadow holds (w - mean) / sqrt(var + eps) per output channel instead of w itself
        self.ws_entries = []
        self.ws_eps = 1e-7

    @staticmethod
    def view_of(buf, off, shape, layout):
        n = 1
        for d in shape:
            n *= d
        flat = buf[off:off + n]
        if layout == "krsc" and len(shape) == 4:
            k, c, r, s = shape
            return flat.view(k, r, s, c).permute(0, 3, 1, 2)
        return flat.view(shape)

    # ---- bf16 shadow maintenance -------------------------------------------------
    def signature(self):
        return sum(p._version for _, p, _, _, _ in self.entries)

    def refresh_shadow(self, force=False):
        """Recast fp32 -> bf16 + repack dgrad filters when parameters were modified by anything
        other than the fused optimizer (load_state_dict, init, a foreign optimizer)."""
        sig = self.signature()
        if force or sig != self._sig:
            ops.cast_bf16(self.flat, self.shadow)
            self.standardize_shadow()
            self.repack_dgrad()
            self._sig = sig

    def repack_dgrad(self):
        if self.pack_count:
            ops.call("sib_pack_dgrad_weights", ops._p(self.shadow), ops._p(self.wdgrad),
                     ops._p(self.pack_table), self.pack_count, self.pack_blocks, ops._stream())
        for p, sub0, sub1 in self.s2.values():
            ops.pack_dgrad_s2(self.dgrad_view(p), sub0, sub1)

    def dgrad_s2_view(self, p):
        e = self.s2.get(id(p))
        return (e[1], e[2]) if e is not None else None

    def enable_weight_standardization(self, params, eps=1e-7):
        """One table for all standardised filters: forward (shadow refresh) and backward (gradient
        transform) are ONE launch each over the flat arenas instead of one per tensor."""
        import numpy as np
        self.ws_eps = eps
        self.ws_entries = []
        total_k = sum(p.shape[0] for p in params)
        self.ws_mi = torch.empty((max(total_k, 1), 2), dtype=torch.float32, device=self.device)
        rec = np.zeros(len(params), dtype=np.dtype([("off", "<i8"), ("mi_off", "<i8"), ("K", "<i4"),
                                                    ("fan", "<i4"), ("bb", "<i4"), ("pad", "<i4")]))
        kb = 0
        for i, p in enumerate(params):
            k = p.shape[0]
            rec[i] = (self.offset_of[id(p)], kb, k, p.numel() // k, kb, 0)
            self.ws_entries.append((p, self.ws_mi[kb:kb + k]))
            kb += k
        self.ws_blocks = kb
        self.ws_table = torch.from_numpy(rec.view(np.uint8).copy()).to(self.device)
        self._sig = None

    def standardize_shadow(self):
        if self.ws_entries:
            ops.call("sib_weight_standardize_batch", ops._p(self.flat), ops._p(self.shadow), ops._p(self.ws_mi),
                     ops._p(self.ws_table), len(self.ws_entries), self.ws_blocks, float(self.ws_eps), ops._stream())

    def standardize_grads(self):
        """dL/dw from dL/d(standardised w), in place in the gradient arena (end of backward)."""
        if self.ws_entries:
            ops.call("sib_weight_standardize_bwd_batch", ops._p(self.flat), ops._p(self.ws_mi), ops._p(self.grad),
                     ops._p(self.ws_table), len(self.ws_entries), self.ws_blocks, ops._stream())

    def mark_fresh(self):
        """Called by the fused optimizer after it rewrote params + shadow itself."""
        self.standardize_shadow()
        self.repack_dgrad()
        self._sig = self.signature()

    def shadow_view(self, p):
        o = self.offset_of[id(p)]
        return self.view_of(self.shadow, o, p.shape, getattr(p, "_sib_layout", "plain"))

    def dgrad_view(self, p):
        k, c, r, s = p.shape
        o = self.dgrad_offset[id(p)]
        return self.wdgrad[o:o + p.numel()].view(c, r, s, k)

    def grad_view(self, p):
        o = self.offset_of[id(p)]
        return self.view_of(self.grad, o, p.shape, getattr(p, "_sib_layout", "plain"))

    # ---- gradient bookkeeping ------------------------------------------------------
    def prepare_grads(self):
        """Start of a backward pass: if grads were dropped (zero_grad(set_to_none=True)) clear the
        flat buffer with one memset and re-attach the views; otherwise keep accumulating."""
        first = self.entries[0][1]
        if first.grad is None:
            self.grad.zero_()
            for _, p, o, _, layout in self.entries:
                p.grad = self.view_of(self.grad, o, p.shape, layout)

    def zero_grad(self):
        self.grad.zero_()
        for _, p, o, _, layout in self.entries:
            if p.grad is None:
                p.grad = self.view_of(self.grad, o, p.shape, layout)

    def intact(self):
        """True while every parameter still aliases the arena (a later .to()/.cuda() breaks it)."""
        base = self.flat.data_ptr()
        for _, p, o, _, _ in self.entries:
            if p.data_ptr() != base + 4 * o:
                return False
        return True
