"""torch.optim.Optimizer drop-ins backed by the multi-tensor kernels.

SGD replaces `torch.optim._multi_tensor.SGD` (reference arg_parser.py:136-138, hyper-parameters
from configs/hydra_exp/1.r50_baseline.yaml:29-31) with torch/optim/sgd.py arithmetic: coupled
L2 weight decay, momentum buffer initialised to the first gradient, optional Nesterov.  One
kernel launch per parameter arena updates every tensor, rewrites the bf16 filter shadows and
(optionally) an EMA copy of the weights (pt_clb.ModelEma, reference train.py:112).

MyNovograd replaces the reference's own `sota_imagenet/optimizers.py:35-161` (three launches per
arena: segmented sum of squares, per-group running norm, elementwise update).
"""
import torch

from . import _lib, ops
from .arena import ParamArena


class SGD(torch.optim.Optimizer):
    def __init__(self, params, lr=0.0, momentum=0.0, dampening=0.0, weight_decay=0.0,
                 nesterov=False, ema_decay=0.0, **unused):
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        defaults = dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                        nesterov=nesterov)
        super().__init__(params, defaults)
        self.ema_decay = ema_decay
        self._arenas = None
        self._steps = 0
        self._seg_cache = {}

    # ------------------------------------------------------------------ arenas
    def _collect_arenas(self):
        arenas, loose = [], []
        for group in self.param_groups:
            for p in group["params"]:
                a = getattr(p, "_sib_arena", None)
                if a is not None and a.intact():
                    if all(a is not b for b in arenas):
                        arenas.append(a)
                else:
                    loose.append(p)
        if loose:
            if not loose[0].is_cuda:
                raise _lib.SibError("fused SGD needs CUDA parameters (no CPU fallback)")
            grads = [p.grad for p in loose]
            a = ParamArena([("loose%d" % i, p) for i, p in enumerate(loose)], loose[0].device)
            for p, g in zip(loose, grads):       # keep autograd-produced grads
                p.grad = g
            a._loose = True
            arenas.append(a)
        for a in arenas:
            if a.momentum is None:
                a.momentum = torch.zeros_like(a.flat)
            if self.ema_decay and a.ema is None:
                a.ema = a.flat.clone()
        self._arenas = arenas
        # expose momentum buffers in the torch.optim.SGD state format.  When the arenas were
        # re-collected (a .cuda()/.to() or an arena rebuilt in registration order) the state
        # accumulated so far MIGRATES into the new buffers instead of being silently zeroed.
        for a in arenas:
            for _, p, o, n, layout in a.entries:
                view = ParamArena.view_of(a.momentum, o, p.shape, layout)
                old = self.state[p].get("momentum_buffer") if p in self.state else None
                if old is not None and old.data_ptr() != view.data_ptr() and old.shape == view.shape:
                    view.copy_(old)
                self.state[p]["momentum_buffer"] = view

    def _segments(self, arena):
        gmap = {}
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                gmap[id(p)] = gi
        recs = []
        for i, (_, p, o, n, _) in enumerate(arena.entries):
            end = arena.entries[i + 1][2] if i + 1 < len(arena.entries) else arena.total
            gi = gmap.get(id(p))
            if gi is None:
                hp = (0.0, 0.0, 0.0, 0.0, False)
            else:
                g = self.param_groups[gi]
                hp = (float(g["lr"]), float(g["weight_decay"]), float(g["momentum"]),
                      float(g["dampening"]), bool(g["nesterov"]))
            if recs and recs[-1][1:] == hp:
                recs[-1] = (end,) + hp
            else:
                recs.append((end,) + hp)
        return recs

    # ------------------------------------------------------------------ API
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._arenas is None or not all(a.intact() for a in self._arenas):
            self._collect_arenas()
        first = self._steps == 0
        for a in self._arenas:
            for _, p, o, n, layout in a.entries:
                if p.grad is None:
                    continue
                gv = ParamArena.view_of(a.grad, o, p.shape, layout)
                if p.grad.data_ptr() != gv.data_ptr():
                    gv.copy_(p.grad)
            cached = self._upload_segments(a)
            recs = cached[0]
            ops.sgd_step(a.flat, a.grad, a.momentum, a.shadow, cached[1], len(recs), first,
                         ema=a.ema if self.ema_decay else None, ema_decay=self.ema_decay)
            a.mark_fresh()
        self._steps += 1
        return loss

    def _upload_segments(self, a):
        """Per-arena hyper-parameter table (lr, weight decay, momentum, ... per segment) in a
        PERSISTENT device buffer: when a value changes, only its contents are rewritten (async copy
        from a pinned host table), the address stays -- so a captured CUDA graph of step() keeps
        reading live values (sync_hyperparams)."""
        recs = self._segments(a)
        key = id(a)
        cached = self._seg_cache.get(key)
        if cached is None or cached[0] != recs:
            host = ops.sgd_segments(recs, "cpu").pin_memory()
            if cached is None or cached[1].numel() != host.numel():
                if cached is not None and torch.cuda.is_current_stream_capturing():
                    raise _lib.SibError("optimizer param-group structure changed inside a CUDA graph capture")
                dev = host.to(a.device, non_blocking=True)
            else:
                dev = cached[1]
                dev.copy_(host, non_blocking=True)
            self._seg_cache[key] = (recs, dev, host)
            cached = self._seg_cache[key]
        return cached

    def sync_hyperparams(self):
        """Push the current param_group values (a per-batch LR schedule) to the device tables WITHOUT
        running step(): call before replaying a CUDA graph that captured step()."""
        for a in self._arenas or []:
            self._upload_segments(a)

    def zero_grad(self, set_to_none=True):
        if self._arenas is None:
            return super().zero_grad(set_to_none)
        for a in self._arenas:
            if getattr(a, "_loose", False):
                for _, p, _, _, _ in a.entries:
                    p.grad = None if set_to_none else (p.grad.zero_() if p.grad is not None else None)
            else:
                a.zero_grad()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._collect_arenas_after_load()

    def _collect_arenas_after_load(self):
        loaded = {p: dict(s) for p, s in self.state.items()}
        self._collect_arenas()
        any_buf = False
        for p, s in loaded.items():
            buf = s.get("momentum_buffer")
            if buf is not None:
                self.state[p]["momentum_buffer"].copy_(buf)
                any_buf = True
        if any_buf:
            self._steps = max(self._steps, 1)

    def ema_state_dict(self, module):
        """EMA weights keyed like module.state_dict() (what ModelEma would hold)."""
        out = {}
        for a in self._arenas or []:
            if a.ema is None:
                continue
            names = {id(p): n for n, p in module.named_parameters()}
            for _, p, o, n, layout in a.entries:
                if id(p) in names:
                    out[names[id(p)]] = ParamArena.view_of(a.ema, o, p.shape, layout).clone()
        return out


FusedSGD = SGD


class MyNovograd(torch.optim.Optimizer):
    """Reference `sota_imagenet/optimizers.py:35-161`, same constructor and state keys.

    ema_norm <- b2*ema_norm + (1-b2)*norm,  norm = sum(p^2) per tensor (the reference feeds the
    WEIGHTS, :136) or ||p||_2 per output unit when `unitwise_norm` (:18-22,133-134);
    ema_grad <- b1*ema_grad + (1-b1)*g;  p <- (p - lr*ema_grad/(sqrt(ema_norm)+eps))*(1-lr*wd).
    `state[p]["ema_norm"]` is an expanded view of one scalar per norm group (the reference
    stores the expanded tensor, :120-121)."""

    def __init__(self, params, lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-2,
                 ema_norm_init=1e-3, unitwise_norm=False, ema_decay=0.0):
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 0: {betas[0]}")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 1: {betas[1]}")
        if not 0.0 <= weight_decay:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        defaults = dict(lr=lr, betas=betas, weight_decay=weight_decay, ema_norm_init=ema_norm_init)
        super().__init__(params, defaults)
        self.eps = eps
        self.unitwise_norm = unitwise_norm
        self.ema_decay = ema_decay
        self._arenas = None
        self._tab_cache = {}

    # ------------------------------------------------------------------ layout
    @staticmethod
    def norm_groups(shape, unitwise):
        """(unit_len, ngroups) of one parameter: one group per tensor, or per output unit
        (first dimension) for >1-D tensors under unitwise_norm (reference :18-22)."""
        numel = 1
        for d in shape:
            numel *= d
        if unitwise and len(shape) > 1 and numel > 0:
            return numel // shape[0], shape[0]
        return max(numel, 1), 1

    def _collect_arenas(self):
        arenas, loose = [], []
        for group in self.param_groups:
            for p in group["params"]:
                a = getattr(p, "_sib_arena", None)
                if a is not None and a.intact():
                    if all(a is not b for b in arenas):
                        arenas.append(a)
                else:
                    loose.append(p)
        if loose:
            if not loose[0].is_cuda:
                raise _lib.SibError("fused MyNovograd needs CUDA parameters (no CPU fallback)")
            grads = [p.grad for p in loose]
            a = ParamArena([("loose%d" % i, p) for i, p in enumerate(loose)], loose[0].device)
            for p, g in zip(loose, grads):
                p.grad = g
            a._loose = True
            arenas.append(a)
        gmap = {id(p): g for g in self.param_groups for p in g["params"]}
        for a in arenas:
            a.novo_ema_grad = torch.zeros_like(a.flat)
            base, layout = 0, []
            for _, p, o, n, lay in a.entries:
                unit, ng = self.norm_groups(tuple(p.shape), self.unitwise_norm)
                layout.append((unit, ng, base))
                base += ng
            a.novo_layout = layout
            a.novo_sumsq = torch.zeros(base, dtype=torch.float32, device=a.device)
            a.novo_denom = torch.zeros(base, dtype=torch.float32, device=a.device)
            init = torch.zeros(base, dtype=torch.float32)
            for (_, p, _, _, _), (unit, ng, nb) in zip(a.entries, layout):
                g = gmap.get(id(p))
                init[nb:nb + ng] = float(g["ema_norm_init"]) if g is not None else 1.0
            a.novo_ema_norm = init.to(a.device)
            if self.ema_decay and a.ema is None:
                a.ema = a.flat.clone()
        self._arenas = arenas
        for a in arenas:
            for (_, p, o, n, lay), (unit, ng, nb) in zip(a.entries, a.novo_layout):
                st = self.state[p]
                st.setdefault("step", 0)
                view = ParamArena.view_of(a.novo_ema_grad, o, p.shape, lay)
                norms = a.novo_ema_norm[nb:nb + ng]
                # re-collection (arena rebuilt / moved): accumulated state migrates, it is not reset
                old_g, old_n = st.get("ema_grad"), st.get("ema_norm")
                if old_g is not None and old_g.data_ptr() != view.data_ptr() and old_g.shape == view.shape:
                    view.copy_(old_g)
                if old_n is not None and old_n.numel() == p.numel() and p.numel() > 0 and \
                        old_n.data_ptr() != norms.data_ptr():
                    norms.copy_(old_n.reshape(ng, -1)[:, 0])
                st["ema_grad"] = view
                if ng == 1:
                    st["ema_norm"] = norms.expand(p.numel()).view(p.shape) if p.dim() == 0 else \
                        norms.view((1,) * p.dim()).expand(p.shape)
                else:
                    st["ema_norm"] = norms.view((ng,) + (1,) * (p.dim() - 1)).expand(p.shape)

    def _records(self, arena):
        gmap = {id(p): g for g in self.param_groups for p in g["params"]}
        recs = []
        for i, ((_, p, o, n, _), (unit, ng, nb)) in enumerate(zip(arena.entries, arena.novo_layout)):
            end = arena.entries[i + 1][2] if i + 1 < len(arena.entries) else arena.total
            g = gmap.get(id(p))
            if g is None or p.grad is None:
                # not optimised (or no gradient, reference :104-110 skips it): identity update
                recs.append((o, end, unit, ng, nb, 0.0, 1.0, 1.0, 0.0, 1.0, 0.0))
                continue
            lr, wd = float(g["lr"]), float(g["weight_decay"])
            b1, b2 = float(g["betas"][0]), float(g["betas"][1])
            recs.append((o, end, unit, ng, nb, lr, 1 - lr * wd, b1, 1 - b1, b2, 1 - b2))
        return recs

    # ------------------------------------------------------------------ API
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._arenas is None or not all(a.intact() for a in self._arenas):
            self._collect_arenas()
        for a in self._arenas:
            for _, p, o, n, layout in a.entries:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("This optimizer does not support sparse gradients")
                gv = ParamArena.view_of(a.grad, o, p.shape, layout)
                if p.grad.data_ptr() != gv.data_ptr():
                    gv.copy_(p.grad)
                self.state[p]["step"] += 1
            cached = self._upload_table(a)
            recs = cached[0]
            ops.novograd_step(a.flat, a.grad, a.novo_ema_grad, a.shadow, cached[1], len(recs),
                              a.novo_sumsq, a.novo_ema_norm, a.novo_denom, self.eps,
                              self.unitwise_norm, ema=a.ema if self.ema_decay else None,
                              ema_decay=self.ema_decay)
            a.mark_fresh()
        return loss

    def _upload_table(self, a):
        """Persistent device table, contents refreshed in place (see SGD._upload_segments)."""
        recs = self._records(a)
        cached = self._tab_cache.get(id(a))
        if cached is None or cached[0] != recs:
            host = ops.novograd_table(recs).pin_memory()
            if cached is None or cached[1].numel() != host.numel():
                dev = host.to(a.device, non_blocking=True)
            else:
                dev = cached[1]
                dev.copy_(host, non_blocking=True)
            cached = self._tab_cache[id(a)] = (recs, dev, host)
        return cached

    def sync_hyperparams(self):
        for a in self._arenas or []:
            self._upload_table(a)

    def zero_grad(self, set_to_none=False):
        if self._arenas is None:
            return super().zero_grad(set_to_none)
        for a in self._arenas:
            if getattr(a, "_loose", False):
                for _, p, _, _, _ in a.entries:
                    p.grad = None if set_to_none else (p.grad.zero_() if p.grad is not None else None)
            else:
                a.zero_grad()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        loaded = {p: dict(s) for p, s in self.state.items()}
        self._collect_arenas()
        for a in self._arenas:
            for (_, p, o, n, lay), (unit, ng, nb) in zip(a.entries, a.novo_layout):
                s = loaded.get(p)
                if not s:
                    continue
                self.state[p]["step"] = int(s.get("step", 0))
                if s.get("ema_grad") is not None:
                    self.state[p]["ema_grad"].copy_(s["ema_grad"])
                en = s.get("ema_norm")
                if en is not None:
                    en = en.to(a.device, torch.float32)
                    a.novo_ema_norm[nb:nb + ng] = en.reshape(ng, -1)[:, 0] if en.numel() else en
