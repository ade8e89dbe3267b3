"""Minimal step driver with the semantics of the external pytorch_tools Runner that the reference
wires in train.py:129-173: fwd -> loss -> backward -> (every
accumulate_steps) optimizer.step + zero_grad, per-batch LR from PhasesScheduler, loss / Acc@1 /
Acc@5 meters, checkpoints {state_dict, epoch, optimizer}.  Loss stays on the device; it is read
back every `log_every` steps only."""
import math
import os

import numpy as np
import torch

from . import ops


class Callback:
    """No-op callback (pytorch_tools.fit_wrapper.callbacks.Callback stand-in)."""

    def set_state(self, state):
        self.state = state

    def on_begin(self): pass
    def on_epoch_begin(self): pass
    def on_loader_begin(self): pass
    def on_loader_end(self): pass
    def on_batch_begin(self): pass
    def on_after_backward(self): pass
    def on_batch_end(self): pass
    def on_epoch_end(self): pass
    def on_end(self): pass


class PhasesScheduler(Callback):
    """Per-batch LR: for the stage covering the current epoch, pct = (epoch - start + step/steps)
    / (end - start); linear: a + (b-a)*pct, cos: b + (a-b)/2 * (1 + cos(pi*pct))."""

    def __init__(self, lr_stages):
        self.stages = lr_stages

    @staticmethod
    def lr_at(stages, epoch_float):
        for st in stages:
            s, e = st["ep"]
            if s <= epoch_float < e or (st is stages[-1] and epoch_float >= e):
                a, b = st["lr"]
                pct = min(max((epoch_float - s) / (e - s), 0.0), 1.0)
                if st.get("mode", "linear") == "cos":
                    return b + (a - b) / 2 * (1 + math.cos(math.pi * pct))
                return a + (b - a) * pct
        return None

    def on_batch_begin(self):
        st = self.state
        lr = self.lr_at(self.stages, st.epoch + st.step / max(st.epoch_size, 1))
        if lr is not None:
            for g in st.optimizer.param_groups:
                g["lr"] = lr


class AverageMeter:
    def __init__(self):
        self.sum, self.n = 0.0, 0

    def update(self, v, n=1):
        self.sum += float(v) * n
        self.n += n

    @property
    def avg(self):
        return self.sum / max(self.n, 1)


def accuracy(output, target, topk=(1, 5)):
    if target.dim() == 2:
        target = target.argmax(1)
    _, pred = output.float().topk(max(topk), 1)
    correct = pred.eq(target.view(-1, 1))
    return [correct[:, :k].any(1).float().mean() * 100 for k in topk]


class GraphStep:
    """One training step (forward, loss, backward, optimizer step, zero_grad, Acc@1/5) captured in a
    CUDA graph per input signature (shapes / dtypes of the batch: a progressive-resize stage or a
    switch between index and soft targets captures a new graph) and replayed afterwards: ~300 kernel
    launches become one graph launch, which is what keeps small-image stages (128 px: 7 ms of GPU
    work per step) from being bound by the host.  Replaces the eager body of the reference's
    training loop (pytorch_tools Runner._run_one_epoch, wired at train.py:121-131).

    * inputs are copied into static buffers (one D2D copy of the batch), outputs (loss, logits,
      Acc@1, Acc@5) live in static tensors that the next replay overwrites;
    * optimizer hyper-parameters are read by the update kernel from a persistent device table;
      `optimizer.sync_hyperparams()` refreshes it before every replay, so a per-batch LR schedule
      (PhasesScheduler) takes effect across replays;
    * anything that cannot be captured (an optimizer without sync_hyperparams, a capture error)
      falls back to the same step launched eagerly -- same kernels, same results."""

    WARMUP = 2

    def __init__(self, model, criterion, optimizer, enabled=True, loss_scale=1.0, after_backward=None):
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.after_backward = after_backward     # e.g. gradient averaging of parameters outside DataParallel
        self.enabled = enabled and hasattr(optimizer, "sync_hyperparams") and torch.cuda.is_available()
        self.loss_scale = loss_scale
        self.graphs = {}         # signature -> (graph, static_x, static_y, outputs) or None (eager)
        self.replays = 0

    def _eager(self, x, y):
        out = self.model(x)
        loss = self.criterion(out, y)
        (loss * self.loss_scale if self.loss_scale != 1.0 else loss).backward()
        if self.after_backward is not None:
            self.after_backward()
        self.optimizer.step()
        self.optimizer.zero_grad()
        a1, a5 = accuracy(out.detach(), y)
        return loss.detach(), out.detach(), a1, a5

    @staticmethod
    def _sig(x, y):
        return (tuple(x.shape), x.dtype, tuple(x.stride()), tuple(y.shape), y.dtype)

    def _capture(self, x, y):
        sx, sy = x.clone(), y.clone()
        self.optimizer.sync_hyperparams()        # no table upload may be captured (it would be replayed)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            outs = self._eager(sx, sy)           # recorded, not executed: the replay below is the step
        return graph, sx, sy, outs

    def __call__(self, x, y):
        """-> (loss, logits, acc1, acc5): device tensors, valid until the next call."""
        if not self.enabled:
            return self._eager(x, y)
        sig = self._sig(x, y)
        entry = self.graphs.get(sig, 0)
        if isinstance(entry, int):
            if entry < self.WARMUP:              # the first steps of a signature run eagerly (lazy
                self.graphs[sig] = entry + 1     # initialisation, arena build, allocator warm-up)
                return self._eager(x, y)
            try:
                entry = self._capture(x, y)
            except Exception as e:               # capture is an optimisation
                import warnings
                warnings.warn("sota_imagenet_b200: CUDA graph capture of the training step failed (%s); "
                              "running eagerly" % repr(e)[:200])
                torch.cuda.synchronize()
                entry = None
            self.graphs[sig] = entry
        if entry is None:
            return self._eager(x, y)
        graph, sx, sy, outs = entry
        sx.copy_(x, non_blocking=True)
        sy.copy_(y, non_blocking=True)
        self.optimizer.sync_hyperparams()
        graph.replay()
        self.replays += 1
        return outs


class RunnerState:
    def __init__(self, model, optimizer, criterion):
        self.model, self.optimizer, self.criterion = model, optimizer, criterion
        self.epoch, self.step, self.epoch_size, self.global_sample_step = 0, 0, 1, 0
        self.input = self.output = self.loss = None
        self.is_train = True
        self.loss_meter = AverageMeter()
        self.metric_meters = {"Acc@1": AverageMeter(), "Acc@5": AverageMeter()}
        self.train_loss = self.val_loss = None
        self.val_metrics = None


class Runner:
    def __init__(self, model, optimizer, criterion, callbacks=(), use_fp16=True, accumulate_steps=1,
                 log_every=50, logger=None, use_graph=None):
        self.state = RunnerState(model, optimizer, criterion)
        # CUDA-graph replay of the whole training step (GraphStep); off with gradient accumulation
        # (the optimizer step is then not part of every batch) and for callbacks that hook between
        # backward and the optimizer step; SIB_GRAPH=0 forces eager launches
        if use_graph is None:
            use_graph = os.environ.get("SIB_GRAPH", "1") != "0"
        hooks_backward = any(type(c).on_after_backward is not Callback.on_after_backward
                             for c in callbacks if c is not None)
        self.graph_step = GraphStep(model, criterion, optimizer,
                                    enabled=use_graph and accumulate_steps == 1 and not hooks_backward)
        self.callbacks = [c for c in callbacks if c is not None]
        for c in self.callbacks:
            c.set_state(self.state)
        self.accumulate_steps = accumulate_steps
        self.log_every = log_every
        self.log = logger or (lambda msg: None)

    def _cb(self, name):
        for c in self.callbacks:
            getattr(c, name)()

    def _run_loader(self, loader, steps=None, train=True):
        st = self.state
        st.is_train = train
        st.model.train(train)
        st.loss_meter = AverageMeter()
        st.metric_meters = {"Acc@1": AverageMeter(), "Acc@5": AverageMeter()}
        st.epoch_size = steps or len(loader)
        self._cb("on_loader_begin")
        pending = []
        for i, batch in enumerate(loader):
            if steps is not None and i >= steps:
                break
            st.step, st.input = i, batch
            self._cb("on_batch_begin")
            data, target = st.input          # callbacks (CutmixMixup) may replace the batch
            accs = None
            if train and self.graph_step.enabled:
                loss, out, a1, a5 = self.graph_step(data, target)
                # static tensors, overwritten by the next replay: keep three scalars, not the logits
                accs = torch.stack([loss.float(), a1, a5])
            elif train:
                out = st.model(data)
                loss = st.criterion(out, target)
                (loss / self.accumulate_steps).backward()
                self._cb("on_after_backward")
                if (i + 1) % self.accumulate_steps == 0:
                    st.optimizer.step()
                    st.optimizer.zero_grad()
            else:
                with torch.no_grad():
                    out = st.model(data)
                    loss = st.criterion(out, target)
            st.output, st.loss = out, loss
            st.global_sample_step += data.shape[0]
            if accs is not None:
                pending.append((accs, None, None, data.shape[0]))
            else:
                pending.append((loss.detach(), out.detach(), target, data.shape[0]))
            if len(pending) >= self.log_every:
                self._drain(pending)
            self._cb("on_batch_end")
        self._drain(pending)
        self._cb("on_loader_end")
        self._reduce_meters()
        return st.loss_meter.avg, {k: m.avg for k, m in st.metric_meters.items()}

    def _reduce_meters(self):
        """Sum the loss / metric meters over the ranks (each rank sees 1 / world of the data; the
        reference Runner reduces them too), so that rank 0 logs whole-dataset numbers."""
        if not (torch.distributed.is_available() and torch.distributed.is_initialized()) or \
                torch.distributed.get_world_size() == 1:
            return
        st = self.state
        meters = [st.loss_meter] + [st.metric_meters[k] for k in sorted(st.metric_meters)]
        dev = "cuda" if torch.distributed.get_backend() == "nccl" else "cpu"
        t = torch.tensor([v for m in meters for v in (m.sum, float(m.n))], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t)
        for i, m in enumerate(meters):
            m.sum, m.n = float(t[2 * i]), int(t[2 * i + 1])

    def _drain(self, pending):
        """One device->host sync for a window of steps (the reference reads every step)."""
        st = self.state
        for loss, out, target, n in pending:
            if out is None:                     # graph-replayed step: (loss, Acc@1, Acc@5) already reduced
                lv, a1, a5 = loss.tolist()
            else:
                a1, a5 = (v.item() for v in accuracy(out, target))
                lv = loss.item()
            st.loss_meter.update(lv, n)
            st.metric_meters["Acc@1"].update(a1, n)
            st.metric_meters["Acc@5"].update(a5, n)
        pending.clear()

    def fit(self, loader, steps_per_epoch=None, val_loader=None, val_steps=None, epochs=1, start_epoch=0):
        self._cb("on_begin")
        for epoch in range(start_epoch, epochs):
            self.state.epoch = epoch
            self._cb("on_epoch_begin")
            tl, tm = self._run_loader(loader, steps_per_epoch, train=True)
            self.state.train_loss = tl
            msg = "Epoch %d | Train loss: %.4f | Acc@1: %.3f | Acc@5: %.3f" % (epoch, tl, tm["Acc@1"], tm["Acc@5"])
            if val_loader is not None:
                vl, vm = self.evaluate(val_loader, val_steps)
                msg += " || Val loss: %.4f | Acc@1: %.3f | Acc@5: %.3f" % (vl, vm["Acc@1"], vm["Acc@5"])
            self.log(msg)
            self._cb("on_epoch_end")
        self._cb("on_end")

    def evaluate(self, loader, steps=None):
        vl, vm = self._run_loader(loader, steps, train=False)
        self.state.val_loss, self.state.val_metrics = vl, vm
        return vl, vm


class _BatchMix(Callback):
    """Shared state of pt_clb.Mixup / pt_clb.Cutmix (pytorch_tools, absent: restated; the
    reference combines them in sota_imagenet/callbacks.py:232-247).  Random draws use the same
    host generators in the same order as the originals (np.random for the gates and the box
    centre, torch CPU RNG for the permutation and the Beta sample); the data path is two
    kernels on the resident batch (ops.mix_batch, ops.mix_targets)."""

    def __init__(self, alpha, num_classes, prob=0.5):
        self.tb = torch.distributions.Beta(alpha, alpha)
        self.num_classes = num_classes
        self.prob = prob
        self.prev_input = None

    def _one_hot(self, data, target):
        if target.dim() == 1:
            return ops.one_hot(target.to(torch.int64), self.num_classes)
        return target

    def _gate(self):
        is_train = self.state.is_train if getattr(self, "state", None) is not None else True
        return is_train and not (np.random.rand() > self.prob)

    def _previous(self, data, target_one_hot):
        prev = (data, target_one_hot) if self.prev_input is None else self.prev_input
        self.prev_input = data.clone(), target_one_hot.clone()
        return prev

    @staticmethod
    def _perm(n, device):
        return torch.randperm(n).to(device=device, dtype=torch.int32)


class Mixup(_BatchMix):
    """out = c*data + (1-c)*prev[perm], targets likewise, c ~ Beta(alpha, alpha)."""

    def on_batch_begin(self):
        self.state.input = self.mixup(*self.state.input)

    @torch.no_grad()
    def mixup(self, data, target):
        target_one_hot = self._one_hot(data, target)
        if not self._gate():
            return data, target_one_hot
        prev_data, prev_target = self._previous(data, target_one_hot)
        perm = self._perm(data.size(0), data.device)
        c = np.float32(self.tb.sample().item())
        omc = np.float32(1.0) - c
        md = ops.mix_batch(data, prev_data, perm, 0, c, omc)
        mt = ops.mix_targets(target_one_hot, prev_target, perm, c, omc)
        return md, mt


class Cutmix(_BatchMix):
    """A random box of the permuted previous batch is pasted into the batch; the targets mix
    with the box's true area fraction."""

    def on_batch_begin(self):
        self.state.input = self.cutmix(*self.state.input)

    @torch.no_grad()
    def cutmix(self, data, target):
        target_one_hot = self._one_hot(data, target)
        if not self._gate():
            return data, target_one_hot
        prev_data, prev_target = self._previous(data, target_one_hot)
        # the previous batch can have another size (progressive resizing): use the common extent
        H, W = min(data.size(2), prev_data.size(2)), min(data.size(3), prev_data.size(3))
        perm = self._perm(data.size(0), data.device)
        lam = float(self.tb.sample())
        lam = min(lam, 1 - lam)
        bbh1, bbw1, bbh2, bbw2 = self.rand_bbox(H, W, lam)
        lam = (bbh2 - bbh1) * (bbw2 - bbw1) / (H * W)      # the clipped box's real share
        md = ops.mix_batch(data, prev_data, perm, 1, box=(bbh1, bbw1, bbh2, bbw2))
        mt = ops.mix_targets(target_one_hot, prev_target, perm, np.float32(1 - lam), np.float32(lam))
        return md, mt

    @staticmethod
    def rand_bbox(H, W, lam):
        """box with area close to lam*H*W around a uniform centre, clipped to the image"""
        cut_rat = np.sqrt(lam)
        cut_h, cut_w = int(H * cut_rat), int(W * cut_rat)
        ch, cw = np.random.randint(H), np.random.randint(W)
        return (int(np.clip(ch - cut_h // 2, 0, H)), int(np.clip(cw - cut_w // 2, 0, W)),
                int(np.clip(ch + cut_h // 2, 0, H)), int(np.clip(cw + cut_w // 2, 0, W)))


class CutmixMixup(Cutmix, Mixup):
    """Reference sota_imagenet/callbacks.py:232-247: CutMix or Mixup, a coin flip per batch."""

    def __init__(self, cutmix_alpha, mixup_alpha, prob=0.5, num_classes=1000):
        self.cutmix_tb = torch.distributions.Beta(cutmix_alpha, cutmix_alpha)
        self.mixup_tb = torch.distributions.Beta(mixup_alpha, mixup_alpha)
        self.prob = prob
        self.prev_input = None
        self.num_classes = num_classes    # the reference leaves it unset (DALI targets are one-hot)

    def on_batch_begin(self):
        if np.random.rand() > 0.5:
            self.tb = self.cutmix_tb
            self.state.input = self.cutmix(*self.state.input)
        else:
            self.tb = self.mixup_tb
            self.state.input = self.mixup(*self.state.input)


class ModelEma(Callback):
    """pt_clb.ModelEma(model, decay) as the reference wires it (train.py:112,138: created after
    .cuda(), placed AFTER CheckpointSaver): an exponential moving average of the weights that is
    swapped into the model when validation starts, so validation and the epoch's checkpoint
    (CheckpointSaver.on_epoch_end runs first) see the EMA weights, and swapped back out at epoch
    end before training continues.  With a fused optimizer (`optimizer` given: SGD / MyNovograd
    keep `arena.ema` up to date inside their update kernel) this callback only swaps; otherwise it
    also maintains the average itself after every batch.  The average covers the parameters
    (BatchNorm running statistics stay the live ones)."""

    def __init__(self, model, decay=0.9999, optimizer=None):
        self.model = model.module if hasattr(model, "module") else model
        self.decay, self.optimizer = decay, optimizer
        self.swapped = False
        self._own = None

    def _arena(self):
        return self.model.ensure_arena()

    def on_batch_end(self):
        if self.optimizer is not None or not self.state.is_train:
            return
        a = self._arena()
        if self._own is None:
            self._own = a.flat.clone()
        self._own.lerp_(a.flat, 1.0 - self.decay)

    def _ema_buffer(self):
        a = self._arena()
        return a.ema if self.optimizer is not None else self._own

    def _swap(self):
        a, ema = self._arena(), self._ema_buffer()
        if ema is None:
            return False
        tmp = a.flat.clone()
        a.flat.copy_(ema)
        ema.copy_(tmp)
        a.refresh_shadow(force=True)
        return True

    def on_loader_begin(self):
        if not self.state.is_train and not self.swapped:
            self.swapped = self._swap()

    def on_epoch_end(self):
        if self.swapped:
            self._swap()
            self.swapped = False

    def state_dict(self):
        """EMA weights keyed like model.state_dict()."""
        from .arena import ParamArena
        a, ema = self._arena(), self._ema_buffer()
        if ema is None:
            return {}
        names = {id(p): n for n, p in self.model.named_parameters()}
        return {names[id(p)]: ParamArena.view_of(ema, o, p.shape, layout).clone()
                for _, p, o, n, layout in a.entries if id(p) in names}


def initialize(model, gamma=1.72):
    """pt.utils.misc.initialize(model, gamma) (reference train.py:70-71, absent package; SURVEY
    App. C): variance-preserving init with gain `gamma` -- conv / linear weights ~ N(0, gamma^2 /
    fan_in)  (gamma = 1.72 compensates the variance a (leaky-)ReLU removes), BatchNorm weight 1 /
    bias 0, linear bias 0.  Parity unpinned (the reference's source is absent)."""
    from .modules import BatchNorm2d, Conv2d, Linear, StemConv
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, (Conv2d, StemConv, Linear, torch.nn.Conv2d, torch.nn.Linear)):
                w = m.weight
                fan_in = w[0].numel()
                w.normal_(0.0, gamma / math.sqrt(fan_in))
                if getattr(m, "bias", None) is not None:
                    m.bias.zero_()
            elif isinstance(m, (BatchNorm2d, torch.nn.modules.batchnorm._BatchNorm)):
                # (a zero-initialised last BN of a residual block is kept)
                if m.weight is not None and float(m.weight.abs().sum()) != 0.0:
                    m.weight.fill_(1.0)
                if m.bias is not None:
                    m.bias.zero_()


class CheckpointSaver(Callback):
    """{state_dict, epoch, optimizer} like pt_clb.CheckpointSaver (reference train.py:134, :101-106)."""

    def __init__(self, save_dir, save_name="model.chpn", include_optimizer=False):
        self.path = os.path.join(save_dir, save_name)
        self.include_optimizer = include_optimizer

    def on_epoch_end(self):
        st = self.state
        model = st.model.module if hasattr(st.model, "module") else st.model
        ckpt = {"state_dict": model.state_dict(), "epoch": st.epoch + 1}
        if self.include_optimizer:
            ckpt["optimizer"] = st.optimizer.state_dict()
        torch.save(ckpt, self.path)


def filter_from_weight_decay(model, skip_list=("bias", "bn", "gain")):
    """pt.utils.misc.filter_from_weight_decay (reference train.py:83-84): two param groups, the
    second with weight_decay=0 for parameter names matching `skip_list`."""
    decay, no_decay = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (no_decay if any(s in name for s in skip_list) else decay).append(p)
    return [{"params": decay}, {"params": no_decay, "weight_decay": 0.0}]


def patch_bn_mom(model, momentum):
    from .modules import BatchNorm2d
    for m in model.modules():
        if isinstance(m, (BatchNorm2d, torch.nn.modules.batchnorm._BatchNorm)):
            m.momentum = momentum
