"""Minimal step driver with the semantics of the external pytorch_tools Runner that the reference
wires in train.py:129-173: fwd -> loss -> backward -> (every
accumulate_steps) optimizer.step + zero_grad, per-batch LR from PhasesScheduler, loss / Acc@1 /
Acc@5 meters, checkpoints {state_dict, epoch, optimizer}.  Loss stays on the device; it is read
back every `log_every` steps only."""
import math
import os

import torch


class Callback:
    """No-op callback (pytorch_tools.fit_wrapper.callbacks.Callback stand-in)."""

    def set_state(self, state):
        self.state = state

    def on_begin(self): pass
    def on_epoch_begin(self): pass
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
                 log_every=50, logger=None):
        self.state = RunnerState(model, optimizer, criterion)
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
        pending = []
        for i, batch in enumerate(loader):
            if steps is not None and i >= steps:
                break
            st.step, st.input = i, batch
            data, target = batch
            self._cb("on_batch_begin")
            if train:
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
            pending.append((loss.detach(), out.detach(), target, data.shape[0]))
            if len(pending) >= self.log_every:
                self._drain(pending)
            self._cb("on_batch_end")
        self._drain(pending)
        return st.loss_meter.avg, {k: m.avg for k, m in st.metric_meters.items()}

    def _drain(self, pending):
        """One device->host sync for a window of steps (the reference reads every step)."""
        st = self.state
        for loss, out, target, n in pending:
            a1, a5 = accuracy(out, target)
            st.loss_meter.update(loss.item(), n)
            st.metric_meters["Acc@1"].update(a1.item(), n)
            st.metric_meters["Acc@5"].update(a5.item(), n)
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
