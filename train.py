#!/usr/bin/env python
"""Training driver with the wiring of the reference train.py (main(), :22-185) on this package:
    torchrun --nproc-per-node N train.py configs/r50_baseline.yaml [key=value ...]
Config files use the reference's live YAML schema; its `_target_` entries resolve to the fused
sm_100a model / criterion / optimizer (sota_imagenet_b200.config.TARGET_REMAP).  Data is the
synthetic GPU pipeline (sota_imagenet_b200.data) unless cfg.loader.root_data_dir / $IMAGENET_DIR holds the
reference's layout (class folders or TFRecord shards + DALI indexes): then data.RecordLoader reads it."""
import os
import sys
import time

import torch

from sota_imagenet_b200 import config as cfglib
from sota_imagenet_b200 import data, parallel, runner as rt


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    path = argv[0] if argv and not ("=" in argv[0]) else None
    cfg = cfglib.load_config(path, [a for a in argv if "=" in a])
    start_time = time.time()
    log = (lambda m: print(time.strftime("[%m-%d %H:%M:%S] - ") + str(m), flush=True)) if cfg.is_master else (lambda m: None)
    if cfg.random_seed is not None:
        torch.manual_seed(cfg.random_seed)
    torch.cuda.set_device(cfg.local_rank)
    if cfg.distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group(backend="nccl", init_method="env://",
                                             world_size=cfg.world_size)
    rank = torch.distributed.get_rank() if cfg.distributed else 0

    log("Loading model")
    model = cfglib.call(cfg.model)                                    # train.py:64
    if cfg.weight_standardization and hasattr(model, "enable_weight_standardization"):
        model.enable_weight_standardization()                         # train.py:66-67
    if cfg.init_gamma is not None:                                    # train.py:70-71
        rt.initialize(model, cfg.init_gamma)
    model = model.cuda()                                              # train.py:73
    if hasattr(model, "ensure_arena"):
        # build the parameter arena in registration order BEFORE the optimizer exists (or loads a
        # checkpoint): the data-parallel bucket plan relies on that order
        model.ensure_arena()
    rt.patch_bn_mom(model, cfg.bn_momentum)                           # train.py:76
    criterion = cfglib.call(cfg.criterion).cuda()                     # train.py:81
    if cfg.filter_from_wd is not None:                                # train.py:83-86
        opt_params = rt.filter_from_weight_decay(model, skip_list=cfg.filter_from_wd)
    else:
        opt_params = [{"params": list(model.parameters())}]
    opt_params[0]["params"].extend(list(criterion.parameters()))      # train.py:89
    # the fused optimizers keep the EMA copy of the weights themselves (one extra stream of the update
    # kernel); any other optimizer target gets the reference's stand-alone ModelEma callback below
    opt_cls = cfglib.resolve_target(cfg.optim["_target_"])
    fused_ema = cfg.run.ema_decay > 0 and "ema_decay" in __import__("inspect").signature(opt_cls.__init__).parameters
    optimizer = cfglib.call(cfg.optim, opt_params, **({"ema_decay": cfg.run.ema_decay} if fused_ema else {}))   # train.py:92
    log("Model params: %.2fM" % (sum(p.numel() for p in model.parameters()) / 1e6))

    if cfg.run.resume:                                                # train.py:98-109
        ckpt = torch.load(cfg.run.resume, map_location="cuda:%d" % cfg.local_rank, weights_only=False)
        model.load_state_dict(ckpt["state_dict"], strict=False)
        if cfg.run.load_start_epoch:
            cfg.run.start_epoch = ckpt["epoch"]
        try:
            optimizer.load_state_dict(ckpt["optimizer"])
        except Exception:
            log("Failed to load state dict into optimizer")

    if cfg.distributed:                                               # train.py:113-114
        model = parallel.DataParallel(model, sync_bn=cfg.sync_bn)

    lr_stages = [dict(ep=(s.start, s.end), lr=s.lr, mode=s.lr_mode) for s in cfg.run.stages if s.lr is not None]
    log("Learning rate stages: %s" % lr_stages)
    callbacks = [rt.PhasesScheduler(lr_stages),
                 rt.CheckpointSaver(os.getcwd(), "model.chpn", cfg.log.save_optim) if cfg.is_master else None]
    if cfg.run.ema_decay > 0:                                         # train.py:112,138: after CheckpointSaver
        callbacks.append(rt.ModelEma(model, cfg.run.ema_decay, optimizer if fused_ema else None))
    callbacks += [cfglib.call(c) for c in cfg.run.extra_callbacks]
    run = rt.Runner(model, optimizer, criterion, callbacks=callbacks,
                    accumulate_steps=cfg.run.accumulate_steps, logger=log)
    dm = data.DataManager(cfg, rank=rank, world_size=cfg.world_size)  # train.py:156

    if cfg.run.evaluate:
        dm.set_stage(0)
        return run.evaluate(dm.val_loader)
    for idx in range(len(dm)):                                        # train.py:164-173
        dm.set_stage(idx)
        steps = 10 if cfg.debug else cfg.steps_per_epoch
        run.fit(dm.loader, steps_per_epoch=steps, val_loader=dm.val_loader,
                val_steps=20 if cfg.debug else cfg.steps_per_epoch, epochs=dm.end_epoch,
                start_epoch=dm.start_epoch)
    m = (time.time() - start_time) / 60
    log("Total time: %dh %.1fm" % (int(m / 60), m % 60))
    if cfg.is_master:
        net = model.module if hasattr(model, "module") else model
        torch.save(net.state_dict(), "model_last.chpn")              # train.py:183-184
    if cfg.distributed:
        torch.distributed.destroy_process_group()
    return run.state.train_loss


if __name__ == "__main__":
    main()
