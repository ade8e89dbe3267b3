"""Config surface of the reference without hydra/omegaconf (both absent offline): the dataclass
schema of sota_imagenet/arg_parser.py:13-160, YAML experiment files in the reference's live
format (configs/hydra_exp/*.yaml: `defaults: [/base@_here_]` + overrides), CLI `key=value`
overrides, `${env:VAR}` interpolation and `_target_` instantiation (hydra.utils.call).  Reference
`_target_` paths are remapped to this package so the reference's own YAML files run unchanged."""
import copy
import importlib
import os
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import yaml

TARGET_REMAP = {
    "pytorch_tools.models.resnet50": "sota_imagenet_b200.models.resnet50",
    "pytorch_tools.models.resnet101": "sota_imagenet_b200.models.resnet101",
    "torchvision.models.resnet50": "sota_imagenet_b200.models.resnet50",
    "torch.optim._multi_tensor.SGD": "sota_imagenet_b200.optimizers.SGD",
    "torch.optim.SGD": "sota_imagenet_b200.optimizers.SGD",
    "pytorch_tools.losses.smooth.CrossEntropyLoss": "sota_imagenet_b200.losses.CrossEntropyLoss",
    "pytorch_tools.losses.CrossEntropyLoss": "sota_imagenet_b200.losses.CrossEntropyLoss",
    "pytorch_tools.fit_wrapper.callbacks.Callback": "sota_imagenet_b200.runner.Callback",
}
PREFIX_REMAP = {
    "sota_imagenet.angular_losses.": "sota_imagenet_b200.losses.",
    "src.angular_losses.": "sota_imagenet_b200.losses.",
    "sota_imagenet.model.CModel": "sota_imagenet_b200.cmodel.CModel",
    "src.model.CModel": "sota_imagenet_b200.cmodel.CModel",
    "sota_imagenet.optimizers.": "sota_imagenet_b200.optimizers.",
    "src.optimizers.": "sota_imagenet_b200.optimizers.",
    "sota_imagenet.callbacks.CutmixMixup": "sota_imagenet_b200.runner.CutmixMixup",
    "src.callbacks.CutmixMixup": "sota_imagenet_b200.runner.CutmixMixup",
    "pytorch_tools.fit_wrapper.callbacks.Cutmix": "sota_imagenet_b200.runner.Cutmix",
    "pytorch_tools.fit_wrapper.callbacks.Mixup": "sota_imagenet_b200.runner.Mixup",
}


@dataclass
class LoaderConfig:
    image_size: int = 224
    batch_size: int = 256
    workers: int = 6
    num_classes: int = 1000
    _is_train: bool = False
    root_data_dir: str = "${env:IMAGENET_DIR}"
    use_tfrecords: bool = False


@dataclass
class TrainLoaderConfig(LoaderConfig):
    _is_train: bool = True
    min_area: float = 0.08
    blur_prob: float = 0
    gray_prob: float = 0
    color_twist_prob: float = 0
    contrast_range: Tuple[float, float] = (0.7, 1.3)
    brightness_range: Tuple[float, float] = (0.7, 1.3)
    random_interpolation: bool = False
    re_prob: float = 0
    re_count: int = 3


@dataclass
class ValLoaderConfig(LoaderConfig):
    batch_size: int = 250
    full_crop: bool = False


@dataclass
class DataStage:
    start: int = 0
    end: int = 90
    lr: Optional[Tuple[float, float]] = None
    lr_mode: Optional[str] = "linear"
    extra_args: Optional[Dict] = None


@dataclass
class RunnerConfig:
    stages: List = field(default_factory=lambda: [DataStage(lr=(0.1, 0))])
    resume: Optional[str] = None
    load_start_epoch: bool = True
    start_epoch: int = 0
    accumulate_steps: int = 1
    ema_decay: float = 0
    fp16: bool = True          # kept for schema compatibility; the kernels compute in bf16
    extra_callbacks: List = field(default_factory=list)
    evaluate: bool = False


@dataclass
class LoggerConfig:
    exp_name: str = "test_run"
    dir: str = "logs"
    print_model: bool = False
    histogram: bool = False
    save_optim: bool = False


@dataclass
class StrictConfig:
    loader: TrainLoaderConfig = field(default_factory=TrainLoaderConfig)
    val_loader: ValLoaderConfig = field(default_factory=ValLoaderConfig)
    model: Dict[str, Any] = field(default_factory=lambda: dict(_target_="pytorch_tools.models.resnet50"))
    weight_standardization: bool = False
    filter_from_wd: Optional[List[str]] = None
    bn_momentum: float = 0.1
    init_gamma: Optional[float] = 1.72
    optim: Dict[str, Any] = field(default_factory=lambda: dict(_target_="torch.optim._multi_tensor.SGD", lr=0, weight_decay=1e-4))
    criterion: Dict[str, Any] = field(default_factory=lambda: dict(_target_="pytorch_tools.losses.smooth.CrossEntropyLoss"))
    run: RunnerConfig = field(default_factory=RunnerConfig)
    log: LoggerConfig = field(default_factory=LoggerConfig)
    debug: bool = False
    random_seed: Optional[int] = None
    world_size: int = 1
    local_rank: int = 0
    distributed: bool = False
    is_master: bool = True
    # extensions of this repo (not in the reference schema)
    sync_bn: bool = True
    steps_per_epoch: Optional[int] = None


_FLOAT_RE = __import__("re").compile(r"^[+-]?(\d+\.?\d*|\.\d+)[eE][+-]?\d+$")


def _coerce(node):
    """PyYAML (YAML 1.1) reads `3e-5` as a string; OmegaConf reads a float.  Follow OmegaConf."""
    if isinstance(node, dict):
        return {k: _coerce(v) for k, v in node.items()}
    if isinstance(node, list):
        return [_coerce(v) for v in node]
    if isinstance(node, str) and _FLOAT_RE.match(node):
        return float(node)
    return node


def _set_path(obj, dotted, value):
    keys = dotted.split(".")
    for k in keys[:-1]:
        obj = obj[k] if isinstance(obj, dict) else getattr(obj, k)
    if isinstance(obj, dict):
        obj[keys[-1]] = value
    else:
        if not hasattr(obj, keys[-1]):
            raise KeyError("unknown config key: %s" % dotted)
        setattr(obj, keys[-1], value)


def _merge(obj, overrides):
    for k, v in overrides.items():
        if k == "defaults":
            continue
        cur = obj.get(k) if isinstance(obj, dict) else getattr(obj, k, None)
        if isinstance(obj, dict):
            if isinstance(v, dict) and isinstance(cur, dict):
                _merge(cur, v)
            else:
                obj[k] = v
        else:
            if not hasattr(obj, k):
                raise KeyError("unknown config key: %s" % k)
            if isinstance(v, dict) and cur is not None and not isinstance(cur, (int, float, str, list, tuple)):
                _merge(cur, v)
            else:
                setattr(obj, k, v)


def load_config(path=None, overrides=()):
    cfg = StrictConfig()
    cfg.world_size = int(os.environ.get("WORLD_SIZE", "1"))
    cfg.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if path is not None:
        with open(path) as f:
            _merge(cfg, _coerce(yaml.safe_load(f) or {}))
    for ov in overrides:
        key, _, val = ov.partition("=")
        _set_path(cfg, key.lstrip("+"), _coerce(yaml.safe_load(val)))
    cfg.run.stages = [s if isinstance(s, DataStage) else DataStage(**s) for s in cfg.run.stages]
    cfg.distributed = cfg.world_size > 1
    cfg.is_master = cfg.local_rank == 0
    return cfg


def resolve_target(name):
    name = TARGET_REMAP.get(name, name)
    for prefix, repl in PREFIX_REMAP.items():
        if name.startswith(prefix):
            name = repl + name[len(prefix):] if prefix.endswith(".") else repl
            break
    module, _, attr = name.rpartition(".")
    return getattr(importlib.import_module(module), attr)


def call(node, *args, **extra):
    """hydra.utils.call: instantiate `_target_` with the remaining keys as kwargs (recursively)."""
    node = copy.deepcopy(dict(node))
    target = resolve_target(node.pop("_target_"))
    kwargs = {k: (call(v) if isinstance(v, dict) and "_target_" in v else v) for k, v in node.items()}
    kwargs.update(extra)
    return target(*args, **kwargs)
