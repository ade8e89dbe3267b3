"""CModel: the contract of the reference's Caffe-style list-of-layers constructor
(sota_imagenet/model.py:1098-1226) over the fused sm_100a modules of this package.

A model is a list of layer entries `{module, args, kwargs, repeat, inputs, tag}`:
  * `module` (and every string in `args` / `kwargs`) is a Python expression evaluated in THIS
    module's namespace (`Conv2d`, `Bottleneck`, `nn.ReLU`, `"'relu'"` for a literal string ...);
  * `repeat > 1` stacks that many fresh copies in an `nn.Sequential`;
  * `inputs` names the producers by `tag` (`"_prev_"` = the previous layer), so U-Net / FPN style
    graphs are expressible; a layer with several inputs receives them as positional arguments;
  * `extra_kwargs[name]` supplies defaults for every layer whose `module` string equals `name`
    (per-layer kwargs win, nested dicts merge).
Attributes the reference exposes are kept (`saved_layers_idx`, per-layer `input_indexes` / `idx`);
the graph executor additionally releases a retained output after its last consumer ran.
"""
from copy import deepcopy
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Union

import torch  # noqa: F401  (available to eval'ed config strings)
import torch.nn as nn

from .modules import (BatchNorm2d, Bottleneck, Concat, Conv2d, GlobalAvgPool, Linear,  # noqa: F401
                      MaxPool3x3s2, StemConv)
from .losses import SphereLinearLayer  # noqa: F401

PREVIOUS = "_prev_"


@dataclass
class ModuleStructure:
    module: Union[str, nn.Module]
    args: List = field(default_factory=tuple)
    kwargs: Dict[str, Any] = field(default_factory=dict)
    repeat: int = 1
    inputs: List[str] = field(default_factory=lambda: [PREVIOUS])
    tag: Optional[str] = None


def listify(x):
    if x is None:
        return []
    return list(x) if isinstance(x, (list, tuple)) else [x]


def _update_dict(to_dict, from_dict):
    """Merge `from_dict` into `to_dict` in place; nested mappings merge key by key instead of being
    replaced (what `dict.update` would do).  Returns `to_dict`."""
    for key, value in from_dict.items():
        nested = hasattr(value, "keys") and key in to_dict
        if nested:
            _update_dict(to_dict[key], value)
        else:
            to_dict[key] = value
    return to_dict


def _resolve(value):
    """Config strings are expressions of this namespace; everything else is taken literally."""
    return eval(value) if isinstance(value, str) else value   # noqa: S307 (the reference's contract)


def _build_layer(spec):
    """One ModuleStructure -> nn.Module (`repeat` fresh instances in a Sequential)."""
    factory = _resolve(spec.module)
    args = [_resolve(a) for a in listify(spec.args)]
    kwargs = {name: _resolve(v) for name, v in spec.kwargs.items()}
    spec.module, spec.args, spec.kwargs = factory, args, kwargs
    copies = [factory(*args, **kwargs) for _ in range(max(int(spec.repeat), 1))]
    return copies[0] if len(copies) == 1 else nn.Sequential(*copies)


class CModel(nn.Sequential):
    def __init__(self, layer_config, extra_kwargs=None):
        specs = [entry if isinstance(entry, ModuleStructure) else ModuleStructure(**entry)
                 for entry in layer_config]
        if extra_kwargs is not None:
            self._update_config_with_extra_params(specs, extra_kwargs)
        layers, retained = self._parse_config(specs)
        super().__init__(*layers)
        self.saved_layers_idx = retained
        # index of the last layer that reads each retained output (-> release point)
        self._last_reader = {}
        for layer in layers:
            for src in layer.input_indexes:
                if src != -1:
                    self._last_reader[src] = layer.idx
        if retained:                       # a plain chain keeps nn.Sequential's own forward
            self.forward = self.custom_forward

    @staticmethod
    def _update_config_with_extra_params(specs, extra_kwargs):
        for spec in specs:
            defaults = extra_kwargs.get(spec.module) if isinstance(spec.module, str) else None
            if defaults is not None:
                spec.kwargs = _update_dict(deepcopy(dict(defaults)), spec.kwargs)

    @staticmethod
    def _parse_config(specs):
        """-> (modules annotated with `input_indexes` / `idx`, indices whose outputs are re-used)."""
        position = {PREVIOUS: -1}
        for idx, spec in enumerate(specs):
            if spec.tag is not None:
                position[spec.tag] = idx
        layers, retained = [], []
        for idx, spec in enumerate(specs):
            module = _build_layer(spec)
            unknown = [name for name in spec.inputs if name not in position]
            if unknown:
                raise KeyError("layer %d reads unknown tag(s) %s" % (idx, unknown))
            module.input_indexes = [position[name] for name in spec.inputs]
            module.idx = idx
            retained += [src for src in module.input_indexes if src != -1]
            layers.append(module)
        return layers, retained

    def custom_forward(self, x):
        kept = {}
        for layer in self.children():
            feeds = [x if src == -1 else kept[src] for src in layer.input_indexes]
            x = layer(*feeds)
            if layer.idx in self._last_reader:
                kept[layer.idx] = x
            for src in layer.input_indexes:            # last consumer done: drop the reference
                if src != -1 and self._last_reader[src] == layer.idx:
                    kept.pop(src, None)
        return x
