"""CModel: the reference's Caffe-style list-of-layers constructor (sota_imagenet/model.py:1098-1226)
with the same contract — a list of {module, args, kwargs, repeat, inputs, tag} entries whose
`module` strings are evaluated in this file's namespace, tagged multi-input graphs, `extra_kwargs`
merged under per-layer kwargs — resolving to the fused sm_100a modules of this package."""
from copy import deepcopy
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Union

import torch  # noqa: F401  (available to eval'ed config strings)
import torch.nn as nn

from .modules import (BatchNorm2d, Bottleneck, Concat, Conv2d, GlobalAvgPool, Linear,  # noqa: F401
                      MaxPool3x3s2, StemConv)
from .losses import SphereLinearLayer  # noqa: F401


@dataclass
class ModuleStructure:
    module: Union[str, nn.Module]
    args: List = field(default_factory=lambda: tuple())
    kwargs: Dict[str, Any] = field(default_factory=dict)
    repeat: int = 1
    inputs: List[str] = field(default_factory=lambda: ["_prev_"])
    tag: Optional[str] = None


def listify(x):
    if x is None:
        return []
    if isinstance(x, (list, tuple)):
        return list(x)
    return [x]


def _update_dict(to_dict, from_dict):
    """`to_dict.update(from_dict)` that recurses into nested dicts (reference model.py:1115-1123)."""
    for k, v in from_dict.items():
        if hasattr(v, "keys") and k in to_dict.keys():
            _update_dict(to_dict[k], v)
        else:
            to_dict[k] = v
    return to_dict


class CModel(nn.Sequential):
    def __init__(self, layer_config, extra_kwargs=None):
        layer_config = [ModuleStructure(**layer) for layer in layer_config]
        if extra_kwargs is not None:
            self._update_config_with_extra_params(layer_config, extra_kwargs)
        layers, self.saved_layers_idx = self._parse_config(layer_config)
        super().__init__(*layers)
        if len(self.saved_layers_idx) > 0:
            self.forward = self.custom_forward

    @staticmethod
    def _update_config_with_extra_params(layer_config, extra_kwargs):
        for extra_layer_name, extra_layer_kwargs in extra_kwargs.items():
            for layer in layer_config:
                if layer.module == extra_layer_name:
                    layer.kwargs = _update_dict(deepcopy(extra_layer_kwargs), layer.kwargs)

    @staticmethod
    def _parse_config(layer_config):
        saved_layers_idx, layers = [], []
        tag_to_idx = {layer.tag: idx for idx, layer in enumerate(layer_config) if layer.tag is not None}
        tag_to_idx["_prev_"] = -1
        maybe_eval = lambda x: eval(x) if isinstance(x, str) else x  # noqa: E731
        for layer_idx, l in enumerate(layer_config):
            l.module = maybe_eval(l.module)
            l.args = [maybe_eval(i) for i in listify(l.args)]
            l.kwargs = {k: maybe_eval(v) for k, v in l.kwargs.items()}
            m = l.module(*l.args, **l.kwargs)
            if l.repeat > 1:
                m = nn.Sequential(*[l.module(*l.args, **l.kwargs) for _ in range(l.repeat)])
            m.input_indexes = [tag_to_idx[inp] for inp in l.inputs]
            m.idx = layer_idx
            layers.append(m)
            saved_layers_idx.extend(idx for idx in m.input_indexes if idx != -1)
        return nn.ModuleList(layers), saved_layers_idx

    def custom_forward(self, x):
        saved_outputs = []
        for layer in self.children():
            inp = [x if j == -1 else saved_outputs[j] for j in layer.input_indexes]
            x = layer(*inp)
            saved_outputs.append(x if layer.idx in self.saved_layers_idx else None)
        return x
