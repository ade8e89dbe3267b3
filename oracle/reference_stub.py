"""Import the reference's own angular_losses.py / optimizers.py in THIS container (they need
`pytorch_tools`, absent): stub the two names it imports.  Used only by oracle/make_golden.py to
generate tests/golden/*.pt; /root/reference does not exist on the GPU box."""
import sys
import types

import torch.nn as nn

REFERENCE_ROOT = "/root/reference"


def load_reference_modules():
    from oracle.torch_ref import smooth_cross_entropy

    class Loss(nn.Module):
        pass

    class CrossEntropyLoss(Loss):
        def __init__(self, mode="multiclass", smoothing=0.0, weight=1.0, reduction="mean",
                     temperature=1.0, normalize=False):
            super().__init__()
            self.smoothing, self.temperature = smoothing, temperature

        def forward(self, y_pred, y_true):
            return smooth_cross_entropy(y_pred, y_true, self.smoothing, self.temperature)

    losses = types.ModuleType("pytorch_tools.losses")
    losses.Loss, losses.CrossEntropyLoss = Loss, CrossEntropyLoss
    pt = types.ModuleType("pytorch_tools")
    pt.losses, pt.__version__ = losses, "stub"
    sys.modules.setdefault("pytorch_tools", pt)
    sys.modules.setdefault("pytorch_tools.losses", losses)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import sota_imagenet.angular_losses as ang
    import sota_imagenet.optimizers as opt
    return ang, opt, CrossEntropyLoss
