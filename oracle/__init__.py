"""CPU oracle for the sota_imagenet training-step hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under sota_imagenet_b200/ imports this package; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as
the checker or the reported CPU baseline, never as the product path.

Pinning: the reference ships no golden vectors.  Pieces whose source is in
/root/reference (angular_losses.py, optimizers.py) are pinned by running the reference file
itself (oracle/make_golden.py -> tests/golden/).  Pieces that live in the absent, unpinned
third-party `pytorch_tools@dev` (ResNet-50 class, smooth CrossEntropyLoss) and in NVIDIA DALI
are restated and anchored on torchvision / torch.nn equivalents: for those, parity is
"unpinned" by the reference and says so in DESIGN.md.
"""
