"""CPU: the C-ABI library loads, exports every symbol include/sib200.h declares, and refuses to
compute without a GPU (no CPU fallback anywhere in the product path)."""
import os
import subprocess

import pytest
import torch

from sota_imagenet_b200 import _lib, ops


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    protos = _lib.header_prototypes()
    assert len(protos) >= 30
    for name in protos:
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert set(protos) <= exported
    assert lib.sib_abi_version() == 2


def test_library_contains_blackwell_kernels():
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass          # tcgen05.mma
    assert "UTMALDG" in sass          # TMA loads (tiled + im2col)
    assert "LDTM" in sass             # tcgen05.ld
    assert "HMMA." not in sass.replace("UTCHMMA", "")   # no legacy mma.sync path


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    x = torch.zeros(1, 64, 4, 4)
    w = torch.zeros(64, 64, 1, 1)
    with pytest.raises(_lib.SibError):
        ops.conv2d_fprop(x, w)
    from sota_imagenet_b200 import models
    with pytest.raises(_lib.SibError):
        models.resnet50()(torch.zeros(1, 3, 32, 32))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: only tests/ (incl. tests/tools), __graft_entry__.smoke()
    and bench.py's CPU-baseline / reference legs may touch it."""
    top = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    roots = [os.path.join(top, "sota_imagenet_b200"), os.path.join(top, "scripts"), os.path.join(top, "include")]
    for root in roots:
        for dirpath, _, files in os.walk(root):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in text and "from oracle" not in text, f
    text = open(os.path.join(top, "train.py")).read()
    assert "oracle" not in text
    # bench.py: the oracle appears only inside the two CPU legs
    import ast
    tree = ast.parse(open(os.path.join(top, "bench.py")).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name not in ("cpu_oracle_rate", "run_reference"):
            assert "oracle" not in {n.module for n in ast.walk(node) if isinstance(n, ast.ImportFrom)}, node.name
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            assert "oracle" not in ast.dump(node)
