"""In-tree build of libsib200.so (sm_100a CUDA kernels + C ABI) and the C oracle.

`python -m sota_imagenet_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles
without a GPU; the resulting .so is git-ignored but travels to the GPU box with the tree.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
OUT_DIR = os.path.join(ROOT, "_build")
LIB = os.path.join(ROOT, "libsib200.so")
SOURCES = ["host.cu", "conv.cu", "norm.cu", "head.cu", "optim.cu", "augment.cu", "extra.cu", "peer.cu", "batchaug.cu", "jpeg.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths):
    h = hashlib.sha1()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_variant(name, extra_flags):
    """Experimental second library (libsib200_<name>.so) with extra nvcc flags, selected at run time
    with SIB_LIB_VARIANT=<name> (A/B measurements on the same box)."""
    global OUT_DIR, LIB, NVCC_FLAGS
    saved = OUT_DIR, LIB, list(NVCC_FLAGS)
    try:
        OUT_DIR = os.path.join(ROOT, "_build_" + name)
        LIB = os.path.join(ROOT, "libsib200_%s.so" % name)
        NVCC_FLAGS = NVCC_FLAGS + list(extra_flags)
        return build()
    finally:
        OUT_DIR, LIB, NVCC_FLAGS = saved


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(ROOT, "..", "include", "sib200.h"))
    stamp = os.path.join(OUT_DIR, "stamp")
    dig = _digest(deps)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".log", "w") as f:
            f.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr[-4000:]))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                  "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
