"""ctypes binding of libsib200.so (the C ABI declared in include/sib200.h).

There is deliberately no fallback: if the shared library cannot be loaded, or the device is
not an sm_100a GPU, every op raises.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsib200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "sib200.h")

_lib = None

c_void_p, c_int, c_long, c_float, c_double = (ctypes.c_void_p, ctypes.c_int, ctypes.c_long,
                                              ctypes.c_float, ctypes.c_double)
c_ull = ctypes.c_ulonglong

_CTYPE = {
    "int": c_int, "long": c_long, "float": c_float, "double": c_double,
    "unsigned long long": c_ull,
}


def header_prototypes(path=HEADER_PATH):
    """Parse `int sib_xxx(args);` prototypes out of the public header -> {name: (restype, [argtypes])}."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const char\*|int|void)\s+(sib_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        argtypes = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(c_void_p)
                else:
                    ty = a.rsplit(" ", 1)[0].replace("const ", "").strip()
                    argtypes.append(_CTYPE[ty])
        restype = {"int": c_int, "void": None, "const char*": ctypes.c_char_p}[ret]
        protos[name] = (restype, argtypes)
    return protos


def load():
    """Load (building first if the .so is absent) and type every exported entry point."""
    global _lib
    if _lib is not None:
        return _lib
    # build.build() is a no-op when the source digest matches the stamp of the built library: a stale
    # .so (missing newer symbols) is rebuilt instead of failing later with AttributeError.  On a box
    # without nvcc the shipped library is used as is.
    try:
        from . import build as _build
        _build.build()
    except Exception:
        if not os.path.exists(LIB_PATH):
            raise
    variant = os.environ.get("SIB_LIB_VARIANT")       # experimental A/B builds (build.build_variant)
    path = os.path.join(_HERE, "libsib200_%s.so" % variant) if variant else LIB_PATH
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in header_prototypes().items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.sib_abi_version() != 2:
        raise RuntimeError("libsib200.so ABI version mismatch")
    _lib = lib
    return lib


class SibError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = load().sib_last_error()
        raise SibError("libsib200: " + (msg.decode() if msg else "error %d" % rc))


CALLS = 0          # number of C-ABI compute calls issued (each launches >= 1 kernel)
PROFILE = None     # when a list: (name, args, start_event, end_event) per call, for bench.py


def call(name, *args):
    global CALLS
    CALLS += 1
    if PROFILE is None:
        check(getattr(load(), name)(*args))
        return
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(getattr(load(), name)(*args))
    e1.record()
    PROFILE.append((name, tuple(a.value if hasattr(a, "value") else a for a in args), e0, e1))


_device_ok = False


def require_device():
    """Raise unless a B200-class (sm_100) device is current.  No CPU path exists."""
    global _device_ok
    if _device_ok:
        return
    import torch
    if not torch.cuda.is_available():
        raise SibError("sota_imagenet_b200 needs a CUDA sm_100a device; no CPU fallback exists")
    check(load().sib_device_check())
    _device_ok = True
