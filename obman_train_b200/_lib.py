"""ctypes binding of libobman_b200.so.

The prototypes are parsed from ``include/obman_b200.h`` (single source of truth for the C ABI), so a
signature change cannot silently desynchronise the Python side.  There is NO fallback: if the library
is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libobman_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "obman_b200.h")

_PROTO_RE = re.compile(r"^\s*(const\s+char\s*\*|int|void)\s+(obman_\w+)\s*\(([^)]*)\)\s*;", re.M | re.S)


def parse_header(path=HEADER_PATH):
    """Return {name: (restype, [argtypes])} for every prototype declared in the header."""
    with open(path) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    protos = {}
    for ret, name, args in _PROTO_RE.findall(text):
        restype = ctypes.c_char_p if "char" in ret else (None if ret == "void" else ctypes.c_int)
        argtypes = []
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                elif a.startswith("float "):
                    argtypes.append(ctypes.c_float)
                elif a.startswith("double "):
                    argtypes.append(ctypes.c_double)
                elif a.startswith("long long "):
                    argtypes.append(ctypes.c_longlong)
                elif a.startswith("int "):
                    argtypes.append(ctypes.c_int)
                else:
                    raise ValueError("unsupported C type in header: " + a)
        protos[name] = (restype, argtypes)
    return protos


_lib = None
_protos = None


def load():
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libobman_b200.so is missing ({}); build it with `python -m obman_train_b200.build` "
            "(there is no CPU or library fallback)".format(LIB_PATH))
    lib = ctypes.CDLL(LIB_PATH)
    _protos = parse_header()
    for name, (restype, argtypes) in _protos.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


launch_count = 0  # kernel-launching C calls issued
kernel_count = 0  # kernels those calls launched (bench.py reads this for its gpu_launches claim)
KERNELS_PER_CALL = {"obman_chamfer_fwd": 2, "obman_contact_fwd": 2, "obman_mano_fwd": 3, "obman_mano_bwd": 3,
                    "obman_pointmlp_l1_bwd": 2, "obman_raycast_hits": 2, "obman_laplacian_fwd": 2, "obman_edge_loss_fwd": 2, "obman_contact_iou": 2, "obman_bn_stats": 2, "obman_bn_bwd": 3}


def call(name, *args):
    """Call an int-returning entry point; raise RuntimeError(obman_get_last_error()) on failure."""
    global launch_count, kernel_count
    lib = load()
    rc = getattr(lib, name)(*args)
    launch_count += 1
    kernel_count += KERNELS_PER_CALL.get(name, 1)
    if rc != 0:
        msg = lib.obman_get_last_error()
        raise RuntimeError("{} failed (rc={}): {}".format(name, rc, msg.decode() if msg else "?"))
    return rc


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
