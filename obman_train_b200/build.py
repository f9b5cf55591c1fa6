"""In-tree build of libobman_b200.so (sm_100a only) with nvcc.

``python -m obman_train_b200.build`` or ``build()``; object files are cached under
``obman_train_b200/csrc/build`` keyed by source mtime.  The library links cudart statically and does
not link libcuda (the driver entry point for tensor-map encoding is fetched at run time), so it can be
dlopen-ed on a machine without a GPU driver for symbol checks.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(CSRC, "build")
LIB_PATH = os.path.join(HERE, "libobman_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    return max([os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC)
                if f.endswith((".cuh", ".h"))] + [0.0])


def _compile(src, log):
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
    newest = max(os.path.getmtime(src), _headers_mtime())
    if os.path.exists(obj) and os.path.getmtime(obj) >= newest:
        return obj
    cmd = [_nvcc()] + NVCC_FLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for {}:\n{}".format(src, r.stdout + r.stderr))
    return obj


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = sources()
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    log = []
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, log), srcs))
    newest_obj = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest_obj:
        cmd = [_nvcc(), "-shared", "-o", LIB_PATH] + objs + ["-cudart", "static",
                                                             "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    if verbose:
        print("\n".join(log))
    with open(os.path.join(OBJ_DIR, "build.log"), "a") as f:
        f.write("\n".join(log))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
