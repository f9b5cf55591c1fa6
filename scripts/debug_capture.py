"""Find the first library call around which a CUDA-graph capture of the training driver's step gets invalidated.

    python scripts/debug_capture.py [img_size] [batch]
"""
import ctypes
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from obman_train_b200 import _lib  # noqa: E402

rt = ctypes.CDLL("libcudart.so.12")
orig = _lib.call
state = {"last": None, "reported": False, "n": 0}


def status():
    st = ctypes.c_int(0)
    rc = rt.cudaStreamIsCapturing(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.byref(st))
    return rc, st.value


def wrapped(name, *args):
    b = status()
    if (b[0] != 0 or b[1] == 2) and not state["reported"]:
        state["reported"] = True
        print("capture invalid BEFORE %s (call #%d); previous library call: %s; status=%s" % (name, state["n"], state["last"], b),
              flush=True)
    r = orig(name, *args)
    a = status()
    if (a[0] != 0 or a[1] == 2) and not state["reported"]:
        state["reported"] = True
        print("capture invalid AFTER %s (call #%d); status=%s args=%s" % (name, state["n"], a, args), flush=True)
    state["last"] = name
    state["n"] += 1
    return r


_lib.call = wrapped
import obman_train_b200  # noqa: E402
import importlib  # noqa: E402
import pkgutil  # noqa: E402

for m in pkgutil.walk_packages(obman_train_b200.__path__, "obman_train_b200."):
    try:
        mod = importlib.import_module(m.name)
    except Exception as e:  # noqa: BLE001
        continue
    if getattr(mod, "call", None) is orig:
        mod.call = wrapped

from obman_train_b200.netscripts import train  # noqa: E402

img = sys.argv[1] if len(sys.argv) > 1 else "64"
bs = sys.argv[2] if len(sys.argv) > 2 else "4"
tmp = tempfile.mkdtemp()
args = train.build_parser().parse_args(["--exp_id", os.path.join(tmp, "run"), "--epochs", "1", "--batch_size", bs,
                                        "--steps_per_epoch", "3", "--img_size", img, "--log_every", "0"])
try:
    train.run(args, model_kwargs=dict(atlas_ico_divisions=2, atlas_separate_encoder=False), out=print)
    print("run OK, %d library calls" % state["n"])
except Exception as e:  # noqa: BLE001
    print("run FAILED:", type(e).__name__, str(e)[:300])
