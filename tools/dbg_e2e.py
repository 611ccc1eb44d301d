import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from side_b200 import ops, _lib
from side_b200.networks import get_pose_net
from side_b200.utils.synthetic import HEADS, realistic_init, make_batch
rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else None
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
m = realistic_init(get_pose_net(34, HEADS, 256), seed=1).eval().cuda()
batch = {k: v.cuda() for k, v in make_batch(1, 64, 1280, seed=3).items()}
lib = _lib.load()
orig = {}
def wrap(name):
    fn = getattr(lib, name)
    def w(*a):
        rc = fn(*a)
        torch.cuda.synchronize()
        print("  call", name, "rc", rc, lib.side_last_error() if rc else "")
        return rc
    return w
class L:  # proxy printing each call
    def __getattr__(self, n):
        f = getattr(lib, n)
        if n.startswith("side_") and n not in ("side_last_error", "side_abi_version"):
            return wrap(n)
        return f
_lib._lib = L()
with torch.no_grad():
    try:
        z = m(batch, True, None, 1.0)[0]
        print("ok", {k: tuple(v.shape) for k, v in z.items()})
    except Exception as e:
        print("EXC", e)
