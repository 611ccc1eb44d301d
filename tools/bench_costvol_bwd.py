"""Time the instance cost volume backward at config #2 (64 RoIs x 48 candidates x 64 channels): gather kernel, scalar-atomic
kernel, gated composition."""
import sys
import torch
sys.path.insert(0, ".")
from side_b200 import _lib, ops  # noqa: E402
from side_b200.utils.synthetic import make_boxes  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(30_000_000)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = "cuda"
    torch.manual_seed(0)
    fL, fR = torch.randn(1, 64, 96, 320, device=dev), torch.randn(1, 64, 96, 320, device=dev)
    left, right, _ = make_boxes(1, 64, seed=0)
    left, right, fb = left.to(dev), right.to(dev), torch.tensor([384.38], device=dev)
    byts = 64 * 192 * 48 * 256 * 4 + 2 * 64 * 96 * 320 * 4
    fLg, fRg = fL.clone().requires_grad_(True), fR.clone().requires_grad_(True)
    for gate in (False, True):
        cost, _ = ops.inst_costvol(fLg, fRg, left, right, fb, 48, 16, 319.0, gate=gate)
        gcost = torch.randn_like(cost)
        fn = lambda: torch.autograd.grad(cost, (fLg, fRg), gcost, retain_graph=True)  # noqa: E731
        t = timeit(fn)
        ops.VOL_BWD_FLAGS = _lib.VOL_BWD_SCALAR
        ts = timeit(fn, 3)
        ops.VOL_BWD_FLAGS = 0
        print(f"inst_costvol_bwd gate={gate}: gather {t*1e3:8.1f} us ({byts/t/1e6:7.0f} GB/s)   scalar atomics {ts*1e3:8.1f} us", flush=True)
        del cost, gcost


if __name__ == "__main__":
    main()
