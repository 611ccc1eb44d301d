"""Time side_dcn_bwd (channels-last path vs the scalar kernel) on the DLA-34 up-path shapes, torchvision beside it.
usage: python tools/bench_dcn_bwd.py [B] [iters]"""
import sys
import torch
sys.path.insert(0, ".")
from side_b200 import _lib, ops  # noqa: E402

SHAPES = [(64, 64, 96, 320), (128, 64, 48, 160), (128, 128, 48, 160), (256, 128, 24, 80), (256, 256, 24, 80),
          (256, 64, 24, 80), (512, 256, 12, 40)]


def timeit(fn, iters):
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
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    import torchvision.ops as tvo
    torch.backends.cuda.matmul.allow_tf32 = False
    for (Cin, Cout, H, W) in SHAPES:
        torch.manual_seed(0)
        x = torch.randn(B, Cin, H, W, device="cuda")
        off = torch.randn(B, 18, H, W, device="cuda") * 2
        mask = torch.sigmoid(torch.randn(B, 9, H, W, device="cuda"))
        w = (torch.rand(Cout, Cin, 3, 3, device="cuda") * 2 - 1) / (9 * Cin) ** 0.5
        gy = torch.randn(B, Cout, H, W, device="cuda")
        t_new = timeit(lambda: ops.dcn_backward_raw(x, off, mask, w, gy, 1, 1, 1, 1), iters)
        t_old = timeit(lambda: ops.dcn_backward_raw(x, off, mask, w, gy, 1, 1, 1, 1, flags=_lib.DCN_BWD_SCALAR), iters)
        xr, offr, mr, wr = (t.clone().requires_grad_(True) for t in (x, off, mask, w))
        y = tvo.deform_conv2d(xr, offr, wr, None, padding=1, mask=mr)
        t_tv = timeit(lambda: torch.autograd.grad(y, (xr, offr, mr, wr), gy, retain_graph=True), iters)
        print(f"dcn_bwd B={B} {Cin:3d}->{Cout:3d} @{H}x{W}: channels-last {t_new*1e3:8.1f} us  scalar {t_old*1e3:8.1f} us  "
              f"torchvision bwd {t_tv*1e3:8.1f} us", flush=True)


if __name__ == "__main__":
    main()
