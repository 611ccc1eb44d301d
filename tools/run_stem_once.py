"""The three DLA-34 stem convolutions at the bench shape (16 images of 384x1280) a few times (for ncu)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from side_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
x = torch.randn(16, 3, 384, 1280, device=dev)
w1, w2, w3 = (torch.randn(16, 3, 7, 7, device=dev) * 0.1, torch.randn(16, 16, 3, 3, device=dev) * 0.1,
              torch.randn(32, 16, 3, 3, device=dev) * 0.1)
sc16, sh16, sc32, sh32 = torch.rand(16, device=dev), torch.randn(16, device=dev), torch.rand(32, device=dev), torch.randn(32, device=dev)
for _ in range(2):
    y = ops.stem_conv(x, w1, sc16, sh16, stride=1)
    y = ops.stem_conv(y, w2, sc16, sh16, stride=1)
    y = ops.stem_conv(y, w3, sc32, sh32, stride=2)
torch.cuda.synchronize()
print("ok")
