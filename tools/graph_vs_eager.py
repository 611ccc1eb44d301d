"""Detector step at the bench's micro-batch: eager launches vs one CUDA-graph replay (how much of the step is launch gaps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from side_b200 import _lib, ops
from side_b200.engine import StereoDetector
from side_b200.utils.synthetic import KITTI_FB
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda")
_lib.load()
ops.set_dcn_precision("3xfp16"); ops.set_tc_format("f16")
model = bench.build_model().to(dev)
det = StereoDetector(model, grid_size=28, K=100)
if os.environ.get('SIDE_NO_IDAUP_FUSE'):
    from side_b200.networks.feature_extraction_dla34 import IDAUp
    IDAUp.fuse_up_add = False
g = torch.Generator().manual_seed(1)
batch = {'input': torch.randn(mb, 3, 384, 1280, generator=g).to(dev), 'input_right': torch.randn(mb, 3, 384, 1280, generator=g).to(dev),
         'fb': torch.full((mb,), KITTI_FB, device=dev)}
def timeit(fn, n=6):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
eager = timeit(lambda: det.process(batch))
det.capture(batch)
graph = timeit(lambda: det.replay(batch))
print("micro-batch %d: eager %.2f ms, graph %.2f ms (%.1f %%), %d library launches" % (mb, eager, graph, 100 * (eager - graph) / eager, det.launches_per_step))
