"""Inference / training plumbing around the hot path: detector step, pair sharding, NCCL collectives.

The reference's only parallelism is single-process ``nn.DataParallel`` (models/data_parallel.py:119-128): per
iteration it broadcasts all parameters, scatters the batch and gathers outputs on GPU 0.  Here the job is
one process per GPU (``torchrun``); stereo pairs are independent, so ranks share nothing on the data path
(SURVEY.md section 8e).  NCCL is used for exactly two things: gathering the fixed-shape detections at the end of an
inference step and averaging gradients in the training configuration.
"""
import torch
import torch.distributed as dist

from . import _lib, ops
from .decode import ddd_decode


def shard_range(total, rank, world):
    """Contiguous block of ``total`` units owned by ``rank`` (32/G pairs per GPU)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class StereoDetector:
    """The detector's ``process`` step (modules/stereoDetector.py:84-103): forward, sigmoid, ddd_decode, append depth.

    ``capture=True`` records the whole step into a CUDA graph (static input / output buffers) -- useful for
    batch-1 latency where ~400 launches would otherwise bound the step; large micro-batches run eagerly.
    """

    def __init__(self, model, grid_size=28, K=100, use_cost_volume=True, wh_scale=1.0):
        self.model = model.eval()
        self.grid, self.K = grid_size, K
        self.use_cost_volume, self.wh_scale = use_cost_volume, wh_scale
        self.model.K = K
        self._graph = None
        self._static = None

    @torch.no_grad()
    def process(self, batch):
        out = self.model(batch, useCostVolume=self.use_cost_volume, wh_scale=self.wh_scale)[-1]
        heat = out['hm'].sigmoid_()
        dets, dets_right, info = ddd_decode(heat, out['kept_type'], out['dim'], out['orien'], wh=out['wh'], reg=out['reg'],
                                            grid_size=self.grid, K=self.K)
        if self.use_cost_volume:
            info = torch.cat([info, out['depth']], dim=2)
        return dets, dets_right, info

    # -- CUDA graph path -----------------------------------------------------------------------
    def capture(self, example_batch, warmup=3):
        dev = example_batch['input'].device
        self._static = {k: v.clone() for k, v in example_batch.items()}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.process(self._static)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(g):
            self._static_out = self.process(self._static)
        self.launches_per_step = _lib.launch_count() - n0
        self._graph = g
        return self

    def replay(self, batch):
        for k in ('input', 'input_right', 'fb'):
            self._static[k].copy_(batch[k], non_blocking=True)
        self._graph.replay()
        return self._static_out


def gather_detections(dets, dets_right, info, group=None):
    """All-gathers the fixed-shape per-rank detections ([B_local,K,6], [B_local,K,6], [B_local,K,9|10]).
    One flat buffer, one NCCL call (8.8 KB per pair).  Returns per-rank lists on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [dets], [dets_right], [info]
    world = dist.get_world_size(group)
    B, K = dets.shape[:2]
    flat = torch.cat([dets.reshape(B, K, -1), dets_right.reshape(B, K, -1), info.reshape(B, K, -1)], dim=2).contiguous()
    out = torch.empty((world * B,) + tuple(flat.shape[1:]), device=flat.device, dtype=flat.dtype)
    dist.all_gather_into_tensor(out, flat, group=group)
    out = out.view((world, B) + tuple(flat.shape[1:]))
    a, b = dets.shape[2], dets.shape[2] + dets_right.shape[2]
    return ([out[r, :, :, :a] for r in range(world)], [out[r, :, :, a:b] for r in range(world)],
            [out[r, :, :, b:] for r in range(world)])


def allreduce_gradients(params, group=None, bucket_bytes=25 << 20):
    """Averages gradients over ranks with bucketed asynchronous all-reduces (replaces DataParallel's reduce-add to
    GPU 0 + per-step parameter broadcast).  BatchNorm statistics stay per replica, as in the reference."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    buckets, cur, size = [], [], 0
    for g in grads:
        cur.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            buckets.append(cur)
            cur, size = [], 0
    if cur:
        buckets.append(cur)
    pending = []
    for bk in buckets:
        flat = torch.cat([g.reshape(-1) for g in bk])
        pending.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True), flat, bk))
    for work, flat, bk in pending:
        work.wait()
        flat.div_(world)
        off = 0
        for g in bk:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
    return len(buckets)


class OpTimer:
    """CUDA-event timing of one operator family inside a running step (bench.py roofline).

    Wraps ``ops.dcn_forward_raw`` (or any named attribute of ``side_b200.ops``): an event pair is recorded on the
    current stream around every call, nothing synchronises until ``summary()``."""

    def __init__(self, name, work_fn, key_fn=None):
        self.name, self.work_fn, self.key_fn = name, work_fn, key_fn
        self.records = []
        self._orig = None

    def __enter__(self):
        self._orig = getattr(ops, self.name)
        orig, recs, work_fn, key_fn = self._orig, self.records, self.work_fn, self.key_fn

        def timed(*a, **k):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            out = orig(*a, **k)
            e1.record()
            recs.append((e0, e1, work_fn(out, *a, **k), key_fn(*a, **k) if key_fn else None))
            return out

        setattr(ops, self.name, timed)
        return self

    def __exit__(self, *exc):
        setattr(ops, self.name, self._orig)

    def summary(self, key=None):
        """Totals over all recorded calls, or over those whose ``key_fn`` value equals ``key``."""
        torch.cuda.synchronize()
        recs = [r for r in self.records if key is None or r[3] == key]
        ms = sum(e0.elapsed_time(e1) for e0, e1, _, _ in recs)
        work = sum(w for _, _, w, _ in recs)
        return dict(calls=len(recs), ms=ms, work=work)
