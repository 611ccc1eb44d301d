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
        self.keep_heads = False      # True: the head maps of the last step stay in .last_heads ('hm' after the sigmoid)
        self.last_heads = None
        self.range_fallbacks = 0     # steps process_checked() had to repeat in 3xTF32

    def pre_process(self, opt, image, image_right, calib, device="cuda"):
        """stereoDetector.pre_process (modules/stereoDetector.py:45-82): raw uint8 H x W x 3 images -> normalised network inputs
        on the device + the ``meta`` dict (one kernel for both images, see side_b200/preprocess.py)."""
        from .preprocess import pre_process
        return pre_process(opt, image, image_right, calib, device)

    @torch.no_grad()
    def process(self, batch):
        out = self.model(batch, useCostVolume=self.use_cost_volume, wh_scale=self.wh_scale)[-1]
        heat = out['hm'].sigmoid_()
        dets, dets_right, info = ddd_decode(heat, out['kept_type'], out['dim'], out['orien'], wh=out['wh'], reg=out['reg'],
                                            grid_size=self.grid, K=self.K)
        if self.use_cost_volume:
            info = torch.cat([info, out['depth']], dim=2)
        if self.keep_heads:
            self.last_heads = out
        return dets, dets_right, info

    def process_checked(self, batch):
        """``process`` + the range guard of the 3xFP16 operand pairs: the kernels that write fp16 pairs report saturation
        (an activation reached 65504) / underflow (a tensor below 2^-18 everywhere) into a device status word; when it is set
        the step is repeated with the 3xTF32 kernels (8-bit exponent, same accuracy class, half the tensor rate).  Costs one
        4-byte device-to-host copy per step -- free in a serving loop that reads the detections anyway."""
        dev = batch['input'].device
        fmt, prec = ops.get_tc_format(), ops.get_dcn_precision()
        if not batch['input'].is_cuda or (fmt != "f16" and prec != "3xfp16"):
            return self.process(batch)
        ops.tc_range_status(dev)                         # clear what earlier work left behind
        out = self.process(batch)
        if ops.tc_range_status(dev):
            self.range_fallbacks += 1
            ops.set_tc_format("tf32")
            if prec == "3xfp16":
                ops.set_dcn_precision("3xtf32")
            try:
                out = self.process(batch)
            finally:
                ops.set_tc_format(fmt)
                ops.set_dcn_precision(prec)
        return out

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


def _buckets(params, bucket_bytes):
    """Rank-independent bucketing: over ALL parameters that require a gradient, in reverse registration order (roughly
    the order backward produces them), never over `p.grad is not None` -- a rank whose batch skipped a branch (no boxes:
    the depth branch of stereo_network.forward does not run) still takes part in every collective with zeros."""
    out, cur, size = [], [], 0
    for p in reversed([p for p in params if p.requires_grad]):
        cur.append(p)
        size += p.numel() * p.element_size()
        if size >= bucket_bytes:
            out.append(cur)
            cur, size = [], 0
    if cur:
        out.append(cur)
    return out


def allreduce_gradients(params, group=None, bucket_bytes=25 << 20):
    """Averages gradients over ranks with bucketed asynchronous all-reduces AFTER backward (replaces DataParallel's
    reduce-add to GPU 0 + per-step parameter broadcast, data_parallel.py:70-72).  Parameters without a gradient on this
    rank contribute zeros and receive the average, so bucket shapes and the collective sequence are the same on every rank.
    BatchNorm statistics stay per replica, as in the reference.  Returns the number of buckets."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    pending = []
    for bk in _buckets(list(params), bucket_bytes):
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bk])
        pending.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True), flat, bk))
    for work, flat, bk in pending:
        work.wait()
        flat.div_(world)
        off = 0
        for p in bk:
            n = p.numel()
            if p.grad is None:
                p.grad = flat[off:off + n].view_as(p).clone()
            else:
                p.grad.copy_(flat[off:off + n].view_as(p))
            off += n
    return len(pending)


class GradientAllReducer:
    """Bucketed gradient averaging OVERLAPPED with backward (BASELINE config #5).

    Every bucket owns one flat buffer and the parameters' ``.grad`` are views into it (no gather / scatter copies).  A
    post-accumulate hook per parameter counts the bucket down; a complete bucket is all-reduced asynchronously on NCCL's
    stream while autograd keeps producing the earlier layers' gradients.  Buckets are launched strictly in index order and
    ``finish()`` launches whatever backward did not complete (parameters that got no gradient this step keep their zeros), so
    every rank issues the same sequence of equally sized collectives whatever its batch looked like.

        red = GradientAllReducer(model.parameters())
        red.zero_grad(); loss.backward(); red.finish(); opt.step()
    """

    def __init__(self, params, group=None, bucket_bytes=25 << 20):
        self.group = group
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.enabled else 1
        self.buckets = _buckets(list(params), bucket_bytes)
        self.flat, self._bucket_of = [], {}
        for bi, bk in enumerate(self.buckets):
            flat = torch.zeros(sum(p.numel() for p in bk), device=bk[0].device, dtype=bk[0].dtype)
            off = 0
            for p in bk:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
                self._bucket_of[p] = bi
                if self.enabled:
                    p.register_post_accumulate_grad_hook(self._hook)
            self.flat.append(flat)
        self.bytes_per_step = sum(f.numel() * f.element_size() for f in self.flat)
        self._reset()

    def _reset(self):
        self._left = [len(bk) for bk in self.buckets]
        self._next = 0
        self._work = []

    def zero_grad(self):
        """Zeroes the flat buffers and re-points ``.grad`` at them (use instead of optimizer.zero_grad(set_to_none=True))."""
        for bk, flat in zip(self.buckets, self.flat):
            flat.zero_()
            off = 0
            for p in bk:
                if p.grad is None or p.grad.data_ptr() != flat.data_ptr() + off * flat.element_size():
                    p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
        self._reset()

    def _launch_ready(self, force=False):
        while self._next < len(self.buckets) and (force or self._left[self._next] <= 0):
            self._work.append(dist.all_reduce(self.flat[self._next], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self._next += 1

    def _hook(self, p):
        self._left[self._bucket_of[p]] -= 1
        self._launch_ready()

    def finish(self):
        """After backward: launch the buckets still waiting (in order), wait for all, divide by the world size.
        Returns the number of buckets whose all-reduce was already in flight when backward returned."""
        if not self.enabled:
            return 0
        early = self._next
        self._launch_ready(force=True)
        for w in self._work:
            w.wait()
        for flat in self.flat:
            flat.div_(self.world)
        self._reset()
        return early


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

    def split_by_work(self, threshold, key=None):
        """(summary of the calls with work >= threshold, summary of the others): large layers vs the many small launches."""
        torch.cuda.synchronize()
        out = []
        for big in (True, False):
            recs = [r for r in self.records if (key is None or r[3] == key) and ((r[2] >= threshold) == big)]
            out.append(dict(calls=len(recs), ms=sum(e0.elapsed_time(e1) for e0, e1, _, _ in recs), work=sum(r[2] for r in recs)))
        return tuple(out)
