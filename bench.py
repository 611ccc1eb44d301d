#!/usr/bin/env python
"""bench.py -- SIDE DLA-34 stereo inference throughput (pairs/s at 384x1280) on N B200s + kernel rooflines.

Contract (see DESIGN.md "Measurement"):
  python bench.py --gpus N --steps K --warmup W            one JSON line on rank 0
  python bench.py --impl reference ...                     the reference's CPU path (port) on the host cores

A step = one pass of the hot path (DLA-34 + DCN neck, heads, decode, instance depth branch, ddd_decode) over
`--pairs` synthetic stereo pairs per GPU in micro-batches of `--micro-batch`; ranks are sharded by pair
(weak scaling: fixed pairs per GPU), the only collective is the all-gather of the fixed-shape detections.
  value : pairs/s with inputs resident in HBM       e2e : same through host buffers (H2D + D2H inside the timing)
The same line also carries, measured in the same process right after the headline number:
  strong      : BASELINE config #4 as written -- 32 pairs TOTAL, 32 / N per rank (strong scaling)
  train       : BASELINE config #5 -- training step, 2 pairs per GPU, StereoLoss, NCCL gradient all-reduce overlapped with backward
  tc_formats  : the headline throughput under both operand formats of the tensor-core convolutions (3xFP16 / strict 3xTF32)
  parity      : agreement of the timed configuration with the CPU port on the pairs the CPU baseline processed
  f16_range   : the fp16 range guard's verdict over the whole timed region
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

H_IN, W_IN = 384, 1280
METRIC = "stereo pairs/sec (384x1280, DLA-34)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="infer", choices=["infer", "train"], help="train: BASELINE config #5 (training step)")
    ap.add_argument("--train-batch", type=int, default=2, help="pairs per GPU per training step (config #5: 16 pairs / 8 GPUs)")
    ap.add_argument("--pairs", type=int, default=32, help="stereo pairs per GPU per step (config #4: batch 32)")
    ap.add_argument("--micro-batch", type=int, default=16, help="pairs per detector call (32 pairs per step = 2 calls; 8: 201, 16: 206, 32: 209 pairs/s resident, but one 32-pair upload no longer overlaps compute end to end)")
    ap.add_argument("--dcn-precision", default=os.environ.get("SIDE_DCN_PRECISION", "3xfp16"), choices=["fp32", "3xtf32", "3xfp16", "tf32"],
                    help="3xfp16 (default): tcgen05 kind::f16 on fp16 hi/lo pairs under the range guard; 3xtf32: the exact tf32 hi/lo split "
                         "(both fp32-class, <= 1e-4 rel); tf32: single pass; fp32: SIMT")
    ap.add_argument("--tc-format", default="f16", choices=["tf32", "f16"],
                    help="operand format of the tensor-core convolutions: 3xTF32 or 3xFP16 (kind::f16 MMAs "
                         "on fp16 hi/lo pairs, the same 22-bit operands at twice the tensor rate)")
    ap.add_argument("--conv-mode", type=int, default=0, help="side_conv_tc_set_mode bit mask: 0 default, 1 halo reuse in the voxel-major kernel, 32 no role-swapped kernel")
    ap.add_argument("--cudnn-only", action="store_true",
                    help="keep the heads and the 3-D aggregation network on cuDNN fp32 (as the reference runs them)")
    ap.add_argument("--allow-tf32", action="store_true", help="let cuDNN use TF32 for the out-of-scope convolutions")
    ap.add_argument("--no-kernels", action="store_true", help="skip the per-kernel roofline microbenchmarks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the in-line training block (config #5)")
    ap.add_argument("--no-latency", action="store_true", help="skip the batch-1 latency block (config #1's GPU counterpart)")
    ap.add_argument("--cpu-pairs", type=int, default=2)
    ap.add_argument("--profiler-range", action="store_true",
                    help="cudaProfilerStart/Stop around the resident timed region (ncu --profile-from-start off)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, "/tmp/side_clocks_%d.csv" % os.getpid()

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.proc.wait()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_model():
    from side_b200.networks import get_pose_net
    from side_b200.utils.synthetic import HEADS, realistic_init
    torch.manual_seed(0)
    return realistic_init(get_pose_net(34, HEADS, 256), seed=1).eval()


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference-style port on the host cores (oracle/torch_port.py)
# ----------------------------------------------------------------------------------------------------
def cpu_reference_pairs_per_s(model_cpu, n_pairs, warmup, seed=100, keep=None):
    """Times the reference-style port; ``keep`` (a list) receives (batch, head maps incl. depth) of every timed pair so the GPU
    arm can be checked against them afterwards (oracle/parity.py)."""
    from oracle import torch_port
    from side_b200.engine import StereoDetector
    from side_b200.utils.synthetic import make_batch
    det = StereoDetector(model_cpu)
    det.keep_heads = keep is not None
    times = []
    with torch_port.reference_ops():
        for i in range(warmup + n_pairs):
            batch = make_batch(1, H_IN, W_IN, seed=seed + i)
            t0 = time.perf_counter()
            det.process(batch)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
                if keep is not None:
                    keep.append((batch, det.last_heads))
    return len(times) / sum(times), times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # all the host threads this box has (torchrun exports OMP_NUM_THREADS=1 to its workers)
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    torch.set_num_threads(max(ncpu, 1))
    model = build_model()
    cores = torch.get_num_threads()
    val, times = cpu_reference_pairs_per_s(model, args.steps, args.warmup)
    sample = "%d steps of 1 pair (batch 1, 384x1280, K=100 RoIs) after %d warm-up, torch CPU ops + torchvision deform_conv2d/roi_align" % (
        args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": "SIDE DLA-34 stereo inference, 384x1280, random-init (calibrated) weights, K=100", "pairs_per_step": 1},
            "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
# per-kernel rooflines (standalone launches at the BASELINE config shapes)
# ----------------------------------------------------------------------------------------------------
def time_op(fn, iters=10, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        torch.cuda._sleep(400000)      # ~0.2 ms of GPU idle spin: the host enqueues fn's launches behind it, so the
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)   # events time the GPU, not Python
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def kernel_rooflines(pk, precision):
    from side_b200 import ops
    from side_b200.utils.synthetic import make_boxes
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    out = {}
    torch.manual_seed(0)
    # config #2: instance volume, 64 RoIs x 48 candidates x 64 ch, P=16  (604 MB written + 15.7 MB read)
    fL, fR = torch.randn(1, 64, 96, 320, device=dev), torch.randn(1, 64, 96, 320, device=dev)
    left, right, _ = make_boxes(1, 64, seed=0)
    left, right, fb = left.to(dev), right.to(dev), torch.tensor([384.38], device=dev)
    byts = 64 * 192 * 48 * 256 * 4 + 2 * 64 * 96 * 320 * 4
    variants = {
        "inst_costvol_fwd_separable": lambda: ops.inst_costvol(fL, fR, left, right, fb, 48, 16, 319.0, separable=True),
        "inst_costvol_fwd_separable_xcross": lambda: ops.inst_costvol_ungated(fL, fR, left, right, fb, 48, 16, 319.0),
        "inst_costvol_fwd_separable_gate_2pass": lambda: ops.inst_costvol(fL, fR, left, right, fb, 48, 16, 319.0, gate=True, separable=True),
        "inst_costvol_fwd_exact": lambda: ops.inst_costvol(fL, fR, left, right, fb, 48, 16, 319.0),
        "inst_costvol_fwd_exact_gate": lambda: ops.inst_costvol(fL, fR, left, right, fb, 48, 16, 319.0, gate=True),
    }
    for name, fn in variants.items():
        ms = time_op(fn, flush=flush)
        out[name] = {"ms": ms, "GBs": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / pk["hbm"], "alg_bytes": byts}
    # config #2 backward (gradient of the volume w.r.t. both feature maps; reads the 604 MB gradient once, atomics into 15.7 MB)
    fLg, fRg = fL.clone().requires_grad_(True), fR.clone().requires_grad_(True)
    cost, _ = ops.inst_costvol(fLg, fRg, left, right, fb, 48, 16, 319.0)
    gcost = torch.randn_like(cost)
    ms = time_op(lambda: torch.autograd.grad(cost, (fLg, fRg), gcost, retain_graph=True), flush=flush, iters=5)
    out["inst_costvol_bwd"] = {"ms": ms, "GBs": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / pk["hbm"], "alg_bytes": byts}
    from side_b200 import _lib as _sl
    ops.VOL_BWD_FLAGS = _sl.VOL_BWD_SCALAR            # the scalar-atomic kernel (torchvision's thread mapping) for comparison
    try:
        ms = time_op(lambda: torch.autograd.grad(cost, (fLg, fRg), gcost, retain_graph=True), flush=flush, iters=3)
    finally:
        ops.VOL_BWD_FLAGS = 0
    out["inst_costvol_bwd_scalar_atomics"] = {"ms": ms, "GBs": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / pk["hbm"],
                                              "alg_bytes": byts}
    del cost
    cost, _ = ops.inst_costvol(fLg, fRg, left, right, fb, 48, 16, 319.0, gate=True)
    ms = time_op(lambda: torch.autograd.grad(cost, (fLg, fRg), gcost, retain_graph=True), flush=flush, iters=5)
    out["inst_costvol_bwd_gate"] = {"ms": ms, "GBs": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / pk["hbm"], "alg_bytes": byts,
                                    "note": "recompute raw volume + gate backward in place + gather backward"}
    del cost, gcost, fLg, fRg
    # reference-shaped volume: 100 RoIs x 16 x 32 ch
    l2, r2, _ = make_boxes(1, 100, seed=1)
    f32L, f32R = fL[:, :32].contiguous(), fR[:, :32].contiguous()
    l2, r2 = l2.to(dev), r2.to(dev)
    byts = 100 * 96 * 16 * 256 * 4 + 2 * 32 * 96 * 320 * 4
    ms = time_op(lambda: ops.inst_costvol_ungated(f32L, f32R, l2, r2, fb, 16, 16, 319.0), flush=flush)
    out["inst_costvol_fwd_separable_xcross_ref_shape"] = {"ms": ms, "GBs": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / pk["hbm"], "alg_bytes": byts}
    ms = time_op(lambda: ops.inst_costvol(f32L, f32R, l2, r2, fb, 16, 16, 319.0, gate=True), flush=flush)
    out["inst_costvol_fwd_exact_gate_ref_shape"] = {"ms": ms, "GBs": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / pk["hbm"], "alg_bytes": byts}
    # full-image concat / gwc volumes: C=64, D=48, 96x320
    ms = time_op(lambda: ops.concat_volume(fL, fR, 48), flush=flush)
    byts = 128 * 48 * 96 * 320 * 4 + 2 * 64 * 96 * 320 * 4
    out["concat_volume_fwd"] = {"ms": ms, "GBs": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / pk["hbm"], "alg_bytes": byts}
    ms = time_op(lambda: ops.gwc_volume(fL, fR, 48, 8), flush=flush)
    byts = 8 * 48 * 96 * 320 * 4 + 2 * 64 * 96 * 320 * 4
    out["gwc_volume_fwd"] = {"ms": ms, "GBs": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / pk["hbm"], "alg_bytes": byts}
    # soft-argmin and decode are latency bound: report microseconds
    lg, db = torch.randn(64, 48, 4, 4, device=dev), torch.rand(64, 48, device=dev) * 80
    out["softargmin_fwd"] = {"us": 1000 * time_op(lambda: ops.softargmin(lg, db))}
    hm = torch.randn(4, 3, 96, 320, device=dev); wh = torch.rand(4, 3, 96, 320, device=dev); reg = torch.rand(4, 3, 96, 320, device=dev)
    out["bbox_decode_B4"] = {"us": 1000 * time_op(lambda: ops.bbox_decode_raw(hm, wh, reg, K=100))}
    # config #3: DCN sweep over the DLA-34 up-path shapes, B=2 (left+right of one pair)
    shapes = [(512, 256, 12, 40), (256, 256, 24, 80), (256, 128, 24, 80), (128, 128, 48, 160), (128, 64, 48, 160), (64, 64, 96, 320),
              (256, 64, 24, 80)]
    dcn = {}
    for (Cin, Cout, H, W) in shapes:
        B = 2
        x = torch.randn(B, Cin, H, W, device=dev); off = torch.randn(B, 18, H, W, device=dev) * 2
        mask = torch.sigmoid(torch.randn(B, 9, H, W, device=dev)); w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.05
        b = torch.rand(Cout, device=dev)
        try:
            ms = time_op(lambda: ops.dcn_forward_raw(x, off, mask, w, b, 1, 1, 1, 1, precision=precision), flush=flush, iters=5)
        except RuntimeError as e:
            dcn["%dx%d@%dx%d" % (Cin, Cout, H, W)] = {"error": str(e)[:80]}
            continue
        fl = 2.0 * B * Cout * Cin * 9 * H * W
        rec = {"ms": ms, "TFLOPs": fl / ms / 1e9, "frac_tensor": fl / ms / 1e9 / pk["tf_burst"]}
        # the same op through torchvision's CUDA deform_conv2d on this GPU (BASELINE config #3), forward and forward+backward
        try:
            import torchvision
            rec["torchvision_fwd_ms"] = time_op(lambda: torchvision.ops.deform_conv2d(x, off, w, b, padding=1, mask=mask), flush=flush, iters=5)
            xg, wg = x.clone().requires_grad_(True), w.clone().requires_grad_(True)

            def tv_fb():
                y = torchvision.ops.deform_conv2d(xg, off, wg, b, padding=1, mask=mask)
                torch.autograd.grad(y, (xg, wg), torch.ones_like(y))

            def our_fb():
                y = ops.dcn_v2_conv(xg, off, mask, wg, b, 1, 1, 1, 1)
                torch.autograd.grad(y, (xg, wg), torch.ones_like(y))

            rec["torchvision_fwd_bwd_ms"] = time_op(tv_fb, flush=flush, iters=3)
            rec["fwd_bwd_ms"] = time_op(our_fb, flush=flush, iters=3)
        except Exception as e:   # torchvision missing on the box: report ours only
            rec["torchvision"] = "unavailable: %s" % str(e)[:60]
        dcn["%dx%d@%dx%d" % (Cin, Cout, H, W)] = rec
    out["dcn_fwd_" + precision] = dcn
    out.update(widened_rows(pk, flush))
    return out


def widened_rows(pk, flush):
    """The rows SURVEY.md 8(f) marks "next", at BASELINE sizes: stereo_network_new voxel volume (F3), dense photometric
    alignment (F2), detector input preparation (F4).  The CPU figures are the oracle's C restatement on one host thread over
    a bounded sample (reported beside the kernel, never the product path)."""
    import time
    import types
    import numpy as np
    from oracle import c_oracle as co
    from side_b200 import dense_align as da, ops, preprocess as pp
    dev = torch.device("cuda")
    out = {}
    rng = np.random.RandomState(0)
    p2 = np.array([[721.54, 0, 609.56, 44.86], [0, 721.54, 172.85, 0.216], [0, 0, 1, 0.00275]], np.float32)
    p3 = p2.copy(); p3[0, 3] = -339.52
    # ---- F3: voxel volume, 8 images x 100 RoIs, features 64 x 96 x 320 ----
    B, N = 8, 800
    c, s = np.array([621., 187.5], np.float32), np.array([1242, 375], np.int32)
    tr = pp.get_affine_transform(c, s, 0, [320, 96]).astype(np.float32)
    tri = pp.get_affine_transform(c, s, 0, [320, 96], inv=1).astype(np.float32)
    st = lambda a: torch.from_numpy(np.ascontiguousarray(np.broadcast_to(a, (B,) + a.shape))).float().to(dev)
    x1 = rng.uniform(0, 290, N); w = rng.uniform(6, 50, N); y1 = rng.uniform(25, 70, N); h = rng.uniform(5, 25, N)
    left = np.stack([np.repeat(np.arange(B), N // B), x1, y1, x1 + w, y1 + h], 1).astype(np.float32)
    right = left.copy(); d = rng.uniform(1.0, 12.0, N).astype(np.float32); right[:, 1] -= d; right[:, 3] -= d
    fL, fR = torch.randn(B, 64, 96, 320, device=dev), torch.randn(B, 64, 96, 320, device=dev)
    tl, trr, fb = torch.from_numpy(left).to(dev), torch.from_numpy(right).to(dev), torch.full((B,), 384.38, device=dev)
    P2, P3, TR, TRI = st(p2), st(p3), st(tr), st(tri)
    ms = time_op(lambda: ops.voxel_volume(fL, fR, tl, trr, P2, P3, fb, TR, TRI), flush=flush)
    byts = N * 192 * 1000 * 4 + 2 * B * 64 * 96 * 320 * 4
    t0 = time.time()
    co.voxel_volume(fL[:1].cpu().numpy(), fR[:1].cpu().numpy(), left[:20], right[:20], p2[None], p3[None], np.array([384.38], np.float32),
                    tr[None], tri[None])
    out["voxel_volume_fwd"] = {"ms": ms, "GBs": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / pk["hbm"], "alg_bytes": byts,
                               "rois": N, "cpu_oracle_ms_per_roi": 1000 * (time.time() - t0) / 20,
                               "note": "fused get_voxel + 2 x grid_sample + mask + cat(L-R, L, R), stereo_network_new.py:160-283,409-449"}
    # ---- F2: dense alignment, one 384 x 1280 frame (sampled at 768 x 2560), 100 RoIs, 50 + 20 hypotheses ----
    Hh, Ww, n = 384, 1280, 100
    img_l = rng.randint(0, 256, (Hh, Ww, 3), dtype=np.uint8)
    img_r = np.ascontiguousarray(np.roll(img_l, -10, axis=1))
    calib = types.SimpleNamespace(p2=p2, p3=p3)
    opt = types.SimpleNamespace(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
    xs = rng.uniform(-10, 10, n); zs = rng.uniform(8, 50, n)
    poses = np.stack([xs, np.full(n, 1.6), zs, np.full(n, 1.6), np.full(n, 1.5), np.full(n, 3.9), rng.uniform(-1.5, 1.5, n)], 1).astype(np.float32)
    u = 721.54 * xs / zs + 609.56; v = 721.54 * 1.6 / zs + 172.85; hw = 721.54 * 2.2 / zs
    box = np.stack([u - hw, v - 1.6 * hw, u + hw, v + 0.1 * hw], 1).astype(np.float32)
    borders = np.stack([u - 0.9 * hw, u + 0.9 * hw], 1).astype(np.float32)
    gl, gr = torch.from_numpy(img_l).to(dev), torch.from_numpy(img_r).to(dev)
    tb, tbd, tp_ = torch.from_numpy(box).to(dev), torch.from_numpy(borders).to(dev), torch.from_numpy(poses).to(dev)
    ms = time_op(lambda: da.align_parallel(calib, opt, gl, gr, tb, tbd, tp_), flush=flush, iters=5)
    L = da.prepare_image(gl, opt.mean, opt.std); R = da.prepare_image(gr, opt.mean, opt.std)
    uvz, wgt, cnt = da._sample_fixed(p2[0, 0] * 2, p2[0, 2] * 2, p2[1, 2] * 2, L.H, L.W, tb * 2, tp_, tbd * 2)
    pix = int(cnt.sum().item())
    m = int(cnt.max().item())
    uvz, wgt = uvz[:, :m].contiguous(), wgt[:, :m].contiguous()
    de = (torch.linspace(-12.5, 12, 50, device=dev).unsqueeze(1) + tp_[:, 2].unsqueeze(0)).clamp_min(1.5).contiguous()
    fbv = float(p2[0, 0] * 2 * ((p2[0, 3] - p3[0, 3]) * 2 / (p2[0, 0] * 2)))
    ms_enum = time_op(lambda: da.enumeration_depth(L, R, uvz, wgt, de, fbv), flush=flush, iters=5)
    Lp, Rp = L.planar()[0].cpu().numpy(), R.planar()[0].cpu().numpy()
    t0 = time.time()
    co.da_enum(Lp, Rp, uvz[:4].cpu().numpy(), wgt[:4].cpu().numpy(), de[:, :4].cpu().numpy(), fbv)
    cpu4 = 1000 * (time.time() - t0)
    out["dense_align"] = {"align_parallel_ms": ms, "rois": n, "valid_pixels": pix, "hypotheses": 70,
                          "enumeration_50_ms": ms_enum, "bilinear_samples_per_s": pix * 51 / (ms_enum / 1000.0),
                          "cpu_oracle_enumeration_50_ms_per_roi": cpu4 / 4,
                          "note": "align_parallel = 2 image preparations + sample + 2 enumerations (dense_align.py:240-312); "
                                  "latency-bound (a few MB touched), reported in ms"}
    # ---- F4: detector input preparation, one KITTI pair 375 x 1242 -> 384 x 1280 ----
    ol = types.SimpleNamespace(input_h=384, input_w=1280, output_h=96, output_w=320, down_ratio=4, keep_res=False,
                               mean=np.array(opt.mean, np.float32).reshape(1, 1, 3), std=np.array(opt.std, np.float32).reshape(1, 1, 3))
    kl = torch.from_numpy(rng.randint(0, 256, (375, 1242, 3), dtype=np.uint8)).to(dev)
    kr = torch.from_numpy(rng.randint(0, 256, (375, 1242, 3), dtype=np.uint8)).to(dev)
    out["pre_process_pair"] = {"us": 1000 * time_op(lambda: pp.pre_process(ol, kl, kr, None), flush=flush),
                               "note": "warpAffine + normalise + transpose of both images in one launch (stereoDetector.py:45-82); "
                                       "images resident as uint8"}
    return out


# ----------------------------------------------------------------------------------------------------
# BASELINE config #5: data-parallel training step (2 synthetic pairs per GPU by default, NCCL gradient all-reduce)
# ----------------------------------------------------------------------------------------------------
def train_setup(args, dev, rank, world):
    """BASELINE config #5: ModelWithLoss-equivalent (GT RoIs built on the device, the reference's StereoLoss terms), Adam,
    bucketed NCCL gradient all-reduce overlapped with backward.  -> (step function, reducer)"""
    from side_b200.engine import GradientAllReducer
    from side_b200.training import ModelWithLoss, StereoLoss
    from side_b200.utils.synthetic import make_batch, make_targets
    model = build_model().to(dev).train()
    B = args.train_batch
    batch = make_batch(B, H_IN, W_IN, seed=50 + rank)
    batch.update(make_targets(B, 8, max_objs=50, seed=60 + rank))        # 8 ground-truth objects per pair (SURVEY config #5)
    batch = {k: v.to(dev) for k, v in batch.items()}
    mwl = ModelWithLoss(model, StereoLoss(grid=28), output_w=W_IN // 4)
    opt = torch.optim.Adam(model.parameters(), lr=1.25e-4)
    red = GradientAllReducer(model.parameters())
    marks = []

    def step():
        red.zero_grad()
        _, loss, _ = mwl(batch)
        loss.backward()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        early = red.finish()              # waits for the buckets still in flight: the time from here is EXPOSED communication
        e1.record()
        marks.append((e0, e1, early))
        opt.step()
        return loss

    return step, red, marks


def time_train(args, dev, rank, world, steps, profiler_range=False):
    from side_b200 import _lib
    step, red, marks = train_setup(args, dev, rank, world)
    for _ in range(max(args.warmup, 5)):      # cuDNN autotuning and the caching allocator settle over the first few steps
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    del marks[:]
    _lib.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if profiler_range:
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    if profiler_range:
        torch.cuda.profiler.stop()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    exposed = sum(a.elapsed_time(b) for a, b, _ in marks) / max(len(marks), 1)
    t = torch.tensor([e0.elapsed_time(e1), exposed], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, exposed = float(t[0].item()), float(t[1].item())
    B = args.train_batch
    return {"value": world * B * steps / (ms / 1000.0), "unit": "pairs/s", "ms_per_step": ms / steps, "steps": steps,
            "pairs_per_gpu_per_step": B, "global_batch": world * B, "scaling": "weak",
            "workload": "SIDE DLA-34 stereo training step (config #5): %d pairs/GPU, 8 GT RoIs/pair built on the device, forward + StereoLoss "
                        "(focal + masked L1 + keypoint cross-entropies + L1 depth) + backward + NCCL gradient all-reduce + Adam" % B,
            "allreduce_bytes_per_step": red.bytes_per_step if world > 1 else 0, "allreduce_buckets": len(red.buckets),
            "buckets_in_flight_before_backward_ended": (sum(e for _, _, e in marks) / max(len(marks), 1)) if world > 1 else 0,
            "exposed_allreduce_ms_per_step": exposed if world > 1 else 0.0,
            "gpu_launches": _lib.launch_count(), "loss": float(loss.detach())}


def run_train(args):
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from side_b200 import ops
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = bool(args.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(args.allow_tf32)
    ops.set_dcn_precision(args.dcn_precision)
    tr = time_train(args, dev, rank, world, args.steps, args.profiler_range)
    if rank == 0:
        line = {"metric": "stereo training pairs/sec (384x1280, DLA-34)", "value": tr["value"], "unit": "pairs/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 5), "ms_per_step": tr["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
                "config": {"workload": tr["workload"], "pairs_per_gpu_per_step": args.train_batch,
                           "parallelism": "data-parallel x%d, bucketed all_reduce overlapped with backward" % world,
                           "dcn_precision": args.dcn_precision},
                "gpu_launches": tr["gpu_launches"], "train": tr}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.mode == "train":
        return run_train(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    from side_b200 import _lib, ops
    from side_b200.engine import OpTimer, StereoDetector, gather_detections
    from side_b200.utils.synthetic import KITTI_FB

    lib = _lib.load()
    if lib.side_device_ok() != 1:
        raise RuntimeError("libside_b200.so cannot run on this device (needs sm_100a)")
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = bool(args.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(args.allow_tf32)
    ops.set_dcn_precision(args.dcn_precision)
    lib.side_conv_tc_set_mode(args.conv_mode)
    pk = peaks()

    cpu_model = build_model()
    cpu_base, cpu_kept = None, []
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            torch.set_num_threads(max(len(os.sched_getaffinity(0)), 1))
        except AttributeError:
            pass
        cpu_kept = []
        v, times = cpu_reference_pairs_per_s(cpu_model, args.cpu_pairs, 1, keep=cpu_kept)
        cpu_base = {"value": v, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
                    "sample": "%d pairs (batch 1, 384x1280, K=100 RoIs) after 1 warm-up; reference-style torch CPU ops "
                              "(torchvision deform_conv2d / roi_align loop, torch.topk decode)" % args.cpu_pairs}
    model = cpu_model.to(dev)
    if args.cudnn_only:
        model.heads_tensor_core = False
        model.depth_estimator.tensor_core = False
    ops.set_tc_format(args.tc_format)
    det = StereoDetector(model, grid_size=28, K=100)

    P, mb = args.pairs, args.micro_batch
    assert P % mb == 0, "--pairs must be a multiple of --micro-batch"
    g = torch.Generator().manual_seed(1234 + rank)
    host_l = torch.randn(P, 3, H_IN, W_IN, generator=g).pin_memory()
    host_r = torch.randn(P, 3, H_IN, W_IN, generator=g).pin_memory()
    dev_l, dev_r = host_l.to(dev), host_r.to(dev)
    fb = torch.full((mb,), KITTI_FB, device=dev)
    out_host = [torch.empty((P, 100, 6)).pin_memory(), torch.empty((P, 100, 6)).pin_memory(), torch.empty((P, 100, 10)).pin_memory()]

    def step_resident(pairs=None, micro=None):
        pairs, micro = pairs or P, micro or mb
        outs = []
        for i in range(0, pairs, micro):
            outs.append(det.process({'input': dev_l[i:i + micro], 'input_right': dev_r[i:i + micro], 'fb': fb[:micro]}))
        d, dr, info = (torch.cat([o[j] for o in outs], 0) for j in range(3))
        return gather_detections(d, dr, info)

    # e2e: pinned host buffers -> double-buffered device staging on a copy stream, so the H2D of micro-batch i+1 overlaps the
    # compute of micro-batch i (every byte still crosses PCIe inside the timed region); detections go back to pinned memory
    stage = [[torch.empty((mb, 3, H_IN, W_IN), device=dev) for _ in range(2)] for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    staged = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def upload(i, slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])            # the step that last read this slot has finished with it
            stage[slot][0].copy_(host_l[i:i + mb], non_blocking=True)
            stage[slot][1].copy_(host_r[i:i + mb], non_blocking=True)
            staged[slot].record(copy_stream)

    def step_e2e():
        outs = []
        cur = torch.cuda.current_stream(dev)
        for k in range(2):
            consumed[k].record(cur)
        upload(0, 0)
        for n, i in enumerate(range(0, P, mb)):
            slot = n & 1
            if i + mb < P:
                upload(i + mb, slot ^ 1)
            cur.wait_event(staged[slot])
            outs.append(det.process({'input': stage[slot][0], 'input_right': stage[slot][1], 'fb': fb}))
            consumed[slot].record(cur)
        d, dr, info = (torch.cat([o[j] for o in outs], 0) for j in range(3))
        res = gather_detections(d, dr, info)
        for dst, src in zip(out_host, (d, dr, info)):
            dst.copy_(src, non_blocking=True)
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step_fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_resident()
    step_e2e()
    barrier()

    clocks = ClockSampler(local)
    clocks.start()
    _lib.launch_count(reset=True)

    def dcn_work(out, x, *a, **k):
        w = a[2]
        return 2.0 * out.numel() * w.shape[1] * w.shape[2] * w.shape[3]

    conv_desc = []                                             # one entry per timed convolution launch (SIDE_BENCH_CONV_TABLE)

    def conv_work(out, x_hi, x_lo, wp, Cout, ksize=(3, 3, 3), **k):
        cin = getattr(wp, "cin_alg", x_hi.shape[-1])          # zero-padded input channels (96 -> 128 for fp16) do not count
        conv_desc.append("%s -> %d k%s s%s%s%s%s" % ("x".join(str(v) for v in x_hi.shape), Cout, "x".join(str(v) for v in ksize),
                                                  k.get("stride", 1), " +res" if k.get("residual") is not None else "",
                                                  " full" if k.get("full", True) else "", " split" if k.get("split") else ""))
        return 2.0 * (x_hi.numel() // x_hi.shape[-1]) * Cout * cin * ksize[0] * ksize[1] * ksize[2]

    def conv_fmt(x_hi, *a, **k):
        return "f16" if x_hi.dtype == torch.float16 else "tf32"

    if args.profiler_range:
        torch.cuda.profiler.start()
    def dcn_cl_work(out, x_nhwc, om_cl, weight, *a, **k):
        return 2.0 * out.numel() * weight.shape[1] * weight.shape[2] * weight.shape[3]

    def vol_work(out, featL, featR, left, right, fb_, D, P_, *a, **k):
        # algorithmic bytes of the volume builder: the pairs written once (hi + lo fp16 = 4 B per element) + both feature maps read once
        return float(out[0].numel() * 4 + 2 * featL.numel() * 4)

    ops.tc_range_status(dev)                                  # clear: the flags below cover exactly the timed regions
    with OpTimer("dcn_forward_raw", dcn_work) as tm, OpTimer("dcn_fwd_cl", dcn_cl_work) as tm2, \
            OpTimer("conv3d_tc", conv_work, conv_fmt) as tmc, OpTimer("inst_costvol_cl", vol_work) as tmv:
        ms_res = timed(step_resident, args.steps)
    if args.profiler_range:
        torch.cuda.profiler.stop()
    dcn, dcn2 = tm.summary(), tm2.summary()
    dcn = {k: dcn[k] + dcn2[k] for k in ("calls", "ms", "work")}       # NCHW entry + channels-last entry of the same kernel
    cv, cv16 = tmc.summary("tf32"), tmc.summary("f16")
    cv_big, cv_small = tmc.split_by_work(1e11, args.tc_format)      # >= 100 GFLOP per launch: aggregation layers, head convolutions
    vol = tmv.summary()
    if os.environ.get("SIDE_BENCH_CONV_TABLE") and rank == 0:
        # per-shape table of the convolution launches of the timed region (profiles/): calls, ms per step, TFLOP/s
        tab = {}
        for (e0, e1, w, _), d in zip(tmc.records, conv_desc):
            t = tab.setdefault(d, [0, 0.0, 0.0])
            t[0] += 1; t[1] += e0.elapsed_time(e1); t[2] += w
        with open(os.environ["SIDE_BENCH_CONV_TABLE"], "w") as f:
            f.write("| input (N x D x H x W x C) -> Cout | calls / step | ms / step | TFLOP/s |\n|---|---|---|---|\n")
            for d, (n, ms_, w) in sorted(tab.items(), key=lambda kv: -kv[1][1]):
                f.write("| %s | %.0f | %.3f | %.1f |\n" % (d, n / args.steps, ms_ / args.steps, w / ms_ / 1e9 if ms_ > 0 else 0.0))
    launches = _lib.launch_count(reset=True)
    ms_e2e = timed(step_e2e, args.steps)
    clk = clocks.stop()
    range_flags = ops.tc_range_status(dev)

    # ---- the same step under the other operand format (strict 3xTF32 next to 3xFP16), a short run ----
    fmt_values = {args.tc_format: world * P * args.steps / (ms_res / 1000.0)}
    if not args.cudnn_only:
        other_fmt = "tf32" if args.tc_format == "f16" else "f16"
        ops.set_tc_format(other_fmt)
        if args.dcn_precision == "3xfp16":                 # the two lines are "everything on fp16 pairs" / "everything on tf32 pairs"
            ops.set_dcn_precision("3xtf32" if other_fmt == "tf32" else "3xfp16")
        step_resident()
        n_o = max(1, min(args.steps, 2))
        fmt_values[other_fmt] = world * P * n_o / (timed(step_resident, n_o) / 1000.0)
        ops.set_tc_format(args.tc_format)
        ops.set_dcn_precision(args.dcn_precision)
        step_resident()

    # ---- BASELINE config #4 as written: 32 pairs TOTAL over the ranks (strong scaling) ----
    strong = None
    if 32 % world == 0 and 32 // world <= P:
        ps = 32 // world
        ms_s = min(mb, ps)
        if ps == P and ms_s == mb:
            strong_ms = ms_res / args.steps
        else:
            step_resident(ps, ms_s)
            strong_ms = timed(lambda: step_resident(ps, ms_s), args.steps) / args.steps
        strong = {"value": 32 / (strong_ms / 1000.0), "unit": "pairs/s", "scaling": "strong", "pairs_total": 32, "pairs_per_rank": ps,
                  "micro_batch": ms_s, "ms_per_step": strong_ms, "steps": args.steps,
                  "workload": "SIDE DLA-34 stereo inference (config #4): 32 pairs in total, contiguous blocks of 32 / N per rank"}

    # ---- BASELINE config #1's GPU counterpart: ONE stereo pair (the batch the real detector runs, stereoDetector.py:84-103) ----
    latency = None
    if world == 1 and not args.no_latency:
        det1 = StereoDetector(model, grid_size=28, K=100)
        b1 = {'input': dev_l[:1].clone(), 'input_right': dev_r[:1].clone(), 'fb': fb[:1].clone()}
        for _ in range(3):
            det1.process(b1)
        n0 = _lib.launch_count()
        eager_ms = timed(lambda: det1.process(b1), 10) / 10
        per_step = (_lib.launch_count() - n0) // 10
        det1.capture(b1)
        graph_ms = timed(lambda: det1.replay(b1), 20) / 20
        latency = {"pairs": 1, "eager_ms": eager_ms, "cuda_graph_ms": graph_ms, "library_launches_per_step": per_step,
                   "note": "one 384x1280 pair, inputs resident: DLA levels, heads and DCN offset convolutions on tcgen05 at batch 1 "
                           "(depth boxes of the 12x40 level hang over the 2-image batch), DCN forward split-K over the taps; whole "
                           "detector step replayed as one CUDA graph"}

    # ---- parity of the timed configuration against the CPU port on the pairs the CPU baseline just processed ----
    par = None
    if cpu_kept:
        from oracle import parity
        nref = min(len(cpu_kept), mb)
        il, ir = dev_l[:mb].clone(), dev_r[:mb].clone()
        for j in range(nref):
            il[j].copy_(cpu_kept[j][0]['input'][0]); ir[j].copy_(cpu_kept[j][0]['input_right'][0])
        det.keep_heads = True
        det.process({'input': il, 'input_right': ir, 'fb': fb})
        det.keep_heads = False
        zg = {k: t[:nref].float().cpu() for k, t in det.last_heads.items()}
        zc = {k: torch.cat([kept[1][k] for kept in cpu_kept[:nref]], 0) for k in cpu_kept[0][1]}
        par = parity.compare(zg, zc, K=100, heat_is_logit=False)
        par["note"] = ("GPU arm exactly as timed (micro-batch %d, %s convs, DCN %s) vs the CPU port on the same weights / inputs; "
                       "heads: max|a-b| / max|b| per head map" % (mb, args.tc_format, args.dcn_precision))
        det.last_heads = None

    # ---- BASELINE config #5 in the same process ----
    train = None
    if not args.no_train:
        torch.cuda.empty_cache()
        train = time_train(args, dev, rank, world, max(3, min(args.steps, 10)))

    total_pairs = world * P * args.steps
    value = total_pairs / (ms_res / 1000.0)
    e2e = total_pairs / (ms_e2e / 1000.0)
    traffic = {}
    for name in ("r1_traffic.json", "r2_traffic.json"):            # dram bytes per launch from the committed ncu --set full captures
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            traffic.update(json.load(open(tpath)))

    def roof(name, label, summ, f16=False):
        tf = summ["work"] / (summ["ms"] / 1000.0) / 1e12 if summ["ms"] > 0 else 0.0
        note = ("3xFP16: fp16 MMAs run at the bf16 rate and every product takes 3 of them, so 1/3 = 0.333 is the ceiling of this "
                "fraction" if f16 else
                "3xTF32: tf32 rate is half of bf16 and every product takes 3 MMAs, so 1/6 = 0.167 is the ceiling of this fraction")
        return {"kernel": "%s, %d launches in the timed region" % (label, summ["calls"]), "bound": "tensor", "achieved": tf,
                "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": tf / pk["tf_sust"],
                "traffic": (traffic.get(name) or {}).get("bytes"), "traffic_detail": traffic.get(name),
                "peak_source": pk["src"] + " bf16 sustained (cuBLAS); this kernel is " + note,
                "share_of_step": summ["ms"] / ms_res, "algorithmic_flop_per_step": summ["work"] / max(args.steps, 1)}

    r_dcn = roof("dcn_fwd_tc_kernel", "dcn_fwd (%s)" % args.dcn_precision, dcn)
    label = ("conv3d_tc %s (every convolution of the step except the 7x7 base layer: aggregation network, heads incl. 1x1 outputs, "
             "DLA levels 0-5 incl. the stem block convolutions, DCN offset convs, strAM, feaRuduce: conv_tct_kernel / conv_tc_kernel)")
    r_cv = roof("conv_tc_kernel", label % "3xtf32", cv)
    r_cv16 = roof("conv_tc_kernel_f16", label % "3xfp16, kind::f16", cv16, f16=True)
    for r in (r_cv16 if args.tc_format == "f16" else r_cv,):
        for name, summ in (("launches_100gflop_and_more", cv_big), ("smaller_launches", cv_small)):
            tf = summ["work"] / (summ["ms"] / 1000.0) / 1e12 if summ["ms"] > 0 else 0.0
            r[name] = {"calls": summ["calls"], "achieved": tf, "frac": tf / pk["tf_sust"], "share_of_step": summ["ms"] / ms_res,
                       "note": "aggregation network and head convolutions" if name.startswith("launches") else
                               "DLA levels 2-5, stem block convolutions (algorithmic FLOPs of the 16-channel pixel-domain layers), "
                               "DCN offset convolutions, 1x1 outputs, strAM / feaRuduce"}
    ranked = sorted([r_cv, r_cv16, r_dcn], key=lambda r: -r["share_of_step"])
    dominant, other, third = ranked
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32" if not args.allow_tf32 else "fp32 (cuDNN convs TF32)", "data": "synthetic",
            "config": {"workload": "SIDE DLA-34 stereo inference (config #4), %d pairs/GPU/step, micro-batch %d, 384x1280, K=100 RoIs/pair, "
                                   "random-init weights with calibrated BatchNorm" % (P, mb),
                       "pairs_per_gpu_per_step": P, "micro_batch": mb, "parallelism": "pair-sharded x%d, all_gather(detections)" % world,
                       "dcn_precision": args.dcn_precision, "cudnn_tf32": bool(args.allow_tf32),
                       "tensor_core_convs": "cuDNN fp32 only" if args.cudnn_only else
                       "3-D aggregation network incl. the strAM gate convolution, heads incl. their 1x1 outputs, feaRuduce, DLA-34 levels "
                       "0-5 (levels 0 / 1 as 2x2 space-to-depth block convolutions) and the DCN offset convolutions on tcgen05 %s "
                       "(fp32-class accuracy: 22-bit operands, fp32 accumulation, <= 1e-4 rel.); the 7x7 base layer as a direct fp32 "
                       "SIMT convolution; no cuDNN / cuBLAS kernel in the step" % ("3xFP16 (kind::f16)" if args.tc_format == "f16" else "3xTF32"),
                       "l2": "inputs larger than L2 (%.0f MB of images per step)" % (2 * P * 3 * H_IN * W_IN * 4 / 1e6)},
            "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": 2 * P * 3 * H_IN * W_IN * 4,
                    "d2h_bytes_per_step": P * 100 * 22 * 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": dominant,
            "roofline_second": other,
            "roofline_third": third,
            "cost_volume": None if not vol["calls"] else {
                "kernel": "inst_costvol_cl_kernel (gated channels-last fp16 pairs), %d launches in the timed region" % vol["calls"],
                "bound": "hbm", "achieved": vol["work"] / (vol["ms"] / 1000.0) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                "frac": vol["work"] / (vol["ms"] / 1000.0) / 1e9 / pk["hbm"],
                "traffic": (traffic.get("inst_costvol_cl_kernel") or {}).get("bytes"), "traffic_detail": traffic.get("inst_costvol_cl_kernel"),
                "ms_per_launch": vol["ms"] / vol["calls"], "algorithmic_bytes_per_launch": vol["work"] / vol["calls"],
                "share_of_step": vol["ms"] / ms_res, "peak_source": pk["src"] + " copy bandwidth"},
            "tc_formats": {("3xfp16" if k == "f16" else "3xtf32") + "_pairs_per_s": v for k, v in fmt_values.items()},
            "f16_range": {"saturated": bool(range_flags & ops.TC_RANGE_SATURATED), "underflow": bool(range_flags & ops.TC_RANGE_UNDERFLOW),
                          "note": "fp16 operand-pair range guard over the timed regions; a set flag means the step must rerun in 3xTF32 "
                                  "(StereoDetector.process_checked)"},
            "parity": par,
            "latency": latency,
            "strong": strong,
            "train": train,
            "cpu_baseline": cpu_base,
        }
        if not args.no_kernels and world == 1:
            line["kernels"] = kernel_rooflines(pk, args.dcn_precision)
            if line["cost_volume"] is not None:
                # BASELINE config #2 (64 RoIs x 48 candidates x 64 channels, reference fp32 layout) beside the in-step builder:
                # stand-alone launches with the L2 flushed, CUDA events around the whole call (see `kernels`)
                kk = line["kernels"]
                line["cost_volume"]["config2"] = {
                    "forward_separable": kk.get("inst_costvol_fwd_separable"), "forward_exact": kk.get("inst_costvol_fwd_exact"),
                    "backward": kk.get("inst_costvol_bwd"), "peak": pk["hbm"], "unit": "GB/s"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
