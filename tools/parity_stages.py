"""Where does the end-to-end difference to the CPU port come from?  Per-stage max|a-b|/max|b| of (a) the GPU fast path,
(b) the GPU with every fast path switched off (cuDNN fp32 + SIMT DCN), (c) the CPU fp32 port -- each against the CPU port
evaluated in float64.  Run on the GPU box: python tools/parity_stages.py [n_pairs]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import torch_port  # noqa: E402
from side_b200 import ops  # noqa: E402
from side_b200.networks import get_pose_net  # noqa: E402
from side_b200.utils.synthetic import HEADS, make_batch, realistic_init  # noqa: E402


def capture(model, batch):
    """-> dict stage -> tensor (first n_ref samples, CPU float64)."""
    out = {}
    fe = model.feature_extraction
    hooks = [fe.base.register_forward_hook(lambda m, i, o: out.update({"base.level%d" % k: t for k, t in enumerate(o)})),
             fe.register_forward_hook(lambda m, i, o: out.update({"features": o})),
             model.feaRuduce.register_forward_hook(lambda m, i, o: out.setdefault("feaRuduce", o))]
    with torch.no_grad():
        z = model(batch, True, None, 1.0)[0]
    for h in hooks:
        h.remove()
    out.update(z)
    return out


def main():
    n_ref = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    model = realistic_init(get_pose_net(34, HEADS, 256), seed=1).eval()
    batch = make_batch(8, 384, 1280, seed=1234)
    small = {k: v[:n_ref] for k, v in batch.items()}
    t0 = time.time()
    with torch_port.reference_ops():
        ref64 = capture(model.double(), {k: v.double() for k, v in small.items()})
    print("fp64 port: %.0f s" % (time.time() - t0), flush=True)
    model = model.float()
    with torch_port.reference_ops():
        cpu32 = capture(model, small)
    m = model.cuda()
    cb = {k: v.cuda() for k, v in batch.items()}
    ops.set_tc_format("f16"); ops.set_dcn_precision("3xtf32")
    fast = capture(m, cb)
    ops.set_tc_format("tf32")
    fast_tf32 = capture(m, cb)
    ops.set_tc_format("f16")
    ops.set_dcn_precision("fp32")
    m.heads_tensor_core = False; m.depth_estimator.tensor_core = False
    m.feature_extraction.base.tensor_core = False; m.feature_extraction.base.direct_stem = False
    plain = capture(m, cb)

    def err(a, b, n):
        a = a.detach().double().cpu()
        if a.dim() == 4 and a.shape[0] == 16:      # batched [left; right] base output: left half first
            a = torch.cat((a[:n], a[8:8 + n]), 0)
            b = b  # the port runs left and right separately only in training; eval batches them the same way
        else:
            a = a[:n]
        b = b.detach().double().cpu()
        if a.shape != b.shape:
            return float("nan")
        return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

    print("%-16s %12s %12s %12s %12s" % ("stage", "gpu f16+3xtf32", "gpu tf32 fmt", "gpu cudnn fp32", "cpu fp32 port"))
    for k in ref64:
        if k == "depth":
            continue
        print("%-16s %12.2e %12.2e %12.2e %12.2e" % (k, err(fast[k], ref64[k], n_ref), err(fast_tf32[k], ref64[k], n_ref),
                                                     err(plain[k], ref64[k], n_ref), err(cpu32[k], ref64[k], n_ref)))


if __name__ == "__main__":
    main()
