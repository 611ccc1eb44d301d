"""TEST INFRASTRUCTURE ONLY -- end-to-end agreement between two runs of the stereo network on the same inputs
(product on the GPU vs the reference-style port on the CPU, ``oracle/torch_port.py``).

Used by ``tests/test_gpu_e2e.py`` and by ``bench.py``'s cpu_baseline leg (as the checker of the numbers it prints, never as
the thing measured).  The detections of the two runs are matched by (image, class, flat peak index) -- the keys the
reference's ``_topk`` (models/decode.py:17-33) produces -- so a pair of near-tied scores that swaps ranks does not count as
a disagreement, while a peak that only one side found does.
"""
import numpy as np
import torch

from . import torch_port

HEADS = ("hm", "wh", "reg", "dim", "orien", "kept_type")


def _keys(z, K, heat_is_logit):
    """-> per image {(cls, ind): depth slot or None}: rows of bbox_decode (decode.py:91-126) in score order; the slot is the row's
    rank among the kept boxes, which is where stereo_network.forward (:378-381) stores that RoI's depth."""
    o = torch_port.bbox_decode_raw(z["hm"].float().cpu(), z["wh"].float().cpu(), z["reg"].float().cpu(), K=K,
                                   heat_is_logit=heat_is_logit)
    B = z["hm"].shape[0]
    keep = o["keep"].view(B, K).bool()
    out = []
    for b in range(B):
        d = {}
        for k in range(K):
            d[(int(o["cls"][b, k]), int(o["ind"][b, k]))] = int(o["slot"][b, k]) if bool(keep[b, k]) else None
        out.append(d)
    return out


def compare(z_test, z_ref, K=100, heat_is_logit=True):
    """z_*: the dict stereo_network.forward returns (heads [B,c,H,W] + depth [B,K,1]).  ``heat_is_logit=False`` when 'hm' already
    went through the detector's in-place sigmoid.  Returns plain floats:
      heads_max_err         max over heads of max|a-b| / max|b|  (error relative to the head's range)
      topk_agreement        fraction of the K (class, peak index) keys per image found by both runs
      depth_max_rel / depth_median_rel / depth_frac_1e-3   |d_a - d_b| / d_b over the matched, kept detections
    """
    res = {"pairs": int(z_ref["hm"].shape[0])}
    errs = {}
    for h in HEADS:
        if h in z_ref and h in z_test:
            a, b = z_test[h].detach().float().cpu().numpy(), z_ref[h].detach().float().cpu().numpy()
            errs[h] = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
    res["heads_err"] = errs
    res["heads_max_err"] = max(errs.values()) if errs else None
    ka, kb = _keys(z_test, K, heat_is_logit), _keys(z_ref, K, heat_is_logit)
    agree, rel = [], []
    da = z_test["depth"].detach().float().cpu().numpy() if "depth" in z_test else None
    db = z_ref["depth"].detach().float().cpu().numpy() if "depth" in z_ref else None
    for b, (a, r) in enumerate(zip(ka, kb)):
        common = set(a) & set(r)
        agree.append(len(common) / float(K))
        if da is None or db is None:
            continue
        for key in common:
            sa, sr = a[key], r[key]
            if sa is None or sr is None:
                continue
            ref = float(db[b, sr, 0])
            if ref != 0.0:
                rel.append(abs(float(da[b, sa, 0]) - ref) / abs(ref))
    res["topk_agreement"] = float(np.mean(agree))
    res["topk_agreement_min"] = float(np.min(agree))
    if rel:
        rel = np.asarray(rel)
        res.update(depth_matched=int(rel.size), depth_max_rel=float(rel.max()), depth_median_rel=float(np.median(rel)),
                   depth_frac_within_1e3=float((rel <= 1e-3).mean()))
    return res
