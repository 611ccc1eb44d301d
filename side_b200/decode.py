"""Drop-in ``bbox_decode`` / ``ddd_decode`` (reference ``models/decode.py:35-126``) on the single-launch
NMS + top-K + gather kernel (``side_bbox_decode`` / ``side_ddd_decode``).

Differences that are deliberate and documented:
  * ties are ordered by lowest flat index (torch.topk leaves them unspecified, Q2);
  * ``kept_type`` is ``floor(argmax / grid)`` -- the reference relied on torch-0.4 integer division (Q1);
  * nothing is staged through the CPU (Q5); ``bbox_decode`` still has to synchronise once because it
    returns a data-dependent number of rows -- the fused network uses ``ops.bbox_decode_raw`` instead.
"""
import torch

from . import ops


def bbox_decode(heat, wh, reg, K=100):
    """heat: pre-sigmoid heat-map [B,cat,H,W]; wh, reg [B,3,H,W]
    -> (bbox_keep [M,5], bbox_right_keep [M,5], torch.Size([B,K,5]))   (decode.py:91-126)"""
    o = ops.bbox_decode_raw(heat, wh, reg, K=K, wh_scale=1.0, heat_is_logit=True)
    keep = o["keep"].bool()
    bbox = o["bbox"].view(-1, 5)
    bbox_right = o["bbox_right"].view(-1, 5)
    return bbox[keep], bbox_right[keep], o["bbox"].shape


def ddd_decode(heat, kept, dim, orien, wh, reg, grid_size, K=40):
    """heat: post-sigmoid heat-map -> (detections [B,K,6], detections_right [B,K,6], info_3d [B,K,9])
    (decode.py:35-89)"""
    return ops.ddd_decode_raw(heat, kept, dim, orien, wh, reg, grid_size, K=K, heat_is_logit=False)
