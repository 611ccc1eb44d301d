"""Dense photometric alignment on the GPU -- drop-in for ``src/lib/dense_align/dense_align.py``.

Same function names, argument meaning and return values as the reference:

* ``sample(calib, scale, f_h, f_w, box_left, poses, borders)``           (dense_align.py:14-70)
* ``enumeration_depth(im_left, im_right, all_uvz, all_weight, depth_enum, fb)``   (:175-237)
* ``align_parallel(calib, opt, im_left, im_right, box_left, borders, poses)``      (:240-312)

What changes is where the work runs.  The reference normalises the images in numpy, loops over RoIs in Python for
``sample`` (8 ``torch.mm`` + masked assignments per RoI) and materialises ``(iter x rois) x pixels`` grids for two
``F.grid_sample`` calls per enumeration.  Here every stage is one kernel of ``libside_b200.so``
(``csrc/dense_align.cu``): image preparation writes packed float4 texels, ``sample`` is one CTA per RoI, and the
enumeration is a gather-and-reduce kernel that never builds the expanded grids.  ``align_parallel`` has no host
synchronisation (the reference's ``if torch.sum(all_weight) == 0`` early exit becomes a ``torch.where``).

``grid_sample`` convention: the reference was written for torch < 1.3 (``align_corners=True`` default) but, executed
under the torch in this image, computes ``align_corners=False``.  ``ALIGN_CORNERS`` (module flag, default False =
what the reference computes today) selects the convention; both are tested.
"""
import numpy as np
import torch

from . import _lib
from .ops import _chk, _stream

ALIGN_CORNERS = False
# Pixels per RoI: the slice steps are max(int(extent / 56), 1), so a RoI yields < 46 x 113 samples for any box that
# fits the image; 8192 leaves slack.  sample() checks the device-side counts against it.
SAMPLE_CAP = 8192


class PackedImage:
    """[H][W] float4 texels (c0, c1, c2, 0) in device memory -- the layout the enumeration kernel gathers from."""

    def __init__(self, data, H, W):
        self.data, self.H, self.W = data, H, W

    def planar(self):
        """Back to the reference's 1 x 3 x H x W tensor."""
        out = torch.empty((1, 3, self.H, self.W), device=self.data.device, dtype=torch.float32)
        _lib.check(_lib.load().side_dense_align_unpack(self.data.data_ptr(), out.data_ptr(), self.H, self.W, _stream()),
                   "side_dense_align_unpack")
        return out


def prepare_image(img, mean, std, device=None):
    """align_parallel's image preparation (dense_align.py:251-266): uint8 HWC (numpy or tensor) -> ((x / 255) - mean)
    / std -> 2x bilinear up-sampling -> PackedImage of size 2H x 2W.  One kernel; the only host work is the upload."""
    if isinstance(img, np.ndarray):
        img = torch.from_numpy(np.ascontiguousarray(img))
    if device is not None:
        img = img.to(device, non_blocking=True)
    img = _chk(img, "image", torch.uint8)
    H, W, C = img.shape
    if C != 3:
        raise RuntimeError("side_b200.dense_align: image must be H x W x 3")
    out = torch.empty((2 * H, 2 * W, 4), device=img.device, dtype=torch.float32)
    m = (_lib._f * 3)(*[float(np.float32(v)) for v in mean])
    s = (_lib._f * 3)(*[float(np.float32(v)) for v in std])
    _lib.check(_lib.load().side_dense_align_prep_u8(img.data_ptr(), out.data_ptr(), H, W, m, s, _stream()),
               "side_dense_align_prep_u8")
    return PackedImage(out, 2 * H, 2 * W)


def _packed(im):
    if isinstance(im, PackedImage):
        return im
    im = _chk(im, "image")
    if im.dim() != 4 or im.shape[0] != 1 or im.shape[1] != 3:
        raise RuntimeError("side_b200.dense_align: image must be 1 x 3 x H x W")
    H, W = int(im.shape[2]), int(im.shape[3])
    out = torch.empty((H, W, 4), device=im.device, dtype=torch.float32)
    _lib.check(_lib.load().side_dense_align_pack(im.data_ptr(), out.data_ptr(), H, W, _stream()), "side_dense_align_pack")
    return PackedImage(out, H, W)


def _sample_fixed(f, cx, cy, f_h, f_w, box_left, poses, borders, cap=SAMPLE_CAP):
    box_left = _chk(box_left, "box_left")
    poses = _chk(poses, "poses")
    borders = _chk(borders, "borders")
    rois = int(box_left.shape[0])
    dev = box_left.device
    uvz = torch.empty((rois, cap, 3), device=dev, dtype=torch.float32)
    weight = torch.empty((rois, cap), device=dev, dtype=torch.float32)
    count = torch.empty((rois,), device=dev, dtype=torch.int32)
    _lib.check(_lib.load().side_dense_align_sample(box_left.data_ptr(), borders.data_ptr(), poses.data_ptr(), rois, float(f),
                                                   float(cx), float(cy), int(f_h), int(f_w), cap, uvz.data_ptr(),
                                                   weight.data_ptr(), count.data_ptr(), _stream()),
               "side_dense_align_sample")
    return uvz, weight, count


def sample(calib, scale, f_h, f_w, box_left, poses, borders):
    """Reference signature and result: ``all_uvz`` rois x pixels x 3, ``all_weight`` rois x pixels, ``pixels`` = the
    largest per-RoI count (one device -> host read, as many as the reference does per RoI)."""
    f = calib.p2[0, 0] * scale
    cx, cy = calib.p2[0, 2] * scale, calib.p2[1, 2] * scale
    uvz, weight, count = _sample_fixed(f, cx, cy, f_h, f_w, box_left, poses, borders)
    m = int(count.max().item()) if count.numel() else 0
    if m > SAMPLE_CAP:
        raise RuntimeError("side_b200.dense_align.sample: %d pixels in one RoI exceed the capacity %d" % (m, SAMPLE_CAP))
    return uvz[:, :m].contiguous(), weight[:, :m].contiguous()


def enumeration_depth(im_left, im_right, all_uvz, all_weight, depth_enum, fb, return_error=False):
    """Reference signature; images are 1 x 3 x H x W tensors or PackedImage.  Returns ``best_depth`` [rois]
    (``return_error=True`` adds the iter x rois photometric error and the argmin index)."""
    L, R = _packed(im_left), _packed(im_right)
    if (L.H, L.W) != (R.H, R.W):
        raise RuntimeError("side_b200.dense_align: left / right image sizes differ")
    all_uvz = _chk(all_uvz, "all_uvz")
    all_weight = _chk(all_weight, "all_weight")
    depth_enum = _chk(depth_enum, "depth_enum")
    iters, rois = int(depth_enum.shape[0]), int(depth_enum.shape[1])
    if all_uvz.shape[0] != rois or all_weight.shape[0] != rois or all_uvz.shape[1] != all_weight.shape[1]:
        raise RuntimeError("side_b200.dense_align: all_uvz / all_weight / depth_enum shapes do not agree")
    pixels = int(all_weight.shape[1])
    dev = depth_enum.device
    err = torch.empty((iters, rois), device=dev, dtype=torch.float32)
    best = torch.empty((rois,), device=dev, dtype=torch.float32)
    idx = torch.empty((rois,), device=dev, dtype=torch.int32)
    flags = _lib.DA_ALIGN_CORNERS if ALIGN_CORNERS else 0
    _lib.check(_lib.load().side_dense_align_enum(L.data.data_ptr(), R.data.data_ptr(), all_uvz.data_ptr(),
                                                 all_weight.data_ptr(), depth_enum.data_ptr(), float(fb), rois, pixels, iters,
                                                 L.H, L.W, flags, err.data_ptr(), best.data_ptr(), idx.data_ptr(), _stream()),
               "side_dense_align_enum")
    return (best, err, idx) if return_error else best


def align_parallel(calib, opt, im_left, im_right, box_left, borders, poses):
    """Reference signature (dense_align.py:240-312): raw uint8 H x W x 3 images (numpy or CUDA tensors), ``box_left``
    rois x 4, ``borders`` rois x 2, ``poses`` rois x 7 on the GPU -> ``(solve_status, best_dis)``."""
    dev = box_left.device
    L = im_left if isinstance(im_left, PackedImage) else prepare_image(im_left, opt.mean, opt.std, dev)
    R = im_right if isinstance(im_right, PackedImage) else prepare_image(im_right, opt.mean, opt.std, dev)
    scale = 2
    f = calib.p2[0, 0] * scale
    bl = (calib.p2[0, 3] - calib.p3[0, 3]) * scale / f
    cx, cy = calib.p2[0, 2] * scale, calib.p2[1, 2] * scale
    box_left = box_left * scale
    borders = borders * scale
    dis_init = f * bl / poses[:, 2]

    all_uvz, all_weight, _ = _sample_fixed(f, cx, cy, L.H, L.W, box_left, poses, borders)
    rois = int(box_left.shape[0])
    solve_status = (all_weight.sum(1) != 0).to(box_left.dtype)
    if rois == 0:
        return solve_status, dis_init

    # initial enumeration: the reference's expressions, on device tensors (same float32 operations, no Python loop)
    iter_num, depth_interval = 50, 0.5
    steps = torch.arange(iter_num, device=dev, dtype=torch.float32).unsqueeze(1)
    depth_enum = (dis_init.reciprocal() * f * bl - iter_num * depth_interval / 2).unsqueeze(0) + depth_interval * steps
    depth_enum = torch.where(depth_enum < 1.5, torch.full_like(depth_enum, 1.5), depth_enum).contiguous()
    best_depth = enumeration_depth(L, R, all_uvz, all_weight, depth_enum, f * bl)

    tune_num = 20
    tune_interval = depth_interval * 2.0 / tune_num
    tsteps = torch.arange(tune_num, device=dev, dtype=torch.float64).unsqueeze(1) * tune_interval
    tune_enum = ((best_depth - tune_num * tune_interval / 2).unsqueeze(0) + tsteps.to(torch.float32)).contiguous()
    best_depth = enumeration_depth(L, R, all_uvz, all_weight, tune_enum, f * bl)

    best_dis = f * bl / (best_depth * scale) + 0.5
    # reference: `if torch.sum(all_weight) == 0: return solve_status(zeros), dis_init`
    none = all_weight.sum() == 0
    return solve_status, torch.where(none, dis_init.to(best_dis.dtype), best_dis)
