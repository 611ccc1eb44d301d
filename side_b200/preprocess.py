"""Detector input preparation on the GPU -- drop-in for ``stereoDetector.pre_process`` (modules/stereoDetector.py:45-82)
and ``get_affine_transform`` (utils/image.py:27-60).

The reference warps, normalises and transposes both images on the host (cv2 + numpy) and uploads two float tensors
(2 x 5.9 MB per pair at 384 x 1280).  Here the raw uint8 images are uploaded (2 x 1.4 MB for a KITTI frame) and ONE kernel
(``side_preprocess_u8``, csrc/preprocess.cu) produces both normalised CHW tensors.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .ops import _chk, _stream


def _get_dir(src_point, rot_rad):
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    return [src_point[0] * cs - src_point[1] * sn, src_point[0] * sn + src_point[1] * cs]


def _third(a, b):
    d = a - b
    return b + np.array([-d[1], d[0]], dtype=np.float32)


def _affine_from_points(src, dst):
    """cv2.getAffineTransform: the 2x3 map taking three source points to three destination points (solved in double)."""
    a = np.zeros((6, 6), np.float64)
    b = np.zeros((6,), np.float64)
    for i in range(3):
        a[i, 0:2], a[i, 2] = src[i], 1.0
        a[i + 3, 3:5], a[i + 3, 5] = src[i], 1.0
        b[i], b[i + 3] = dst[i, 0], dst[i, 1]
    return np.linalg.solve(a, b).reshape(2, 3)


def get_affine_transform(center, scale, rot, output_size, shift=np.array([0, 0], dtype=np.float32), inv=0):
    """utils/image.py:27-60 with cv2.getAffineTransform replaced by its definition (cv2 is not a dependency here)."""
    if not isinstance(scale, np.ndarray) and not isinstance(scale, list):
        scale = np.array([scale, scale], dtype=np.float32)
    src_w, dst_w, dst_h = scale[0], output_size[0], output_size[1]
    rot_rad = np.pi * rot / 180
    src_dir = _get_dir([0, src_w * -0.5], rot_rad)
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center + scale * shift
    src[1, :] = center + src_dir + scale * shift
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5], np.float32) + dst_dir
    src[2:, :] = _third(src[0, :], src[1, :])
    dst[2:, :] = _third(dst[0, :], dst[1, :])
    return _affine_from_points(dst, src) if inv else _affine_from_points(src, dst)


def invert_affine(m):
    """The inversion cv2.warpAffine applies to a forward map (imgwarp.cpp), same operation order, double precision."""
    m = np.asarray(m, np.float64).copy()
    d = m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[1, 1] * d, m[0, 0] * d
    m[0, 0], m[0, 1], m[1, 0], m[1, 1] = a11, m[0, 1] * -d, m[1, 0] * -d, a22
    b1 = -m[0, 0] * m[0, 2] - m[0, 1] * m[1, 2]
    b2 = -m[1, 0] * m[0, 2] - m[1, 1] * m[1, 2]
    m[0, 2], m[1, 2] = b1, b2
    return m


def warp_normalize(image, image_right, trans_input, out_hw, mean, std, device="cuda"):
    """uint8 H x W x 3 images (numpy or CUDA tensors) -> two float32 1 x 3 x h x w tensors: warpAffine(INTER_LINEAR) +
    (x / 255 - mean) / std + transpose, one launch."""
    def up(img):
        if img is None:
            return None
        if isinstance(img, np.ndarray):
            img = torch.from_numpy(np.ascontiguousarray(img)).to(device, non_blocking=True)
        return _chk(img, "image", torch.uint8)
    L, R = up(image), up(image_right)
    sh, sw = int(L.shape[0]), int(L.shape[1])
    if L.dim() != 3 or L.shape[2] != 3 or (R is not None and R.shape != L.shape):
        raise RuntimeError("side_b200.preprocess: images must be H x W x 3 uint8 of one size")
    h, w = int(out_hw[0]), int(out_hw[1])
    outL = torch.empty((1, 3, h, w), device=L.device, dtype=torch.float32)
    outR = torch.empty((1, 3, h, w), device=L.device, dtype=torch.float32) if R is not None else None
    inv = invert_affine(trans_input).reshape(-1)
    m6 = (ctypes.c_double * 6)(*[float(v) for v in inv])
    mean3 = (_lib._f * 3)(*[float(np.float32(v)) for v in np.asarray(mean).reshape(-1)[:3]])
    std3 = (_lib._f * 3)(*[float(np.float32(v)) for v in np.asarray(std).reshape(-1)[:3]])
    _lib.check(_lib.load().side_preprocess_u8(L.data_ptr(), None if R is None else R.data_ptr(), sh, sw, m6, mean3, std3,
                                              outL.data_ptr(), None if outR is None else outR.data_ptr(), h, w, _stream()),
               "side_preprocess_u8")
    return outL, outR


def pre_process(opt, image, image_right, calib, device="cuda"):
    """``StereoDetector.pre_process(image, image_right, calib)`` (stereoDetector.py:45-82) with ``opt`` explicit:
    returns ``(inp, inp_right, meta)`` -- the tensors already on the device."""
    height, width = image.shape[0:2]
    inp_height, inp_width = opt.input_h, opt.input_w
    c = np.array([width / 2, height / 2], dtype=np.float32)
    s = np.array([inp_width, inp_height], dtype=np.int32) if opt.keep_res else np.array([width, height], dtype=np.int32)
    trans_input = get_affine_transform(c, s, 0, [inp_width, inp_height])
    inp, inp_right = warp_normalize(image, image_right, trans_input, (inp_height, inp_width), opt.mean, opt.std, device)
    trans = get_affine_transform(c, s, 0, [opt.output_w, opt.output_h])
    trans_inv = get_affine_transform(c, s, 0, [opt.output_w, opt.output_h], inv=1)
    meta = {'c': c, 's': s, 'out_height': inp_height // opt.down_ratio, 'out_width': inp_width // opt.down_ratio,
            'calib': calib, 'trans': trans, 'trans_inv': trans_inv}
    return inp, inp_right, meta
