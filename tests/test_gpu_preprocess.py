"""GPU: detector input preparation (csrc/preprocess.cu) against the numpy restatement of cv2.warpAffine + normalisation
(oracle/torch_port.py).  cv2 is not in the image: parity with the real library is unpinned (stated in the header and in
DESIGN.md); what is checked is the published algorithm, bit for bit, and exact pass-through properties."""
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import torch_port as tp  # noqa: E402

MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def _opt(keep_res=False):
    return types.SimpleNamespace(input_h=384, input_w=1280, output_h=96, output_w=320, down_ratio=4, keep_res=keep_res,
                                 mean=np.array(MEAN, np.float32).reshape(1, 1, 3), std=np.array(STD, np.float32).reshape(1, 1, 3))


@pytest.mark.parametrize("hw", [(375, 1242), (370, 1224), (384, 1280)])
def test_pre_process_kitti_frame(lib, hw):
    from side_b200 import preprocess as pp
    rng = np.random.RandomState(hw[0])
    img_l = rng.randint(0, 256, hw + (3,), dtype=np.uint8)
    img_r = rng.randint(0, 256, hw + (3,), dtype=np.uint8)
    opt = _opt()
    inp, inp_r, meta = pp.pre_process(opt, img_l, img_r, calib=None)
    assert inp.shape == (1, 3, 384, 1280) and inp_r.shape == inp.shape and inp.is_cuda
    c = np.array([hw[1] / 2, hw[0] / 2], dtype=np.float32)
    s = np.array([hw[1], hw[0]], dtype=np.int32)
    t_in = pp.get_affine_transform(c, s, 0, [1280, 384])
    for got, img in ((inp, img_l), (inp_r, img_r)):
        ref = tp.pre_process_ref(img, t_in, (384, 1280), MEAN, STD)
        assert np.array_equal(got.cpu().numpy(), ref)                     # integer warp + the same three float ops: bit-exact
    # the maps to and from the 1/4-scale output are inverse to each other
    full = np.vstack([meta['trans'], [0, 0, 1]]) @ np.vstack([meta['trans_inv'], [0, 0, 1]])
    assert np.allclose(full, np.eye(3), atol=1e-9)
    assert meta['out_height'] == 96 and meta['out_width'] == 320


def test_identity_warp_passes_pixels_through(lib):
    """Source size == input size: the warp is the identity, every output value is ((p / 255) - mean) / std of its own pixel."""
    from side_b200 import preprocess as pp
    rng = np.random.RandomState(1)
    img = rng.randint(0, 256, (384, 1280, 3), dtype=np.uint8)
    inp, none, _ = pp.pre_process(_opt(), img, None, calib=None)
    assert none is None
    ref = ((img.astype(np.float32) / 255.) - np.array(MEAN, np.float32)) / np.array(STD, np.float32)
    assert np.array_equal(inp[0].cpu().numpy(), ref.transpose(2, 0, 1))


def test_keep_res_crops_and_pads_with_border(lib):
    """keep_res: scale 1 about the image centre; pixels outside the source are the constant border 0 before normalisation."""
    from side_b200 import preprocess as pp
    img = np.full((300, 1000, 3), 200, np.uint8)
    inp, _, _ = pp.pre_process(_opt(keep_res=True), img, None, calib=None)
    out = inp[0].cpu().numpy()
    border = ((0.0 - np.array(MEAN, np.float32)) / np.array(STD, np.float32)).astype(np.float32)
    inside = ((np.float32(200) / np.float32(255.) - np.array(MEAN, np.float32)) / np.array(STD, np.float32)).astype(np.float32)
    assert np.array_equal(out[:, 0, 0], border) and np.array_equal(out[:, 192, 640], inside)
    t_in = pp.get_affine_transform(np.array([500., 150.], np.float32), np.array([1280, 384], np.int32), 0, [1280, 384])
    assert np.array_equal(out, tp.pre_process_ref(img, t_in, (384, 1280), MEAN, STD)[0])


def test_rejects_host_pointer(lib):
    from side_b200 import preprocess as pp
    with pytest.raises(RuntimeError):
        pp.warp_normalize(torch.zeros(8, 8, 3, dtype=torch.uint8), None, np.eye(2, 3), (8, 8), MEAN, STD, device="cpu")
