"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference Python.

Imports the reference's own ``stereo_network_old``, ``decode`` and ``DCNv2/dcn_v2.py``
from ``/root/reference/src/lib`` so that golden vectors can be generated from the
reference itself (``oracle/gen_golden.py``) and the C / torch restatements in this
directory can be pinned against it.  ``/root/reference`` exists only in the build
container, never on the GPU box: nothing under ``tests/ -m gpu``, ``bench.py`` or
``__graft_entry__.smoke()`` may import this module.

Shims -- none of them changes the reference's arithmetic:
  (a) ``_ext``: the legacy DCNv2 extension cannot be built (THC removed from
      torch >= 1.11, ``DCNv2/src/cuda/dcn_v2_cuda.cu:7-12``) and its CPU path only
      raises (``DCNv2/src/cpu/dcn_v2_cpu.cpp:8-23``); ``dcn_v2_forward`` is mapped
      to ``torchvision.ops.deform_conv2d`` as BASELINE.json allows, and
      ``dcn_v2_backward`` to its autograd.
  (b) ``matplotlib.pyplot`` stub (``stereo_network_old.py:13`` imports it, never uses it).
  (c) ``DLA.load_pretrained_model`` neutralised (hard-coded ``pretrained=True`` would
      download, ``stereo_network_old.py:267``).
  (d) ``torch.Tensor.cuda`` -> identity and ``torch.cuda.FloatTensor`` -> ``torch.FloatTensor``
      (hard-coded ``.cuda()`` at ``stereo_network_old.py:39,133,233,370,383``, ``decode.py:64-75,126``).
"""
import os
import sys
import types
import importlib

import torch

REF_ROOT = os.environ.get("SIDE_REFERENCE_ROOT", "/root/reference")
REF_LIB = os.path.join(REF_ROOT, "src", "lib")


def available():
    return os.path.isdir(REF_LIB)


def _make_ext():
    import torchvision.ops as tvo

    ext = types.ModuleType("_ext")

    def dcn_v2_forward(input, weight, bias, offset, mask, kh, kw, sh, sw, ph, pw, dh, dw, dg):
        return tvo.deform_conv2d(input, offset, weight, bias, stride=(sh, sw), padding=(ph, pw),
                                 dilation=(dh, dw), mask=mask)

    def dcn_v2_backward(input, weight, bias, offset, mask, grad_output, kh, kw, sh, sw, ph, pw, dh, dw, dg):
        with torch.enable_grad():
            leaves = [t.detach().clone().requires_grad_(True) for t in (input, offset, mask, weight, bias)]
            out = tvo.deform_conv2d(leaves[0], leaves[1], leaves[3], leaves[4], stride=(sh, sw),
                                    padding=(ph, pw), dilation=(dh, dw), mask=leaves[2])
            gi, go, gm, gw, gb = torch.autograd.grad(out, leaves, grad_output)
        return [gi, go, gm, gw, gb]

    ext.dcn_v2_forward = dcn_v2_forward
    ext.dcn_v2_backward = dcn_v2_backward
    return ext


_loaded = {}


def load():
    """Returns a namespace with the reference modules (imported once)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_LIB)

    sys.modules.setdefault("_ext", _make_ext())
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    if "cv2" not in sys.modules:
        try:
            import cv2  # noqa: F401
        except Exception:
            sys.modules["cv2"] = types.ModuleType("cv2")

    # (d) device shims -- CPU-only container
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        torch.cuda.FloatTensor = torch.FloatTensor

    if REF_LIB not in sys.path:
        sys.path.insert(0, REF_LIB)
    dcn_dir = os.path.join(REF_LIB, "models", "networks", "DCNv2")

    fe = importlib.import_module("models.networks.feature_extraction_dla34")
    fe.DLA.load_pretrained_model = lambda self, *a, **k: None  # (c)
    net = importlib.import_module("models.networks.stereo_network_old")
    dec = importlib.import_module("models.decode")
    dcn = importlib.import_module("models.networks.DCNv2.dcn_v2")
    utils = importlib.import_module("models.utils")
    _loaded.update(dict(net=net, decode=dec, dcn=dcn, fe=fe, utils=utils, dcn_dir=dcn_dir))
    return types.SimpleNamespace(**_loaded)
