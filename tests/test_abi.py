"""CPU: the C-ABI library loads, exports exactly what include/side_b200.h declares, and refuses to compute
without device memory (no CPU fallback anywhere in the product)."""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "side_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(side_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from side_b200 import _lib
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.SO_PATH]).decode()
    exported = sorted(set(re.findall(r" T (side_[a-z0-9_]+)", out)))
    assert exported == _header_symbols()
    assert sorted(_lib.SIGNATURES) == _header_symbols()


def test_library_is_sm100a_only(lib):
    from side_b200 import _lib
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", _lib.SO_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_compute_without_device_memory(lib):
    from side_b200 import _lib
    assert lib.side_abi_version() == 1
    a = np.zeros(16, np.float32)
    p = a.ctypes.data
    rc = lib.side_softargmin_fwd(p, p, p, None, 1, 1, 1, None)
    assert rc == -2 and b"device pointer" in lib.side_last_error()       # SIDE_ERR_NOT_DEVICE
    assert lib.side_inst_costvol_fwd(p, p, p, p, p, None, p, p, None, 3, 1, 4, 8, 8, 1, 16, 7.0, 0, None, 0, None) == -1  # D < 2
    assert lib.side_dcn_fwd(p, p, p, p, None, None, None, p, 1, 4, 4, 4, 4, 3, 3, 1, 1, 1, 1, 1, 1, 3, 0, 0, 0, None, 0, None) == -1
    assert lib.side_decode_ws_bytes(2, 3, 100) == 256 + 8 * 600
    assert lib.side_dcn_fwd_ws_bytes(1, 64, 96, 320, 64, 3, 3, 0) == 4 * 64 * 64 * 9


def test_ops_reject_cpu_tensors(lib):
    from side_b200 import ops
    from side_b200.decode import bbox_decode
    x = torch.zeros(1, 4, 5, 5)
    with pytest.raises(RuntimeError, match="CPU"):
        ops.softargmin(torch.zeros(2, 4, 4, 4), torch.zeros(2, 4))
    with pytest.raises(RuntimeError, match="CPU"):
        bbox_decode(torch.zeros(1, 3, 8, 8), torch.zeros(1, 3, 8, 8), torch.zeros(1, 3, 8, 8), K=4)
    with pytest.raises(RuntimeError, match="CPU"):
        ops.concat_volume(x, x, 2)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from side_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "SO_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU"):
        _lib.load()


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "side_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "torch_port" not in src and "c_oracle" not in src, f


def test_state_dict_contract():
    """24 995 606 parameters, key names / shapes of the canonical model."""
    from side_b200.networks import get_pose_net
    from side_b200.utils.synthetic import HEADS
    m = get_pose_net(34, HEADS, 256)
    assert sum(p.numel() for p in m.parameters()) == 24995606
    sd = m.state_dict()
    assert tuple(sd["feature_extraction.dla_up.ida_0.proj_1.conv.weight"].shape) == (256, 512, 3, 3)
    assert tuple(sd["feature_extraction.dla_up.ida_0.proj_1.conv.conv_offset_mask.weight"].shape) == (27, 512, 3, 3)
    assert tuple(sd["feaRuduce.0.weight"].shape) == (32, 64, 1, 1)
    assert tuple(sd["depth_estimator.dres0.0.weight"].shape) == (64, 96, 3, 3, 3)
    assert tuple(sd["depth_estimator.classify.3.weight"].shape) == (1, 64, 3, 3, 3)
    assert float(sd["hm.2.bias"][0]) == pytest.approx(-2.19)
    assert float(sd["feature_extraction.ida_up.node_1.conv.conv_offset_mask.weight"].abs().max()) == 0.0   # init_offset
