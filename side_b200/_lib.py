"""ctypes binding of ``libside_b200.so`` (the C ABI declared in ``include/side_b200.h``).

There is deliberately no fallback: if the shared object is missing or an entry point
returns an error, a ``RuntimeError`` is raised -- the same way the reference's ``_ext``
raises through ``AT_ASSERTM`` / ``AT_ERROR`` (DCNv2/src/cuda/dcn_v2_cuda.cu:61-85,
DCNv2/src/dcn_v2.h:35-39 "Not implemented on the CPU").
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libside_b200.so")

_vp, _i, _f, _ll, _sz = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_size_t

# name -> (restype, argtypes); mirrors include/side_b200.h one to one
SIGNATURES = {
    "side_abi_version": (_i, []),
    "side_last_error": (C.c_char_p, []),
    "side_device_ok": (_i, []),
    "side_launch_count": (_ll, [_i]),
    "side_dcn_fwd_ws_bytes": (_sz, [_i] * 8),
    "side_dcn_fwd": (_i, [_vp] * 8 + [_i] * 14 + [_ll, _ll, _i, _vp, _sz, _vp]),
    "side_dcn_fwd_cl": (_i, [_vp, _vp, _i] + [_vp] * 5 + [_i] * 14 + [_vp, _sz, _vp]),
    "side_dcn_bwd_ws_bytes": (_sz, [_i] * 8),
    "side_dcn_bwd": (_i, [_vp] * 10 + [_i] * 14 + [_ll, _ll, _i, _vp, _sz, _vp]),
    "side_proposal_shift": (_i, [_vp] * 3 + [_i] * 3 + [_f] + [_vp] * 4),
    "side_inst_costvol_ws_bytes": (_sz, [_i] * 4),
    "side_inst_costvol_fast_ws_bytes": (_sz, [_i] * 6),
    "side_inst_costvol_fwd": (_i, [_vp] * 9 + [_i] * 7 + [_f, _i, _vp, _sz, _vp]),
    "side_inst_costvol_cl_ws_bytes": (_sz, [_i] * 4),
    "side_inst_costvol_fwd_cl": (_i, [_vp] * 10 + [_i] * 7 + [_f, _i, _vp, _sz, _vp]),
    "side_inst_costvol_bwd": (_i, [_vp] * 9 + [_i] * 7 + [_f, _i, _vp]),
    "side_inst_costvol_bwd_fast_ws_bytes": (_sz, [_i] * 7),
    "side_inst_costvol_bwd_fast": (_i, [_vp] * 9 + [_i] * 7 + [_f, _i, _vp, _sz, _vp]),
    "side_xcross_gate_fwd": (_i, [_vp] * 3 + [_i] * 4 + [_vp]),
    "side_xcross_gate_bwd": (_i, [_vp] * 3 + [_i] * 4 + [_vp]),
    "side_softargmin_fwd": (_i, [_vp] * 4 + [_i] * 3 + [_vp]),
    "side_softargmin_bwd": (_i, [_vp] * 6 + [_i] * 3 + [_vp]),
    "side_decode_ws_bytes": (_sz, [_i] * 3),
    "side_bbox_decode": (_i, [_vp] * 11 + [_i] * 5 + [_f, _i, _vp, _sz, _vp]),
    "side_ddd_decode": (_i, [_vp] * 12 + [_i] * 7 + [_vp, _sz, _vp]),
    "side_concat_volume_fwd": (_i, [_vp] * 3 + [_i] * 5 + [_vp]),
    "side_concat_volume_bwd": (_i, [_vp] * 3 + [_i] * 5 + [_vp]),
    "side_gwc_volume_fwd": (_i, [_vp] * 3 + [_i] * 6 + [_vp]),
    "side_gwc_volume_bwd": (_i, [_vp] * 5 + [_i] * 6 + [_vp]),
    "side_dw_deconv_fwd": (_i, [_vp] * 3 + [_i] * 7 + [_vp]),
    "side_dw_deconv_bwd": (_i, [_vp] * 5 + [_i] * 7 + [_vp]),
    "side_conv_tc_weight_bytes": (_sz, [_i] * 3),
    "side_conv_tc_prep_weights": (_i, [_vp] * 2 + [_i] * 3 + [_vp]),
    "side_conv3d_tc_fwd": (_i, [_vp] * 9 + [_i] * 11 + [_vp]),
    "side_conv_tc_prep_weights_f16": (_i, [_vp] * 2 + [_i] * 3 + [_vp]),
    "side_conv3d_tc_fwd_f16": (_i, [_vp] * 9 + [_i] * 11 + [_vp]),
    "side_ncdhw_to_cl_split_f16": (_i, [_vp] * 5 + [_i] * 2 + [_ll, _i, _i, _vp]),
    "side_idaup_fuse_cl_f16": (_i, [_vp] * 6 + [_i] * 6 + [_vp]),
    "side_gate_mul_split_f16": (_i, [_vp] * 4 + [_i] * 5 + [_vp]),
    "side_maxpool_hw2_cl_f16": (_i, [_vp] * 4 + [_i] * 5 + [_vp]),
    "side_ncdhw_to_cl_split": (_i, [_vp] * 5 + [_i] * 2 + [_ll, _i, _vp]),
    "side_tf32_split": (_i, [_vp] * 3 + [_ll, _vp]),
    "side_f16_split": (_i, [_vp] * 3 + [_ll, _vp]),
    "side_gate_mul_split": (_i, [_vp] * 4 + [_i] * 5 + [_vp]),
    "side_maxpool_hw2_cl": (_i, [_vp] * 4 + [_i] * 5 + [_vp]),
    "side_conv3d_c1_cl": (_i, [_vp] * 3 + [_i] * 5 + [_vp]),
    "side_cl_to_nchw": (_i, [_vp] * 2 + [_i] * 2 + [_ll, _vp]),
    "side_cl_concat": (_i, [C.POINTER(_vp), C.POINTER(_i), _i, _vp, _ll, _vp]),
    "side_cl_to_nchw_ld": (_i, [_vp, _i, _vp] + [_i] * 2 + [_ll, _vp]),
    "side_conv_tc_set_mode": (_i, [_i]),
    "side_tc_range_guard": (_i, [_vp, _i]),
    "side_stem_conv_fwd": (_i, [_vp] * 5 + [_i] * 8 + [_vp]),
    "side_stem_conv_fwd_s2d": (_i, [_vp] * 6 + [_i] * 4 + [_vp]),
    "side_voxel_coords": (_i, [_vp] * 8 + [_i] * 7 + [_vp] * 8),
    "side_voxel_volume_ws_bytes": (_sz, [_i] * 4),
    "side_voxel_volume_fwd": (_i, [_vp] * 11 + [_i] * 8 + [_vp, _sz, _vp]),
    "side_voxel_volume_bwd": (_i, [_vp] * 10 + [_i] * 8 + [_vp, _sz, _vp]),
    "side_bn_train_ws_bytes": (_sz, [_i, _i, _ll]),
    "side_bn_train_fwd": (_i, [_vp] * 8 + [_i, _i, _ll, _f, _f, _vp, _sz, _vp]),
    "side_bn_train_bwd": (_i, [_vp] * 8 + [_i, _i, _ll, _vp, _sz, _vp]),
    "side_pow2_range_scale": (_i, [_vp, _ll, _vp, _i, _vp, _i, _vp, _vp]),
    "side_conv_wgrad_tc_ws_bytes": (_sz, [_i] * 10),
    "side_conv_wgrad_tc": (_i, [_vp] * 4 + [_i] * 10 + [_vp, _sz, _vp]),
    "side_preprocess_u8": (_i, [_vp] * 2 + [_i] * 2 + [C.POINTER(C.c_double)] + [C.POINTER(_f)] * 2 + [_vp] * 2 + [_i] * 2 + [_vp]),
    "side_dense_align_prep_u8": (_i, [_vp] * 2 + [_i] * 2 + [C.POINTER(_f)] * 2 + [_vp]),
    "side_dense_align_up2_pack": (_i, [_vp] * 2 + [_i] * 2 + [_vp]),
    "side_dense_align_pack": (_i, [_vp] * 2 + [_i] * 2 + [_vp]),
    "side_dense_align_unpack": (_i, [_vp] * 2 + [_i] * 2 + [_vp]),
    "side_dense_align_sample": (_i, [_vp] * 3 + [_i] + [_f] * 3 + [_i] * 3 + [_vp] * 4),
    "side_dense_align_enum": (_i, [_vp] * 5 + [_f] + [_i] * 6 + [_vp] * 4),
}

# flag values (include/side_b200.h)
DCN_MASK_IS_LOGIT = 1 << 0
DCN_FUSE_AFFINE = 1 << 1
DCN_FUSE_RELU = 1 << 2
DCN_PREC_FP32 = 0 << 4
DCN_PREC_3XTF32 = 1 << 4
DCN_PREC_TF32 = 2 << 4
DCN_PREC_3XFP16 = 3 << 4
DCN_BWD_SCALAR = 1 << 8
DCN_BWD_SIMT_GEMM = 1 << 9
VOL_GATE = 1 << 0
VOL_FMA = 1 << 1
VOL_SEPARABLE = 1 << 2
VOL_XCROSS = 1 << 3
VOL_BWD_SCALAR = 1 << 4
VOL_FEAT_NHWC = 1 << 5
VOL_NO_DIFF = 1 << 6
DECODE_HEAT_IS_LOGIT = 1 << 0
DA_ALIGN_CORNERS = 1 << 0
VOXEL_ALIGN_CORNERS = 1 << 0

_lib = None


def load():
    """Loads the shared object (once).  Raises RuntimeError when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            "side_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C side_b200/csrc`).  There is no CPU / PyTorch fallback." % SO_PATH)
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.side_abi_version() != 1:
        raise RuntimeError("side_b200: ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().side_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg))


def launch_count(reset=False):
    return int(load().side_launch_count(1 if reset else 0))
