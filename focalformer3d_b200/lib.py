"""ctypes binding of libff3d.so (the C ABI declared in include/ff3d.h).

The product path has no CPU or PyTorch fallback: if the shared library is missing this module raises at import,
and every op raises ``Ff3dError`` on a non-zero return code.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libff3d.so")

ACT_NONE, ACT_RELU, ACT_RELU6 = 0, 1, 2
GEMM_ROWS, GEMM_CONV2D, GEMM_SPARSE = 0, 1, 2


class Ff3dError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("mode", C.c_int), ("M", C.c_int), ("m_dev", C.c_void_p),
        ("cin", C.c_int), ("cout", C.c_int), ("taps", C.c_int),
        ("x", C.c_void_p), ("ldx", C.c_int), ("x2", C.c_void_p),
        ("w", C.c_void_p), ("ldw", C.c_int), ("bias", C.c_void_p),
        ("res", C.c_void_p), ("ldres", C.c_int),
        ("y", C.c_void_p), ("ldy", C.c_int), ("act", C.c_int),
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Ho", C.c_int), ("Wo", C.c_int),
        ("kh", C.c_int), ("kw", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
        ("x_bstride", C.c_longlong), ("y_bstride", C.c_longlong), ("y_row0", C.c_longlong),
        ("ux", C.c_int), ("uy", C.c_int), ("dx", C.c_int), ("dy", C.c_int),
        ("nbr", C.c_void_p), ("nbr_stride", C.c_int), ("y_off", C.c_void_p), ("res_after_act", C.c_int),
        ("tile_mask", C.c_void_p),
        ("xs", C.c_void_p), ("ldxs", C.c_int), ("xs_lo", C.c_int), ("xs_rows", C.c_int),
        ("res_s", C.c_void_p), ("ldres_s", C.c_int), ("res_s_lo", C.c_int),
        ("ys", C.c_void_p), ("ldys", C.c_int), ("ys_lo", C.c_int),
        ("y_row", C.c_void_p), ("zero_row", C.c_int),
    ]


class DecoderLayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w_qkv", "b_qkv", "w_o", "b_o", "w_oa", "b_oa", "w_op", "b_op", "w_f1", "b_f1",
                                          "w_f2", "b_f2")] + [("ln_gamma", C.c_void_p * 3), ("ln_beta", C.c_void_p * 3)]


class DecoderStageDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("nq", C.c_int), ("n_layers", C.c_int), ("hidden", C.c_int), ("heads", C.c_int),
        ("n_levels", C.c_int), ("n_points", C.c_int), ("ffn", C.c_int),
        ("lvl_h", C.c_int * 4), ("lvl_w", C.c_int * 4), ("lvl_start", C.c_int * 4),
        ("x_in", C.c_void_p), ("qpe", C.c_void_p), ("q_pos", C.c_void_p),
        ("ref_w", C.c_float), ("ref_h", C.c_float),
        ("value", C.c_void_p), ("ldv", C.c_int), ("v_bstride", C.c_longlong),
        ("layers", DecoderLayerWeights * 4),
        ("w_h1", C.c_void_p), ("b_h1", C.c_void_p), ("n_h1", C.c_int),
        ("w_h2", C.c_void_p), ("b_h2", C.c_void_p), ("n_pred", C.c_int),
        ("x_out", C.c_void_p), ("pred", C.c_void_p), ("ld_pred", C.c_int), ("pred_cols", C.c_int),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("overflow_dev", C.c_void_p),
    ]


# name -> (restype, argtypes); mirrors include/ff3d.h one to one
_P, _I, _F, _LL, _SZ = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_size_t
_IP = C.POINTER(C.c_int)
_FP = C.POINTER(C.c_float)
SIGNATURES = {
    "ff3d_last_error": (C.c_char_p, []),
    "ff3d_version": (_I, []),
    "ff3d_voxelize_workspace_bytes": (_SZ, [_I, _I, _I, _I]),
    "ff3d_voxelize_hard": (_I, [_P, _I, _I, _IP, _I, _FP, _FP, _I, _I, _P, _P, _P, _P, _I, _P, _P, _SZ, _P]),
    "ff3d_vfe_hard": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _I, _I, _P]),
    "ff3d_sp_hash_build": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P]),
    "ff3d_sp_subm_map": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P]),
    "ff3d_sp_down_build": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _IP, _IP, _IP, _P, _P, _I, _I, _I, _I, _P, _P,
                                _I, _P, _P, _P]),
    "ff3d_sp_down_sites": (_I, [_P, _P, _I, _I, _I, _I, _I, _IP, _IP, _IP, _P, _P, _I, _I, _I, _I, _P, _P, _I, _P, _P, _P]),
    "ff3d_sp_down_sites_scratch_ints": (_I, [_I]),
    "ff3d_sp_tap_keys": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _I, _IP, _IP, _IP, _P, _P, _P]),
    "ff3d_sp_nbr_permute": (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "ff3d_sort_workspace_bytes": (_SZ, [_I]),
    "ff3d_sort_pairs": (_I, [_P, _P, _P, _I, _I, _P, _P, _P, _SZ, _P]),
    "ff3d_sp_level_permute": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _I, _P, _P]),
    "ff3d_sp_nbr_build": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _I, _IP, _IP, _IP, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "ff3d_sp_gather_rows": (_I, [_P, _I, _P, _P, _I, _P, _I, _I, _P]),
    "ff3d_sp_bev_offsets": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "ff3d_igemm": (_I, [C.POINTER(GemmDesc), _P]),
    "ff3d_tcgemm": (_I, [C.POINTER(GemmDesc), _P, _P]),
    "ff3d_tcgemm_bn": (_I, [C.POINTER(GemmDesc), _P, _I, _P]),
    "ff3d_tcgemm_f16": (_I, [C.POINTER(GemmDesc), _P, _I, _P, _P]),
    "ff3d_tcgemm_f16_ntile": (_I, [_I, _I]),
    "ff3d_tcgemm_f16_stages": (_I, [_I, _I]),
    "ff3d_tmagemm": (_I, [C.POINTER(GemmDesc), _P, _I, _P, _P]),
    "ff3d_tmagemm_supported": (_I, [C.POINTER(GemmDesc)]),
    "ff3d_tmagemm_conv_patch": (None, [_I, _I, _IP, _IP]),
    "ff3d_split_rows": (_I, [_P, _I, _P, _LL, _I, _P, _I, _I, _P, _P]),
    "ff3d_unsplit_rows": (_I, [_P, _I, _I, _P, _LL, _I, _P, _I, _P]),
    "ff3d_tcgemm_ntile": (_I, [_I, _I]),
    "ff3d_tcgemm_stages": (_I, [_I, _I]),
    "ff3d_dwconv3x3": (_I, [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "ff3d_dwconv3x3_split": (_I, [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "ff3d_layernorm": (_I, [_P, _P, _P, _P, _I, _I, _F, _P]),
    "ff3d_hip_workspace_bytes": (_SZ, [_I, _I, _I, _I]),
    "ff3d_hip_stage": (_I, [_P, _I, _P, _I, _P, _P, _P, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P,
                            _P, _P, _SZ, _P]),
    "ff3d_sine_embed": (_I, [_P, _F, _F, _P, _P, _I, _P]),
    "ff3d_mha_core": (_I, [_P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _P]),
    "ff3d_msda": (_I, [_P, _I, _I, _LL, _IP, _IP, _IP, _I, _I, _P, _F, _F, _P, _I, _P, _I, _P, _I, _I, _I, _I, _P]),
    "ff3d_roi_sample": (_I, [_P, _I, _P, _I, _LL, _IP, _IP, _IP, _I, _I, _I, _F, _F, _F, _F, _F, _FP, _P, _I, _I, _P]),
    "ff3d_roi_sample_split": (_I, [_P, _I, _P, _I, _LL, _IP, _IP, _IP, _I, _I, _I, _F, _F, _F, _F, _F, _FP, _P, _I, _I, _P, _P]),
    "ff3d_decoder_stage_workspace_bytes": (_SZ, [_I, _I, _I]),
    "ff3d_decoder_stage": (_I, [C.POINTER(DecoderStageDesc), _P]),
    "ff3d_head_update": (_I, [_P, _I, _P, _P, _I, _I, _P]),
    "ff3d_box_decode": (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _F, _F, _F, _F, _FP, _P, _P, _P, _P, _P]),
    "ff3d_assemble_sweeps": (_I, [_P, _I, _IP, _I, C.POINTER(C.c_double), C.POINTER(C.c_double), _FP, C.POINTER(C.c_ubyte),
                                  C.POINTER(C.c_ubyte), _F, _FP, _F, _P, _P]),
    "ff3d_image_preprocess": (_I, [_P, _I, _I, _I, _I, _I, _FP, _FP, _I, _I, _P, _I, _I, _P]),
    "ff3d_nms_tasks": (_I, [_P, _I, _P, _P, _P, _I, _I, _I, C.POINTER(C.c_uint), _FP, _I, _I, _I, _P, _P]),
    "ff3d_boxes_iou_bev": (_I, [_P, _I, _I, _P, _I, _I, _P, _P]),
    "ff3d_box_voting": (_I, [_P, _I, _I, _P, _I, _I, _F, _P, _P]),
    "ff3d_boxes_map_back": (_I, [_P, _I, _I, _I, _F, _I, _I, _P]),
    "ff3d_add_rows": (_I, [_P, _P, _P, _LL, _P]),
    "ff3d_add_bcast_rows": (_I, [_P, _P, _P, _I, _LL, _I, _P]),
    "ff3d_add_bcast_rows_split": (_I, [_P, _P, _P, _I, _LL, _I, _P, _P]),
    "ff3d_class_select": (_I, [_P, _I, _P, _P, _I, _I, _I, _P, _I, _I, _P]),
    "ff3d_local_attention": (_I, [_P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _P]),
    "ff3d_nchw_to_nhwc": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "ff3d_maxpool3x3s2": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "ff3d_upsample_add": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "ff3d_lss_splat": (_I, [_P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _I, _I, _P]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise Ff3dError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C focalformer3d_b200/csrc`). There is no CPU / PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc, what=""):
    if rc != 0:
        msg = lib.ff3d_last_error()
        raise Ff3dError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def int_array(vals):
    return (C.c_int * len(vals))(*[int(v) for v in vals])


def float_array(vals):
    return (C.c_float * len(vals))(*[float(v) for v in vals])
