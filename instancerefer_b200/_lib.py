"""ctypes binding of ``libinstancerefer_b200.so`` (the C ABI declared in
``include/instancerefer_b200.h``).  There is no fallback of any kind: if the shared library is
missing or a call fails, an exception is raised."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libinstancerefer_b200.so")

ENC_LAYERS = 13
ENC_LEVELS = 5

p = C.c_void_p
i32 = C.c_int32
i64 = C.c_int64
f64 = C.c_double
f32 = C.c_float


class EncoderParams(C.Structure):
    _fields_ = [("cin", i32), ("use_tc", i32),
                ("weight", p * ENC_LAYERS), ("wprep", p * ENC_LAYERS),
                ("bn_scale", p * ENC_LAYERS), ("bn_shift", p * ENC_LAYERS)]


class EncoderLayout(C.Structure):
    _fields_ = [("n_max", i64), ("cap", i64), ("total_bytes", i64),
                ("off_nlvl", i64), ("off_kcount", i64), ("off_scan", i64), ("scan_stride", i64), ("off_sync", i64),
                ("zero_bytes", i64), ("off_keys", i64), ("off_vals", i64),
                ("off_coords", i64 * ENC_LEVELS), ("off_pslot", i64),
                ("off_k3_in", i64 * ENC_LEVELS), ("off_k3_slot", i64 * ENC_LEVELS),
                ("off_k2_in", i64 * 4), ("off_k2_slot", i64 * 4),
                ("off_feat0", i64), ("off_feat", i64 * 3), ("off_T", i64)]


class EncoderTrainParams(C.Structure):
    _fields_ = [("cin", i32), ("use_tc", i32),
                ("weight", p * ENC_LAYERS), ("gamma", p * ENC_LAYERS), ("beta", p * ENC_LAYERS),
                ("running_mean", p * ENC_LAYERS), ("running_var", p * ENC_LAYERS),
                ("momentum", f32 * ENC_LAYERS), ("eps", f32)]


class EncoderTrainGrads(C.Structure):
    _fields_ = [("dweight", p * ENC_LAYERS), ("dgamma", p * ENC_LAYERS), ("dbeta", p * ENC_LAYERS)]


class EncoderTrainLayout(C.Structure):
    _fields_ = [("total_bytes", i64), ("off_y", i64 * ENC_LAYERS), ("off_out", i64 * ENC_LAYERS),
                ("off_mean", i64 * ENC_LAYERS), ("off_rstd", i64 * ENC_LAYERS), ("off_bn_scratch", i64),
                ("off_tr_out", i64 * 9), ("off_tr_slot", i64 * 9), ("off_grad", i64 * 6), ("off_wt", i64),
                ("off_absmax", i64)]


class MlpHead(C.Structure):
    _fields_ = [("M", i32), ("K", i32), ("N1", i32), ("N2", i32), ("norm", i32),
                ("eps", f32), ("momentum", f32), ("drop_p", f32), ("seed", C.c_uint64),
                ("w1", p), ("b1", p), ("gamma", p), ("beta", p), ("w2", p), ("b2", p),
                ("running_mean", p), ("running_var", p)]


class LangParams(C.Structure):
    _fields_ = [("B", i32), ("L", i32), ("E_in", i32), ("D", i32), ("H", i32), ("n_cls", i32),
                ("drop_p", f32), ("seed", C.c_uint64),
                ("w0", p), ("b0", p), ("w3", p), ("b3", p),
                ("wih", (p * 2) * 2), ("bih", (p * 2) * 2), ("whh", (p * 2) * 2), ("bhh", (p * 2) * 2),
                ("fcw", p * 4), ("fcb", p * 4), ("wc", p), ("bc", p)]


class LangGrads(C.Structure):
    _fields_ = [("dw0", p), ("db0", p), ("dw3", p), ("db3", p),
                ("dwih", (p * 2) * 2), ("dwhh", (p * 2) * 2), ("dbih", p * 2), ("dbhh", p * 2),
                ("dfcw", p), ("dfcb", p), ("dwc", p), ("dbc", p)]


class EdgeConvParams(C.Structure):
    _fields_ = [("nq", i32), ("k", i32), ("F", i32), ("ncls", i32), ("H1", i32), ("Fout", i32),
                ("ww1", p), ("bw1", p), ("ww2", p), ("bw2", p), ("wm1", p), ("bm1", p), ("wm2", p), ("bm2", p)]


class EdgeConvGrads(C.Structure):
    _fields_ = [(n, p) for n in ("dww1", "dbw1", "dww2", "dbw2", "dwm1", "dbm1", "dwm2", "dbm2")]


class SceneTail(C.Structure):
    _fields_ = [("B", i32), ("n_rows", i64), ("eps", f32), ("mom0", f32), ("mom1", f32), ("drop_p", f32), ("seed", C.c_uint64),
                ("kernel", p), ("g0", p), ("be0", p), ("w1", p), ("b1", p), ("g1", p), ("be1", p), ("w2", p), ("b2", p),
                ("rm0", p), ("rv0", p), ("rm1", p), ("rv1", p)]


class SceneTailGrads(C.Structure):
    _fields_ = [(n, p) for n in ("dkernel", "dg0", "dbe0", "dw1", "db1", "dg1", "dbe1", "dw2", "db2")]


# name -> (restype, argtypes); mirrors include/instancerefer_b200.h one to one
SIGNATURES = {
    "ir_version": (i32, []),
    "ir_last_error": (C.c_char_p, []),
    "ir_check_device": (i32, [i32]),
    "ir_launch_count": (i64, []),
    "ir_profile_enable": (i32, [i32]),
    "ir_profile_read": (i32, [p, p, p, i32, p]),
    "ir_debug_set": (i32, [i32]),
    "ir_debug_stamp": (i32, [p, i32, p]),
    "ir_conv_stamps_set": (i32, [p]),
    "ir_conv_stamps_meta": (i32, [p, i32, p]),
    "ir_encoder_persist_debug": (i32, [p]),
    "ir_encoder_persist_occupancy": (i32, []),
    "ir_encoder_mode_set": (i32, [i32]),
    "ir_gather_mode_set": (i32, [i32]),
    "ir_tune_set": (i32, [i32, i32]),
    "ir_encoder_layout": (i32, [i64, C.POINTER(EncoderLayout)]),
    "ir_encoder_workspace_bytes": (C.c_size_t, [i64]),
    "ir_encoder_reset": (i32, [p, i64, p]),
    "ir_voxelize": (i32, [p, p, i32, i32, i32, f64, p, i64, p]),
    "ir_voxelize_points_scratch_bytes": (C.c_size_t, [i64]),
    "ir_voxelize_points": (i32, [p, p, i32, i32, i32, f64, p, p, p, p, p]),
    "ir_encoder_build_maps": (i32, [p, i32, p, p, i64, p]),
    "ir_encoder_features": (i32, [C.POINTER(EncoderParams), p, p, i64, p, p]),
    "ir_encoder_features_pair": (i32, [C.POINTER(EncoderParams), p, p, i64, p, C.POINTER(EncoderParams), p, p, i64, p, p]),
    "ir_spconv_layer": (i32, [p, i32, i32, i32, p, i64, p, p, p, i64, p, p, i32, p, p, p, i32, p, p, p]),
    "ir_spconv_wprep_floats": (i64, [i32, i32, i32]),
    "ir_spconv_prepare_weights": (i32, [p, i32, i32, i32, p, p]),
    "ir_segmax": (i32, [p, p, p, i64, i32, i32, p, p, p]),
    "ir_bev": (i32, [p, p, p, i64, i32, p, p, p, i32, p, p, p, p, p]),
    "ir_conv2d_3x3_tc": (i32, [p, p, p, p, p, i64, p, p, p, i32, p, p, p, p, p]),
    "ir_conv2d_3x3": (i32, [p, i32, i32, i32, i32, p, p, p, p, i32, p, p]),
    "ir_scene_attention": (i32, [p, p, i32, i32, i32, p, p, p]),
    "ir_linear": (i32, [p, i32, i32, p, p, i32, i32, p, p]),
    "ir_gru_layer": (i32, [p, p, p, p, i32, i32, i32, p, p]),
    "ir_token_attention": (i32, [p, p, i64, p, p, p, i32, i32, i32, i32, p, p, p]),
    "ir_mlp_head": (i32, [p, i32, i32, p, p, i32, i32, p, p, p, p, i32, i32, p, p, p, p, p]),
    "ir_candidate_softmax": (i32, [p, p, p, p, i32, p, p, p]),
    "ir_instance_mean": (i32, [p, i32, i32, i32, p, p]),
    "ir_knn": (i32, [p, p, p, p, i32, i32, p, p]),
    "ir_edgeconv": (i32, [p, p, p, p, i32, i32, i32, i32, p, p, p, p, p, p, p, p, i32, p, p]),
    # training step
    "ir_rulebook_transpose": (i32, [p, p, i32, i64, p, i64, p, p, p]),
    "ir_spconv_wgrad": (i32, [p, i32, p, i32, i32, p, p, p, i64, p, p]),
    "ir_spconv_wgrad_scaled": (i32, [p, i32, p, p, i32, i32, p, p, p, i64, i32, p, p]),
    "ir_bn_scratch_floats": (i64, [i32]),
    "ir_bn_train_fwd": (i32, [p, p, i32, i32, p, p, p, i32, f32, f32, p, p, p, p, p, p, p]),
    "ir_bn_train_bwd": (i32, [p, p, p, p, i32, i32, p, p, p, i32, p, p, p, p, p, p, p]),
    "ir_spconv_layer_scaled": (i32, [p, p, i32, i32, i32, p, i64, p, p, p, i64, p, i32, p, p, p, p]),
    "ir_segmax_bwd": (i32, [p, p, p, i64, i32, i32, p, p, p, p, p]),
    "ir_cross_entropy": (i32, [p, p, i32, i32, p, p, p]),
    "ir_region_label": (i32, [p, p, p, i32, i32, p, p]),
    "ir_ref_loss": (i32, [p, p, p, p, i32, p, p, p, f32, f32, f64, p, p, p, p, p]),
    "ir_ref_eval": (i32, [p, p, p, p, i32, p, p, p, p, p, p, p, p, p, p]),
    "ir_adam_step": (i32, [p, p, p, p, i64, f32, f32, f32, f32, f32, i32, f32, p, p]),
    "ir_adam_hyper": (i32, [f32, f32, f32, i32, p]),
    "ir_adam_step_dev": (i32, [p, p, p, p, i64, p, f32, f32, f32, f32, f32, p, p]),
    "ir_encoder_train_layout": (i32, [i64, p, i32, C.POINTER(EncoderTrainLayout)]),
    "ir_encoder_train_forward": (i32, [C.POINTER(EncoderTrainParams), p, p, i64, p, p, p]),
    "ir_encoder_train_backward": (i32, [C.POINTER(EncoderTrainParams), p, p, i64, p, p, p, C.POINTER(EncoderTrainGrads), p]),
    "ir_encoder_train_backward_range": (i32, [C.POINTER(EncoderTrainParams), p, p, i64, p, p, p, C.POINTER(EncoderTrainGrads), i32, i32, p]),
    "ir_gemm": (i32, [i32, i32, i32, p, i32, i32, p, i32, i32, p, i32, p, i32, i32, p]),
    "ir_colsum": (i32, [p, i32, i32, p, p]),
    "ir_relu_bwd": (i32, [p, p, i64, p, p]),
    "ir_dropout_fwd": (i32, [p, i64, f32, C.c_uint64, p, p, p]),
    "ir_dropout_seed_step": (i32, [p]),
    "ir_dropout_bwd": (i32, [p, p, i64, f32, p, p]),
    "ir_layernorm_fwd": (i32, [p, i32, i32, p, p, f32, i32, p, p, p, p]),
    "ir_layernorm_bwd": (i32, [p, p, p, i32, i32, p, p, p, i32, p, p, p, p]),
    "ir_l2norm_fwd": (i32, [p, i32, i32, p, p]),
    "ir_l2norm_bwd": (i32, [p, p, i32, i32, p, p]),
    "ir_match_fwd": (i32, [p, p, p, i32, i32, i32, p, p]),
    "ir_match_bwd": (i32, [p, p, p, p, p, i32, i32, i32, i32, p, p, p]),
    "ir_im2col_3x3": (i32, [p, i32, i32, i32, i32, p, p]),
    "ir_col2im_3x3": (i32, [p, i32, i32, i32, i32, p, p]),
    "ir_bev_bwd": (i32, [p, p, p, p, p, i64, i32, p, i32, p, p, p]),
    "ir_scene_attention_bwd": (i32, [p, p, p, p, p, i32, i32, i32, p, p, p]),
    "ir_token_attention_bwd": (i32, [p, p, i64, p, p, p, p, p, i32, i32, i32, i32, p, p, p, p, p]),
    "ir_gru_layer_bwd": (i32, [p, p, p, p, p, p, i32, i32, i32, p, p, p, p]),
    "ir_edge_inputs": (i32, [p, p, p, p, i32, i32, i32, i32, p, p, p, p]),
    "ir_edge_max_fwd": (i32, [p, p, i32, i32, i32, p, p, p]),
    "ir_edge_max_bwd": (i32, [p, p, i32, i32, i32, p, p]),
    "ir_lang_train_arena_bytes": (i64, [C.POINTER(LangParams)]),
    "ir_lang_train_fwd": (i32, [C.POINTER(LangParams), p, p, p, p, p, p]),
    "ir_lang_train_bwd": (i32, [C.POINTER(LangParams), p, p, p, p, p, p, C.POINTER(LangGrads), p]),
    "ir_lang_train_view": (i32, [C.POINTER(LangParams), C.POINTER(i64), C.POINTER(i64)]),
    "ir_edgeconv_train_arena_bytes": (i64, [C.POINTER(EdgeConvParams)]),
    "ir_edgeconv_train_fwd": (i32, [C.POINTER(EdgeConvParams), p, p, p, p, p, p, p]),
    "ir_edgeconv_train_bwd": (i32, [C.POINTER(EdgeConvParams), p, p, C.POINTER(EdgeConvGrads), p]),
    "ir_prepare_scratch_bytes": (i64, [i64, i64, i32]),
    "ir_mesh_normals": (i32, [p, i64, p, i64, p, p]),
    "ir_align_vertices": (i32, [p, i64, p, p, p]),
    "ir_vertex_labels": (i32, [p, i64, p, p, i32, p, p, p]),
    "ir_instance_boxes": (i32, [p, p, i64, i32, p, p, p, p]),
    "ir_pointgroup_labels": (i32, [p, p, i32, i64, p, p, p]),
    "ir_keep_index": (i32, [p, i64, p, i32, p, p, p, p]),
    "ir_gather_rows": (i32, [p, i64, p, p, i64, p, p]),
    "ir_scene_tail_arena_bytes": (i64, [i64, i32]),
    "ir_scene_tail_train_fwd": (i32, [C.POINTER(SceneTail), p, p, p, p, p, p]),
    "ir_scene_tail_train_bwd": (i32, [C.POINTER(SceneTail), p, p, p, p, p, p, C.POINTER(SceneTailGrads), p]),
    "ir_mlp_head_arena_bytes": (i64, [i32, i32]),
    "ir_mlp_head_train_fwd": (i32, [C.POINTER(MlpHead), p, p, p, p]),
    "ir_mlp_head_train_bwd": (i32, [C.POINTER(MlpHead), p, p, p, p, p, p, p, p, p, p, p]),
}

_lib = None


class IrError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built — the product path never
    falls back to PyTorch/CPU code."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise IrError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                      f"g.build()'` (nvcc, sm_100a). There is no CPU/PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_fn = {}


def call(name, *args):
    """Invoke an ``int ir_*`` entry point; raise IrError with the library's message on failure."""
    fn = _fn.get(name)
    if fn is None:
        fn = _fn[name] = getattr(load(), name)
    rc = fn(*args)
    if rc != 0:
        raise IrError(f"{name} failed ({rc}): {load().ir_last_error().decode(errors='replace')}")
    return rc
