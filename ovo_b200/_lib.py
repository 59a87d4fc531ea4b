"""ctypes binding of libovo_b200.so (include/ovo_b200.h).  There is no CPU fallback: if the shared
library is missing or a call fails, a RuntimeError is raised."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OVO_B200_LIB") or os.path.join(_HERE, "libovo_b200.so")   # the override is an A/B measurement aid

c_void_p, c_int, c_float, c_int64 = C.c_void_p, C.c_int, C.c_float, C.c_int64


class VitCfg(C.Structure):
    _fields_ = [("image_size", c_int), ("patch_size", c_int), ("width", c_int), ("layers", c_int),
                ("heads", c_int), ("mlp_width", c_int), ("output_dim", c_int), ("ln_eps", c_float),
                ("text_ctx", c_int), ("text_width", c_int), ("text_heads", c_int), ("text_layers", c_int),
                ("text_mlp_width", c_int), ("vocab_size", c_int), ("text_output_dim", c_int)]


class BlockWeights(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("ln1_w", "ln1_b", "qkv_w", "qkv_b", "out_w", "out_b",
                                        "ln2_w", "ln2_b", "fc_w", "fc_b", "proj_w", "proj_b")]


class VitWeights(C.Structure):
    _fields_ = [("patch_w", c_void_p), ("patch_kpad", c_int), ("cls_pos0", c_void_p), ("pos", c_void_p),
                ("ln_pre_w", c_void_p), ("ln_pre_b", c_void_p), ("ln_post_w", c_void_p), ("ln_post_b", c_void_p),
                ("blocks", C.POINTER(BlockWeights)), ("pool_w", c_void_p), ("pool_b", c_void_p), ("pool_b_empty", c_void_p),
                ("tok_emb", c_void_p), ("text_pos", c_void_p), ("text_blocks", C.POINTER(BlockWeights)),
                ("ln_final_w", c_void_p), ("ln_final_b", c_void_p), ("text_proj_w", c_void_p)]


class PoolHeadWeights(C.Structure):
    _fields_ = [("heads", c_int), ("mlp_width", c_int)] + \
               [(n, c_void_p) for n in ("q", "kv_w", "kv_b", "out_w", "out_b", "ln_w", "ln_b", "fc_w", "fc_b", "proj_w", "proj_b",
                                        "vis_proj_w")]


class CropParams(C.Structure):
    _fields_ = [("embed_type", c_int), ("return_all", c_int), ("mask_res", c_int), ("bbox_margin", c_int),
                ("w_masked", c_float), ("w_global", c_float)]


class MergerLayer(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("in_w", "in_b", "out_w", "out_b", "ln1_w", "ln1_b", "ff1_w", "ff1_b", "ff2_w", "ff2_b",
                                        "ln2_w", "ln2_b")]


class MergerWeights(C.Structure):
    _fields_ = [("d_model", c_int), ("nhead", c_int), ("dim_feedforward", c_int), ("n_layers", c_int),
                ("layers", C.POINTER(MergerLayer)), ("n_linear", c_int), ("mlp_w", C.POINTER(c_void_p)),
                ("mlp_b", C.POINTER(c_void_p)), ("mlp_out", C.POINTER(c_int)), ("ln_eps", c_float)]


EMBED_TYPES = {"vanilla": 0, "fixed_weights": 1, "hovsg": 2, "adaptive_weights": 3, "concept_fusion": 4}


class VoteRow(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_matched", "n_assigned", "n_unassigned", "mode_id", "ins_id",
                                         "is_new", "area", "reserved")]


class Frame(C.Structure):
    _fields_ = [("depth_dev", c_void_p), ("h", c_int), ("w", c_int), ("seg_map_dev", c_void_p), ("H", c_int),
                ("W", c_int), ("n_masks", c_int), ("c2w", c_float * 16), ("w2c", c_float * 16), ("K", c_float * 9),
                ("match_th", c_float), ("track_th", c_int), ("depth_filter", c_int), ("has_ratio", c_int),
                ("ratio_h", c_float), ("ratio_w", c_float), ("crop_edge", c_int), ("depth_range_dev", c_void_p)]


class HieraBlock(C.Structure):
    _fields_ = [(n, c_int) for n in ("dim", "dim_out", "heads", "window", "q_pool", "grid_in")] + \
               [(n, c_void_p) for n in ("norm1_w", "norm1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "norm2_w", "norm2_b",
                                        "fc1_w", "fc1_b", "fc2_w", "fc2_b", "short_w", "short_b")]


class SamAttn(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "o_w", "o_b")]


class SamDecLayer(C.Structure):
    _fields_ = [("self_attn", SamAttn), ("t2i", SamAttn), ("i2t", SamAttn), ("norm_w", c_void_p * 4), ("norm_b", c_void_p * 4),
                ("mlp0_w", c_void_p), ("mlp0_b", c_void_p), ("mlp1_w", c_void_p), ("mlp1_b", c_void_p)]


class SamCfg(C.Structure):
    _fields_ = [("image_size", c_int), ("n_blocks", c_int), ("embed_dim", c_int), ("stage_end", c_int * 4),
                ("decoder_depth", c_int), ("trunk_ln_eps", c_float), ("max_batch", c_int)]


class SamWeights(C.Structure):
    _fields_ = [("patch_w", c_void_p), ("patch_kpad", c_int), ("patch_b", c_void_p), ("pos", c_void_p),
                ("blocks", C.POINTER(HieraBlock)),
                ("neck3_w", c_void_p), ("neck3_b", c_void_p), ("neck2_w", c_void_p), ("neck2_b", c_void_p),
                ("s1_w", c_void_p), ("s1_b", c_void_p), ("s0_w", c_void_p), ("s0_b", c_void_p),
                ("gauss", c_void_p), ("point_embed", c_void_p), ("not_a_point", c_void_p), ("dense_pe", c_void_p),
                ("no_mask_embed", c_void_p), ("out_tokens", c_void_p), ("layers", C.POINTER(SamDecLayer)),
                ("final_attn", SamAttn), ("norm_final_w", c_void_p), ("norm_final_b", c_void_p),
                ("up0_w", c_void_p), ("up0_b", c_void_p), ("up_ln_w", c_void_p), ("up_ln_b", c_void_p),
                ("up1_w", c_void_p), ("up1_b", c_void_p),
                ("hyper_w", (c_void_p * 3) * 4), ("hyper_b", (c_void_p * 3) * 4), ("iou_w", c_void_p * 3), ("iou_b", c_void_p * 3)]


class AmgParams(C.Structure):
    _fields_ = [("points_per_side", c_int), ("pred_iou_thresh", c_float), ("stability_thresh", c_float),
                ("stability_offset", c_float), ("box_nms_thresh", c_float), ("nms_iou_th", c_float),
                ("nms_score_th", c_float), ("nms_inner_th", c_float)]


# name -> (restype, argtypes); mirrors include/ovo_b200.h one to one
SIGNATURES = {
    "ovo_last_error": (C.c_char_p, []),
    "ovo_version": (c_int, []),
    "ovo_launch_count": (C.c_longlong, [c_int]),
    "ovo_encoder_create": (c_int, [C.POINTER(VitCfg), C.POINTER(VitWeights), c_int, c_int, c_int, c_int,
                                   C.POINTER(c_void_p)]),
    "ovo_encoder_destroy": (None, [c_void_p]),
    "ovo_encoder_preprocess": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, C.POINTER(c_int), c_void_p]),
    "ovo_encoder_load_pixels": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "ovo_encoder_forward": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ovo_encoder_pool_regions": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "ovo_encode_regions": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, C.POINTER(c_int), c_void_p,
                                   c_void_p]),
    "ovo_encode_text": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "ovo_encoder_set_pool_head": (c_int, [c_void_p, C.POINTER(PoolHeadWeights)]),
    "ovo_encode_images": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "ovo_mask_boxes": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ovo_encode_crops": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, C.POINTER(CropParams), c_void_p, c_void_p,
                                 c_void_p]),
    "ovo_fuse_clips": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p]),
    "ovo_merge_clips_learned": (c_int, [C.POINTER(MergerWeights), c_void_p, c_int, c_void_p, c_void_p]),
    "ovo_siglip_similarity": (c_int, [c_void_p, c_int64, c_float, c_float, c_void_p]),
    "ovo_knn": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "ovo_knn_stats": (None, [C.POINTER(c_float), C.POINTER(c_int), C.POINTER(c_int)]),
    "ovo_knn_mode": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "ovo_map_create": (c_int, [C.POINTER(c_void_p)]),
    "ovo_map_destroy": (None, [c_void_p]),
    "ovo_map_reserve": (c_int, [c_void_p, c_int64, c_int, c_int, c_int64]),
    "ovo_depth_filter": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ovo_depth_filter_batch": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ovo_depth_range": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "ovo_map_associate": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, C.POINTER(Frame), C.POINTER(c_int),
                                  C.POINTER(VoteRow), C.POINTER(c_int), c_int, c_void_p]),
    "ovo_map_associate_launch": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, C.POINTER(Frame), c_int, c_int, c_void_p]),
    "ovo_map_associate_wait": (c_int, [c_void_p, C.POINTER(c_int), C.POINTER(VoteRow), C.POINTER(c_int)]),
    "ovo_map_vote": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, C.POINTER(Frame), c_int, c_void_p, c_int, c_void_p]),
    "ovo_map_apply": (c_int, [c_void_p, c_void_p, c_void_p, C.POINTER(c_int), C.POINTER(VoteRow), C.POINTER(c_int), c_void_p]),
    "ovo_map_get_matches": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "ovo_map_fuse_dense": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_int,
                                   c_void_p]),
    "ovo_map_fuse_dense_batch": (c_int, [c_void_p, C.POINTER(c_int), c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int,
                                         c_void_p, c_int, c_void_p]),
    "ovo_map_associate_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, C.POINTER(Frame), c_int, C.POINTER(c_int), C.POINTER(c_int),
                                        C.POINTER(VoteRow), c_int, C.POINTER(c_int), c_void_p, c_void_p]),
    "ovo_map_batch_begin": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, C.POINTER(Frame), c_int, C.POINTER(c_int), c_int, c_void_p,
                                    c_int64, c_void_p]),
    "ovo_map_batch_vote": (c_int, [c_void_p, c_int, c_void_p, C.POINTER(c_void_p), C.POINTER(c_int), c_void_p]),
    "ovo_map_batch_decide": (c_int, [c_void_p, c_int, c_void_p]),
    "ovo_map_batch_info": (c_int, [c_void_p, c_int, C.POINTER(c_int), C.POINTER(c_void_p)]),
    "ovo_xchg_create": (c_int, [c_int, c_int, c_int, c_int64, C.POINTER(c_void_p)]),
    "ovo_xchg_ipc_handle": (c_int, [c_void_p, c_void_p]),
    "ovo_xchg_open_peers": (c_int, [c_void_p, c_void_p]),
    "ovo_xchg_exchange": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    "ovo_xchg_destroy": (None, [c_void_p]),
    "ovo_route_pack": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "ovo_map_associate_batch_sharded": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, C.POINTER(Frame), c_int, C.POINTER(c_int),
                                                C.POINTER(c_int), C.POINTER(VoteRow), c_int, C.POINTER(c_int), c_void_p, c_void_p]),
    "ovo_map_batch_end": (c_int, [c_void_p, c_void_p, C.POINTER(c_int), C.POINTER(VoteRow), c_int, C.POINTER(c_int), c_void_p, c_void_p]),
    "ovo_bank_add_views": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "ovo_bank_update_mean": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "ovo_query_dense": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "ovo_query_instances": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "ovo_merge_masks": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "ovo_fuse_views": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ovo_text_bank": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ovo_map_integrate": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int, c_int,
                                  C.POINTER(c_float), C.POINTER(c_float), C.POINTER(c_float), c_float, c_int, c_int, c_int,
                                  C.POINTER(c_int), c_void_p]),
    "ovo_mask_nms": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p, c_void_p]),
    "ovo_mask2segmap": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ovo_sam_create": (c_int, [C.POINTER(SamCfg), C.POINTER(SamWeights), c_int, c_int, c_int, C.POINTER(c_void_p)]),
    "ovo_sam_destroy": (None, [c_void_p]),
    "ovo_sam_set_image": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "ovo_sam_set_images": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ovo_sam_select_image": (c_int, [c_void_p, c_int]),
    "ovo_sam_generate_batch": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, C.POINTER(AmgParams), c_void_p, c_void_p, c_int,
                                       C.POINTER(c_int), c_void_p]),
    "ovo_sam_set_pixels": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "ovo_sam_predict": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "ovo_sam_postprocess": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, C.POINTER(AmgParams), c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_int, C.POINTER(c_int), c_void_p]),
    "ovo_sam_override_logits": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "ovo_sam_generate": (c_int, [c_void_p, c_void_p, c_int, c_int, C.POINTER(AmgParams), c_void_p, c_void_p, c_int, C.POINTER(c_int),
                                 c_void_p]),
    "ovo_classify": (c_int, [c_void_p, c_int64, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "ovo_profile_begin": (None, []),
    "ovo_profile_report": (c_int, [c_int, C.POINTER(c_float), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(c_int)]),
    "ovo_set_gemm_cluster": (None, [c_int]),
    "ovo_attn_trace": (None, [c_void_p]),
    "ovo_gemm_bench": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, C.POINTER(c_float), c_void_p]),
    "ovo_gemm_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                              c_int, c_void_p]),
}

_lib = None


def lib():
    """The loaded library; raises if it has not been built (python -m ovo_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: build it with `python -m ovo_b200.build` "
                               "(there is no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(code: int, what: str = "") -> int:
    if code < 0:
        msg = lib().ovo_last_error()
        raise RuntimeError(f"ovo_b200 {what} failed ({code}): {msg.decode() if msg else ''}")
    return code


def ptr(t):
    """Device (or host) pointer of a torch tensor, or None."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    """torch's current stream of `device` (default: the current device) as a cudaStream_t.  Kernels launch on the CURRENT
    device, so a handle that lives on another device is refused instead of being driven on the wrong stream."""
    import torch
    if device is not None:
        idx = torch.device(device).index
        if idx is not None and idx != torch.cuda.current_device():
            raise RuntimeError(f"ovo_b200: this object lives on cuda:{idx} but the current device is cuda:{torch.cuda.current_device()} "
                               "(one process per GPU: call torch.cuda.set_device first, or wrap the call in torch.cuda.device(...))")
        return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


PROF_CLASSES = ("gemm", "attention", "layernorm", "preprocess", "pool", "associate", "fuse_dense", "query", "other")


def profile_begin():
    lib().ovo_profile_begin()


def profile_report():
    """-> {class: dict(ms, flops, bytes, launches)} for the launches since profile_begin()."""
    n = len(PROF_CLASSES)
    ms, fl, by, ct = (c_float * n)(), (C.c_double * n)(), (C.c_double * n)(), (c_int * n)()
    check(lib().ovo_profile_report(n, ms, fl, by, ct), "ovo_profile_report")
    return {PROF_CLASSES[i]: dict(ms=ms[i], flops=fl[i], bytes=by[i], launches=ct[i]) for i in range(n)}
