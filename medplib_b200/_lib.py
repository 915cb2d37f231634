"""ctypes binding of libmedplib_b200.so (C ABI: include/medplib_b200.h). Fails loudly when the library is missing."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmedplib_b200.so")

MPL_OK = 0
_ERR = {-1: "MPL_ERR_ARG", -2: "MPL_ERR_ALIGN", -3: "MPL_ERR_DRIVER", -4: "MPL_ERR_CUDA", -5: "MPL_ERR_UNSUPPORTED"}

ACT_NONE, ACT_GELU, ACT_QUICK_GELU, ACT_RELU, ACT_SILU, ACT_SIGMOID = range(6)
DT_BF16, DT_F32 = 0, 1

c_void_p, c_int, c_ll, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float


class MplError(RuntimeError):
    pass


class GemmArgs(ctypes.Structure):
    """mpl_gemm_args (include/medplib_b200.h)."""
    _fields_ = [
        ("A", c_void_p), ("lda", c_ll),
        ("B", c_void_p * 3), ("B2", c_void_p), ("ldb", c_ll),
        ("C", c_void_p * 3), ("ldc", c_ll),
        ("bias", c_void_p * 3), ("residual", c_void_p), ("ldr", c_ll),
        ("row_scale", c_void_p), ("m_dev", c_void_p),
        ("M", c_int), ("N", c_int), ("K", c_int),
        ("nb", c_int), ("act", c_int), ("out_dtype", c_int), ("tile_n", c_int),
        ("ln_weight", c_void_p), ("ln_eps", c_float),
        ("lora_u", c_void_p * 2), ("lora_b", c_void_p * 2), ("lora_scale", c_float * 2), ("lora_u_f32", c_int * 2),
        ("lora_mat", c_int * 2), ("lora_r", c_int), ("ext_a", c_void_p), ("ext_b", c_void_p),
        ("dual_g", c_void_p), ("dual_u", c_void_p), ("silu_bwd_g", c_void_p), ("silu_bwd_u", c_void_p),
    ]


class GroupedGemmArgs(ctypes.Structure):
    """mpl_grouped_gemm_args (include/medplib_b200.h)."""
    _fields_ = [
        ("A", c_void_p), ("lda", c_ll), ("a_group_stride", c_ll),
        ("B", c_void_p * 8), ("B2", c_void_p * 8), ("ldb", c_ll),
        ("C", c_void_p), ("ldc", c_ll), ("c_group_stride", c_ll),
        ("residual", c_void_p), ("ldr", c_ll), ("m_dev", c_void_p), ("a_row_map", c_void_p),
        ("row_map", c_void_p), ("row_gate", c_void_p), ("map_group_stride", c_ll),
        ("groups", c_int), ("M", c_int), ("N", c_int), ("K", c_int), ("act", c_int), ("out_dtype", c_int),
        ("m_dev_stable", c_int), ("m_total_hint", c_int), ("tile_n", c_int),
    ]


class AttnArgs(ctypes.Structure):
    """mpl_attn_args (include/medplib_b200.h)."""
    _fields_ = [
        ("q", c_void_p), ("k", c_void_p), ("v", c_void_p), ("o", c_void_p),
        ("q_stride", c_ll * 3), ("k_stride", c_ll * 3), ("v_stride", c_ll * 3), ("o_stride", c_ll * 3),
        ("B", c_int), ("H", c_int), ("Tq", c_int), ("Tk", c_int), ("head_dim", c_int),
        ("scale", c_float), ("causal", c_int),
        ("kv_mask", c_void_p), ("kv_mask_stride", c_ll), ("rel_h", c_void_p), ("rel_w", c_void_p),
        ("rel_kh", c_int), ("rel_kw", c_int), ("tk_dev", c_void_p), ("scratch", c_void_p), ("scratch_bytes", c_ll),
        ("lse", c_void_p),
    ]


class AttnBwdArgs(ctypes.Structure):
    """mpl_attn_bwd_args (include/medplib_b200.h)."""
    _fields_ = [
        ("q", c_void_p), ("k", c_void_p), ("v", c_void_p), ("o", c_void_p), ("d_o", c_void_p),
        ("q_stride", c_ll * 3), ("k_stride", c_ll * 3), ("v_stride", c_ll * 3), ("o_stride", c_ll * 3),
        ("lse", c_void_p), ("delta", c_void_p), ("dq_f32", c_void_p), ("dk", c_void_p), ("dv", c_void_p),
        ("dk_stride", c_ll * 3), ("dv_stride", c_ll * 3),
        ("B", c_int), ("H", c_int), ("T", c_int), ("head_dim", c_int), ("scale", c_float), ("causal", c_int),
        ("kv_mask", c_void_p), ("kv_mask_stride", c_ll),
    ]


class MoeRouteArgs(ctypes.Structure):
    """mpl_moe_route_args."""
    _fields_ = [
        ("h", c_void_p), ("ldh", c_ll), ("wg", c_void_p), ("noise", c_void_p),
        ("S", c_int), ("D", c_int), ("E", c_int), ("k", c_int), ("capacity", c_int),
        ("logits", c_void_p), ("gates", c_void_p), ("expert", c_void_p), ("gate", c_void_p), ("slot", c_void_p),
        ("kept", c_void_p), ("exp_counts", c_void_p), ("l_aux", c_void_p),
    ]


MAX_EXPERTS = 8


class LlamaLayer(ctypes.Structure):
    """mpl_llama_layer."""
    _fields_ = [
        ("input_ln", c_void_p), ("wq", c_void_p), ("wk", c_void_p), ("wv", c_void_p), ("wo", c_void_p),
        ("post_ln", c_void_p), ("wg", c_void_p), ("n_experts", c_int),
        ("w_gate", c_void_p * MAX_EXPERTS), ("w_up", c_void_p * MAX_EXPERTS), ("w_down", c_void_p * MAX_EXPERTS),
    ]


class LlamaModel(ctypes.Structure):
    """mpl_llama_model."""
    _fields_ = [
        ("n_layers", c_int), ("hidden", c_int), ("n_heads", c_int), ("ffn", c_int), ("rms_eps", c_float),
        ("top_k", c_int), ("capacity_factor", c_float), ("min_capacity", c_int),
        ("layers", ctypes.POINTER(LlamaLayer)), ("final_norm", c_void_p),
        ("rope_cos", c_void_p), ("rope_sin", c_void_p), ("rope_len", c_int),
    ]


class LlamaIO(ctypes.Structure):
    """mpl_llama_io."""
    _fields_ = [
        ("x", c_void_p), ("out_norm", c_void_p), ("hidden_states", ctypes.POINTER(c_void_p)),
        ("B", c_int), ("T", c_int), ("past_len", c_int),
        ("k_cache", c_void_p), ("v_cache", c_void_p), ("Tmax", c_int),
        ("kv_mask", c_void_p), ("kv_mask_stride", c_ll), ("pos_dev", c_void_p), ("tk_dev", c_void_p),
        ("attn_scratch", c_void_p), ("attn_scratch_bytes", c_ll), ("decode_plan", c_void_p),
        ("moe_noise", ctypes.POINTER(c_void_p)), ("gate_logits", c_void_p), ("l_aux", c_void_p),
        ("exp_counts", c_void_p), ("workspace", c_void_p), ("workspace_bytes", c_ll), ("rope_pos", c_void_p),
    ]


class ClipLayer(ctypes.Structure):
    """mpl_clip_layer."""
    _fields_ = [(n, c_void_p) for n in (
        "ln1_w", "ln1_b", "wq", "bq", "wk", "bk", "wv", "bv", "wo", "bo", "ln2_w", "ln2_b",
        "fc1_w", "fc1_b", "fc2_w", "fc2_b")]


class ClipModel(ctypes.Structure):
    """mpl_clip_model."""
    _fields_ = [
        ("n_layers", c_int), ("hidden", c_int), ("n_heads", c_int), ("mlp", c_int), ("image_size", c_int),
        ("patch", c_int), ("k_pad", c_int), ("ln_eps", c_float),
        ("patch_w", c_void_p), ("cls", c_void_p), ("pos", c_void_p), ("pre_ln_w", c_void_p), ("pre_ln_b", c_void_p),
        ("layers", ctypes.POINTER(ClipLayer)),
    ]


class SamBlock(ctypes.Structure):
    """mpl_sam_block."""
    _fields_ = [
        ("ln1_w", c_void_p), ("ln1_b", c_void_p), ("qkv_w", c_void_p), ("qkv_b", c_void_p), ("proj_w", c_void_p),
        ("proj_b", c_void_p), ("rel_pos_h", c_void_p), ("rel_pos_w", c_void_p), ("window", c_int),
        ("ln2_w", c_void_p), ("ln2_b", c_void_p), ("lin1_w", c_void_p), ("lin1_b", c_void_p), ("lin2_w", c_void_p),
        ("lin2_b", c_void_p), ("ad_ch0", c_void_p), ("ad_ch2", c_void_p), ("ad_conv", c_void_p),
        ("ad_convt", c_void_p), ("ad_norm_w", c_void_p), ("ad_norm_b", c_void_p),
    ]


class SamEncoder(ctypes.Structure):
    """mpl_sam_encoder."""
    _fields_ = [
        ("depth", c_int), ("hidden", c_int), ("n_heads", c_int), ("mlp", c_int), ("image_size", c_int),
        ("patch", c_int), ("out_chans", c_int),
        ("patch_w", c_void_p), ("patch_b", c_void_p), ("pos_embed", c_void_p), ("blocks", ctypes.POINTER(SamBlock)),
        ("neck0_w", c_void_p), ("neck1_w", c_void_p), ("neck1_b", c_void_p), ("neck2_w", c_void_p),
        ("neck3_w", c_void_p), ("neck3_b", c_void_p),
    ]


class SamAttn(ctypes.Structure):
    """mpl_sam_attn."""
    _fields_ = [(n, c_void_p) for n in ("q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "o_w", "o_b")]


class SamTwoWayLayer(ctypes.Structure):
    """mpl_sam_twoway_layer."""
    _fields_ = [("self_attn", SamAttn), ("t2i", SamAttn), ("i2t", SamAttn)] + [(n, c_void_p) for n in (
        "n1_w", "n1_b", "n2_w", "n2_b", "n3_w", "n3_b", "n4_w", "n4_b", "lin1_w", "lin1_b", "lin2_w", "lin2_b")]


class SamMaskDecoder(ctypes.Structure):
    """mpl_sam_mask_decoder."""
    _fields_ = [
        ("dim", c_int), ("n_heads", c_int), ("mlp", c_int), ("depth", c_int), ("n_mask_tokens", c_int),
        ("grid", c_int),
        ("iou_token", c_void_p), ("mask_tokens", c_void_p), ("no_mask", c_void_p), ("dense_pe", c_void_p),
        ("layers", ctypes.POINTER(SamTwoWayLayer)), ("final_attn", SamAttn), ("nf_w", c_void_p), ("nf_b", c_void_p),
        ("up0_w", c_void_p), ("up0_b", c_void_p), ("up_ln_w", c_void_p), ("up_ln_b", c_void_p),
        ("up1_w", c_void_p), ("up1_b", c_void_p), ("shuffle_idx", c_void_p),
        ("hyper_w", c_void_p * 3), ("hyper_b", c_void_p * 3), ("iou_w", c_void_p * 3), ("iou_b", c_void_p * 3),
    ]


# Every symbol include/medplib_b200.h declares; tests/test_abi.py checks the built library exports all of them.
EXPORTS = [
    "mpl_version", "mpl_device_info", "mpl_launch_count", "mpl_gemm_bf16", "mpl_profile_gemm", "mpl_profile_gemm_read", "mpl_profile_decode", "mpl_profile_decode_read", "mpl_skinny_gemm_bf16", "mpl_linear_bf16", "mpl_grouped_gemm_bf16", "mpl_moe_route_small", "mpl_llama_decode_plan_bytes", "mpl_llama_decode_plan_build", "mpl_debug_decode_timing",
    "mpl_rmsnorm", "mpl_layernorm", "mpl_pool_layernorm", "mpl_attention",
    "mpl_moe_route", "mpl_moe_dispatch", "mpl_moe_combine", "mpl_rope_kv", "mpl_gather_rows", "mpl_argmax_f32",
    "mpl_im2col_patch", "mpl_im2col_nhwc", "mpl_clip_embed", "mpl_sam_relpos", "mpl_col_mean",
    "mpl_convt4s2_col2im", "mpl_add", "mpl_bilinear_resize", "mpl_region_sample_mean",
    "mpl_llama_workspace_bytes", "mpl_llama_forward", "mpl_clip_workspace_bytes", "mpl_clip_forward",
    "mpl_sam_encoder_workspace_bytes", "mpl_sam_encoder_forward", "mpl_sam_mask_decoder_workspace_bytes",
    "mpl_sam_mask_decoder_forward",
    "mpl_transpose_bf16", "mpl_lora_down", "mpl_lora_up_add", "mpl_rank_wgrad", "mpl_rmsnorm_bwd", "mpl_silu_mul",
    "mpl_silu_mul_bwd", "mpl_attention_bwd", "mpl_rope_bwd", "mpl_moe_combine_bwd", "mpl_moe_router_bwd", "mpl_ce_fwd",
    "mpl_ce_bwd", "mpl_scatter_add_rows", "mpl_sumsq_f32", "mpl_adamw", "mpl_adamw_multi", "mpl_mask_losses",
    "mpl_gemm_small", "mpl_col_sum", "mpl_layernorm_bwd", "mpl_act_fwd", "mpl_act_bwd", "mpl_attn_small_bwd",
    "mpl_bilinear_resize_bwd", "mpl_mask_losses_bwd", "mpl_mask_scale_bf16", "mpl_token_pool", "mpl_token_pool_bwd", "mpl_col2im_nhwc", "mpl_zero_tail_rows", "mpl_lora_down_ext", "mpl_lora_pack",
    "mpl_preprocess_images", "mpl_preprocess_band_rows",
    "mpl_geo_point_table", "mpl_geo_fps", "mpl_geo_knn", "mpl_geo_group", "mpl_geo_ln_pool",
]
_LL_RET = {"mpl_launch_count", "mpl_llama_workspace_bytes", "mpl_clip_workspace_bytes", "mpl_sam_encoder_workspace_bytes",
           "mpl_sam_mask_decoder_workspace_bytes"}

_lib = None


def load():
    """Load the shared library once. Raises MplError if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MplError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C medplib_b200/csrc`. medplib_b200 has no CPU / eager fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name in EXPORTS:
        getattr(lib, name).restype = c_ll if name in _LL_RET else c_int
    _lib = lib
    return lib


def check(rc, what):
    if rc != MPL_OK:
        raise MplError(f"{what} failed: {_ERR.get(rc, rc)}")
