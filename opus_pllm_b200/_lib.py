"""ctypes binding of libopus_b200.so (the C ABI declared in include/opus_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``python -m opus_pllm_b200.build``. There is no CPU
fallback: if the shared object is missing, or a compute entry point fails, the caller gets an exception.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# OPUS_B200_LIB points at another build of the same ABI (A/B measurements of kernel variants); default = the in-tree build
LIB_PATH = os.environ.get("OPUS_B200_LIB") or os.path.join(_HERE, "libopus_b200.so")

OK = 0
EPI_BF16, EPI_BF16_GELU, EPI_RES_F32, EPI_RES_BF16, EPI_SWIGLU, EPI_PARTIAL_F32, EPI_F32, EPI_BF16_RELU = range(8)
ARCH_LLAMA, ARCH_OPT = 0, 1
ABI_VERSION = 4

c_void_p, c_int, c_float, c_size_t, c_longlong = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_longlong


class OpusError(RuntimeError):
    pass


# ---------------------------------------------------------------------------------------------- structs
class Esm2Layer(C.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "ln1_g", "ln1_b", "wqkv", "bqkv", "wo", "bo", "ln2_g", "ln2_b", "w1", "b1", "w2", "b2")]


class Esm2Model(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("dim", C.c_int32), ("n_heads", C.c_int32), ("ffn_dim", C.c_int32),
                ("vocab", C.c_int32), ("rope_max_pos", C.c_int32), ("ln_eps", c_float),
                ("embed", c_void_p), ("layers", C.POINTER(Esm2Layer)), ("lnf_g", c_void_p), ("lnf_b", c_void_p),
                ("rope_cos", c_void_p), ("rope_sin", c_void_p)]


class Esm2Workspace(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("x", "xn", "qkv", "attn", "ffn")]


class ProjectorModel(C.Structure):
    _fields_ = [("in_dim", C.c_int32), ("cstp_dim", C.c_int32), ("hidden_dim", C.c_int32),
                ("w_cstp", c_void_p), ("b_cstp", c_void_p), ("w0", c_void_p), ("b0", c_void_p),
                ("w2", c_void_p), ("b2", c_void_p)]


class LlamaLayer(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("ln1_w", "wqkv", "wo", "ln2_w", "wgu", "wdown", "bqkv",
                                        "ln1_g", "ln1_b", "ln2_g", "ln2_b", "bo", "b1", "b2")]   # OPT family tail


class LlamaModel(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("dim", C.c_int32), ("n_q_heads", C.c_int32), ("n_kv_heads", C.c_int32),
                ("head_dim", C.c_int32), ("ffn_dim", C.c_int32), ("vocab", C.c_int32), ("rope_max_pos", C.c_int32),
                ("rms_eps", c_float), ("embed", c_void_p), ("layers", C.POINTER(LlamaLayer)),
                ("norm_w", c_void_p), ("lm_head", c_void_p), ("rope_cos", c_void_p), ("rope_sin", c_void_p),
                ("arch", C.c_int32), ("opt_act", C.c_int32), ("pos_embed", c_void_p), ("pos_rows", C.c_int32),
                ("head_dim_real", C.c_int32), ("norm_g", c_void_p), ("norm_b", c_void_p)]


class KvCache(C.Structure):
    _fields_ = [("k", c_void_p), ("v", c_void_p), ("num_blocks", C.c_int32), ("block_size", C.c_int32)]


class LlamaWorkspace(C.Structure):
    _fields_ = [("h", c_void_p), ("xn", c_void_p), ("qkv", c_void_p), ("attn", c_void_p), ("act", c_void_p),
                ("partial", c_void_p), ("partial_bytes", c_size_t), ("last_h", c_void_p), ("logits", c_void_p)]


class DecodeState(C.Structure):
    _fields_ = [("next_tok", c_void_p), ("ctx_len", c_void_p), ("pos", c_void_p), ("slot", c_void_p),
                ("block_table", c_void_p), ("max_blocks", C.c_int32), ("finished", c_void_p),
                ("n_unfinished", c_void_p), ("step", c_void_p), ("out_ids", c_void_p), ("out_ld", C.c_int32),
                ("eos_ids", c_void_p), ("n_eos", C.c_int32), ("pad_id", C.c_int32),
                ("seed", C.c_uint64), ("temperature", c_float), ("top_p", c_float), ("do_sample", C.c_int32),
                ("reserved_", C.c_int32),
                ("stop_seqs", c_void_p), ("stop_lens", c_void_p), ("n_stop", C.c_int32), ("stop_ld", C.c_int32),
                ("seed_ptr", c_void_p)]


# ---------------------------------------------------------------------------------------------- signatures
_P = c_void_p
_SIGNATURES = {
    "opus_abi_version": (c_int, []),
    "opus_last_error": (C.c_char_p, []),
    "opus_device_check": (c_int, []),
    "opus_ctx_create": (c_int, [C.POINTER(c_void_p)]),
    "opus_ctx_destroy": (c_int, [c_void_p]),
    "opus_ctx_set_current": (c_int, [c_void_p]),
    "opus_gemm_bf16": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P, _P, c_int,
                               c_int, c_int, _P]),
    "opus_gemm_bf16_fused": (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P, _P, c_int, c_int,
                                     c_int, _P, c_int, _P, c_int, c_int, _P, c_float, _P]),
    "opus_gemm_suggest_split_k": (c_int, [c_int, c_int, c_int, c_int]),
    "opus_splitk_reduce_bf16": (c_int, [_P, c_int, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "opus_esm_embed": (c_int, [_P, _P, _P, _P, c_int, c_int, _P]),
    "opus_layernorm_f32_bf16": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_float, _P]),
    "opus_rmsnorm_bf16": (c_int, [_P, _P, c_int, _P, _P, _P, _P, c_int, c_int, c_float, _P]),
    "opus_rope_esm_bf16": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, _P]),
    "opus_rope_llama_kvappend_bf16": (c_int, [_P, _P, c_int, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int,
                                              c_int, c_int, _P]),
    "opus_final_ln_meanpool": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_float, _P]),
    "opus_l2norm_f32_bf16": (c_int, [_P, _P, c_int, c_int, _P]),
    "opus_splice_gather_bf16": (c_int, [_P, _P, _P, _P, c_int, c_int, _P]),
    "opus_argmax_eos": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_int, c_int, _P, _P, c_int, c_int, _P, _P]),
    "opus_sample_top_p": (c_int, [_P, c_int, c_int, c_int, c_float, c_float, C.c_uint64, _P, _P, c_int, c_int, _P, _P,
                                  c_int, c_int, _P, _P, _P]),
    "opus_cross_entropy_bf16": (c_int, [_P, c_int, c_int, _P, _P, c_int, _P]),
    "opus_embed_gather_bf16": (c_int, [_P, _P, _P, c_int, c_int, _P]),
    "opus_lora_merge_bf16": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_float, _P]),
    "opus_stop_sequences": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, c_int, c_int, _P, _P, _P]),
    "opus_layernorm_bf16": (c_int, [_P, _P, c_int, _P, _P, _P, _P, _P, _P, c_int, c_int, c_float, _P]),
    "opus_add_pos_embed_bf16": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "opus_attn_varlen_bf16": (c_int, [_P, c_int, _P, c_int, _P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int,
                                      c_int, c_int, c_int, c_float, _P]),
    "opus_attn_decode_paged_bf16": (c_int, [_P, c_int, _P, _P, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int,
                                            c_int, c_float, _P]),
    "opus_esm2_forward": (c_int, [C.POINTER(Esm2Model), C.POINTER(Esm2Workspace), _P, _P, _P, _P, c_int, c_int,
                                  c_int, _P, _P, _P, _P]),
    "opus_projector_forward": (c_int, [C.POINTER(ProjectorModel), _P, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "opus_llama_prefill": (c_int, [C.POINTER(LlamaModel), C.POINTER(KvCache), C.POINTER(LlamaWorkspace), _P, _P, _P,
                                   _P, _P, c_int, c_int, c_int, _P]),
    "opus_llama_decode_step": (c_int, [C.POINTER(LlamaModel), C.POINTER(KvCache), C.POINTER(LlamaWorkspace),
                                       C.POINTER(DecodeState), c_int, _P]),
    "opus_llama_select": (c_int, [C.POINTER(LlamaModel), C.POINTER(LlamaWorkspace), C.POINTER(DecodeState), c_int,
                                  _P]),
    "opus_llama_decode_loop": (c_int, [C.POINTER(LlamaModel), C.POINTER(KvCache), C.POINTER(LlamaWorkspace),
                                       C.POINTER(DecodeState), c_int, c_int, c_int, c_int, _P]),
    "opus_release_graphs": (c_int, []),
    "opus_set_tunable": (c_int, [C.c_char_p, c_int]),
    "opus_chain_trace": (c_int, [c_int, _P, c_int]),
    "opus_trace_begin": (c_int, [_P]),
    "opus_trace_end": (c_int, [C.c_char_p, c_int]),
    "opus_launch_count": (c_longlong, [c_int]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load() -> C.CDLL:
    """Load libopus_b200.so (once). Raises OpusError when it has not been built — never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OpusError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). opus_pllm_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.opus_abi_version() != ABI_VERSION:
        raise OpusError("libopus_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> int:
    if rc < 0:
        msg = load().opus_last_error().decode("utf-8", "replace")
        raise OpusError(f"{what or 'opus_b200 call'} failed with code {rc}: {msg}")
    return rc
