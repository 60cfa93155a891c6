"""Tensor-level wrappers over the C ABI: PyTorch supplies device memory and the current stream, nothing else.

Every function validates device/dtype/contiguity, calls exactly one `opus_*` entry point and raises on failure.
"""
from __future__ import annotations

import torch

from . import _lib as L

BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t) -> int | None:
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, dtype, name: str, contiguous: bool = True):
    if not t.is_cuda:
        raise L.OpusError(f"{name}: expected a CUDA tensor (opus_pllm_b200 has no CPU path)")
    if t.dtype != dtype:
        raise L.OpusError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if contiguous and not t.is_contiguous():
        raise L.OpusError(f"{name}: expected a contiguous tensor")
    return t


XFER = {"h2d_bytes": 0, "d2h_bytes": 0}  # host<->device traffic issued by the package (bench.py reports it)


def h2d(arr, device) -> torch.Tensor:
    """numpy array / CPU tensor -> device tensor through pinned memory, asynchronous on the current stream."""
    import numpy as np
    t = torch.from_numpy(np.ascontiguousarray(arr)) if not isinstance(arr, torch.Tensor) else arr.contiguous()
    XFER["h2d_bytes"] += t.numel() * t.element_size()
    if not t.is_pinned():
        t = t.pin_memory()
    return t.to(device, non_blocking=True)


def d2h(t: torch.Tensor) -> torch.Tensor:
    if t.is_cuda:
        XFER["d2h_bytes"] += t.numel() * t.element_size()
    return t.cpu()


def device_check():
    L.check(L.load().opus_device_check(), "opus_device_check")


def launch_count(reset: bool = False) -> int:
    return int(L.load().opus_launch_count(1 if reset else 0))


# ------------------------------------------------------------------------------------------------ GEMM
def gemm_fused(x: torch.Tensor, w: torch.Tensor, *, epilogue: int = L.EPI_BF16, bias: torch.Tensor | None = None,
               residual: torch.Tensor | None = None, out: torch.Tensor | None = None, split_k: int = 0,
               splitk_fixup: bool = False, sumsq_out: torch.Tensor | None = None, norm_sumsq: torch.Tensor | None = None,
               norm_gamma: torch.Tensor | None = None, norm_eps: float = 1e-5) -> torch.Tensor:
    """Swap-AB y = epilogue(f(x) @ w.T) with the decode-step fusions of opus_gemm_bf16_fused: in-kernel split-K reduce,
    per-slab sums of squares of an EPI_RES_BF16 result (sumsq_out fp32 [N_out/32, ld >= rows]), RMSNorm of the activation
    operand on load (norm_sumsq fp32 [slabs, ld >= rows], norm_gamma bf16 [K]; x is then the raw residual stream)."""
    _chk(x, BF16, "x"); _chk(w, BF16, "w")
    rows, K = x.shape
    N = w.shape[0]
    n_out = N // 2 if epilogue == L.EPI_SWIGLU else N
    if out is None:
        out = (torch.empty((max(split_k, 1), rows, N), dtype=F32, device=x.device) if epilogue == L.EPI_PARTIAL_F32
               else torch.empty((rows, n_out), dtype=BF16, device=x.device))
    rc = L.load().opus_gemm_bf16_fused(
        _p(w), w.stride(0), _p(x), x.stride(0), N, rows, K, epilogue, _p(out), out.stride(-2), _p(bias), _p(residual),
        0 if residual is None else residual.stride(0), split_k, int(splitk_fixup),
        _p(sumsq_out), 0 if sumsq_out is None else sumsq_out.stride(0),
        _p(norm_sumsq), 0 if norm_sumsq is None else norm_sumsq.shape[0], 0 if norm_sumsq is None else norm_sumsq.stride(0),
        _p(norm_gamma), norm_eps, _stream())
    L.check(rc, "opus_gemm_bf16_fused")
    return out



def gemm(x: torch.Tensor, w: torch.Tensor, *, epilogue: int = L.EPI_BF16, bias: torch.Tensor | None = None,
         residual: torch.Tensor | None = None, out: torch.Tensor | None = None, transposed: bool | None = None,
         split_k: int = 0, block_n: int = 0) -> torch.Tensor:
    """y = epilogue(x @ w.T): x bf16 [rows, K], w bf16 [N, K] (torch Linear layout).

    transposed=None picks the weight-streaming swap-AB form for rows <= 256. For EPI_PARTIAL_F32 the result is
    fp32 [split_k, rows, N].
    """
    _chk(x, BF16, "x"); _chk(w, BF16, "w")
    rows, K = x.shape
    N, K2 = w.shape
    if K != K2:
        raise L.OpusError("gemm: K mismatch")
    if transposed is None:
        transposed = rows <= 256
    n_out = N // 2 if epilogue == L.EPI_SWIGLU else N
    if out is None:
        if epilogue == L.EPI_PARTIAL_F32:
            out = torch.empty((max(split_k, 1), rows, N), dtype=F32, device=x.device)
        elif epilogue in (L.EPI_RES_F32, L.EPI_F32):
            out = torch.empty((rows, n_out), dtype=F32, device=x.device)
        else:
            out = torch.empty((rows, n_out), dtype=BF16, device=x.device)
    if bias is not None:
        _chk(bias, F32, "bias")
    ldr = 0
    if residual is not None:
        _chk(residual, F32 if epilogue == L.EPI_RES_F32 else BF16, "residual")
        ldr = residual.stride(0)
    ldo = out.stride(-2)
    lib = L.load()
    if transposed:
        rc = lib.opus_gemm_bf16(_p(w), w.stride(0), _p(x), x.stride(0), N, rows, K, 1, epilogue, _p(out), ldo,
                                _p(bias), _p(residual), ldr, split_k, block_n, _stream())
    else:
        rc = lib.opus_gemm_bf16(_p(x), x.stride(0), _p(w), w.stride(0), rows, N, K, 0, epilogue, _p(out), ldo,
                                _p(bias), _p(residual), ldr, split_k, block_n, _stream())
    L.check(rc, "opus_gemm_bf16")
    return out


def splitk_reduce(partial: torch.Tensor, bias: torch.Tensor | None = None, gelu: bool = False) -> torch.Tensor:
    _chk(partial, F32, "partial")
    s, rows, cols = partial.shape
    out = torch.empty((rows, cols), dtype=BF16, device=partial.device)
    L.check(L.load().opus_splitk_reduce_bf16(_p(partial), s, _p(bias), _p(out), rows, cols, cols, int(gelu),
                                             _stream()), "opus_splitk_reduce_bf16")
    return out


# ------------------------------------------------------------------------------------------------ bandwidth ops
def esm_embed(tok: torch.Tensor, scale: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    _chk(tok, I32, "tok"); _chk(scale, F32, "scale"); _chk(table, F32, "table")
    x = torch.empty((tok.numel(), table.shape[1]), dtype=F32, device=tok.device)
    L.check(L.load().opus_esm_embed(_p(tok), _p(scale), _p(table), _p(x), tok.numel(), table.shape[1], _stream()),
            "opus_esm_embed")
    return x


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
              delta: torch.Tensor | None = None) -> torch.Tensor:
    """y = LN(x (+ delta)); with delta, x is updated in place (x += delta)."""
    _chk(x, F32, "x"); _chk(gamma, F32, "gamma"); _chk(beta, F32, "beta")
    if delta is not None:
        _chk(delta, BF16, "delta")
    y = torch.empty(x.shape, dtype=BF16, device=x.device)
    L.check(L.load().opus_layernorm_f32_bf16(_p(x), _p(delta), _p(gamma), _p(beta), _p(y), x.shape[0], x.shape[1], eps,
                                             _stream()), "opus_layernorm_f32_bf16")
    return y


def rmsnorm(x: torch.Tensor | None, w: torch.Tensor | None, eps: float = 1e-5, *, partial: torch.Tensor | None = None,
            residual: torch.Tensor | None = None, h_out: torch.Tensor | None = None, normalise: bool = True):
    """Returns y (bf16) — and writes h = x(+residual) into h_out when given."""
    src = x if x is not None else partial
    rows, cols = src.shape[-2], src.shape[-1]
    if x is not None:
        _chk(x, BF16, "x")
    else:
        _chk(partial, F32, "partial")
    y = torch.empty((rows, cols), dtype=BF16, device=src.device) if normalise else None
    L.check(L.load().opus_rmsnorm_bf16(_p(x), _p(partial), 0 if partial is None else partial.shape[0], _p(residual),
                                       _p(h_out), _p(w), _p(y), rows, cols, eps, _stream()), "opus_rmsnorm_bf16")
    return y


def layernorm_bf16(x: torch.Tensor | None, gamma: torch.Tensor | None, beta: torch.Tensor | None, eps: float = 1e-5, *,
                   partial: torch.Tensor | None = None, red_bias: torch.Tensor | None = None,
                   residual: torch.Tensor | None = None, h_out: torch.Tensor | None = None, normalise: bool = True):
    """OPT-family LayerNorm over bf16 rows (fp32 gamma / beta), same fusions as rmsnorm plus the reduced linear's bias."""
    src = x if x is not None else partial
    rows, cols = src.shape[-2], src.shape[-1]
    if x is not None:
        _chk(x, BF16, "x")
    else:
        _chk(partial, F32, "partial")
    for t, n in ((gamma, "gamma"), (beta, "beta"), (red_bias, "red_bias")):
        if t is not None:
            _chk(t, F32, n)
    y = torch.empty((rows, cols), dtype=BF16, device=src.device) if normalise else None
    L.check(L.load().opus_layernorm_bf16(_p(x), _p(partial), 0 if partial is None else partial.shape[0], _p(red_bias),
                                         _p(residual), _p(h_out), _p(gamma), _p(beta), _p(y), rows, cols, eps,
                                         _stream()), "opus_layernorm_bf16")
    return y


def add_pos_embed_(h: torch.Tensor, table: torch.Tensor, pos: torch.Tensor, offset: int = 2) -> torch.Tensor:
    """h[i] += table[pos[i] + offset] in place (OPTLearnedPositionalEmbedding)."""
    _chk(h, BF16, "h"); _chk(table, BF16, "table"); _chk(pos, I32, "pos")
    L.check(L.load().opus_add_pos_embed_bf16(_p(h), _p(table), _p(pos), offset, table.shape[0], h.shape[0], h.shape[1],
                                             _stream()), "opus_add_pos_embed_bf16")
    return h


def rope_esm_(qkv: torch.Tensor, pos: torch.Tensor, cos_t: torch.Tensor, sin_t: torch.Tensor, n_heads: int,
              head_dim: int, q_scale: float) -> torch.Tensor:
    _chk(qkv, BF16, "qkv"); _chk(pos, I32, "pos"); _chk(cos_t, F32, "cos"); _chk(sin_t, F32, "sin")
    L.check(L.load().opus_rope_esm_bf16(_p(qkv), _p(pos), _p(cos_t), _p(sin_t), qkv.shape[0], n_heads, head_dim,
                                        qkv.stride(0), q_scale, _stream()), "opus_rope_esm_bf16")
    return qkv


def rope_llama_kvappend_(qkv: torch.Tensor, pos: torch.Tensor, slot: torch.Tensor | None, cos_t: torch.Tensor,
                         sin_t: torch.Tensor, kcache: torch.Tensor | None, vcache: torch.Tensor | None, n_q_heads: int,
                         n_kv_heads: int, head_dim: int, block_size: int = 16,
                         partial: torch.Tensor | None = None) -> torch.Tensor:
    _chk(qkv, BF16, "qkv"); _chk(pos, I32, "pos"); _chk(cos_t, BF16, "cos"); _chk(sin_t, BF16, "sin")
    L.check(L.load().opus_rope_llama_kvappend_bf16(
        _p(qkv), _p(partial), 0 if partial is None else partial.shape[0], _p(pos), _p(slot), _p(cos_t), _p(sin_t),
        _p(kcache), _p(vcache), qkv.shape[0], n_q_heads, n_kv_heads, head_dim, qkv.stride(0), block_size, _stream()),
        "opus_rope_llama_kvappend_bf16")
    return qkv


def final_ln_meanpool(x: torch.Tensor, cu_seqlens: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                      eps: float = 1e-5, want_hidden: bool = False, delta: torch.Tensor | None = None):
    _chk(x, F32, "x"); _chk(cu_seqlens, I32, "cu_seqlens")
    n_seqs, dim = cu_seqlens.numel() - 1, x.shape[1]
    pooled = torch.empty((n_seqs, dim), dtype=F32, device=x.device)
    pooled_l2 = torch.empty((n_seqs, dim), dtype=BF16, device=x.device)
    hidden = torch.empty_like(x) if want_hidden else None
    L.check(L.load().opus_final_ln_meanpool(_p(x), _p(delta), _p(cu_seqlens), _p(gamma), _p(beta), _p(pooled), _p(pooled_l2),
                                            _p(hidden), n_seqs, dim, eps, _stream()), "opus_final_ln_meanpool")
    return pooled, pooled_l2, hidden


def l2norm(x: torch.Tensor) -> torch.Tensor:
    _chk(x, F32, "x")
    y = torch.empty(x.shape, dtype=BF16, device=x.device)
    L.check(L.load().opus_l2norm_f32_bf16(_p(x), _p(y), x.shape[0], x.shape[1], _stream()), "opus_l2norm_f32_bf16")
    return y


def splice_gather(src: torch.Tensor, embed: torch.Tensor, soft: torch.Tensor, out: torch.Tensor | None = None):
    _chk(src, I32, "src"); _chk(embed, BF16, "embed"); _chk(soft, BF16, "soft")
    dim = embed.shape[1]
    if out is None:
        out = torch.empty((src.numel(), dim), dtype=BF16, device=src.device)
    L.check(L.load().opus_splice_gather_bf16(_p(src), _p(embed), _p(soft), _p(out), src.numel(), dim, _stream()),
            "opus_splice_gather_bf16")
    return out


def argmax_eos(logits: torch.Tensor, finished: torch.Tensor, eos_ids: torch.Tensor | None, pad_id: int,
               next_tok: torch.Tensor, out_ids: torch.Tensor, step: int, n_unfinished: torch.Tensor | None = None):
    _chk(logits, BF16, "logits"); _chk(finished, I32, "finished"); _chk(out_ids, I32, "out_ids")
    L.check(L.load().opus_argmax_eos(_p(logits), logits.stride(0), logits.shape[1], logits.shape[0], _p(finished),
                                     _p(eos_ids), 0 if eos_ids is None else eos_ids.numel(), pad_id, _p(next_tok),
                                     _p(out_ids), out_ids.stride(0), step, _p(n_unfinished), _stream()),
            "opus_argmax_eos")



def cross_entropy_rows(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """fp32 per-row cross entropy over bf16 logits [rows, vocab]; rows with target < 0 contribute 0."""
    _chk(logits, BF16, "logits"); _chk(target, I32, "target")
    loss = torch.empty((logits.shape[0],), dtype=F32, device=logits.device)
    L.check(L.load().opus_cross_entropy_bf16(_p(logits), logits.stride(0), logits.shape[1], _p(target), _p(loss),
                                             logits.shape[0], _stream()), "opus_cross_entropy_bf16")
    return loss


def sample_top_p(logits: torch.Tensor, temperature: float, top_p: float, seed: int, finished: torch.Tensor,
                 eos_ids: torch.Tensor | None, pad_id: int, next_tok: torch.Tensor, out_ids: torch.Tensor, step: int,
                 n_unfinished: torch.Tensor | None = None, kept_count: torch.Tensor | None = None):
    """Temperature / nucleus sampling of one token per row (HF do_sample=True semantics), EOS bookkeeping included."""
    _chk(logits, BF16, "logits"); _chk(finished, I32, "finished"); _chk(out_ids, I32, "out_ids")
    L.check(L.load().opus_sample_top_p(_p(logits), logits.stride(0), logits.shape[1], logits.shape[0],
                                       float(temperature), float(top_p), int(seed) & (2 ** 64 - 1), _p(finished),
                                       _p(eos_ids), 0 if eos_ids is None else eos_ids.numel(), pad_id, _p(next_tok),
                                       _p(out_ids), out_ids.stride(0), step, _p(n_unfinished), _p(kept_count),
                                       _stream()), "opus_sample_top_p")

def embed_gather(tok: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    _chk(tok, I32, "tok"); _chk(table, BF16, "table")
    x = torch.empty((tok.numel(), table.shape[1]), dtype=BF16, device=tok.device)
    L.check(L.load().opus_embed_gather_bf16(_p(tok), _p(table), _p(x), tok.numel(), table.shape[1], _stream()),
            "opus_embed_gather_bf16")
    return x


def lora_merge_(W: torch.Tensor, A: torch.Tensor, B: torch.Tensor, scale: float) -> torch.Tensor:
    _chk(W, BF16, "W"); _chk(A, BF16, "A"); _chk(B, BF16, "B")
    L.check(L.load().opus_lora_merge_bf16(_p(W), _p(A), _p(B), W.shape[0], W.shape[1], A.shape[0], scale, _stream()),
            "opus_lora_merge_bf16")
    return W


# ------------------------------------------------------------------------------------------------ attention
def attn_varlen(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cu_seqlens: torch.Tensor, max_len: int,
                n_q_heads: int, n_kv_heads: int, head_dim: int, causal: bool, scale: float) -> torch.Tensor:
    """q/k/v: 2-D bf16 views [n_tok, heads*head_dim] (may be column slices of a fused qkv buffer)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _chk(t, BF16, n, contiguous=False)
        if t.stride(1) != 1:
            raise L.OpusError(f"{n}: inner stride must be 1")
    o = torch.empty((q.shape[0], n_q_heads * head_dim), dtype=BF16, device=q.device)
    L.check(L.load().opus_attn_varlen_bf16(_p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(o),
                                           o.stride(0), _p(cu_seqlens), cu_seqlens.numel() - 1, q.shape[0], max_len, n_q_heads,
                                           n_kv_heads, head_dim, int(causal), scale, _stream()),
            "opus_attn_varlen_bf16")
    return o


def attn_decode_paged(q: torch.Tensor, kcache: torch.Tensor, vcache: torch.Tensor, block_table: torch.Tensor,
                      ctx_len: torch.Tensor, n_q_heads: int, n_kv_heads: int, head_dim: int, scale: float,
                      block_size: int = 16) -> torch.Tensor:
    _chk(q, BF16, "q", contiguous=False); _chk(kcache, BF16, "kcache"); _chk(vcache, BF16, "vcache")
    _chk(block_table, I32, "block_table"); _chk(ctx_len, I32, "ctx_len")
    o = torch.empty((q.shape[0], n_q_heads * head_dim), dtype=BF16, device=q.device)
    L.check(L.load().opus_attn_decode_paged_bf16(_p(q), q.stride(0), _p(kcache), _p(vcache), _p(block_table),
                                                 block_table.shape[1], _p(ctx_len), _p(o), o.stride(0), q.shape[0],
                                                 n_q_heads, n_kv_heads, head_dim, block_size, scale, _stream()),
            "opus_attn_decode_paged_bf16")
    return o
