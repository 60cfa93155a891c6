"""ORACLE (test infrastructure only — never imported by the product package).

Plain-PyTorch fp32 restatements of the individual operators on the OPUS-PLLM generation path, with the rounding
points of the reference stack (torch autocast for the encoder/projectors, bf16 HF Llama) made explicit. Each function
cites the reference (or third-party) lines it follows; `HF:` = transformers/models in this image, used only as a
cross-check of arithmetic the reference does not vendor.

Parity status: the reference has no tests/golden vectors for this path. These restatements are pinned
instead against the reference's own classes / HF modules executed in the build container, see
oracle/make_golden.py and tests/golden/.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def bf16r(x: torch.Tensor) -> torch.Tensor:
    """round-trip through bf16 (the storage dtype of activations), result fp32"""
    return x.to(torch.bfloat16).to(torch.float32)


def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    # fair-esm `gelu`; HF:esm/modeling_esm.py:57-61 — exact erf form, NOT the tanh approximation
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def linear_ref(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None) -> torch.Tensor:
    """fp32 accumulate of bf16 operands: what cuBLAS bf16 GEMM computes before its output rounding."""
    y = x.to(torch.float32) @ w.to(torch.float32).t()
    if bias is not None:
        y = y + bias.to(torch.float32)
    return y


def swiglu_interleaved_ref(acc: torch.Tensor) -> torch.Tensor:
    """acc fp32 [rows, 2*ffn] with columns (gate_j, up_j) interleaved. HF:llama/modeling_llama.py:182-183
    `down_proj(act_fn(gate_proj(x)) * up_proj(x))` in bf16: each linear output, silu and the product round to bf16."""
    g = bf16r(acc[:, 0::2])
    u = bf16r(acc[:, 1::2])
    return bf16r(bf16r(F.silu(g)) * u)


def layernorm_ref(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    return F.layer_norm(x.float(), (x.shape[-1],), gamma.float(), beta.float(), eps)


def rmsnorm_ref(x: torch.Tensor, w: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """HF:llama/modeling_llama.py:62-67 LlamaRMSNorm for bf16 input; returns fp32 holding bf16-representable values."""
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(-1, keepdim=True)
    xn = bf16r(xf * torch.rsqrt(var + eps))
    return bf16r(w.to(torch.float32) * xn)


def esm_rope_tables(max_pos: int, head_dim: int = 64, device="cpu"):
    """fair-esm RotaryEmbedding / HF:esm/modeling_esm.py:81-123: inv_freq = 1/10000^(2i/d); returns cos,sin [max_pos, d/2]"""
    inv_freq = 1.0 / (10000 ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    t = torch.arange(max_pos, dtype=torch.float32)
    freqs = torch.outer(t, inv_freq)
    return freqs.cos().to(device), freqs.sin().to(device)


def rope_half_ref(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x [..., T, d]; cos/sin [T, d/2]; rotate_half convention cat(-x2, x1) (HF:esm/modeling_esm.py:43-55)."""
    d = x.shape[-1]
    x1, x2 = x[..., : d // 2], x[..., d // 2:]
    c = torch.cat([cos, cos], -1)
    s = torch.cat([sin, sin], -1)
    return x * c + torch.cat([-x2, x1], -1) * s


def llama_rope_tables(max_pos: int, head_dim: int = 128, theta: float = 500000.0, device="cpu"):
    """HF:llama/modeling_llama.py:124-168 default rope: cos/sin of cat(freqs, freqs), computed in fp32 then cast to the
    activation dtype (bf16). Returns bf16 [max_pos, head_dim]."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    t = torch.arange(max_pos, dtype=torch.float32)
    freqs = torch.outer(t, inv_freq)
    emb = torch.cat([freqs, freqs], -1)
    return emb.cos().to(torch.bfloat16).to(device), emb.sin().to(torch.bfloat16).to(device)


def llama_rope_bf16_ref(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x bf16 [T, H, d]; cos/sin bf16 [T, d]. bf16 arithmetic exactly as torch evaluates
    (q * cos) + (rotate_half(q) * sin) on bf16 tensors."""
    d = x.shape[-1]
    x1, x2 = x[..., : d // 2], x[..., d // 2:]
    rot = torch.cat([-x2, x1], -1)
    return (x * cos[:, None, :]) + (rot * sin[:, None, :])


def attention_ref(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool, scale: float) -> torch.Tensor:
    """Single sequence. q [T, Hq, d], k/v [T, Hkv, d] (any float dtype) -> fp32 [T, Hq, d]; softmax in fp32.
    fair-esm MultiheadAttention (bmm / softmax(float32) / bmm) and HF:llama/modeling_llama.py:199-222 eager attention."""
    T, Hq, d = q.shape
    Hkv = k.shape[1]
    qf, kf, vf = q.float(), k.float(), v.float()
    if Hkv != Hq:
        rep = Hq // Hkv
        kf = kf.repeat_interleave(rep, dim=1)
        vf = vf.repeat_interleave(rep, dim=1)
    s = torch.einsum("thd,shd->hts", qf, kf) * scale
    if causal:
        mask = torch.ones(T, k.shape[0], dtype=torch.bool, device=q.device).tril(diagonal=k.shape[0] - T)
        s = s.masked_fill(~mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    return torch.einsum("hts,shd->thd", p, vf)


def meanpool_ref(hidden: torch.Tensor, cu_seqlens) -> torch.Tensor:
    """cstp_v3/modelling.py:53-55: mean over residues [1, len-1) of each sequence (drops <cls>/<eos>)."""
    out = []
    for b in range(len(cu_seqlens) - 1):
        s, e = int(cu_seqlens[b]), int(cu_seqlens[b + 1])
        out.append(hidden[s + 1: e - 1].mean(0))
    return torch.stack(out).float()


def greedy_select_ref(logits: torch.Tensor, finished: torch.Tensor, eos_ids, pad_id: int):
    """HF:generation/utils.py greedy branch of _sample: argmax over float32 logits; finished rows emit pad; rows that
    emit EOS become finished."""
    tok = torch.argmax(logits.float(), dim=-1)
    tok = torch.where(finished.bool(), torch.full_like(tok, pad_id), tok)
    is_eos = torch.zeros_like(tok, dtype=torch.bool)
    for e in eos_ids:
        is_eos |= tok == e
    return tok, finished.bool() | is_eos


def top_p_keep_mask(logits: torch.Tensor, temperature: float, top_p: float) -> torch.Tensor:
    """Nucleus of every row, as HF builds it (transformers generation/logits_process.py: TemperatureLogitsWarper then
    TopPLogitsWarper with min_tokens_to_keep=1; reached from the reference through generate(do_sample=True,
    temperature, top_p), run_opus_ddp.py:126-128). logits [B, V] -> bool [B, V], True = kept."""
    scores = logits.float() / temperature
    sorted_logits, sorted_indices = torch.sort(scores, descending=False)
    cumulative_probs = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
    remove = cumulative_probs <= (1 - top_p)
    remove[..., -1:] = False
    remove = remove.scatter(1, sorted_indices, remove)
    return ~remove


def top_p_probs(logits: torch.Tensor, temperature: float, top_p: float) -> torch.Tensor:
    """Distribution HF samples the next token from: softmax of the warped scores (-inf outside the nucleus)."""
    scores = (logits.float() / temperature).masked_fill(~top_p_keep_mask(logits, temperature, top_p), float("-inf"))
    return scores.softmax(-1)
