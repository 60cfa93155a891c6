"""ORACLE (test infrastructure only). Plain-PyTorch restatement of the LLM half of the reference's generate path.

The reference's `OpusLlamaForCausalLM.generate` (multi_modality_v1/model/language_model/opus_llama.py:95-132) hands
`inputs_embeds` + `attention_mask` to third-party **transformers ~= 4.46.3** (requirements.txt) `LlamaForCausalLM` /
`GenerationMixin`; that code is not under /root/reference. This file restates its published algorithm
(HF: models/llama/modeling_llama.py — RMSNorm :53-67, rotary :124-168, MLP :182-183, eager attention :199-222,
decoder layer :292-333; generation/utils.py greedy branch of `_sample`) in plain torch, written op-for-op in the same
order so that running it on bf16 tensors reproduces HF's bf16 rounding points, and running it on fp32 tensors gives
the "truth".

Pinned in this container by oracle/make_golden.py against (a) transformers 5.5.0 `LlamaForCausalLM` and (b) the
reference's own `OpusLlamaForCausalLM.generate` imported from /root/reference (with the `cache_position` shim that
transformers >= 4.47 needs), fixtures in tests/golden/llama_small.pt. Parity status: pinned against the reference run.

Weights use the HF state-dict names (model.embed_tokens.weight, model.layers.{i}.self_attn.q_proj.weight, ...,
model.norm.weight, lm_head.weight).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class LlamaCfg:
    n_layers: int = 32
    dim: int = 4096
    n_q_heads: int = 32
    n_kv_heads: int = 8
    head_dim: int = 128
    ffn_dim: int = 14336
    vocab: int = 128256
    rms_eps: float = 1e-5
    rope_theta: float = 500000.0

    @property
    def q_dim(self):
        return self.n_q_heads * self.head_dim

    @property
    def kv_dim(self):
        return self.n_kv_heads * self.head_dim


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    dt = x.dtype
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(-1, keepdim=True)
    xf = xf * torch.rsqrt(var + eps)
    return w * xf.to(dt)


def rope_cos_sin(position_ids: torch.Tensor, head_dim: int, theta: float, dtype) -> tuple[torch.Tensor, torch.Tensor]:
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float().to(position_ids.device)
                                / head_dim))
    freqs = position_ids[..., None].float() * inv_freq  # [B, T, hd/2]
    emb = torch.cat([freqs, freqs], -1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def llama_forward(w: dict, cfg: LlamaCfg, embeds: torch.Tensor, attention_mask: torch.Tensor,
                  position_ids: torch.Tensor, past=None, all_positions: bool = False):
    """embeds [B, T, D]; attention_mask bool [B, S] over (past + current) keys, S = past_len + T; position_ids [B, T].
    Returns (last-position logits fp32 [B, V], new past). `past` = list of (k, v) each [B, Hkv, S_past, hd]."""
    B, T, D = embeds.shape
    dt = embeds.dtype
    Hq, Hkv, hd = cfg.n_q_heads, cfg.n_kv_heads, cfg.head_dim
    cos, sin = rope_cos_sin(position_ids, hd, cfg.rope_theta, dt)
    cos, sin = cos[:, None], sin[:, None]
    S = attention_mask.shape[1]
    past_len = S - T
    # HF 4-D mask: causal over absolute key index, AND key-padding
    q_idx = torch.arange(past_len, S, device=embeds.device)[:, None]
    k_idx = torch.arange(S, device=embeds.device)[None, :]
    allowed = (k_idx <= q_idx)[None, None] & attention_mask.bool()[:, None, None, :]
    bias = torch.zeros(B, 1, T, S, dtype=dt, device=embeds.device).masked_fill(~allowed, torch.finfo(dt).min)
    h = embeds
    new_past = []
    for i in range(cfg.n_layers):
        p = f"model.layers.{i}."
        res = h
        x = rmsnorm(h, w[p + "input_layernorm.weight"], cfg.rms_eps)
        # q/k/v projection biases exist in the Qwen2 family (HF modeling_qwen2.py: Linear(..., bias=True)), not in Llama
        q = F.linear(x, w[p + "self_attn.q_proj.weight"], w.get(p + "self_attn.q_proj.bias")).view(B, T, Hq, hd).transpose(1, 2)
        k = F.linear(x, w[p + "self_attn.k_proj.weight"], w.get(p + "self_attn.k_proj.bias")).view(B, T, Hkv, hd).transpose(1, 2)
        v = F.linear(x, w[p + "self_attn.v_proj.weight"], w.get(p + "self_attn.v_proj.bias")).view(B, T, Hkv, hd).transpose(1, 2)
        q = (q * cos) + (rotate_half(q) * sin)
        k = (k * cos) + (rotate_half(k) * sin)
        if past is not None:
            k = torch.cat([past[i][0], k], dim=2)
            v = torch.cat([past[i][1], v], dim=2)
        new_past.append((k, v))
        kr = k.repeat_interleave(Hq // Hkv, dim=1)
        vr = v.repeat_interleave(Hq // Hkv, dim=1)
        s = torch.matmul(q, kr.transpose(2, 3)) * (1.0 / math.sqrt(hd)) + bias
        pr = torch.softmax(s, dim=-1, dtype=torch.float32).to(dt)
        a = torch.matmul(pr, vr).transpose(1, 2).reshape(B, T, Hq * hd)
        h = res + F.linear(a, w[p + "self_attn.o_proj.weight"])
        res = h
        x = rmsnorm(h, w[p + "post_attention_layernorm.weight"], cfg.rms_eps)
        g = F.linear(x, w[p + "mlp.gate_proj.weight"])
        u = F.linear(x, w[p + "mlp.up_proj.weight"])
        h = res + F.linear(F.silu(g) * u, w[p + "mlp.down_proj.weight"])
    lm_head = w["lm_head.weight"] if "lm_head.weight" in w else w["model.embed_tokens.weight"]   # tied embeddings
    if all_positions:   # teacher-forced scoring: HF returns logits for every position when labels are given
        return F.linear(rmsnorm(h, w["model.norm.weight"], cfg.rms_eps), lm_head).float(), new_past
    h = rmsnorm(h[:, -1:, :], w["model.norm.weight"], cfg.rms_eps)
    logits = F.linear(h, lm_head)[:, -1, :]
    return logits.float(), new_past


def causal_lm_loss(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
    """HF LlamaForCausalLM loss (reached from opus_llama.py:82-93 when labels are passed): position t predicts label
    t+1; mean fp32 cross entropy over the labels != ignore_index. logits [B, T, V], labels [B, T]."""
    shift_logits = logits[..., :-1, :].float().reshape(-1, logits.shape[-1])
    shift_labels = labels[..., 1:].reshape(-1)
    return F.cross_entropy(shift_logits, shift_labels, ignore_index=ignore_index)


def greedy_generate(w: dict, cfg: LlamaCfg, embeds: torch.Tensor, attention_mask: torch.Tensor, max_new_tokens: int,
                    eos_ids=(), pad_id: int = 0, return_logits: bool = False):
    """HF GenerationMixin greedy loop started from inputs_embeds (left-padded): position ids = cumsum(mask)-1 with pads
    filled by 1 (generation/utils.py), new tokens only are returned, finished rows emit pad, stop when all finished."""
    B = embeds.shape[0]
    mask = attention_mask.bool()
    pos = (mask.long().cumsum(-1) - 1).masked_fill(~mask, 1)
    logits, past = llama_forward(w, cfg, embeds, mask, pos)
    unfinished = torch.ones(B, dtype=torch.bool, device=embeds.device)
    out, all_logits = [], []
    for step in range(max_new_tokens):
        if return_logits:
            all_logits.append(logits)
        tok = torch.argmax(logits, dim=-1)
        tok = torch.where(unfinished, tok, torch.full_like(tok, pad_id))
        out.append(tok)
        for e in eos_ids:
            unfinished = unfinished & (tok != e)
        if not bool(unfinished.any()) or step == max_new_tokens - 1:
            break
        mask = torch.cat([mask, torch.ones(B, 1, dtype=torch.bool, device=mask.device)], dim=1)
        pos = pos[:, -1:] + 1
        x = w["model.embed_tokens.weight"][tok][:, None, :]
        logits, past = llama_forward(w, cfg, x, mask, pos, past)
    ids = torch.stack(out, dim=1)
    return (ids, torch.stack(all_logits, 1)) if return_logits else ids


def random_llama_weights(cfg: LlamaCfg, seed: int = 0, dtype=torch.float32, device="cpu", peaked: bool = False) -> dict:
    """Seeded synthetic weights (HF names). Default: N(0, 0.02) like HF `_init_weights`, RMSNorm gains ~1.

    peaked=True is the token-parity recipe (SURVEY.md §7 hard part 1): residual branches are damped and lm_head is
    tied to a fixed random permutation of the embedding rows with a large gain, so the greedy argmax has a margin far
    above bf16 noise and both implementations must agree token-for-token unless one of them is wrong.
    """
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std=0.02: (torch.randn(*s, generator=g) * std)  # noqa: E731
    D, F_ = cfg.dim, cfg.ffn_dim
    w = {"model.embed_tokens.weight": rn(cfg.vocab, D, std=1.0 if peaked else 0.02)}
    out_std = 0.02 / math.sqrt(2 * cfg.n_layers) if peaked else 0.02
    for i in range(cfg.n_layers):
        p = f"model.layers.{i}."
        w[p + "input_layernorm.weight"] = 1.0 + rn(D, std=0.05)
        w[p + "post_attention_layernorm.weight"] = 1.0 + rn(D, std=0.05)
        w[p + "self_attn.q_proj.weight"] = rn(cfg.q_dim, D)
        w[p + "self_attn.k_proj.weight"] = rn(cfg.kv_dim, D)
        w[p + "self_attn.v_proj.weight"] = rn(cfg.kv_dim, D)
        w[p + "self_attn.o_proj.weight"] = rn(D, cfg.q_dim, std=out_std)
        w[p + "mlp.gate_proj.weight"] = rn(F_, D)
        w[p + "mlp.up_proj.weight"] = rn(F_, D)
        w[p + "mlp.down_proj.weight"] = rn(D, F_, std=out_std)
    w["model.norm.weight"] = 1.0 + rn(D, std=0.05)
    if peaked:
        perm = torch.randperm(cfg.vocab, generator=g)
        w["lm_head.weight"] = w["model.embed_tokens.weight"][perm] * (8.0 / math.sqrt(D))
    else:
        w["lm_head.weight"] = rn(cfg.vocab, D)
    return {k: v.to(dtype).to(device) for k, v in w.items()}


def lora_merge_ref(W: torch.Tensor, A: torch.Tensor, B: torch.Tensor, alpha: float, r: int) -> torch.Tensor:
    """peft 0.11.1 LoraLayer.merge (Linear): W += (lora_alpha / r) * (B @ A); call site model/builder.py:107-109."""
    return W + (alpha / r) * (B @ A)
