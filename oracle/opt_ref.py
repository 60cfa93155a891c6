"""ORACLE (test infrastructure only). Plain-PyTorch restatement of the OPT / Galactica half of the reference's generate
path (sibling family).

The reference's `OpusOPTForCausalLM.generate` (multi_modality_v1/model/language_model/opus_opt.py:95-132; family chosen
at model/builder.py:71-81 for 'opt' / 'galactica' base paths) hands `inputs_embeds` + `attention_mask` to third-party
**transformers ~= 4.46.3** `OPTForCausalLM`; that code is not under /root/reference. This file restates its published
algorithm (HF models/opt/modeling_opt.py: learned positions with offset 2 from `cumsum(mask)*mask - 1`, pre-LN decoder
layer -- 350m's post-LN order is kept behind `pre_ln=False` --, q scaled by head_dim^-0.5 right after its projection,
fp32 softmax, ReLU (OPT) or erf-GELU (Galactica) MLP, final LayerNorm, lm_head without bias) op-for-op in the same order,
so bf16 tensors reproduce HF's bf16 rounding points and fp32 tensors give the "truth".

Pinned in this container by oracle/make_golden.py against the reference's own `OpusOPTForCausalLM.generate` imported
from /root/reference and run over transformers 5.5.0: fixture tests/golden/opt_small.pt. Parity status: pinned against
the reference run.

Weights use the HF state-dict names (model.decoder.embed_tokens.weight, model.decoder.embed_positions.weight,
model.decoder.layers.{i}.self_attn.q_proj.{weight,bias}, ..., model.decoder.final_layer_norm.*, lm_head.weight).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class OptCfg:
    n_layers: int = 32
    dim: int = 4096
    n_heads: int = 32
    ffn_dim: int = 16384
    vocab: int = 50272
    max_pos: int = 2048
    activation: str = "relu"      # "relu" (OPT) | "gelu" (Galactica)
    pre_ln: bool = True           # config.do_layer_norm_before
    ln_eps: float = 1e-5

    @property
    def head_dim(self):
        return self.dim // self.n_heads


P0 = "model.decoder."


def _ln(x, w, key, eps):
    return F.layer_norm(x, (x.shape[-1],), w[key + ".weight"], w.get(key + ".bias"), eps)


def positions_from_mask(mask: torch.Tensor) -> torch.Tensor:
    """modeling_opt.py: `cumsum(attention_mask) * attention_mask - 1` (pads get -1 -> table row 1 after the offset)."""
    m = mask.long()
    return m.cumsum(-1) * m - 1


def opt_forward(w: dict, cfg: OptCfg, embeds: torch.Tensor, attention_mask: torch.Tensor, position_ids: torch.Tensor,
                past=None, all_positions: bool = False):
    """embeds [B, T, D]; attention_mask bool [B, S] over past + current keys; position_ids [B, T] (before the +2 offset).
    Returns (last-position logits fp32 [B, V], new past); past = list of (k, v) each [B, H, S_past, hd]."""
    B, T, D = embeds.shape
    dt = embeds.dtype
    H, hd = cfg.n_heads, cfg.head_dim
    S = attention_mask.shape[1]
    past_len = S - T
    q_idx = torch.arange(past_len, S, device=embeds.device)[:, None]
    k_idx = torch.arange(S, device=embeds.device)[None, :]
    allowed = (k_idx <= q_idx)[None, None] & attention_mask.bool()[:, None, None, :]
    bias = torch.zeros(B, 1, T, S, dtype=dt, device=embeds.device).masked_fill(~allowed, torch.finfo(dt).min)
    act = F.relu if cfg.activation == "relu" else F.gelu
    h = embeds + w[P0 + "embed_positions.weight"][position_ids + 2]
    new_past = []
    for i in range(cfg.n_layers):
        p = f"{P0}layers.{i}."
        res = h
        x = _ln(h, w, p + "self_attn_layer_norm", cfg.ln_eps) if cfg.pre_ln else h
        q = F.linear(x, w[p + "self_attn.q_proj.weight"], w.get(p + "self_attn.q_proj.bias")) * (hd ** -0.5)
        k = F.linear(x, w[p + "self_attn.k_proj.weight"], w.get(p + "self_attn.k_proj.bias"))
        v = F.linear(x, w[p + "self_attn.v_proj.weight"], w.get(p + "self_attn.v_proj.bias"))
        q, k, v = (t.view(B, T, H, hd).transpose(1, 2) for t in (q, k, v))
        if past is not None:
            k = torch.cat([past[i][0], k], dim=2)
            v = torch.cat([past[i][1], v], dim=2)
        new_past.append((k, v))
        s = torch.matmul(q, k.transpose(2, 3)) * 1.0 + bias
        pr = torch.softmax(s, dim=-1, dtype=torch.float32).to(dt)
        a = torch.matmul(pr, v).transpose(1, 2).reshape(B, T, D)
        h = res + F.linear(a, w[p + "self_attn.out_proj.weight"], w.get(p + "self_attn.out_proj.bias"))
        if not cfg.pre_ln:
            h = _ln(h, w, p + "self_attn_layer_norm", cfg.ln_eps)
        res = h
        x = _ln(h, w, p + "final_layer_norm", cfg.ln_eps) if cfg.pre_ln else h
        x = act(F.linear(x, w[p + "fc1.weight"], w.get(p + "fc1.bias")))
        h = res + F.linear(x, w[p + "fc2.weight"], w.get(p + "fc2.bias"))
        if not cfg.pre_ln:
            h = _ln(h, w, p + "final_layer_norm", cfg.ln_eps)
    lm_head = w["lm_head.weight"] if "lm_head.weight" in w else w[P0 + "embed_tokens.weight"]   # tied in OPT checkpoints
    if not all_positions:
        h = h[:, -1:, :]
    if cfg.pre_ln:
        h = _ln(h, w, P0 + "final_layer_norm", cfg.ln_eps)
    logits = F.linear(h, lm_head)
    return (logits.float() if all_positions else logits[:, -1, :].float()), new_past


def greedy_generate(w: dict, cfg: OptCfg, embeds: torch.Tensor, attention_mask: torch.Tensor, max_new_tokens: int,
                    eos_ids=(), pad_id: int = 0, return_logits: bool = False):
    """HF greedy loop from inputs_embeds (left-padded), see llama_ref.greedy_generate; positions follow the mask."""
    B = embeds.shape[0]
    mask = attention_mask.bool()
    pos = positions_from_mask(mask)
    logits, past = opt_forward(w, cfg, embeds, mask, pos)
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
        pos = positions_from_mask(mask)[:, -1:]
        x = w[P0 + "embed_tokens.weight"][tok][:, None, :]
        logits, past = opt_forward(w, cfg, x, mask, pos, past)
    ids = torch.stack(out, dim=1)
    return (ids, torch.stack(all_logits, 1)) if return_logits else ids
