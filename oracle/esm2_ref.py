"""ORACLE (test infrastructure only). Plain-PyTorch restatement of the ESM-2 encoder as the reference uses it.

The reference calls third-party **fair-esm ~= 2.0.0** (requirements.txt:6), which is NOT under /root/reference and not
installed here; call sites: cstp_v3/modelling.py:21 (esm2_t33_650M_UR50D), :44 (batch converter), :48 (forward with
repr_layers=[33]), :53-55 (mean pool). This file restates fair-esm's published algorithm (esm/model/esm2.py,
esm/modules.py, esm/multihead_attention.py, esm/rotary_embedding.py, esm/data.py of fair-esm 2.0.0) and is pinned in
this container against HuggingFace `EsmModel` (transformers 5.5.0, a port of the same model) by
oracle/make_golden.py -> tests/golden/esm2_small.pt, and checked again in tests/test_oracle_cpu.py.

Parity status: unpinned by the reference itself (it has no tests); pinned here against HF EsmModel.

Weights use the fair-esm state-dict names:
  embed_tokens.weight, layers.{i}.self_attn.{q,k,v,out}_proj.{weight,bias}, layers.{i}.self_attn_layer_norm.*,
  layers.{i}.fc1.*, layers.{i}.fc2.*, layers.{i}.final_layer_norm.*, emb_layer_norm_after.*
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .ops_ref import esm_rope_tables, gelu_erf, rope_half_ref

# fair-esm Alphabet.from_architecture("ESM-1b"): prepend <cls> <pad> <eos> <unk>, 25 residue symbols + '.', '-',
# then <null_1> <mask>; prepend_bos = append_eos = True (esm/data.py)
ESM_TOKS = ["<cls>", "<pad>", "<eos>", "<unk>", "L", "A", "G", "V", "S", "E", "R", "T", "I", "D", "P", "K", "Q", "N",
            "F", "Y", "M", "H", "W", "C", "X", "B", "U", "Z", "O", ".", "-", "<null_1>", "<mask>"]
TOK = {t: i for i, t in enumerate(ESM_TOKS)}
CLS, PAD, EOS, UNK, MASK = 0, 1, 2, 3, 32


def tokenize(seqs: list[str]) -> torch.Tensor:
    """BatchConverter semantics (esm/data.py): <cls> + residues + <eos>, right-padded with <pad> to the longest."""
    T = max(len(s) for s in seqs) + 2
    out = torch.full((len(seqs), T), PAD, dtype=torch.int64)
    for i, s in enumerate(seqs):
        ids = [CLS] + [TOK.get(ch, UNK) for ch in s] + [EOS]
        out[i, : len(ids)] = torch.tensor(ids)
    return out


def esm2_forward(w: dict, tokens: torch.Tensor, n_layers: int, n_heads: int, matmul_dtype=None) -> torch.Tensor:
    """Returns representations[n_layers] = emb_layer_norm_after(x): fp32 [B, T, D].

    matmul_dtype = torch.bfloat16/float16 reproduces torch.autocast as the reference runs the encoder
    (opus_arch.py:107): Linear/bmm operands are cast, LayerNorm / softmax / residual stream stay fp32.
    """
    def lin(x, name):
        W, b = w[name + ".weight"], w[name + ".bias"]
        if matmul_dtype is None:
            return F.linear(x, W, b)
        return F.linear(x.to(matmul_dtype), W.to(matmul_dtype), b.to(matmul_dtype))

    B, T = tokens.shape
    D = w["embed_tokens.weight"].shape[1]
    hd = D // n_heads
    padding_mask = tokens.eq(PAD)
    x = w["embed_tokens.weight"][tokens]                               # embed_scale = 1
    # token_dropout (ESM2.forward): zero <mask> embeddings and rescale by (1-0.15*0.8)/(1-observed mask ratio)
    x = x.masked_fill((tokens == MASK).unsqueeze(-1), 0.0)
    src_lengths = (~padding_mask).sum(-1)
    mask_ratio_observed = (tokens == MASK).sum(-1).to(x.dtype) / src_lengths
    x = x * (1 - 0.15 * 0.8) / (1 - mask_ratio_observed)[:, None, None]
    x = x * (1 - padding_mask.unsqueeze(-1).type_as(x))
    has_pad = bool(padding_mask.any())
    cos, sin = esm_rope_tables(T, hd, x.device)
    for i in range(n_layers):
        p = f"layers.{i}."
        res = x
        h = F.layer_norm(x, (D,), w[p + "self_attn_layer_norm.weight"], w[p + "self_attn_layer_norm.bias"], 1e-5)
        q = lin(h, p + "self_attn.q_proj") * hd ** -0.5                # scale BEFORE rotary
        k = lin(h, p + "self_attn.k_proj")
        v = lin(h, p + "self_attn.v_proj")
        q = q.view(B, T, n_heads, hd).transpose(1, 2)
        k = k.view(B, T, n_heads, hd).transpose(1, 2)
        v = v.view(B, T, n_heads, hd).transpose(1, 2)
        q = rope_half_ref(q.float(), cos, sin)                         # fp32 cos/sin promote q,k to fp32
        k = rope_half_ref(k.float(), cos, sin)
        if matmul_dtype is not None:
            q, k = q.to(matmul_dtype), k.to(matmul_dtype)
        s = torch.matmul(q, k.transpose(-1, -2))
        if has_pad:
            s = s.masked_fill(padding_mask[:, None, None, :], float("-inf"))
        pr = torch.softmax(s.float(), dim=-1).type_as(s)
        a = torch.matmul(pr, v).transpose(1, 2).reshape(B, T, D)
        x = res + lin(a, p + "self_attn.out_proj")
        res = x
        h = F.layer_norm(x, (D,), w[p + "final_layer_norm.weight"], w[p + "final_layer_norm.bias"], 1e-5)
        h = gelu_erf(lin(h, p + "fc1"))
        x = res + lin(h, p + "fc2")
        x = x.float()
    return F.layer_norm(x, (D,), w["emb_layer_norm_after.weight"], w["emb_layer_norm_after.bias"], 1e-5)


def get_protein_seq_embeddings(w: dict, seqs: list[str], n_layers: int, n_heads: int, matmul_dtype=None):
    """cstp_v3/modelling.py:37-57: tokenise, forward, mean over residues [1, len-1) per sequence -> fp32 [B, D]."""
    tokens = tokenize(seqs).to(w["embed_tokens.weight"].device)
    lens = (tokens != PAD).sum(1)
    rep = esm2_forward(w, tokens, n_layers, n_heads, matmul_dtype)
    return torch.stack([rep[i, 1: int(lens[i]) - 1].mean(0) for i in range(len(seqs))]).float()


def random_esm2_weights(n_layers: int, dim: int, ffn: int, vocab: int = 33, seed: int = 0, device="cpu") -> dict:
    """Seeded synthetic weights in fair-esm naming. Linear ~ N(0, 0.02) as HF/fair-esm init; LayerNorm gains ~1 with
    a little noise and small random biases so that affine terms and biases are actually exercised."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std=0.02: torch.randn(*s, generator=g) * std  # noqa: E731
    w = {"embed_tokens.weight": rn(vocab, dim, std=1.0)}
    w["embed_tokens.weight"][PAD] = 0
    for i in range(n_layers):
        p = f"layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            w[p + f"self_attn.{n}.weight"] = rn(dim, dim, std=dim ** -0.5)
            w[p + f"self_attn.{n}.bias"] = rn(dim, std=0.02)
        w[p + "fc1.weight"] = rn(ffn, dim, std=dim ** -0.5)
        w[p + "fc1.bias"] = rn(ffn, std=0.02)
        w[p + "fc2.weight"] = rn(dim, ffn, std=ffn ** -0.5 * 0.5)
        w[p + "fc2.bias"] = rn(dim, std=0.02)
        for n in ("self_attn_layer_norm", "final_layer_norm"):
            w[p + n + ".weight"] = 1.0 + rn(dim, std=0.05)
            w[p + n + ".bias"] = rn(dim, std=0.05)
    w["emb_layer_norm_after.weight"] = 1.0 + rn(dim, std=0.05)
    w["emb_layer_norm_after.bias"] = rn(dim, std=0.05)
    return {k: v.to(device) for k, v in w.items()}


def to_hf_esm_state_dict(w: dict, n_layers: int) -> dict:
    """fair-esm names -> HF EsmModel names (for the cross-check in make_golden.py / tests)."""
    m = {"embeddings.word_embeddings.weight": w["embed_tokens.weight"],
         "encoder.emb_layer_norm_after.weight": w["emb_layer_norm_after.weight"],
         "encoder.emb_layer_norm_after.bias": w["emb_layer_norm_after.bias"]}
    for i in range(n_layers):
        s, d = f"layers.{i}.", f"encoder.layer.{i}."
        for a, b in (("self_attn.q_proj", "attention.self.query"), ("self_attn.k_proj", "attention.self.key"),
                     ("self_attn.v_proj", "attention.self.value"), ("self_attn.out_proj", "attention.output.dense"),
                     ("self_attn_layer_norm", "attention.LayerNorm"), ("fc1", "intermediate.dense"),
                     ("fc2", "output.dense"), ("final_layer_norm", "LayerNorm")):
            m[d + b + ".weight"] = w[s + a + ".weight"]
            m[d + b + ".bias"] = w[s + a + ".bias"]
    return m
