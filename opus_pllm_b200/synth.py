"""Deterministic synthetic weights and inputs (no checkpoints or datasets exist offline).

Values come from a 64-bit integer hash evaluated with torch int64 ops, so they are bit-identical on every machine, torch
version and device (CPU or GPU) — the same tensors feed the oracle, the golden-fixture generator, the CUDA path and bench.py.
Shapes follow SURVEY.md §8(d): proteins are i.i.d. uniform over the 20 standard amino acids, prompts are uniform token
ids with one `-200` protein sentinel.
"""
from __future__ import annotations

import math
import random

import torch

def _i64(v: int) -> int:
    """reinterpret an unsigned 64-bit constant as the signed value torch.int64 holds"""
    v &= 0xFFFFFFFFFFFFFFFF
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(x: torch.Tensor, k: int) -> torch.Tensor:
    # logical shift right on int64 (torch's >> is arithmetic)
    return (x >> k) & ((1 << (64 - k)) - 1)


def _mix(x: torch.Tensor) -> torch.Tensor:
    # splitmix64 finaliser; int64 multiplication wraps modulo 2^64 on CPU and CUDA alike
    x = (x ^ _lsr(x, 30)) * _i64(0xBF58476D1CE4E5B9)
    x = (x ^ _lsr(x, 27)) * _i64(0x94D049BB133111EB)
    return x ^ _lsr(x, 31)


def _seed_of(name: str, seed: int) -> int:
    h = 1469598103934665603
    for ch in name.encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return (h ^ (seed * 0x9E3779B97F4A7C15)) & 0xFFFFFFFFFFFFFFFF


def hash_uniform(shape, name: str, seed: int = 0, device="cpu", chunk: int = 1 << 26) -> torch.Tensor:
    """fp32 tensor of `shape`, i.i.d. uniform in [-1, 1), fully determined by (name, seed, flat index): pure integer
    arithmetic, so the values are bit-identical whether generated on the CPU or on the GPU."""
    n = 1
    for d in shape:
        n *= int(d)
    out = torch.empty(n, dtype=torch.float32, device=device)
    s = _i64(_seed_of(name, seed))
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        idx = torch.arange(lo, hi, dtype=torch.int64, device=device)
        bits = _mix(idx * _i64(0x9E3779B97F4A7C15) + s)
        out[lo:hi] = _lsr(bits, 40).to(torch.float32) * (2.0 ** -23) - 1.0
    return out.reshape(tuple(shape))


def weight(shape, name: str, std: float, seed: int = 0, mean: float = 0.0, device="cpu") -> torch.Tensor:
    """uniform with the requested standard deviation (uniform[-1,1) has std 1/sqrt(3))"""
    return hash_uniform(shape, name, seed, device) * (std * math.sqrt(3.0)) + mean


AMINO = "ACDEFGHIKLMNPQRSTVWY"


def proteins(n: int, min_len: int, max_len: int | None = None, seed: int = 1234) -> list[str]:
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        ln = min_len if max_len is None else rng.randint(min_len, max_len)
        out.append("".join(rng.choice(AMINO) for _ in range(ln)))
    return out


def prompt_ids(n: int, text_len: int, vocab: int = 128256, seed: int = 1234, bos: int = 128000,
               sentinel_at: int = 40, ragged: int = 0) -> list[torch.Tensor]:
    """n prompts of `text_len` ids each (BOS first, one -200 sentinel); ragged > 0 shortens prompt i by (i % ragged)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    hi = min(120000, vocab)
    lo = min(1000, hi // 2)
    for i in range(n):
        ln = text_len - (i % ragged if ragged else 0)
        ids = torch.randint(lo, hi, (ln,), generator=g, dtype=torch.int64)
        ids[0] = bos if bos < vocab else 1
        ids[min(sentinel_at, ln - 1)] = -200
        out.append(ids)
    return out


# ---------------------------------------------------------------------------------------------- model weights
def esm2_weights(n_layers: int, dim: int, ffn: int, vocab: int = 33, seed: int = 0, device="cpu") -> dict:
    """fair-esm state-dict names; Linear std ~ 1/sqrt(fan_in) keeps activations O(1) through 33 random layers."""
    w = {"embed_tokens.weight": weight((vocab, dim), "esm.embed", 1.0, seed, device=device)}
    w["embed_tokens.weight"][1] = 0  # <pad>
    for i in range(n_layers):
        p = f"layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            w[p + f"self_attn.{n}.weight"] = weight((dim, dim), f"esm.{i}.{n}.w", dim ** -0.5, seed, device=device)
            w[p + f"self_attn.{n}.bias"] = weight((dim,), f"esm.{i}.{n}.b", 0.02, seed, device=device)
        w[p + "fc1.weight"] = weight((ffn, dim), f"esm.{i}.fc1.w", dim ** -0.5, seed, device=device)
        w[p + "fc1.bias"] = weight((ffn,), f"esm.{i}.fc1.b", 0.02, seed, device=device)
        w[p + "fc2.weight"] = weight((dim, ffn), f"esm.{i}.fc2.w", 0.5 * ffn ** -0.5, seed, device=device)
        w[p + "fc2.bias"] = weight((dim,), f"esm.{i}.fc2.b", 0.02, seed, device=device)
        for n in ("self_attn_layer_norm", "final_layer_norm"):
            w[p + n + ".weight"] = weight((dim,), f"esm.{i}.{n}.g", 0.05, seed, mean=1.0, device=device)
            w[p + n + ".bias"] = weight((dim,), f"esm.{i}.{n}.b", 0.05, seed, device=device)
    w["emb_layer_norm_after.weight"] = weight((dim,), "esm.lnf.g", 0.05, seed, mean=1.0, device=device)
    w["emb_layer_norm_after.bias"] = weight((dim,), "esm.lnf.b", 0.05, seed, device=device)
    return w


def projector_weights(in_dim: int, cstp_dim: int, hidden: int, seed: int = 0, device="cpu") -> dict:
    """CSTP projection (`protein_projection.linear.*`) and switch projector (`0.*`, `2.*`) weights."""
    return {
        "protein_projection.linear.weight": weight((cstp_dim, in_dim), "proj.cstp.w", 1.0, seed, device=device),
        "protein_projection.linear.bias": weight((cstp_dim,), "proj.cstp.b", 0.02, seed, device=device),
        "0.weight": weight((hidden, cstp_dim), "proj.0.w", cstp_dim ** -0.5, seed, device=device),
        "0.bias": weight((hidden,), "proj.0.b", 0.02, seed, device=device),
        "2.weight": weight((hidden, hidden), "proj.2.w", hidden ** -0.5, seed, device=device),
        "2.bias": weight((hidden,), "proj.2.b", 0.02, seed, device=device),
    }


def llama_weights(n_layers: int, dim: int, n_q_heads: int, n_kv_heads: int, head_dim: int, ffn: int, vocab: int,
                  seed: int = 0, peaked: bool = False, dtype=torch.float32, device="cpu", qkv_bias: bool = False) -> dict:
    """HF state-dict names. Default = HF init statistics (std 0.02). peaked=True is the token-parity recipe: damped
    residual branches and lm_head tied to a permutation of the embedding rows, so greedy argmax margins dwarf bf16
    noise (hard part 1)."""
    qd, kd = n_q_heads * head_dim, n_kv_heads * head_dim
    emb_std = 1.0 if peaked else 0.02
    out_std = 0.02 / math.sqrt(2 * n_layers) if peaked else 0.02
    w = {"model.embed_tokens.weight": weight((vocab, dim), "llama.embed", emb_std, seed, device=device).to(dtype)}
    for i in range(n_layers):
        p = f"model.layers.{i}."
        w[p + "input_layernorm.weight"] = weight((dim,), f"llama.{i}.ln1", 0.05, seed, mean=1.0, device=device).to(dtype)
        w[p + "post_attention_layernorm.weight"] = weight((dim,), f"llama.{i}.ln2", 0.05, seed, mean=1.0, device=device).to(dtype)
        w[p + "self_attn.q_proj.weight"] = weight((qd, dim), f"llama.{i}.q", 0.02, seed, device=device).to(dtype)
        w[p + "self_attn.k_proj.weight"] = weight((kd, dim), f"llama.{i}.k", 0.02, seed, device=device).to(dtype)
        w[p + "self_attn.v_proj.weight"] = weight((kd, dim), f"llama.{i}.v", 0.02, seed, device=device).to(dtype)
        if qkv_bias:   # Qwen2 family: biases on the q/k/v projections
            w[p + "self_attn.q_proj.bias"] = weight((qd,), f"llama.{i}.qb", 0.1, seed, device=device).to(dtype)
            w[p + "self_attn.k_proj.bias"] = weight((kd,), f"llama.{i}.kb", 0.1, seed, device=device).to(dtype)
            w[p + "self_attn.v_proj.bias"] = weight((kd,), f"llama.{i}.vb", 0.1, seed, device=device).to(dtype)
        w[p + "self_attn.o_proj.weight"] = weight((dim, qd), f"llama.{i}.o", out_std, seed, device=device).to(dtype)
        w[p + "mlp.gate_proj.weight"] = weight((ffn, dim), f"llama.{i}.gate", 0.02, seed, device=device).to(dtype)
        w[p + "mlp.up_proj.weight"] = weight((ffn, dim), f"llama.{i}.up", 0.02, seed, device=device).to(dtype)
        w[p + "mlp.down_proj.weight"] = weight((dim, ffn), f"llama.{i}.down", out_std, seed, device=device).to(dtype)
    w["model.norm.weight"] = weight((dim,), "llama.norm", 0.05, seed, mean=1.0, device=device).to(dtype)
    if peaked:
        g = torch.Generator().manual_seed(seed + 17)
        perm = torch.randperm(vocab, generator=g).to(device)
        w["lm_head.weight"] = (w["model.embed_tokens.weight"].float()[perm] * (8.0 / math.sqrt(dim))).to(dtype)
    else:
        w["lm_head.weight"] = weight((vocab, dim), "llama.lm_head", 0.02, seed, device=device).to(dtype)
    return w


def lora_adapters(w: dict, n_layers: int, r: int = 16, seed: int = 0, device="cpu",
                  targets=("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj")) -> dict:
    """Random NON-zero LoRA A/B for every target linear (peft names lora_A / lora_B), std 0.02."""
    out = {}
    for i in range(n_layers):
        for t in targets:
            sub = "self_attn" if t in ("q_proj", "k_proj", "v_proj", "o_proj") else "mlp"
            key = f"model.layers.{i}.{sub}.{t}"
            o, k = w[key + ".weight"].shape
            out[key + ".lora_A.weight"] = weight((r, k), f"lora.{i}.{t}.A", 0.02, seed, device=device)
            out[key + ".lora_B.weight"] = weight((o, r), f"lora.{i}.{t}.B", 0.02, seed, device=device)
    return out


def opt_weights(n_layers: int, dim: int, n_heads: int, ffn: int, vocab: int, max_pos: int = 2048, seed: int = 0,
                peaked: bool = False, dtype=torch.float32, device="cpu", bias: bool = True) -> dict:
    """HF OPT / Galactica state-dict names (`model.decoder.*`, learned positions with the offset-2 table, LayerNorm with
    bias, fc1/fc2; the sibling family of language_model/opus_opt.py). `lm_head.weight` is tied to the embedding in real
    OPT checkpoints; here it is a separate tensor so that peaked=True can use the token-parity recipe of llama_weights.
    bias=False is Galactica's `enable_bias: false`."""
    emb_std = 1.0 if peaked else 0.02
    out_std = 0.02 / math.sqrt(2 * n_layers) if peaked else 0.02
    p0 = "model.decoder."
    w = {p0 + "embed_tokens.weight": weight((vocab, dim), "opt.embed", emb_std, seed, device=device).to(dtype),
         p0 + "embed_positions.weight": weight((max_pos + 2, dim), "opt.pos", 0.3 * emb_std, seed, device=device).to(dtype)}
    for i in range(n_layers):
        p = f"{p0}layers.{i}."
        for n, std in (("q_proj", 0.02), ("k_proj", 0.02), ("v_proj", 0.02), ("out_proj", out_std)):
            w[p + f"self_attn.{n}.weight"] = weight((dim, dim), f"opt.{i}.{n}.w", std, seed, device=device).to(dtype)
            if bias:
                w[p + f"self_attn.{n}.bias"] = weight((dim,), f"opt.{i}.{n}.b", 0.05, seed, device=device).to(dtype)
        w[p + "fc1.weight"] = weight((ffn, dim), f"opt.{i}.fc1.w", 0.02, seed, device=device).to(dtype)
        w[p + "fc2.weight"] = weight((dim, ffn), f"opt.{i}.fc2.w", out_std, seed, device=device).to(dtype)
        if bias:
            w[p + "fc1.bias"] = weight((ffn,), f"opt.{i}.fc1.b", 0.05, seed, device=device).to(dtype)
            w[p + "fc2.bias"] = weight((dim,), f"opt.{i}.fc2.b", 0.05 * (out_std / 0.02), seed, device=device).to(dtype)
        for n in ("self_attn_layer_norm", "final_layer_norm"):
            w[p + n + ".weight"] = weight((dim,), f"opt.{i}.{n}.g", 0.05, seed, mean=1.0, device=device).to(dtype)
            w[p + n + ".bias"] = weight((dim,), f"opt.{i}.{n}.b", 0.05, seed, device=device).to(dtype)
    w[p0 + "final_layer_norm.weight"] = weight((dim,), "opt.lnf.g", 0.05, seed, mean=1.0, device=device).to(dtype)
    w[p0 + "final_layer_norm.bias"] = weight((dim,), "opt.lnf.b", 0.05, seed, device=device).to(dtype)
    if peaked:
        g = torch.Generator().manual_seed(seed + 17)
        perm = torch.randperm(vocab, generator=g).to(device)
        w["lm_head.weight"] = (w[p0 + "embed_tokens.weight"].float()[perm] * (8.0 / math.sqrt(dim))).to(dtype)
    else:
        w["lm_head.weight"] = weight((vocab, dim), "opt.lm_head", 0.02, seed, device=device).to(dtype)
    return w
