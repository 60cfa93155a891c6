"""OPT / Galactica decoder behind the same prefill + paged-cache decode machinery as `B200Llama`.

Sibling family of the reference: `OpusOPTForCausalLM` (multi_modality_v1/model/language_model/opus_opt.py), picked by
`load_pretrained_model` for 'opt' / 'galactica' base paths (model/builder.py:71-81). What HF's `OPTDecoder` adds over the
Llama blocks is handled inside libopus_b200.so (`csrc/models.cu::opt_prefill / opt_decode_step`): learned positions
(`embed_positions`, offset 2), LayerNorm with bias, biased q/k/v/out/fc1/fc2, ReLU (OPT) or erf-GELU (Galactica) MLP.
Only `do_layer_norm_before=True` models with `word_embed_proj_dim == hidden_size` are accepted; heads narrower than the
kernels' 128 columns (OPT-125m / 1.3B / 2.7B, Galactica-1.3B: 64 or 80) are stored zero-padded to 128 (llama.pad_heads).
The 350m post-LN variant cannot be loaded by the reference either (its lm_head is built hidden_size wide, opus_opt.py:35).
Weights use the HF state-dict names (`model.decoder.*`, `lm_head.weight` optional = tied).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from . import ops
from .llama import KERNEL_HEAD_DIM, B200Llama, pad_heads

P0 = "model.decoder."


class B200Opt(B200Llama):
    def __init__(self, weights: dict, n_layers: int, dim: int, n_heads: int, ffn_dim: int, vocab: int,
                 max_pos: int = 2048, activation: str = "relu", ln_eps: float = 1e-5, device="cuda",
                 lora: dict | None = None, lora_alpha: float = 32.0, lora_r: int = 16):
        L.load()
        if dim % n_heads or dim // n_heads > KERNEL_HEAD_DIM or (dim // n_heads) % 2:
            raise NotImplementedError(f"OPT family: head_dim {dim / n_heads:g} is not supported (even values up to 128)")
        # narrower heads (OPT-125m / 1.3B, Galactica-1.3B: 64; OPT-2.7B: 80) are stored zero-padded to the kernels' 128
        # columns ([head | zeros]; no rotary pairing to preserve); the scores are scaled by the real width
        self.hd_real = hr = dim // n_heads
        pad_out = lambda W: pad_heads(W, n_heads, hr, 0, rotary=False)   # noqa: E731
        pad_in = lambda W: pad_heads(W, n_heads, hr, 1, rotary=False)    # noqa: E731
        if activation not in ("relu", "gelu"):
            raise NotImplementedError(f"OPT family: activation_function {activation!r} (relu | gelu)")
        # older facebook/opt-* checkpoints store the decoder without the `model.` prefix
        weights = {("model." + k if k.startswith("decoder.") else k): v for k, v in weights.items()}
        self.device = torch.device(device)
        self.n_layers, self.dim, self.Hq, self.Hkv, self.hd = n_layers, dim, n_heads, n_heads, 128
        self.ffn, self.vocab, self.rms_eps, self.rope_theta = ffn_dim, vocab, ln_eps, None
        self.activation = activation
        self.qkv_n = 3 * n_heads * 128
        b16 = lambda t: t.detach().to(self.device, torch.bfloat16).contiguous()  # noqa: E731
        f32 = lambda t: None if t is None else t.detach().to(self.device, torch.float32).contiguous()  # noqa: E731

        def merged(key: str) -> torch.Tensor:
            W = b16(weights[key + ".weight"])
            if lora is not None and key + ".lora_A.weight" in lora:
                if W.data_ptr() == weights[key + ".weight"].data_ptr():
                    W = W.clone()
                ops.lora_merge_(W, b16(lora[key + ".lora_A.weight"]), b16(lora[key + ".lora_B.weight"]),
                                lora_alpha / lora_r)
            return W

        emb_key = P0 + "embed_tokens.weight" if P0 + "embed_tokens.weight" in weights else "model.embed_tokens.weight"
        self.embed = b16(weights[emb_key])
        if self.embed.shape[1] != dim or P0 + "project_in.weight" in weights:
            raise NotImplementedError("OPT family: word_embed_proj_dim != hidden_size (opt-350m) is not supported")
        if P0 + "final_layer_norm.weight" not in weights:
            raise NotImplementedError("OPT family: do_layer_norm_before=False (opt-350m) is not supported")
        self.pos_embed = b16(weights[P0 + "embed_positions.weight"])          # [max_pos + 2, dim]
        self.max_pos = self.pos_embed.shape[0] - 2
        self._keep = []
        layers = (L.LlamaLayer * n_layers)()
        for i in range(n_layers):
            p = f"{P0}layers.{i}."
            q, k, v = (pad_out(merged(p + f"self_attn.{n}")) for n in ("q_proj", "k_proj", "v_proj"))
            bq = [weights.get(p + f"self_attn.{n}.bias") for n in ("q_proj", "k_proj", "v_proj")]
            if any(b is not None for b in bq) and any(b is None for b in bq):
                raise L.OpusError(f"layer {i}: q/k/v projection biases must be given together")
            bq = [None if b is None else pad_out(f32(b)) for b in bq]
            t = dict(wqkv=torch.cat([q, k, v], 0).contiguous(),
                     wo=pad_in(merged(p + "self_attn.out_proj")),
                     wgu=merged(p + "fc1"),                                     # plain fc1 rows (no gate)
                     wdown=merged(p + "fc2"),
                     ln1_g=f32(weights[p + "self_attn_layer_norm.weight"]),
                     ln1_b=f32(weights.get(p + "self_attn_layer_norm.bias")),
                     ln2_g=f32(weights[p + "final_layer_norm.weight"]),
                     ln2_b=f32(weights.get(p + "final_layer_norm.bias")),
                     bqkv=None if bq[0] is None else torch.cat([f32(b) for b in bq]).contiguous(),
                     bo=f32(weights.get(p + "self_attn.out_proj.bias")),
                     b1=f32(weights.get(p + "fc1.bias")),
                     b2=f32(weights.get(p + "fc2.bias")))
            del q, k, v
            t = {kk: vv for kk, vv in t.items() if vv is not None}
            self._keep.append(t)
            for kk, vv in t.items():
                setattr(layers[i], kk, vv.data_ptr())
        self._layers = layers
        self.norm_g = f32(weights[P0 + "final_layer_norm.weight"])
        self.norm_b = f32(weights.get(P0 + "final_layer_norm.bias"))
        self.norm_w = None
        self.lm_head = b16(weights["lm_head.weight"]) if "lm_head.weight" in weights else self.embed   # tied in OPT
        self._build_rope(self.max_pos)
        self._cache = None
        self._alloc = None
        self._ws_rows = 0
        self._ws_seqs = 0

    def _build_rope(self, max_pos: int):
        """OPT has no rotary embedding: the tables hold cos = 1, sin = 0, which turns the fused RoPE + KV-append epilogues
        into a plain append. Positions beyond the learned table cannot exist (HF raises an index error there too)."""
        if max_pos > self.max_pos:
            raise L.OpusError(f"OPT family: {max_pos} positions requested, embed_positions holds {self.max_pos}")
        self.rope_cos = torch.ones((max_pos, self.hd), dtype=torch.bfloat16, device=self.device)
        self.rope_sin = torch.zeros((max_pos, self.hd), dtype=torch.bfloat16, device=self.device)
        self.max_positions = max_pos
        m = L.LlamaModel()
        m.n_layers, m.dim, m.n_q_heads, m.n_kv_heads, m.head_dim = self.n_layers, self.dim, self.Hq, self.Hkv, self.hd
        m.ffn_dim, m.vocab, m.rope_max_pos, m.rms_eps = self.ffn, self.vocab, max_pos, self.rms_eps
        m.head_dim_real = 0 if self.hd_real == self.hd else self.hd_real
        m.embed = self.embed.data_ptr()
        m.layers = C.cast(self._layers, C.POINTER(L.LlamaLayer))
        m.lm_head = self.lm_head.data_ptr()
        m.rope_cos, m.rope_sin = self.rope_cos.data_ptr(), self.rope_sin.data_ptr()
        m.arch, m.opt_act = L.ARCH_OPT, 1 if self.activation == "gelu" else 0
        m.pos_embed, m.pos_rows = self.pos_embed.data_ptr(), self.pos_embed.shape[0]
        m.norm_g = self.norm_g.data_ptr()
        m.norm_b = None if self.norm_b is None else self.norm_b.data_ptr()
        self._model = m
        L.load().opus_release_graphs()

    def _final_norm(self, hidden: torch.Tensor) -> torch.Tensor:
        return ops.layernorm_bf16(hidden, self.norm_g, self.norm_b, self.rms_eps)
