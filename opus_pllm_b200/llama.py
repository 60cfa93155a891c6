"""Llama-3 prefill + greedy decode over a paged KV cache, driven through the C ABI.

Stands in for what `OpusLlamaForCausalLM.generate` delegates to HF (`LlamaForCausalLM.forward` + `GenerationMixin`,
multi_modality_v1/model/language_model/opus_llama.py:127-132): prefill over the spliced prompt embeddings, then a
greedy loop that is one CUDA-graph replay per token with argmax / EOS bookkeeping on the device.
Weights use HF state-dict names; LoRA adapters are merged at load (model/builder.py:107-109).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L
from . import ops

BLOCK = 16  # tokens per KV-cache page
KERNEL_HEAD_DIM = 128  # head width of the attention / paged-cache kernels


def pad_heads(t: torch.Tensor, n_heads: int, hd_real: int, dim: int, rotary: bool) -> torch.Tensor:
    """Re-lay the head axis (`dim` of `t`, length n_heads * hd_real) out as n_heads * 128 with zero padding.
    rotary=True keeps HF's rotate_half pairing (j, j + hd/2) on the kernels' pairing (j, j + 64):
    [first half | zeros | second half | zeros]; rotary=False (OPT) is [head | zeros]. No-op for 128-wide heads."""
    if hd_real == KERNEL_HEAD_DIM:
        return t
    t = t.movedim(dim, 0)
    v = t.reshape(n_heads, hd_real, *t.shape[1:])
    out = v.new_zeros((n_heads, KERNEL_HEAD_DIM) + tuple(v.shape[2:]))
    if rotary:
        h = hd_real // 2
        out[:, :h] = v[:, :h]
        out[:, KERNEL_HEAD_DIM // 2: KERNEL_HEAD_DIM // 2 + h] = v[:, h:]
    else:
        out[:, :hd_real] = v
    return out.reshape((n_heads * KERNEL_HEAD_DIM,) + tuple(v.shape[2:])).movedim(0, dim).contiguous()


class BlockAllocator:
    """Free-list page allocator for the paged KV cache (pages are recycled between generate() calls)."""

    def __init__(self, num_blocks: int):
        self.num_blocks = num_blocks
        self.free = list(range(num_blocks - 1, -1, -1))

    def alloc(self, n: int) -> list[int]:
        if n > len(self.free):
            raise L.OpusError(f"KV cache exhausted: need {n} pages, {len(self.free)} free")
        out = self.free[-n:][::-1]
        del self.free[-n:]
        return out

    def release(self, blocks):
        self.free.extend(reversed(list(blocks)))


class B200Llama:
    def __init__(self, weights: dict, n_layers: int, dim: int, n_q_heads: int, n_kv_heads: int, head_dim: int,
                 ffn_dim: int, vocab: int, rms_eps: float = 1e-5, rope_theta: float = 500000.0, device="cuda",
                 max_positions: int = 8192, lora: dict | None = None, lora_alpha: float = 32.0, lora_r: int = 16):
        L.load()
        self.device = torch.device(device)
        # The attention / KV-cache kernels work on 128-column heads. Narrower heads (Qwen2-0.5B: 64) are stored zero-padded
        # to 128 columns in the rotate_half layout [first half | 0 | second half | 0]: q.k^T, the rotary pairing (j, j+64)
        # and P.V are unchanged by the zero columns, the softmax scale uses the real width (opus_llama_model.head_dim_real).
        self.hd_real = head_dim
        if head_dim != KERNEL_HEAD_DIM:
            if head_dim > KERNEL_HEAD_DIM or head_dim % 2:
                raise NotImplementedError(f"head_dim {head_dim} is not supported (even values up to {KERNEL_HEAD_DIM})")
            head_dim = KERNEL_HEAD_DIM
        self.n_layers, self.dim, self.Hq, self.Hkv, self.hd = n_layers, dim, n_q_heads, n_kv_heads, head_dim
        self.ffn, self.vocab, self.rms_eps, self.rope_theta = ffn_dim, vocab, rms_eps, rope_theta
        self.qkv_n = (n_q_heads + 2 * n_kv_heads) * head_dim
        b16 = lambda t: t.detach().to(self.device, torch.bfloat16).contiguous()  # noqa: E731
        hr = self.hd_real
        pad_out = lambda W, nh: pad_heads(W, nh, hr, 0, rotary=True)        # noqa: E731  rows / entries = head outputs
        pad_in = lambda W, nh: pad_heads(W, nh, hr, 1, rotary=True)         # noqa: E731  columns = head inputs (o_proj)

        def merged(key: str) -> torch.Tensor:
            W = b16(weights[key + ".weight"])
            if lora is not None and key + ".lora_A.weight" in lora:
                if W.data_ptr() == weights[key + ".weight"].data_ptr():
                    W = W.clone()
                ops.lora_merge_(W, b16(lora[key + ".lora_A.weight"]), b16(lora[key + ".lora_B.weight"]),
                                lora_alpha / lora_r)
            return W

        self.embed = b16(weights["model.embed_tokens.weight"])
        self._keep = []
        layers = (L.LlamaLayer * n_layers)()
        for i in range(n_layers):
            p = f"model.layers.{i}."
            q, k, v = (merged(p + f"self_attn.{n}") for n in ("q_proj", "k_proj", "v_proj"))
            q, k, v = pad_out(q, n_q_heads), pad_out(k, n_kv_heads), pad_out(v, n_kv_heads)
            g, u = merged(p + "mlp.gate_proj"), merged(p + "mlp.up_proj")
            bq = [weights.get(p + f"self_attn.{n}.bias") for n in ("q_proj", "k_proj", "v_proj")]
            t = dict(ln1_w=b16(weights[p + "input_layernorm.weight"]),
                     wqkv=torch.cat([q, k, v], 0).contiguous(),
                     wo=pad_in(merged(p + "self_attn.o_proj"), n_q_heads),
                     ln2_w=b16(weights[p + "post_attention_layernorm.weight"]),
                     # rows interleaved (gate_0, up_0, gate_1, up_1, ...) so SwiGLU lives in the GEMM epilogue
                     wgu=torch.stack([g, u], 1).reshape(2 * ffn_dim, dim).contiguous(),
                     wdown=merged(p + "mlp.down_proj"))
            if any(b is not None for b in bq):   # Qwen2 family: q/k/v projections carry a bias (o_proj does not)
                if any(b is None for b in bq):
                    raise L.OpusError(f"layer {i}: q/k/v projection biases must be given together")
                t["bqkv"] = torch.cat([pad_out(b.detach().to(self.device, torch.float32), nh)
                                       for b, nh in zip(bq, (n_q_heads, n_kv_heads, n_kv_heads))]).contiguous()
            del q, k, v, g, u
            self._keep.append(t)
            for kk, vv in t.items():
                setattr(layers[i], kk, vv.data_ptr())
        self._layers = layers
        self.norm_w = b16(weights["model.norm.weight"])
        # tie_word_embeddings (small Qwen2 variants): no separate lm_head tensor
        self.lm_head = b16(weights["lm_head.weight"]) if "lm_head.weight" in weights else self.embed
        self._build_rope(max_positions)
        self._cache = None
        self._alloc = None
        self._ws_rows = 0
        self._ws_seqs = 0

    # HF LlamaRotaryEmbedding (default rope): fp32 angles, cos/sin cast to the activation dtype (bf16)
    def _build_rope(self, max_pos: int):
        hd = self.hd_real
        inv_freq = 1.0 / (self.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
        fr = torch.outer(torch.arange(max_pos, dtype=torch.float32), inv_freq)
        if hd != self.hd:   # padded heads: angle 0 (cos 1, sin 0) on the zero columns
            fr = torch.cat([fr, torch.zeros(max_pos, (self.hd - hd) // 2)], -1)
        emb = torch.cat([fr, fr], -1)
        self.rope_cos = emb.cos().to(torch.bfloat16).to(self.device).contiguous()
        self.rope_sin = emb.sin().to(torch.bfloat16).to(self.device).contiguous()
        self.max_positions = max_pos
        m = L.LlamaModel()
        m.n_layers, m.dim, m.n_q_heads, m.n_kv_heads, m.head_dim = self.n_layers, self.dim, self.Hq, self.Hkv, self.hd
        m.head_dim_real = 0 if self.hd_real == self.hd else self.hd_real
        m.ffn_dim, m.vocab, m.rope_max_pos, m.rms_eps = self.ffn, self.vocab, max_pos, self.rms_eps
        m.embed = self.embed.data_ptr()
        m.layers = C.cast(self._layers, C.POINTER(L.LlamaLayer))
        m.norm_w, m.lm_head = self.norm_w.data_ptr(), self.lm_head.data_ptr()
        m.rope_cos, m.rope_sin = self.rope_cos.data_ptr(), self.rope_sin.data_ptr()
        self._model = m
        L.load().opus_release_graphs()

    # ------------------------------------------------------------------------------------------ buffers
    def _ensure_cache(self, need_blocks: int):
        """Make sure `need_blocks` pages are FREE. Growing re-allocates the (zero-initialised) cache: page ids held by
        live plans stay valid, their contents do not survive — plans are re-prefilled before every decode."""
        free = len(self._alloc.free) if self._alloc is not None else 0
        if self._cache is None or free < need_blocks:
            L.load().opus_release_graphs()
            old = self._cache_blocks if self._cache is not None else 0
            nb = max(old + need_blocks - free, 64)
            self._cache = None
            self._k = self._v = None
            shape = (self.n_layers, nb, self.Hkv, BLOCK, self.hd)
            self._k = torch.zeros(shape, dtype=torch.bfloat16, device=self.device)
            self._v = torch.zeros(shape, dtype=torch.bfloat16, device=self.device)
            kv = L.KvCache()
            kv.k, kv.v, kv.num_blocks, kv.block_size = self._k.data_ptr(), self._v.data_ptr(), nb, BLOCK
            self._cache, self._cache_blocks = kv, nb
            if self._alloc is None:
                self._alloc = BlockAllocator(nb)
            else:
                self._alloc.free[:0] = list(range(nb - 1, old - 1, -1))
                self._alloc.num_blocks = nb
        return self._cache

    def _ensure_ws(self, rows: int, n_seqs: int):
        if rows > self._ws_rows or n_seqs > self._ws_seqs:
            L.load().opus_release_graphs()
            rows = max(rows, self._ws_rows, 256)
            n_seqs = max(n_seqs, self._ws_seqs, 8)
            dev, bf = self.device, torch.bfloat16
            # split-K partial sums (up to 16 slices) + the per-slab sums of squares of the norm-fused decode path
            part_bytes = 16 * n_seqs * max(self.qkv_n, self.dim) * 4 + (self.dim // 32) * 64 * 4
            b = dict(h=torch.empty((rows, self.dim), dtype=bf, device=dev),
                     xn=torch.empty((rows, self.dim), dtype=bf, device=dev),
                     qkv=torch.empty((rows, self.qkv_n), dtype=bf, device=dev),
                     attn=torch.empty((rows, self.Hq * self.hd), dtype=bf, device=dev),
                     act=torch.empty((rows, self.ffn), dtype=bf, device=dev),
                     partial=torch.empty((part_bytes // 4,), dtype=torch.float32, device=dev),
                     last_h=torch.empty((n_seqs, self.dim), dtype=bf, device=dev),
                     logits=torch.empty((n_seqs, self.vocab), dtype=bf, device=dev))
            ws = L.LlamaWorkspace()
            for k, v in b.items():
                setattr(ws, k, v.data_ptr())
            ws.partial_bytes = part_bytes
            self._ws, self._ws_bufs, self._ws_rows, self._ws_seqs = ws, b, rows, n_seqs
        return self._ws

    def _upload(self, arr: np.ndarray) -> torch.Tensor:
        return ops.h2d(arr, self.device)

    # ------------------------------------------------------------------------------------------ forward
    def make_plan(self, cu_seqlens, max_new_tokens: int) -> dict:
        """Host-side planning for one batch: KV pages, slot mapping and positions, uploaded once (pinned -> device).
        The plan owns its KV pages until `release_plan`."""
        cu = np.asarray(cu_seqlens, dtype=np.int32)
        n_seqs, n_tok = len(cu) - 1, int(cu[-1])
        lens = np.diff(cu)
        max_len = int(lens.max())
        total = max_len + max_new_tokens
        if total > self.max_positions:
            self._build_rope(1 << (total - 1).bit_length())
        max_blocks = (total + BLOCK - 1) // BLOCK
        self._ensure_cache(n_seqs * max_blocks)
        self._ensure_ws(max(n_tok, n_seqs), n_seqs)
        blocks = self._alloc.alloc(n_seqs * max_blocks)
        bt = np.asarray(blocks, dtype=np.int32).reshape(n_seqs, max_blocks)
        seq_of = np.repeat(np.arange(n_seqs, dtype=np.int32), lens)
        pos = (np.arange(n_tok, dtype=np.int32) - cu[:-1][seq_of]).astype(np.int32)
        slot = (bt[seq_of, pos // BLOCK] * BLOCK + pos % BLOCK).astype(np.int32)
        last_rows = (cu[1:] - 1).astype(np.int32)
        return dict(pos=self._upload(pos), slot=self._upload(slot), cu=self._upload(cu), last=self._upload(last_rows),
                    bt=self._upload(bt), ctx=self._upload(lens.astype(np.int32)), n_seqs=n_seqs, n_tok=n_tok,
                    max_len=max_len, max_blocks=max_blocks, blocks=blocks, max_new=max_new_tokens)

    def release_plan(self, plan: dict):
        if plan.get("blocks") is not None:
            self._alloc.release(plan["blocks"])
            plan["blocks"] = None

    @torch.no_grad()
    def prefill(self, embeds: torch.Tensor, cu_seqlens=None, max_new_tokens: int = 1, plan: dict | None = None):
        """embeds bf16 [n_tok, dim] packed. Runs the prompt through all layers, fills the paged cache, leaves the logits
        of the last prompt token of every sequence in plan['logits']."""
        d = plan if plan is not None else self.make_plan(cu_seqlens, max_new_tokens)
        n_seqs, n_tok = d["n_seqs"], d["n_tok"]
        embeds = embeds.contiguous()
        assert embeds.dtype == torch.bfloat16 and embeds.shape == (n_tok, self.dim)
        rc = L.load().opus_llama_prefill(C.byref(self._model), C.byref(self._cache), C.byref(self._ws),
                                         embeds.data_ptr(), d["pos"].data_ptr(), d["slot"].data_ptr(),
                                         d["cu"].data_ptr(), d["last"].data_ptr(), n_seqs, n_tok, d["max_len"],
                                         torch.cuda.current_stream().cuda_stream)
        L.check(rc, "opus_llama_prefill")
        d["logits"] = self._ws_bufs["logits"][:n_seqs]
        d["embeds"] = embeds
        return d

    @torch.no_grad()
    def score_packed(self, embeds: torch.Tensor, cu_seqlens, targets: torch.Tensor, return_logits: bool = False,
                     chunk_rows: int = 4096):
        """Teacher-forced scoring of packed sequences: `targets` int32 [n_tok] holds, for every row, the id the model
        should predict NEXT (already shifted; < 0 = not counted). Returns (per-row fp32 losses, logits bf16 | None).
        All-position logits are produced chunk by chunk (final RMSNorm -> lm_head -> fp32 cross entropy), so the
        [n_tok, vocab] matrix only exists when return_logits is set."""
        st = self.prefill(embeds, cu_seqlens, 1)
        try:
            n_tok = st["n_tok"]
            hidden = self._ws_bufs["h"][:n_tok]                    # residual stream after the last layer
            losses = torch.empty((n_tok,), dtype=torch.float32, device=self.device)
            all_logits = torch.empty((n_tok, self.vocab), dtype=torch.bfloat16, device=self.device) if return_logits else None
            targets = targets.to(self.device, torch.int32).contiguous()
            for r0 in range(0, n_tok, chunk_rows):
                r1 = min(n_tok, r0 + chunk_rows)
                xn = self._final_norm(hidden[r0:r1])
                logits = ops.gemm(xn, self.lm_head, out=None if all_logits is None else all_logits[r0:r1])
                losses[r0:r1] = ops.cross_entropy_rows(logits, targets[r0:r1])
        finally:
            self.release_plan(st)
        return losses, all_logits

    def _final_norm(self, hidden: torch.Tensor) -> torch.Tensor:
        return ops.rmsnorm(hidden, self.norm_w, self.rms_eps)

    def _decode_state(self, st: dict, max_new_tokens: int, eos_ids, pad_id: int, sampling=None, stop_sequences=()):
        """sampling = None (greedy) or (temperature, top_p, seed); stop_sequences = token-id sequences that finish a row."""
        n, dev = st["n_seqs"], self.device
        stops = tuple(tuple(int(t) for t in q) for q in (stop_sequences or ()) if len(q))
        # the seed travels through a device buffer (state.seed_ptr), so one state / one captured graph serves every
        # sampled call: a fresh seed per generate() (the reference's default decode) re-captures nothing
        key = (n, max_new_tokens, tuple(eos_ids), pad_id, None if sampling is None else tuple(sampling[:2]), stops)
        cached = getattr(self, "_state_cache", None)
        if cached is None or cached[0] != key:
            i32 = lambda *s: torch.zeros(s, dtype=torch.int32, device=dev)  # noqa: E731
            bufs = dict(next_tok=i32(n), ctx_len=i32(n), pos=i32(n), slot=i32(n), block_table=i32(n, st["max_blocks"]),
                        finished=i32(n), n_unfinished=i32(1), step=i32(1), out_ids=i32(n, max_new_tokens),
                        eos=torch.tensor(list(eos_ids) or [-1], dtype=torch.int32, device=dev),
                        seed=torch.zeros(1, dtype=torch.int64, device=dev))
            if stops:
                ld = max(len(q) for q in stops)
                bufs["stop_seqs"] = torch.tensor([list(q) + [-1] * (ld - len(q)) for q in stops], dtype=torch.int32,
                                                 device=dev)
                bufs["stop_lens"] = torch.tensor([len(q) for q in stops], dtype=torch.int32, device=dev)
            self._state_cache = (key, bufs)
            L.load().opus_release_graphs()
        bufs = self._state_cache[1]
        if bufs["block_table"].shape != st["bt"].shape:
            bufs["block_table"] = torch.zeros_like(st["bt"])
            L.load().opus_release_graphs()
        bufs["block_table"].copy_(st["bt"])
        bufs["ctx_len"].copy_(st["ctx"])
        bufs["finished"].zero_()
        bufs["step"].zero_()
        bufs["n_unfinished"].fill_(n)
        bufs["out_ids"].fill_(pad_id)
        s = L.DecodeState()
        for k in ("next_tok", "ctx_len", "pos", "slot", "block_table", "finished", "n_unfinished", "step", "out_ids"):
            setattr(s, k, bufs[k].data_ptr())
        s.max_blocks, s.out_ld = st["max_blocks"], max_new_tokens
        s.eos_ids, s.n_eos, s.pad_id = bufs["eos"].data_ptr(), len(eos_ids), pad_id
        if sampling is not None:
            s.do_sample, s.temperature, s.top_p, s.seed = 1, float(sampling[0]), float(sampling[1]), 0
            bufs["seed"].fill_(u64_as_i64(int(sampling[2])))
            s.seed_ptr = bufs["seed"].data_ptr()
        if stops:
            s.stop_seqs, s.stop_lens = bufs["stop_seqs"].data_ptr(), bufs["stop_lens"].data_ptr()
            s.n_stop, s.stop_ld = len(stops), bufs["stop_seqs"].shape[1]
        return s, bufs

    @torch.no_grad()
    def generate_packed(self, embeds: torch.Tensor, cu_seqlens, max_new_tokens: int, eos_ids=(), pad_id: int = 0,
                        use_graph: bool = True, check_every: int = 16, return_prefill_logits: bool = False,
                        plan: dict | None = None, sampling=None, stop_sequences=()):
        """Greedy generation from packed prompt embeddings. Returns int64 [n_seqs, n_new] (new tokens only; finished
        rows padded with pad_id; trimmed at the step where every row had finished, like HF). stop_sequences: token-id
        sequences (e.g. the ids of "###") after which a row counts as finished -- checked on the device every step."""
        own_plan = plan is None
        st = self.prefill(embeds, cu_seqlens, max_new_tokens, plan=plan)
        try:
            prefill_logits = st["logits"].clone() if return_prefill_logits else None
            out = self.generate_from_prefill(st, max_new_tokens, eos_ids, pad_id, use_graph, check_every, sampling,
                                             stop_sequences)
        finally:
            if own_plan:
                self.release_plan(st)
        return (out, prefill_logits) if return_prefill_logits else out

    @torch.no_grad()
    def generate_from_prefill(self, st: dict, max_new_tokens: int, eos_ids=(), pad_id: int = 0,
                              use_graph: bool = True, check_every: int = 16, sampling=None,
                              stop_sequences=()) -> torch.Tensor:
        """Select the first token from the prefill logits, then run the decode loop (CUDA-graph replays). Greedy unless
        sampling = (temperature, top_p, seed): HF do_sample=True semantics, drawn on the device."""
        lib = L.load()
        s, bufs = self._decode_state(st, max_new_tokens, eos_ids, pad_id, sampling, stop_sequences)
        stream = torch.cuda.current_stream().cuda_stream
        L.check(lib.opus_llama_select(C.byref(self._model), C.byref(self._ws), C.byref(s), st["n_seqs"], stream),
                "opus_llama_select")
        if max_new_tokens > 1:
            rc = lib.opus_llama_decode_loop(C.byref(self._model), C.byref(self._cache), C.byref(self._ws),
                                            C.byref(s), st["n_seqs"], max_new_tokens - 1,
                                            check_every if (len(eos_ids) or s.n_stop) else 0, int(use_graph), stream)
            L.check(rc, "opus_llama_decode_loop")
        out = bufs["out_ids"].to(torch.int64)
        if len(eos_ids):
            out = _trim_like_hf(out, eos_ids)
        return out


def u64_as_i64(v: int) -> int:
    """the bit pattern of an unsigned 64-bit seed as the signed value a torch.int64 buffer holds"""
    v &= 0xFFFFFFFFFFFFFFFF
    return v - (1 << 64) if v >= (1 << 63) else v


def _trim_like_hf(out: torch.Tensor, eos_ids) -> torch.Tensor:
    """HF stops at the first step where every row has emitted an EOS; later columns do not exist in its output."""
    is_eos = torch.zeros_like(out, dtype=torch.bool)
    for e in eos_ids:
        is_eos |= out == e
    has = is_eos.any(1)
    if not bool(has.all()):
        return out
    first = is_eos.float().argmax(1)
    return out[:, : int(first.max()) + 1]
