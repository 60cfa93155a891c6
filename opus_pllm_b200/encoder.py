"""Protein encoder: drop-in for the reference's `ProteinSeqEmbeddingExtractor` (cstp_v3/modelling.py:18-77).

Same public surface — ``get_protein_seq_embeddings(list[str]) -> torch.float32[B, 1280]`` on CUDA, mean of the
final-LayerNorm residue states with <cls>/<eos>/<pad> excluded — but the ESM-2 forward runs as ONE call into
``opus_esm2_forward`` (packed variable-length tokens, tcgen05 GEMMs, flash attention, fused LN+pool), not fair-esm.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L

# fair-esm Alphabet "ESM-1b" (esm/data.py): <cls> <pad> <eos> <unk> + 25 residue symbols + '.' '-' + <null_1> <mask>
ESM_TOKS = ["<cls>", "<pad>", "<eos>", "<unk>", "L", "A", "G", "V", "S", "E", "R", "T", "I", "D", "P", "K", "Q", "N",
            "F", "Y", "M", "H", "W", "C", "X", "B", "U", "Z", "O", ".", "-", "<null_1>", "<mask>"]
CLS, PAD, EOS, UNK, MASK = 0, 1, 2, 3, 32
_LUT = np.full(256, UNK, dtype=np.int32)
for _i, _t in enumerate(ESM_TOKS):
    if len(_t) == 1:
        _LUT[ord(_t)] = _i


def _ptr(t):
    return None if t is None else t.data_ptr()


class PackedTokens:
    """Host-side result of tokenisation: packed ids, per-token position and token-dropout scale, cu_seqlens."""

    def __init__(self, seqs: list[str]):
        lens = np.fromiter((len(s) + 2 for s in seqs), dtype=np.int64, count=len(seqs))
        self.cu = np.zeros(len(seqs) + 1, dtype=np.int32)
        np.cumsum(lens, out=self.cu[1:])
        n = int(self.cu[-1])
        self.tokens = np.empty(n, dtype=np.int32)
        self.pos = np.empty(n, dtype=np.int32)
        self.scale = np.empty(n, dtype=np.float32)
        for i, s in enumerate(seqs):
            a, b = int(self.cu[i]), int(self.cu[i + 1])
            self.tokens[a] = CLS
            if len(s):
                self.tokens[a + 1: b - 1] = _LUT[np.frombuffer(s.encode("latin-1", "replace"), dtype=np.uint8)]
            self.tokens[b - 1] = EOS
            self.pos[a:b] = np.arange(b - a, dtype=np.int32)
            # token dropout (ESM2.forward): x *= (1 - 0.15*0.8) / (1 - n_mask / src_len); <mask> rows are zeroed
            seg = self.tokens[a:b]
            n_mask = int((seg == MASK).sum())
            sc = (1 - 0.15 * 0.8) / (1 - n_mask / float(b - a))
            self.scale[a:b] = np.where(seg == MASK, 0.0, sc)
        self.n_tok = n
        self.n_seqs = len(seqs)
        self.max_len = int(lens.max()) if len(seqs) else 0
        self.n_residues = int(sum(len(s) for s in seqs))


class B200ProteinEncoder:
    """ESM-2 (t33 650M by default) on the B200 kernels. `weights` uses fair-esm state-dict names (fp32 tensors)."""

    def __init__(self, weights: dict, n_layers: int = 33, dim: int = 1280, n_heads: int = 20, ffn_dim: int = 5120,
                 device="cuda", max_positions: int = 4096):
        L.load()
        self.device = torch.device(device)
        self.n_layers, self.dim, self.n_heads, self.ffn_dim = n_layers, dim, n_heads, ffn_dim
        self.head_dim = dim // n_heads
        f32 = lambda t: t.detach().to(self.device, torch.float32).contiguous()  # noqa: E731
        b16 = lambda t: t.detach().to(self.device, torch.bfloat16).contiguous()  # noqa: E731
        self._keep = []
        self.embed = f32(weights["embed_tokens.weight"])
        self.vocab = self.embed.shape[0]
        layers = (L.Esm2Layer * n_layers)()
        for i in range(n_layers):
            p = f"layers.{i}."
            wqkv = b16(torch.cat([weights[p + f"self_attn.{n}.weight"] for n in ("q_proj", "k_proj", "v_proj")], 0))
            bqkv = f32(torch.cat([weights[p + f"self_attn.{n}.bias"] for n in ("q_proj", "k_proj", "v_proj")], 0))
            t = dict(
                ln1_g=f32(weights[p + "self_attn_layer_norm.weight"]), ln1_b=f32(weights[p + "self_attn_layer_norm.bias"]),
                wqkv=wqkv, bqkv=bqkv,
                wo=b16(weights[p + "self_attn.out_proj.weight"]), bo=f32(weights[p + "self_attn.out_proj.bias"]),
                ln2_g=f32(weights[p + "final_layer_norm.weight"]), ln2_b=f32(weights[p + "final_layer_norm.bias"]),
                w1=b16(weights[p + "fc1.weight"]), b1=f32(weights[p + "fc1.bias"]),
                w2=b16(weights[p + "fc2.weight"]), b2=f32(weights[p + "fc2.bias"]))
            self._keep.append(t)
            for k, v in t.items():
                setattr(layers[i], k, v.data_ptr())
        self._layers = layers
        self.lnf_g = f32(weights["emb_layer_norm_after.weight"])
        self.lnf_b = f32(weights["emb_layer_norm_after.bias"])
        self._build_rope(max_positions)
        self._ws_tok = 0
        self._ws = None

    # fair-esm RotaryEmbedding: inv_freq = 1/10000^(2i/d), angle = position * inv_freq (fp32)
    def _build_rope(self, max_pos: int):
        hd = self.head_dim
        inv_freq = 1.0 / (10000 ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
        fr = torch.outer(torch.arange(max_pos, dtype=torch.float32), inv_freq)
        self.rope_cos = fr.cos().to(self.device).contiguous()
        self.rope_sin = fr.sin().to(self.device).contiguous()
        self.max_positions = max_pos
        m = L.Esm2Model()
        m.n_layers, m.dim, m.n_heads, m.ffn_dim, m.vocab = self.n_layers, self.dim, self.n_heads, self.ffn_dim, self.vocab
        m.rope_max_pos, m.ln_eps = max_pos, 1e-5
        m.embed = self.embed.data_ptr()
        m.layers = C.cast(self._layers, C.POINTER(L.Esm2Layer))
        m.lnf_g, m.lnf_b = self.lnf_g.data_ptr(), self.lnf_b.data_ptr()
        m.rope_cos, m.rope_sin = self.rope_cos.data_ptr(), self.rope_sin.data_ptr()
        self._model = m

    def _workspace(self, n_tok: int):
        if n_tok > self._ws_tok:
            self._ws = None
            cap = max(n_tok, 1024)
            d, f = self.dim, self.ffn_dim
            bufs = dict(x=torch.empty((cap, d), dtype=torch.float32, device=self.device),
                        xn=torch.empty((cap, d), dtype=torch.bfloat16, device=self.device),
                        qkv=torch.empty((cap, 3 * d), dtype=torch.bfloat16, device=self.device),
                        attn=torch.empty((cap, d), dtype=torch.bfloat16, device=self.device),
                        ffn=torch.empty((cap, f), dtype=torch.bfloat16, device=self.device))
            ws = L.Esm2Workspace()
            for k, v in bufs.items():
                setattr(ws, k, v.data_ptr())
            self._ws, self._ws_bufs, self._ws_tok = ws, bufs, cap
        return self._ws

    def load_model(self):  # opus_arch.py:63 calls this on an existing encoder; nothing to do here
        return self

    @staticmethod
    def tokenize(seqs: list[str]) -> PackedTokens:
        return PackedTokens(seqs)

    def _upload(self, arr: np.ndarray) -> torch.Tensor:
        from .ops import h2d
        return h2d(arr, self.device)

    @torch.no_grad()
    def encode(self, seqs: list[str], want_hidden: bool = False, packed: PackedTokens | None = None):
        """-> (pooled fp32 [B, D], pooled_l2 bf16 [B, D], hidden fp32 [n_tok, D] | None, PackedTokens)"""
        pk = packed if packed is not None else PackedTokens(seqs)
        if pk.max_len > self.max_positions:
            self._build_rope(1 << (pk.max_len - 1).bit_length())
        if getattr(pk, "device_arrays", None) is None:
            pk.device_arrays = tuple(self._upload(a) for a in (pk.tokens, pk.pos, pk.scale, pk.cu))
        tok, pos, scale, cu = pk.device_arrays
        ws = self._workspace(pk.n_tok)
        pooled = torch.empty((pk.n_seqs, self.dim), dtype=torch.float32, device=self.device)
        pooled_l2 = torch.empty((pk.n_seqs, self.dim), dtype=torch.bfloat16, device=self.device)
        hidden = torch.empty((pk.n_tok, self.dim), dtype=torch.float32, device=self.device) if want_hidden else None
        with torch.cuda.device(self.device):
            rc = L.load().opus_esm2_forward(C.byref(self._model), C.byref(ws), tok.data_ptr(), scale.data_ptr(),
                                            pos.data_ptr(), cu.data_ptr(), pk.n_seqs, pk.n_tok, pk.max_len,
                                            pooled.data_ptr(), pooled_l2.data_ptr(), _ptr(hidden),
                                            torch.cuda.current_stream().cuda_stream)
        L.check(rc, "opus_esm2_forward")
        if packed is None:
            self._last_inputs = pk.device_arrays  # keep alive until the stream has consumed them
            pk.device_arrays = None
        return pooled, pooled_l2, hidden, pk

    def get_protein_seq_embeddings(self, data: list[str]) -> torch.Tensor:
        """cstp_v3/modelling.py:37-57 contract: fp32 [B, dim] on CUDA."""
        return self.encode(list(data))[0]
