"""ORACLE (test infrastructure only). Restatement of the reference's own multimodal glue — projectors and the soft-token
splice — in plain torch. Unlike the encoder/LLM arithmetic these functions DO live under /root/reference, so they are
pinned directly against the reference classes executed in the build container (oracle/make_golden.py ->
tests/golden/mm_small.pt).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .ops_ref import gelu_erf

SEQ_TOKEN_INDEX = -200  # multi_modality_v1/constants.py:8
IGNORE_INDEX = -100     # multi_modality_v1/constants.py:7


def protein_forward(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, matmul_dtype=None) -> torch.Tensor:
    """CSTPBase.protein_forward (cstp_v3/modelling.py:396-400): F.normalize(dim=-1) then Linear(1280, 5120)."""
    x = F.normalize(x.float(), dim=-1)
    if matmul_dtype is not None:
        return F.linear(x.to(matmul_dtype), w.to(matmul_dtype), b.to(matmul_dtype))
    return F.linear(x, w, b)


def switch_projector(x: torch.Tensor, sd: dict, hidden_size: int, matmul_dtype=None) -> torch.Tensor:
    """build_switch_projector 'mlp2x_gelu' (protein_mlp/builder.py:11-25) + reshape to [B, 8, H]
    (opus_arch.py:122-131). sd keys: '0.weight','0.bias','2.weight','2.bias' ('linear' type: 'weight','bias')."""
    c = (lambda t: t.to(matmul_dtype)) if matmul_dtype is not None else (lambda t: t)
    if "weight" in sd:
        y = F.linear(c(x), c(sd["weight"]), c(sd["bias"]))
    else:
        h = F.linear(c(x), c(sd["0.weight"]), c(sd["0.bias"]))
        h = gelu_erf(h.float()).to(h.dtype) if matmul_dtype is not None else gelu_erf(h)  # nn.GELU() = exact erf
        y = F.linear(h, c(sd["2.weight"]), c(sd["2.bias"]))
    return y.reshape(x.shape[0], -1, hidden_size)


def splice(input_ids: torch.Tensor, attention_mask: torch.Tensor | None, soft: torch.Tensor, embed: torch.Tensor,
           inference_mode: bool = True, max_length: int | None = None):
    """prepare_inputs_labels_for_multimodal (opus_arch.py:133-294), embedding part only.

    input_ids int64 [B, L] with -200 sentinels; soft [n_seq, 8, H]; embed [V, H].
    Returns (inputs_embeds [B, Lmax', H], attention_mask bool [B, Lmax'], position_ids [B, Lmax'], lengths list).
    Rows without a sentinel still consume one soft-token slot (:196-203).
    """
    B, L = input_ids.shape
    mask = torch.ones_like(input_ids, dtype=torch.bool) if attention_mask is None else attention_mask.bool()
    rows, seq_idx = [], 0
    for b in range(B):
        ids = input_ids[b][mask[b]]
        n_prot = int((ids == SEQ_TOKEN_INDEX).sum())
        if n_prot == 0:
            rows.append(embed[ids])
            seq_idx += 1
            continue
        cuts = [-1] + torch.where(ids == SEQ_TOKEN_INDEX)[0].tolist() + [ids.shape[0]]
        parts = []
        for i in range(len(cuts) - 1):
            parts.append(embed[ids[cuts[i] + 1: cuts[i + 1]]])
            if i < n_prot:
                parts.append(soft[seq_idx].to(embed.dtype))
                seq_idx += 1
        rows.append(torch.cat(parts))
    if max_length is not None:
        rows = [r[:max_length] for r in rows]
    lens = [r.shape[0] for r in rows]
    Lm = max(lens)
    H = embed.shape[1]
    out = torch.zeros(B, Lm, H, dtype=embed.dtype, device=embed.device)
    new_mask = torch.zeros(B, Lm, dtype=torch.bool, device=embed.device)
    pos = torch.zeros(B, Lm, dtype=torch.long, device=embed.device)
    for b, r in enumerate(rows):
        n = lens[b]
        if n == 0:
            continue
        sl = slice(Lm - n, Lm) if inference_mode else slice(0, n)
        out[b, sl] = r
        new_mask[b, sl] = True
        pos[b, sl] = torch.arange(n, device=embed.device)
    return out, new_mask, pos, lens


def splice_labels(input_ids: torch.Tensor, attention_mask: torch.Tensor | None, labels: torch.Tensor, n_soft: int = 8,
                  ignore_index: int = -100, max_length: int | None = None) -> torch.Tensor:
    """Label side of prepare_inputs_labels_for_multimodal in training / scoring mode (opus_arch.py:176-269 with
    inference_mode=False): pad positions dropped, every <seq> sentinel replaced by n_soft ignore_index labels, rows
    right-padded with ignore_index. Returns int64 [B, Lmax']."""
    B = input_ids.shape[0]
    mask = torch.ones_like(input_ids, dtype=torch.bool) if attention_mask is None else attention_mask.bool()
    rows = []
    for b in range(B):
        ids, lab = input_ids[b][mask[b]], labels[b][mask[b]]
        out = []
        for t, l in zip(ids.tolist(), lab.tolist()):
            out.extend([ignore_index] * n_soft if t == SEQ_TOKEN_INDEX else [l])
        rows.append(torch.tensor(out[:max_length] if max_length is not None else out, dtype=torch.int64))
    Lm = max(r.numel() for r in rows)
    return torch.stack([torch.cat([r, torch.full((Lm - r.numel(),), ignore_index, dtype=torch.int64)]) for r in rows])
