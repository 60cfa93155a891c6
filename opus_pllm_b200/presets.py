"""Model presets of the OPUS-PLLM-Llama3-8B path and synthetic (random-init) model construction."""
from __future__ import annotations

import torch

from . import synth

ESM2_650M = dict(n_layers=33, dim=1280, n_heads=20, ffn_dim=5120)          # esm2_t33_650M_UR50D (modelling.py:21)
LLAMA3_8B = dict(n_layers=32, dim=4096, n_q_heads=32, n_kv_heads=8, head_dim=128, ffn_dim=14336, vocab=128256)
CSTP_DIM = 5120                                                             # protein_projector/builder.py:7-10
N_SOFT = 8                                                                  # protein_mlp/builder.py:11

ESM2_TINY = dict(n_layers=2, dim=128, n_heads=2, ffn_dim=512)
LLAMA_TINY = dict(n_layers=2, dim=512, n_q_heads=4, n_kv_heads=2, head_dim=128, ffn_dim=1024, vocab=2048)


def synthetic_state_dicts(size: str = "full", device="cuda", peaked: bool = False, seed: int = 0,
                          llama_dtype=torch.bfloat16, with_lora: bool = False):
    """Random-init weights of the named architecture (there are no checkpoints offline)."""
    esm_cfg, lcfg = (ESM2_650M, LLAMA3_8B) if size == "full" else (ESM2_TINY, LLAMA_TINY)
    cstp = CSTP_DIM if size == "full" else 256
    esm_sd = synth.esm2_weights(esm_cfg["n_layers"], esm_cfg["dim"], esm_cfg["ffn_dim"], seed=seed, device=device)
    proj_sd = synth.projector_weights(esm_cfg["dim"], cstp, N_SOFT * lcfg["dim"], seed=seed, device=device)
    llama_sd = synth.llama_weights(lcfg["n_layers"], lcfg["dim"], lcfg["n_q_heads"], lcfg["n_kv_heads"],
                                   lcfg["head_dim"], lcfg["ffn_dim"], lcfg["vocab"], seed=seed, peaked=peaked,
                                   dtype=llama_dtype, device=device)
    lora_sd = synth.lora_adapters(llama_sd, lcfg["n_layers"], r=16, seed=seed, device=device) if with_lora else None
    return dict(esm_cfg=esm_cfg, llama_cfg=lcfg, esm=esm_sd, proj=proj_sd, llama=llama_sd, lora=lora_sd)


def build_synthetic_model(size: str = "full", device="cuda", peaked: bool = False, seed: int = 0,
                          with_lora: bool = False, keep_state: bool = False):
    from .model import build_from_state_dicts
    sd = synthetic_state_dicts(size, device, peaked, seed, with_lora=with_lora)
    model = build_from_state_dicts(sd["llama"], sd["llama_cfg"], sd["esm"], sd["esm_cfg"], sd["proj"], sd["proj"],
                                   lora_sd=sd["lora"], device=device)
    return (model, sd) if keep_state else model
