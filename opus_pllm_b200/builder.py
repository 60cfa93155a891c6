"""Model loading for real OPUS-PLLM weights: counterpart of `multi_modality_v1/model/builder.py::load_pretrained_model`.

Same signature and return value `(tokenizer, model, context_len)`; the files read are the ones the reference reads
("Weight formats to load"):
  1. HF Llama-3 (or Qwen2 / Qwen2.5: same layout + q/k/v biases, builder.py:83-94) directory: `config.json` +
     `*.safetensors` shards                                                              (builder.py:61-65);
     or an HF OPT / Galactica directory (`pytorch_model*.bin`, builder.py:71-82) -> `opt.py::B200Opt`
  2. `<weights>/lora_adapter/{adapter_config.json, adapter_model.safetensors|.bin}`     (builder.py:105-109, peft)
  3. `<weights>/modality_refinement_projector/modality_refinement_projection.bin`       (builder.py:111, opus_arch.py:85-89)
  4. `<weights>/modality_encoder/modality_encoding_adapter.ckpt` (Lightning checkpoint)  (protein_projector/builder.py:16-25)
  5. ESM-2 t33 650M weights: a fair-esm `.pt` (hub cache) or an HF `EsmModel` directory  (cstp_v3/modelling.py:21)
No peft / lightning / fair-esm import is needed: the readers below only parse the state dicts.
`load_8bit` / `load_4bit` are accepted and ignored (this backend is bf16, the north star excludes bitsandbytes).
"""
from __future__ import annotations

import glob
import json
import os
import re
import warnings

import torch


def return_cstp_path(args_path: str, file_name: str) -> str:
    return f"{args_path}{file_name}" if args_path.endswith("/") else f"{args_path}/{file_name}"


# ------------------------------------------------------------------------------------------------ readers (CPU only)
def _load_any(path: str) -> dict:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    return torch.load(path, map_location="cpu", weights_only=False)


def read_hf_llama(model_dir: str) -> tuple[dict, dict]:
    """-> (state dict with HF names, config kwargs for B200Llama)"""
    cfg = json.load(open(os.path.join(model_dir, "config.json")))
    sd = {}
    shards = sorted(glob.glob(os.path.join(model_dir, "*.safetensors"))) or \
        sorted(glob.glob(os.path.join(model_dir, "pytorch_model*.bin")))
    if not shards:
        raise FileNotFoundError(f"no *.safetensors / pytorch_model*.bin under {model_dir}")
    for sh in shards:
        sd.update(_load_any(sh))
    if "lm_head.weight" not in sd and cfg.get("tie_word_embeddings", False):
        sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
    n_heads = cfg["num_attention_heads"]
    kw = dict(n_layers=cfg["num_hidden_layers"], dim=cfg["hidden_size"], n_q_heads=n_heads,
              n_kv_heads=cfg.get("num_key_value_heads", n_heads),
              head_dim=cfg.get("head_dim") or cfg["hidden_size"] // n_heads, ffn_dim=cfg["intermediate_size"],
              vocab=cfg["vocab_size"], rms_eps=cfg.get("rms_norm_eps", 1e-5),
              rope_theta=cfg.get("rope_theta", 500000.0))
    scaling = cfg.get("rope_scaling")
    if scaling and scaling.get("rope_type", scaling.get("type")) not in (None, "default"):
        warnings.warn(f"rope_scaling {scaling} is not applied (Llama-3-8B base uses the default rope)")
    extra = dict(eos_token_id=cfg.get("eos_token_id"), max_sequence_length=cfg.get("max_sequence_length"))
    return sd, dict(kw, **{"_extra": extra})


def read_hf_opt(model_dir: str) -> tuple[dict, dict]:
    """HF OPT / Galactica directory -> (state dict with HF names, kwargs for B200Opt). The reference loads these with
    `use_safetensors=False` (builder.py:73-76), i.e. from pytorch_model*.bin; safetensors shards are read when present."""
    cfg = json.load(open(os.path.join(model_dir, "config.json")))
    shards = sorted(glob.glob(os.path.join(model_dir, "pytorch_model*.bin"))) or \
        sorted(glob.glob(os.path.join(model_dir, "*.safetensors")))
    if not shards:
        raise FileNotFoundError(f"no pytorch_model*.bin / *.safetensors under {model_dir}")
    sd = {}
    for sh in shards:
        sd.update(_load_any(sh))
    if not cfg.get("do_layer_norm_before", True) or cfg.get("word_embed_proj_dim", cfg["hidden_size"]) != cfg["hidden_size"]:
        raise NotImplementedError("OPT post-LN / projected-embedding variants (opt-350m) are not supported")
    kw = dict(n_layers=cfg["num_hidden_layers"], dim=cfg["hidden_size"], n_heads=cfg["num_attention_heads"],
              ffn_dim=cfg["ffn_dim"], vocab=cfg["vocab_size"], max_pos=cfg.get("max_position_embeddings", 2048),
              activation=cfg.get("activation_function", "relu"))
    extra = dict(eos_token_id=cfg.get("eos_token_id"), max_sequence_length=cfg.get("max_sequence_length"))
    return sd, dict(kw, **{"_extra": extra})


def read_peft_lora(adapter_dir: str) -> tuple[dict, float, int]:
    """peft adapter dir -> ({'model.layers.N.<mod>.lora_A.weight': ..., '...lora_B.weight': ...}, alpha, r)"""
    cfg = json.load(open(os.path.join(adapter_dir, "adapter_config.json")))
    f = [p for p in (os.path.join(adapter_dir, "adapter_model.safetensors"), os.path.join(adapter_dir, "adapter_model.bin"))
         if os.path.exists(p)]
    if not f:
        raise FileNotFoundError(f"no adapter_model.safetensors/.bin under {adapter_dir}")
    raw = _load_any(f[0])
    out = {}
    for k, v in raw.items():
        k = re.sub(r"^base_model\.model\.", "", k)
        k = k.replace(".lora_A.default.", ".lora_A.").replace(".lora_B.default.", ".lora_B.")
        if ".lora_A." in k or ".lora_B." in k:
            out[k] = v
    if cfg.get("fan_in_fan_out"):
        raise NotImplementedError("fan_in_fan_out LoRA adapters are not supported")
    return out, float(cfg["lora_alpha"]), int(cfg["r"])


def read_switch_projector(path: str) -> dict:
    """`.bin` with keys containing 'switch_projector.' -> {'0.weight','0.bias','2.weight','2.bias'} (opus_arch.py:85-89)"""
    raw = _load_any(path)
    return {k.split("switch_projector.")[1]: v for k, v in raw.items() if "switch_projector" in k}


def read_cstp_checkpoint(path: str) -> dict:
    """Lightning checkpoint -> {'protein_projection.linear.weight', 'protein_projection.linear.bias'}"""
    raw = _load_any(path)
    sd = raw.get("state_dict", raw)
    out = {k: v for k, v in sd.items() if k.startswith("protein_projection.linear.")}
    if len(out) != 2:
        raise KeyError(f"{path}: protein_projection.linear.{{weight,bias}} not found")
    return out


_HF_ESM = (("attention.self.query", "self_attn.q_proj"), ("attention.self.key", "self_attn.k_proj"),
           ("attention.self.value", "self_attn.v_proj"), ("attention.output.dense", "self_attn.out_proj"),
           ("attention.LayerNorm", "self_attn_layer_norm"), ("intermediate.dense", "fc1"), ("output.dense", "fc2"),
           ("LayerNorm", "final_layer_norm"))


def read_esm2(path: str) -> tuple[dict, dict]:
    """fair-esm `.pt` (keys optionally prefixed 'encoder.sentence_encoder.') or HF EsmModel dir/file ->
    (state dict with fair-esm names, config kwargs for B200ProteinEncoder)"""
    if os.path.isdir(path):
        raw = {}
        for sh in sorted(glob.glob(os.path.join(path, "*.safetensors"))) or sorted(glob.glob(os.path.join(path, "*.bin"))):
            raw.update(_load_any(sh))
    else:
        raw = _load_any(path)
        raw = raw.get("model", raw)
    sd = {}
    for k, v in raw.items():
        k = re.sub(r"^(encoder\.sentence_encoder\.|esm\.|model\.)", "", k)
        if k.startswith("embeddings.word_embeddings."):
            k = k.replace("embeddings.word_embeddings.", "embed_tokens.")
        elif k.startswith("encoder.emb_layer_norm_after."):
            k = k.replace("encoder.emb_layer_norm_after.", "emb_layer_norm_after.")
        elif k.startswith("encoder.layer."):
            m = re.match(r"encoder\.layer\.(\d+)\.(.*)\.(weight|bias)$", k)
            if not m:
                continue
            for hf, fe in _HF_ESM:
                if m.group(2) == hf:
                    k = f"layers.{m.group(1)}.{fe}.{m.group(3)}"
                    break
            else:
                continue
        if k.startswith(("embed_tokens.", "layers.", "emb_layer_norm_after.")):
            sd[k] = v
    n_layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("layers."))
    dim = sd["embed_tokens.weight"].shape[1]
    ffn = sd["layers.0.fc1.weight"].shape[0]
    return sd, dict(n_layers=n_layers, dim=dim, n_heads=dim // 64, ffn_dim=ffn)


# ------------------------------------------------------------------------------------------------ component seams
FAIR_ESM_HUB = "~/.cache/torch/hub/checkpoints/esm2_t33_650M_UR50D.pt"   # what esm.pretrained.esm2_t33_650M_UR50D() caches


def default_esm_path() -> str:
    """$OPUS_ESM_PATH, else the file fair-esm's `esm.pretrained.esm2_t33_650M_UR50D()` downloads into the torch hub cache
    ($TORCH_HOME/hub/checkpoints, default ~/.cache/torch/hub/checkpoints)."""
    env = os.environ.get("OPUS_ESM_PATH")
    if env:
        return env
    home = os.environ.get("TORCH_HOME")
    if home:
        return os.path.join(home, "hub", "checkpoints", os.path.basename(FAIR_ESM_HUB))
    return os.path.expanduser(FAIR_ESM_HUB)


def _eos_ids(model_dir: str, config_eos) -> list[int]:
    """EOS ids HF `generate()` stops on: `generation_config.json` (what the reference's `model.generate` reads;
    Meta-Llama-3-8B-Instruct lists [128001, 128009] there) merged with `config.json`'s `eos_token_id`."""
    ids: list[int] = []
    sources = [config_eos]
    gpath = os.path.join(model_dir, "generation_config.json")
    if os.path.exists(gpath):
        try:
            sources.insert(0, json.load(open(gpath)).get("eos_token_id"))
        except (OSError, ValueError):
            pass
    for src in sources:
        if src is None:
            continue
        for e in (src if isinstance(src, (list, tuple)) else [src]):
            if e is not None and int(e) not in ids:
                ids.append(int(e))
    return ids


def build_protein_encoder(ckpt=None, esm_path: str | None = None, device="cuda"):
    """`protein_encoder/builder.py:3-6` -> `ProteinSeqEmbeddingExtractor(ckpt)` (cstp_v3/modelling.py:19-36): the ESM-2
    t33 650M weights (fair-esm hub cache, or `esm_path`), optionally overridden by the `protein_model.model.*` tensors of
    a CSTP training checkpoint (`torch.load(ckpt)['model']`, loaded non-strictly like the reference). Returns an object
    with `get_protein_seq_embeddings(list[str]) -> float32[B, 1280]` on the GPU (used by opus_arch.py:53 and
    scripts/generate_esm_*.py)."""
    from .encoder import B200ProteinEncoder
    sd, cfg = read_esm2(esm_path or default_esm_path())
    if ckpt is not None:
        raw = _load_any(ckpt)["model"]
        pref = "protein_model.model."
        over = {k[len(pref):]: v for k, v in raw.items() if k.startswith(pref)}
        sd.update({k: v for k, v in over.items() if k in sd})          # strict=False: unknown keys are ignored
    return B200ProteinEncoder(sd, device=device, **cfg)


def build_protein_projector(cstp_chackpoint_path: str, device="cuda"):
    """`protein_projector/builder.py:15-29`: the CSTP Lightning checkpoint -> object with `.protein_forward(x[B,1280]) ->
    [B,5120]`, `.to(device)`, `.parameters()`, `.eval()` (only the protein projection is on the generation path)."""
    from .projector import B200ProteinProjector
    sd = read_cstp_checkpoint(cstp_chackpoint_path)
    return B200ProteinProjector(sd["protein_projection.linear.weight"], sd["protein_projection.linear.bias"], device=device)


def build_switch_projector(model_args, n_tokens: int = 8, device="cuda"):
    """`protein_mlp/builder.py:11-25`: `model_args.hidden_size`, `.switch_projector_type` ('linear' | 'mlp{N}x_gelu',
    default 'mlp2x_gelu'), `.pretrain_protein_projector_ckpt` (None -> the projector reads the 1280-wide ESM embedding,
    else the 5120-wide CSTP one) -> callable with `.load_state_dict({'0.weight', '0.bias', '2.weight', '2.bias'})`."""
    from .projector import B200SwitchProjector
    ptype = getattr(model_args, "switch_projector_type", "mlp2x_gelu")
    in_dim = 5120 if getattr(model_args, "pretrain_protein_projector_ckpt", None) is not None else 1280
    try:
        return B200SwitchProjector(in_dim, model_args.hidden_size * n_tokens, ptype, device=device)
    except NotImplementedError:
        return None   # the reference falls off the end of the function for unknown types (builder.py:25)


# ------------------------------------------------------------------------------------------------ reference entry point
def load_pretrained_model(model_base_path, adapter_path, model_name, load_8bit=False, load_4bit=False,
                          accelerator=None, switch_projector_type="mlp2x_gelu", cstp_path=True, esm_path=None,
                          tokenizer=None, **kwargs):
    """multi_modality_v1/model/builder.py:29-131 contract -> (tokenizer, model, context_len)."""
    from .model import build_from_state_dicts
    if model_name is None or not model_base_path:
        raise NotImplementedError
    name = model_base_path.lower()
    if "llama" in name or "qwen" in name:                                               # builder.py:59-70, 83-94
        family = "llama"
    elif "opt" in name or "galactica" in name:                                          # builder.py:71-82
        family = "opt"
    else:
        raise NotImplementedError                                                       # builder.py:95-96
    if load_8bit or load_4bit:
        warnings.warn("load_8bit/load_4bit are ignored: opus_pllm_b200 runs bf16 weights")
    device = "cuda:0" if accelerator is None else f"cuda:{accelerator.process_index}"
    llama_sd, llama_cfg = read_hf_llama(model_base_path) if family == "llama" else read_hf_opt(model_base_path)
    extra = llama_cfg.pop("_extra")
    if tokenizer is None:
        import transformers
        tokenizer = transformers.AutoTokenizer.from_pretrained(model_base_path, use_fast=False)
        if family == "opt":
            tokenizer.pad_token, tokenizer.unk_token, tokenizer.eos_token = "<pad>", "<unk>", "</s>"   # builder.py:80-82
        elif "llama" in name:
            tokenizer.pad_token = tokenizer.unk_token = tokenizer.eos_token             # builder.py:69-70
            tokenizer.pad_token_id = tokenizer.unk_token_id = tokenizer.eos_token_id
        # the Qwen branch (builder.py:83-94) leaves the tokenizer as loaded
    if accelerator is not None:
        accelerator.wait_for_everyone()                                                 # builder.py:102-103
    lora_sd, alpha, r, switch_sd = None, 32.0, 16, None
    if adapter_path is not None:
        lora_sd, alpha, r = read_peft_lora(return_cstp_path(adapter_path, "lora_adapter"))
        switch_sd = read_switch_projector(return_cstp_path(
            adapter_path, "modality_refinement_projector/modality_refinement_projection.bin"))
    else:
        print("No adapter path!")
    cstp_sd = read_cstp_checkpoint(cstp_path) if isinstance(cstp_path, str) else None
    esm_sd, esm_cfg = read_esm2(esm_path or default_esm_path())
    model = build_from_state_dicts(llama_sd, llama_cfg, esm_sd, esm_cfg, cstp_sd, switch_sd,
                                   switch_type=switch_projector_type, lora_sd=lora_sd, lora_alpha=alpha, lora_r=r,
                                   eos_token_id=_eos_ids(model_base_path, extra["eos_token_id"]),
                                   device=device, family=family)
    context_len = extra["max_sequence_length"] or 512                                  # builder.py:126-131
    return tokenizer, model, context_len
