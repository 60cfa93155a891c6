"""Drop-in counterpart of the reference's `OpusLlamaForCausalLM` for the generation path.

Keeps the call surface the eval scripts use (multi_modality_v1/eval/run_opus_ddp.py:118-132 and siblings):
``model.generate(input_ids, seq, attention_mask=..., pad_token_id=..., do_sample=False, max_new_tokens=...)`` returning
``LongTensor[B, n_new]`` (new tokens only), plus the mixin methods of `OpusMetaModelForCauselLM`
(multi_modality_v1/model/opus_arch.py:103-294): encode_seq2embedding / encode_projector_embedding /
switch_projector_embedding / prepare_inputs_labels_for_multimodal. Error conventions follow the reference
(NotImplementedError for non-str `seq`, for `inputs_embeds=` and for beam search / top-k, which this backend does not
implement). Greedy and temperature / top-p sampling are both drawn on the device.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

from . import _lib as L
from . import ops
from .constants import DEFAULT_SEQ_TOKEN_INDEX, IGNORE_INDEX
from .encoder import B200ProteinEncoder
from .llama import B200Llama
from .projector import B200ProteinProjector, B200SwitchProjector, FusedProjectors, IdentityProjector

INT32_MIN = -(2 ** 31)


class SplicePlan:
    """Host-side plan of the soft-token splice (opus_arch.py:176-270): for every output embedding row either a vocab
    id (>= 0) or -(soft_row + 1); rows are packed per prompt (no pad rows)."""

    def __init__(self, input_ids: np.ndarray, mask: np.ndarray | None, n_soft: int, n_seq_avail: int,
                 max_length: int | None = None):
        B, _ = input_ids.shape
        rows, lens, seq_idx = [], [], 0
        for b in range(B):
            ids = input_ids[b] if mask is None else input_ids[b][mask[b]]
            sent = np.flatnonzero(ids == DEFAULT_SEQ_TOKEN_INDEX)
            if sent.size == 0:
                row = ids.astype(np.int64)
                seq_idx += 1  # a prompt without <seq> still consumes one protein slot (opus_arch.py:196-203)
            else:
                parts, prev = [], 0
                for s in sent:
                    parts.append(ids[prev:s])
                    if seq_idx >= n_seq_avail:
                        raise IndexError("more <seq> sentinels than protein sequences")
                    parts.append(-(seq_idx * n_soft + np.arange(n_soft, dtype=np.int64)) - 1)
                    seq_idx += 1
                    prev = s + 1
                parts.append(ids[prev:])
                row = np.concatenate(parts)
            if max_length is not None:
                row = row[:max_length]
            rows.append(row)
            lens.append(row.shape[0])
        self.lens = np.asarray(lens, dtype=np.int32)
        self.cu = np.zeros(B + 1, dtype=np.int32)
        np.cumsum(self.lens, out=self.cu[1:])
        self.src = (np.concatenate(rows) if rows else np.zeros(0)).astype(np.int32)
        self.n_seq_used = seq_idx

    def padded_src(self, left: bool) -> tuple[np.ndarray, np.ndarray]:
        """[B, Lmax] map with INT32_MIN on pad rows + bool mask (for the HF-style padded return)."""
        B, Lm = len(self.lens), int(self.lens.max()) if len(self.lens) else 0
        src = np.full((B, Lm), INT32_MIN, dtype=np.int32)
        m = np.zeros((B, Lm), dtype=bool)
        for b in range(B):
            n = int(self.lens[b])
            sl = slice(Lm - n, Lm) if left else slice(0, n)
            src[b, sl] = self.src[self.cu[b]: self.cu[b + 1]]
            m[b, sl] = True
        return src, m


class B200OpusLlama:
    """Encoder + projectors + Llama behind the reference's generate() contract."""

    def __init__(self, llama: B200Llama, encoder: B200ProteinEncoder | None, protein_projector, switch_projector,
                 n_soft_tokens: int = 8, eos_token_id=(), tokenizer_model_max_length: int | None = None):
        self.llama, self.protein_encoder = llama, encoder
        self.protein_projector = protein_projector if protein_projector is not None else IdentityProjector()
        self.switch_projector = switch_projector
        self.n_soft = n_soft_tokens
        self.device = llama.device
        eos = eos_token_id if isinstance(eos_token_id, (list, tuple)) else [eos_token_id]
        self.config = SimpleNamespace(hidden_size=llama.dim, vocab_size=llama.vocab, has_switch_projector=True,
                                      has_protein_encoder=encoder is not None, model_type="opus_llama",
                                      tokenizer_model_max_length=tokenizer_model_max_length,
                                      eos_token_id=[int(e) for e in eos if e is not None])
        self.generation_config = SimpleNamespace(eos_token_id=self.config.eos_token_id)
        self._fused = None

    # ---- nn.Module-ish conveniences the eval scripts touch
    def eval(self):
        return self

    def get_model(self):
        return self

    def get_protein_encoder(self):
        return self.protein_encoder

    def initialize_protein_modules(self, model_args, fsdp=None):
        """opus_arch.py:46-90 (called by load_pretrained_model, builder.py:118-119): (re)build the protein encoder, the CSTP
        projector and the switch projector from `model_args` -- `.esm_ckpt` (None = base ESM-2; `.esm_path` optionally
        names the base weights file), `.pretrain_protein_projector_ckpt` (None = identity), `.has_switch_projector`,
        `.switch_projector_type`, `.pretrain_switch_projector_ckpt` -- through the three builder seams."""
        from . import builder as B
        self.config.device = getattr(model_args, "device", self.device)
        self.config.has_protein_encoder = getattr(model_args, "has_protein_encoder", True)
        self.config.has_switch_projector = getattr(model_args, "has_switch_projector", False)
        if self.protein_encoder is None:
            self.protein_encoder = B.build_protein_encoder(getattr(model_args, "esm_ckpt", None),
                                                           esm_path=getattr(model_args, "esm_path", None),
                                                           device=self.device)
        else:
            self.protein_encoder.load_model()                                            # opus_arch.py:63
        ckpt = getattr(model_args, "pretrain_protein_projector_ckpt", None)
        self.protein_projector = (B.build_protein_projector(ckpt, device=self.device).to(self.device)
                                  if ckpt is not None else IdentityProjector())
        if self.config.has_switch_projector:
            if not hasattr(model_args, "hidden_size"):
                model_args.hidden_size = self.config.hidden_size
            self.switch_projector = B.build_switch_projector(model_args, self.n_soft, device=self.device)
            # the reference hard-codes 5120 / 1280 (protein_mlp/builder.py:14); take the width that is really there
            self.switch_projector.in_dim = (self.protein_projector.out_dim if ckpt is not None
                                            else getattr(self.protein_encoder, "dim", self.switch_projector.in_dim))
            sw_ckpt = getattr(model_args, "pretrain_switch_projector_ckpt", None)
            if sw_ckpt is not None:
                self.switch_projector.load_state_dict(B.read_switch_projector(sw_ckpt))
        self._fused = None

    def embed_tokens(self, ids: torch.Tensor) -> torch.Tensor:
        flat = ids.reshape(-1).to(self.device, torch.int32)
        return ops.embed_gather(flat, self.llama.embed).reshape(*ids.shape, self.llama.dim)

    # ---- OpusMetaModelForCauselLM mixin
    def encode_seq2embedding(self, seq):
        if type(seq) is not list:
            seq = [seq]
        if type(seq[0]) is not str:
            raise NotImplementedError
        return self.protein_encoder.get_protein_seq_embeddings(seq)

    def encode_projector_embedding(self, extractor_embedding):
        return self.protein_projector.protein_forward(extractor_embedding)

    def switch_projector_embedding(self, seq_embedding):
        out = self.switch_projector(seq_embedding)
        return out.reshape(out.shape[0], -1, self.config.hidden_size)

    def _soft_tokens(self, seq, seq_embedding):
        """-> bf16 [n_seq, n_soft, H]"""
        if seq_embedding is None:
            if type(seq) is not list:
                seq = [seq]
            if type(seq[0]) is not str:
                raise NotImplementedError
            if (self._fused is None and isinstance(self.switch_projector, B200SwitchProjector)
                    and isinstance(self.protein_projector, B200ProteinProjector)
                    and self.switch_projector.depth <= 2):
                self._fused = FusedProjectors(self.protein_projector, self.switch_projector)
            if self._fused is not None:
                _, pooled_l2, _, _ = self.protein_encoder.encode(seq)
                return self._fused(pooled_l2).reshape(len(seq), -1, self.config.hidden_size)
            emb = self.encode_seq2embedding(seq)
        else:
            emb = seq_embedding  # pre-computed ESM mean embeddings (opus_arch.py:151-161)
        emb = self.encode_projector_embedding(emb)
        emb = self.switch_projector_embedding(emb)
        if emb.ndim == 2:
            emb = emb.unsqueeze(1)
        elif emb.ndim != 3:
            raise NotImplementedError
        return emb

    def _plan(self, input_ids, attention_mask, n_seq):
        ids = ops.d2h(input_ids.detach()).numpy()
        mask = None if attention_mask is None else ops.d2h(attention_mask.detach().bool()).numpy()
        return SplicePlan(ids, mask, self.n_soft, n_seq, self.config.tokenizer_model_max_length)

    def prepare_inputs_labels_for_multimodal(self, input_ids, position_ids, attention_mask, past_key_values, labels,
                                             seq, seq_embedding=None, inference_mode=False):
        """opus_arch.py:133-294 contract: returns (None, position_ids, attention_mask, past_key_values, inputs_embeds,
        labels) with left padding when inference_mode else right padding."""
        if seq is None or self.protein_encoder is None or input_ids.shape[1] == 1:
            return input_ids, position_ids, attention_mask, past_key_values, None, labels
        soft = self._soft_tokens(seq, seq_embedding)
        plan = self._plan(input_ids, attention_mask, soft.shape[0])
        src, m = plan.padded_src(left=inference_mode)
        B, Lm = src.shape
        src_d = torch.from_numpy(src.reshape(-1)).to(self.device)
        embeds = ops.splice_gather(src_d, self.llama.embed, soft.reshape(-1, soft.shape[-1]).contiguous())
        embeds = embeds.reshape(B, Lm, -1)
        new_mask = torch.from_numpy(m).to(self.device)
        new_pos = None
        if position_ids is not None:
            p = np.zeros((B, Lm), dtype=np.int64)
            for b in range(B):
                n = int(plan.lens[b])
                p[b, (Lm - n if inference_mode else 0): (Lm if inference_mode else n)] = np.arange(n)
            new_pos = torch.from_numpy(p).to(self.device)
        new_labels = None
        if labels is not None:
            lab = np.full((B, Lm), IGNORE_INDEX, dtype=np.int64)
            lab_in = labels.detach().cpu().numpy()
            ids_in = input_ids.detach().cpu().numpy()
            mk = np.ones_like(ids_in, dtype=bool) if attention_mask is None else attention_mask.bool().cpu().numpy()
            for b in range(B):
                row_ids, row_lab = ids_in[b][mk[b]], lab_in[b][mk[b]]
                outl = []
                for t, lb in zip(row_ids, row_lab):
                    outl.extend([IGNORE_INDEX] * self.n_soft if t == DEFAULT_SEQ_TOKEN_INDEX else [lb])
                outl = outl[: int(plan.lens[b])]
                n = len(outl)
                lab[b, (Lm - n if inference_mode else 0): (Lm if inference_mode else n)] = outl
            new_labels = torch.from_numpy(lab).to(self.device)
        out_mask = None if attention_mask is None else new_mask.to(attention_mask.dtype)
        return None, new_pos, out_mask, past_key_values, embeds, new_labels

    # ---- teacher-forced scoring
    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, labels=None,
                use_cache=None, output_attentions=None, output_hidden_states=None, seq=None, input_embed=None,
                return_dict=None, return_logits: bool = True, **kwargs):
        """opus_llama.py:41-93 contract for the scoring use (`model(input_ids, labels=..., seq=...)`): splice the soft
        tokens with RIGHT padding semantics (opus_arch.py:259-269), run the prompt, and return an object with `.loss`
        (HF: mean fp32 cross entropy of position t predicting label t+1 over labels != -100) and `.logits`
        ([B, L', V] bf16, right-padded with zeros; None when return_logits=False to save the 2*V bytes per token).
        Inference only: there is no autograd graph behind the loss."""
        if kwargs.get("inputs_embeds") is not None or past_key_values is not None:
            raise NotImplementedError("forward(): `inputs_embeds` / `past_key_values` are not supported")
        if output_attentions or output_hidden_states:
            raise NotImplementedError("forward(): attention maps / hidden states are not materialised")
        n_prompts = input_ids.shape[0]
        if seq is not None and self.protein_encoder is not None and input_ids.shape[1] != 1:
            soft = self._soft_tokens(seq, input_embed)
            plan = self._plan(input_ids, attention_mask, soft.shape[0])
            soft2d = soft.reshape(-1, soft.shape[-1]).contiguous()
        else:
            ids = input_ids.detach().cpu().numpy()
            mask = None if attention_mask is None else attention_mask.detach().bool().cpu().numpy()
            plan = SplicePlan(np.where(ids == DEFAULT_SEQ_TOKEN_INDEX, 0, ids), mask, self.n_soft, 1 << 30)
            soft2d = torch.zeros((1, self.llama.dim), dtype=torch.bfloat16, device=self.device)
        # labels follow the splice (IGNORE_INDEX under the soft tokens), then shift by one inside every sequence
        n_tok = int(plan.cu[-1])
        tgt = np.full((n_tok,), IGNORE_INDEX, dtype=np.int32)
        if labels is not None:
            lab_in = labels.detach().cpu().numpy()
            ids_in = input_ids.detach().cpu().numpy()
            mk = np.ones_like(ids_in, dtype=bool) if attention_mask is None else attention_mask.detach().bool().cpu().numpy()
            for b in range(n_prompts):
                row_ids, row_lab = ids_in[b][mk[b]], lab_in[b][mk[b]]
                reps = np.where(row_ids == DEFAULT_SEQ_TOKEN_INDEX, self.n_soft, 1)
                spliced = np.repeat(np.where(row_ids == DEFAULT_SEQ_TOKEN_INDEX, IGNORE_INDEX, row_lab), reps)
                n = int(plan.lens[b])
                spliced = spliced[:n]
                tgt[plan.cu[b]: plan.cu[b] + n - 1] = spliced[1:n]
        embeds = ops.splice_gather(ops.h2d(plan.src, self.device), self.llama.embed, soft2d)
        losses, logits = self.llama.score_packed(embeds, plan.cu, torch.from_numpy(tgt), return_logits=return_logits)
        counted = int((tgt >= 0).sum())
        loss = (losses.sum() / counted) if (labels is not None and counted > 0) else None
        padded = None
        if logits is not None:
            Lm = int(plan.lens.max())
            padded = torch.zeros((n_prompts, Lm, logits.shape[-1]), dtype=logits.dtype, device=self.device)
            for b in range(n_prompts):
                padded[b, : int(plan.lens[b])] = logits[plan.cu[b]: plan.cu[b + 1]]
        return SimpleNamespace(loss=loss, logits=padded, token_losses=losses, past_key_values=None)

    __call__ = forward

    # ---- generation
    @torch.no_grad()
    def generate(self, inputs=None, seq=None, seq_embedding=None, **kwargs):
        """opus_llama.py:95-132 contract. do_sample=False: greedy (north star). do_sample=True: temperature / top-p
        sampling on the device (the reference's default eval setting, run_opus_ddp.py:126-128: temperature 0.1,
        top_p 0.7); `seed=` pins the draw, otherwise torch.initial_seed() and a per-call counter are used."""
        kwargs.pop("position_ids", None)
        attention_mask = kwargs.pop("attention_mask", None)
        if "inputs_embeds" in kwargs:
            raise NotImplementedError("`inputs_embeds` is not supported")
        do_sample = bool(kwargs.pop("do_sample", False))
        if int(kwargs.pop("num_beams", 1) or 1) != 1:
            raise NotImplementedError("beam search is not implemented")
        temperature = kwargs.pop("temperature", None)
        top_p = kwargs.pop("top_p", None)
        seed = kwargs.pop("seed", None)
        if kwargs.pop("top_k", None) not in (None, 0):
            raise NotImplementedError("top_k sampling is not implemented")
        kwargs.pop("use_cache", None)
        sampling = None
        if do_sample:
            temperature = 1.0 if temperature is None else float(temperature)
            top_p = 1.0 if top_p is None else float(top_p)
            if not temperature > 0.0:
                raise ValueError("`temperature` has to be a strictly positive float")   # HF's message
            if not 0.0 < top_p <= 1.0:
                raise ValueError("`top_p` has to be a float > 0 and <= 1")
            if seed is None:
                self._sample_calls = getattr(self, "_sample_calls", 0) + 1
                seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + self._sample_calls) & (2 ** 64 - 1)
            sampling = (temperature, top_p, int(seed) & (2 ** 64 - 1))
        max_new = int(kwargs.pop("max_new_tokens", 20))
        pad_id = kwargs.pop("pad_token_id", None)
        eos = kwargs.pop("eos_token_id", None)
        eos_ids = self.config.eos_token_id if eos is None else ([eos] if isinstance(eos, int) else list(eos))
        if pad_id is None:
            pad_id = eos_ids[0] if eos_ids else 0
        use_graph = bool(kwargs.pop("use_graph", True))
        return_logits = bool(kwargs.pop("return_prefill_logits", False))
        # stop keywords: `stop_sequences=[[ids of "###"], ...]` or the reference's KeywordsStoppingCriteria objects
        # (mm_utils.py:43-75) in `stopping_criteria=[...]`; rows stop individually, on the device
        stops = [list(map(int, q)) for q in (kwargs.pop("stop_sequences", None) or [])]
        for crit in kwargs.pop("stopping_criteria", None) or []:
            if not hasattr(crit, "keyword_ids"):
                raise NotImplementedError("only KeywordsStoppingCriteria-style criteria (keyword_ids) are supported")
            stops.extend([int(t) for t in torch.as_tensor(q).reshape(-1).tolist()] for q in crit.keyword_ids)
        if kwargs:
            raise TypeError(f"generate(): unsupported arguments {sorted(kwargs)}")

        if seq is not None and self.protein_encoder is not None and inputs.shape[1] != 1:
            soft = self._soft_tokens(seq, seq_embedding)
            plan = self._plan(inputs, attention_mask, soft.shape[0])
            soft2d = soft.reshape(-1, soft.shape[-1]).contiguous()
        else:
            ids = inputs.detach().cpu().numpy()
            mask = None if attention_mask is None else attention_mask.detach().bool().cpu().numpy()
            plan = SplicePlan(np.where(ids == DEFAULT_SEQ_TOKEN_INDEX, 0, ids), mask, self.n_soft, 1 << 30)
            soft2d = torch.zeros((1, self.llama.dim), dtype=torch.bfloat16, device=self.device)
        src_d = ops.h2d(plan.src, self.device)
        embeds = ops.splice_gather(src_d, self.llama.embed, soft2d)
        return self.llama.generate_packed(embeds, plan.cu, max_new, eos_ids=eos_ids, pad_id=int(pad_id),
                                          use_graph=use_graph, return_prefill_logits=return_logits, sampling=sampling,
                                          stop_sequences=stops)


def build_from_state_dicts(llama_sd: dict, llama_cfg: dict, esm_sd: dict | None, esm_cfg: dict | None,
                           projector_sd: dict | None, switch_sd: dict | None, switch_type: str = "mlp2x_gelu",
                           lora_sd: dict | None = None, lora_alpha: float = 32.0, lora_r: int = 16,
                           eos_token_id=(), device="cuda", n_soft_tokens: int = 8,
                           family: str = "llama") -> B200OpusLlama:
    """Assemble the model from plain state dicts (HF / fair-esm / Lightning / .bin key names). family = "llama" (Llama,
    Qwen2: opus_llama.py / opus_qwen.py) or "opt" (OPT, Galactica: opus_opt.py; llama_cfg then holds B200Opt's kwargs)."""
    if family == "opt":
        from .opt import B200Opt
        llama = B200Opt(llama_sd, device=device, lora=lora_sd, lora_alpha=lora_alpha, lora_r=lora_r, **llama_cfg)
    elif family == "llama":
        llama = B200Llama(llama_sd, device=device, lora=lora_sd, lora_alpha=lora_alpha, lora_r=lora_r, **llama_cfg)
    else:
        raise NotImplementedError(f"unknown decoder family {family!r}")
    enc = B200ProteinEncoder(esm_sd, device=device, **esm_cfg) if esm_sd is not None else None
    pp = None
    in_dim = esm_cfg["dim"] if esm_cfg else 1280
    if projector_sd is not None:
        pp = B200ProteinProjector(projector_sd["protein_projection.linear.weight"],
                                  projector_sd["protein_projection.linear.bias"], device=device)
        in_dim = pp.out_dim
    sw = B200SwitchProjector(in_dim, llama.dim * n_soft_tokens, switch_type, device=device)
    if switch_sd is not None:
        sw.load_state_dict(switch_sd)
    model = B200OpusLlama(llama, enc, pp, sw, n_soft_tokens, eos_token_id)
    if family == "opt":
        model.config.model_type = "opus_opt"
    return model
