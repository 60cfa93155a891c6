"""GPU: the CUDA path against the COMMITTED golden fixtures (tests/golden/*.pt) that oracle/make_golden.py produced
from the reference's own classes (/root/reference) and HF modules in the build container."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from opus_pllm_b200 import synth  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def _cos(a, b):
    a, b = a.float().flatten().cpu(), b.float().flatten().cpu()
    return float(torch.dot(a, b) / (a.norm() * b.norm()))


def test_encoder_matches_hf_esm_golden():
    from opus_pllm_b200.encoder import B200ProteinEncoder
    g = _load("esm2_small.pt")
    c = g["cfg"]
    enc = B200ProteinEncoder(synth.esm2_weights(c["n_layers"], c["dim"], c["ffn"], seed=g["seed"]), c["n_layers"],
                             c["dim"], c["n_heads"], c["ffn"])
    pooled, _, hidden, pk = enc.encode(g["seqs"], want_hidden=True)
    assert pooled.shape == g["pooled"].shape
    assert _cos(pooled, g["pooled"]) >= 0.9995 and float((pooled.cpu() - g["pooled"]).abs().max()) <= 3e-2
    # per-residue states (representations[33]) too: un-pad the golden [B, T, D] into the packed layout
    valid = g["tokens"] != 1
    assert _cos(hidden, g["hidden"][valid]) >= 0.9995


def test_projectors_and_splice_match_reference_golden():
    from opus_pllm_b200.model import build_from_state_dicts
    g, lg = _load("mm_small.pt"), _load("llama_small.pt")
    c = lg["cfg"]
    H = c["dim"]
    pw = synth.projector_weights(g["esm_cfg"]["dim"], 5120, 8 * H, seed=g["proj_seed"])
    lw = synth.llama_weights(c["n_layers"], H, c["n_q_heads"], c["n_kv_heads"], c["head_dim"], c["ffn_dim"], c["vocab"],
                             seed=lg["seed"])
    ec = g["esm_cfg"]
    model = build_from_state_dicts(lw, c, synth.esm2_weights(ec["n_layers"], ec["dim"], ec["ffn"], seed=11),
                                   dict(n_layers=ec["n_layers"], dim=ec["dim"], n_heads=ec["n_heads"], ffn_dim=ec["ffn"]),
                                   pw, pw)
    cstp = model.encode_projector_embedding(g["pooled"].cuda())
    assert _cos(cstp, g["cstp_out"]) >= 0.9995
    soft = model.switch_projector_embedding(cstp)
    assert soft.shape == g["soft"].shape and _cos(soft, g["soft"]) >= 0.999
    # whole reference-shaped call: proteins -> soft tokens -> spliced, left-padded embeddings
    out = model.prepare_inputs_labels_for_multimodal(g["input_ids"].cuda(), None, g["attention_mask"].cuda(), None, None,
                                                     g["seqs"], inference_mode=True)
    assert torch.equal(out[2].bool().cpu(), g["mask_left"])
    assert _cos(out[4], g["embeds_left"]) >= 0.999
    text_rows = (g["embeds_left"].abs().sum(-1) > 0) & g["mask_left"]
    # text rows are pure gathers of the bf16 embedding table: exact
    is_text = torch.zeros_like(g["mask_left"])
    for b in range(g["input_ids"].shape[0]):
        ids = g["input_ids"][b][g["attention_mask"][b]]
        flags = []
        for t in ids.tolist():
            flags.extend([False] * 8 if t == -200 else [True])
        n = len(flags)
        is_text[b, is_text.shape[1] - n:] = torch.tensor(flags)
    want_text = g["embeds_left"][is_text].to(torch.bfloat16)
    assert torch.equal(out[4].cpu()[is_text], want_text) and bool(text_rows.any())


def test_llama_prefill_matches_reference_generate_golden():
    from opus_pllm_b200.llama import B200Llama
    g = _load("llama_small.pt")
    c = g["cfg"]
    lw = synth.llama_weights(c["n_layers"], c["dim"], c["n_q_heads"], c["n_kv_heads"], c["head_dim"], c["ffn_dim"],
                             c["vocab"], seed=g["seed"])
    model = B200Llama(lw, **c)
    mask = g["mask"]
    lens = mask.sum(1).tolist()
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    packed = g["embeds"][mask].cuda().to(torch.bfloat16)
    out, logits = model.generate_packed(packed, cu, g["max_new_tokens"], eos_ids=[g["eos"]], pad_id=g["pad"],
                                        return_prefill_logits=True)
    want = g["prefill_logits"]
    assert _cos(logits, want) >= 0.999
    assert float((logits.float().cpu() - want).abs().max()) <= 0.06 * float(want.std()) + 1e-3
    # first generated token: identical unless the reference's own top-2 margin is within bf16 noise
    top2 = want.topk(2, dim=-1).values
    margin = top2[:, 0] - top2[:, 1]
    same = out[:, 0].cpu() == g["tokens"][:, 0]
    assert bool((same | (margin < 0.05 * float(want.std()))).all())


def test_forward_labels_scoring_matches_reference_golden():
    """model(input_ids, labels=..., seq=..., input_embed=...) — the teacher-forced scoring path (opus_llama.py:41-93,
    right-padded splice opus_arch.py:259-269, pre-computed ESM embeddings opus_arch.py:151-161) against the loss and
    all-position logits the reference's own forward produced (tests/golden/score_small.pt)."""
    from opus_pllm_b200.model import build_from_state_dicts
    g, mm = _load("score_small.pt"), _load("mm_small.pt")
    c = g["cfg"]
    H = c["dim"]
    pw = synth.projector_weights(mm["esm_cfg"]["dim"], 5120, 8 * H, seed=mm["proj_seed"])
    lw = synth.llama_weights(c["n_layers"], H, c["n_q_heads"], c["n_kv_heads"], c["head_dim"], c["ffn_dim"], c["vocab"],
                             seed=g["seed"])
    ec = mm["esm_cfg"]
    model = build_from_state_dicts(lw, c, synth.esm2_weights(ec["n_layers"], ec["dim"], ec["ffn"], seed=11),
                                   dict(n_layers=ec["n_layers"], dim=ec["dim"], n_heads=ec["n_heads"], ffn_dim=ec["ffn"]),
                                   pw, pw)
    ids, mask, labels = g["input_ids"].cuda(), g["attention_mask"].cuda(), g["labels"].cuda()
    out = model(ids, attention_mask=mask, labels=labels, seq=g["seqs"], input_embed=mm["pooled"].cuda())
    want_logits, valid = g["logits"].float(), g["mask_right"]
    assert out.logits.shape == want_logits.shape
    assert _cos(out.logits.cpu()[valid], want_logits[valid]) >= 0.999
    assert abs(float(out.loss) - float(g["loss"])) <= 0.03, (float(out.loss), float(g["loss"]))
    assert bool((out.logits.cpu()[~valid] == 0).all())
    # same through the protein encoder instead of pre-computed embeddings, and without materialising the logits
    out2 = model(ids, attention_mask=mask, labels=labels, seq=g["seqs"], return_logits=False)
    assert out2.logits is None and abs(float(out2.loss) - float(g["loss"])) <= 0.05
    # no labels -> no loss (HF returns loss=None)
    assert model(ids, attention_mask=mask, seq=g["seqs"], return_logits=False).loss is None


def test_qwen_family_matches_reference_wrapper_golden():
    """Qwen2 / Qwen2.5 (model/language_model/opus_qwen.py) through the same kernels: q/k/v projection biases (prefill
    epilogue, decode split-K reduce), GQA group 3, rope theta 1e6, eps 1e-6 — against the reference wrapper's own
    generate output (tests/golden/qwen_small.pt)."""
    from opus_pllm_b200.llama import B200Llama
    g = _load("qwen_small.pt")
    c = g["cfg"]
    lw = synth.llama_weights(c["n_layers"], c["dim"], c["n_q_heads"], c["n_kv_heads"], c["head_dim"], c["ffn_dim"],
                             c["vocab"], seed=g["seed"], qkv_bias=True)
    model = B200Llama(lw, **c, rms_eps=g["rms_eps"], rope_theta=g["rope_theta"])
    mask = g["mask"]
    lens = mask.sum(1).tolist()
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    packed = lw["model.embed_tokens.weight"][g["input_ids"]][mask].cuda().to(torch.bfloat16)
    out, logits = model.generate_packed(packed, cu, g["max_new_tokens"], eos_ids=[g["eos"]], pad_id=g["pad"],
                                        return_prefill_logits=True)
    want = g["prefill_logits"]
    assert _cos(logits, want) >= 0.999
    assert float((logits.float().cpu() - want).abs().max()) <= 0.06 * float(want.std()) + 1e-3
    top2 = want.topk(2, dim=-1).values
    same = out[:, 0].cpu() == g["tokens"][:, 0]
    assert bool((same | ((top2[:, 0] - top2[:, 1]) < 0.05 * float(want.std()))).all())
    # without the biases the logits must move: the bias path is really exercised
    nb = {k: v for k, v in lw.items() if not k.endswith(".bias")}
    _, logits_nb = B200Llama(nb, **c, rms_eps=g["rms_eps"], rope_theta=g["rope_theta"]).generate_packed(
        packed, cu, 2, return_prefill_logits=True)
    assert _cos(logits_nb, want) < 0.999
