"""ORACLE tooling (test infrastructure only): pin the restatements in oracle/ against the real thing and write the
golden fixtures under tests/golden/.  Run in the BUILD container (needs /root/reference; the GPU box never runs it):

    python -m oracle.make_golden

What is executed here, on CPU fp32, tiny shapes, weights from opus_pllm_b200.synth (hash-seeded, reproducible):
  1. HuggingFace `EsmModel` (transformers, port of fair-esm ESM-2; fair-esm itself is not installable offline)
       -> tests/golden/esm2_small.pt       pins oracle.esm2_ref
  2. the REFERENCE's own classes imported from /root/reference: `CSTPBase.protein_forward`
     (cstp_v3/modelling.py:396-400), `build_switch_projector` (protein_mlp/builder.py:11-25) and
     `OpusLlamaForCausalLM.prepare_inputs_labels_for_multimodal` (opus_arch.py:133-294)
       -> tests/golden/mm_small.pt         pins oracle.mm_ref
  3. the REFERENCE's `OpusLlamaForCausalLM.generate` (opus_llama.py:95-132) over transformers' LlamaForCausalLM, greedy,
     with a duck-typed protein encoder that returns the oracle's ESM embeddings
       -> tests/golden/llama_small.pt      pins oracle.llama_ref (prefill logits + generated tokens)
Two stub modules (`esm`, `pytorch_lightning`) satisfy the reference's imports and one 3-line subclass drops the
`cache_position` pop that only exists for transformers 4.46 (opus_llama.py:141); nothing else is modified.
Every fixture stores the inputs, the expected outputs and the max deviation of the oracle restatement at creation time.
"""
from __future__ import annotations

import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")

from opus_pllm_b200 import synth  # noqa: E402
from oracle import esm2_ref, llama_ref, mm_ref  # noqa: E402


def _install_stubs():
    import torch.nn as nn
    esm = types.ModuleType("esm")
    esm.pretrained = types.SimpleNamespace(esm2_t33_650M_UR50D=lambda: (_ for _ in ()).throw(RuntimeError("stub")))
    sys.modules["esm"] = esm
    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = nn.Module
    util = types.ModuleType("pytorch_lightning.utilities")
    util.rank_zero_info = print
    pl.utilities = util
    sys.modules["pytorch_lightning"] = pl
    sys.modules["pytorch_lightning.utilities"] = util
    if REF not in sys.path:
        sys.path.insert(0, REF)


ESM_SMALL = dict(n_layers=2, dim=128, n_heads=2, ffn=256)
LLAMA_SMALL = dict(n_layers=2, dim=256, n_q_heads=2, n_kv_heads=1, head_dim=128, ffn_dim=512, vocab=1024)


def golden_esm():
    from transformers import EsmConfig, EsmModel
    c = ESM_SMALL
    w = synth.esm2_weights(c["n_layers"], c["dim"], c["ffn"], seed=11)
    cfg = EsmConfig(vocab_size=33, hidden_size=c["dim"], num_hidden_layers=c["n_layers"],
                    num_attention_heads=c["n_heads"], intermediate_size=c["ffn"], position_embedding_type="rotary",
                    token_dropout=True, emb_layer_norm_before=False, mask_token_id=32, pad_token_id=1,
                    layer_norm_eps=1e-5, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    hf = EsmModel(cfg, add_pooling_layer=False).eval()
    missing, unexpected = hf.load_state_dict(esm2_ref.to_hf_esm_state_dict(w, c["n_layers"]), strict=False)
    assert not [m for m in missing if "inv_freq" not in m and "position_ids" not in m and "contact_head" not in m], missing
    assert not unexpected, unexpected
    seqs = synth.proteins(5, 3, 60, seed=21) + ["MKT", "A"]
    tokens = esm2_ref.tokenize(seqs)
    with torch.no_grad():
        hidden = hf(input_ids=tokens, attention_mask=(tokens != 1)).last_hidden_state
    lens = (tokens != 1).sum(1)
    pooled = torch.stack([hidden[i, 1: int(lens[i]) - 1].mean(0) for i in range(len(seqs))]).float()  # modelling.py:53-55
    mine = esm2_ref.get_protein_seq_embeddings(w, seqs, c["n_layers"], c["n_heads"])
    ok = ~torch.isnan(pooled).any(1)   # "A" has 1 residue -> fine; empty slices would be NaN in the reference too
    dev = float((mine[ok] - pooled[ok]).abs().max())
    assert dev < 2e-5, dev
    torch.save(dict(cfg=c, seed=11, seqs=seqs, tokens=tokens, pooled=pooled, hidden=hidden, oracle_dev=dev,
                    source="transformers.EsmModel " + __import__("transformers").__version__),
               os.path.join(GOLD, "esm2_small.pt"))
    print(f"esm2_small.pt: oracle vs HF EsmModel max |diff| = {dev:.3g}")
    return w, seqs, pooled


def golden_mm_and_llama(esm_w, seqs, pooled):
    _install_stubs()
    from transformers import LlamaConfig
    from multi_modality_model.cstp_v3.modelling import CSTPBase
    from multi_modality_model.multi_modality_v1.model.language_model.opus_llama import (OpusLlamaConfig,
                                                                                        OpusLlamaForCausalLM)
    from multi_modality_model.multi_modality_v1.model.protein_mlp.builder import build_switch_projector

    c = LLAMA_SMALL
    H = c["dim"]
    pw = synth.projector_weights(ESM_SMALL["dim"], 5120, 8 * H, seed=12)

    # ---- reference projector classes with our weights
    cstp = CSTPBase(ESM_SMALL["dim"], 5120, 5120, 5120, 8, 1, 0.5).eval()
    cstp.protein_projection.linear.weight.data.copy_(pw["protein_projection.linear.weight"])
    cstp.protein_projection.linear.bias.data.copy_(pw["protein_projection.linear.bias"])
    margs = types.SimpleNamespace(hidden_size=H, pretrain_protein_projector_ckpt="x", switch_projector_type="mlp2x_gelu")
    switch = build_switch_projector(margs).eval()
    switch.load_state_dict({k: pw[k] for k in ("0.weight", "0.bias", "2.weight", "2.bias")})

    # ---- reference LLM wrapper on a tiny Llama config with our weights
    class Shim(OpusLlamaForCausalLM):  # transformers >= 4.47 no longer returns `cache_position` (opus_llama.py:141)
        def prepare_inputs_for_generation(self, input_ids, past_key_values=None, inputs_embeds=None, **kwargs):
            kwargs.pop("seq", None)
            return super(OpusLlamaForCausalLM, self).prepare_inputs_for_generation(
                input_ids, past_key_values=past_key_values, inputs_embeds=inputs_embeds, **kwargs)

    hf_cfg = OpusLlamaConfig(vocab_size=c["vocab"], hidden_size=H, intermediate_size=c["ffn_dim"],
                             num_hidden_layers=c["n_layers"], num_attention_heads=c["n_q_heads"],
                             num_key_value_heads=c["n_kv_heads"], head_dim=c["head_dim"], rms_norm_eps=1e-5,
                             rope_theta=500000.0, max_position_embeddings=512, tie_word_embeddings=False,
                             attention_bias=False, mlp_bias=False, bos_token_id=1, eos_token_id=2, pad_token_id=None)
    hf_cfg._attn_implementation = "eager"
    lw = synth.llama_weights(c["n_layers"], H, c["n_q_heads"], c["n_kv_heads"], c["head_dim"], c["ffn_dim"], c["vocab"],
                             seed=13)
    model = Shim(hf_cfg).eval()
    missing, unexpected = model.load_state_dict(lw, strict=False)
    assert not [m for m in missing if "rotary" not in m and "inv_freq" not in m], missing
    assert not unexpected, unexpected

    class FakeEncoder:  # duck-typed `protein_encoder` (opus_arch.py:35,111): returns the pinned ESM embeddings
        def get_protein_seq_embeddings(self, data):
            assert list(data) == list(seqs)
            return pooled.clone()

    model.model.protein_encoder = FakeEncoder()
    model.model.protein_projector = cstp
    model.model.switch_projector = switch
    model.config.has_switch_projector = True
    model.config.has_protein_encoder = True

    B = len(seqs)
    prompts = synth.prompt_ids(B, 20, vocab=c["vocab"], seed=31, bos=1, sentinel_at=5, ragged=4)
    prompts[2] = prompts[2][prompts[2] != -200]            # a row without <seq> still consumes a protein slot
    pad = 0
    Lm = max(p.numel() for p in prompts)
    ids = torch.stack([torch.cat([torch.full((Lm - p.numel(),), pad), p]) for p in prompts])
    mask = ids != pad
    ok_rows = ~torch.isnan(pooled).any(1)
    assert bool(ok_rows.all())

    with torch.no_grad():
        c_ref = model.encode_projector_embedding(pooled)
        s_ref = model.switch_projector_embedding(c_ref)
        _, _, m_ref, _, e_ref, _ = model.prepare_inputs_labels_for_multimodal(ids, None, mask, None, None, list(seqs),
                                                                               inference_mode=True)
        _, p_r, m_r, _, e_r, _ = model.prepare_inputs_labels_for_multimodal(
            ids, torch.arange(Lm), mask, None, None, list(seqs), inference_mode=False)
        out = model.generate(ids, list(seqs), attention_mask=mask, pad_token_id=2, do_sample=False, max_new_tokens=12,
                             use_cache=True)
        # prefill logits of the last position straight from the reference forward
        logits = model(inputs_embeds=e_ref, attention_mask=m_ref).logits[:, -1, :].float()

        # teacher-forced scoring through the reference forward (opus_llama.py:41-93; right padding, labels spliced)
        labels = ids.clone()
        labels[~mask] = -100
        labels[:, :Lm - 12] = -100                          # only the last dozen positions are scored
        labels[ids == -200] = -100
        scored = model(input_ids=ids, attention_mask=mask, labels=labels, seq=list(seqs))
        score_loss, score_logits = scored.loss.float(), scored.logits.float()

    # ---- oracle restatements on the same inputs
    c_mine = mm_ref.protein_forward(pooled, pw["protein_projection.linear.weight"], pw["protein_projection.linear.bias"])
    s_mine = mm_ref.switch_projector(c_mine, pw, H)
    e_mine, m_mine, _, lens = mm_ref.splice(ids, mask, s_mine, lw["model.embed_tokens.weight"], True)
    e_mine_r, m_mine_r, p_mine_r, _ = mm_ref.splice(ids, mask, s_mine, lw["model.embed_tokens.weight"], False)
    d_c = float((c_mine - c_ref.float()).abs().max())
    d_s = float((s_mine - s_ref.float()).abs().max())
    d_e = float((e_mine - e_ref.float()).abs().max())
    assert d_c < 1e-5 and d_s < 1e-4 and d_e < 1e-4, (d_c, d_s, d_e)
    assert torch.equal(m_mine, m_ref.bool()) and torch.equal(m_mine_r, m_r.bool()) and torch.equal(p_mine_r, p_r)
    assert float((e_mine_r - e_r.float()).abs().max()) < 1e-4
    ocfg = llama_ref.LlamaCfg(n_layers=c["n_layers"], dim=H, n_q_heads=c["n_q_heads"], n_kv_heads=c["n_kv_heads"],
                              head_dim=c["head_dim"], ffn_dim=c["ffn_dim"], vocab=c["vocab"])
    pos = (m_mine.long().cumsum(-1) - 1).masked_fill(~m_mine, 1)
    lg_mine, _ = llama_ref.llama_forward(lw, ocfg, e_ref.float(), m_mine, pos)
    d_l = float((lg_mine - logits).abs().max())
    out_mine = llama_ref.greedy_generate(lw, ocfg, e_ref.float(), m_mine, 12, eos_ids=(2,), pad_id=2)
    assert d_l < 2e-4, d_l
    assert out_mine.shape == out.shape and torch.equal(out_mine, out), (out_mine, out)
    # scoring path: right-padded splice + all-position logits + shifted cross entropy
    lab_mine = mm_ref.splice_labels(ids, mask, labels)
    pos_r = (m_mine_r.long().cumsum(-1) - 1).masked_fill(~m_mine_r, 1)
    lg_all, _ = llama_ref.llama_forward(lw, ocfg, e_mine_r, m_mine_r, pos_r, all_positions=True)
    loss_mine = llama_ref.causal_lm_loss(lg_all, lab_mine)
    d_loss = float((loss_mine - score_loss).abs())
    d_lg = float(((lg_all - score_logits) * m_mine_r[..., None]).abs().max())
    assert d_loss < 1e-4 and d_lg < 5e-4, (d_loss, d_lg)
    torch.save(dict(cfg=c, seed=13, input_ids=ids, attention_mask=mask, labels=labels, seqs=list(seqs), loss=score_loss,
                    logits=score_logits.to(torch.bfloat16), mask_right=m_mine_r, labels_spliced=lab_mine,
                    oracle_dev=dict(loss=d_loss, logits=d_lg),
                    source="reference OpusLlamaForCausalLM.forward(labels=...) over transformers " +
                           __import__("transformers").__version__),
               os.path.join(GOLD, "score_small.pt"))
    print(f"score_small.pt: oracle vs reference forward(labels): loss |diff| {d_loss:.3g}, logits max |diff| {d_lg:.3g}")
    torch.save(dict(esm_cfg=ESM_SMALL, proj_seed=12, seqs=list(seqs), pooled=pooled, cstp_out=c_ref.float(),
                    soft=s_ref.float(), input_ids=ids, attention_mask=mask, embeds_left=e_ref.float(),
                    mask_left=m_ref.bool(), embeds_right=e_r.float(), mask_right=m_r.bool(), pos_right=p_r, lens=lens,
                    oracle_dev=dict(cstp=d_c, switch=d_s, splice=d_e), source="reference classes from /root/reference"),
               os.path.join(GOLD, "mm_small.pt"))
    torch.save(dict(cfg=c, seed=13, embeds=e_ref.float(), mask=m_ref.bool(), prefill_logits=logits, tokens=out,
                    eos=2, pad=2, max_new_tokens=12, oracle_dev=dict(logits=d_l, tokens=0),
                    source="reference OpusLlamaForCausalLM.generate over transformers " +
                           __import__("transformers").__version__),
               os.path.join(GOLD, "llama_small.pt"))
    print(f"mm_small.pt: oracle vs reference classes max |diff| cstp {d_c:.3g} switch {d_s:.3g} splice {d_e:.3g}")
    print(f"llama_small.pt: oracle vs reference generate: logits max |diff| {d_l:.3g}, tokens identical {tuple(out.shape)}")


def golden_qwen():
    """Sibling family (model/language_model/opus_qwen.py): the reference's OpusQwenForCausalLM (Qwen2: Llama blocks +
    q/k/v biases, GQA group 3 here, rope theta 1e6, eps 1e-6) on a tiny config with hash-seeded weights; text-only
    greedy generate (`seq=None` branch, opus_qwen.py:124-125) and last-position prefill logits."""
    from multi_modality_model.multi_modality_v1.model.language_model.opus_qwen import OpusQwenConfig, OpusQwenForCausalLM

    class Shim(OpusQwenForCausalLM):   # same transformers-version shim as for the Llama wrapper
        def prepare_inputs_for_generation(self, input_ids, past_key_values=None, inputs_embeds=None, **kwargs):
            kwargs.pop("seq", None)
            return super(OpusQwenForCausalLM, self).prepare_inputs_for_generation(
                input_ids, past_key_values=past_key_values, inputs_embeds=inputs_embeds, **kwargs)

    c = dict(n_layers=2, dim=384, n_q_heads=3, n_kv_heads=1, head_dim=128, ffn_dim=768, vocab=1024)
    hf_cfg = OpusQwenConfig(vocab_size=c["vocab"], hidden_size=c["dim"], intermediate_size=c["ffn_dim"],
                            num_hidden_layers=c["n_layers"], num_attention_heads=c["n_q_heads"],
                            num_key_value_heads=c["n_kv_heads"], rms_norm_eps=1e-6, rope_theta=1000000.0,
                            max_position_embeddings=512, tie_word_embeddings=False, use_sliding_window=False,
                            bos_token_id=1, eos_token_id=2, pad_token_id=None)
    hf_cfg._attn_implementation = "eager"
    lw = synth.llama_weights(c["n_layers"], c["dim"], c["n_q_heads"], c["n_kv_heads"], c["head_dim"], c["ffn_dim"],
                             c["vocab"], seed=23, qkv_bias=True)
    model = Shim(hf_cfg).eval()
    missing, unexpected = model.load_state_dict(lw, strict=False)
    assert not [m for m in missing if "rotary" not in m and "inv_freq" not in m and "protein" not in m and "switch" not in m], missing
    assert not unexpected, unexpected
    B, L = 5, 19
    g = torch.Generator().manual_seed(41)
    ids = torch.randint(3, c["vocab"], (B, L), generator=g)
    pad = 0
    for b in range(B):                                       # left padding, ragged
        ids[b, : b * 2] = pad
    mask = ids != pad
    with torch.no_grad():
        out = model.generate(ids, None, attention_mask=mask, pad_token_id=2, do_sample=False, max_new_tokens=10,
                             use_cache=True)
        emb = model.get_model().embed_tokens(ids)
        logits = model(inputs_embeds=emb, attention_mask=mask).logits[:, -1, :].float()
    ocfg = llama_ref.LlamaCfg(n_layers=c["n_layers"], dim=c["dim"], n_q_heads=c["n_q_heads"], n_kv_heads=c["n_kv_heads"],
                              head_dim=c["head_dim"], ffn_dim=c["ffn_dim"], vocab=c["vocab"], rms_eps=1e-6,
                              rope_theta=1000000.0)
    pos = (mask.long().cumsum(-1) - 1).masked_fill(~mask, 1)
    lg_mine, _ = llama_ref.llama_forward(lw, ocfg, emb.float(), mask, pos)
    out_mine = llama_ref.greedy_generate(lw, ocfg, emb.float(), mask, 10, eos_ids=(2,), pad_id=2)
    d_l = float((lg_mine - logits).abs().max())
    assert d_l < 2e-4, d_l
    assert out_mine.shape == out.shape and torch.equal(out_mine, out), (out_mine, out)
    torch.save(dict(cfg=c, seed=23, rms_eps=1e-6, rope_theta=1000000.0, input_ids=ids, mask=mask, prefill_logits=logits,
                    tokens=out, eos=2, pad=2, max_new_tokens=10, oracle_dev=dict(logits=d_l, tokens=0),
                    source="reference OpusQwenForCausalLM.generate over transformers " +
                           __import__("transformers").__version__),
               os.path.join(GOLD, "qwen_small.pt"))
    print(f"qwen_small.pt: oracle vs reference OpusQwenForCausalLM: logits max |diff| {d_l:.3g}, tokens identical {tuple(out.shape)}")


def golden_opt():
    """Sibling family (model/language_model/opus_opt.py, chosen at builder.py:71-81): the reference's OpusOPTForCausalLM
    (OPT: learned positions, LayerNorm, biased linears, ReLU MLP; Galactica: same blocks, erf-GELU, no biases) on tiny
    configs with hash-seeded weights; text-only greedy generate (`seq=None` branch, opus_opt.py:124-125) and
    last-position prefill logits. The OPT wrapper needs no transformers-version shim."""
    from multi_modality_model.multi_modality_v1.model.language_model.opus_opt import OpusOPTConfig, OpusOPTForCausalLM
    from oracle import opt_ref
    cases = {}
    for name, act, bias, seed in (("opt", "relu", True, 29), ("galactica", "gelu", False, 31)):
        c = dict(n_layers=2, dim=256, n_heads=2, ffn_dim=512, vocab=1024, max_pos=256)
        hf_cfg = OpusOPTConfig(vocab_size=c["vocab"], hidden_size=c["dim"], ffn_dim=c["ffn_dim"],
                               num_hidden_layers=c["n_layers"], num_attention_heads=c["n_heads"],
                               max_position_embeddings=c["max_pos"], word_embed_proj_dim=c["dim"],
                               do_layer_norm_before=True, activation_function=act, enable_bias=bias, dropout=0.0,
                               bos_token_id=0, eos_token_id=2, pad_token_id=1, tie_word_embeddings=False)
        hf_cfg._attn_implementation = "eager"
        lw = synth.opt_weights(c["n_layers"], c["dim"], c["n_heads"], c["ffn_dim"], c["vocab"], c["max_pos"], seed=seed,
                               bias=bias)
        model = OpusOPTForCausalLM(hf_cfg).eval()
        sd = dict(lw)
        sd["model.embed_tokens.weight"] = lw["model.decoder.embed_tokens.weight"]   # opus_opt.py:24 aliases the table
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not [m for m in missing if "protein" not in m and "switch" not in m], missing
        assert not unexpected, unexpected
        B, L = 5, 21
        g = torch.Generator().manual_seed(43)
        ids = torch.randint(3, c["vocab"], (B, L), generator=g)
        pad = 1
        for b in range(B):                                       # left padding, ragged
            ids[b, : b * 2] = pad
        mask = ids != pad
        with torch.no_grad():
            out = model.generate(ids, None, attention_mask=mask, pad_token_id=2, do_sample=False, max_new_tokens=10,
                                 use_cache=True)
            emb = model.get_model().embed_tokens(ids)
            logits = model(inputs_embeds=emb, attention_mask=mask).logits[:, -1, :].float().clone()
        ocfg = opt_ref.OptCfg(n_layers=c["n_layers"], dim=c["dim"], n_heads=c["n_heads"], ffn_dim=c["ffn_dim"],
                              vocab=c["vocab"], max_pos=c["max_pos"], activation=act)
        lg_mine, _ = opt_ref.opt_forward(lw, ocfg, emb.float(), mask, opt_ref.positions_from_mask(mask))
        out_mine = opt_ref.greedy_generate(lw, ocfg, emb.float(), mask, 10, eos_ids=(2,), pad_id=2)
        d_l = float((lg_mine - logits).abs().max())
        assert d_l < 2e-4, d_l
        assert out_mine.shape == out.shape and torch.equal(out_mine, out), (out_mine, out)
        cases[name] = dict(cfg=c, seed=seed, activation=act, bias=bias, input_ids=ids, mask=mask, prefill_logits=logits,
                           tokens=out, eos=2, pad=2, max_new_tokens=10, oracle_dev=dict(logits=d_l, tokens=0))
        print(f"opt_small.pt[{name}]: oracle vs reference OpusOPTForCausalLM: logits max |diff| {d_l:.3g}, "
              f"tokens identical {tuple(out.shape)}")
    cases["source"] = "reference OpusOPTForCausalLM.generate over transformers " + __import__("transformers").__version__
    torch.save(cases, os.path.join(GOLD, "opt_small.pt"))


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    w, seqs, pooled = golden_esm()
    golden_mm_and_llama(w, seqs, pooled)
    golden_qwen()
    golden_opt()
