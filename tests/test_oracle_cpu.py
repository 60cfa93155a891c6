"""CPU tests (no GPU): the oracle restatements against the committed golden fixtures (tests/golden/, produced by
oracle/make_golden.py from the reference's own classes and HF modules) and against live HF modules of this image."""
import os

import pytest
import torch

from opus_pllm_b200 import synth
from oracle import esm2_ref, llama_ref, mm_ref, ops_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def test_esm2_oracle_matches_golden():
    g = _load("esm2_small.pt")
    c = g["cfg"]
    w = synth.esm2_weights(c["n_layers"], c["dim"], c["ffn"], seed=g["seed"])
    assert torch.equal(esm2_ref.tokenize(g["seqs"]), g["tokens"])
    pooled = esm2_ref.get_protein_seq_embeddings(w, g["seqs"], c["n_layers"], c["n_heads"])
    assert torch.allclose(pooled, g["pooled"], rtol=0, atol=2e-5), float((pooled - g["pooled"]).abs().max())
    hidden = esm2_ref.esm2_forward(w, g["tokens"], c["n_layers"], c["n_heads"])
    valid = (g["tokens"] != 1)
    assert torch.allclose(hidden[valid], g["hidden"][valid], rtol=0, atol=3e-5)


def test_esm2_batching_invariance_and_autocast_noise():
    w = synth.esm2_weights(2, 128, 256, seed=3)
    seqs = synth.proteins(4, 5, 40, seed=9)
    together = esm2_ref.get_protein_seq_embeddings(w, seqs, 2, 2)
    alone = torch.cat([esm2_ref.get_protein_seq_embeddings(w, [s], 2, 2) for s in seqs])
    assert torch.allclose(together, alone, atol=1e-5)
    ac = esm2_ref.get_protein_seq_embeddings(w, seqs, 2, 2, torch.bfloat16)
    assert 0 < float((ac - together).abs().max()) < 0.1


def test_esm2_token_dropout_scale_with_mask_token():
    w = synth.esm2_weights(1, 128, 256, seed=4)
    tok = esm2_ref.tokenize(["ACDE"])
    tok[0, 2] = esm2_ref.MASK
    out = esm2_ref.esm2_forward(w, tok, 1, 2)
    assert torch.isfinite(out).all()


def test_mm_oracle_matches_reference_golden():
    g = _load("mm_small.pt")
    lg = _load("llama_small.pt")
    c, H = lg["cfg"], lg["cfg"]["dim"]
    pw = synth.projector_weights(g["esm_cfg"]["dim"], 5120, 8 * H, seed=g["proj_seed"])
    lw_embed = synth.weight((c["vocab"], H), "llama.embed", 0.02, lg["seed"])
    cstp = mm_ref.protein_forward(g["pooled"], pw["protein_projection.linear.weight"],
                                  pw["protein_projection.linear.bias"])
    assert torch.allclose(cstp, g["cstp_out"], atol=1e-5)
    soft = mm_ref.switch_projector(cstp, pw, H)
    assert soft.shape == (len(g["seqs"]), 8, H) and torch.allclose(soft, g["soft"], atol=1e-4)
    e, m, p, lens = mm_ref.splice(g["input_ids"], g["attention_mask"], soft, lw_embed, True)
    assert lens == g["lens"] and torch.equal(m, g["mask_left"]) and torch.allclose(e, g["embeds_left"], atol=1e-4)
    e, m, p, _ = mm_ref.splice(g["input_ids"], g["attention_mask"], soft, lw_embed, False)
    assert torch.equal(m, g["mask_right"]) and torch.equal(p, g["pos_right"])
    assert torch.allclose(e, g["embeds_right"], atol=1e-4)


def test_llama_oracle_matches_reference_generate_golden():
    g = _load("llama_small.pt")
    c = g["cfg"]
    w = synth.llama_weights(c["n_layers"], c["dim"], c["n_q_heads"], c["n_kv_heads"], c["head_dim"], c["ffn_dim"],
                            c["vocab"], seed=g["seed"])
    ocfg = llama_ref.LlamaCfg(n_layers=c["n_layers"], dim=c["dim"], n_q_heads=c["n_q_heads"],
                              n_kv_heads=c["n_kv_heads"], head_dim=c["head_dim"], ffn_dim=c["ffn_dim"], vocab=c["vocab"])
    mask = g["mask"]
    pos = (mask.long().cumsum(-1) - 1).masked_fill(~mask, 1)
    logits, _ = llama_ref.llama_forward(w, ocfg, g["embeds"], mask, pos)
    assert torch.allclose(logits, g["prefill_logits"], atol=5e-4), float((logits - g["prefill_logits"]).abs().max())
    out = llama_ref.greedy_generate(w, ocfg, g["embeds"], mask, g["max_new_tokens"], eos_ids=(g["eos"],),
                                    pad_id=g["pad"])
    assert torch.equal(out, g["tokens"])


def test_llama_oracle_matches_live_hf_llama():
    tf = pytest.importorskip("transformers")
    c = dict(n_layers=2, dim=256, n_q_heads=2, n_kv_heads=1, head_dim=128, ffn_dim=512, vocab=512)
    w = synth.llama_weights(c["n_layers"], c["dim"], 2, 1, 128, 512, 512, seed=5)
    cfg = tf.LlamaConfig(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                         num_attention_heads=2, num_key_value_heads=1, head_dim=128, rms_norm_eps=1e-5,
                         rope_theta=500000.0, max_position_embeddings=256, tie_word_embeddings=False)
    cfg._attn_implementation = "eager"
    hf = tf.LlamaForCausalLM(cfg).eval()
    hf.load_state_dict(w, strict=False)
    ids = torch.randint(0, 512, (3, 17), generator=torch.Generator().manual_seed(1))
    mask = torch.ones(3, 17, dtype=torch.bool)
    mask[1, :5] = False
    with torch.no_grad():
        want = hf(input_ids=ids, attention_mask=mask).logits[:, -1].float()
    ocfg = llama_ref.LlamaCfg(n_layers=2, dim=256, n_q_heads=2, n_kv_heads=1, head_dim=128, ffn_dim=512, vocab=512)
    pos = (mask.long().cumsum(-1) - 1).masked_fill(~mask, 1)
    got, _ = llama_ref.llama_forward(w, ocfg, w["model.embed_tokens.weight"][ids], mask, pos)
    assert torch.allclose(got, want, atol=5e-4), float((got - want).abs().max())


def test_llama_decode_with_cache_equals_full_recompute():
    c = llama_ref.LlamaCfg(n_layers=2, dim=256, n_q_heads=2, n_kv_heads=1, head_dim=128, ffn_dim=512, vocab=300)
    w = synth.llama_weights(2, 256, 2, 1, 128, 512, 300, seed=6)
    emb = w["model.embed_tokens.weight"][torch.arange(20).view(2, 10)]
    mask = torch.ones(2, 10, dtype=torch.bool)
    mask[0, :3] = False
    out = llama_ref.greedy_generate(w, c, emb, mask, 5)
    # recompute step 3 from scratch
    full = torch.cat([emb, w["model.embed_tokens.weight"][out[:, :3]]], 1)
    m2 = torch.cat([mask, torch.ones(2, 3, dtype=torch.bool)], 1)
    pos = (m2.long().cumsum(-1) - 1).masked_fill(~m2, 1)
    logits, _ = llama_ref.llama_forward(w, c, full, m2, pos)
    assert torch.equal(logits.argmax(-1), out[:, 3])


def test_lora_merge_and_op_refs():
    W, A, B = torch.randn(8, 6), torch.randn(2, 6), torch.randn(8, 2)
    assert torch.allclose(llama_ref.lora_merge_ref(W, A, B, 32.0, 16), W + 2.0 * B @ A)
    x = torch.randn(5, 64)
    g, u = torch.randn(5, 32), torch.randn(5, 32)
    inter = torch.stack([g, u], -1).reshape(5, 64)
    assert torch.allclose(ops_ref.swiglu_interleaved_ref(inter),
                          ops_ref.bf16r(ops_ref.bf16r(torch.nn.functional.silu(ops_ref.bf16r(g))) * ops_ref.bf16r(u)))
    assert torch.allclose(ops_ref.gelu_erf(x), torch.nn.functional.gelu(x), atol=1e-6)
    tok, fin = ops_ref.greedy_select_ref(torch.tensor([[0., 3., 3.], [9., 1., 0.]]), torch.tensor([0, 1]), [1], 7)
    assert tok.tolist() == [1, 7] and fin.tolist() == [True, True]


def test_top_p_oracle_matches_hf_warpers():
    """oracle/ops_ref.py::top_p_keep_mask / top_p_probs restate HF's TemperatureLogitsWarper + TopPLogitsWarper (what
    the reference's generate(do_sample=True, temperature, top_p) applies): pin them against transformers itself."""
    from transformers.generation.logits_process import TemperatureLogitsWarper, TopPLogitsWarper
    from oracle import ops_ref as R
    g = torch.Generator().manual_seed(0)
    for T, P in [(0.1, 0.7), (1.0, 0.9), (0.7, 0.5), (2.0, 1.0)]:
        lg = (torch.randn(6, 3000, generator=g) * 3).bfloat16()
        sc = TopPLogitsWarper(top_p=P)(None, TemperatureLogitsWarper(T)(None, lg.float()))
        assert torch.equal(torch.isfinite(sc), R.top_p_keep_mask(lg, T, P))
        assert torch.allclose(sc.softmax(-1), R.top_p_probs(lg, T, P))


def test_scoring_oracle_matches_reference_forward_golden():
    """oracle restatement of forward(labels=...) (right-padded splice, spliced labels, all-position logits, shifted
    cross entropy) against the reference's own forward output pinned in tests/golden/score_small.pt."""
    g, mm = _load("score_small.pt"), _load("mm_small.pt")
    c = g["cfg"]
    w = synth.llama_weights(c["n_layers"], c["dim"], c["n_q_heads"], c["n_kv_heads"], c["head_dim"], c["ffn_dim"],
                            c["vocab"], seed=g["seed"])
    ocfg = llama_ref.LlamaCfg(n_layers=c["n_layers"], dim=c["dim"], n_q_heads=c["n_q_heads"],
                              n_kv_heads=c["n_kv_heads"], head_dim=c["head_dim"], ffn_dim=c["ffn_dim"], vocab=c["vocab"])
    lab = mm_ref.splice_labels(g["input_ids"], g["attention_mask"], g["labels"])
    assert torch.equal(lab, g["labels_spliced"])
    e, m, _, _ = mm_ref.splice(g["input_ids"], g["attention_mask"], mm["soft"], w["model.embed_tokens.weight"], False)
    pos = (m.long().cumsum(-1) - 1).masked_fill(~m, 1)
    logits, _ = llama_ref.llama_forward(w, ocfg, e, m, pos, all_positions=True)
    assert float((logits - g["logits"].float())[m].abs().max()) <= 0.05          # fixture logits are stored in bf16
    assert abs(float(llama_ref.causal_lm_loss(logits, lab)) - float(g["loss"])) <= 1e-4


def test_qwen_oracle_matches_reference_wrapper_golden():
    """Sibling family (opus_qwen.py): the Llama restatement + q/k/v biases against the reference's own
    OpusQwenForCausalLM.generate output pinned in tests/golden/qwen_small.pt."""
    g = _load("qwen_small.pt")
    c = g["cfg"]
    w = synth.llama_weights(c["n_layers"], c["dim"], c["n_q_heads"], c["n_kv_heads"], c["head_dim"], c["ffn_dim"],
                            c["vocab"], seed=g["seed"], qkv_bias=True)
    ocfg = llama_ref.LlamaCfg(n_layers=c["n_layers"], dim=c["dim"], n_q_heads=c["n_q_heads"],
                              n_kv_heads=c["n_kv_heads"], head_dim=c["head_dim"], ffn_dim=c["ffn_dim"], vocab=c["vocab"],
                              rms_eps=g["rms_eps"], rope_theta=g["rope_theta"])
    mask = g["mask"]
    emb = w["model.embed_tokens.weight"][g["input_ids"]]
    pos = (mask.long().cumsum(-1) - 1).masked_fill(~mask, 1)
    logits, _ = llama_ref.llama_forward(w, ocfg, emb, mask, pos)
    assert torch.allclose(logits, g["prefill_logits"], atol=5e-4)
    out = llama_ref.greedy_generate(w, ocfg, emb, mask, g["max_new_tokens"], eos_ids=(g["eos"],), pad_id=g["pad"])
    assert torch.equal(out, g["tokens"])


@pytest.mark.parametrize("case", ["opt", "galactica"])
def test_opt_oracle_matches_reference_wrapper_golden(case):
    """Sibling family (opus_opt.py; builder.py:71-81): oracle/opt_ref.py against the reference's own
    OpusOPTForCausalLM.generate output pinned in tests/golden/opt_small.pt (OPT: ReLU + biases; Galactica: GELU, none)."""
    from oracle import opt_ref
    g = _load("opt_small.pt")[case]
    c = g["cfg"]
    w = synth.opt_weights(c["n_layers"], c["dim"], c["n_heads"], c["ffn_dim"], c["vocab"], c["max_pos"], seed=g["seed"],
                          bias=g["bias"])
    ocfg = opt_ref.OptCfg(n_layers=c["n_layers"], dim=c["dim"], n_heads=c["n_heads"], ffn_dim=c["ffn_dim"],
                          vocab=c["vocab"], max_pos=c["max_pos"], activation=g["activation"])
    mask = g["mask"]
    emb = w["model.decoder.embed_tokens.weight"][g["input_ids"]]
    logits, _ = opt_ref.opt_forward(w, ocfg, emb, mask, opt_ref.positions_from_mask(mask))
    assert torch.allclose(logits, g["prefill_logits"], atol=5e-4)
    out = opt_ref.greedy_generate(w, ocfg, emb, mask, g["max_new_tokens"], eos_ids=(g["eos"],), pad_id=g["pad"])
    assert torch.equal(out, g["tokens"])
