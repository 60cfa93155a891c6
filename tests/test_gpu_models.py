"""GPU parity of the composite forwards (encoder, projectors, splice, prefill logits, greedy decode) vs the oracle.

Tolerances (north star: "within a stated bf16 tolerance, max-abs and cosine"):
  * encoder pooled embedding  : cosine >= 0.9995, max-abs <= 3e-2 (values are post-LayerNorm, O(1))
  * projector soft tokens     : cosine >= 0.999
  * prefill last-token logits : cosine >= 0.999 vs the fp32 oracle, max-abs <= 6 % of the logit std
  * greedy decode             : token-identical over 32 new tokens on >= 99 % of prompts (peaked-logit weights)
"""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import esm2_ref, llama_ref, mm_ref  # noqa: E402
from opus_pllm_b200 import synth  # noqa: E402


def _cos(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm()))


def _dev(d, dtype=None):
    return {k: (v.cuda() if dtype is None else v.cuda().to(dtype)) for k, v in d.items()}


# ------------------------------------------------------------------------------------------------ encoder
@pytest.mark.parametrize("n_layers,dim,heads,ffn,lens", [
    (2, 128, 2, 512, [5, 1, 40, 130, 258]),
    (3, 1280, 20, 5120, [256, 17, 300]),
])
def test_encoder_vs_oracle(n_layers, dim, heads, ffn, lens):
    from opus_pllm_b200.encoder import B200ProteinEncoder
    w = synth.esm2_weights(n_layers, dim, ffn)
    seqs = [s[:n] for s, n in zip(synth.proteins(len(lens), max(lens)), lens)]
    seqs[1] = seqs[1][:-1] + "X" if len(seqs[1]) > 0 else seqs[1]  # a non-standard residue goes through the LUT
    enc = B200ProteinEncoder(w, n_layers, dim, heads, ffn)
    pooled, pooled_l2, hidden, pk = enc.encode(seqs, want_hidden=True)
    wd = _dev(w)
    want = esm2_ref.get_protein_seq_embeddings(wd, seqs, n_layers, heads)            # fp32 truth
    want_ac = esm2_ref.get_protein_seq_embeddings(wd, seqs, n_layers, heads, torch.bfloat16)  # reference-style autocast
    assert torch.equal(torch.from_numpy(pk.tokens).long(),
                       torch.cat([esm2_ref.tokenize([s])[0] for s in seqs]))
    # the oracle's own bf16-autocast run bounds what "bf16 tolerance" means for this depth
    noise = float((want_ac - want).abs().max())
    err = float((pooled - want).abs().max())
    assert _cos(pooled, want) >= 0.9995, _cos(pooled, want)
    assert err <= max(3e-2, 3 * noise), (err, noise)
    for b in range(len(seqs)):
        assert _cos(pooled[b], want[b]) >= 0.9995
    l2 = torch.nn.functional.normalize(want, dim=-1)
    assert _cos(pooled_l2, l2) >= 0.9995
    assert enc.get_protein_seq_embeddings(seqs).dtype == torch.float32


def test_encoder_fused_rope_epilogue_matches_separate_kernel():
    """The q|k|v GEMM applies ESM's rotary embedding in its TMA-store epilogue; with the tunable off the same forward
    runs the plain epilogue + rope_esm_kernel. Both must agree to bf16 rounding (one FMA contraction apart)."""
    from opus_pllm_b200 import _lib as L
    from opus_pllm_b200.encoder import B200ProteinEncoder
    n_layers, dim, heads, ffn = 3, 1280, 20, 5120
    w = synth.esm2_weights(n_layers, dim, ffn, seed=21, device="cuda")
    enc = B200ProteinEncoder(w, n_layers, dim, heads, ffn)
    seqs = synth.proteins(5, 60, 300, seed=17)
    lib = L.load()
    try:
        fused, _, hid_f, _ = enc.encode(seqs, want_hidden=True)
        fused, hid_f = fused.clone(), hid_f.clone()
        L.check(lib.opus_set_tunable(b"tma_store", 0))
        plain, _, hid_p, _ = enc.encode(seqs, want_hidden=True)
    finally:
        L.check(lib.opus_set_tunable(b"tma_store", 1))
    assert _cos(hid_f, hid_p) >= 0.99999 and float((fused - plain).abs().max()) <= 5e-3
    want = esm2_ref.get_protein_seq_embeddings(w, seqs, n_layers, heads)
    assert _cos(fused, want) >= 0.9995 and float((fused - want).abs().max()) <= 3e-2


def test_encoder_padded_equals_varlen():
    """key-padding mask semantics: a sequence's embedding does not depend on what it is batched with."""
    from opus_pllm_b200.encoder import B200ProteinEncoder
    w = synth.esm2_weights(2, 128, 512)
    enc = B200ProteinEncoder(w, 2, 128, 2, 512)
    seqs = synth.proteins(6, 10, 200, seed=7)
    together = enc.get_protein_seq_embeddings(seqs)
    alone = torch.cat([enc.get_protein_seq_embeddings([s]) for s in seqs])
    assert torch.allclose(together, alone, rtol=0, atol=2e-3), float((together - alone).abs().max())


# ------------------------------------------------------------------------------------------------ projectors
def test_projectors_vs_oracle():
    from opus_pllm_b200.projector import B200ProteinProjector, B200SwitchProjector, FusedProjectors
    from opus_pllm_b200 import ops
    H = 256
    pw = synth.projector_weights(1280, 5120, 8 * H)
    x = synth.weight((12, 1280), "pooled", 1.0).cuda()
    pp = B200ProteinProjector(pw["protein_projection.linear.weight"], pw["protein_projection.linear.bias"])
    sw = B200SwitchProjector(5120, 8 * H).load_state_dict(pw)
    pwd = _dev(pw)
    c_want = mm_ref.protein_forward(x, pwd["protein_projection.linear.weight"], pwd["protein_projection.linear.bias"])
    c_got = pp.protein_forward(x)
    assert c_got.shape == (12, 5120) and _cos(c_got, c_want) >= 0.9995
    s_want = mm_ref.switch_projector(c_want, pwd, H)
    s_got = sw(c_got).reshape(12, 8, H)
    assert _cos(s_got, s_want) >= 0.999, _cos(s_got, s_want)
    fused = FusedProjectors(pp, sw)(ops.l2norm(x)).reshape(12, 8, H)
    assert torch.equal(fused, s_got)
    # 'linear' projector type and identity CSTP projector
    lin = B200SwitchProjector(1280, 8 * H, "linear").load_state_dict(
        {"weight": pw["0.weight"][:, :1280].contiguous(), "bias": pw["0.bias"]})
    l_want = torch.nn.functional.linear(x, pwd["0.weight"][:, :1280], pwd["0.bias"])
    assert _cos(lin(x.to(torch.bfloat16)), l_want) >= 0.999


# ------------------------------------------------------------------------------------------------ splice
def test_splice_vs_oracle_and_left_padding():
    from opus_pllm_b200.model import build_from_state_dicts
    cfg = dict(n_layers=1, dim=256, n_q_heads=2, n_kv_heads=1, head_dim=128, ffn_dim=512, vocab=512)
    lw = synth.llama_weights(cfg["n_layers"], cfg["dim"], 2, 1, 128, 512, 512)
    esm_cfg = dict(n_layers=1, dim=128, n_heads=2, ffn_dim=256)
    ew = synth.esm2_weights(1, 128, 256)
    sw = synth.projector_weights(128, 128, 8 * 256)
    model = build_from_state_dicts(lw, cfg, ew, esm_cfg, None,
                                   {k: v for k, v in sw.items() if k[0] in "02"}, eos_token_id=[3])
    ids = torch.tensor([[7, 7, 5, -200, 9, 10, 11],
                        [7, 7, 7, 7, 20, 21, 22],      # no sentinel: consumes a protein slot
                        [7, 30, -200, 31, -200, 32, 33]]).cuda()
    mask = ids != 7
    seqs = synth.proteins(4, 12, 30, seed=3)
    soft = model._soft_tokens(seqs, None)
    out = model.prepare_inputs_labels_for_multimodal(ids, None, mask, None, None, seqs, inference_mode=True)
    assert out[0] is None and out[1] is None and out[5] is None
    want_e, want_m, want_p, lens = mm_ref.splice(ids, mask, soft.float(), model.llama.embed.float(), True)
    assert lens == [12, 3, 20]
    assert torch.equal(out[4].float(), want_e) and torch.equal(out[2].bool(), want_m)
    # right padding + position ids + labels (training-side contract, opus_arch.py:259-269)
    labels = torch.where(ids == -200, torch.full_like(ids, -100), ids)
    pos_in = torch.arange(ids.shape[1]).cuda()
    out_r = model.prepare_inputs_labels_for_multimodal(ids, pos_in, mask, None, labels, seqs, inference_mode=False)
    want_e, want_m, want_p, _ = mm_ref.splice(ids, mask, soft.float(), model.llama.embed.float(), False)
    assert torch.equal(out_r[4].float(), want_e) and torch.equal(out_r[1], want_p)
    assert out_r[5].shape == want_m.shape and out_r[5][0].tolist()[:12] == [5] + [-100] * 8 + [9, 10, 11]
    with pytest.raises(NotImplementedError):
        model.generate(ids, seqs, inputs_embeds=out[4])
    with pytest.raises(NotImplementedError):
        model.encode_seq2embedding([1, 2, 3])


# ------------------------------------------------------------------------------------------------ llama
SMALL = dict(n_layers=2, dim=512, n_q_heads=4, n_kv_heads=2, head_dim=128, ffn_dim=1024, vocab=2048)


def _oracle_cfg(c):
    return llama_ref.LlamaCfg(n_layers=c["n_layers"], dim=c["dim"], n_q_heads=c["n_q_heads"],
                              n_kv_heads=c["n_kv_heads"], head_dim=c["head_dim"], ffn_dim=c["ffn_dim"],
                              vocab=c["vocab"])


def _padded(embeds_packed, cu, dim):
    lens = np.diff(cu)
    B, Lm = len(lens), int(lens.max())
    e = torch.zeros(B, Lm, dim, dtype=embeds_packed.dtype, device=embeds_packed.device)
    m = torch.zeros(B, Lm, dtype=torch.bool, device=embeds_packed.device)
    for b in range(B):
        e[b, Lm - lens[b]:] = embeds_packed[cu[b]: cu[b + 1]]
        m[b, Lm - lens[b]:] = True
    return e, m


@pytest.mark.parametrize("cfg,lens", [(SMALL, [40, 7, 129, 64]),
                                      (dict(n_layers=2, dim=4096, n_q_heads=32, n_kv_heads=8, head_dim=128,
                                            ffn_dim=14336, vocab=8192), [300, 33])])
def test_prefill_logits_vs_oracle(cfg, lens):
    from opus_pllm_b200.llama import B200Llama
    w = synth.llama_weights(seed=1, device="cuda", **{k if k != "ffn_dim" else "ffn": v for k, v in cfg.items()})
    model = B200Llama(w, **cfg)
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    emb = synth.weight((int(cu[-1]), cfg["dim"]), "prompt_embeds", 0.02).cuda().to(torch.bfloat16)
    st = model.prefill(emb, cu, 4)
    got = st["logits"].float().clone()
    model.release_plan(st)
    e_pad, m_pad = _padded(emb, cu, cfg["dim"])
    pos = (m_pad.long().cumsum(-1) - 1).masked_fill(~m_pad, 1)
    ocfg = _oracle_cfg(cfg)
    want32, _ = llama_ref.llama_forward(_dev(w, torch.float32), ocfg, e_pad.float(), m_pad, pos)
    want16, _ = llama_ref.llama_forward(_dev(w, torch.bfloat16), ocfg, e_pad, m_pad, pos)
    sigma = float(want32.std())
    noise = float((want16 - want32).abs().max())          # the reference bf16 path's own distance from fp32
    err = float((got - want32).abs().max())
    assert _cos(got, want32) >= 0.999, _cos(got, want32)
    assert err <= max(0.06 * sigma, 2.0 * noise), (err, sigma, noise)
    for b in range(len(lens)):
        assert _cos(got[b], want32[b]) >= 0.999


@pytest.mark.parametrize("cfg,lens", [(SMALL, [140, 37, 129, 64]),
                                      (dict(n_layers=2, dim=4096, n_q_heads=32, n_kv_heads=8, head_dim=128,
                                            ffn_dim=14336, vocab=8192), [300, 33, 200]),
                                      (dict(n_layers=2, dim=4096, n_q_heads=32, n_kv_heads=8, head_dim=128,
                                            ffn_dim=14336, vocab=8192), [700, 512, 130])])   # >= 1024 rows: CTA-pair GEMMs
def test_prefill_fused_rope_kvappend_epilogue_is_bit_identical(cfg, lens):
    """The prefill q|k|v GEMM rotates q/k and appends k/v to the paged cache in its epilogue; with the tunable off the
    plain epilogue + rope_llama_kvappend_kernel run instead. Same rounding points -> identical logits and cache."""
    from opus_pllm_b200 import _lib as L
    from opus_pllm_b200.llama import B200Llama
    w = synth.llama_weights(seed=9, device="cuda", **{k if k != "ffn_dim" else "ffn": v for k, v in cfg.items()})
    model = B200Llama(w, **cfg)
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    emb = synth.weight((int(cu[-1]), cfg["dim"]), "prompt_embeds_rope", 0.02).cuda().to(torch.bfloat16)
    plan = model.make_plan(cu, 2)
    lib = L.load()
    try:
        model._k.zero_(); model._v.zero_()
        fused = model.prefill(emb, plan=plan)["logits"].clone()
        k_f, v_f = model._k.clone(), model._v.clone()
        L.check(lib.opus_set_tunable(b"tma_store", 0))
        model._k.zero_(); model._v.zero_()
        plain = model.prefill(emb, plan=plan)["logits"].clone()
        k_p, v_p = model._k.clone(), model._v.clone()
    finally:
        L.check(lib.opus_set_tunable(b"tma_store", 1))
        model.release_plan(plan)
    assert bool(k_p.abs().sum() > 0) and torch.equal(k_f, k_p) and torch.equal(v_f, v_p)
    assert torch.equal(fused, plain)


def test_greedy_decode_token_parity_peaked():
    """>= 99 % of prompts token-identical over 32 new tokens (north star), peaked-logit synthetic weights."""
    from opus_pllm_b200.llama import B200Llama
    cfg = SMALL
    w = synth.llama_weights(seed=2, peaked=True, device="cuda", **{k if k != "ffn_dim" else "ffn": v for k, v in cfg.items()})
    model = B200Llama(w, **cfg)
    lens = [30 + (i * 7) % 50 for i in range(32)]
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    tok = torch.randint(0, cfg["vocab"], (int(cu[-1]),), generator=torch.Generator().manual_seed(5))
    emb = w["model.embed_tokens.weight"][tok].cuda().to(torch.bfloat16)
    got = model.generate_packed(emb, cu, 32)
    got_nograph = model.generate_packed(emb, cu, 32, use_graph=False)
    assert torch.equal(got, got_nograph)
    e_pad, m_pad = _padded(emb, cu, cfg["dim"])
    want = llama_ref.greedy_generate(_dev(w, torch.bfloat16), _oracle_cfg(cfg), e_pad, m_pad, 32)
    same = (got.cpu() == want.cpu()).all(1).float().mean()
    assert got.shape == (32, 32) and float(same) >= 0.99, float(same)
    assert len(torch.unique(got)) > 32  # not a degenerate constant stream


@pytest.mark.parametrize("cfg,batch", [(SMALL, 32), (SMALL, 5),
                                       (dict(n_layers=2, dim=4096, n_q_heads=32, n_kv_heads=8, head_dim=128,
                                             ffn_dim=14336, vocab=32768), 64)])
def test_decode_fused_chain_matches_kernel_per_op(cfg, batch):
    """The fused decode chain (one persistent kernel for o_proj .. next qkv / lm_head, device-wide barriers between its
    phases) must reproduce the kernel-per-op path: same tokens, and the same last-step logits up to the one place where
    the two differ (the order of the RMSNorm sum of squares). The fused form is opt-in (tunable decode_fused)."""
    from opus_pllm_b200 import _lib as L
    from opus_pllm_b200.llama import B200Llama
    w = synth.llama_weights(seed=4, peaked=True, device="cuda", **{k if k != "ffn_dim" else "ffn": v for k, v in cfg.items()})
    model = B200Llama(w, **cfg)
    lens = [17 + (i * 5) % 40 for i in range(batch)]
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    tok = torch.randint(0, cfg["vocab"], (int(cu[-1]),), generator=torch.Generator().manual_seed(6))
    emb = w["model.embed_tokens.weight"][tok].cuda().to(torch.bfloat16)
    lib = L.load()
    try:
        L.check(lib.opus_set_tunable(b"decode_fused", 1))
        fused = model.generate_packed(emb, cu, 12)
        fused_logits = model._ws_bufs["logits"][:batch].float().clone()
        fused_eager = model.generate_packed(emb, cu, 12, use_graph=False)
        L.check(lib.opus_set_tunable(b"decode_fused", 0))
        plain = model.generate_packed(emb, cu, 12)
        plain_logits = model._ws_bufs["logits"][:batch].float().clone()
    finally:
        L.check(lib.opus_set_tunable(b"decode_fused", 0))
    assert torch.equal(fused, fused_eager)
    assert torch.equal(fused, plain), float((fused == plain).float().mean())
    # 12 steps of bf16 one-ulp differences (RMSNorm reduction order) compound through the KV cache
    assert _cos(fused_logits, plain_logits) >= 0.995


@pytest.mark.parametrize("cfg,batch,qkv_bias", [(SMALL, 32, False), (SMALL, 5, True), (SMALL, 50, False),
                                                 (dict(n_layers=2, dim=4096, n_q_heads=32, n_kv_heads=8, head_dim=128,
                                                       ffn_dim=14336, vocab=32768), 64, False)])
def test_decode_norm_fused_matches_kernel_per_op(cfg, batch, qkv_bias):
    """Norm-fused decode (tunable decode_norm_fused): o_proj / down reduce their split-K sums inside the GEMM and emit the
    per-slab sums of squares, q|k|v and gate/up normalise the raw residual stream on its way into the tensor cores. Same
    rounding points as the kernel-per-op path except the order of the RMSNorm sum of squares: tokens equal (peaked
    weights), last-step logits equal to bf16 noise; graph and eager launches bit-identical; and the oracle agrees."""
    from opus_pllm_b200 import _lib as L
    from opus_pllm_b200.llama import B200Llama
    w = synth.llama_weights(seed=4, peaked=True, device="cuda", qkv_bias=qkv_bias,
                            **{k if k != "ffn_dim" else "ffn": v for k, v in cfg.items()})
    model = B200Llama(w, **cfg)
    lens = [17 + (i * 5) % 40 for i in range(batch)]
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    tok = torch.randint(0, cfg["vocab"], (int(cu[-1]),), generator=torch.Generator().manual_seed(6))
    emb = w["model.embed_tokens.weight"][tok.cuda()].to(torch.bfloat16)
    lib = L.load()
    try:
        L.check(lib.opus_set_tunable(b"decode_norm_fused", 1))
        fused = model.generate_packed(emb, cu, 12)
        fused_logits = model._ws_bufs["logits"][:batch].float().clone()
        fused_eager = model.generate_packed(emb, cu, 12, use_graph=False)
        L.check(lib.opus_set_tunable(b"decode_norm_fused", 0))
        plain = model.generate_packed(emb, cu, 12)
        plain_logits = model._ws_bufs["logits"][:batch].float().clone()
    finally:
        L.check(lib.opus_set_tunable(b"decode_norm_fused", 0))
    assert torch.equal(fused, fused_eager)
    assert torch.equal(fused, plain), float((fused == plain).float().mean())
    assert _cos(fused_logits, plain_logits) >= 0.995
    if cfg is SMALL:
        e_pad, m_pad = _padded(emb, cu, cfg["dim"])
        want = llama_ref.greedy_generate(_dev(w, torch.bfloat16), _oracle_cfg(cfg), e_pad, m_pad, 12)
        assert float((fused.cpu() == want.cpu()).all(1).float().mean()) >= 0.99


@pytest.mark.parametrize("cfg,batch", [(SMALL, 9), (dict(n_layers=2, dim=4096, n_q_heads=32, n_kv_heads=8, head_dim=128,
                                                        ffn_dim=14336, vocab=8192), 64)])
def test_decode_rope_fused_into_attention_is_bit_identical(cfg, batch):
    """Decode: the attention CTAs reduce the q|k|v split-K partials of their heads, apply RoPE and append K/V themselves
    (default) — same arithmetic as the separate rope_llama_kvappend kernel, so tokens, cache and logits are identical."""
    from opus_pllm_b200 import _lib as L
    from opus_pllm_b200.llama import B200Llama
    w = synth.llama_weights(seed=12, device="cuda", **{k if k != "ffn_dim" else "ffn": v for k, v in cfg.items()})
    model = B200Llama(w, **cfg)
    lens = [21 + (i * 11) % 60 for i in range(batch)]
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    emb = synth.weight((int(cu[-1]), cfg["dim"]), "rope_fused_prompt", 0.02).cuda().to(torch.bfloat16)
    lib = L.load()
    plan = model.make_plan(cu, 10)
    try:
        res = {}
        for mode in (1, 0):
            L.check(lib.opus_set_tunable(b"decode_rope_fused", mode))
            model._k.zero_(); model._v.zero_()
            st = model.prefill(emb, plan=plan)
            toks = model.generate_from_prefill(st, 10)
            res[mode] = (toks.clone(), model._ws_bufs["logits"][:batch].clone(), model._k.clone(), model._v.clone())
    finally:
        L.check(lib.opus_set_tunable(b"decode_rope_fused", 1))
        model.release_plan(plan)
    for a, b in zip(res[1], res[0]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("n_q,n_kv,batch", [(3, 1, 7), (12, 2, 33), (28, 4, 64)])
def test_qwen_style_decode_token_parity(n_q, n_kv, batch):
    """Qwen2-style blocks (q/k/v biases; GQA groups 3 / 6 / 7; theta 1e6) through prefill + the decode loop: tokens vs
    the oracle on peaked-logit weights, CUDA graph == plain launches, fused RoPE paths == separate kernels."""
    from opus_pllm_b200 import _lib as L
    from opus_pllm_b200.llama import B200Llama
    cfg = dict(n_layers=2, dim=n_q * 128, n_q_heads=n_q, n_kv_heads=n_kv, head_dim=128, ffn_dim=1024, vocab=2048)
    w = synth.llama_weights(seed=6, peaked=True, device="cuda", qkv_bias=True,
                            **{k if k != "ffn_dim" else "ffn": v for k, v in cfg.items()})
    model = B200Llama(w, **cfg, rms_eps=1e-6, rope_theta=1e6)
    lens = [12 + (i * 13) % 70 for i in range(batch)]
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    tok = torch.randint(0, cfg["vocab"], (int(cu[-1]),), generator=torch.Generator().manual_seed(8))
    emb = w["model.embed_tokens.weight"][tok].cuda().to(torch.bfloat16)
    got = model.generate_packed(emb, cu, 16)
    assert torch.equal(got, model.generate_packed(emb, cu, 16, use_graph=False))
    lib = L.load()
    try:
        L.check(lib.opus_set_tunable(b"decode_rope_fused", 0))
        L.check(lib.opus_set_tunable(b"tma_store", 0))
        assert torch.equal(got, model.generate_packed(emb, cu, 16))
    finally:
        L.check(lib.opus_set_tunable(b"decode_rope_fused", 1))
        L.check(lib.opus_set_tunable(b"tma_store", 1))
    e_pad, m_pad = _padded(emb, cu, cfg["dim"])
    ocfg = llama_ref.LlamaCfg(n_layers=2, dim=cfg["dim"], n_q_heads=n_q, n_kv_heads=n_kv, head_dim=128, ffn_dim=1024,
                              vocab=2048, rms_eps=1e-6, rope_theta=1e6)
    want = llama_ref.greedy_generate(_dev(w, torch.bfloat16), ocfg, e_pad, m_pad, 16)
    assert float((got.cpu() == want.cpu()).all(1).float().mean()) >= 0.99


def test_greedy_decode_default_init_margin_aware():
    """HF-init statistics: random logits have tiny top-1 margins, so token identity is fragile by construction
    (SURVEY.md §7); every disagreement must be explained by a near-tie in the oracle's own logits."""
    from opus_pllm_b200.llama import B200Llama
    cfg = SMALL
    w = synth.llama_weights(seed=3, device="cuda", **{k if k != "ffn_dim" else "ffn": v for k, v in cfg.items()})
    model = B200Llama(w, **cfg)
    lens = [20 + i for i in range(16)]
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    emb = synth.weight((int(cu[-1]), cfg["dim"]), "e", 0.02).cuda().to(torch.bfloat16)
    got, logits = model.generate_packed(emb, cu, 8, return_prefill_logits=True)
    e_pad, m_pad = _padded(emb, cu, cfg["dim"])
    want, wl = llama_ref.greedy_generate(_dev(w, torch.float32), _oracle_cfg(cfg), e_pad.float(), m_pad, 8,
                                         return_logits=True)
    first = got[:, 0].cpu() == want[:, 0].cpu()
    top2 = wl[:, 0].topk(2, dim=-1).values
    margin = (top2[:, 0] - top2[:, 1]).cpu()
    assert bool((first | (margin < 0.05 * float(wl[:, 0].std()))).all())
    assert _cos(logits, wl[:, 0]) >= 0.999


def test_generate_sampling_default_eval_settings():
    """generate(do_sample=True, temperature=0.1, top_p=0.7) — the reference eval scripts' default (run_opus_ddp.py:
    126-128,156-157): runs on the device, is reproducible for a given seed, graph replay == plain launches, and at this
    temperature the peaked-logit model's draws coincide with greedy on nearly every position."""
    from opus_pllm_b200 import presets
    model = presets.build_synthetic_model("tiny", "cuda", peaked=True)
    B, new = 6, 10
    seqs = synth.proteins(B, 20, 40, seed=3)
    ids = torch.stack(synth.prompt_ids(B, 30, vocab=2048, sentinel_at=5)).cuda()
    kw = dict(pad_token_id=1, max_new_tokens=new)
    greedy = model.generate(ids, seqs, do_sample=False, **kw)
    a = model.generate(ids, seqs, do_sample=True, temperature=0.1, top_p=0.7, seed=11, **kw)
    b = model.generate(ids, seqs, do_sample=True, temperature=0.1, top_p=0.7, seed=11, use_graph=False, **kw)
    assert a.shape == (B, new) and a.dtype == torch.int64 and torch.equal(a, b)
    assert float((a == greedy).float().mean()) >= 0.9
    flat = presets.build_synthetic_model("tiny", "cuda", peaked=False)      # HF-init statistics: flat next-token distribution
    flat_greedy = flat.generate(ids, seqs, do_sample=False, **kw)
    hot = flat.generate(ids, seqs, do_sample=True, temperature=1.0, top_p=1.0, seed=11, **kw)
    hot2 = flat.generate(ids, seqs, do_sample=True, temperature=1.0, top_p=1.0, seed=12, **kw)
    assert not torch.equal(hot, flat_greedy) and not torch.equal(hot, hot2)
    with pytest.raises(ValueError):
        model.generate(ids, seqs, do_sample=True, temperature=0.0, **kw)
    with pytest.raises(NotImplementedError):
        model.generate(ids, seqs, do_sample=False, num_beams=4, **kw)


def test_eos_and_pad_semantics():
    from opus_pllm_b200.llama import B200Llama
    cfg = SMALL
    w = synth.llama_weights(seed=2, peaked=True, device="cuda", **{k if k != "ffn_dim" else "ffn": v for k, v in cfg.items()})
    model = B200Llama(w, **cfg)
    lens = [12, 20, 9, 15]
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    tok = torch.randint(0, cfg["vocab"], (int(cu[-1]),), generator=torch.Generator().manual_seed(6))
    emb = w["model.embed_tokens.weight"][tok].cuda().to(torch.bfloat16)
    free = model.generate_packed(emb, cu, 12)
    eos = [int(free[0, 3]), int(free[1, 6])]            # make rows 0 and 1 stop early
    got = model.generate_packed(emb, cu, 12, eos_ids=eos, pad_id=eos[0], check_every=2)
    e_pad, m_pad = _padded(emb, cu, cfg["dim"])
    want = llama_ref.greedy_generate(_dev(w, torch.bfloat16), _oracle_cfg(cfg), e_pad, m_pad, 12, eos_ids=eos,
                                     pad_id=eos[0])
    assert got.shape == want.shape and torch.equal(got.cpu(), want.cpu())


def test_lora_merge_at_load():
    from opus_pllm_b200.llama import B200Llama
    cfg = SMALL
    kw = {k if k != "ffn_dim" else "ffn": v for k, v in cfg.items()}
    w = synth.llama_weights(seed=4, device="cuda", **kw)
    lora = synth.lora_adapters(w, cfg["n_layers"], r=16, device="cuda")
    model = B200Llama(w, lora=lora, lora_alpha=32.0, lora_r=16, **cfg)
    wm = dict(w)
    for key in [k[: -len(".lora_A.weight")] for k in lora if k.endswith(".lora_A.weight")]:
        wm[key + ".weight"] = llama_ref.lora_merge_ref(w[key + ".weight"], lora[key + ".lora_A.weight"],
                                                       lora[key + ".lora_B.weight"], 32.0, 16)
    lens = [33, 21]
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    emb = synth.weight((int(cu[-1]), cfg["dim"]), "e2", 0.02).cuda().to(torch.bfloat16)
    st = model.prefill(emb, cu, 2)
    got = st["logits"].float().clone()
    model.release_plan(st)
    e_pad, m_pad = _padded(emb, cu, cfg["dim"])
    pos = (m_pad.long().cumsum(-1) - 1).masked_fill(~m_pad, 1)
    want, _ = llama_ref.llama_forward(_dev(wm, torch.float32), _oracle_cfg(cfg), e_pad.float(), m_pad, pos)
    base, _ = llama_ref.llama_forward(_dev(w, torch.float32), _oracle_cfg(cfg), e_pad.float(), m_pad, pos)
    assert _cos(got, want) >= 0.999 and _cos(got, want) > _cos(got, base)


@pytest.mark.parametrize("qkv_bias", [True, False])
def test_narrow_heads_llama_family_vs_oracle(qkv_bias):
    """head_dim 64 (Qwen2-0.5B: 14 q heads / 2 kv heads x 64, biased q/k/v): heads are stored zero-padded to the kernels' 128
    columns in the rotate_half layout; prefill logits and greedy tokens must match the oracle run at the real head width."""
    from opus_pllm_b200.llama import B200Llama
    cfg = dict(n_layers=3, dim=448, n_q_heads=7, n_kv_heads=1, head_dim=64, ffn_dim=1024, vocab=2048)
    kw = {k if k != "ffn_dim" else "ffn": v for k, v in cfg.items()}
    w = synth.llama_weights(seed=12, peaked=True, device="cuda", qkv_bias=qkv_bias, **kw)
    model = B200Llama(w, **cfg, rope_theta=1000000.0)
    assert model.hd == 128 and model.hd_real == 64
    ocfg = llama_ref.LlamaCfg(n_layers=3, dim=448, n_q_heads=7, n_kv_heads=1, head_dim=64, ffn_dim=1024, vocab=2048,
                              rope_theta=1000000.0)
    for lens, new in (([40, 7, 129, 64], 12), ([30 + (i * 7) % 50 for i in range(40)], 8)):
        cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        tok = torch.randint(0, cfg["vocab"], (int(cu[-1]),), generator=torch.Generator().manual_seed(5))
        emb = w["model.embed_tokens.weight"][tok.cuda()].to(torch.bfloat16)
        got, logits = model.generate_packed(emb, cu, new, return_prefill_logits=True)
        assert torch.equal(got, model.generate_packed(emb, cu, new, use_graph=False))
        e_pad, m_pad = _padded(emb, cu, cfg["dim"])
        pos = (m_pad.long().cumsum(-1) - 1).masked_fill(~m_pad, 1)
        want32, _ = llama_ref.llama_forward(_dev(w, torch.float32), ocfg, e_pad.float(), m_pad, pos)
        assert _cos(logits, want32) >= 0.999
        want = llama_ref.greedy_generate(_dev(w, torch.bfloat16), ocfg, e_pad, m_pad, new)
        same = (got.cpu() == want.cpu()).all(1).float().mean()
        assert float(same) >= 0.99, float(same)


# ------------------------------------------------------------------------------------------------ whole path
def test_generate_end_to_end_vs_oracle_pipeline():
    """proteins + prompts with -200 sentinels -> tokens, through the reference-shaped generate() call."""
    from opus_pllm_b200.model import build_from_state_dicts
    cfg = SMALL
    kw = {k if k != "ffn_dim" else "ffn": v for k, v in cfg.items()}
    lw = synth.llama_weights(seed=2, peaked=True, device="cuda", **kw)
    esm_cfg = dict(n_layers=2, dim=128, n_heads=2, ffn_dim=512)
    ew = synth.esm2_weights(2, 128, 512)
    pw = synth.projector_weights(128, 256, 8 * cfg["dim"])
    model = build_from_state_dicts(lw, cfg, ew, esm_cfg, pw, pw)
    B = 8
    seqs = synth.proteins(B, 20, 90, seed=11)
    prompts = synth.prompt_ids(B, 48, vocab=cfg["vocab"], ragged=5, sentinel_at=10)
    L = max(p.numel() for p in prompts)
    pad = 1
    ids = torch.stack([torch.cat([torch.full((L - p.numel(),), pad), p]) for p in prompts]).cuda()
    mask = ids != pad
    got = model.generate(ids, seqs, attention_mask=mask, pad_token_id=pad, do_sample=False, temperature=0,
                         top_p=0.7, num_beams=1, max_new_tokens=32, use_cache=True)
    assert got.dtype == torch.int64 and got.shape == (B, 32)
    # oracle pipeline on the same weights
    pooled = esm2_ref.get_protein_seq_embeddings(_dev(ew), seqs, 2, 2)
    pwd = _dev(pw)
    c = mm_ref.protein_forward(pooled, pwd["protein_projection.linear.weight"], pwd["protein_projection.linear.bias"])
    soft = mm_ref.switch_projector(c, pwd, cfg["dim"])
    emb, m, _, _ = mm_ref.splice(ids, mask, soft.to(torch.bfloat16), lw["model.embed_tokens.weight"].cuda().bfloat16())
    want = llama_ref.greedy_generate(_dev(lw, torch.bfloat16), _oracle_cfg(cfg), emb, m, 32)
    same = (got.cpu() == want.cpu()).all(1).float().mean()
    assert float(same) >= 0.99, float(same)


# ------------------------------------------------------------------------------------------------ continuous batching
def test_continuous_batching_matches_plain_generate():
    """BASELINE config 5 semantics at test scale: sequences finish at different steps, slots are refilled from the
    queue, outputs come back in input order and equal the per-prompt greedy result."""
    from opus_pllm_b200.model import build_from_state_dicts
    from opus_pllm_b200.scheduler import ContinuousBatcher
    cfg = SMALL
    kw = {k if k != "ffn_dim" else "ffn": v for k, v in cfg.items()}
    lw = synth.llama_weights(seed=2, peaked=True, device="cuda", **kw)
    esm_cfg = dict(n_layers=2, dim=128, n_heads=2, ffn_dim=512)
    model = build_from_state_dicts(lw, cfg, synth.esm2_weights(2, 128, 512), esm_cfg,
                                   synth.projector_weights(128, 256, 8 * cfg["dim"]),
                                   synth.projector_weights(128, 256, 8 * cfg["dim"]))
    n, new = 13, 24
    seqs = synth.proteins(n, 10, 60, seed=21)
    prompts = synth.prompt_ids(n, 40, vocab=cfg["vocab"], ragged=7, sentinel_at=5, seed=77)
    # reference: every prompt on its own through the plain call, no EOS
    free = [model.generate(p[None].cuda(), [s], do_sample=False, max_new_tokens=new)[0].cpu() for p, s in zip(prompts, seqs)]
    eos = sorted({int(free[0][5]), int(free[3][11]), int(free[7][2])})
    want = []
    for f in free:
        hit = [i for i, t in enumerate(f.tolist()) if t in eos]
        want.append(f[: hit[0] + 1] if hit else f)
    assert len({len(w) for w in want}) > 2          # genuinely ragged finish times
    cb = ContinuousBatcher(model, max_slots=4, round_steps=5)
    got = cb.generate(prompts, seqs, new, eos_ids=eos, pad_id=eos[0])
    assert len(got) == n
    same = sum(int(torch.equal(g, w)) for g, w in zip(got, want))
    assert same >= n - 0, [(g.tolist(), w.tolist()) for g, w in zip(got, want) if not torch.equal(g, w)][:2]
    # the allocator got every page back
    assert len(model.llama._alloc.free) == model.llama._alloc.num_blocks
    # sampling through the scheduler (the eval scripts' default decode): a nucleus that only holds the top token must
    # reproduce the greedy result; a real nucleus is reproducible per seed and moves with it
    got_s = cb.generate(prompts, seqs, new, eos_ids=eos, pad_id=eos[0], sampling=(1.0, 1e-6, 11))
    assert all(torch.equal(g, w) for g, w in zip(got_s, want))
    hot = [cb.generate(prompts, seqs, new, eos_ids=eos, pad_id=eos[0], sampling=(50.0, 1.0, sd)) for sd in (5, 5, 6)]   # peaked logits need a hot softmax
    assert all(torch.equal(a, b) for a, b in zip(hot[0], hot[1]))
    assert any(not torch.equal(a, b) for a, b in zip(hot[0], hot[2]))
    assert any(not torch.equal(a, w) for a, w in zip(hot[0], want))
    # rounds of one request must not repeat the same draws: with a flat nucleus the tokens of consecutive rounds differ
    long = max(hot[0], key=len).tolist()
    assert long[1:6] != long[6:11] or len(long) < 11
    assert len(model.llama._alloc.free) == model.llama._alloc.num_blocks


def test_continuous_batching_compacts_into_smaller_batch_tiers():
    """Once the queue is empty the live requests move into the smallest decode batch tier (32 / 64 / 128 / 256) that holds
    them; results stay those of the per-prompt greedy call, in input order, and every page returns to the allocator."""
    from opus_pllm_b200.model import build_from_state_dicts
    from opus_pllm_b200.scheduler import ContinuousBatcher
    cfg = SMALL
    kw = {k if k != "ffn_dim" else "ffn": v for k, v in cfg.items()}
    lw = synth.llama_weights(seed=2, peaked=True, device="cuda", **kw)
    esm_cfg = dict(n_layers=2, dim=128, n_heads=2, ffn_dim=512)
    pw = synth.projector_weights(128, 256, 8 * cfg["dim"])
    model = build_from_state_dicts(lw, cfg, synth.esm2_weights(2, 128, 512), esm_cfg, pw, pw)
    n = 90
    seqs = synth.proteins(n, 10, 40, seed=31)
    prompts = synth.prompt_ids(n, 30, vocab=cfg["vocab"], ragged=7, sentinel_at=5, seed=78)
    lens = [3 + (i * 7) % 30 for i in range(n)]                       # ragged per-request budgets: a long ramp-down
    ids = torch.stack([torch.cat([torch.full((30 - p.numel(),), 1), p]) for p in prompts]).cuda()
    full = model.generate(ids, seqs, attention_mask=ids != 1, pad_token_id=1, do_sample=False, max_new_tokens=max(lens)).cpu()
    cb = ContinuousBatcher(model, max_slots=70, round_steps=4)
    got = cb.generate(prompts, seqs, lens)
    assert cb.stats["compactions"] >= 2 and cb.stats["slots"] == 70      # 70 -> 64 -> 32
    assert all(torch.equal(g, full[i, : lens[i]]) for i, g in enumerate(got))
    assert len(model.llama._alloc.free) == model.llama._alloc.num_blocks


def test_two_contexts_two_threads_generate_concurrently():
    """the C ABI is re-entrant per context. Two host threads, each with its own opus_ctx, CUDA stream and model,
    generate at the same time (decode graphs, stream-K scratch and tunables are per context; error strings per thread);
    both reproduce what they produce alone, and a tunable set in one context does not leak into the other."""
    import ctypes as C
    import threading
    from opus_pllm_b200 import _lib as L
    from opus_pllm_b200.llama import B200Llama
    lib = L.load()
    cfg = SMALL
    kw = {k if k != "ffn_dim" else "ffn": v for k, v in cfg.items()}
    models, inputs, alone = [], [], []
    for i in range(2):
        w = synth.llama_weights(seed=20 + i, peaked=True, device="cuda", **kw)
        m = B200Llama(w, **cfg)
        lens = [20 + (j * (3 + i)) % 30 for j in range(24 + 8 * i)]
        cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        tok = torch.randint(0, cfg["vocab"], (int(cu[-1]),), generator=torch.Generator().manual_seed(40 + i))
        emb = w["model.embed_tokens.weight"][tok.cuda()].to(torch.bfloat16)
        models.append(m); inputs.append((emb, cu))
        alone.append(m.generate_packed(emb, cu, 16).cpu())
    torch.cuda.synchronize()
    out, err = [None, None], [None, None]

    def worker(i):
        try:
            ctx = C.c_void_p()
            L.check(lib.opus_ctx_create(C.byref(ctx)))
            L.check(lib.opus_ctx_set_current(ctx))
            if i == 1:
                L.check(lib.opus_set_tunable(b"decode_rope_fused", 0))      # this context only
            with torch.cuda.stream(torch.cuda.Stream()):
                res = [models[i].generate_packed(*inputs[i], 16) for _ in range(3)]
                torch.cuda.current_stream().synchronize()
            out[i] = [r.cpu() for r in res]
            L.check(lib.opus_release_graphs())
            L.check(lib.opus_ctx_set_current(None))
            L.check(lib.opus_ctx_destroy(ctx))
        except Exception as e:  # noqa: BLE001
            err[i] = e
    ts = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert err == [None, None], err
    for i in range(2):
        assert all(torch.equal(r, alone[i]) for r in out[i])
    # the default context still has its own (default) tunables and works afterwards
    assert torch.equal(models[0].generate_packed(*inputs[0], 16).cpu(), alone[0])


def test_opt_in_tunables_keep_the_results():
    """Measured-and-parked alternatives stay correct: split-KV decode attention (2 / 4 parts, last-arriver merge), L2
    lookahead and the epilogue warm-up pass in the swap-AB GEMM (bit-identical by construction), grouped-N raster of the
    CTA-pair kernel (bit-identical: same k order per tile)."""
    from opus_pllm_b200 import _lib as L, ops
    from opus_pllm_b200.llama import B200Llama
    lib = L.load()
    cfg = dict(n_layers=2, dim=4096, n_q_heads=32, n_kv_heads=8, head_dim=128, ffn_dim=14336, vocab=8192)
    w = synth.llama_weights(seed=4, peaked=True, device="cuda", **{k if k != "ffn_dim" else "ffn": v for k, v in cfg.items()})
    model = B200Llama(w, **cfg)
    lens = [150 + (i * 37) % 200 for i in range(24)]
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    tok = torch.randint(0, cfg["vocab"], (int(cu[-1]),), generator=torch.Generator().manual_seed(6))
    emb = w["model.embed_tokens.weight"][tok.cuda()].to(torch.bfloat16)
    base = model.generate_packed(emb, cu, 10)
    base_logits = model._ws_bufs["logits"][:24].float().clone()
    try:
        for name, val, exact in (("attn_split", 2, False), ("attn_split", 4, False), ("attn_split", -1, False),
                                 ("l2_ahead", 8, True), ("epi_warm", 1, True)):
            L.check(lib.opus_set_tunable(name.encode(), val))
            got = model.generate_packed(emb, cu, 10)
            logits = model._ws_bufs["logits"][:24].float().clone()
            L.check(lib.opus_set_tunable(name.encode(), 0))
            assert torch.equal(got, base), (name, val)
            if exact:
                assert torch.equal(logits, base_logits), (name, val)
            else:
                assert _cos(logits, base_logits) >= 0.9999, (name, val)
        x = (torch.randn(2048, 14336, device="cuda") * 0.05).bfloat16()
        wd = (torch.randn(4096, 14336, device="cuda") * 0.02).bfloat16()
        res = torch.randn(2048, 4096, device="cuda").bfloat16()
        ref = ops.gemm(x, wd, epilogue=L.EPI_RES_BF16, residual=res, transposed=False)
        L.check(lib.opus_set_tunable(b"group_n", 8))
        assert torch.equal(ops.gemm(x, wd, epilogue=L.EPI_RES_BF16, residual=res, transposed=False), ref)
    finally:
        for name in ("attn_split", "l2_ahead", "epi_warm", "group_n"):
            L.check(lib.opus_set_tunable(name.encode(), 0))


# ------------------------------------------------------------------------------------------------ loaders + eval driver
def test_load_pretrained_model_and_eval_driver(tmp_path):
    """A fake OPUS-PLLM release on disk (HF safetensors dir, peft adapter, switch .bin, Lightning ckpt, fair-esm .pt)
    -> load_pretrained_model -> the run_opus_ddp-shaped driver; equals the model assembled from the same tensors."""
    import json
    from types import SimpleNamespace
    from tests.test_loaders_cpu import CFG, ESM, ToyTokenizer, write_fake_release
    from opus_pllm_b200 import builder, eval_ddp
    from opus_pllm_b200.model import build_from_state_dicts
    rel = write_fake_release(str(tmp_path), CFG, ESM)
    tok = ToyTokenizer()
    _, model, ctx = builder.load_pretrained_model(
        rel["base"], rel["weights"], "Meta-Llama-3-tiny", load_4bit=True, switch_projector_type="mlp2x_gelu",
        cstp_path=builder.return_cstp_path(rel["weights"], "modality_encoder/modality_encoding_adapter.ckpt"),
        esm_path=rel["esm"], tokenizer=tok)
    assert ctx == 512 and model.config.eos_token_id == [2, 3]
    direct = build_from_state_dicts(rel["llama"], CFG, rel["esm_sd"], ESM, rel["proj"], rel["proj"], lora_sd=rel["lora"],
                                    lora_alpha=8.0, lora_r=4, eos_token_id=[2, 3])
    data = [{"input": s, "instruction": f"Describe protein number {i} please", "output": "gt"}
            for i, s in enumerate(synth.proteins(5, 12, 40, seed=9))]
    inp, outp = tmp_path / "function_test.json", tmp_path / "out.json"
    json.dump(data, open(inp, "w"))
    args = SimpleNamespace(input_path=str(inp), save_path=str(outp), temperature=0.0, top_p=0.7, num_beams=1,
                           max_new_tokens=6, max_new_tokens_fixed=True, batch_size=2, continuous_batching=False,
                           system_prompt="You are a helpful protein assistant.", load_8bit=False, load_4bit=False,
                           switch_projector_type="mlp2x_gelu", esm_path=None)
    res = eval_ddp.eval_model(args, tokenizer=tok, model=model)
    assert len(res) == 5 and json.load(open(outp)) == res
    args.continuous_batching = True
    res_cb = eval_ddp.eval_model(args, tokenizer=tok, model=model)
    assert res_cb == res
    # the reference's default decode (temperature 0.1, top_p 0.7) through the same driver: reproducible per --seed
    args.continuous_batching, args.temperature, args.seed = False, 0.1, 3
    res_s = eval_ddp.eval_model(args, tokenizer=tok, model=model)
    assert len(res_s) == 5 and res_s == eval_ddp.eval_model(args, tokenizer=tok, model=model)
    # same tokens from the model assembled directly from the tensors
    prompt = eval_ddp.build_prompt(data[0]["instruction"], args.system_prompt, str(inp))
    from opus_pllm_b200.mm_utils import tokenizer_seq_token
    ids = tokenizer_seq_token(prompt, tok, return_tensors="pt")[None].cuda()
    a = model.generate(ids, [data[0]["input"]], do_sample=False, max_new_tokens=6, pad_token_id=2)
    b = direct.generate(ids, [data[0]["input"]], do_sample=False, max_new_tokens=6, pad_token_id=2)
    assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------------ stop keywords
def test_stop_sequences_kernel_and_generate(tmp_path):
    """Device-side counterpart of KeywordsStoppingCriteria (mm_utils.py:43-75): a row finishes when its emitted tail equals
    a stop sequence; afterwards it emits pad, and the loop ends early once every row has finished."""
    from opus_pllm_b200 import _lib as L, ops
    from opus_pllm_b200.llama import B200Llama
    # kernel: hand-made histories
    out = torch.tensor([[5, 6, 7, 0], [9, 6, 7, 0], [6, 7, 8, 0], [1, 2, 7, 0]], dtype=torch.int32).cuda()
    fin = torch.tensor([0, 1, 0, 0], dtype=torch.int32).cuda()
    left = torch.tensor([3], dtype=torch.int32).cuda()
    seqs = torch.tensor([[6, 7], [8, -1]], dtype=torch.int32).cuda()
    lens = torch.tensor([2, 1], dtype=torch.int32).cuda()
    st = torch.cuda.current_stream().cuda_stream
    L.check(L.load().opus_stop_sequences(out.data_ptr(), 4, 4, 2, None, seqs.data_ptr(), lens.data_ptr(), 2, 2,
                                         fin.data_ptr(), left.data_ptr(), st))
    assert fin.tolist() == [1, 1, 1, 0] and int(left) == 1      # row 1 was already finished: not counted twice
    # generate: take a greedy run, then ask to stop at a 2-token sequence that row 0 emitted
    cfg = SMALL
    kw = {k if k != "ffn_dim" else "ffn": v for k, v in cfg.items()}
    lw = synth.llama_weights(seed=2, peaked=True, device="cuda", **kw)
    model = B200Llama(lw, **cfg)
    lens_p = [30, 21, 17]
    cu = np.concatenate([[0], np.cumsum(lens_p)]).astype(np.int32)
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(3, cfg["vocab"], (sum(lens_p),), generator=g)
    emb = lw["model.embed_tokens.weight"][ids.cuda()].to(torch.bfloat16)
    pad = 0
    base = model.generate_packed(emb, cu, 12, pad_id=pad)
    stop = base[0, 3:5].tolist()
    got = model.generate_packed(emb, cu, 12, pad_id=pad, stop_sequences=[stop])
    want = base.clone()
    for b in range(3):
        row = base[b].tolist()
        hits = [i for i in range(1, 12) if row[i - 1: i + 1] == stop]
        if hits:
            want[b, hits[0] + 1:] = pad
    assert torch.equal(got, want)
    assert (got[0, 5:] == pad).all() and torch.equal(got[0, :5], base[0, :5])
    assert torch.equal(model.generate_packed(emb, cu, 12, pad_id=pad, stop_sequences=[stop], use_graph=False), got)
    assert torch.equal(model.generate_packed(emb, cu, 12, pad_id=pad), base)       # and off again: state is rebuilt


# ------------------------------------------------------------------------------------------------ builder seams
def test_component_builder_seams_on_gpu(tmp_path):
    """`build_protein_encoder(ckpt)` / `build_protein_projector(path)` / `build_switch_projector(model_args)`:
    objects built from files behave like the ones assembled from the tensors; a CSTP checkpoint's
    `protein_model.model.*` tensors override the base ESM-2 weights (cstp_v3/modelling.py:22-31); and the embedding
    producer (scripts/generate_esm_embedding.py) returns the encoder's embeddings regardless of batching."""
    from types import SimpleNamespace
    from tests.test_loaders_cpu import CFG, ESM, write_fake_release
    from opus_pllm_b200 import builder, generate_esm_embedding as G
    from opus_pllm_b200.encoder import B200ProteinEncoder
    rel = write_fake_release(str(tmp_path), CFG, ESM)
    seqs = synth.proteins(6, 5, 70, seed=4)
    enc = builder.build_protein_encoder(None, esm_path=rel["esm"])
    direct = B200ProteinEncoder(rel["esm_sd"], **ESM)
    want = direct.get_protein_seq_embeddings(seqs)
    assert torch.equal(enc.get_protein_seq_embeddings(seqs), want)
    # CSTP checkpoint override of one layer's fc1 (+ an unknown key that must be ignored)
    over = {k: v.clone() for k, v in rel["esm_sd"].items()}
    over["layers.0.fc1.weight"] = over["layers.0.fc1.weight"] * 0.5
    ck = tmp_path / "cstp_train.ckpt"
    torch.save({"model": {"protein_model.model.layers.0.fc1.weight": over["layers.0.fc1.weight"],
                          "protein_model.model.contact_head.regression.weight": torch.zeros(1, 4),
                          "text_model.whatever": torch.zeros(2)}}, ck)
    enc2 = builder.build_protein_encoder(str(ck), esm_path=rel["esm"])
    got2 = enc2.get_protein_seq_embeddings(seqs)
    assert torch.equal(got2, B200ProteinEncoder(over, **ESM).get_protein_seq_embeddings(seqs))
    assert not torch.equal(got2, want)
    # projector seams
    pp = builder.build_protein_projector(builder.return_cstp_path(rel["weights"], "modality_encoder/modality_encoding_adapter.ckpt"))
    c = pp.to("cuda").protein_forward(want)
    assert c.shape == (6, 256) and _cos(c, mm_ref.protein_forward(want, rel["proj"]["protein_projection.linear.weight"].cuda(),
                                                                  rel["proj"]["protein_projection.linear.bias"].cuda())) >= 0.999
    # producer: one protein per call (the reference's loop) == packed, length-sorted batches
    emb = G.embed_all(enc, seqs + seqs[:2], max_tokens=100)
    assert set(emb) == set(seqs)
    for s in seqs:
        one = enc.get_protein_seq_embeddings([s])[0].cpu()
        assert _cos(torch.tensor(emb[s]), one) >= 0.99999 and float((torch.tensor(emb[s]) - one).abs().max()) <= 2e-2


@pytest.mark.parametrize("batch", [160, 256])
def test_decode_batch_above_128_pair_kernel_is_bit_identical(batch):
    """Decode at batch 129..256: the q|k|v, o_proj and down split-K GEMMs run on the CTA-pair swap-AB kernel
    (`gemm_2cta_tr`). With the pair kernel off the same launches go to the single-CTA kernel: tokens, logits and the KV
    cache must not change by a bit (Llama-3-8B-wide layers, reduced depth, ragged prompts)."""
    from opus_pllm_b200 import _lib as L
    from opus_pllm_b200.llama import B200Llama
    cfg = dict(n_layers=2, dim=4096, n_q_heads=32, n_kv_heads=8, head_dim=128, ffn_dim=14336, vocab=8192)
    w = synth.llama_weights(seed=13, device="cuda", dtype=torch.bfloat16,
                            **{k if k != "ffn_dim" else "ffn": v for k, v in cfg.items()})
    model = B200Llama(w, **cfg)
    lens = [9 + (i * 7) % 40 for i in range(batch)]
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    emb = synth.weight((int(cu[-1]), cfg["dim"]), "pair_decode_prompt", 0.02).cuda().to(torch.bfloat16)
    lib = L.load()
    plan = model.make_plan(cu, 6)
    try:
        res = {}
        for mode in (1, 0):
            L.check(lib.opus_set_tunable(b"gemm_2cta_tr", mode))
            model._k.zero_(); model._v.zero_()
            st = model.prefill(emb, plan=plan)
            toks = model.generate_from_prefill(st, 6)
            res[mode] = (toks.clone(), model._ws_bufs["logits"][:batch].clone(), model._k.clone(), model._v.clone())
    finally:
        L.check(lib.opus_set_tunable(b"gemm_2cta_tr", 2))
        model.release_plan(plan)
    for a, b in zip(res[1], res[0]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("batch", [300, 512])
def test_decode_batch_257_to_512_vs_oracle(batch):
    """Decode at batch 257..512 (C3's 512 prompts per GPU as one batch): two 256-wide batch tiles per weight tile on the
    CTA-pair swap-AB kernel, split-K chosen for whole waves of pairs (`gemm_pick_split_k_wide`), gate/up with a two-tile
    stream-K tail. Llama-3-8B-wide layers at reduced depth, peaked logits, ragged prompts: tokens against the oracle, and
    against the single-CTA kernels (pair kernel off)."""
    from opus_pllm_b200 import _lib as L
    from opus_pllm_b200.llama import B200Llama
    cfg = dict(n_layers=2, dim=4096, n_q_heads=32, n_kv_heads=8, head_dim=128, ffn_dim=14336, vocab=8192)
    w = synth.llama_weights(seed=17, peaked=True, device="cuda", dtype=torch.bfloat16,
                            **{k if k != "ffn_dim" else "ffn": v for k, v in cfg.items()})
    model = B200Llama(w, **cfg)
    lens = [5 + (i * 11) % 28 for i in range(batch)]
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    tok = torch.randint(0, cfg["vocab"], (int(cu[-1]),), generator=torch.Generator().manual_seed(6))
    emb = w["model.embed_tokens.weight"][tok.cuda()].to(torch.bfloat16)
    lib = L.load()
    n_new = 8
    try:
        got = model.generate_packed(emb, cu, n_new)
        assert torch.equal(got, model.generate_packed(emb, cu, n_new, use_graph=False))
        L.check(lib.opus_set_tunable(b"gemm_2cta_tr", 0))
        single = model.generate_packed(emb, cu, n_new)
    finally:
        L.check(lib.opus_set_tunable(b"gemm_2cta_tr", 2))
    e_pad, m_pad = _padded(emb, cu, cfg["dim"])
    want = llama_ref.greedy_generate(_dev(w, torch.bfloat16), _oracle_cfg(cfg), e_pad, m_pad, n_new)
    assert got.shape == (batch, n_new)
    assert float((got.cpu() == want.cpu()).all(1).float().mean()) >= 0.99
    assert float((got == single).all(1).float().mean()) >= 0.99
    assert len(torch.unique(got)) > 32


def test_initialize_protein_modules_rebuilds_the_seams(tmp_path):
    """opus_arch.py:46-90: the model's protein modules (re)built from `model_args` through the builder seams give the
    same generation as the model assembled by load_pretrained_model from the same files."""
    from types import SimpleNamespace
    from tests.test_loaders_cpu import CFG, ESM, ToyTokenizer, write_fake_release
    from opus_pllm_b200 import builder
    from opus_pllm_b200.model import build_from_state_dicts
    rel = write_fake_release(str(tmp_path), CFG, ESM)
    cstp = builder.return_cstp_path(rel["weights"], "modality_encoder/modality_encoding_adapter.ckpt")
    sw = builder.return_cstp_path(rel["weights"], "modality_refinement_projector/modality_refinement_projection.bin")
    _, ref_model, _ = builder.load_pretrained_model(rel["base"], rel["weights"], "Meta-Llama-3-tiny", cstp_path=cstp,
                                                    esm_path=rel["esm"], tokenizer=ToyTokenizer())
    bare = build_from_state_dicts(rel["llama"], CFG, None, None, None, None, lora_sd=rel["lora"], lora_alpha=8.0, lora_r=4,
                                  eos_token_id=[2, 3])
    assert bare.get_protein_encoder() is None
    bare.get_model().initialize_protein_modules(SimpleNamespace(
        device="cuda", has_protein_encoder=True, has_switch_projector=True, esm_ckpt=None, esm_path=rel["esm"],
        pretrain_protein_projector_ckpt=cstp, pretrain_switch_projector_ckpt=sw, switch_projector_type="mlp2x_gelu",
        hidden_size=CFG["dim"]), fsdp=None)
    seqs = synth.proteins(3, 10, 50, seed=2)
    ids = torch.stack(synth.prompt_ids(3, 20, vocab=CFG["vocab"], sentinel_at=5)).cuda()
    a = ref_model.generate(ids, seqs, do_sample=False, max_new_tokens=5, pad_token_id=2)
    b = bare.generate(ids, seqs, do_sample=False, max_new_tokens=5, pad_token_id=2)
    assert torch.equal(a, b)
