"""GPU: BASELINE-sized models (ESM-2-650M, Llama-3-8B shapes, random init) checked through size-independent properties,
plus reduced-scale versions of the other BASELINE configs (C3 large-batch decode, C4 long-protein varlen encoder)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from opus_pllm_b200 import presets, synth  # noqa: E402
from oracle import esm2_ref, llama_ref  # noqa: E402


@pytest.fixture(scope="module")
def full_model():
    return presets.build_synthetic_model("full", "cuda", peaked=True, with_lora=True)


def _ids(n, length, vocab, seed):
    return torch.stack(synth.prompt_ids(n, length, vocab=vocab, seed=seed)).cuda()


def test_fullsize_generate_permutation_prefix_and_graph_invariance(full_model):
    m = full_model
    B = 12
    seqs = synth.proteins(B, 120, 300, seed=31)
    ids = _ids(B, 96, 128256, 32)
    out = m.generate(ids, seqs, do_sample=False, max_new_tokens=8, pad_token_id=128001)
    assert out.shape == (B, 8) and out.dtype == torch.int64 and int(out.min()) >= 0 and int(out.max()) < 128256
    # idempotence + CUDA-graph replay == plain launches
    assert torch.equal(out, m.generate(ids, seqs, do_sample=False, max_new_tokens=8, pad_token_id=128001))
    assert torch.equal(out, m.generate(ids, seqs, do_sample=False, max_new_tokens=8, pad_token_id=128001,
                                       use_graph=False))
    # prefix property: fewer new tokens = a prefix of the longer generation
    assert torch.equal(out[:, :3], m.generate(ids, seqs, do_sample=False, max_new_tokens=3, pad_token_id=128001))
    # permutation equivariance over the prompt order (prompts are independent rows of the batch)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0))
    outp = m.generate(ids[perm.cuda()], [seqs[i] for i in perm.tolist()], do_sample=False, max_new_tokens=8,
                      pad_token_id=128001)
    assert torch.equal(outp, out[perm.cuda()])
    # the protein matters: swapping the proteins changes the soft tokens the LLM sees
    soft_a = m._soft_tokens(seqs[:2], None)
    soft_b = m._soft_tokens(seqs[2:4], None)
    assert soft_a.shape == (2, 8, 4096) and not torch.allclose(soft_a.float(), soft_b.float())


def test_fullsize_encoder_batch_invariance_and_statistics(full_model):
    enc = full_model.protein_encoder
    seqs = synth.proteins(6, 40, 500, seed=33)
    together = enc.get_protein_seq_embeddings(seqs)
    assert together.shape == (6, 1280) and together.dtype == torch.float32 and bool(torch.isfinite(together).all())
    alone = torch.cat([enc.get_protein_seq_embeddings([s]) for s in seqs])
    assert torch.allclose(together, alone, rtol=0, atol=5e-3), float((together - alone).abs().max())
    # reversing a protein changes its embedding (positions matter), duplicating it does not
    rev = enc.get_protein_seq_embeddings([seqs[0][::-1], seqs[0], seqs[0]])
    assert float((rev[0] - rev[1]).abs().max()) > 1e-2 and torch.equal(rev[1], rev[2])


def test_c4_long_protein_varlen_encoder_vs_oracle():
    """BASELINE config 4 at reduced depth/batch: lengths 1024-2048, packed varlen, positions beyond 1024."""
    from opus_pllm_b200.encoder import B200ProteinEncoder
    n_layers, dim, heads, ffn = 2, 1280, 20, 5120
    w = synth.esm2_weights(n_layers, dim, ffn, seed=8, device="cuda")
    enc = B200ProteinEncoder(w, n_layers, dim, heads, ffn)
    seqs = synth.proteins(3, 1024, 2048, seed=44)
    got = enc.get_protein_seq_embeddings(seqs)
    want = torch.cat([esm2_ref.get_protein_seq_embeddings(w, [s], n_layers, heads) for s in seqs])
    cos = torch.nn.functional.cosine_similarity(got, want, dim=-1)
    assert float(cos.min()) >= 0.9995 and float((got - want).abs().max()) <= 3e-2


def test_c3_large_batch_decode_vs_oracle():
    """BASELINE config 3 at reduced size: 256 prompts per GPU in one decode batch (swap-AB tiles with N = 256)."""
    from opus_pllm_b200.llama import B200Llama
    cfg = dict(n_layers=2, dim=512, n_q_heads=4, n_kv_heads=2, head_dim=128, ffn_dim=1024, vocab=2048)
    w = synth.llama_weights(seed=2, peaked=True, device="cuda", **{k if k != "ffn_dim" else "ffn": v for k, v in cfg.items()})
    model = B200Llama(w, **cfg)
    B = 256
    lens = [16 + (i * 5) % 40 for i in range(B)]
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    tok = torch.randint(0, cfg["vocab"], (int(cu[-1]),), generator=torch.Generator().manual_seed(9))
    emb = w["model.embed_tokens.weight"][tok.cuda()].to(torch.bfloat16)
    got = model.generate_packed(emb, cu, 12)
    Lm = max(lens)
    e = torch.zeros(B, Lm, cfg["dim"], dtype=torch.bfloat16, device="cuda")
    m = torch.zeros(B, Lm, dtype=torch.bool, device="cuda")
    for b in range(B):
        e[b, Lm - lens[b]:] = emb[cu[b]: cu[b + 1]]
        m[b, Lm - lens[b]:] = True
    ocfg = llama_ref.LlamaCfg(n_layers=2, dim=512, n_q_heads=4, n_kv_heads=2, head_dim=128, ffn_dim=1024, vocab=2048)
    want = llama_ref.greedy_generate({k: v.to(torch.bfloat16) for k, v in w.items()}, ocfg, e, m, 12)
    same = (got.cpu() == want.cpu()).all(1).float().mean()
    assert float(same) >= 0.99, float(same)
