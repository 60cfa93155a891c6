"""CPU tests of the weight-file readers (builder.py) and prompt helpers against the reference's semantics."""
import json
import os

import torch
from safetensors.torch import save_file

from opus_pllm_b200 import builder, mm_utils, synth
from oracle import esm2_ref


def _write_llama_dir(d, cfg, sd):
    os.makedirs(d, exist_ok=True)
    json.dump(dict(num_hidden_layers=cfg["n_layers"], hidden_size=cfg["dim"], num_attention_heads=cfg["n_q_heads"],
                   num_key_value_heads=cfg["n_kv_heads"], head_dim=cfg["head_dim"], intermediate_size=cfg["ffn_dim"],
                   vocab_size=cfg["vocab"], rms_norm_eps=1e-5, rope_theta=500000.0, eos_token_id=[2, 3],
                   model_type="llama"), open(os.path.join(d, "config.json"), "w"))
    keys = sorted(sd)
    save_file({k: sd[k].contiguous() for k in keys[: len(keys) // 2]}, os.path.join(d, "model-00001-of-00002.safetensors"))
    save_file({k: sd[k].contiguous() for k in keys[len(keys) // 2:]}, os.path.join(d, "model-00002-of-00002.safetensors"))


def write_fake_release(root, cfg, esm_cfg, seed=3):
    """Lay out tiny synthetic weights exactly like an OPUS-PLLM release + its Llama base dir."""
    lw = synth.llama_weights(cfg["n_layers"], cfg["dim"], cfg["n_q_heads"], cfg["n_kv_heads"], cfg["head_dim"],
                             cfg["ffn_dim"], cfg["vocab"], seed=seed, peaked=True)
    base = os.path.join(root, "Meta-Llama-3-tiny")
    _write_llama_dir(base, cfg, lw)
    w = os.path.join(root, "opus_weights")
    os.makedirs(os.path.join(w, "lora_adapter")); os.makedirs(os.path.join(w, "modality_refinement_projector"))
    os.makedirs(os.path.join(w, "modality_encoder"))
    lora = synth.lora_adapters(lw, cfg["n_layers"], r=4, seed=seed)
    json.dump(dict(r=4, lora_alpha=8, target_modules=["q_proj", "v_proj"], fan_in_fan_out=False, peft_type="LORA"),
              open(os.path.join(w, "lora_adapter", "adapter_config.json"), "w"))
    save_file({"base_model.model." + k: v.contiguous() for k, v in lora.items()},
              os.path.join(w, "lora_adapter", "adapter_model.safetensors"))
    pw = synth.projector_weights(esm_cfg["dim"], 256, 8 * cfg["dim"], seed=seed)
    torch.save({"model.switch_projector." + k: pw[k] for k in ("0.weight", "0.bias", "2.weight", "2.bias")},
               os.path.join(w, "modality_refinement_projector", "modality_refinement_projection.bin"))
    torch.save({"state_dict": {"protein_projection.linear.weight": pw["protein_projection.linear.weight"],
                               "protein_projection.linear.bias": pw["protein_projection.linear.bias"],
                               "text_projection.linear.weight": torch.zeros(2, 2)}, "epoch": 1},
               os.path.join(w, "modality_encoder", "modality_encoding_adapter.ckpt"))
    ew = synth.esm2_weights(esm_cfg["n_layers"], esm_cfg["dim"], esm_cfg["ffn_dim"], seed=seed)
    esm_pt = os.path.join(root, "esm2_tiny.pt")
    torch.save({"model": {"encoder.sentence_encoder." + k: v for k, v in ew.items()}, "cfg": {}}, esm_pt)
    return dict(base=base, weights=w, esm=esm_pt, llama=lw, lora=lora, proj=pw, esm_sd=ew)


CFG = dict(n_layers=2, dim=256, n_q_heads=2, n_kv_heads=1, head_dim=128, ffn_dim=512, vocab=512)
ESM = dict(n_layers=2, dim=128, n_heads=2, ffn_dim=256)


def test_readers_roundtrip(tmp_path):
    rel = write_fake_release(str(tmp_path), CFG, ESM)
    sd, kw = builder.read_hf_llama(rel["base"])
    extra = kw.pop("_extra")
    assert kw == dict(CFG, rms_eps=1e-5, rope_theta=500000.0) and extra["eos_token_id"] == [2, 3]
    assert set(sd) == set(rel["llama"]) and all(torch.equal(sd[k], rel["llama"][k]) for k in sd)
    lora, alpha, r = builder.read_peft_lora(os.path.join(rel["weights"], "lora_adapter"))
    assert (alpha, r) == (8.0, 4) and set(lora) == set(rel["lora"])
    sw = builder.read_switch_projector(os.path.join(rel["weights"], "modality_refinement_projector",
                                                    "modality_refinement_projection.bin"))
    assert set(sw) == {"0.weight", "0.bias", "2.weight", "2.bias"} and torch.equal(sw["2.bias"], rel["proj"]["2.bias"])
    cs = builder.read_cstp_checkpoint(builder.return_cstp_path(rel["weights"], "modality_encoder/modality_encoding_adapter.ckpt"))
    assert set(cs) == {"protein_projection.linear.weight", "protein_projection.linear.bias"}
    esd, ecfg = builder.read_esm2(rel["esm"])
    assert ecfg == ESM and set(esd) == set(rel["esm_sd"])


def test_read_esm2_accepts_hf_names(tmp_path):
    ew = synth.esm2_weights(2, 128, 256, seed=1)
    hf = {"esm." + k: v.contiguous() for k, v in esm2_ref.to_hf_esm_state_dict(ew, 2).items()}
    d = tmp_path / "hf_esm"
    d.mkdir()
    save_file(hf, str(d / "model.safetensors"))
    sd, cfg = builder.read_esm2(str(d))
    assert cfg == dict(n_layers=2, dim=128, n_heads=2, ffn_dim=256)
    assert set(sd) == set(ew) and all(torch.equal(sd[k], ew[k]) for k in ew)


class ToyTokenizer:
    """whitespace tokenizer with a BOS, enough for the prompt-helper semantics"""
    bos_token_id, eos_token_id, pad_token_id = 1, 2, 2

    def __call__(self, text):
        from types import SimpleNamespace
        return SimpleNamespace(input_ids=[self.bos_token_id] + [3 + (hash(w) % 400) for w in text.split()])

    def batch_decode(self, ids, skip_special_tokens=True):
        return [" ".join(f"t{int(t)}" for t in row if int(t) not in (1, 2)) for row in ids]


def test_prompt_helpers_match_reference_semantics():
    tok = ToyTokenizer()
    ids = mm_utils.tokenizer_seq_token("a b <seq>\nc d <seq> e", tok)
    assert ids[0] == 1 and ids.count(-200) == 2 and ids.count(1) == 1
    a, b = tok("a b").input_ids[1:], tok("\nc d").input_ids[1:]
    assert ids[1:3] == a and ids[3] == -200 and ids[4:6] == b
    padded = mm_utils.left_pad_sequence([torch.tensor([5, 6, 7]), torch.tensor([8])], 2, batch_first=True)
    assert padded.tolist() == [[5, 6, 7], [2, 2, 8]]
    assert mm_utils.after_process_output("  Nucleus ### Student: next", "###") == "Nucleus"
    assert mm_utils.after_process_output("Cytoplasm", "###") == "Cytoplasm"
    assert mm_utils.get_model_name_from_path("/x/llama3/checkpoint-12/") == "llama3_checkpoint-12"
    from opus_pllm_b200 import eval_ddp
    p = eval_ddp.build_prompt("Where is it?", "SYS", "data/localization.json")
    assert p == "SYS\n\n### Student: <seq>\nWhere is it?Kindly reply with only one word.\n### Professor:"
    assert eval_ddp.max_new_tokens_for("x/keywords.json", 32) == 128 and eval_ddp.max_new_tokens_for("x/f.json", 32) == 256


def test_read_hf_opt_directory_and_family_dispatch(tmp_path):
    """OPT / Galactica base dirs (model/builder.py:71-82: `use_safetensors=False`, i.e. pytorch_model*.bin)."""
    import pytest
    d = str(tmp_path / "opt-tiny")
    os.makedirs(d)
    sd = synth.opt_weights(2, 256, 2, 512, 512, max_pos=64, seed=4)
    json.dump(dict(num_hidden_layers=2, hidden_size=256, num_attention_heads=2, ffn_dim=512, vocab_size=512,
                   max_position_embeddings=64, word_embed_proj_dim=256, do_layer_norm_before=True,
                   activation_function="relu", eos_token_id=2, model_type="opt"), open(os.path.join(d, "config.json"), "w"))
    torch.save({k[len("model."):]: v for k, v in sd.items() if k != "lm_head.weight"},    # old-style `decoder.*` keys
               os.path.join(d, "pytorch_model.bin"))
    got, kw = builder.read_hf_opt(d)
    extra = kw.pop("_extra")
    assert kw == dict(n_layers=2, dim=256, n_heads=2, ffn_dim=512, vocab=512, max_pos=64, activation="relu")
    assert extra["eos_token_id"] == 2
    assert torch.equal(got["decoder.layers.1.fc1.weight"], sd["model.decoder.layers.1.fc1.weight"])
    cfg = json.load(open(os.path.join(d, "config.json")))
    cfg.update(do_layer_norm_before=False, word_embed_proj_dim=128)                       # opt-350m
    json.dump(cfg, open(os.path.join(d, "config.json"), "w"))
    with pytest.raises(NotImplementedError):
        builder.read_hf_opt(d)
    with pytest.raises(NotImplementedError):                                              # builder.py:95-96
        builder.load_pretrained_model("/nonexistent/mistral-7b", None, "mistral-7b")   # (tmp_path itself contains "opt")


def test_component_builder_seams_and_embedding_producer_host_logic(tmp_path):
    """The reference's three builder seams keep their names / arguments, and the embedding producer keeps the
    scripts' file formats (scripts/generate_esm_embedding.py, generate_esm_for_each_seq.py). Host logic only."""
    import inspect
    from types import SimpleNamespace
    from opus_pllm_b200 import generate_esm_embedding as G
    assert list(inspect.signature(builder.build_protein_encoder).parameters)[0] == "ckpt"
    assert list(inspect.signature(builder.build_protein_projector).parameters)[0] == "cstp_chackpoint_path"
    assert list(inspect.signature(builder.build_switch_projector).parameters)[:2] == ["model_args", "n_tokens"]
    sw = builder.build_switch_projector(SimpleNamespace(hidden_size=64, pretrain_protein_projector_ckpt="x"), device="cpu")
    assert (sw.in_dim, sw.hidden_dim, sw.depth) == (5120, 512, 2)
    sw = builder.build_switch_projector(SimpleNamespace(hidden_size=64, pretrain_protein_projector_ckpt=None,
                                                        switch_projector_type="linear"), n_tokens=4, device="cpu")
    assert (sw.in_dim, sw.hidden_dim, sw.depth) == (1280, 256, 1)
    assert builder.build_switch_projector(SimpleNamespace(hidden_size=64, pretrain_protein_projector_ckpt=None,
                                                          switch_projector_type="conv"), device="cpu") is None
    # batching: longest first, token budget respected, every index exactly once
    seqs = ["A" * n for n in (5, 300, 12, 298, 7, 4000)]
    bs = G.batches_by_length(seqs, 620)
    assert sorted(i for b in bs for i in b) == list(range(6)) and bs[0] == [5]
    assert all(sum(len(seqs[i]) + 2 for i in b) <= 620 or len(b) == 1 for b in bs)

    class FakeEncoder:                                           # embedding = [len, #A, 0, ...] so results are checkable
        calls = 0

        def get_protein_seq_embeddings(self, data):
            FakeEncoder.calls += 1
            return torch.tensor([[float(len(s)), float(s.count("A"))] + [0.0] * 2 for s in data])

    recs = [dict(instruction="i", input=s, output="o", extra=1) for s in ("MKA", "AAAA", "MKA", "G" * 4001, "K" * 4000)]
    src, dct = tmp_path / "d.json", tmp_path / "known.json"
    json.dump(recs, open(src, "w"))
    json.dump({"AAAA": [9.0, 9.0, 9.0, 9.0]}, open(dct, "w"))
    args = SimpleNamespace(file_path=str(src), save_path=str(tmp_path / "out.jsonl"), dict_path=str(dct), esm_path=None,
                           ckpt=None, max_tokens=4096, dict_only=False)
    assert G.generate_esm_embedding(args, FakeEncoder()) == 4   # the 4001-residue protein is skipped (> 4000)
    lines = [json.loads(l) for l in open(args.save_path)]
    assert [l["input"] for l in lines] == ["MKA", "AAAA", "MKA", "K" * 4000]
    assert lines[0]["input_embed"] == [3.0, 1.0, 0.0, 0.0] and lines[1]["input_embed"] == [9.0] * 4   # dict entry reused
    assert set(lines[0]) == {"instruction", "input", "output", "input_embed"}
    args.dict_only, args.save_path = True, str(tmp_path / "seq2embed.json")
    assert G.generate_esm_embedding(args, FakeEncoder()) == 2   # unique sequences shorter than 4000
    assert set(json.load(open(args.save_path))) == {"MKA", "AAAA"}
