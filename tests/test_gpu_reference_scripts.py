"""GPU: the reference's OWN eval scripts, unmodified, running on this backend (north star: "the eval scripts stay drop-in
compatible").

The scripts are executed from baseline/_ref/ — the three scripts and the helper modules they import, staged there
unmodified by oracle/stage_reference.py
(git-ignored, travels to the GPU box with the snapshot; /root/reference does not exist there). `opus_pllm_b200.compat`
registers this package as `multi_modality_model.multi_modality_v1.model.builder`, so the scripts' own import statement
yields this backend's `load_pretrained_model`; everything else they import from the reference (constants, conversation
templates, mm_utils.tokenizer_seq_token, utils) is the reference's code. The model is a fake OPUS-PLLM release on disk
(HF safetensors dir with a real HF tokenizer, peft adapter, switch-projector .bin, Lightning ckpt, fair-esm .pt).

  eval/run_opus_ddp.py:47-148          -> JSON equal to opus_pllm_b200.eval_ddp's (temperature 0), batches of 8
  eval/eval_run_multichoice.py:52-219  -> chat template + conv_vicuna_v3, answers equal to direct generate() calls
  eval/run_opus_online.py:16-102       -> interactive loop, bs 1, attention_mask=None, with and without a protein
"""
import builtins
import json
import os
import sys
from types import SimpleNamespace

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
EVAL = os.path.join(REF, "multi_modality_model", "multi_modality_v1", "eval")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.isfile(os.path.join(EVAL, "run_opus_ddp.py")),
                                 reason="baseline/_ref is not staged (python oracle/stage_reference.py where /root/reference exists)")]

from opus_pllm_b200 import synth  # noqa: E402
from tests.test_loaders_cpu import CFG, ESM, write_fake_release  # noqa: E402


def write_byte_tokenizer(model_dir: str, vocab_size: int):
    """A real HF fast tokenizer (byte-level, no merges) saved next to the fake Llama weights: every id < vocab_size decodes."""
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers
    from tokenizers.processors import TemplateProcessing
    vocab = {"<unk>": 0, "<s>": 1, "</s>": 2}
    for ch in sorted(pre_tokenizers.ByteLevel.alphabet()):
        vocab[ch] = len(vocab)
    tok = Tokenizer(models.BPE(vocab=vocab, merges=[], unk_token="<unk>"))
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)
    tok.decoder = decoders.ByteLevel()
    tok.post_processor = TemplateProcessing(single="<s> $A", special_tokens=[("<s>", 1)])
    tok.add_tokens([f"<x{i}>" for i in range(len(vocab), vocab_size)])
    tok.save(os.path.join(model_dir, "tokenizer.json"))
    json.dump({"tokenizer_class": "PreTrainedTokenizerFast", "bos_token": "<s>", "eos_token": "</s>", "unk_token": "<unk>"},
              open(os.path.join(model_dir, "tokenizer_config.json"), "w"))


@pytest.fixture(scope="module")
def release(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("release"))
    rel = write_fake_release(root, CFG, ESM)
    write_byte_tokenizer(rel["base"], CFG["vocab"])
    json.dump({"eos_token_id": [2, 7]}, open(os.path.join(rel["base"], "generation_config.json"), "w"))
    old_env = os.environ.get("OPUS_ESM_PATH")
    os.environ["OPUS_ESM_PATH"] = rel["esm"]
    saved_modules = dict(sys.modules)
    sys.path.insert(0, REF)
    sys.path.insert(0, EVAL)                 # the scripts are run from their own directory (`import metrics_computing_opi`)
    from opus_pllm_b200 import compat
    compat.install(force_stubs=True)
    yield rel
    sys.path.remove(REF); sys.path.remove(EVAL)
    for k in set(sys.modules) - set(saved_modules):
        if k.startswith(("multi_modality_model", "accelerate", "metrics_computing_opi")):
            del sys.modules[k]
    if old_env is None:
        os.environ.pop("OPUS_ESM_PATH", None)
    else:
        os.environ["OPUS_ESM_PATH"] = old_env


def _args(rel, **kw):
    base = dict(model_base_path=rel["base"], opus_pllm_weights_path=rel["weights"], is_json=True, temperature=0.0,
                top_p=0.7, num_beams=1, max_new_tokens=8, switch_projector_type="mlp2x_gelu", load_4bit=True,
                load_8bit=False)
    base.update(kw)
    return SimpleNamespace(**base)


def test_compat_registers_the_backend_under_the_reference_module_name(release):
    from multi_modality_model.multi_modality_v1.model.builder import load_pretrained_model, return_cstp_path
    from opus_pllm_b200 import builder
    assert load_pretrained_model is builder.load_pretrained_model and return_cstp_path is builder.return_cstp_path
    # the host-side helpers still come from the unmodified reference tree
    import multi_modality_model.multi_modality_v1.mm_utils as ref_mm
    assert os.path.abspath(ref_mm.__file__).startswith(os.path.abspath(REF))


def test_run_opus_ddp_unmodified_equals_eval_ddp(release, tmp_path):
    from opus_pllm_b200 import compat, eval_ddp
    script = compat.load_script(os.path.join(EVAL, "run_opus_ddp.py"), "ref_run_opus_ddp")
    prots = synth.proteins(11, 12, 60, seed=9)
    # 11 entries = one full batch of 8 + a ragged one; the first instructions carry <seq> themselves (user's
    # max_new_tokens applies), a later one does not (the script then forces the localization value for the rest of the run)
    data = [{"input": s, "output": f"gt{i}",
             "instruction": (f"<seq>\nWhere does protein {i} live?" if i < 9 else f"Where does protein {i} live?")}
            for i, s in enumerate(prots)]
    data.insert(4, {"input": None, "instruction": "dropped", "output": "x"})      # run_opus_ddp.py:65 filters these
    inp = tmp_path / "localization_test.json"
    json.dump(data, open(inp, "w"))
    a = _args(release, input_path=str(inp), save_path=str(tmp_path / "ref.json"))
    script.eval_model(a)
    ref = json.load(open(a.save_path))
    assert len(ref) == 11 and [r["ground_truth"] for r in ref] == [d["output"] for d in data if d["input"] is not None]
    assert a.max_new_tokens == 32                                        # the script mutated it (run_opus_ddp.py:94)
    b = _args(release, input_path=str(inp), save_path=str(tmp_path / "b200.json"), max_new_tokens_fixed=False,
              batch_size=8, continuous_batching=False, system_prompt=None, esm_path=None, seed=0, stop_keyword=False)
    ours = eval_ddp.eval_model(b)
    assert ours == ref and json.load(open(b.save_path)) == ref
    assert len({r["generated"] for r in ref}) > 1
    # continuous batching through the same driver returns the same answers
    b.continuous_batching, b.save_path = True, str(tmp_path / "b200_cb.json")
    assert eval_ddp.eval_model(b) == ref
    # sampled decode (the script's default temperature 0.1 / top_p 0.7) runs end to end too
    a2 = _args(release, input_path=str(inp), save_path=str(tmp_path / "ref_s.json"), temperature=0.1)
    script.eval_model(a2)
    assert len(json.load(open(a2.save_path))) == 11


def test_eval_run_multichoice_unmodified(release, tmp_path, capsys):
    from opus_pllm_b200 import builder, compat
    script = compat.load_script(os.path.join(EVAL, "eval_run_multichoice.py"), "ref_eval_run_multichoice")
    prots = synth.proteins(10, 15, 50, seed=4)
    data = [{"question": f"Which compartment holds protein {i}?", "options": ["A) nucleus", "B) cytosol", "C) membrane", "D) none"],
             "input": s, "answer": "B"} for i, s in enumerate(prots)]
    inp = tmp_path / "mc.json"
    json.dump(data, open(inp, "w"))
    a = _args(release, input_path=str(inp), save_path=str(tmp_path / "mc_out.json"), max_new_tokens=6)
    script.eval_model(a)
    got = json.load(open(a.save_path))
    assert len(got) == 10 and all(g["ground_truth"] == "B" for g in got)
    assert "Accuracy" in capsys.readouterr().out
    # the same prompts through direct generate() calls (chat template + conv_vicuna_v3 exactly as the script builds them)
    import multi_modality_model.multi_modality_v1.conversation as conv_lib
    from multi_modality_model.multi_modality_v1.mm_utils import tokenizer_seq_token
    tok, model, _ = builder.load_pretrained_model(release["base"], release["weights"], "Meta-Llama-3-tiny",
                                                  cstp_path=builder.return_cstp_path(release["weights"], "modality_encoder/modality_encoding_adapter.ckpt"))
    assert model.config.eos_token_id == [2, 7, 3]            # generation_config.json first, then config.json's extra id
    # the script installs its own template on tokenizers that have none (eval_run_multichoice.py:59-72): same string here
    tok.chat_template = next(c for c in script.eval_model.__code__.co_consts if isinstance(c, str) and "im_start" in c)
    want = []
    for d in data:
        conv = conv_lib.conv_vicuna_v3.copy()
        conv.tokenizer = tok
        conv.append_message("system", conv.system)
        opts = "\n".join(d["options"])
        q = f"""Question: {d['question']}

        Options:
        {opts}

        Please carefully read the question and select the single correct answer from A-D.
        You can only output one option from A), B), C), D) with format 'The correct answer is' without explanation."""
        conv.append_message("user", "<seq>\n" + q)
        ids = tokenizer_seq_token(conv.get_prompt_eval(), tok, -200, return_tensors="pt")[None].cuda()
        out = model.generate(ids, [d["input"]], attention_mask=None, pad_token_id=tok.eos_token_id, do_sample=False,
                             max_new_tokens=6)
        want.append(script.after_process_output(tok.batch_decode(out, skip_special_tokens=True)[0], conv.sep))
    assert [g["generated"] for g in got] == want


def test_run_opus_online_unmodified(release, monkeypatch, capsys):
    from opus_pllm_b200 import builder, compat
    script = compat.load_script(os.path.join(EVAL, "run_opus_online.py"), "ref_run_opus_online")
    prot = synth.proteins(1, 40, seed=2)[0]
    feed = iter(["Describe the function of this protein.", "not a protein 123", prot,      # invalid sequence is re-asked
                 "What is a kinase?", ""])                                                  # text-only turn (seq=None)

    def fake_input(prompt=""):
        try:
            return next(feed)
        except StopIteration:
            raise KeyboardInterrupt
    monkeypatch.setattr(builtins, "input", fake_input)
    a = _args(release, max_new_tokens=8)
    with pytest.raises(KeyboardInterrupt):
        script.eval_model(a)
    out = capsys.readouterr().out
    assert "Invalid sequence!" in out and out.count("Output:") == 2
    printed = [l[len("Output: "):] for l in out.splitlines() if l.startswith("Output: ")]
    # the same two turns through direct calls
    from multi_modality_model.multi_modality_v1.conversation import conv_vicuna_v0
    from multi_modality_model.multi_modality_v1.mm_utils import tokenizer_seq_token
    tok, model, _ = builder.load_pretrained_model(release["base"], release["weights"], "Meta-Llama-3-tiny",
                                                  cstp_path=builder.return_cstp_path(release["weights"], "modality_encoder/modality_encoding_adapter.ckpt"))
    head = f"{conv_vicuna_v0.system}\n\n### Student: "
    p1 = head + "<seq>\nDescribe the function of this protein.\n### Professor:"
    ids1 = tokenizer_seq_token(p1, tok, -200, return_tensors="pt").unsqueeze(0).cuda()
    o1 = model.generate(ids1, prot, attention_mask=None, pad_token_id=tok.eos_token_id, do_sample=False, max_new_tokens=8)
    p2 = head + "What is a kinase?\n### Professor:"
    ids2 = torch.as_tensor(tok([p2]).input_ids).cuda()
    o2 = model.generate(ids2, None, attention_mask=None, pad_token_id=tok.eos_token_id, do_sample=False, max_new_tokens=8)

    def post(o):
        s = tok.batch_decode(o, skip_special_tokens=True)[0].strip()
        s += "" if "###" in s[2:] else "###"
        return s[: s.index("###", 2)].strip()
    assert printed == [post(o1), post(o2)]
