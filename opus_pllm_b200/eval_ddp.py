"""Batch annotation driver: counterpart of `multi_modality_v1/eval/run_opus_ddp.py` on the B200 backend.

Same flags, same prompt construction, same data-parallel scheme (contiguous shards of the prompt list, one gather at the
end; `--temperature 0.1 --top_p 0.7` sampling by default like the reference, `--temperature 0` = greedy). Differences:
the collective is a tensor all-gather of token ids instead of `gather_object`, `--seed` pins the sampling stream, and
`--continuous-batching` (greedy or sampled, `--stop-keyword` honoured) keeps a fixed number of decode slots busy instead
of walking fixed batches of 8. Launch with torchrun (one process per GPU) or plain python (one GPU).

  torchrun --nproc-per-node 8 -m opus_pllm_b200.eval_ddp --model-base-path <llama3 dir> \
      --opus-pllm-weights-path <weights dir> --input_path data.json --save_path out.json
"""
from __future__ import annotations

import argparse
import json
import os
import time

import torch
import torch.distributed as dist

from .constants import DEFAULT_SEQ_TOKEN, DEFAULT_SEQ_TOKEN_INDEX
from .dp import gather_token_ids, split_between_processes
from .mm_utils import after_process_output, get_model_name_from_path, left_pad_sequence, tokenizer_seq_token

# conv_vicuna_v0 of the reference (multi_modality_v1/conversation.py:159-170): roles and separator; the system text is
# taken from the reference package when it is importable so the prompt is byte-identical to the one the adapter was
# trained with, otherwise --system-prompt must be given.
ROLES = ("Student", "Professor")
SEP = "###"


def default_system_prompt() -> str | None:
    try:
        from multi_modality_model.multi_modality_v1.conversation import conv_vicuna_v0
        return conv_vicuna_v0.system
    except Exception:
        return None


def max_new_tokens_for(input_path: str, default: int | None = None) -> int:
    """the value run_opus_ddp.py:92-101 FORCES (whatever --max_new_tokens says) once it meets an instruction that does
    not carry `<seq>` itself: 32 for localization sets, 128 for keyword sets, 256 otherwise"""
    if "localization" in input_path:
        return 32
    if "keywords" in input_path:
        return 128
    return 256


def max_new_tokens_schedule(instructions, input_path: str, user_value: int, fixed: bool = False) -> list[int]:
    """max_new_tokens in force after each instruction, as the reference's loop leaves it (run_opus_ddp.py:90-101 mutates
    `args.max_new_tokens` while it builds the prompts, so the override is decided per instruction and sticks for the rest
    of the run): the user's value until the first instruction without `<seq>`, the forced value from then on."""
    cur, out = int(user_value), []
    for ins in instructions:
        if not fixed and DEFAULT_SEQ_TOKEN not in ins:
            cur = max_new_tokens_for(input_path)
        out.append(cur)
    return out


def build_prompt(instruct: str, system: str, input_path: str) -> str:
    """run_opus_ddp.py:90-108"""
    if DEFAULT_SEQ_TOKEN not in instruct:
        instruct = DEFAULT_SEQ_TOKEN + "\n" + instruct
        if "localization" in input_path:
            instruct += "Kindly reply with only one word."
    return f"{system}\n\n### {ROLES[0]}: {instruct}\n### {ROLES[1]}:"


def eval_model(args, tokenizer=None, model=None):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if model is None:
        from types import SimpleNamespace
        from .builder import load_pretrained_model, return_cstp_path
        acc = SimpleNamespace(process_index=local_rank, wait_for_everyone=(dist.barrier if world > 1 else lambda: None))
        cstp = return_cstp_path(args.opus_pllm_weights_path, "modality_encoder/modality_encoding_adapter.ckpt")
        tokenizer, model, _ = load_pretrained_model(args.model_base_path, args.opus_pllm_weights_path,
                                                    get_model_name_from_path(args.model_base_path), args.load_8bit,
                                                    args.load_4bit, accelerator=acc,
                                                    switch_projector_type=args.switch_projector_type, cstp_path=cstp,
                                                    esm_path=args.esm_path)
    system = args.system_prompt or default_system_prompt()
    if system is None:
        raise SystemExit("--system-prompt is required when the reference package is not importable")
    qs = [q for q in json.load(open(args.input_path)) if q["input"] is not None]
    seqs_all, instr_all, gt_all = [q["input"] for q in qs], [q["instruction"] for q in qs], [q["output"] for q in qs]
    seqs, instrs = split_between_processes(seqs_all, rank, world), split_between_processes(instr_all, rank, world)
    sched = max_new_tokens_schedule(instrs, args.input_path, args.max_new_tokens, args.max_new_tokens_fixed)
    # every rank pads its rows to the same width for the gather: the largest value the schedule can reach anywhere
    max_new = max([args.max_new_tokens] + max_new_tokens_schedule(instr_all, args.input_path, args.max_new_tokens,
                                                                 args.max_new_tokens_fixed))
    dev = model.device
    prompts = [build_prompt(i, system, args.input_path) for i in instrs]
    ids = [tokenizer_seq_token(p, tokenizer, DEFAULT_SEQ_TOKEN_INDEX, return_tensors="pt") for p in prompts]
    stop_kw, stop_ids = {}, []
    if getattr(args, "stop_keyword", False):   # opt-in: stop a row once it has emitted the "###" separator (device-side)
        from .mm_utils import KeywordsStoppingCriteria
        crit = KeywordsStoppingCriteria([SEP], tokenizer)
        stop_kw = {"stopping_criteria": [crit]}
        stop_ids = [[int(t) for t in torch.as_tensor(q).reshape(-1).tolist()] for q in crit.keyword_ids]
    t0 = time.time()
    rows = []
    pad_row = lambda o: torch.nn.functional.pad(o, (0, max_new - o.numel()), value=tokenizer.eos_token_id)  # noqa: E731
    if args.continuous_batching:
        from .scheduler import ContinuousBatcher
        # per request: the value the reference's fixed batch holding it would have been generated with
        per_req = [sched[min((i // args.batch_size + 1) * args.batch_size, len(ids)) - 1] for i in range(len(ids))]
        outs = ContinuousBatcher(model, max_slots=args.batch_size).generate(
            ids, seqs, per_req, eos_ids=model.config.eos_token_id, pad_id=tokenizer.eos_token_id, stop_sequences=stop_ids,
            sampling=(args.temperature, args.top_p, args.seed + 1000003 * rank) if args.temperature > 0 else None)
        rows = [pad_row(o) for o in outs]
    else:
        for i in range(0, len(ids), args.batch_size):
            batch = left_pad_sequence(ids[i: i + args.batch_size], tokenizer.pad_token_id, batch_first=True)
            batch_new = sched[min(i + args.batch_size, len(ids)) - 1]      # the value in force when this batch is generated
            out = model.generate(batch, seqs[i: i + args.batch_size], attention_mask=batch != tokenizer.pad_token_id,
                                 pad_token_id=tokenizer.eos_token_id, do_sample=args.temperature > 0,
                                 temperature=args.temperature, top_p=args.top_p, num_beams=args.num_beams,
                                 max_new_tokens=batch_new, use_cache=True, **stop_kw,
                                 **({"seed": args.seed + 1000003 * rank + i} if args.temperature > 0 else {}))
            rows.extend(pad_row(o) for o in out.cpu())
    local = torch.stack(rows).to(dev) if rows else torch.zeros((0, max_new), dtype=torch.int64, device=dev)
    gathered = gather_token_ids(local, tokenizer.eos_token_id)
    if rank == 0:
        texts = tokenizer.batch_decode(gathered.cpu(), skip_special_tokens=True)
        answers = [after_process_output(t, SEP) for t in texts]
        dt = time.time() - t0
        print(f"entries/sec: {len(qs) / dt}, time elapsed: {dt}")
        result = [{"ground_truth": g, "generated": a} for g, a in zip(gt_all, answers)]
        with open(args.save_path, "w") as f:
            json.dump(result, f)
        return result
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model-base-path", type=str, required=True)
    ap.add_argument("--opus-pllm-weights-path", type=str, required=True)
    ap.add_argument("--is_json", type=bool, default=True)
    ap.add_argument("--input_path", type=str, required=True)
    ap.add_argument("--save_path", type=str, required=True)
    ap.add_argument("--temperature", type=float, default=0.1)      # run_opus_ddp.py:156
    ap.add_argument("--seed", type=int, default=0, help="sampling stream (per rank and batch offsets are added)")
    ap.add_argument("--top_p", type=float, default=0.7)
    ap.add_argument("--num_beams", type=int, default=1)
    ap.add_argument("--max_new_tokens", type=int, default=32)
    ap.add_argument("--max_new_tokens_fixed", action="store_true", help="do not apply the dataset-name heuristics")
    ap.add_argument("--switch_projector_type", type=str, default="mlp2x_gelu")
    ap.add_argument("--load-4bit", type=bool, default=False)
    ap.add_argument("--load-8bit", type=bool, default=False)
    ap.add_argument("--batch_size", type=int, default=64)
    ap.add_argument("--continuous-batching", action="store_true")
    ap.add_argument("--stop-keyword", action="store_true",
                    help='finish a row when it emits the "###" separator (mm_utils.KeywordsStoppingCriteria, on the device)')
    ap.add_argument("--esm-path", type=str, default=None)
    ap.add_argument("--system-prompt", type=str, default=None)
    eval_model(ap.parse_args())


if __name__ == "__main__":
    main()
