"""Pre-computed protein embeddings: counterpart of `scripts/generate_esm_embedding.py` and
`scripts/generate_esm_for_each_seq.py` (the producers of the `input_embed` field that `forward(input_embed=...)` /
`generate(seq_embedding=...)` consume, opus_arch.py:151-161, opus_llama.py:52-56).

Same inputs and outputs as the reference scripts -- a JSON list of {instruction, input, output} records in, a JSONL file
with an added `input_embed` (float list, 1280) out, or (`--dict-only`) a {sequence: embedding} JSON dictionary; sequences
already present in `--dict_path` are reused; sequences longer than 4000 residues are skipped -- but the encoder runs on
length-sorted, packed batches (`--max-tokens` residues per forward) instead of one protein per call with an
`empty_cache()` after each (generate_esm_embedding.py:17-27).

    python -m opus_pllm_b200.generate_esm_embedding --file_path data.json --save_path data_embed.jsonl \\
        [--dict_path seq2embed.json] [--esm-path esm2_t33_650M_UR50D.pt] [--ckpt cstp.ckpt] [--dict-only]
"""
from __future__ import annotations

import argparse
import json

MAX_LEN = 4000   # generate_esm_embedding.py:19 (`> 4000` skipped); generate_esm_for_each_seq.py:16 keeps `< 4000`


def batches_by_length(seqs: list[str], max_tokens: int) -> list[list[int]]:
    """indices of `seqs` grouped so that a batch holds at most `max_tokens` residues (+2 per protein); longest first, so
    the varlen attention sees similar lengths together."""
    order = sorted(range(len(seqs)), key=lambda i: -len(seqs[i]))
    out, cur, tok = [], [], 0
    for i in order:
        n = len(seqs[i]) + 2
        if cur and tok + n > max_tokens:
            out.append(cur)
            cur, tok = [], 0
        cur.append(i)
        tok += n
    if cur:
        out.append(cur)
    return out


def embed_all(encoder, seqs: list[str], max_tokens: int = 65536) -> dict[str, list[float]]:
    uniq = sorted(set(seqs))
    res = {}
    for idx in batches_by_length(uniq, max_tokens):
        emb = encoder.get_protein_seq_embeddings([uniq[i] for i in idx]).cpu().numpy()
        for j, i in enumerate(idx):
            res[uniq[i]] = emb[j].tolist()
    return res


def generate_esm_embedding(args, encoder=None):
    data = json.load(open(args.file_path))
    data = [{"instruction": d["instruction"], "input": d["input"], "output": d["output"]} for d in data]
    known = json.load(open(args.dict_path)) if args.dict_path else {}
    limit_ok = (lambda s: len(s) < MAX_LEN) if args.dict_only else (lambda s: len(s) <= MAX_LEN)
    todo = [d["input"] for d in data if limit_ok(d["input"]) and d["input"] not in known]
    if todo:
        if encoder is None:
            from .builder import build_protein_encoder
            encoder = build_protein_encoder(args.ckpt, esm_path=args.esm_path)
        known = dict(known, **embed_all(encoder, todo, args.max_tokens))
    if args.dict_only:                                           # generate_esm_for_each_seq.py
        out = {d["input"]: known[d["input"]] for d in data if limit_ok(d["input"])}
        json.dump(out, open(args.save_path, "w"))
        return len(out)
    n = 0
    with open(args.save_path, "w") as f:                         # generate_esm_embedding.py: one JSON object per line
        for d in data:
            if not limit_ok(d["input"]):
                continue
            f.write(json.dumps(dict(d, input_embed=known[d["input"]])) + "\n")
            n += 1
    return n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--file_path", type=str, required=True)
    ap.add_argument("--save_path", type=str, required=True)
    ap.add_argument("--dict_path", type=str, default=None)
    ap.add_argument("--esm-path", type=str, default=None, help="fair-esm .pt or HF EsmModel dir (default: fair-esm hub cache)")
    ap.add_argument("--ckpt", type=str, default=None, help="CSTP checkpoint whose protein_model.model.* tensors override ESM-2")
    ap.add_argument("--max-tokens", type=int, default=65536)
    ap.add_argument("--dict-only", action="store_true", help="write a {sequence: embedding} dictionary instead of JSONL")
    print(generate_esm_embedding(ap.parse_args()), "records written")


if __name__ == "__main__":
    main()
