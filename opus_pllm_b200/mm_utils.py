"""Prompt helpers with the reference's semantics (multi_modality_v1/mm_utils.py:12-41, eval/run_opus_ddp.py:19-44)."""
from __future__ import annotations

import torch

from .constants import DEFAULT_SEQ_TOKEN, DEFAULT_SEQ_TOKEN_INDEX


def tokenizer_seq_token(prompt: str, tokenizer, seq_token_index: int = DEFAULT_SEQ_TOKEN_INDEX, return_tensors=None):
    """Tokenise the text around every `<seq>` marker and put `seq_token_index` (-200) where the protein goes; a BOS
    produced for every chunk is kept only once (mm_utils.py:12-32)."""
    chunks = [tokenizer(c).input_ids for c in prompt.split(DEFAULT_SEQ_TOKEN)]
    ids: list[int] = []
    offset = 0
    if chunks and chunks[0] and chunks[0][0] == tokenizer.bos_token_id:
        offset = 1
        ids.append(chunks[0][0])
    for i, c in enumerate(chunks):
        if i > 0:
            ids.append(seq_token_index)
        ids.extend(c[offset:])
    if return_tensors is None:
        return ids
    if return_tensors == "pt":
        return torch.tensor(ids, dtype=torch.long)
    raise ValueError(f"Unsupported tensor type: {return_tensors}")


def get_model_name_from_path(model_path: str) -> str:
    parts = model_path.strip("/").split("/")
    return parts[-2] + "_" + parts[-1] if parts[-1].startswith("checkpoint-") else parts[-1]


def left_pad_sequence(sequences, padding_value, batch_first: bool = False) -> torch.Tensor:
    """run_opus_ddp.py:30-44"""
    n = max(s.size(0) for s in sequences)
    out = torch.stack([torch.cat([torch.full((n - s.size(0),), padding_value, dtype=s.dtype, device=s.device), s])
                       for s in sequences])
    return out if batch_first else out.transpose(0, 1)


def after_process_output(text: str, sep: str) -> str:
    """cut the generation at the conversation separator (run_opus_ddp.py:19-27)"""
    text = text.strip()
    i = text.find(sep)
    return (text if i < 0 else text[:i]).strip()


class KeywordsStoppingCriteria:
    """Host half of the reference's stop-keyword criterion (mm_utils.py:43-61): keywords -> token-id sequences, a
    leading BOS dropped. Pass it as `model.generate(..., stopping_criteria=[crit])`: the matching runs on the device,
    once per decode step and per row (`opus_stop_sequences`), instead of the reference's per-step host loop over
    `output_ids[0, -len:]` plus a `batch_decode` substring search (mm_utils.py:63-75). Rows stop individually -- the
    reference's `all(outputs)` only stops when every row matches in the same step -- which leaves the text up to and
    including the keyword unchanged, and `after_process_output` cuts there anyway."""

    def __init__(self, keywords, tokenizer, input_ids=None):
        self.keywords = list(keywords)
        self.keyword_ids = []
        self.max_keyword_len = 0
        for keyword in self.keywords:
            ids = list(tokenizer(keyword).input_ids)
            if len(ids) > 1 and ids[0] == tokenizer.bos_token_id:
                ids = ids[1:]
            self.max_keyword_len = max(self.max_keyword_len, len(ids))
            self.keyword_ids.append(torch.tensor(ids))
        self.tokenizer = tokenizer
        self.start_len = 0 if input_ids is None else input_ids.shape[1]
