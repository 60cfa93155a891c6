"""Continuous batching over the paged KV cache (BASELINE config 5: long generations, 512 new tokens).

The reference has no scheduler: `run_opus_ddp.py` walks the prompt list in fixed batches of 8 and every batch runs until
its slowest row stops (`eval/run_opus_ddp.py:75,88-134`). Here a fixed number of decode *slots* is kept busy instead:
whenever a sequence emits EOS (or reaches max_new_tokens) its pages go back to the allocator and the slot is refilled by
prefilling the next waiting prompt, while the other slots keep decoding. Outputs are returned in input order, so the
call stays a drop-in for the per-batch `generate` loop of the eval scripts. Greedy, or temperature / top-p sampling
(`sampling=(temperature, top_p, seed)`, the eval scripts' default decode): every round draws from its own seed stream
(the draw is hash(seed, row, step) and steps restart every round); the seed is handed over in device memory
(`opus_decode_state.seed_ptr`), so sampled rounds replay the same CUDA graph as greedy ones.

All device work reuses the C ABI entry points (`opus_llama_prefill`, `opus_llama_select`, `opus_llama_decode_loop`); the
scheduler itself is host logic over the decode-state arrays.
"""
from __future__ import annotations

import ctypes as C
import time
from collections import deque

import numpy as np
import torch

from . import _lib as L
from . import ops
from .llama import BLOCK, u64_as_i64
from .model import B200OpusLlama, SplicePlan


def cut_round(toks: np.ndarray, room: np.ndarray, eos_ids) -> tuple[np.ndarray, np.ndarray]:
    """Per slot of one round's tokens [S, R]: how many tokens the request keeps (up to and including its first EOS, at most
    `room[s]` = what is left of its max_new_tokens) and whether that completes it. Stop keywords are matched separately."""
    S, R = toks.shape
    first = np.full(S, R + 1, dtype=np.int64)                # tokens up to and including the first EOS; R + 1 = none
    if len(eos_ids) and S:
        hit = np.isin(toks, np.asarray(list(eos_ids), dtype=toks.dtype))
        any_hit = hit.any(1)
        first[any_hit] = hit[any_hit].argmax(1) + 1
    take = np.minimum(np.minimum(first, R), np.maximum(room, 0))
    done = (take >= room) | (first <= take)
    return take, done


class ContinuousBatcher:
    def __init__(self, model: B200OpusLlama, max_slots: int = 64, round_steps: int = 8, encoder_chunk: int = 64):
        self.model, self.llama = model, model.llama
        self.S, self.R, self.enc_chunk = max_slots, round_steps, encoder_chunk
        self.dev = model.device
        self.stats = {}

    # ------------------------------------------------------------------------------------------ helpers
    def _soft_tokens(self, seqs):
        outs = []
        for i in range(0, len(seqs), self.enc_chunk):
            outs.append(self.model._soft_tokens(list(seqs[i: i + self.enc_chunk]), None))
        return torch.cat(outs, 0)  # [n, n_soft, H]

    def _state(self, n: int, max_blocks: int, out_ld: int, eos_ids, pad_id: int, sampling=None):
        i32 = lambda *s: torch.zeros(s, dtype=torch.int32, device=self.dev)  # noqa: E731
        bufs = dict(next_tok=i32(n), ctx_len=i32(n), pos=i32(n), slot=i32(n), block_table=i32(n, max_blocks),
                    finished=i32(n), n_unfinished=i32(1), step=i32(1), out_ids=i32(n, out_ld),
                    eos=torch.tensor(list(eos_ids) or [-1], dtype=torch.int32, device=self.dev),
                    seed=torch.zeros(1, dtype=torch.int64, device=self.dev))
        s = L.DecodeState()
        for k in ("next_tok", "ctx_len", "pos", "slot", "block_table", "finished", "n_unfinished", "step", "out_ids"):
            setattr(s, k, bufs[k].data_ptr())
        s.max_blocks, s.out_ld = max_blocks, out_ld
        s.eos_ids, s.n_eos, s.pad_id = bufs["eos"].data_ptr(), len(eos_ids), pad_id
        if sampling is not None:
            # the seed lives in device memory (seed_ptr): the per-round streams below do not invalidate the decode graph
            s.do_sample, s.temperature, s.top_p, s.seed = 1, float(sampling[0]), float(sampling[1]), 0
            bufs["seed"].fill_(u64_as_i64(int(sampling[2])))
            s.seed_ptr = bufs["seed"].data_ptr()
        return s, bufs

    @staticmethod
    def _mix(seed: int, n: int) -> int:
        """independent 64-bit stream per (seed, round / admission index): splitmix64 finaliser"""
        x = (seed + 0x9E3779B97F4A7C15 * (n + 1)) & 0xFFFFFFFFFFFFFFFF
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return x ^ (x >> 31)

    # ------------------------------------------------------------------------------------------ main entry
    @torch.no_grad()
    def generate(self, prompts: list[torch.Tensor], seqs: list[str], max_new_tokens, eos_ids=(), pad_id: int = 0,
                 use_graph: bool = True, sampling=None, stop_sequences=()) -> list[torch.Tensor]:
        """prompts[i]: 1-D int64 ids with one -200 sentinel (no padding); seqs[i]: its protein. Returns, per prompt, the
        new tokens (int64, cut after the first EOS, at most max_new_tokens) — same content as generate() row by row.
        max_new_tokens: one int, or one int per request. stop_sequences: token-id sequences after which a request is
        complete (the `"###"` keyword of mm_utils.KeywordsStoppingCriteria), matched on the host when a round is read back."""
        lib, ll = L.load(), self.llama
        n_req = len(prompts)
        if n_req == 0:                     # e.g. a rank whose shard of the prompt list is empty
            return []
        lib.opus_release_graphs()  # decode graphs are keyed by buffer addresses; ours are private to this call
        eos_set = set(int(e) for e in eos_ids)
        stops = [tuple(int(t) for t in q) for q in (stop_sequences or ()) if len(q)]
        new_of = [int(max_new_tokens)] * n_req if np.isscalar(max_new_tokens) else [int(v) for v in max_new_tokens]
        assert len(new_of) == n_req
        max_new_tokens = max(new_of)
        n_soft = self.model.n_soft
        # per-request splice plans against protein slot 0; the slot of a request inside its admission chunk is added when
        # it is admitted (the encoder + projectors run per admission chunk, not for the whole queue up front)
        plans = [SplicePlan(p.detach().cpu().numpy()[None, :], None, n_soft, 1).src for p in prompts]
        lens = np.array([len(s) for s in plans], dtype=np.int64)
        margin = self.R + 1
        max_blocks = int((lens.max() + max_new_tokens + margin + BLOCK - 1) // BLOCK)
        need_pages = lambda ln: int((ln + max_new_tokens + margin + BLOCK - 1) // BLOCK)  # noqa: E731
        S = min(self.S, n_req)
        # positions reach prompt + new tokens + the steps a finished slot keeps spinning until its round is read back
        total = int(lens.max()) + max_new_tokens + margin
        if total > ll.max_positions:
            ll._build_rope(1 << (total - 1).bit_length())     # OPT: raises (its learned position table cannot grow)
        # worst case all slots hold the longest prompts, +1 scratch page for idle slots
        ll._ensure_cache(S * max_blocks + 1)
        ll._ensure_ws(int(max(lens.max() * min(S, self.enc_chunk), S)), S)
        scratch = ll._alloc.alloc(1)[0]

        st, bufs = self._state(S, max_blocks, self.R, eos_ids, pad_id, sampling)
        n_round = n_adm = 0
        round_events, adm_events, n_rounds, prefill_tokens = [], [], 0, 0
        t_wall0 = time.perf_counter()
        bufs["block_table"].fill_(scratch)
        bufs["finished"].fill_(1)
        slot_req = [-1] * S                   # request id held by each slot
        slot_pages: list[list[int]] = [[] for _ in range(S)]
        results: list[list[int]] = [[] for _ in range(n_req)]
        done = [False] * n_req
        waiting = deque(range(n_req))
        stream = torch.cuda.current_stream().cuda_stream
        active = 0

        retired: list[int] = []               # slots freed since the last flush of the device-side state

        def retire(slot):
            nonlocal active
            r = slot_req[slot]
            done[r] = True
            ll._alloc.release(slot_pages[slot])
            slot_pages[slot], slot_req[slot] = [], -1
            retired.append(slot)
            active -= 1

        def flush_retired():
            """one indexed update for all slots retired since the last call: scratch page, finished, empty context"""
            if retired:
                idx = torch.tensor(retired, dtype=torch.long, device=self.dev)
                bufs["block_table"][idx] = scratch
                bufs["finished"][idx] = 1
                bufs["ctx_len"][idx] = 0
                retired.clear()

        def absorb(r, toks) -> bool:
            """append tokens of request r; True when the request is complete"""
            for t in toks:
                results[r].append(int(t))
                if int(t) in eos_set or len(results[r]) >= new_of[r]:
                    return True
                for q in stops:
                    if len(results[r]) >= len(q) and tuple(results[r][-len(q):]) == q:
                        return True
            return False

        def admit_chunk() -> bool:
            """prefill one chunk of waiting prompts into free slots (bounded by the workspace rows and the encoder chunk)"""
            nonlocal n_adm, prefill_tokens, active
            flush_retired()
            free = [s for s in range(S) if slot_req[s] < 0]
            adm = []
            budget = ll._ws_rows
            while free and waiting and lens[waiting[0]] <= budget and len(adm) < self.enc_chunk:
                r = waiting.popleft()
                adm.append((free.pop(0), r))
                budget -= int(lens[r])
            if not adm:
                return False
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            adm_events.append((a0, a1))
            cu = np.zeros(len(adm) + 1, dtype=np.int32)
            np.cumsum([lens[r] for _, r in adm], out=cu[1:])
            n_tok = int(cu[-1])
            bt = np.full((len(adm), max_blocks), scratch, dtype=np.int32)
            for j, (s, r) in enumerate(adm):
                pages = ll._alloc.alloc(need_pages(int(lens[r])))
                slot_pages[s] = pages
                bt[j, : len(pages)] = pages
            seq_of = np.repeat(np.arange(len(adm), dtype=np.int32), np.diff(cu))
            pos = (np.arange(n_tok, dtype=np.int32) - cu[:-1][seq_of]).astype(np.int32)
            slot_map = (bt[seq_of, pos // BLOCK] * BLOCK + pos % BLOCK).astype(np.int32)
            soft = self.model._soft_tokens([seqs[r] for _, r in adm], None)          # [len(adm), n_soft, H]
            soft2d = soft.reshape(-1, soft.shape[-1]).contiguous()
            parts = []
            for j, (_, r) in enumerate(adm):
                src = plans[r].copy()
                src[src < 0] -= j * n_soft                                           # -(k+1) -> -(j*n_soft + k + 1)
                parts.append(src)
            src = np.concatenate(parts).astype(np.int32)
            embeds = ops.splice_gather(ops.h2d(src, self.dev), ll.embed, soft2d)
            d_pos, d_slot, d_cu = (ops.h2d(a, self.dev) for a in (pos, slot_map, cu))
            d_last = ops.h2d((cu[1:] - 1).astype(np.int32), self.dev)
            L.check(lib.opus_llama_prefill(C.byref(ll._model), C.byref(ll._cache), C.byref(ll._ws),
                                           embeds.data_ptr(), d_pos.data_ptr(), d_slot.data_ptr(), d_cu.data_ptr(),
                                           d_last.data_ptr(), len(adm), n_tok, int(np.diff(cu).max()), stream),
                    "opus_llama_prefill")
            n_adm += 1
            prefill_tokens += n_tok
            adm_sampling = None if sampling is None else (sampling[0], sampling[1], self._mix(int(sampling[2]) ^ 0x5DEECE66D, n_adm))
            ast, ab = self._state(len(adm), max_blocks, 1, eos_ids, pad_id, adm_sampling)
            ab["n_unfinished"].fill_(len(adm))
            L.check(lib.opus_llama_select(C.byref(ll._model), C.byref(ll._ws), C.byref(ast), len(adm), stream),
                    "opus_llama_select")
            a1.record()
            first = ops.d2h(ab["next_tok"]).tolist()
            idx = torch.tensor([s for s, _ in adm], dtype=torch.long, device=self.dev)
            bufs["next_tok"][idx] = ab["next_tok"]
            bufs["ctx_len"][idx] = torch.tensor([int(lens[r]) for _, r in adm], dtype=torch.int32, device=self.dev)
            bufs["block_table"][idx] = ops.h2d(bt, self.dev)
            bufs["finished"][idx] = 0
            for j, (s, r) in enumerate(adm):
                slot_req[s] = r
                active += 1
                if absorb(r, [first[j]]):
                    retire(s)
            return True

        # decode batch tiers of the kernels (batch tile 32 / 64 / 128 / 256, two batch tiles above that): once the queue is
        # empty the live requests are compacted into the smallest tier that holds them, so the ramp-down of a job does not
        # pay full-batch steps
        tiers = sorted({t for t in (32, 64, 128, 256, 384, 512) if t < S} | {S})
        n_compact = 0
        while waiting or active:
            # ---- admit chunk after chunk until the slots are full or the queue is empty; only then decode a round
            while waiting and any(r < 0 for r in slot_req):
                if not admit_chunk():
                    break
            if not waiting and 0 < active and min(t for t in tiers if t >= active) < S:
                tgt = min(t for t in tiers if t >= active)
                retired.clear()                     # the compacted state below is rebuilt from the live slots only
                live = [i for i in range(S) if slot_req[i] >= 0]
                st2, bufs2 = self._state(tgt, max_blocks, self.R, eos_ids, pad_id, sampling)
                idx = torch.tensor(live, dtype=torch.long, device=self.dev)
                n_live = len(live)
                bufs2["block_table"].fill_(scratch)
                bufs2["finished"].fill_(1)
                for k in ("next_tok", "ctx_len", "block_table"):
                    bufs2[k][:n_live] = bufs[k][idx]
                bufs2["finished"][:n_live] = 0
                slot_req = [slot_req[i] for i in live] + [-1] * (tgt - n_live)
                slot_pages = [slot_pages[i] for i in live] + [[] for _ in range(tgt - n_live)]
                st, bufs, S = st2, bufs2, tgt
                n_compact += 1
            if not active:
                continue
            # ---- one round of R decode steps over all slots (idle slots spin on the scratch page)
            flush_retired()
            idle = [s for s in range(S) if slot_req[s] < 0]
            if idle:    # idle slots keep stepping on the scratch page: their context restarts every round
                bufs["ctx_len"][torch.tensor(idle, device=self.dev)] = 0
            bufs["step"].fill_(-1)
            bufs["n_unfinished"].fill_(active)
            if sampling is not None:
                n_round += 1
                bufs["seed"].fill_(u64_as_i64(self._mix(int(sampling[2]), n_round)))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.opus_llama_decode_loop(C.byref(ll._model), C.byref(ll._cache), C.byref(ll._ws), C.byref(st), S,
                                            self.R, 0, int(use_graph), stream)
            L.check(rc, "opus_llama_decode_loop")
            e1.record()
            round_events.append((e0, e1))
            n_rounds += 1
            toks = ops.d2h(bufs["out_ids"]).numpy()          # [S, R]; synchronises
            if stops:
                for s in range(S):
                    r = slot_req[s]
                    if r >= 0 and absorb(r, toks[s]):
                        retire(s)
            else:       # no stop keywords: the cut positions of all slots at once
                live = [s for s in range(S) if slot_req[s] >= 0]
                room = np.array([new_of[slot_req[s]] - len(results[slot_req[s]]) for s in live], dtype=np.int64)
                take, fin = cut_round(toks[live], room, eos_set)
                for j, s in enumerate(live):
                    results[slot_req[s]].extend(toks[s, : take[j]].tolist())
                    if fin[j]:
                        retire(s)
        ll._alloc.release([scratch])
        torch.cuda.current_stream().synchronize()
        # bookkeeping of the last call (bench.py: HBM roofline of the decode rounds)
        self.stats = dict(rounds=n_rounds, decode_steps=n_rounds * self.R, slots=min(self.S, n_req), compactions=n_compact,
                          admissions=n_adm, prefill_tokens=prefill_tokens,
                          decode_ms=float(sum(a.elapsed_time(b) for a, b in round_events)),
                          admission_ms=float(sum(a.elapsed_time(b) for a, b in adm_events)),
                          wall_ms=(time.perf_counter() - t_wall0) * 1e3)
        lib.opus_release_graphs()
        return [torch.tensor(r, dtype=torch.int64) for r in results]
