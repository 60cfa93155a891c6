#!/usr/bin/env python
"""bench.py — OPUS-PLLM protein-conditioned generation hot path on B200 (contract: see the task statement / DESIGN.md §5).

One "step" = one pass of the whole hot path over one batch of synthetic prompts:
    ESM-2-650M encoder -> mean-pool/L2 -> CSTP + switch projector -> soft-token splice -> Llama-3-8B(+LoRA merged)
    prefill -> greedy decode.
Default workload = BASELINE.json configs[1] ("c2"): 64 prompts/GPU, protein length 256, spliced prompt length 512,
32 new tokens. metric = generated tokens/s over the whole job (all ranks), weak scaling (fixed work per GPU).

  value : device-timed, inputs (token ids, splice map, KV plan) already resident in HBM
  e2e   : same metric through the reference-shaped public call model.generate(input_ids, seqs, ...) with HOST inputs
          (pinned ids + python strings) and a device->host read of the generated ids inside the timed region
  roofline     : the dominant kernel (prefill tcgen05 GEMM), timed live with CUDA events, vs MEASURED_PEAKS.json
  cpu_baseline : the oracle (reference path restated, oracle/) on this box's host cores, bounded sample, rank 0, N=1

`--impl reference` times the reference's own CPU implementation of the path (the oracle port: the reference is pure
Python over torch / fair-esm / transformers and fair-esm is not installable offline) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # generate-type steps: encoder -> projectors -> splice -> prefill -> greedy decode of one batch
    "c2": dict(kind="generate", batch=64, protein_len=256, prompt_len=512, new_tokens=32,
               desc="OPUS-PLLM-Llama3-8B random-init, subcellular-localization prompts, bs 64, seq 512, 32 new tokens"),
    # BASELINE configs[2]: 4096 prompts over 8 GPUs = one shard of 512 prompts per GPU, decoded as ONE batch
    "c3": dict(kind="generate", batch=512, protein_len=256, prompt_len=128, new_tokens=128,
               desc="Llama3-8B+LoRA GO-term generation shard, 512 prompts/GPU in one batch, prompt 128, 128 new tokens"),
    "tiny": dict(kind="generate", batch=8, protein_len=64, prompt_len=64, new_tokens=8,
                 desc="tiny smoke workload (not a bench line)"),
    # BASELINE configs[0]: encoder + projector forward only (metric 2, residues/s)
    "c1": dict(kind="encode", batch=64, protein_len=256,
               desc="cstp_v3 protein encoder (ESM-2-650M) + CSTP/switch projector fwd, 64 synthetic seqs len 256"),
    # BASELINE configs[3]: long-protein stress, packed varlen encoder + prefill, no decode
    "c4": dict(kind="encode_prefill", batch=256, protein_len=(1024, 2048), prompt_len=128,
               desc="long-protein encoder stress: 256 seqs len U[1024,2048] varlen, encoder + projectors + prefill only"),
    # BASELINE configs[4]: long generations with ragged stops, continuous batching over the paged KV cache
    "c5": dict(kind="continuous", requests=1024, slots=512, protein_len=(64, 1024), prompt_len=128, new_tokens=512,
               stop_len=(64, 512),
               desc="functional-description generation, up to 512 new tokens, continuous batching (512 slots) with paged KV; "
                    "a random-init model has no meaningful EOS, so every request carries a synthetic stop length "
                    "U[64,512] (deterministic work)"),
}


def _shrink(wl: dict) -> dict:
    """--size tiny: the same step on toy sizes (CPU-side plumbing checks, smoke runs)"""
    wl = dict(wl)
    for k, cap in (("batch", 8), ("requests", 12), ("slots", 4), ("prompt_len", 48), ("new_tokens", 12)):
        if k in wl:
            wl[k] = min(wl[k], cap)
    pl = wl.get("protein_len")
    wl["protein_len"] = (16, 48) if isinstance(pl, tuple) else min(pl, 48)
    if "stop_len" in wl:
        wl["stop_len"] = (3, wl["new_tokens"])
    return wl


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained",
                    d["bf16_tflops"]), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc = index, None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        self.result = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out = ""
        sm, mx, power, reasons = [], [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            # samples under load = the upper half of the power draw distribution
            thr = statistics.median(power) if power else 0
            load = [s for s, p in zip(sm, power) if p >= thr] or sm
            self.result = {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                           "power_w_max": max(power) if power else None, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
import contextlib


@contextlib.contextmanager
def _fast_cpu_weights(torch, synth):
    """synth.weight replaced by a seeded CPU uniform generator with the same mean / std for the duration of the block"""
    import math
    g = torch.Generator().manual_seed(0)
    orig = synth.weight

    def fast(shape, name, std, seed=0, mean=0.0, device="cpu"):
        a = std * math.sqrt(3.0)
        return torch.empty(tuple(shape), dtype=torch.float32).uniform_(-a, a, generator=g).add_(mean)
    synth.weight = fast
    try:
        yield
    finally:
        synth.weight = orig


METRIC_TOKENS = "generated tokens/s (whole job)"
METRIC_RESIDUES = "encoder residues/s (whole job)"


def cpu_reference_runner(sd, wl, torch):
    """Returns (run_once() -> units done, sample description, cores, unit, metric). Full-size weights, bounded sample of
    the workload: generate-type workloads time encoder -> projectors -> splice -> prefill -> greedy decode and count
    generated tokens; encoder-type workloads (c1, c4) time encoder -> projectors and count residues."""
    import psutil
    from oracle import esm2_ref, llama_ref, mm_ref
    from opus_pllm_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    lcfg, ecfg = sd["llama_cfg"], sd["esm_cfg"]
    ew = {k: v.detach().to("cpu", torch.float32) for k, v in sd["esm"].items()}
    pw = {k: v.detach().to("cpu", torch.float32) for k, v in sd["proj"].items()}
    if wl["kind"] in ("encode", "encode_prefill"):
        pl = wl["protein_len"]
        n = 16 if not isinstance(pl, tuple) else 3
        seqs = synth.proteins(n, *(pl if isinstance(pl, tuple) else (pl,)))
        n_res = sum(len(s) for s in seqs)

        def run_enc():
            with torch.no_grad():
                pooled = esm2_ref.get_protein_seq_embeddings(ew, seqs, ecfg["n_layers"], ecfg["n_heads"])
                c = mm_ref.protein_forward(pooled, pw["protein_projection.linear.weight"],
                                           pw["protein_projection.linear.bias"])
                mm_ref.switch_projector(c, pw, lcfg["dim"])
            return n_res
        sample = (f"oracle port of the reference path, fp32, {cores} threads: ESM-2-650M encoder + CSTP/switch projectors on "
                  f"{n} proteins ({n_res} residues) of this workload's length distribution")
        return run_enc, sample, cores, "residues/s", METRIC_RESIDUES
    avail_gb = psutil.virtual_memory().available / 2 ** 30
    per_layer_gb = 4 * (lcfg["dim"] * (lcfg["n_q_heads"] + 2 * lcfg["n_kv_heads"]) * lcfg["head_dim"] +
                        lcfg["dim"] * lcfg["n_q_heads"] * lcfg["head_dim"] + 3 * lcfg["dim"] * lcfg["ffn_dim"]) / 2 ** 30
    fixed_gb = 4 * (2 * lcfg["vocab"] * lcfg["dim"]) / 2 ** 30 + 8.0
    n_layers = int(max(1, min(lcfg["n_layers"], (avail_gb * 0.6 - fixed_gb) // per_layer_gb)))
    keep = lambda k: (not k.startswith("model.layers.")) or int(k.split(".")[2]) < n_layers  # noqa: E731
    lw = {k: v.detach().to("cpu", torch.float32) for k, v in sd["llama"].items() if keep(k)}
    if sd.get("lora"):
        for k in [k[: -len(".lora_A.weight")] for k in sd["lora"] if k.endswith(".lora_A.weight") and keep(k)]:
            lw[k + ".weight"] = llama_ref.lora_merge_ref(lw[k + ".weight"], sd["lora"][k + ".lora_A.weight"].float().cpu(),
                                                         sd["lora"][k + ".lora_B.weight"].float().cpu(), 32.0, 16)
    ocfg = llama_ref.LlamaCfg(n_layers=n_layers, dim=lcfg["dim"], n_q_heads=lcfg["n_q_heads"],
                              n_kv_heads=lcfg["n_kv_heads"], head_dim=lcfg["head_dim"], ffn_dim=lcfg["ffn_dim"],
                              vocab=lcfg["vocab"])
    pl = wl["protein_len"]
    B, T, new = 4, min(128, wl["prompt_len"]), min(16, wl["new_tokens"])
    seqs = synth.proteins(B, *((pl[0], min(pl[1], 256)) if isinstance(pl, tuple) else (pl,)))
    prompts = synth.prompt_ids(B, T - 7, vocab=lcfg["vocab"])
    ids = torch.stack(prompts)

    def run_once():
        with torch.no_grad():
            pooled = esm2_ref.get_protein_seq_embeddings(ew, seqs, ecfg["n_layers"], ecfg["n_heads"])
            c = mm_ref.protein_forward(pooled, pw["protein_projection.linear.weight"],
                                       pw["protein_projection.linear.bias"])
            soft = mm_ref.switch_projector(c, pw, lcfg["dim"])
            emb, mask, _, _ = mm_ref.splice(ids, None, soft, lw["model.embed_tokens.weight"])
            out = llama_ref.greedy_generate(lw, ocfg, emb, mask, new)
        return int(out.numel())

    layers_note = "" if n_layers == lcfg["n_layers"] else f", Llama depth cut to {n_layers}/{lcfg['n_layers']} layers (host RAM)"
    sample = (f"oracle port of the reference path, fp32, {cores} threads: {B} prompts x (protein {len(seqs[0])} aa, "
              f"prompt {T} tokens, {new} new tokens), full-width ESM-2-650M + projectors + Llama-3-8B{layers_note}")
    return run_once, sample, cores, "tokens/s", METRIC_TOKENS


def _cpu_baseline(sd, wl, torch):
    run_once, sample, cores, unit, _ = cpu_reference_runner(sd, wl, torch)
    run_once()
    t0 = time.perf_counter()
    units = run_once()
    dt = time.perf_counter() - t0
    return {"value": units / dt, "unit": unit, "cores": cores, "kind": "port", "sample": sample, "seconds": dt}


def gpu_reference_leg(torch, wl):
    """SURVEY 8d, last row: the reference's GPU PyTorch path on the same box as the practical bar. Stock transformers
    bf16 sdpa Llama + HF EsmModel under fp16 autocast + torch projectors (tools/bench_hf_gpu.py; none of this repo's
    kernels), random init, the C2 shapes, at the eval scripts' batch of 8 (run_opus_ddp.py:75) and at the config's
    batch of 64. Runs in a subprocess after this arm's numbers are taken; failure is reported, not fatal."""
    import subprocess
    tool = os.path.join(ROOT, "tools", "bench_hf_gpu.py")
    try:
        r = subprocess.run([sys.executable, tool, "8", "64"], capture_output=True, text=True, timeout=420)
        rows = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
        if not rows:
            return {"unavailable": (r.stderr or "no output").strip().splitlines()[-1][:200]}
        return {"impl": rows[0]["impl"], "unit": "tokens/s",
                "by_batch": {str(x["batch"]): {"tokens_per_s": x["tokens_per_s"], "ms_per_step": x["ms_per_step"]} for x in rows},
                "note": "same 64 prompts x 512 tokens x 32 new tokens per step; device-timed, inputs created on the device"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def run_other_workload(args, wl, model, sd, lcfg, peaks, dev, rank, world, warmup, timed, sync_all, torch, dist):
    """c1 (encoder + projectors), c4 (long-protein encoder + projectors + prefill), c5 (continuous batching)."""
    import numpy as np
    from opus_pllm_b200 import ops, synth
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    kind, full = wl["kind"], args.size == "full"
    pl = wl["protein_len"]
    n_prot = wl.get("batch", wl.get("requests"))
    seqs = synth.proteins(n_prot, *(pl if isinstance(pl, tuple) else (pl,)), seed=1234 + rank)
    n_res = sum(len(s) for s in seqs)
    model._soft_tokens(seqs[:2], None)            # builds the fused projector object
    enc = model.protein_encoder
    cfg = {"workload": f"{args.workload}: {wl['desc']}", "proteins_per_gpu": n_prot, "protein_len": list(pl) if isinstance(pl, tuple) else pl,
           "weights": "random-init (hash-seeded) ESM-2-650M + CSTP/switch projectors + Llama-3-8B, LoRA r=16 merged at load",
           "parallelism": f"dp{world} (replica per GPU, independent shards, no collective on the model path)",
           "l2": "inputs larger than L2 (activations of one step exceed the 126 MB L2)"}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    extra = {}

    def enc_flops(pk):
        lens = (pk.cu[1:] - pk.cu[:-1]).astype("float64")
        return 2.0 * pk.n_tok * 648806400 + 4.0 * 1280 * 33 * float((lens ** 2).sum())

    if kind in ("encode", "encode_prefill"):
        pk = enc.tokenize(seqs)
        pk.device_arrays = tuple(ops.h2d(a, dev) for a in (pk.tokens, pk.pos, pk.scale, pk.cu))
        if kind == "encode_prefill":
            prompts = synth.prompt_ids(n_prot, wl["prompt_len"] - 7, vocab=lcfg["vocab"], seed=1234 + rank)
            ids_host = torch.stack(prompts).pin_memory()
            plan_sp = model._plan(ids_host, None, n_prot)
            src_d = ops.h2d(plan_sp.src, dev)
            plan = model.llama.make_plan(plan_sp.cu, 1)
            n_ptok = int(plan_sp.cu[-1])

        def step_device(record=False):
            if record: ev[0].record()
            _, pooled_l2, _, _ = enc.encode(None, packed=pk)
            if record: ev[1].record()
            soft = model._fused(pooled_l2)
            if kind == "encode_prefill":
                embeds = ops.splice_gather(src_d, model.llama.embed, soft.reshape(-1, lcfg["dim"]))
                if record: ev[2].record()
                model.llama.prefill(embeds, plan=plan)
            elif record:
                ev[2].record()
            if record: ev[3].record()
            return soft

        def step_e2e():
            # public calls with HOST inputs: python strings (+ pinned prompt ids) in, a device->host read of the result out
            if kind == "encode":
                return ops.d2h(model._soft_tokens(seqs, None))
            st = None
            try:
                soft = model._soft_tokens(seqs, None)
                sp = model._plan(ids_host, None, n_prot)
                embeds = ops.splice_gather(ops.h2d(sp.src, dev), model.llama.embed, soft.reshape(-1, lcfg["dim"]))
                st = model.llama.prefill(embeds, sp.cu, 1)
                return ops.d2h(st["logits"].float().argmax(-1))
            finally:
                if st is not None:
                    model.llama.release_plan(st)

        for _ in range(warmup):
            step_device()
        ops.launch_count(reset=True)
        with ClockSampler(local_rank) as clk:
            ms = timed(step_device, args.steps)
        launches = ops.launch_count(reset=True)
        sync_all(); step_device(record=True); sync_all()
        phases = {"encoder_ms": ev[0].elapsed_time(ev[1]), "projector_splice_ms": ev[1].elapsed_time(ev[2]),
                  "prefill_ms": ev[2].elapsed_time(ev[3])}
        value = world * n_res * args.steps / (ms / 1e3)
        for _ in range(2):
            step_e2e()
        ops.XFER["h2d_bytes"] = ops.XFER["d2h_bytes"] = 0
        ms_e2e = timed(step_e2e, args.steps)
        e2e = {"value": world * n_res * args.steps / (ms_e2e / 1e3), "unit": "residues/s",
               "h2d_bytes_per_step": ops.XFER["h2d_bytes"] // args.steps,
               "d2h_bytes_per_step": ops.XFER["d2h_bytes"] // args.steps, "ms_per_step": ms_e2e / args.steps}
        fl = enc_flops(pk) if full else 0.0
        # c1: the step IS the encoder (+ 0.4 ms of projectors), so the roofline uses the timed back-to-back loop (sustained,
        # power-capped clock) rather than the single event-bracketed step that follows an idle moment at boost clock
        enc_ms = (ms / args.steps - phases["projector_splice_ms"]) if kind == "encode" else phases["encoder_ms"]
        tf = fl / enc_ms / 1e9
        extra["encoder_single_step"] = {"encoder_ms": phases["encoder_ms"], "tflops": fl / phases["encoder_ms"] / 1e9,
                                        "frac_of_sustained_peak": fl / phases["encoder_ms"] / 1e9 / peaks["tf_sustained"],
                                        "note": "one event-bracketed step after an idle moment (boost clock)"}
        roofline = {"kernel": "ESM-2 encoder forward (tcgen05 GEMMs 2*N*648.8M + attention 4*1280*33*sum T^2, SURVEY 8d), "
                              "timed in the step with CUDA events",
                    "bound": "tensor", "achieved": tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                    "frac": tf / peaks["tf_sustained"], "traffic": None,
                    "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)"}
        cfg.update(encoder_tokens=int(pk.n_tok), residues=n_res)
        if kind == "encode_prefill":
            pre_fl = (2.0 * n_ptok * 6979321856 + 2.0 * n_prot * 4096 * 128256 +
                      2.0 * 4096 * 32 * float((np.diff(plan_sp.cu).astype("float64") ** 2).sum())) if full else 0.0
            extra["roofline_prefill"] = {"bound": "tensor", "achieved": pre_fl / phases["prefill_ms"] / 1e9,
                                         "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                                         "frac": pre_fl / phases["prefill_ms"] / 1e9 / peaks["tf_sustained"],
                                         "prompt_tokens": n_ptok}
            cfg.update(prompt_len=wl["prompt_len"], prompt_tokens=n_ptok)
        metric, unit = METRIC_RESIDUES, "residues/s"
        extra["phases_ms"] = phases
    else:
        # ---- c5: continuous batching
        from opus_pllm_b200.scheduler import ContinuousBatcher
        import random
        rng = random.Random(4321 + rank)
        stop_len = [rng.randint(*wl["stop_len"]) for _ in range(n_prot)]
        prompts = synth.prompt_ids(n_prot, wl["prompt_len"] - 7, vocab=lcfg["vocab"], seed=1234 + rank, ragged=17)
        cb = ContinuousBatcher(model, max_slots=wl["slots"], round_steps=16)
        n_tok = sum(stop_len)

        def step():
            return cb.generate(prompts, seqs, stop_len, eos_ids=(), pad_id=128001, use_graph=not args.no_graph)

        for _ in range(min(warmup, 3)):
            step()
        ops.launch_count(reset=True)
        ops.XFER["h2d_bytes"] = ops.XFER["d2h_bytes"] = 0
        with ClockSampler(local_rank) as clk:
            ms = timed(step, args.steps)
        launches = ops.launch_count(reset=True)
        value = world * n_tok * args.steps / (ms / 1e3)
        # host prompts / strings in, host token lists out: the scheduler's public call already is the end-to-end path
        e2e = {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": ops.XFER["h2d_bytes"] // args.steps,
               "d2h_bytes_per_step": ops.XFER["d2h_bytes"] // args.steps, "ms_per_step": ms / args.steps,
               "note": "value IS the public ContinuousBatcher.generate(host prompts, host strings) call: its rounds read "
                       "their tokens back to the host and upload every admission's plan inside the timed region"}
        stats = dict(cb.stats)
        # HBM roofline of the decode rounds: weights once per step + KV of the live context per step
        plens = [int(p.numel()) + 7 for p in prompts]
        kv_bytes = 131072.0 * sum(n * pl_ + n * (n + 1) / 2.0 for n, pl_ in zip(stop_len, plens))
        w_bytes = 15009316864.0 * stats["decode_steps"] if full else 0.0
        dec_ms = stats["decode_ms"]
        ach = (w_bytes + kv_bytes) / dec_ms / 1e6 if dec_ms > 0 else 0.0
        roofline = {"kernel": "decode rounds of the continuous batcher (swap-AB tcgen05 GEMMs + paged attention), CUDA-event "
                              "time of the rounds of the last step", "bound": "hbm", "achieved": ach, "peak": peaks["hbm"],
                    "unit": "GB/s", "frac": ach / peaks["hbm"], "traffic": None, "peak_source": peaks["source"],
                    "bytes": w_bytes + kv_bytes, "decode_ms": dec_ms}
        # the tensor-pipe view of the same rounds (useful work only: one weight pass per generated token + its attention);
        # with 512 slots the full-batch rounds are tensor-bound, the drained tiers HBM-bound
        dec_flops = (2.0 * 7504658432 * n_tok + 4.0 * 4096 * 32 * kv_bytes / 131072.0) if full else 0.0
        roofline["tensor"] = {"flops": dec_flops, "achieved": dec_flops / dec_ms / 1e9 if dec_ms > 0 else 0.0,
                              "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                              "frac": dec_flops / dec_ms / 1e9 / peaks["tf_sustained"] if dec_ms > 0 else 0.0}
        extra["scheduler"] = stats
        cfg.update(requests_per_gpu=n_prot, slots=wl["slots"], prompt_len=wl["prompt_len"], max_new_tokens=wl["new_tokens"],
                   generated_tokens_per_step=n_tok, cuda_graph_decode=not args.no_graph)
        metric, unit = METRIC_TOKENS, "tokens/s"

    cpu_baseline = None
    if sd is not None:
        cpu_baseline = _cpu_baseline(sd, wl, torch)
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": cfg, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clk.result, "roofline": roofline, "cpu_baseline": cpu_baseline}
    line.update(extra)
    return line


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--size", default="full", choices=["full", "tiny"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true",
                    help="skip the stock-transformers GPU leg (gpu_reference key of the default c2 line)")
    ap.add_argument("--set", action="append", default=[], metavar="KEY=INT",
                    help="experiments only: override an integer field of the workload (e.g. --set batch=256); the "
                         "override is recorded in config.overrides")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload] if args.size == "full" else _shrink(WORKLOADS[args.workload])
    if args.set:
        wl = dict(wl)
        for kv in args.set:
            k, v = kv.split("=")
            if k not in wl or not isinstance(wl[k], int):
                ap.error(f"--set {kv}: {k!r} is not an integer field of workload {args.workload}")
            wl[k] = int(v)
        wl["desc"] += " [overrides: " + ", ".join(args.set) + "]"
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    import torch
    from opus_pllm_b200 import presets, synth

    if args.impl == "reference":
        if rank != 0:
            return 0
        # the reference arm never touches the GPU: weights of the same shapes / statistics come from a seeded CPU
        # generator (hashing 8 G values with int64 ops on the host would take minutes; only the timing matters here)
        with _fast_cpu_weights(torch, synth):
            sd = presets.synthetic_state_dicts(args.size, "cpu", with_lora=True)
        run_once, sample, cores, unit, metric = cpu_reference_runner(sd, wl, torch)
        del sd
        for _ in range(args.warmup):
            run_once()
        t0 = time.perf_counter()
        units = 0
        for _ in range(args.steps):
            units += run_once()
        dt = time.perf_counter() - t0
        v = units / dt
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": v, "unit": unit,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {wl['desc']}", "sample": sample},
            "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return 0

    # ---------------------------------------------------------------- B200 arm
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL writes its "NCCL version ..." banner to fd 1 when the first
        # communicator is created, so fd 1 points at stderr until that has happened
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    from opus_pllm_b200 import ops
    ops.device_check()
    peaks = _peaks()

    model, sd = presets.build_synthetic_model(args.size, dev, with_lora=True, keep_state=True)
    lcfg = sd["llama_cfg"]
    if args.no_cpu_baseline or rank != 0 or world != 1:
        sd = None
    torch.cuda.empty_cache()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        sync_all()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    if wl["kind"] != "generate":
        line = run_other_workload(args, wl, model, sd, lcfg, peaks, dev, rank, world, warmup, timed, sync_all, torch, dist)
        if rank == 0:
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return 0

    B, new = wl["batch"], wl["new_tokens"]
    seqs = synth.proteins(B, wl["protein_len"], seed=1234 + rank)
    prompts = synth.prompt_ids(B, wl["prompt_len"] - 7, vocab=lcfg["vocab"], seed=1234 + rank)
    ids_host = torch.stack(prompts).pin_memory()                    # [B, L] int64, one -200 sentinel per row
    use_graph = not args.no_graph

    # device-resident inputs for `value`
    pk = model.protein_encoder.tokenize(seqs)
    pk.device_arrays = tuple(ops.h2d(a, dev) for a in (pk.tokens, pk.pos, pk.scale, pk.cu))
    plan_sp = model._plan(ids_host, None, B)
    src_d = ops.h2d(plan_sp.src, dev)
    plan = model.llama.make_plan(plan_sp.cu, new)
    fused = None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    gathered = torch.empty((world * B, new), dtype=torch.int64, device=dev) if world > 1 else None

    def step_device(record=False):
        if record: ev[0].record()
        _, pooled_l2, _, _ = model.protein_encoder.encode(None, packed=pk)
        if record: ev[1].record()
        soft = model._fused(pooled_l2) if model._fused is not None else None
        if soft is None:
            soft = model._soft_tokens(seqs, None)
        embeds = ops.splice_gather(src_d, model.llama.embed, soft.reshape(-1, lcfg["dim"]))
        if record: ev[2].record()
        st = model.llama.prefill(embeds, plan=plan)
        if record: ev[3].record()
        out = model.llama.generate_from_prefill(st, new, use_graph=use_graph)
        if record: ev[4].record()
        return out

    def step_e2e():
        out = model.generate(ids_host, seqs, attention_mask=None, pad_token_id=128001, do_sample=False,
                             temperature=0, top_p=0.7, num_beams=1, max_new_tokens=new, use_cache=True,
                             use_graph=use_graph)
        return ops.d2h(out)

    def job_e2e(steps):
        def run():
            for _ in range(steps):
                last_out[:] = [step_e2e()]
            if world > 1:
                dist.all_gather_into_tensor(gathered, last_out[0].to(dev))
        return run

    model._soft_tokens(seqs[:2], None)  # builds the fused projector object

    # The path's ONE collective (an end-of-job gather of the generated ids, run_opus_ddp.py:138): issued once,
    # after the last step, inside the timed region. A per-step gather would make every step wait for the slowest of N
    # independently power-capped GPUs.
    last_out = []

    def job(steps):
        def run():
            for _ in range(steps):
                last_out[:] = [step_device()]
            if world > 1:
                dist.all_gather_into_tensor(gathered, last_out[0])
        return run

    for _ in range(warmup):
        step_device()
    ops.launch_count(reset=True)
    with ClockSampler(local_rank) as clk:
        ms = timed(job(args.steps), 1)
    launches = ops.launch_count(reset=True)
    tokens_per_step = world * B * new
    value = tokens_per_step * args.steps / (ms / 1e3)

    # phase breakdown of one more step (events inside the step; same stream)
    sync_all(); step_device(record=True); sync_all()
    phases = {"encoder_ms": ev[0].elapsed_time(ev[1]), "projector_splice_ms": ev[1].elapsed_time(ev[2]),
              "prefill_ms": ev[2].elapsed_time(ev[3]), "decode_ms": ev[3].elapsed_time(ev[4])}

    # e2e: host inputs, H2D + D2H inside the timed region
    for _ in range(2):
        step_e2e()
    ops.XFER["h2d_bytes"] = ops.XFER["d2h_bytes"] = 0
    ms_e2e = timed(job_e2e(args.steps), 1)
    e2e = {"value": tokens_per_step * args.steps / (ms_e2e / 1e3), "unit": "tokens/s",
           "h2d_bytes_per_step": ops.XFER["h2d_bytes"] // args.steps,
           "d2h_bytes_per_step": ops.XFER["d2h_bytes"] // args.steps, "ms_per_step": ms_e2e / args.steps}

    # dominant kernel: the prefill tcgen05 GEMM (4 launches per layer at M = B*prompt_len), timed in isolation
    from opus_pllm_b200 import _lib as L
    M = B * wl["prompt_len"]
    d, f, qn = lcfg["dim"], lcfg["ffn_dim"], (lcfg["n_q_heads"] + 2 * lcfg["n_kv_heads"]) * lcfg["head_dim"]
    layer0 = model.llama._keep[0]
    bufs = model.llama._ws_bufs
    shapes = [("qkv", bufs["xn"][:M], layer0["wqkv"], L.EPI_BF16, None, bufs["qkv"][:M]),
              ("o_proj", bufs["attn"][:M], layer0["wo"], L.EPI_RES_BF16, bufs["h"][:M], bufs["h"][:M]),
              ("gate_up", bufs["xn"][:M], layer0["wgu"], L.EPI_SWIGLU, None, bufs["act"][:M]),
              ("down", bufs["act"][:M], layer0["wdown"], L.EPI_RES_BF16, bufs["h"][:M], bufs["h"][:M])]
    for t in (bufs["xn"], bufs["attn"], bufs["act"], bufs["h"]):
        t[:M].normal_(0, 0.05)
    flops_total, ms_total, per_shape = 0.0, 0.0, {}
    for name, x, w, epi, res, out in shapes:
        run = lambda: ops.gemm(x, w, epilogue=epi, residual=res, out=out, transposed=False)  # noqa: E731
        for _ in range(3):
            run()
        reps = 5
        t = timed(lambda: run(), reps) / reps
        fl = 2.0 * M * w.shape[0] * w.shape[1]
        per_shape[name] = {"ms": t, "tflops": fl / t / 1e9}
        flops_total += fl; ms_total += t
    achieved = flops_total / ms_total / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("gemm_prefill_dram_bytes_per_launch")
    roofline = {"kernel": "tcgen05 GEMM, prefill linears at M=%d: gemm_bf16_2cta_kernel (cta_group::2; qkv, o_proj, down) + "
                          "gemm_bf16_tcgen05_kernel<256> (gate/up, SwiGLU epilogue)" % M, "bound": "tensor",
                "achieved": achieved, "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_burst"],
                "traffic": traffic, "peak_source": peaks["source"] + ", burst figure (kernel timed alone)",
                "per_shape": per_shape}
    # decode step vs HBM roofline (SURVEY.md §8d bytes): weights once per step + KV read + KV write
    steps_dec = max(new - 1, 1)
    ctx_mean = wl["prompt_len"] + (new + 1) / 2.0
    dec_bytes = 15009316864 * (lcfg["n_layers"] / 32.0 if args.size == "full" else 0) + 131072.0 * B * ctx_mean + 131072.0 * B
    dec_ms = phases["decode_ms"] / steps_dec
    roofline_decode = {"bound": "hbm", "achieved": dec_bytes / dec_ms / 1e6, "peak": peaks["hbm"], "unit": "GB/s",
                       "frac": dec_bytes / dec_ms / 1e6 / peaks["hbm"], "ms_per_decode_step": dec_ms,
                       "bytes_per_step": dec_bytes}
    # a decode step is 2 x 7.505 G FLOP per sequence (layer + lm_head weights) + the attention over the context: above
    # ~batch 300 the tensor time at the sustained peak exceeds the HBM time of the same step, so both bounds are reported
    dec_flops = (2.0 * 7504658432 * B + 4.0 * 4096 * 32 * B * ctx_mean) * (lcfg["n_layers"] / 32.0 if args.size == "full" else 0)
    t_hbm, t_tensor = dec_bytes / peaks["hbm"] / 1e6, dec_flops / peaks["tf_sustained"] / 1e9      # ms
    roofline_decode["tensor"] = {"flops_per_step": dec_flops, "achieved": dec_flops / dec_ms / 1e9, "peak": peaks["tf_sustained"],
                                 "unit": "TFLOP/s", "frac": dec_flops / dec_ms / 1e9 / peaks["tf_sustained"]}
    roofline_decode["tighter_bound"] = "tensor" if t_tensor > t_hbm else "hbm"
    roofline_decode["frac_of_tighter_bound"] = max(t_hbm, t_tensor) / dec_ms
    if new > 1 and world == 1:
        # the same decode loop (graph replays against the last prefill's cache) timed ALONE after a pause: inside the step
        # it inherits the power-capped SM clock of the prefill that precedes it, and ~35 % of a decode step is
        # launch / latency bound, i.e. scales with that clock. Not part of `value`.
        st_alone = model.llama.prefill(ops.splice_gather(src_d, model.llama.embed,
                                                         model._soft_tokens(seqs, None).reshape(-1, lcfg["dim"])), plan=plan)
        sync_all()
        time.sleep(1.5)
        for _ in range(2):
            model.llama.generate_from_prefill(st_alone, new, use_graph=use_graph)
        reps = 3
        t_alone = timed(lambda: model.llama.generate_from_prefill(st_alone, new, use_graph=use_graph), reps) / reps / steps_dec
        roofline_decode["alone"] = {"ms_per_decode_step": t_alone, "achieved": dec_bytes / t_alone / 1e6,
                                    "frac": dec_bytes / t_alone / 1e6 / peaks["hbm"],
                                    "note": "decode loop timed by itself (no prefill in front): SM clock not power-capped"}
    pre_flops = (2.0 * M * 6979321856 + 2.0 * B * 4096 * 128256 + 2.0 * 4096 * 32 * B * wl["prompt_len"] ** 2) if args.size == "full" else 0
    roofline_prefill = {"bound": "tensor", "achieved": pre_flops / phases["prefill_ms"] / 1e9, "peak": peaks["tf_sustained"],
                        "unit": "TFLOP/s", "frac": pre_flops / phases["prefill_ms"] / 1e9 / peaks["tf_sustained"]}
    n_res = pk.n_residues
    enc_flops = 2.0 * pk.n_tok * 648806400 + 4.0 * 1280 * 33 * float((pk.cu[1:] - pk.cu[:-1]).astype("float64").__pow__(2).sum())
    encoder = {"residues_per_s": world * n_res / (phases["encoder_ms"] / 1e3),
               "tflops": (enc_flops / phases["encoder_ms"] / 1e9) if args.size == "full" else None,
               "frac_of_sustained_peak": (enc_flops / phases["encoder_ms"] / 1e9 / peaks["tf_sustained"]) if args.size == "full" else None}

    cpu_baseline = None
    if sd is not None:
        cpu_baseline = _cpu_baseline(sd, wl, torch)
        del sd
    gpu_reference = None
    if rank == 0 and world == 1 and args.size == "full" and args.workload == "c2" and not args.no_gpu_reference:
        del model
        torch.cuda.empty_cache()
        gpu_reference = gpu_reference_leg(torch, wl)

    if rank == 0:
        print(json.dumps({
            "metric": METRIC_TOKENS, "value": value, "unit": "tokens/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {wl['desc']}", "prompts_per_gpu": B,
                       "protein_len": wl["protein_len"], "prompt_len": wl["prompt_len"], "new_tokens": new,
                       "weights": "random-init (hash-seeded) ESM-2-650M + CSTP/switch projectors + Llama-3-8B, LoRA r=16 merged at load",
                       "parallelism": f"dp{world} (replica per GPU, no collective on the model path, one NCCL all-gather "
                                      "of the generated ids at the end of the timed job)",
                       "l2": "inputs larger than L2 (15 GB of weights streamed per decode step, 126 MB L2)",
                       "cuda_graph_decode": use_graph},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk.result, "roofline": roofline,
            "roofline_decode": roofline_decode, "roofline_prefill": roofline_prefill, "encoder": encoder,
            "phases_ms": phases, "cpu_baseline": cpu_baseline, "gpu_reference": gpu_reference}))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
