"""Decode-only timing (CUDA-graph replays) of the Llama-3-8B greedy loop, with run-time sweeps of the L2-prefetch depths.

usage: python tools/bench_decode.py [B] [T] [new]     env SWEEP=1 sweeps the opus_set_tunable knobs.
"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from opus_pllm_b200 import _lib as L, presets, synth
from opus_pllm_b200.llama import B200Llama

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 512
new = int(sys.argv[3]) if len(sys.argv) > 3 else 32
cfg = presets.LLAMA3_8B
sd = synth.llama_weights(cfg["n_layers"], cfg["dim"], cfg["n_q_heads"], cfg["n_kv_heads"], cfg["head_dim"],
                         cfg["ffn_dim"], cfg["vocab"], seed=0, peaked=False, dtype=torch.bfloat16, device="cuda")
ll = B200Llama(sd, **cfg, device="cuda")
del sd
torch.cuda.empty_cache()
lib = L.load()
cu = np.arange(B + 1, dtype=np.int32) * T
emb = (torch.randn(B * T, cfg["dim"], device="cuda") * 0.02).bfloat16()
plan = ll.make_plan(cu, new)
st = ll.prefill(emb, plan=plan)
bytes_step = 15009316864 + 131072.0 * B * (T + (new + 1) / 2.0) + 131072.0 * B


def run(label):
    for _ in range(2):
        ll.generate_from_prefill(st, new)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    a.record()
    for _ in range(reps):
        ll.generate_from_prefill(st, new)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps / (new - 1)
    print(f"{label:40s} {ms:7.3f} ms/step  {bytes_step / ms / 1e6:7.0f} GB/s  frac {bytes_step / ms / 1e6 / 6551.4:.3f}", flush=True)
    return ms


def setk(**kw):
    for k, v in kw.items():
        L.check(lib.opus_set_tunable(k.encode(), v), k)


def after_prefill(label, gap_ms=0.0):
    # the same decode loop timed right after a prefill, as inside bench.py (clock / power state carried over);
    # gap_ms > 0: the GPU idles that long between the two (is the power cap released faster when idle?) -- the gap is
    # INSIDE the timed region
    import time
    tot = 0.0
    for _ in range(3):
        st2 = ll.prefill(emb, plan=plan)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if gap_ms > 0:
            torch.cuda.synchronize()
            time.sleep(gap_ms / 1e3)
        ll.generate_from_prefill(st2, new)
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    ms = tot / 3 / (new - 1)
    print(f"{label:40s} {ms:7.3f} ms/step  {bytes_step / ms / 1e6:7.0f} GB/s  frac {bytes_step / ms / 1e6 / 6551.4:.3f}", flush=True)


if os.environ.get("SETK"):      # e.g. SETK="wide_overhead=4,gemm_2cta_tr=0"
    setk(**{k: int(v) for k, v in (kv.split("=") for kv in os.environ["SETK"].split(","))})
run("defaults (kernel per op, PDL chain)" if not os.environ.get("SETK") else os.environ["SETK"])
if os.environ.get("AFTER_PREFILL"):
    after_prefill("decode timed right after a prefill")
if os.environ.get("L2AHEAD"):
    for d in (4, 8, 16, 32):
        setk(l2_ahead=d); run(f"l2_ahead={d}")
        if os.environ.get("AFTER_PREFILL"):
            after_prefill(f"l2_ahead={d} right after a prefill")
    setk(l2_ahead=0)
if os.environ.get("GAPS"):
    for gms in (2, 5, 10, 20, 40):
        after_prefill(f"after a prefill + {gms} ms idle gap (gap included)", gms)
if os.environ.get("PAIR_GU"):
    setk(gemm_2cta_tr=2); run("gemm_2cta_tr=2 (gate/up on the CTA-pair kernel)"); setk(gemm_2cta_tr=1)
if os.environ.get("NORMF"):
    setk(decode_norm_fused=1); run("decode_norm_fused=1 (5 launches per layer)")
    if os.environ.get("AFTER_PREFILL"):
        after_prefill("norm-fused right after a prefill")
    setk(decode_norm_fused=0)
if os.environ.get("AFTER_PREFILL"):
    pass
    if os.environ.get("FUSED"):
        setk(decode_fused=1); run("decode_fused=1 (chain kernel)"); after_prefill("chain kernel right after a prefill")
        setk(decode_fused=0)
if os.environ.get("FUSED") and not os.environ.get("AFTER_PREFILL"):
    setk(decode_fused=1); run("decode_fused=1 (chain kernel)"); setk(decode_fused=0)
if os.environ.get("SWEEP"):
    setk(pf_qkv=0, pf_o=0, pf_gu=0, pf_down=0, pf_lm=0); base = run("all off")
    for name in ("pf_qkv", "pf_o", "pf_gu", "pf_down", "pf_lm"):
        for v in (8, 16, 32, 64):
            setk(**{name: v}); run(f"only {name}={v}")
        setk(**{name: 0})
    for combo in [dict(pf_qkv=64, pf_o=64, pf_gu=8, pf_down=8, pf_lm=8), dict(pf_qkv=64, pf_o=64, pf_gu=16, pf_down=16, pf_lm=16),
                  dict(pf_qkv=64, pf_o=64, pf_gu=32, pf_down=32, pf_lm=16), dict(pf_qkv=64, pf_o=64, pf_gu=64, pf_down=56, pf_lm=32),
                  dict(pf_qkv=32, pf_o=32, pf_gu=32, pf_down=32, pf_lm=16), dict(pf_qkv=16, pf_o=16, pf_gu=16, pf_down=16, pf_lm=16)]:
        setk(**combo); run(str(combo).replace("'", "").replace("pf_", ""))
