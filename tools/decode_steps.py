"""Per-step time series of the decode loop right after a prefill (as inside bench.py) and after a pause: does the step time
recover while the decode runs (power-cap lag), or is it flat?  usage: python tools/decode_steps.py [B] [T] [new]"""
import ctypes as C, os, sys, time
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
stream = torch.cuda.current_stream().cuda_stream


def series(label, pause):
    for _ in range(2):                      # warm: two prefills back to back (power-capped state)
        st = ll.prefill(emb, plan=plan)
    s, bufs = ll._decode_state(st, new, (), 0)
    L.check(lib.opus_llama_select(C.byref(ll._model), C.byref(ll._ws), C.byref(s), B, stream))
    if pause:
        torch.cuda.synchronize(); time.sleep(pause)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(new)]
    ev[0].record()
    for i in range(1, new):
        L.check(lib.opus_llama_decode_loop(C.byref(ll._model), C.byref(ll._cache), C.byref(ll._ws), C.byref(s), B, 1, 0, 1,
                                           stream))
        ev[i].record()
    torch.cuda.synchronize()
    ms = [ev[i - 1].elapsed_time(ev[i]) for i in range(1, new)]
    print(f"{label}: mean {sum(ms) / len(ms):.3f} ms | " + " ".join(f"{m:.2f}" for m in ms), flush=True)


series("warm-up", 0)
series("right after prefill", 0)
series("right after prefill", 0)
series("after a 2 s pause", 2.0)
