"""A/B of the CTA-pair swap-AB kernel (tunable gemm_2cta_tr) on the decode loop at large batch.
usage: python tools/decode_pair.py [B] [T] [new]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from opus_pllm_b200 import _lib as L, presets, synth
from opus_pllm_b200.llama import B200Llama

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 128
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
outs = {}
for mode in (0, 1, 0, 1):
    L.check(lib.opus_set_tunable(b"gemm_2cta_tr", mode))
    for _ in range(2):
        out = ll.generate_from_prefill(st, new)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        out = ll.generate_from_prefill(st, new)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3 / (new - 1)
    outs[mode] = out.clone()
    print(f"B={B} ctx={T} gemm_2cta_tr={mode}: {ms:7.3f} ms/step  {bytes_step / ms / 1e6:7.0f} GB/s  "
          f"frac {bytes_step / ms / 1e6 / 6551.4:.3f}", flush=True)
print("tokens identical:", bool(torch.equal(outs[0], outs[1])))
