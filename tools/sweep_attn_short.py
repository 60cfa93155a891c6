"""Short-sequence attention: the mma.sync kernel (attn_mode 1) against the tcgen05 kernel (attn_mode 2), CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opus_pllm_b200 import ops, _lib as L
lib = L.load()
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for name, B, T, Hq, Hkv, D, causal in [("prefill T=64", 256, 64, 32, 8, 128, True), ("prefill T=96", 256, 96, 32, 8, 128, True),
                                       ("prefill T=128", 256, 128, 32, 8, 128, True), ("prefill T=160", 256, 160, 32, 8, 128, True),
                                       ("prefill T=110 ragged", 256, 110, 32, 8, 128, True),
                                       ("encoder T=66", 64, 66, 20, 20, 64, False), ("encoder T=130", 64, 130, 20, 20, 64, False)]:
    n = B * T
    qkv = torch.randn(n, (Hq + 2 * Hkv) * D, device="cuda").bfloat16()
    cu = (torch.arange(B + 1, dtype=torch.int32) * T).cuda()
    q, k, v = qkv[:, : Hq * D], qkv[:, Hq * D: (Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D:]
    res = {}
    outs = {}
    for mode in (1, 2):
        L.check(lib.opus_set_tunable(b"attn_mode", mode))
        outs[mode] = ops.attn_varlen(q, k, v, cu, T, Hq, Hkv, D, causal, D ** -0.5).float()
        res[mode] = t(lambda: ops.attn_varlen(q, k, v, cu, T, Hq, Hkv, D, causal, D ** -0.5))
    L.check(lib.opus_set_tunable(b"attn_mode", 0))
    print(f"{name:22s} mma.sync {res[1]:7.1f} us   tcgen05 {res[2]:7.1f} us   max|diff| {float((outs[1] - outs[2]).abs().max()):.4f}", flush=True)
