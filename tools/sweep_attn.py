"""Attention kernel timing (CUDA events): prefill (causal GQA hd 128), encoder (hd 64, T = 258) and long proteins."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opus_pllm_b200 import ops

def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

out = {}
for name, B, T, Hq, Hkv, D, causal in [("prefill", 64, 512, 32, 8, 128, True), ("encoder258", 64, 258, 20, 20, 64, False),
                                       ("encoder2048", 16, 2048, 20, 20, 64, False), ("prefill2048", 8, 2048, 32, 8, 128, True)]:
    n = B * T
    qkv = torch.randn(n, (Hq + 2 * Hkv) * D, device="cuda").bfloat16()
    cu = (torch.arange(B + 1, dtype=torch.int32) * T).cuda()
    q, k, v = qkv[:, : Hq * D], qkv[:, Hq * D: (Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D:]
    us = t(lambda: ops.attn_varlen(q, k, v, cu, T, Hq, Hkv, D, causal, D ** -0.5))
    fl = 4.0 * B * Hq * T * T * D * (0.5 if causal else 1.0)
    out[name] = (round(us, 1), round(fl / us / 1e6))
print(os.environ.get("OPUS_ATTN", "auto"), out)
