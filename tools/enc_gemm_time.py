"""CUDA-event timing of the four encoder GEMM launches of one ESM-2 layer at C1 size (M = 16512), L2-warm like in the
encoder (the A operand was just written by the previous kernel). OPUS_B200_LIB selects another build for A/B runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opus_pllm_b200 import ops, _lib as L
M = int(sys.argv[1]) if len(sys.argv) > 1 else 16512
x1 = torch.randn(M, 1280, device="cuda").bfloat16(); x5 = torch.randn(M, 5120, device="cuda").bfloat16()
shapes = [("qkv", x1, 3840, L.EPI_BF16), ("out", x1, 1280, L.EPI_BF16), ("fc1", x1, 5120, L.EPI_BF16_GELU), ("fc2", x5, 1280, L.EPI_BF16)]
ws = {n: (torch.randn(N, x.shape[1], device="cuda").bfloat16() * 0.02, torch.randn(N, device="cuda")) for n, x, N, _ in shapes}
outs = {n: torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for n, x, N, _ in shapes}
res = []
for n, x, N, epi in shapes:
    run = lambda: ops.gemm(x, ws[n][0], epilogue=epi, bias=ws[n][1], out=outs[n], transposed=False)  # noqa: E731
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(40):
        run()
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 40 * 1e3
    res.append(f"{n} {us:6.1f} us {2.0 * M * N * x.shape[1] / us / 1e6:5.0f} TF")
print(os.path.basename(L.LIB_PATH), " | ".join(res), flush=True)
