"""The four encoder GEMM launches of one ESM-2 layer at C1 size (M = 16512), as the encoder issues them (bf16 epilogues)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opus_pllm_b200 import ops, _lib as L
M = 16512
x1 = torch.randn(M, 1280, device="cuda").bfloat16(); x5 = torch.randn(M, 5120, device="cuda").bfloat16()
shapes = [("qkv", x1, 3840, L.EPI_BF16), ("out", x1, 1280, L.EPI_BF16), ("fc1", x1, 5120, L.EPI_BF16_GELU), ("fc2", x5, 1280, L.EPI_BF16)]
ws = {n: (torch.randn(N, x.shape[1], device="cuda").bfloat16() * 0.02, torch.randn(N, device="cuda")) for n, x, N, _ in shapes}
for rep in range(3):
    for n, x, N, epi in shapes:
        ops.gemm(x, ws[n][0], epilogue=epi, bias=ws[n][1], transposed=False)
torch.cuda.synchronize()
print("done")
