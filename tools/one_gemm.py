"""A handful of decode-sized gate/up (SwiGLU) swap-AB GEMM launches, for ncu source-level captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opus_pllm_b200 import ops, _lib as L
rows = int(os.environ.get("ROWS", "64"))
ws = [torch.randn(28672, 4096, device="cuda").bfloat16() * 0.02 for _ in range(3)]
x = torch.randn(rows, 4096, device="cuda").bfloat16()
for i in range(6):
    ops.gemm(x, ws[i % 3], epilogue=L.EPI_SWIGLU, transposed=True)
torch.cuda.synchronize()
print("done")
