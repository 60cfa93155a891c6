"""GPU-side cost of swap-AB GEMM launches by epilogue (run under ncu): 32 tiles, K = 4096."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opus_pllm_b200 import ops, _lib as L
K = 4096
x = torch.randn(64, K, device="cuda").bfloat16()
w = torch.randn(4096, K, device="cuda").bfloat16()
res = torch.randn(64, 4096, device="cuda").bfloat16()
for rep in range(3):
    ops.gemm(x, w, epilogue=L.EPI_PARTIAL_F32, transposed=True, split_k=1)   # 0
    ops.gemm(x, w, epilogue=L.EPI_F32, transposed=True)                      # 1
    ops.gemm(x, w, epilogue=L.EPI_BF16, transposed=True)                     # 2
    ops.gemm(x, w, epilogue=L.EPI_SWIGLU, transposed=True)                   # 3
    ops.gemm(x, w, epilogue=L.EPI_RES_BF16, residual=res, transposed=True)   # 4
torch.cuda.synchronize()
print("done")
