"""down-proj prefill GEMM (K = 14336) once per mode for ncu: single-CTA vs CTA-pair."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opus_pllm_b200 import ops, _lib as L
lib = L.load()
x = torch.randn(32768, 14336, device="cuda").bfloat16() * 0.05
w = torch.randn(4096, 14336, device="cuda").bfloat16() * 0.02
for mode in (0, 1, 0, 1):
    L.check(lib.opus_set_tunable(b"gemm_2cta", mode))
    ops.gemm(x, w, epilogue=L.EPI_BF16, transposed=False)
torch.cuda.synchronize()
print("done")
