"""Isolated timing of the prefill / encoder GEMM shapes (CUDA events), for raster / tile-shape sweeps via env vars."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opus_pllm_b200 import ops, _lib as L

def t(fn, reps=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

def main():
    dev = "cuda"
    out = {}
    M = 32768
    x4 = torch.randn(M, 4096, device=dev).bfloat16() * 0.05
    x14 = torch.randn(M, 14336, device=dev).bfloat16() * 0.05
    res = torch.randn(M, 4096, device=dev).bfloat16()
    shapes = [("qkv", x4, 6144, L.EPI_BF16, None), ("o", x4, 4096, L.EPI_RES_BF16, res),
              ("gate_up", x4, 28672, L.EPI_SWIGLU, None), ("down", x14, 4096, L.EPI_RES_BF16, res)]
    for name, x, N, epi, r in shapes:
        w = torch.randn(N, x.shape[1], device=dev).bfloat16() * 0.02
        ms = t(lambda: ops.gemm(x, w, epilogue=epi, residual=r, transposed=False))
        out[name] = (round(ms, 3), round(2.0 * M * N * x.shape[1] / ms / 1e9))
    Me = 16512
    e1 = torch.randn(Me, 1280, device=dev).bfloat16(); e5 = torch.randn(Me, 5120, device=dev).bfloat16()
    rf = torch.randn(Me, 1280, device=dev)
    for name, x, N, epi, r in [("e_qkv", e1, 3840, L.EPI_BF16, None), ("e_out", e1, 1280, L.EPI_RES_F32, rf),
                               ("e_fc1", e1, 5120, L.EPI_BF16_GELU, None), ("e_fc2", e5, 1280, L.EPI_RES_F32, rf)]:
        w = torch.randn(N, x.shape[1], device=dev).bfloat16() * 0.02
        b = torch.randn(N, device=dev)
        for bn in ([0] if "BN" not in os.environ else [int(os.environ["BN"])]):
            ms = t(lambda: ops.gemm(x, w, epilogue=epi, bias=b, residual=r, out=(r if r is not None else None),
                                    transposed=False, block_n=bn))
            out[name] = (round(ms, 3), round(2.0 * Me * N * x.shape[1] / ms / 1e9))
    print(os.environ.get("OPUS_GEMM_GROUP_M", "16"), os.environ.get("BN", "auto"), out)

main()
