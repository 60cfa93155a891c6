"""In-graph (PDL chain) timing of the decode-sized swap-AB GEMMs: N launches of one shape over distinct weights, replayed
as a CUDA graph. Reports us/launch and GB/s, i.e. what each GEMM costs inside the decode step rather than alone."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opus_pllm_b200 import ops, _lib as L

rows = int(os.environ.get("ROWS", "64"))
NL = 16


def bench(name, n_out, K, split, epi, block_n=0, nw=None):
    nw = nw or max(2, int(600e6 // (n_out * K * 2)) + 1)
    ws = [torch.randn(n_out, K, device="cuda").bfloat16() * 0.02 for _ in range(nw)]
    x = torch.randn(rows, K, device="cuda").bfloat16()
    if epi == L.EPI_PARTIAL_F32:
        out = torch.empty((max(split, 1), rows, n_out), dtype=torch.float32, device="cuda")
    else:
        out = torch.empty((rows, n_out // 2 if epi == L.EPI_SWIGLU else n_out), dtype=torch.bfloat16, device="cuda")
    def body():
        for i in range(NL):
            ops.gemm(x, ws[i % nw], epilogue=epi, transposed=True, split_k=split, out=out, block_n=block_n)
    body(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): g.replay()
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 10 / NL * 1e3
    print(f"{name:34s} tiles {n_out // 128:5d} K {K:5d} split {split:2d} bn {block_n:3d}: {us:7.1f} us  {n_out * K * 2 / us / 1e3:6.0f} GB/s", flush=True)
    del ws


P = L.EPI_PARTIAL_F32
if os.environ.get("L2_TEST"):
    # same weight every launch (L2-resident after the first) vs distinct weights (HBM): what does a boundary cost?
    for nw in (1, None):
        print("weights:", "one buffer (L2 hits)" if nw else "distinct buffers (HBM)")
        bench("o_proj s4", 4096, 4096, 4, P, nw=nw)
        bench("qkv s3", 6144, 4096, 3, P, nw=nw)
        bench("16 tiles s8", 2048, 4096, 8, P, nw=nw)
    sys.exit(0)
if os.environ.get("BN_TEST"):
    # batch 129..256: one 256-wide batch tile (4 stages) vs two 128-wide tiles sharing the weight panel through L2
    for bn in (256, 128):
        print("block_n", bn)
        bench("qkv s3", 6144, 4096, 3, P, block_n=bn)
        bench("o_proj s4", 4096, 4096, 4, P, block_n=bn)
        bench("gate_up swiglu", 28672, 4096, 1, L.EPI_SWIGLU, block_n=bn)
        bench("down s4", 4096, 14336, 4, P, block_n=bn)
        bench("lm_head", 128256, 4096, 1, L.EPI_BF16, block_n=bn)
    sys.exit(0)
if os.environ.get("DEPTH_TEST"):
    # weight bytes in flight per SM: bn 64 -> 8 stages x 16 KB, bn 128 -> 6 x 16 KB, bn 256 -> 4 x 16 KB
    for bn in (64, 128, 256):
        bench("lm_head", 128256, 4096, 1, L.EPI_BF16, block_n=bn)
        bench("gate_up bf16", 28672, 4096, 1, L.EPI_BF16, block_n=bn)
        bench("down s4", 4096, 14336, 4, P, block_n=bn)
    sys.exit(0)
if os.environ.get("SK_ONLY"):
    for fill in (0, 90):
        L.check(L.load().opus_set_tunable(b"streamk_fill", fill))
        print("streamk_fill", fill)
        bench("gate_up swiglu", 28672, 4096, 1, L.EPI_SWIGLU)
        bench("gate_up bf16", 28672, 4096, 1, L.EPI_BF16)
        bench("lm_head", 128256, 4096, 1, L.EPI_BF16)
        bench("150 tiles", 150 * 128, 4096, 1, L.EPI_BF16)
        bench("185 tiles", 185 * 128, 4096, 1, L.EPI_BF16)
        bench("280 tiles", 280 * 128, 4096, 1, L.EPI_BF16)
    sys.exit(0)
for split in (1, 2, 3, 4, 6):
    bench("qkv", 6144, 4096, split, P)
for split in (1, 2, 4, 8):
    bench("o_proj", 4096, 4096, split, P)
bench("gate_up swiglu", 28672, 4096, 1, L.EPI_SWIGLU)
bench("gate_up bf16", 28672, 4096, 1, L.EPI_BF16)
bench("gate_up partial s2", 28672, 4096, 2, P)
bench("296 tiles bf16", 296 * 128, 4096, 1, L.EPI_BF16)
bench("148 tiles bf16", 148 * 128, 4096, 1, L.EPI_BF16)
for split in (1, 2, 4, 8, 9):
    bench("down", 4096, 14336, split, P)
bench("lm_head", 128256, 4096, 1, L.EPI_BF16)
