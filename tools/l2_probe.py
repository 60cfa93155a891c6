"""Is the swap-AB decode GEMM limited by L2 -> SM traffic (weights + re-read activation tiles) rather than by HBM?
Times a CUDA graph of 16 gate/up-sized swap-AB launches over distinct weights at several batch sizes (activation bytes
per 16 KB weight tile: 4 KB at rows <= 32, 8 KB at 64, 16 KB at 128, 32 KB at 256), once with a cool GPU (boost clock)
and once right after 300 ms of tensor-bound GEMMs (the power-capped clock the decode loop inherits from the prefill)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opus_pllm_b200 import ops, _lib as L

NL = 16
big_a = torch.randn(16384, 4096, device="cuda").bfloat16()
big_b = torch.randn(8192, 4096, device="cuda").bfloat16()
big_o = torch.empty(16384, 8192, device="cuda", dtype=torch.bfloat16)


def heat(ms=300):
    t0 = time.time()
    while (time.time() - t0) * 1e3 < ms:
        for _ in range(20):
            ops.gemm(big_a, big_b, out=big_o)
        torch.cuda.synchronize()


def graph_of(rows, n_out, K, epi, split):
    nw = max(3, int(700e6 // (n_out * K * 2)) + 1)
    ws = [torch.randn(n_out, K, device="cuda").bfloat16() * 0.02 for _ in range(nw)]
    x = torch.randn(rows, K, device="cuda").bfloat16()
    if epi == L.EPI_PARTIAL_F32:
        out = torch.empty((max(split, 1), rows, n_out), dtype=torch.float32, device="cuda")
    else:
        out = torch.empty((rows, n_out // 2 if epi == L.EPI_SWIGLU else n_out), dtype=torch.bfloat16, device="cuda")

    def body():
        for i in range(NL):
            ops.gemm(x, ws[i % nw], epilogue=epi, transposed=True, split_k=split, out=out)
    body(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    g.replay(); torch.cuda.synchronize()
    return g, (ws, x, out)


def timed(g, reps=6):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / NL * 1e3


for name, n_out, K, epi, split in (("gate_up", 28672, 4096, L.EPI_SWIGLU, 1), ("down s4", 4096, 14336, L.EPI_PARTIAL_F32, 4),
                                   ("o_proj s4", 4096, 4096, L.EPI_PARTIAL_F32, 4)):
    for rows in (16, 32, 64, 128, 256):
        g, keep = graph_of(rows, n_out, K, epi, split)
        time.sleep(1.0)
        timed(g, 2)
        cool = timed(g)
        heat()
        hot = timed(g)
        gb = n_out * K * 2 / 1e3
        print(f"{name:10s} rows {rows:3d}: cool {cool:7.1f} us {gb / cool:6.0f} GB/s | right after 300 ms of GEMMs {hot:7.1f} us {gb / hot:6.0f} GB/s",
              flush=True)
        del g, keep
        torch.cuda.empty_cache()
