"""What does each dependent step of a decode layer cost inside a PDL-chained CUDA graph? Times graphs of 16 repetitions of
growing kernel sequences (distinct weights per repetition, batch 64) and prints the increments."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opus_pllm_b200 import ops, _lib as L

B, D, F, QKV = int(os.environ.get("ROWS", "64")), 4096, 14336, 6144
NL = 16
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: (torch.randn(*s, device="cuda", generator=g) * 0.02).bfloat16()
wq = [rnd(QKV, D) for _ in range(NL)]
wo = [rnd(D, D) for _ in range(NL)]
wgu = [rnd(2 * F, D) for _ in range(NL)]
wd = [rnd(D, F) for _ in range(NL)]
gam = torch.ones(D, device="cuda").bfloat16()
x = rnd(B, D); attn = rnd(B, D); act = rnd(B, F)
h = rnd(B, D); xn = rnd(B, D)
part = torch.empty((4, B, QKV), dtype=torch.float32, device="cuda")
sumsq = torch.rand((D // 32, 64), device="cuda") * 32
P = L.EPI_PARTIAL_F32


cur = [xn]


def seq(steps):
    def body():
        cur[0] = xn
        for i in range(NL):
            for s in steps:
                if s == "qkv":
                    ops.gemm(x, wq[i], epilogue=P, transposed=True, split_k=3, out=part[:3, :, :QKV])
                elif s == "o":
                    ops.gemm(attn, wo[i], epilogue=P, transposed=True, split_k=4, out=part.view(-1)[: 4 * B * D].view(4, B, D))
                elif s == "norm":
                    cur[0] = ops.rmsnorm(None, gam, partial=part.view(-1)[: 4 * B * D].view(4, B, D), residual=h, h_out=h)
                elif s == "gu":
                    ops.gemm(cur[0], wgu[i], epilogue=L.EPI_SWIGLU, transposed=True, out=act)
                elif s == "o_fix":      # in-kernel split-K reduce + residual + sumsq
                    ops.gemm_fused(attn, wo[i], epilogue=L.EPI_RES_BF16, residual=h, out=h, split_k=4, splitk_fixup=True,
                                   sumsq_out=sumsq)
                elif s == "down_fix":
                    ops.gemm_fused(act, wd[i], epilogue=L.EPI_RES_BF16, residual=h, out=h, split_k=4, splitk_fixup=True,
                                   sumsq_out=sumsq)
                elif s == "gu_n":       # RMSNorm of the activation operand on load
                    ops.gemm_fused(h, wgu[i], epilogue=L.EPI_SWIGLU, out=act, norm_sumsq=sumsq, norm_gamma=gam)
                elif s == "qkv_n":
                    ops.gemm_fused(h, wq[i], epilogue=P, split_k=3, out=part[:3, :, :QKV], norm_sumsq=sumsq, norm_gamma=gam)
                elif s == "down":
                    ops.gemm(act, wd[i], epilogue=P, transposed=True, split_k=4, out=part.view(-1)[: 4 * B * D].view(4, B, D))
    body(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        body()
    for _ in range(3): gr.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): gr.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 10 / NL * 1e3


lib = L.load()
for warm in ((0, 1) if os.environ.get("WARM_AB") else (None,)):
    if warm is not None:
        L.check(lib.opus_set_tunable(b"epi_warm", warm))
        print(f"-- epi_warm = {warm}")
    for steps in (["o"], ["o", "norm"], ["gu"], ["o", "norm", "gu"], ["down"], ["gu", "down"], ["gu", "down", "norm"],
                  ["o", "norm", "gu", "down", "norm"], ["qkv"], ["o", "norm", "gu", "down", "norm", "qkv"],
                  ["o_fix"], ["gu_n"], ["down_fix"], ["qkv_n"], ["o_fix", "gu_n"], ["o_fix", "gu_n", "down_fix", "qkv_n"]):
        us = seq(steps)
        print(f"{'+'.join(steps):36s} {us:7.1f} us per repetition", flush=True)
