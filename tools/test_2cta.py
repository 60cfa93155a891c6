"""CTA-pair (cta_group::2) GEMM: correctness against torch and timing against the single-CTA kernel (tunable gemm_2cta)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opus_pllm_b200 import ops, _lib as L

lib = L.load()
def setk(v): L.check(lib.opus_set_tunable(b"gemm_2cta", v))

def check(M, N, K, epi=L.EPI_BF16, bias=False):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    x = (torch.randn(M, K, generator=g) * 1.0).bfloat16().cuda()
    w = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16().cuda()
    b = torch.randn(N, generator=g).cuda() if bias else None
    want = x.float() @ w.float().t() + (b if bias else 0)
    setk(1); got = ops.gemm(x, w, epilogue=epi, bias=b, transposed=False).float(); torch.cuda.synchronize()
    setk(0); ref = ops.gemm(x, w, epilogue=epi, bias=b, transposed=False).float(); torch.cuda.synchronize()
    if epi == L.EPI_BF16_GELU:
        want = torch.nn.functional.gelu(want)
    err = float((got - want).abs().max()); err1 = float((ref - want).abs().max())
    print(f"M {M} N {N} K {K} epi {epi} bias {bias}: 2cta max err {err:.4f} (1cta {err1:.4f}) equal-to-1cta {bool(torch.equal(got, ref))}", flush=True)
    assert err <= max(2 * err1, 0.05), "2-CTA result is wrong"

for shape in [(1024, 256, 64), (2048, 1024, 512), (1100, 768, 1280), (4096, 5120, 1280), (16512, 1280, 1280)]:
    check(*shape)
check(2048, 1024, 512, L.EPI_BF16, True)
check(3000, 2560, 1280, L.EPI_BF16_GELU, True)

def t(fn, reps=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

for name, M, N, K, epi in [("qkv", 32768, 6144, 4096, L.EPI_BF16), ("gate_up", 32768, 28672, 4096, L.EPI_SWIGLU),
                           ("down", 32768, 4096, 14336, L.EPI_BF16), ("e_qkv", 16512, 3840, 1280, L.EPI_BF16),
                           ("e_fc1", 16512, 5120, 1280, L.EPI_BF16_GELU), ("e_fc2", 16512, 1280, 5120, L.EPI_BF16)]:
    x = torch.randn(M, K, device="cuda").bfloat16() * 0.05
    w = torch.randn(N, K, device="cuda").bfloat16() * 0.02
    res = {}
    for mode in (0, 1):
        setk(mode)
        ms = t(lambda: ops.gemm(x, w, epilogue=epi, transposed=False))
        res[mode] = (round(ms, 3), round(2.0 * M * N * K / ms / 1e9))
    print(name, "1cta", res[0], "2cta", res[1], flush=True)
setk(0)
