"""Raster sweeps of the two big prefill GEMMs at M = 32768 (CUDA events, back-to-back launches):
gate/up (single-CTA kernel, grouped-M raster, tunable group_m) and down (CTA-pair kernel, grouped-M vs grouped-N bands)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opus_pllm_b200 import ops, _lib as L

lib = L.load()
def setk(**kw):
    for k, v in kw.items():
        L.check(lib.opus_set_tunable(k.encode(), v), k)

def t(fn, reps=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

M = 32768
x4 = torch.randn(M, 4096, device="cuda").bfloat16() * 0.05
x14 = torch.randn(M, 14336, device="cuda").bfloat16() * 0.05
res = torch.randn(M, 4096, device="cuda").bfloat16()
wgu = torch.randn(28672, 4096, device="cuda").bfloat16() * 0.02
wd = torch.randn(4096, 14336, device="cuda").bfloat16() * 0.02
ref = None
for gm in (0, 48, 64):
    setk(group_m=gm)
    ms = t(lambda: ops.gemm(x4, wgu, epilogue=L.EPI_SWIGLU, transposed=False))
    print(f"gate_up group_m={gm or 32:3d}: {ms:.3f} ms {2.0 * M * 28672 * 4096 / ms / 1e9:.0f} TFLOP/s", flush=True)
setk(group_m=0)
for gn, hints in ((0, 1), (8, 1), (8, 0), (4, 1), (16, 1)):
    setk(group_n=gn, group_n_hints=hints)
    out = res.clone()
    y = ops.gemm(x14, wd, epilogue=L.EPI_RES_BF16, residual=res, out=out, transposed=False)
    if ref is None: ref = y.clone()
    same = bool(torch.equal(y, ref))
    ms = t(lambda: ops.gemm(x14, wd, epilogue=L.EPI_RES_BF16, residual=res, out=out, transposed=False))
    print(f"down group_n={gn:2d} hints={hints}: {ms:.3f} ms {2.0 * M * 4096 * 14336 / ms / 1e9:.0f} TFLOP/s bit-identical={same}", flush=True)
setk(group_n=0)
