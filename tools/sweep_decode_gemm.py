"""Weight-streaming (swap-AB) GEMM timing at decode batch sizes: GB/s vs number of 128-row tiles / split-K."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opus_pllm_b200 import ops, _lib as L

def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

rows = int(os.environ.get("ROWS", "64"))
res = []
for n_out, K, split, epi in [(148 * 128, 4096, 1, L.EPI_BF16), (2 * 148 * 128, 4096, 1, L.EPI_BF16),
                             (224 * 128, 4096, 1, L.EPI_BF16), (224 * 128, 4096, 1, L.EPI_SWIGLU),
                             (224 * 128, 4096, 2, L.EPI_PARTIAL_F32), (4 * 148 * 128, 4096, 1, L.EPI_BF16),
                             (4096, 14336, 4, L.EPI_PARTIAL_F32), (4096, 14336, 9, L.EPI_PARTIAL_F32),
                             (6144, 4096, 3, L.EPI_PARTIAL_F32), (6144, 4096, 6, L.EPI_PARTIAL_F32),
                             (4096, 4096, 4, L.EPI_PARTIAL_F32), (4096, 4096, 9, L.EPI_PARTIAL_F32),
                             (128256, 4096, 1, L.EPI_BF16)]:
    # several distinct weight buffers so consecutive launches never hit L2
    ws = [torch.randn(n_out, K, device="cuda").bfloat16() * 0.02 for _ in range(max(2, int(400e6 // (n_out * K * 2)) + 1))]
    x = torch.randn(rows, K, device="cuda").bfloat16()
    i = [0]
    def run():
        w = ws[i[0] % len(ws)]; i[0] += 1
        ops.gemm(x, w, epilogue=epi, transposed=True, split_k=split)
    us = t(run)
    res.append((n_out // 128, K, split, epi, round(us, 1), round(n_out * K * 2 / us / 1e3)))
for r in res: print("tiles %5d K %5d split %2d epi %d : %7.1f us  %5d GB/s" % r)
