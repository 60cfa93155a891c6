"""Encoder throughput (BASELINE metric 2, residues/s) at C1 (64 x 256 aa) and C4 (256 x U[1024,2048] aa, packed varlen):
ESM-2-650M forward + final LN + mean-pool + L2, device-resident packed tokens, CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opus_pllm_b200 import ops, presets, synth
from opus_pllm_b200.encoder import B200ProteinEncoder

cfg = presets.ESM2_650M
enc = B200ProteinEncoder(synth.esm2_weights(cfg["n_layers"], cfg["dim"], cfg["ffn_dim"], seed=0, device="cuda"), **cfg)
for name, seqs in [("c1: 64 x 256 aa", synth.proteins(64, 256)),
                   ("c4: 256 x U[1024,2048] aa", synth.proteins(256, 1024, 2048)),
                   ("c4/4: 64 x U[1024,2048] aa", synth.proteins(64, 1024, 2048))]:
    pk = enc.tokenize(seqs)
    pk.device_arrays = tuple(ops.h2d(a, "cuda") for a in (pk.tokens, pk.pos, pk.scale, pk.cu))
    for _ in range(2):
        enc.encode(None, packed=pk)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    a.record()
    for _ in range(reps):
        enc.encode(None, packed=pk)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    lens = (pk.cu[1:] - pk.cu[:-1]).astype("float64")
    flops = 2.0 * pk.n_tok * 648806400 + 4.0 * 1280 * 33 * float((lens ** 2).sum())
    attn = 4.0 * 1280 * 33 * float((lens ** 2).sum()) / flops
    print(f"{name:28s} {pk.n_tok:7d} tokens  {ms:8.1f} ms  {pk.n_residues / ms * 1e3 / 1e3:8.1f} k residues/s  "
          f"{flops / ms / 1e9:6.0f} TFLOP/s ({100 * flops / ms / 1e9 / 1383.8:.0f} % of sustained peak; attention = {100 * attn:.0f} % of the FLOPs)", flush=True)
