"""Steady-state (power-capped) timing of the Llama-3-8B prefill alone: B x T prompt embeddings, repeated back to back.
Env knobs are read by the library at first use (OPUS_GEMM_GROUP_M, OPUS_ATTN, OPUS_PDL)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from opus_pllm_b200 import presets, synth
from opus_pllm_b200.llama import B200Llama

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 512
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
cfg = presets.LLAMA3_8B
sd = synth.llama_weights(cfg["n_layers"], cfg["dim"], cfg["n_q_heads"], cfg["n_kv_heads"], cfg["head_dim"],
                         cfg["ffn_dim"], cfg["vocab"], seed=0, peaked=False, dtype=torch.bfloat16, device="cuda")
ll = B200Llama(sd, **cfg, device="cuda")
del sd
torch.cuda.empty_cache()
cu = np.arange(B + 1, dtype=np.int32) * T
emb = (torch.randn(B * T, cfg["dim"], device="cuda") * 0.02).bfloat16()
plan = ll.make_plan(cu, 4)
for _ in range(3):
    ll.prefill(emb, plan=plan)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    ll.prefill(emb, plan=plan)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
flops = 2.0 * B * T * 6979321856 + 2.0 * B * 4096 * 128256 + 2.0 * 4096 * 32 * B * T * T
print(f"group_m={os.environ.get('OPUS_GEMM_GROUP_M', 'default')} attn={os.environ.get('OPUS_ATTN', 'auto')}: "
      f"prefill {B}x{T}: {ms:.1f} ms  {flops / ms / 1e9:.0f} TFLOP/s", flush=True)
