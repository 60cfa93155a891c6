"""Short profiling driver (for ncu): build the full-size synthetic model, run ONE hot-path step with a short decode,
without CUDA graphs, so a launch list / --set full capture stays cheap.

  python tools/profile_step.py [--new 4] [--batch 64]
Kernel launch order (opus:: kernels only): 224 LoRA merges at load, then per step: encoder (33*8+2), projector 3,
splice 1, prefill 32*8+3, select 1, then (new-1) decode steps of 32*8+5 launches each.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from opus_pllm_b200 import ops, presets, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--new", type=int, default=4)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--prompt", type=int, default=512)
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
model = presets.build_synthetic_model("full", "cuda", with_lora=True)
seqs = synth.proteins(a.batch, 256)
ids = torch.stack(synth.prompt_ids(a.batch, a.prompt - 7))
torch.cuda.synchronize()
ops.launch_count(reset=True)
for _ in range(a.steps):
    out = model.generate(ids, seqs, pad_token_id=128001, do_sample=False, max_new_tokens=a.new, use_graph=False)
torch.cuda.synchronize()
print("launches per step", ops.launch_count() // a.steps, "out", tuple(out.shape))
