"""Per-phase timeline of the fused decode chain kernel (globaltimer stamps written by the kernel itself)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from opus_pllm_b200 import _lib as L, presets, synth
from opus_pllm_b200.llama import B200Llama

B, T, new = 64, 512, 8
if os.environ.get("SK_OFF"): L.check(L.load().opus_set_tunable(b"streamk_fill", 0))
cfg = dict(presets.LLAMA3_8B); cfg["n_layers"] = 4
sd = synth.llama_weights(cfg["n_layers"], cfg["dim"], cfg["n_q_heads"], cfg["n_kv_heads"], cfg["head_dim"],
                         cfg["ffn_dim"], cfg["vocab"], seed=0, peaked=False, dtype=torch.bfloat16, device="cuda")
ll = B200Llama(sd, **cfg, device="cuda")
lib = L.load()
L.check(lib.opus_set_tunable(b"decode_fused", 1))
cu = np.arange(B + 1, dtype=np.int32) * T
emb = (torch.randn(B * T, cfg["dim"], device="cuda") * 0.02).bfloat16()
plan = ll.make_plan(cu, new)
st = ll.prefill(emb, plan=plan)
ll.generate_from_prefill(st, new)
torch.cuda.synchronize()
n_sms = lib.opus_chain_trace(1, None, 0)
ll.generate_from_prefill(st, new)      # graphs were captured without the trace pointer: force a re-capture
lib.opus_release_graphs()
ll.generate_from_prefill(st, new)
buf = (C.c_ulonglong * (n_sms * 6 * 4))()
lib.opus_chain_trace(1, buf, len(buf))
t = np.frombuffer(buf, dtype=np.uint64).reshape(n_sms, 6, 4).astype(np.int64)
t0 = t[:, 0, 0].min()
names = ["o_proj", "norm", "gate_up", "down", "norm", "lm_head/qkv"]
print("last chain launch of the step (phase 5 = lm_head); times in us relative to the first barrier pass")
for ph in range(6):
    s = t[:, ph, :]
    def rng(slot):
        v = s[:, slot]; v = v[v > 0]
        return "      -      " if v.size == 0 else f"{(v.min() - t0) / 1e3:6.1f}-{(v.max() - t0) / 1e3:6.1f}"
    print(f"phase {ph} {names[ph]:12s} barrier passed {rng(0)}  first acc {rng(1)}  norm start {rng(3)}  phase end {rng(2)}")
end2 = (t[:, 2, 2] - t0) / 1e3
acc2 = (t[:, 2, 1] - t0) / 1e3
print("gate_up phase end per CTA (us):")
for i in range(0, n_sms, 16):
    print("  cta %3d.. " % i + " ".join(f"{v:5.1f}" for v in end2[i:i + 16]))
last2 = (t[:, 2, 3] - t0) / 1e3
print("gate_up: last item acc-ready -> phase end per CTA (first 16):", [(round(a, 1), round(b, 1)) for a, b in zip(last2[:16], end2[:16])])
print("norm phase 1: start/end of CTAs 0..7:", [(round((t[i,1,3]-t0)/1e3,1), round((t[i,1,2]-t0)/1e3,1)) for i in range(8)], " CTA 100:", round((t[100,1,3]-t0)/1e3,1), round((t[100,1,2]-t0)/1e3,1))
print("o_proj: first acc / end of CTAs 0..7:", [(round((t[i,0,1]-t0)/1e3,1), round((t[i,0,2]-t0)/1e3,1)) for i in range(8)])
