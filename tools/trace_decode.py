"""In-situ timeline of one decode step / one prefill (CUDA events between launches, no graph, warm caches, real clocks)."""
import collections, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from opus_pllm_b200 import _lib as L, ops, presets, synth

model = presets.build_synthetic_model("full", "cuda", with_lora=True)
ll = model.llama
B, T, new = 64, 512, 6
cu = np.arange(B + 1, dtype=np.int32) * T
emb = (torch.randn(B * T, 4096, device="cuda") * 0.02).bfloat16()
lib = L.load()
stream = torch.cuda.current_stream().cuda_stream
plan = ll.make_plan(cu, new)
def show(title, txt, top=12):
    agg = collections.OrderedDict()
    for line in txt.strip().splitlines():
        k, v = line.split("\t"); agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += float(v)
    tot = sum(v for _, v in agg.values())
    print(f"== {title}: total {tot/1e3:.3f} ms")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"   {k:28s} n={n:4d} total={v/1e3:8.3f} ms avg={v/n:8.1f} us")
buf = C.create_string_buffer(1 << 20)
for rep in range(2):
    st = ll.prefill(emb, plan=plan)
    torch.cuda.synchronize()
lib.opus_trace_begin(stream)
st = ll.prefill(emb, plan=plan)
n = lib.opus_trace_end(buf, len(buf)); show("prefill (64 x 512)", buf.raw[:n].decode())
s, bufs = ll._decode_state(st, new, (), 0)
lib.opus_llama_select(C.byref(ll._model), C.byref(ll._ws), C.byref(s), B, stream)
for rep in range(2):
    lib.opus_llama_decode_step(C.byref(ll._model), C.byref(ll._cache), C.byref(ll._ws), C.byref(s), B, stream)
torch.cuda.synchronize()
lib.opus_trace_begin(stream)
lib.opus_llama_decode_step(C.byref(ll._model), C.byref(ll._cache), C.byref(ll._ws), C.byref(s), B, stream)
n = lib.opus_trace_end(buf, len(buf)); txt = buf.raw[:n].decode(); show("decode step (B=64, ctx 515), eager launches + events", txt)
print("first layer:", [tuple(l.split("\t")) for l in txt.strip().splitlines()[3:11]])
