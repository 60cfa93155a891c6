"""The reference's *GPU* path on this box ("the practical bar"): stock PyTorch / transformers kernels, none of
this repo's. ESM-2-650M fp32 weights under fp16 autocast (opus_arch.py:107) via HF EsmModel (fair-esm is absent),
projectors under autocast, then HF LlamaForCausalLM (bf16, sdpa) `generate(inputs_embeds=..., do_sample=False)` the way
opus_llama.py:126-132 calls it. Random-init weights, C2 shapes: 64 prompts x 512 tokens, 32 new tokens, as 8 batches of
8 (run_opus_ddp.py:75) and as one batch of 64. Prints one JSON line per batch size. Not a bench.py arm: reported beside it."""
import json, sys, time
import torch
from transformers import EsmConfig, EsmModel, LlamaConfig, LlamaForCausalLM

import os
TINY = os.environ.get("TINY") == "1"                     # CPU syntax check of the script only
NP, T, NEW, PLEN = (4, 64, 4, 16) if TINY else (64, 512, 32, 256)
dev = "cpu" if TINY else "cuda"
L_E, L_L = (1, 1) if TINY else (33, 32)
torch.manual_seed(0)


def build():
    ecfg = EsmConfig(vocab_size=33, hidden_size=1280, num_hidden_layers=L_E, num_attention_heads=20,
                     intermediate_size=5120, position_embedding_type="rotary", token_dropout=True,
                     emb_layer_norm_before=False, mask_token_id=32, pad_token_id=1, layer_norm_eps=1e-5,
                     hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    with torch.device(dev):
        esm = EsmModel(ecfg, add_pooling_layer=False).eval()
        proj = torch.nn.Linear(1280, 5120)
        switch = torch.nn.Sequential(torch.nn.Linear(5120, 32768), torch.nn.GELU(), torch.nn.Linear(32768, 32768))
        lcfg = LlamaConfig(vocab_size=128256, hidden_size=4096, intermediate_size=14336, num_hidden_layers=L_L,
                           num_attention_heads=32, num_key_value_heads=8, rope_theta=500000.0, rms_norm_eps=1e-5,
                           max_position_embeddings=8192, bos_token_id=128000, eos_token_id=128001,
                           attn_implementation="sdpa")
        prev = torch.get_default_dtype()
        torch.set_default_dtype(torch.bfloat16)
        try:
            llama = LlamaForCausalLM(lcfg).eval()
        finally:
            torch.set_default_dtype(prev)
    return esm, proj, switch.to(torch.bfloat16), llama


@torch.inference_mode()
def step(esm, proj, switch, llama, bs):
    """One pass over the 64 prompts in batches of `bs`; returns generated token count."""
    n = 0
    for s in range(0, NP, bs):
        tok = torch.randint(4, 24, (bs, PLEN + 2), device=dev)
        tok[:, 0], tok[:, -1] = 0, 2
        with torch.autocast(dev, dtype=torch.bfloat16 if TINY else torch.float16):
            h = esm(input_ids=tok, attention_mask=torch.ones_like(tok)).last_hidden_state
            pooled = h[:, 1:-1].float().mean(1)
            c = proj(torch.nn.functional.normalize(pooled, dim=-1))
        soft = switch(c.to(torch.bfloat16)).view(bs, 8, 4096)
        ids = torch.randint(1000, 120000, (bs, T - 8), device=dev)
        emb = llama.get_input_embeddings()(ids)
        emb = torch.cat([emb[:, :40], soft, emb[:, 40:]], 1)
        mask = torch.ones(bs, T, dtype=torch.bool, device=dev)
        out = llama.generate(inputs_embeds=emb, attention_mask=mask, do_sample=False, max_new_tokens=NEW,
                             min_new_tokens=NEW, pad_token_id=128001, use_cache=True)
        n += out.numel()
    return n


def main():
    esm, proj, switch, llama = build()
    for bs in [int(a) for a in sys.argv[1:]] or [8, 64]:
        for _ in range(2):
            step(esm, proj, switch, llama, bs)
        reps, n = 3, 0
        if TINY:
            t = time.perf_counter()
            n = sum(step(esm, proj, switch, llama, bs) for _ in range(reps))
            ms = 1e3 * (time.perf_counter() - t) / reps
        else:
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                n += step(esm, proj, switch, llama, bs)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
        print(json.dumps({"impl": "reference-gpu (stock transformers " + __import__("transformers").__version__ +
                          ", bf16 sdpa, HF EsmModel fp16 autocast)", "batch": bs, "prompts": NP, "prompt_len": T,
                          "new_tokens": NEW, "ms_per_step": ms, "tokens_per_s": n / reps / (ms / 1e3)}), flush=True)


if __name__ == "__main__":
    t0 = time.time()
    main()
    print(f"# total {time.time() - t0:.0f} s", file=sys.stderr)
