"""CPU tests (no GPU) of the host logic, the C ABI surface and the data-parallel plumbing (gloo, world_size 2)."""
import os
import re
import socket

import numpy as np
import pytest
import torch

from opus_pllm_b200 import synth
from oracle import esm2_ref, mm_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------------ C ABI
def test_shared_library_exports_every_declared_symbol():
    from opus_pllm_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "opus_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*|long long)\s+(opus_\w+)\s*\(", header, flags=re.M))
    assert len(declared) >= 25
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.opus_abi_version() == 4 == _lib.ABI_VERSION
    assert isinstance(lib.opus_last_error(), bytes)


def test_contexts_and_thread_local_errors():
    """re-entrant per context. Tunables belong to the calling thread's current context (unknown names fail
    without touching it), and the error message of a failing call is visible to the thread that made it only."""
    import ctypes as C
    import threading
    from opus_pllm_b200 import _lib
    lib = _lib.load()
    ctx = C.c_void_p()
    assert lib.opus_ctx_create(C.byref(ctx)) == 0 and ctx.value
    assert lib.opus_ctx_set_current(ctx) == 0
    assert lib.opus_set_tunable(b"decode_fused", 1) == 0          # this context only
    assert lib.opus_set_tunable(b"no_such_knob", 1) < 0
    assert b"unknown name" in lib.opus_last_error()
    seen = {}

    def other():
        seen["msg"] = lib.opus_last_error()                       # another thread: its own (empty) message
        seen["rc"] = lib.opus_set_tunable(None, 0)
        seen["msg2"] = lib.opus_last_error()
    t = threading.Thread(target=other); t.start(); t.join()
    assert seen["msg"] == b"" and seen["rc"] < 0 and b"null name" in seen["msg2"]
    assert b"unknown name" in lib.opus_last_error()              # untouched by the other thread's failure
    assert lib.opus_ctx_set_current(None) == 0 and lib.opus_ctx_destroy(ctx) == 0


def test_no_cpu_fallback_ops_fail_loudly_without_cuda():
    from opus_pllm_b200 import _lib, ops
    x = torch.zeros(4, 64, dtype=torch.bfloat16)
    with pytest.raises(_lib.OpusError):
        ops.gemm(x, x)
    with pytest.raises(_lib.OpusError):
        ops.layernorm(torch.zeros(4, 64), torch.ones(64), torch.zeros(64))
    if not torch.cuda.is_available():
        with pytest.raises(_lib.OpusError):
            ops.device_check()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "opus_pllm_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


# ------------------------------------------------------------------------------------------------ tokeniser
def test_packed_tokens_match_batch_converter_semantics():
    from opus_pllm_b200.encoder import PackedTokens
    seqs = ["MKTAYIAKQR", "A", "", "XBUZO.-", "acd"]
    pk = PackedTokens(seqs)
    want = esm2_ref.tokenize(seqs)
    lens = [len(s) + 2 for s in seqs]
    assert pk.cu.tolist() == [0] + np.cumsum(lens).tolist() and pk.max_len == max(lens) and pk.n_residues == 21
    for i, n in enumerate(lens):
        assert pk.tokens[pk.cu[i]: pk.cu[i + 1]].tolist() == want[i, :n].tolist()
        assert pk.pos[pk.cu[i]: pk.cu[i + 1]].tolist() == list(range(n))
    assert np.allclose(pk.scale, 0.88)


# ------------------------------------------------------------------------------------------------ splice plan
@pytest.mark.parametrize("seed", range(5))
def test_splice_plan_matches_oracle(seed):
    from opus_pllm_b200.model import SplicePlan
    rng = np.random.default_rng(seed)
    B, L, H, n_soft = 6, 14, 8, 8
    ids = rng.integers(2, 50, size=(B, L)).astype(np.int64)
    mask = np.ones((B, L), dtype=bool)
    for b in range(B):
        pad = rng.integers(0, 6)
        mask[b, :pad] = False
        for _ in range(rng.integers(0, 3)):
            ids[b, rng.integers(pad, L)] = -200
    n_slots = int(sum(max(1, int((ids[b][mask[b]] == -200).sum())) for b in range(B)))
    embed = torch.randn(50, H)
    soft = torch.randn(n_slots, n_soft, H)
    plan = SplicePlan(ids, mask, n_soft, n_slots)
    table = torch.cat([embed, soft.reshape(-1, H)])
    idx = torch.from_numpy(np.where(plan.src >= 0, plan.src, 50 + (-plan.src - 1)))
    packed = table[idx]
    for left in (True, False):
        e, m, p, lens = mm_ref.splice(torch.from_numpy(ids), torch.from_numpy(mask), soft, embed, left)
        assert lens == plan.lens.tolist()
        src, pm = plan.padded_src(left)
        assert np.array_equal(pm, m.numpy())
        for b in range(B):
            assert torch.equal(e[b][m[b]], packed[plan.cu[b]: plan.cu[b + 1]])
    assert plan.n_seq_used == n_slots
    with pytest.raises(IndexError):
        SplicePlan(np.array([[-200, -200]]), None, n_soft, 1)


def test_splice_plan_truncation():
    from opus_pllm_b200.model import SplicePlan
    plan = SplicePlan(np.array([[5, -200, 6, 7]]), None, 8, 1, max_length=6)
    assert plan.lens.tolist() == [6] and plan.src.tolist() == [5, -1, -2, -3, -4, -5]


# ------------------------------------------------------------------------------------------------ small host helpers
def test_block_allocator_and_trim():
    from opus_pllm_b200.llama import BlockAllocator, _trim_like_hf
    from opus_pllm_b200._lib import OpusError
    a = BlockAllocator(8)
    x = a.alloc(3)
    y = a.alloc(5)
    assert sorted(x + y) == list(range(8))
    with pytest.raises(OpusError):
        a.alloc(1)
    a.release(x)
    assert sorted(a.alloc(3)) == sorted(x)
    out = torch.tensor([[4, 9, 9, 9], [4, 5, 9, 9]])
    assert _trim_like_hf(out, [9]).shape == (2, 3)
    assert _trim_like_hf(torch.tensor([[4, 9, 9], [4, 5, 6]]), [9]).shape == (2, 3)


def test_synth_is_deterministic_known_answers():
    a = synth.hash_uniform((4,), "x", 3)
    b = synth.hash_uniform((2, 2), "x", 3).flatten()
    assert torch.equal(a, b)
    assert [round(float(v), 6) for v in synth.hash_uniform((1000, 1000), "x", 3)[0, :3]] == [-0.351248, -0.427369,
                                                                                            -0.160979]
    assert synth.proteins(2, 5)[0] == synth.proteins(1, 5)[0] and set(synth.proteins(1, 50)[0]) <= set(synth.AMINO)
    p = synth.prompt_ids(3, 20, vocab=1000)
    assert all(int((x == -200).sum()) == 1 for x in p)


# ------------------------------------------------------------------------------------------------ data parallel
def test_shard_bounds_match_accelerate_semantics():
    from opus_pllm_b200.dp import shard_bounds, split_between_processes
    items = list(range(10))
    got = [split_between_processes(items, r, 4) for r in range(4)]
    assert got == [[0, 1, 2], [3, 4, 5], [6, 7], [8, 9]]
    assert [shard_bounds(3, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 3), (3, 3)]


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    from opus_pllm_b200.dp import gather_token_ids, split_between_processes
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    prompts = list(range(7))
    mine = split_between_processes(prompts)
    ids = torch.tensor([[p * 10 + t for t in range(3 + rank)] for p in mine], dtype=torch.int64)
    out = gather_token_ids(ids, pad_id=-1)
    q.put((rank, out.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_gather_world_size_2_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [[p * 10 + t for t in range(3)] + [-1] for p in range(4)] + [[p * 10 + t for t in range(4)] for p in range(4, 7)]
    assert res[0] == want and res[1] == want


def test_streamk_tail_partition_covers_every_unit_once():
    """Python restatement of gemm_tcgen05.cu::get_work's stream-K tail partition: every k-block unit of the tail tiles is
    computed exactly once, a CTA touches at most two tiles, and the owner's [first_cta, owner) range is exactly the set
    of CTAs that dump a partial for that tile (so the counter it waits on reaches the expected value)."""
    def sim(sk_tiles, kb, g, share=0):
        U = sk_tiles * kb
        ge = min(g, U)
        if 0 < share < ge:          # GemmParams::sk_share: a tail of a few tiles is cut into 8 pieces per tile only
            ge = share
        cover, contrib, owners = [0] * U, {}, {}
        for b in range(ge):
            u0, u1 = b * U // ge, (b + 1) * U // ge
            assert 0 < u1 - u0 <= kb
            tA, tB = u0 // kb, (u1 - 1) // kb
            segs = [(tA, u0 - tA * kb, u1 - tA * kb)] if tA == tB else [(tB, 0, u1 - tB * kb), (tA, u0 - tA * kb, kb)]
            assert tB - tA <= 1
            for t, k0, k1 in segs:
                for k in range(k0, k1):
                    cover[t * kb + k] += 1
                if k1 == kb:
                    owners[t] = (b, ((t * kb + 1) * ge - 1) // U)
                else:
                    contrib.setdefault(t, []).append(b)
        assert all(c == 1 for c in cover)
        for t in range(sk_tiles):
            b, first = owners[t]
            assert contrib.get(t, []) == list(range(first, b))
    for g in (148, 132):
        for sk in (1, 2, 3, 53, 76, 114, g - 1):
            for kb in (1, 2, 5, 16, 20, 64, 80, 224):
                sim(sk, kb, g)
    for g, sk in ((148, 4), (148, 14), (74, 2), (74, 7), (74, 1)):      # small tails: rem * 10 <= g, share = 8 per tile
        for kb in (5, 16, 64, 224):
            sim(sk, kb, g, share=min(g, sk * 8))


def test_keywords_stopping_criteria_host_half():
    """mm_utils.py:43-61: keywords -> id sequences, a leading BOS dropped only when something follows it."""
    from opus_pllm_b200.mm_utils import KeywordsStoppingCriteria

    class Tok:
        bos_token_id = 1

        def __call__(self, text):
            table = {"###": [1, 835], "</s>": [1], "Human:": [1, 12968, 29901]}
            return type("E", (), {"input_ids": table[text]})()

    crit = KeywordsStoppingCriteria(["###", "</s>", "Human:"], Tok(), torch.zeros(2, 7, dtype=torch.long))
    assert [k.tolist() for k in crit.keyword_ids] == [[835], [1], [12968, 29901]]
    assert crit.max_keyword_len == 2 and crit.start_len == 7


def test_compat_shim_loads_the_unmodified_reference_scripts():
    """opus_pllm_b200.compat registers the backend as `multi_modality_model.multi_modality_v1.model.builder`; the three
    reference eval scripts then import (no GPU needed for that) and bind this backend's loader. Skipped when the
    reference tree has not been staged under baseline/_ref (oracle/stage_reference.py)."""
    import os
    import sys
    import pytest
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = os.path.join(root, "baseline", "_ref")
    ev = os.path.join(ref, "multi_modality_model", "multi_modality_v1", "eval")
    if not os.path.isfile(os.path.join(ev, "run_opus_ddp.py")):
        pytest.skip("baseline/_ref not staged")
    saved = dict(sys.modules)
    sys.path[:0] = [ref, ev]
    try:
        from opus_pllm_b200 import builder, compat
        compat.install(force_stubs=True)
        for name in ("run_opus_ddp.py", "eval_run_multichoice.py", "run_opus_online.py"):
            mod = compat.load_script(os.path.join(ev, name), "ref_" + name[:-3])
            assert mod.load_pretrained_model is builder.load_pretrained_model
            assert mod.tokenizer_seq_token.__module__ == "multi_modality_model.multi_modality_v1.mm_utils"
        import accelerate
        acc = accelerate.Accelerator()
        with acc.split_between_processes(list(range(5))) as part:
            assert part == list(range(5)) and acc.is_main_process
        assert accelerate.utils.gather_object([[1, 2]]) == [[1, 2]]
    finally:
        sys.path.remove(ref); sys.path.remove(ev)
        for k in set(sys.modules) - set(saved):
            if k.startswith(("multi_modality_model", "accelerate", "metrics_computing_opi", "ref_")):
                del sys.modules[k]


def test_eval_ddp_max_new_tokens_schedule_follows_the_reference_loop():
    """run_opus_ddp.py:90-101: the override is decided per instruction (only when `<seq>` is absent), always forces
    32 / 128 / 256, and sticks for the rest of the run because the script mutates args.max_new_tokens."""
    from opus_pllm_b200 import eval_ddp
    ins = ["<seq>\nq0", "<seq>\nq1", "q2", "<seq>\nq3"]
    assert eval_ddp.max_new_tokens_schedule(ins, "d/keywords.json", 50) == [50, 50, 128, 128]
    assert eval_ddp.max_new_tokens_schedule(ins, "d/localization.json", 50) == [50, 50, 32, 32]
    assert eval_ddp.max_new_tokens_schedule(ins, "d/function.json", 32) == [32, 32, 256, 256]
    assert eval_ddp.max_new_tokens_schedule(ins[:2], "d/function.json", 77) == [77, 77]
    assert eval_ddp.max_new_tokens_schedule(ins, "d/function.json", 9, fixed=True) == [9] * 4


def test_eos_ids_merge_generation_config(tmp_path):
    import json
    from opus_pllm_b200 import builder
    assert builder._eos_ids(str(tmp_path), 2) == [2] and builder._eos_ids(str(tmp_path), None) == []
    json.dump({"eos_token_id": [128001, 128009]}, open(tmp_path / "generation_config.json", "w"))
    assert builder._eos_ids(str(tmp_path), 128001) == [128001, 128009]
    assert builder._eos_ids(str(tmp_path), [5]) == [128001, 128009, 5]


def test_pad_heads_layout_preserves_rotary_pairs_and_dot_products():
    """llama.pad_heads: narrow heads stored in 128 columns. With rotary=True the halves land at columns [0, h) and
    [64, 64 + h), so the kernels' rotate_half pairing (j, j + 64) acts on the original pairs (j, j + h); dot products over
    the padded axis equal the originals; projecting back through the padded o_proj columns is the identity."""
    from opus_pllm_b200.llama import pad_heads
    g = torch.Generator().manual_seed(0)
    H, hr, D = 3, 64, 40
    W = torch.randn(H * hr, D, generator=g)
    Wp = pad_heads(W, H, hr, 0, rotary=True)
    assert Wp.shape == (H * 128, D)
    v, vp = W.view(H, hr, D), Wp.view(H, 128, D)
    assert torch.equal(vp[:, :32], v[:, :32]) and torch.equal(vp[:, 64:96], v[:, 32:])
    assert float(vp[:, 32:64].abs().max()) == 0 and float(vp[:, 96:].abs().max()) == 0
    x = torch.randn(5, D, generator=g)
    q, qp = (x @ W.T).view(5, H, hr), (x @ Wp.T).view(5, H, 128)
    assert torch.allclose((q * q.flip(0)).sum(-1), (qp * qp.flip(0)).sum(-1), atol=1e-4)
    # rotate_half on the padded layout == padded rotate_half of the original
    rot = torch.cat([-q[..., 32:], q[..., :32]], -1)
    rot_p = torch.cat([-qp[..., 64:], qp[..., :64]], -1)
    assert torch.allclose(pad_heads(rot.reshape(5, H * hr), H, hr, 1, rotary=True).view(5, H, 128), rot_p)
    # o_proj columns: out = attn @ Wo^T is unchanged when both are padded the same way
    Wo = torch.randn(D, H * hr, generator=g)
    a = torch.randn(5, H * hr, generator=g)
    for rotary in (True, False):
        ap, Wop = pad_heads(a, H, hr, 1, rotary=rotary), pad_heads(Wo, H, hr, 1, rotary=rotary)
        assert torch.allclose(a @ Wo.T, ap @ Wop.T, atol=1e-4)
    assert pad_heads(W, H * hr // 128, 128, 0, rotary=True) is W
    W80 = torch.randn(2 * 80, D, generator=g)
    assert pad_heads(W80, 2, 80, 0, rotary=False).view(2, 128, D)[:, 80:].abs().max() == 0


def test_bench_reference_arm_runs_on_cpu_for_every_workload_kind():
    """`bench.py --impl reference` (the driver's second arm) never touches the GPU: tiny size, one generate-type and one
    encoder-type workload, one JSON line each with the contract's keys."""
    import json
    import subprocess
    import sys
    for wl, unit in (("c2", "tokens/s"), ("c1", "residues/s"), ("c5", "tokens/s")):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "tiny",
                            "--workload", wl, "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300,
                           env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
        assert r.returncode == 0, r.stderr[-2000:]
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        assert len(lines) == 1
        d = json.loads(lines[0])
        assert d["impl"] == "reference" and d["unit"] == unit and d["value"] > 0 and d["gpu_launches"] == 0
        assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
        assert d["e2e"] == {"value": d["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        assert d["config"]["workload"].startswith(wl + ":")


def test_scheduler_cut_round_matches_the_token_by_token_rule():
    """scheduler.cut_round (all slots of a round at once) against the scalar rule of the reference loop: keep tokens up to
    and including the first EOS, never more than what is left of max_new_tokens; complete on EOS or on the limit."""
    import numpy as np
    from opus_pllm_b200.scheduler import cut_round
    rng = np.random.default_rng(5)
    for eos in ((), (7,), (7, 3)):
        toks = rng.integers(0, 12, size=(200, 16)).astype(np.int32)
        room = rng.integers(1, 40, size=200)
        take, done = cut_round(toks, room, eos)
        for s in range(200):
            kept, fin = [], False
            for t in toks[s]:
                kept.append(int(t))
                if int(t) in eos or len(kept) >= room[s]:
                    fin = True
                    break
            assert take[s] == len(kept) and bool(done[s]) == fin, (s, eos)
    take, done = cut_round(np.zeros((0, 16), dtype=np.int32), np.zeros(0, dtype=np.int64), (1,))
    assert take.shape == (0,) and done.shape == (0,)
