"""GPU parity tests of every kernel behind the C ABI against the oracle (oracle/ops_ref.py), same seeded inputs."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ops_ref as R  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    from opus_pllm_b200 import ops as o
    o.device_check()
    return o


def _randn(shape, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).cuda()


def _close_bf16(got: torch.Tensor, want_f32: torch.Tensor, ulps: float = 1.0, atol: float = 1e-3):
    """bf16 result vs fp32 oracle: within `ulps` bf16 ulps (2^-8 relative) + atol (accumulation-order noise)."""
    got = got.float()
    err = (got - want_f32).abs()
    tol = atol + ulps * (2.0 ** -8) * want_f32.abs()
    bad = err > tol
    assert not bad.any(), (
        f"{int(bad.sum())}/{bad.numel()} elements off; max err {float(err.max()):.4g} at "
        f"{tuple(int(i) for i in torch.nonzero(bad)[0])}; got {float(got[bad][0]):.6g} want {float(want_f32[bad][0]):.6g}")


# ------------------------------------------------------------------------------------------------ GEMM
GEMM_SHAPES = [
    # rows, N, K
    (128, 256, 64),
    (256, 512, 128),
    (300, 1280, 1280),     # M tail, encoder out-proj shape
    (1000, 3840, 1280),    # encoder qkv
    (516, 5120, 1280),
    (384, 1280, 5120),
    (520, 6144, 4096),     # llama qkv (prefill)
    (130, 128, 4096),      # small N tile
]


@pytest.mark.parametrize("rows,N,K", GEMM_SHAPES)
def test_gemm_plain_and_bias(ops, rows, N, K):
    from opus_pllm_b200._lib import EPI_BF16
    x = _randn((rows, K), 1)
    w = _randn((N, K), 2, scale=K ** -0.5)
    bias = _randn((N,), 3, dtype=torch.float32)
    want = R.linear_ref(x, w, bias)
    got = ops.gemm(x, w, epilogue=EPI_BF16, bias=bias, transposed=False)
    _close_bf16(got, want)
    got_nb = ops.gemm(x, w, epilogue=EPI_BF16, transposed=False)
    _close_bf16(got_nb, R.linear_ref(x, w))


@pytest.mark.parametrize("bn", [64, 128, 256])
def test_gemm_block_n_variants(ops, bn):
    from opus_pllm_b200._lib import EPI_BF16
    x = _randn((400, 1024), 4)
    w = _randn((768, 1024), 5, scale=1 / 32)
    got = ops.gemm(x, w, epilogue=EPI_BF16, transposed=False, block_n=bn)
    _close_bf16(got, R.linear_ref(x, w))


def test_gemm_gelu(ops):
    from opus_pllm_b200._lib import EPI_BF16_GELU
    x = _randn((260, 1280), 6)
    w = _randn((5120, 1280), 7, scale=1280 ** -0.5)
    b = _randn((5120,), 8, dtype=torch.float32)
    got = ops.gemm(x, w, epilogue=EPI_BF16_GELU, bias=b, transposed=False)
    _close_bf16(got, R.gelu_erf(R.linear_ref(x, w, b)))


def test_gemm_residual_f32_inplace(ops):
    from opus_pllm_b200._lib import EPI_RES_F32
    x = _randn((300, 5120), 9)
    w = _randn((1280, 5120), 10, scale=5120 ** -0.5)
    b = _randn((1280,), 11, dtype=torch.float32)
    res = _randn((300, 1280), 12, dtype=torch.float32)
    want = res + R.linear_ref(x, w, b)
    out = res.clone()
    ops.gemm(x, w, epilogue=EPI_RES_F32, bias=b, residual=out, out=out, transposed=False)
    assert torch.allclose(out, want, rtol=1e-4, atol=2e-3), float((out - want).abs().max())


def test_gemm_residual_bf16(ops):
    from opus_pllm_b200._lib import EPI_RES_BF16
    x = _randn((257, 4096), 13)
    w = _randn((4096, 4096), 14, scale=1 / 64)
    res = _randn((257, 4096), 15)
    want = R.bf16r(res.float() + R.bf16r(R.linear_ref(x, w)))
    got = ops.gemm(x, w, epilogue=EPI_RES_BF16, residual=res, transposed=False)
    _close_bf16(got, want, ulps=2.0, atol=2e-2)


def test_gemm_swiglu(ops):
    from opus_pllm_b200._lib import EPI_SWIGLU
    x = _randn((300, 1024), 16)
    w = _randn((2 * 1536, 1024), 17, scale=1 / 32)
    want = R.swiglu_interleaved_ref(R.linear_ref(x, w))
    got = ops.gemm(x, w, epilogue=EPI_SWIGLU, transposed=False)
    assert got.shape == (300, 1536)
    _close_bf16(got, want, ulps=3.0, atol=2e-2)


@pytest.mark.parametrize("rows", [1, 8, 33, 64, 100, 256])
def test_gemm_transposed_weight_streaming(ops, rows):
    from opus_pllm_b200._lib import EPI_BF16, EPI_BF16_GELU
    x = _randn((rows, 2048), 18)
    w = _randn((1000 + 24, 2048), 19, scale=2048 ** -0.5)  # 1024 features
    b = _randn((1024,), 20, dtype=torch.float32)
    got = ops.gemm(x, w, epilogue=EPI_BF16, bias=b, transposed=True)
    _close_bf16(got, R.linear_ref(x, w, b))
    got = ops.gemm(x, w, epilogue=EPI_BF16_GELU, bias=b, transposed=True)
    _close_bf16(got, R.gelu_erf(R.linear_ref(x, w, b)))


def test_gemm_transposed_swiglu_and_residual(ops):
    from opus_pllm_b200._lib import EPI_SWIGLU, EPI_RES_BF16
    x = _randn((64, 1024), 21)
    w = _randn((2 * 1024, 1024), 22, scale=1 / 32)
    got = ops.gemm(x, w, epilogue=EPI_SWIGLU, transposed=True)
    _close_bf16(got, R.swiglu_interleaved_ref(R.linear_ref(x, w)), ulps=3.0, atol=2e-2)
    res = _randn((64, 2048), 23)
    got = ops.gemm(x, w, epilogue=EPI_RES_BF16, residual=res, transposed=True)
    _close_bf16(got, R.bf16r(res.float() + R.bf16r(R.linear_ref(x, w))), ulps=2.0, atol=2e-2)


@pytest.mark.parametrize("transposed,rows", [(True, 64), (False, 300)])
def test_gemm_split_k_partials(ops, transposed, rows):
    from opus_pllm_b200._lib import EPI_PARTIAL_F32
    x = _randn((rows, 4096), 24)
    w = _randn((512, 4096), 25, scale=1 / 64)
    part = ops.gemm(x, w, epilogue=EPI_PARTIAL_F32, transposed=transposed, split_k=4)
    assert part.shape == (4, rows, 512)
    want = R.linear_ref(x, w)
    assert torch.allclose(part.sum(0), want, rtol=1e-4, atol=2e-3), float((part.sum(0) - want).abs().max())
    red = ops.splitk_reduce(part)
    _close_bf16(red, want)


# Stream-K tail (gemm_tcgen05.cu: get_work): shapes whose tile count is not a multiple of the SM count, so the last
# partial wave is cut along K across all CTAs and fixed up by the tile owners. Tail sizes: 76, 2, 140 (off: > 90 %),
# 114 tiles (transposed); 645 / 1935 tiles (plain form, the encoder's out_proj / qkv shapes at C1).
@pytest.mark.parametrize("tiles,K,rows", [(224, 4096, 64), (150, 1024, 64), (288, 512, 16), (262, 2048, 256),
                                          (299, 320, 33)])
def test_gemm_streamk_tail_transposed(ops, tiles, K, rows):
    from opus_pllm_b200._lib import EPI_BF16, EPI_F32, EPI_SWIGLU
    x = _randn((rows, K), 40)
    w = _randn((tiles * 128, K), 41, scale=K ** -0.5)
    b = _randn((tiles * 128,), 42, dtype=torch.float32)
    want = R.linear_ref(x, w, b)
    for rep in range(3):   # the tail-tile counters must be re-armed by every launch
        got32 = ops.gemm(x, w, epilogue=EPI_F32, bias=b, transposed=True)
        assert torch.allclose(got32, want, rtol=1e-4, atol=2e-3), float((got32 - want).abs().max())
    _close_bf16(ops.gemm(x, w, epilogue=EPI_BF16, bias=b, transposed=True), want)
    got = ops.gemm(x, w, epilogue=EPI_SWIGLU, transposed=True)
    _close_bf16(got, R.swiglu_interleaved_ref(R.linear_ref(x, w)), ulps=3.0, atol=2e-2)
    # deterministic: partials are added in CTA order
    assert torch.equal(ops.gemm(x, w, epilogue=EPI_F32, bias=b, transposed=True), got32)


@pytest.fixture
def streamk_plain():
    """The plain-form stream-K tail is opt-in (it trades bitwise batch invariance for wave balance)."""
    from opus_pllm_b200 import _lib as L
    L.check(L.load().opus_set_tunable(b"streamk_plain", 1))
    yield
    L.check(L.load().opus_set_tunable(b"streamk_plain", 0))


@pytest.mark.parametrize("rows,N,K", [(16512, 1280, 1280), (16512, 3840, 1280), (5000, 1280, 5120), (2100, 2560, 192)])
def test_gemm_streamk_tail_plain(ops, rows, N, K, streamk_plain):
    from opus_pllm_b200._lib import EPI_BF16, EPI_BF16_GELU, EPI_RES_BF16
    x = _randn((rows, K), 43)
    w = _randn((N, K), 44, scale=K ** -0.5)
    b = _randn((N,), 45, dtype=torch.float32)
    want = R.linear_ref(x, w, b)
    for rep in range(2):
        _close_bf16(ops.gemm(x, w, epilogue=EPI_BF16, bias=b, transposed=False), want)
    _close_bf16(ops.gemm(x, w, epilogue=EPI_BF16_GELU, bias=b, transposed=False), R.gelu_erf(want))
    res = _randn((rows, N), 46)
    h = res.clone()
    ops.gemm(x, w, epilogue=EPI_RES_BF16, residual=h, out=h, transposed=False)   # in place, like o_proj / down_proj
    _close_bf16(h, R.bf16r(res.float() + R.bf16r(R.linear_ref(x, w))), ulps=2.0, atol=2e-2)


@pytest.mark.parametrize("rows,N,K", [(512, 28672, 4096), (384, 28672, 4096), (300, 128256, 1024)])
def test_gemm_streamk_small_tail_plain_decode_shapes(ops, rows, N, K, streamk_plain):
    """Decode at batch 257..512 runs gate/up and lm_head in the plain form with the stream-K tail on; a tail of a few tiles
    (448 = 3 waves + 4) is shared by 8 CTAs per tile. Against whole tiles: same values up to the fp32 summation order of
    the tail tiles, deterministic."""
    from opus_pllm_b200 import _lib as L
    x = _randn((rows, K), 47)
    w = _randn((N, K), 48, scale=K ** -0.5)
    got = ops.gemm(x, w, epilogue=L.EPI_SWIGLU, transposed=False)
    assert torch.equal(got, ops.gemm(x, w, epilogue=L.EPI_SWIGLU, transposed=False))
    got16 = ops.gemm(x, w, epilogue=L.EPI_BF16, transposed=False)
    L.check(L.load().opus_set_tunable(b"streamk_plain", 0))
    want = ops.gemm(x, w, epilogue=L.EPI_SWIGLU, transposed=False)
    want16 = ops.gemm(x, w, epilogue=L.EPI_BF16, transposed=False)
    for a, b in ((got, want), (got16, want16)):
        assert float((a != b).float().mean()) <= 0.35          # only tail tiles may differ at all
        assert float((a.float() - b.float()).abs().max()) <= 2.0 ** -6 * float(b.float().abs().max())
    _close_bf16(got16, R.linear_ref(x, w))


@pytest.mark.parametrize("M,N,K", [(1024, 256, 64), (1100, 768, 1280), (4096, 5120, 1280), (16512, 1280, 1280),
                                   (2500, 6272, 512)])
def test_gemm_cta_pair_form_is_bit_identical(ops, M, N, K):
    """cta_group::2 kernel (two CTAs share a 256 x 256 tile, UMMA M = 256, B split across the pair) against the
    single-CTA kernel: same k order per output element, so every epilogue must agree bit for bit."""
    from opus_pllm_b200 import _lib as L
    x = _randn((M, K), 50)
    w = _randn((N, K), 51, scale=K ** -0.5)
    b = _randn((N,), 52, dtype=torch.float32)
    res = _randn((M, N), 53)
    lib = L.load()
    out = {}
    try:
        for mode in (1, 0):
            L.check(lib.opus_set_tunable(b"gemm_2cta", mode))
            h = res.clone()
            ops.gemm(x, w, epilogue=L.EPI_RES_BF16, residual=h, out=h, transposed=False)
            out[mode] = (ops.gemm(x, w, epilogue=L.EPI_BF16, bias=b, transposed=False),
                         ops.gemm(x, w, epilogue=L.EPI_BF16_GELU, bias=b, transposed=False),
                         ops.gemm(x, w, epilogue=L.EPI_SWIGLU, transposed=False) if N % 16 == 0 else None, h)
    finally:
        L.check(lib.opus_set_tunable(b"gemm_2cta", 2))
    for a, c in zip(out[1], out[0]):
        if a is not None:
            assert torch.equal(a, c)
    _close_bf16(out[1][0], R.linear_ref(x, w, b))


@pytest.mark.parametrize("rows", [129, 200, 256, 300, 512])          # above 256: two batch tiles per weight tile
@pytest.mark.parametrize("N,K,split", [(4096, 4096, 4), (6144, 4096, 3), (4096, 14336, 4), (18816, 1024, 1), (8320, 512, 2)])
def test_gemm_cta_pair_swap_ab_form_is_bit_identical(ops, rows, N, K, split):
    """Swap-AB (weight-streaming) launches at batch 129..256 run on the CTA-pair kernel (each CTA holds half of the
    256-row activation tile; (256-feature tile, k-split) work items): same k order per element as the single-CTA
    kernel, so split-K partials and every fused epilogue must agree bit for bit (decode shapes of Llama-3-8B: o_proj,
    qkv, down; a 74-tile single wave; an odd 128-feature tail)."""
    from opus_pllm_b200 import _lib as L
    x = _randn((rows, K), 54)
    w = _randn((N, K), 55, scale=K ** -0.5)
    b = _randn((N,), 56, dtype=torch.float32)
    res = _randn((rows, N), 57)
    lib = L.load()
    out = {}
    try:
        for mode in (1, 0):
            L.check(lib.opus_set_tunable(b"gemm_2cta_tr", mode))
            h = res.clone()
            ops.gemm(x, w, epilogue=L.EPI_RES_BF16, bias=b, residual=h, out=h, transposed=True)
            out[mode] = (ops.gemm(x, w, epilogue=L.EPI_PARTIAL_F32, transposed=True, split_k=split),
                         ops.gemm(x, w, epilogue=L.EPI_BF16, bias=b, transposed=True),
                         ops.gemm(x, w, epilogue=L.EPI_BF16_GELU, bias=b, transposed=True), h)
    finally:
        L.check(lib.opus_set_tunable(b"gemm_2cta_tr", 2))
    for a, c in zip(out[1], out[0]):
        assert torch.equal(a, c)
    assert out[1][0].shape == (split, rows, N)
    _close_bf16(out[1][1], R.linear_ref(x, w, b))
    assert torch.allclose(out[1][0].sum(0), R.linear_ref(x, w), rtol=1e-4, atol=2e-3)


def test_gemm_linearity_full_size(ops):
    """size-independent property at a BASELINE-sized weight: f(x1 + x2) == f(x1) + f(x2) up to bf16 rounding."""
    from opus_pllm_b200._lib import EPI_F32
    w = _randn((28672, 4096), 26, scale=1 / 64)
    x1 = _randn((64, 4096), 27)
    x2 = _randn((64, 4096), 28)
    xs = (x1.float() + x2.float()).to(torch.bfloat16)
    f = lambda x: ops.gemm(x, w, epilogue=EPI_F32, transposed=True)  # noqa: E731
    lhs = f(xs)
    rhs = f(x1) + f(x2)
    # xs is rounded to bf16, so allow the corresponding perturbation
    assert torch.allclose(lhs, rhs, rtol=0, atol=0.15), float((lhs - rhs).abs().max())
    want = R.linear_ref(x1, w)
    assert torch.allclose(f(x1), want, rtol=1e-4, atol=2e-3)


# ------------------------------------------------------------------------------------------------ norms
def test_layernorm(ops):
    x = _randn((1000, 1280), 30, scale=3.0, dtype=torch.float32) + 0.5
    g = _randn((1280,), 31, dtype=torch.float32)
    b = _randn((1280,), 32, dtype=torch.float32)
    _close_bf16(ops.layernorm(x, g, b, 1e-5), R.layernorm_ref(x, g, b, 1e-5), ulps=1.0, atol=1e-4)


def test_layernorm_folds_pending_delta_in_place(ops):
    x = _randn((600, 1280), 130, scale=3.0, dtype=torch.float32)
    d = _randn((600, 1280), 131)
    g = _randn((1280,), 132, dtype=torch.float32)
    b = _randn((1280,), 133, dtype=torch.float32)
    want_x = x + d.float()
    got = ops.layernorm(x, g, b, 1e-5, delta=d)
    assert torch.equal(x, want_x)
    _close_bf16(got, R.layernorm_ref(want_x, g, b, 1e-5), ulps=1.0, atol=1e-4)
    # final LN + pool with a pending delta
    cu = torch.tensor([0, 200, 600], dtype=torch.int32).cuda()
    pooled, _, hidden = ops.final_ln_meanpool(x, cu, g, b, 1e-5, want_hidden=True, delta=d)
    h_want = R.layernorm_ref(x + d.float(), g, b)
    assert torch.allclose(hidden, h_want, rtol=1e-4, atol=1e-4)
    assert torch.allclose(pooled, R.meanpool_ref(h_want, cu.tolist()), rtol=1e-4, atol=1e-4)


def test_rmsnorm_exact(ops):
    x = _randn((777, 4096), 33, scale=2.0)
    w = _randn((4096,), 34)
    got = ops.rmsnorm(x, w, 1e-5)
    want = R.rmsnorm_ref(x, w, 1e-5)
    # same rounding points as HF -> allow only rare 1-ulp flips from reduction order
    diff = (got.float() - want).abs() > (2.0 ** -7) * want.abs() + 1e-6
    assert diff.float().mean() < 1e-3, float(diff.float().mean())


def test_rmsnorm_residual_partials(ops):
    part = _randn((3, 64, 4096), 35, dtype=torch.float32)
    res = _randn((64, 4096), 36)
    w = _randn((4096,), 37)
    h_out = torch.empty_like(res)
    y = ops.rmsnorm(None, w, 1e-5, partial=part, residual=res, h_out=h_out)
    h_want = R.bf16r(res.float() + R.bf16r(part.sum(0)))
    _close_bf16(h_out, h_want, ulps=1.0, atol=1e-6)
    y_want = R.rmsnorm_ref(h_out, w, 1e-5)
    diff = (y.float() - y_want).abs() > (2.0 ** -7) * y_want.abs() + 1e-6
    assert diff.float().mean() < 1e-3


# ------------------------------------------------------------------------------------------------ embeddings / rope
def test_esm_embed(ops):
    table = _randn((33, 1280), 40, dtype=torch.float32)
    tok = torch.randint(0, 33, (500,), dtype=torch.int32).cuda()
    scale = torch.full((500,), 0.88, dtype=torch.float32).cuda()
    scale[7] = 0.0
    got = ops.esm_embed(tok, scale, table)
    want = table[tok.long()] * scale[:, None]
    assert torch.equal(got, want)


def test_rope_esm(ops):
    T, H, D = 300, 20, 64
    qkv = _randn((T, 3 * H * D), 41)
    pos = torch.arange(T, dtype=torch.int32).cuda()
    pos[100:] -= 100  # two packed sequences
    cos, sin = R.esm_rope_tables(512, D, "cuda")
    want_q = R.rope_half_ref((qkv[:, : H * D].float() * 0.125).view(T, H, D).transpose(0, 1), cos[pos.long()],
                             sin[pos.long()]).transpose(0, 1).reshape(T, H * D)
    want_k = R.rope_half_ref(qkv[:, H * D: 2 * H * D].float().view(T, H, D).transpose(0, 1), cos[pos.long()],
                             sin[pos.long()]).transpose(0, 1).reshape(T, H * D)
    v_before = qkv[:, 2 * H * D:].clone()
    ops.rope_esm_(qkv, pos, cos.contiguous(), sin.contiguous(), H, D, 0.125)
    _close_bf16(qkv[:, : H * D], want_q, ulps=1.0, atol=1e-6)
    _close_bf16(qkv[:, H * D: 2 * H * D], want_k, ulps=1.0, atol=1e-6)
    assert torch.equal(qkv[:, 2 * H * D:], v_before)


def test_rope_llama_kvappend_bit_exact(ops):
    T, Hq, Hkv, D, BS = 200, 32, 8, 128, 16
    qkv = _randn((T, (Hq + 2 * Hkv) * D), 42)
    orig = qkv.clone()
    pos = torch.cat([torch.arange(120), torch.arange(80)]).to(torch.int32).cuda()
    nblk = 32
    perm = torch.randperm(nblk)[: (T + BS - 1) // BS]
    slot = (perm[torch.arange(T) // BS] * BS + torch.arange(T) % BS).to(torch.int32).cuda()
    slot[5] = -1
    cos, sin = R.llama_rope_tables(256, D, device="cuda")
    kc = torch.zeros((nblk, Hkv, BS, D), dtype=torch.bfloat16, device="cuda")
    vc = torch.zeros_like(kc)
    ops.rope_llama_kvappend_(qkv, pos, slot, cos, sin, kc, vc, Hq, Hkv, D, BS)
    q_want = R.llama_rope_bf16_ref(orig[:, : Hq * D].view(T, Hq, D), cos[pos.long()], sin[pos.long()])
    k_want = R.llama_rope_bf16_ref(orig[:, Hq * D: (Hq + Hkv) * D].view(T, Hkv, D), cos[pos.long()], sin[pos.long()])
    v_want = orig[:, (Hq + Hkv) * D:].view(T, Hkv, D)
    assert torch.equal(qkv[:, : Hq * D].view(T, Hq, D), q_want)
    assert torch.equal(qkv[:, Hq * D: (Hq + Hkv) * D].view(T, Hkv, D), k_want)
    for t in (0, 5, 17, 199):
        s = int(slot[t])
        if s < 0:
            continue
        assert torch.equal(kc[s // BS, :, s % BS, :], k_want[t])
        assert torch.equal(vc[s // BS, :, s % BS, :], v_want[t])
    written = (kc.abs().sum((1, 3)) > 0).sum()
    assert int(written) == T - 1


def test_final_ln_meanpool(ops):
    lens = [258, 3, 40, 130]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32).cuda()
    x = _randn((sum(lens), 1280), 43, scale=2.0, dtype=torch.float32)
    g = _randn((1280,), 44, dtype=torch.float32)
    b = _randn((1280,), 45, dtype=torch.float32)
    pooled, pooled_l2, hidden = ops.final_ln_meanpool(x, cu, g, b, 1e-5, want_hidden=True)
    h_want = R.layernorm_ref(x, g, b)
    assert torch.allclose(hidden, h_want, rtol=1e-4, atol=1e-4)
    p_want = R.meanpool_ref(h_want, cu.tolist())
    assert torch.allclose(pooled, p_want, rtol=1e-4, atol=1e-4), float((pooled - p_want).abs().max())
    l2_want = torch.nn.functional.normalize(p_want, dim=-1)
    _close_bf16(pooled_l2, l2_want, ulps=1.0, atol=1e-5)


def test_splice_gather(ops):
    embed = _randn((1000, 4096), 46)
    soft = _randn((16, 4096), 47)
    src = torch.tensor([5, 7, -1, -2, -16, 999, -(2 ** 31), 0], dtype=torch.int32).cuda()
    out = ops.splice_gather(src, embed, soft)
    assert torch.equal(out[0], embed[5]) and torch.equal(out[1], embed[7])
    assert torch.equal(out[2], soft[0]) and torch.equal(out[3], soft[1]) and torch.equal(out[4], soft[15])
    assert torch.equal(out[5], embed[999]) and bool((out[6] == 0).all()) and torch.equal(out[7], embed[0])


def test_argmax_eos(ops):
    B, V = 9, 128256
    logits = _randn((B, V), 48)
    logits[0, 77] = 100.0
    logits[0, 5000] = 100.0      # tie -> lowest index
    logits[1, V - 1] = 200.0     # last column
    logits[2, 128001] = 300.0    # EOS
    logits[3, 42] = 50.0
    finished = torch.zeros(B, dtype=torch.int32).cuda()
    finished[3] = 1
    eos = torch.tensor([128001, 128009], dtype=torch.int32).cuda()
    nxt = torch.zeros(B, dtype=torch.int32).cuda()
    out_ids = torch.full((B, 4), -7, dtype=torch.int32).cuda()
    n_unf = torch.tensor([B - 1], dtype=torch.int32).cuda()
    want_tok, want_fin = R.greedy_select_ref(logits, finished, [128001, 128009], 128001)
    ops.argmax_eos(logits, finished, eos, 128001, nxt, out_ids, 2, n_unf)
    assert torch.equal(nxt.long(), want_tok) and int(nxt[0]) == 77 and int(nxt[3]) == 128001
    assert torch.equal(out_ids[:, 2].long(), want_tok) and bool((out_ids[:, [0, 1, 3]] == -7).all())
    assert torch.equal(finished.bool(), want_fin)
    assert int(n_unf) == B - 2


@pytest.mark.parametrize("temperature,top_p,vocab", [(0.1, 0.7, 128256), (1.0, 0.9, 1000), (0.7, 0.5, 4099), (1.5, 1.0, 517)])
def test_sample_top_p_nucleus_and_frequencies(ops, temperature, top_p, vocab):
    """do_sample=True path (the reference's default decode, run_opus_ddp.py:126-128): the nucleus must be HF's, every
    draw must come from it, and the draw frequencies must follow the renormalised probabilities."""
    g = torch.Generator().manual_seed(vocab)
    base = (torch.randn(4, vocab, generator=g) * (2.0 if temperature >= 0.5 else 0.3)).to(torch.bfloat16)
    reps = 2048
    logits = base.repeat_interleave(reps, 0).cuda().contiguous()          # 4 distinct rows x 2048 independent draws
    n = logits.shape[0]
    fin = torch.zeros(n, dtype=torch.int32, device="cuda")
    nxt = torch.zeros(n, dtype=torch.int32, device="cuda")
    out = torch.zeros(n, 2, dtype=torch.int32, device="cuda")
    kept = torch.zeros(n, dtype=torch.int32, device="cuda")
    ops.sample_top_p(logits, temperature, top_p, 1234, fin, None, 0, nxt, out, 1, kept_count=kept)
    tok = out[:, 1].long().cpu()
    assert torch.equal(tok, nxt.long().cpu())
    keep = R.top_p_keep_mask(base, temperature, top_p)                    # [4, V]
    probs = R.top_p_probs(base, temperature, top_p)
    for r in range(4):
        t = tok[r * reps:(r + 1) * reps]
        # nucleus size (ties between equal bf16 logits at the boundary may be resolved differently: +-1)
        assert abs(int(kept[r * reps]) - int(keep[r].sum())) <= 1, (int(kept[r * reps]), int(keep[r].sum()))
        # every draw is at least as probable as the least probable token of HF's nucleus
        floor = base[r].float()[keep[r]].min()
        assert bool((base[r].float()[t] >= floor).all())
        # frequencies follow the renormalised distribution: total variation distance of 2048 draws
        freq = torch.bincount(t, minlength=vocab).float() / reps
        tv = 0.5 * float((freq - probs[r]).abs().sum())
        support = int(keep[r].sum())
        assert tv <= 0.04 + 0.6 * (support / reps) ** 0.5, (tv, support)
    # deterministic in the seed, different across seeds (when the nucleus has more than one token)
    out2 = torch.zeros_like(out)
    ops.sample_top_p(logits, temperature, top_p, 1234, fin, None, 0, nxt, out2, 1)
    assert torch.equal(out2[:, 1], out[:, 1])
    ops.sample_top_p(logits, temperature, top_p, 99, fin, None, 0, nxt, out2, 1)
    if int(keep.sum(1).min()) > 1:
        assert not torch.equal(out2[:, 1], out[:, 1])


def test_sample_top_p_limits_and_eos(ops):
    g = torch.Generator().manual_seed(7)
    logits = (torch.randn(16, 3001, generator=g) * 2).to(torch.bfloat16).cuda()
    fin = torch.zeros(16, dtype=torch.int32, device="cuda"); fin[3] = 1
    nxt = torch.zeros(16, dtype=torch.int32, device="cuda")
    out = torch.full((16, 1), -1, dtype=torch.int32, device="cuda")
    left = torch.tensor([15], dtype=torch.int32, device="cuda")
    greedy = logits.float().argmax(-1)
    eos = torch.tensor([int(greedy[5])], dtype=torch.int32, device="cuda")
    # temperature -> 0 (or a nucleus of one token) degenerates to greedy; finished rows emit pad; EOS rows finish
    ops.sample_top_p(logits, 1e-3, 0.9, 5, fin, eos, 77, nxt, out, 0, left)
    want = greedy.clone(); want[3] = 77
    assert torch.equal(out[:, 0].long(), want)
    assert int(fin[5]) == 1 and int(left) == 15 - int((greedy == greedy[5]).sum() - (1 if int(greedy[3]) == int(greedy[5]) else 0))
    out.fill_(-1); fin.zero_()
    ops.sample_top_p(logits, 1.0, 1e-6, 5, fin, None, 0, nxt, out, 0)
    assert torch.equal(out[:, 0].long(), greedy)


def test_lora_merge(ops):
    W = _randn((512, 256), 49, scale=0.05)
    A = _randn((16, 256), 50, scale=0.02)
    Bm = _randn((512, 16), 51, scale=0.02)
    want = R.bf16r(W.float() + 2.0 * (Bm.float() @ A.float()))
    ops.lora_merge_(W, A, Bm, 2.0)
    _close_bf16(W, want, ulps=1.0, atol=1e-6)


# ------------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("lens", [[258, 258], [1, 64, 65, 127, 128, 129, 300], [700]])
def test_attn_encoder_varlen(ops, lens):
    H, D = 20, 64
    T = sum(lens)
    qkv = _randn((T, 3 * H * D), 60)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32).cuda()
    q, k, v = qkv[:, : H * D], qkv[:, H * D: 2 * H * D], qkv[:, 2 * H * D:]
    got = ops.attn_varlen(q, k, v, cu, max(lens), H, H, D, False, 0.125)
    for b in range(len(lens)):
        s, e = int(cu[b]), int(cu[b + 1])
        want = R.attention_ref(q[s:e].view(-1, H, D), k[s:e].view(-1, H, D), v[s:e].view(-1, H, D), False, 0.125)
        _close_bf16(got[s:e].view(-1, H, D), want, ulps=2.0, atol=8e-3)


@pytest.mark.parametrize("lens", [[512, 512], [1, 17, 128, 200, 333], [515]])
def test_attn_prefill_causal_gqa(ops, lens):
    Hq, Hkv, D = 32, 8, 128
    T = sum(lens)
    qkv = _randn((T, (Hq + 2 * Hkv) * D), 61)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32).cuda()
    q, k, v = qkv[:, : Hq * D], qkv[:, Hq * D: (Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D:]
    sc = 1 / math.sqrt(D)
    got = ops.attn_varlen(q, k, v, cu, max(lens), Hq, Hkv, D, True, sc)
    for b in range(len(lens)):
        s, e = int(cu[b]), int(cu[b + 1])
        want = R.attention_ref(q[s:e].view(-1, Hq, D), k[s:e].view(-1, Hkv, D), v[s:e].view(-1, Hkv, D), True, sc)
        _close_bf16(got[s:e].view(-1, Hq, D), want, ulps=2.0, atol=8e-3)


@pytest.mark.parametrize("ctxs,Hq,Hkv", [([1, 16, 17, 528, 100, 33], 32, 8), ([544] * 8, 32, 8),
                                        ([5, 40, 300], 3, 1), ([70, 129, 16, 31], 28, 4), ([33, 257], 12, 2)])
def test_attn_decode_paged(ops, ctxs, Hq, Hkv):
    """GQA groups 4 (Llama-3-8B) and 3 / 7 / 6 (Qwen2.5 sizes): the group's query heads share one m16 tile."""
    D, BS = 128, 16
    B = len(ctxs)
    max_blocks = (max(ctxs) + BS - 1) // BS
    nblk = B * max_blocks + 3
    g = torch.Generator().manual_seed(62)
    perm = torch.randperm(nblk, generator=g)
    bt = perm[: B * max_blocks].view(B, max_blocks).to(torch.int32).cuda()
    kc = _randn((nblk, Hkv, BS, D), 63)
    vc = _randn((nblk, Hkv, BS, D), 64)
    q = _randn((B, Hq * D), 65)
    ctx = torch.tensor(ctxs, dtype=torch.int32).cuda()
    sc = 1 / math.sqrt(D)
    got = ops.attn_decode_paged(q, kc, vc, bt, ctx, Hq, Hkv, D, sc, BS)
    for b in range(B):
        n = ctxs[b]
        blocks = bt[b, : (n + BS - 1) // BS].long()
        kk = kc[blocks].permute(0, 2, 1, 3).reshape(-1, Hkv, D)[:n]
        vv = vc[blocks].permute(0, 2, 1, 3).reshape(-1, Hkv, D)[:n]
        want = R.attention_ref(q[b].view(1, Hq, D), kk, vv, False, sc)
        _close_bf16(got[b].view(1, Hq, D), want, ulps=2.0, atol=8e-3)


# ------------------------------------------------------------------------------------------------ norm-fused decode GEMMs
@pytest.mark.parametrize("rows,n_out,K,split", [(64, 4096, 4096, 4), (50, 4096, 14336, 4), (7, 512, 1024, 2), (32, 1024, 512, 2)])
def test_gemm_inkernel_splitk_residual_and_sumsq(rows, n_out, K, split):
    """o_proj / down of the norm-fused decode step: split-K reduced inside the kernel (same slice order as the reduce
    kernel), h = bf16(residual + bf16(sum)) in place, plus the per-32-feature-slab sums of squares of the stored values."""
    from opus_pllm_b200 import _lib as L, ops
    g = torch.Generator(device="cuda").manual_seed(rows + K)
    x = (torch.randn(rows, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(n_out, K, device="cuda", generator=g) * 0.05).bfloat16()
    res = torch.randn(rows, n_out, device="cuda", generator=g).bfloat16()
    # reference: the existing two-kernel path (split-K partials + fused reduce/residual)
    part = ops.gemm(x, w, epilogue=L.EPI_PARTIAL_F32, transposed=True, split_k=split)
    want_h = res.clone()
    ops.rmsnorm(None, None, partial=part, residual=res, h_out=want_h, normalise=False)
    h = res.clone()
    sumsq = torch.full((n_out // 32, 64), -1.0, device="cuda")
    ops.gemm_fused(x, w, epilogue=L.EPI_RES_BF16, residual=h, out=h, split_k=split, splitk_fixup=True, sumsq_out=sumsq)
    assert torch.equal(h, want_h)                                     # bit-identical: same partial sums, same order
    want_sq = want_h.float().pow(2).reshape(rows, n_out // 32, 32).sum(-1).T           # [slab, row]
    assert torch.allclose(sumsq[:, :rows], want_sq, rtol=1e-5, atol=1e-6)
    assert bool((sumsq[:, rows:] == -1.0).all())                      # rows beyond the batch are not touched


@pytest.mark.parametrize("rows,n_out,K,epi", [(64, 6144, 4096, "partial"), (64, 28672, 4096, "swiglu"), (19, 1024, 512, "bf16"),
                                              (40, 2048, 1024, "swiglu")])
def test_gemm_rmsnorm_on_load(rows, n_out, K, epi):
    """q|k|v / gate-up of the norm-fused decode step: the activation operand is the raw residual stream, normalised in
    shared memory on its way to the tensor cores. Against rmsnorm kernel + plain GEMM: identical except where the
    different order of the sum of squares moves rstd by an ulp."""
    from opus_pllm_b200 import _lib as L, ops
    g = torch.Generator(device="cuda").manual_seed(rows * 3 + K)
    h = (torch.randn(rows, K, device="cuda", generator=g) * 1.5).bfloat16()
    w = (torch.randn(n_out, K, device="cuda", generator=g) * 0.03).bfloat16()
    gamma = (1.0 + 0.1 * torch.randn(K, device="cuda", generator=g)).bfloat16()
    code = {"partial": L.EPI_PARTIAL_F32, "swiglu": L.EPI_SWIGLU, "bf16": L.EPI_BF16}[epi]
    split = 3 if epi == "partial" else 0
    xn = ops.rmsnorm(h, gamma, 1e-5)
    want = ops.gemm(xn, w, epilogue=code, transposed=True, split_k=split)
    sumsq = torch.zeros((K // 32, 64), device="cuda")
    sumsq[:, :rows] = h.float().pow(2).reshape(rows, K // 32, 32).sum(-1).T
    got = ops.gemm_fused(h, w, epilogue=code, split_k=split, norm_sumsq=sumsq, norm_gamma=gamma, norm_eps=1e-5)
    assert got.shape == want.shape
    a, b = got.float(), want.float()
    assert float(torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0)) >= 0.99999
    assert float((a - b).abs().max()) <= 2.0 ** -6 * float(b.abs().max())
    assert float((a != b).float().mean()) <= 0.02            # only rstd-ulp rows may differ at all


@pytest.mark.parametrize("rows,n_out,K,epi", [(256, 28672, 4096, "swiglu"), (200, 28672, 4096, "swiglu"), (256, 128256, 4096, "bf16"),
                                              (512, 28672, 4096, "swiglu"), (400, 128256, 4096, "bf16")])
def test_pair_kernel_streamk_tail_matches_single_cta(rows, n_out, K, epi):
    """Batch 129..256 swap-AB launches run on the CTA-pair kernel; when their 256-feature tiles do not fill whole waves of
    the 74 pairs the last partial wave is cut along K over all pairs (stream-K, per-rank fix-up). Against the single-CTA
    kernel (tunable gemm_2cta_tr = 0): same values up to the fp32 summation order of the split tiles."""
    from opus_pllm_b200 import _lib as L, ops
    g = torch.Generator(device="cuda").manual_seed(n_out + rows)
    x = (torch.randn(rows, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(n_out, K, device="cuda", generator=g) * 0.05).bfloat16()
    code = {"swiglu": L.EPI_SWIGLU, "bf16": L.EPI_BF16}[epi]
    lib = L.load()
    try:
        L.check(lib.opus_set_tunable(b"gemm_2cta_tr", 2))
        L.check(lib.opus_set_tunable(b"pair_streamk", 1))
        got = ops.gemm(x, w, epilogue=code, transposed=True)
        again = ops.gemm(x, w, epilogue=code, transposed=True)
        L.check(lib.opus_set_tunable(b"pair_streamk", 0))
        two_waves = ops.gemm(x, w, epilogue=code, transposed=True)      # the default: whole tiles only
        L.check(lib.opus_set_tunable(b"gemm_2cta_tr", 0))
        L.check(lib.opus_set_tunable(b"streamk_fill", 0))              # single-CTA kernel, whole tiles only
        want = ops.gemm(x, w, epilogue=code, transposed=True)
    finally:
        L.check(lib.opus_set_tunable(b"gemm_2cta_tr", 2))
        L.check(lib.opus_set_tunable(b"pair_streamk", 0))
        L.check(lib.opus_set_tunable(b"streamk_fill", 90))
    rem = (-(-n_out // 256) * -(-rows // 256)) % 74
    if 0 < rem * 10 <= 74:      # a tail of a few tiles is always cut along K (gate/up at batch 512: 3 waves + 2 tiles)
        assert float((two_waves != want).float().mean()) <= rem / (n_out // 256 * -(-rows // 256)) + 1e-6
        assert float((two_waves.float() - want.float()).abs().max()) <= 2.0 ** -7 * float(want.float().abs().max())
    else:
        assert torch.equal(two_waves, want)                          # same k order per tile: bit-identical
    assert torch.equal(got, again)                                   # deterministic (fixed fix-up order)
    a, b = got.float(), want.float()
    assert float(torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0)) >= 0.999999
    assert float((a - b).abs().max()) <= 2.0 ** -7 * float(b.abs().max())
    ref = (x.float() @ w.float().T)
    if epi == "swiglu":
        gte, up = ref[:, 0::2].bfloat16().float(), ref[:, 1::2].bfloat16().float()
        ref = torch.nn.functional.silu(gte).bfloat16().float() * up
    assert float(torch.nn.functional.cosine_similarity(a.flatten(), ref.flatten(), dim=0)) >= 0.9999
