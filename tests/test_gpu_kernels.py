"""Kernel-level parity through the C ABI (mmcm_gemm_bf16 / mmcm_layernorm / mmcm_attention) against a plain
PyTorch fp32 statement of the same op on the same bf16-rounded operands.  Tolerances are written per test."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from mmcm_b200 import lib as L
    return L.load()


def _check(code):
    from mmcm_b200 import lib as L
    L.check(code)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)


GEMM_SHAPES = [
    # (M, N, K)   every (N, K) pair the CLIP / SigLIP towers use, with ragged and tiny M
    (6400, 2304, 768), (6400, 768, 768), (6400, 3072, 768), (6400, 768, 3072),
    (9856, 1536, 512), (9856, 512, 512), (9856, 2048, 512), (9856, 512, 2048),
    (392, 768, 3072), (1568, 768, 768), (1568, 1536, 768),
    (8, 768, 768), (1, 512, 512), (129, 256, 64), (300, 128, 128), (20000, 512, 512),
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("impl", [0, 1, 2])   # 0 = tcgen05 CTA pair, 1 = SIMT validation, 2 = tcgen05 single CTA
def test_gemm_bias_bf16(lib, M, N, K, impl):
    if impl == 1 and M * N * K > 3e10:
        pytest.skip("SIMT validation kernel only on small shapes")
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    _check(lib.mmcm_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, K, 0, 0, out.data_ptr(), None, None,
                              0, 0, impl, _stream()))
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t() + bias
    # fp32 accumulation of exact bf16 products; the only rounding is the final bf16 store (rel 2^-9)
    err = (out.float() - ref).abs()
    assert (err <= 1e-2 * ref.abs() + 2e-2).all(), f"max err {err.max().item()}"
    assert torch.isfinite(out.float()).all()


@pytest.mark.parametrize("impl", [0, 2])
@pytest.mark.parametrize("act", [1, 2])
def test_gemm_bias_act(lib, act, impl):
    M, N, K = 1000, 3072, 768
    g = torch.Generator(device="cuda").manual_seed(act)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5 * 2).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    _check(lib.mmcm_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, K, 1, act, out.data_ptr(), None, None,
                              0, 0, impl, _stream()))
    torch.cuda.synchronize()
    pre = A.float() @ W.float().t() + bias
    ref = _quick_gelu(pre) if act == 1 else torch.nn.functional.gelu(pre, approximate="tanh")
    err = (out.float() - ref).abs()
    assert (err <= 1e-2 * ref.abs() + 1e-2).all(), f"max err {err.max().item()}"


@pytest.mark.parametrize("impl", [0, 2])
def test_gemm_bias_residual_in_place(lib, impl):
    M, N, K = 777, 768, 3072
    g = torch.Generator(device="cuda").manual_seed(11)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    x = torch.randn(M, N, device="cuda", generator=g)
    ref = x + A.float() @ W.float().t() + bias
    _check(lib.mmcm_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, K, 2, 0, x.data_ptr(), x.data_ptr(),
                              None, 0, 0, impl, _stream()))
    torch.cuda.synchronize()
    assert (x - ref).abs().max().item() < 2e-3   # fp32 out: summation order only


@pytest.mark.parametrize("impl", [0, 2])
def test_gemm_patch_epilogue(lib, impl):
    """rows of the im2col GEMM land at token 1..49 of each sample with the position embedding added."""
    Bn, P, T, N, K = 5, 49, 50, 768, 3072
    M = Bn * P
    g = torch.Generator(device="cuda").manual_seed(12)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.02).bfloat16()
    pos = torch.randn(T, N, device="cuda", generator=g)
    out = torch.full((Bn * T, N), 123.0, device="cuda")
    _check(lib.mmcm_gemm_bf16(A.data_ptr(), W.data_ptr(), None, M, N, K, 3, 0, out.data_ptr(), None, pos.data_ptr(),
                              P, T, impl, _stream()))
    torch.cuda.synchronize()
    ref = (A.float() @ W.float().t()).view(Bn, P, N) + pos[1:][None]
    o = out.view(Bn, T, N)
    assert (o[:, 1:] - ref).abs().max().item() < 2e-3
    assert (o[:, 0] == 123.0).all()                # class rows are not touched by the GEMM


# ---------------------------------------------------------------------------------------------- LN fold
# LayerNorm(x) W^T + b computed as rstd * (bf16(x) (W*gamma)^T - mean * colsum) + b' (include/mmcm.h, "LayerNorm
# folded into the GEMMs").  Reference: torch fp32 layer_norm + linear on the fp32 x.

def _slab_stats_ref(x):
    """float2 [D/128][rows]: (sum, M2 about the slab mean) of every 128-column slab."""
    rows, D = x.shape
    xs = x.double().view(rows, D // 128, 128)
    s = xs.sum(-1)
    m2 = ((xs - xs.mean(-1, keepdim=True)) ** 2).sum(-1)
    return torch.stack([s, m2], -1).permute(1, 0, 2).contiguous()      # [slab, row, 2]


@pytest.mark.parametrize("rows,D,affine", [(1, 512, False), (333, 512, False), (640, 768, True), (77, 1024, True)])
def test_prep_rows(lib, rows, D, affine):
    g = torch.Generator(device="cuda").manual_seed(rows + D)
    x = torch.randn(rows, D, device="cuda", generator=g) * 2 + 0.7
    gam = torch.randn(D, device="cuda", generator=g)
    bet = torch.randn(D, device="cuda", generator=g)
    ref = torch.nn.functional.layer_norm(x, (D,), gam, bet, 1e-5) if affine else x.clone()
    xb = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    st = torch.empty(D // 128, rows, 2, device="cuda")
    _check(lib.mmcm_prep_rows(x.data_ptr(), gam.data_ptr() if affine else None, bet.data_ptr() if affine else None,
                              1e-5, rows, D, xb.data_ptr(), st.data_ptr(), _stream()))
    torch.cuda.synchronize()
    assert (x - ref).abs().max().item() < 1e-4                         # in-place LayerNorm (or untouched rows)
    assert torch.equal(xb, x.bfloat16())                               # bf16 copy of exactly the fp32 rows left in x
    want = _slab_stats_ref(x)
    assert (st.double() - want).abs().max().item() <= 1e-4 * want.abs().max().item()


@pytest.mark.parametrize("M,N,K", [(9856, 512, 512), (6400, 768, 768), (777, 768, 3072), (300, 512, 2048),
                                   (31, 1024, 1024), (1, 512, 512)])
def test_gemm_resid_stats(lib, M, N, K):
    """x += A W^T + b in place, plus the bf16 copy and the slab statistics of the UPDATED rows."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    x = torch.randn(M, N, device="cuda", generator=g) * 1.5 + 0.3
    ref = x + A.float() @ W.float().t() + bias
    # bit-identity with the plain residual epilogue (TMA reduce-add): same (acc + bias) + x in fp32
    x2 = x.clone()
    _check(lib.mmcm_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, K, 2, 0, x2.data_ptr(), x2.data_ptr(),
                              None, 0, 0, 0, _stream()))
    xb = torch.full((M, N), 9.0, device="cuda", dtype=torch.bfloat16)
    st = torch.full((N // 128, M, 2), -1.0, device="cuda")
    _check(lib.mmcm_gemm_resid_stats(A.data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, K, x.data_ptr(), xb.data_ptr(),
                                     st.data_ptr(), _stream()))
    torch.cuda.synchronize()
    assert (x - ref).abs().max().item() < 2e-3
    assert torch.equal(x, x2)
    assert torch.equal(xb, x.bfloat16())
    want = _slab_stats_ref(x)
    assert (st.double() - want).abs().max().item() <= 1e-4 * max(want.abs().max().item(), 1.0)


@pytest.mark.parametrize("M,N,K,act", [(9856, 1536, 512, 0), (9856, 2048, 512, 1), (6400, 2304, 768, 0),
                                       (1000, 3072, 768, 2), (40, 256, 1024, 1), (1, 1536, 512, 0)])
def test_gemm_lnfold_matches_layernorm_linear(lib, M, N, K, act):
    """fold_ln + prep_rows + gemm_lnfold == act(Linear(LayerNorm(x))) up to the bf16 rounding of the operands."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K + act)
    x = torch.randn(M, K, device="cuda", generator=g) * (0.5 + 2.0 * torch.rand(M, 1, device="cuda", generator=g)) \
        + 0.5 * torch.randn(M, 1, device="cuda", generator=g)            # per-row scale and mean offsets
    W = torch.randn(N, K, device="cuda", generator=g) * K ** -0.5
    b = torch.randn(N, device="cuda", generator=g)
    gam = 1 + 0.1 * torch.randn(K, device="cuda", generator=g)
    bet = 0.1 * torch.randn(K, device="cuda", generator=g)
    q_rows, qs = (N // 3, 0.125) if N % 3 == 0 else (0, 1.0)
    wf = torch.empty(N, K, device="cuda", dtype=torch.bfloat16)
    rs = torch.empty(N, device="cuda")
    bf = torch.empty(N, device="cuda")
    _check(lib.mmcm_fold_ln(W.data_ptr(), b.data_ptr(), gam.data_ptr(), bet.data_ptr(), N, K, q_rows, qs, wf.data_ptr(),
                            rs.data_ptr(), bf.data_ptr(), _stream()))
    xb = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
    st = torch.empty(K // 128, M, 2, device="cuda")
    _check(lib.mmcm_prep_rows(x.data_ptr(), None, None, 1e-5, M, K, xb.data_ptr(), st.data_ptr(), _stream()))
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    _check(lib.mmcm_gemm_lnfold(xb.data_ptr(), wf.data_ptr(), bf.data_ptr(), st.data_ptr(), M, N, K, 1e-5, act,
                                out.data_ptr(), _stream()))
    torch.cuda.synchronize()
    scale = torch.ones(N, device="cuda")
    scale[:q_rows] = qs
    # weight side: rows of scale * W * gamma, centred, rounded once to bf16
    g_ = W * gam[None] * scale[:, None]
    want_w = g_ - g_.mean(-1, keepdim=True)
    assert (wf.float() - want_w).abs().max().item() <= 2.0 ** -8 * want_w.abs().max().item()
    assert (rs - wf.float().sum(-1)).abs().max().item() < 1e-4
    # what rounding leaves of the zero row sum: ~ sqrt(K) half-ulps of the entries; it multiplies mean(x) * rstd
    assert rs.abs().max().item() <= 4 * K ** 0.5 * 2.0 ** -9 * want_w.abs().max().item()
    assert (bf - scale * (b + W @ bet)).abs().max().item() < 1e-4
    pre = (torch.nn.functional.layer_norm(x, (K,), gam, bet, 1e-5) @ W.t() + b) * scale
    ref = pre if act == 0 else (_quick_gelu(pre) if act == 1 else torch.nn.functional.gelu(pre, approximate="tanh"))
    # operands rounded to bf16 (2^-9 each) over a K-term sum of O(1) normalised values: error ~ 2^-9 * |w|_2 * |xhat|_2
    # / sqrt(K)-ish; the bound below is 4x what the un-folded path (bf16(LN(x)) @ bf16(W)) measures on the same data
    err = (out.float() - ref).abs()
    assert (err <= 2e-2 * ref.abs() + 4e-2).all(), f"max err {err.max().item()}"
    h = torch.nn.functional.layer_norm(x, (K,), gam, bet, 1e-5).bfloat16().float()
    base = ((h @ W.bfloat16().float().t() + b) * scale - pre).abs().max().item()      # the separate-pass design
    if act == 0:
        assert err.max().item() <= max(4 * base, 4e-2), f"fold {err.max().item()} vs separate pass {base}"


@pytest.mark.parametrize("M", [1, 50, 77, 256])
@pytest.mark.parametrize("epi,N,K", [(0, 2304, 768), (1, 2048, 512), (2, 768, 3072), (2, 512, 512), (3, 768, 3072)])
def test_gemm_narrow_tiles_bit_identical(lib, M, epi, N, K):
    """GEMMs with one 256-row block run 64 / 128-column tiles (more CTA pairs share the weight stream).  The k order of
    every output element is the same as with 256-column tiles, so the results must be identical bits -- which is what
    keeps the pooled-rows last layer (M = B) bit-identical to the all-rows path."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K + epi)
    if epi == 3:
        M = max(49, (M // 49) * 49)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    pos = torch.randn(50, N, device="cuda", generator=g)
    x0 = torch.randn(M + M // 49 + 1, N, device="cuda", generator=g)
    outs = []
    for narrow in (1, 0):
        _check(lib.mmcm_set_option(None, b"narrow_tiles", narrow))
        if epi in (0, 1):
            out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            _check(lib.mmcm_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, K, epi, 1, out.data_ptr(), None,
                                      None, 0, 0, 0, _stream()))
        elif epi == 2:
            out = x0[:M].clone()
            _check(lib.mmcm_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, K, 2, 0, out.data_ptr(),
                                      out.data_ptr(), None, 0, 0, 0, _stream()))
        else:
            out = x0.clone()
            _check(lib.mmcm_gemm_bf16(A.data_ptr(), W.data_ptr(), None, M, N, K, 3, 0, out.data_ptr(), None,
                                      pos.data_ptr(), 49, 50, 0, _stream()))
        torch.cuda.synchronize()
        outs.append(out)
    _check(lib.mmcm_set_option(None, b"narrow_tiles", 1))
    assert torch.equal(outs[0], outs[1])
    if epi == 2:
        ref = x0[:M] + A.float() @ W.float().t() + bias
        assert (outs[0] - ref).abs().max().item() < 2e-3


def test_gemm_rejects_bad_shapes(lib):
    A = torch.zeros(4, 96, device="cuda", dtype=torch.bfloat16)
    W = torch.zeros(128, 96, device="cuda", dtype=torch.bfloat16)
    out = torch.zeros(4, 128, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        _check(lib.mmcm_gemm_bf16(A.data_ptr(), W.data_ptr(), None, 4, 128, 96, 0, 0, out.data_ptr(), None, None, 0, 0,
                                  0, _stream()))


@pytest.mark.parametrize("D", [512, 768])
@pytest.mark.parametrize("rows", [1, 50, 12800])
def test_layernorm(lib, D, rows):
    g = torch.Generator(device="cuda").manual_seed(rows + D)
    x = torch.randn(rows, D, device="cuda", generator=g) * 3 + 1
    gam = torch.randn(D, device="cuda", generator=g)
    bet = torch.randn(D, device="cuda", generator=g)
    ob = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    of = torch.empty(rows, D, device="cuda")
    _check(lib.mmcm_layernorm(x.data_ptr(), gam.data_ptr(), bet.data_ptr(), 1e-5, rows, D, ob.data_ptr(), of.data_ptr(),
                              _stream()))
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x, (D,), gam, bet, 1e-5)
    assert (of - ref).abs().max().item() < 1e-4               # fp32 in, fp32 statistics, fp32 out
    assert (ob.float() - ref).abs().max().item() <= 8e-3 * ref.abs().max().item() + 1e-3   # bf16 store


def _ref_attention(qkv, B, T, H, causal, kvalid):
    D = H * 64
    q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)   # [B,H,T,64] each; q is pre-scaled
    s = q @ k.transpose(-1, -2)
    allow = torch.ones(B, 1, T, T, dtype=torch.bool, device=qkv.device)
    if causal:
        allow = allow & torch.ones(T, T, dtype=torch.bool, device=qkv.device).tril()
    if kvalid is not None:
        allow = allow & (kvalid.view(B, 1, 1, T) != 0)
    s = s.masked_fill(~allow, float("-inf"))
    dead = ~allow.any(-1, keepdim=True)
    p = torch.softmax(s.masked_fill(dead, 0.0), -1).masked_fill(dead, 0.0)
    return (p @ v).transpose(1, 2).reshape(B * T, D)


@pytest.mark.parametrize("T,H,causal,masked", [
    (50, 12, 0, False), (77, 8, 1, False), (77, 8, 1, True), (64, 12, 0, True), (196, 12, 0, False),
    (11, 8, 1, True), (33, 8, 0, True), (128, 4, 1, True), (250, 2, 0, True)])
@pytest.mark.parametrize("impl", [1, 2])     # 1 = mma.sync kernel, 2 = tcgen05 kernel (T <= 128; falls back above)
def test_attention(lib, T, H, causal, masked, impl):
    _check(lib.mmcm_set_option(None, b"attention_impl", impl))     # defaults of the stand-alone kernels (no handle)
    B = 6
    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(T * 13 + H)
    qkv = torch.randn(B * T, 3 * D, device="cuda", generator=g)
    qkv[:, :D] *= 0.125 * 2.0     # scale folded into q; x2 for peaky softmaxes
    qkv = qkv.bfloat16()
    kvalid = None
    if masked:
        lens = torch.randint(1, T + 1, (B,), device="cuda", generator=g)
        kvalid = (torch.arange(T, device="cuda")[None] < lens[:, None]).to(torch.uint8)
        kvalid[0] = 0             # fully masked sample: every query row must come out exactly 0
        kvalid = kvalid.contiguous()
    out = torch.full((B * T, D), 7.0, device="cuda", dtype=torch.bfloat16)
    _check(lib.mmcm_attention(qkv.data_ptr(), None if kvalid is None else kvalid.data_ptr(), B, T, H, causal,
                              out.data_ptr(), _stream()))
    torch.cuda.synchronize()
    ref = _ref_attention(qkv, B, T, H, causal, kvalid)
    # P is rounded to bf16 before P V and the output is stored as bf16: 2^-8 relative on O(1) values
    assert (out.float() - ref).abs().max().item() < 3e-2
    if masked:
        assert (out.view(B, T, D)[0] == 0).all()
    _check(lib.mmcm_set_option(None, b"attention_impl", 0))


def _attention_inputs(B, T, H, masked, seed):
    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(seed)
    qkv = torch.randn(B * T, 3 * D, device="cuda", generator=g)
    qkv[:, :D] *= 0.25
    qkv = qkv.bfloat16()
    kvalid = None
    if masked:
        lens = torch.randint(1, T + 1, (B,), device="cuda", generator=g)
        kvalid = (torch.arange(T, device="cuda")[None] < lens[:, None]).to(torch.uint8)
        kvalid[0] = 0
        kvalid[1, ::3] = 0           # holes, not only a padded tail
        kvalid = kvalid.contiguous()
    return qkv, kvalid


def _run_attention(lib, impl, qkv, kvalid, B, T, H, causal):
    _check(lib.mmcm_set_option(None, b"attention_impl", impl))
    out = torch.full((B * T, H * 64), 7.0, device="cuda", dtype=torch.bfloat16)
    try:
        _check(lib.mmcm_attention(qkv.data_ptr(), None if kvalid is None else kvalid.data_ptr(), B, T, H, causal,
                                  out.data_ptr(), _stream()))
        torch.cuda.synchronize()
    finally:
        _check(lib.mmcm_set_option(None, b"attention_impl", 0))
    return out


@pytest.mark.parametrize("T,H,causal,masked", [
    (77, 8, 1, True), (77, 8, 1, False), (77, 8, 0, True), (50, 12, 0, False), (64, 12, 0, True), (40, 8, 1, True),
    (80, 8, 0, True), (33, 8, 1, False), (48, 12, 1, True)])
def test_attention_ring_is_bit_identical_to_the_cp_async_kernel(lib, T, H, causal, masked):
    """attention_ring.cuh keeps the fragments and the order of operations of attention.cuh and only skips key tiles
    whose probabilities are exact zeros: same bits.  B * H > 148 SMs x 8 stages, so every stage of the ring is reused
    and the mbarrier phases wrap."""
    B = 170
    qkv, kvalid = _attention_inputs(B, T, H, masked, T * 7 + H + causal)
    a = _run_attention(lib, 1, qkv, kvalid, B, T, H, causal)
    b = _run_attention(lib, 3, qkv, kvalid, B, T, H, causal)
    assert torch.equal(a, b)
    ref = _ref_attention(qkv[: 6 * T], 6, T, H, causal, None if kvalid is None else kvalid[:6])
    assert (b[: 6 * T].float() - ref).abs().max().item() < 3e-2
    if masked:
        assert (b.view(B, T, H * 64)[0] == 0).all()


@pytest.mark.parametrize("T,causal", [(77, 1), (77, 0), (50, 0)])
def test_attention_ring_isolates_non_finite_neighbours(lib, T, causal):
    """The TMA box of a sample covers 16 * ceil(T / 16) rows: the rows behind the sequence belong to the NEXT sample.
    A corrupt neighbour (NaN / inf in its q, k, v) must not reach this sample through 0 * NaN."""
    B, H = 40, 8
    qkv, _ = _attention_inputs(B, T, H, False, 99 + T)
    clean = _run_attention(lib, 3, qkv, None, B, T, H, causal)
    bad = qkv.clone().view(B, T, -1)
    bad[1::2] = float("nan")
    bad[3, 0] = float("inf")
    got = _run_attention(lib, 3, bad.view(B * T, -1).contiguous(), None, B, T, H, causal).view(B, T, -1)
    assert torch.equal(got[0::2], clean.view(B, T, -1)[0::2])
    assert torch.isfinite(got[0::2].float()).all()
