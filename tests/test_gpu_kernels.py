"""GPU: every kernel K0-K7 through the C ABI against a plain torch fp32 computation of the same
op on seeded tensors (SURVEY.md §4 'kernel parity'). Tolerances are the 16-bit output rounding:
bf16 2^-8 relative, fp16 2^-11; fp32 outputs 1e-5."""
import math

import pytest
import torch

from arxiv_rag_b200 import _lib

pytestmark = pytest.mark.gpu

DT = {"bf16": (torch.bfloat16, _lib.ARB_DTYPE_BF16, 6e-3), "fp16": (torch.float16, _lib.ARB_DTYPE_F16, 8e-4)}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _rel_err(got, ref):
    return ((got.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("dt", ["bf16", "fp16"])
@pytest.mark.parametrize("shape", [(128, 256, 64), (300, 256, 768), (1000, 768, 768), (513, 768, 3072),
                                   (4096, 2304, 768), (1, 32, 8), (129, 96, 200)])
def test_gemm_mainloop_fp32_out(lib, cuda, dt, shape):
    """tcgen05 main loop alone: bf16/fp16 products are exact in fp32, only the summation order
    differs -> 2e-5 of the output range. Covers ragged M, N < tile, K tail (TMA zero fill)."""
    M, N, K = shape
    tdt, code, _ = DT[dt]
    torch.manual_seed(0)
    A = (torch.randn(M, K, device=cuda) * 0.5).to(tdt)
    B = (torch.randn(N, K, device=cuda) * 0.5).to(tdt)
    C = torch.full((M, N), float("nan"), device=cuda)
    _lib.check(lib.arb_gemm16_f32out(A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), N, M, N, K, code, _stream()))
    assert _rel_err(C, A.float() @ B.float().T) < 2e-5


@pytest.mark.parametrize("dt", ["bf16", "fp16"])
@pytest.mark.parametrize("epi", [0, 1, 2])
def test_gemm_epilogues(lib, cuda, dt, epi):
    """bias (modeling_mpnet.py:145-159), bias+GELU(erf) (:225-228), bias+residual (:183,:239-243)."""
    tdt, code, tol = DT[dt]
    torch.manual_seed(1)
    M, N, K = 777, 768, 768
    A = (torch.randn(M, K, device=cuda) * 0.3).to(tdt)
    B = (torch.randn(N, K, device=cuda) * 0.05).to(tdt)
    bias = torch.randn(N, device=cuda)
    R = torch.randn(M, N, device=cuda).to(tdt)
    ref = A.float() @ B.float().T + bias
    ref = [ref, torch.nn.functional.gelu(ref), ref + R.float()][epi]
    C = torch.zeros(M, N, device=cuda, dtype=tdt)
    _lib.check(lib.arb_gemm16(A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), N, bias.data_ptr(),
                              R.data_ptr() if epi == 2 else 0, N, M, N, K, epi, code, _stream()))
    assert _rel_err(C, ref) < tol


@pytest.fixture
def pair_schedule(lib):
    """Force the CTA-pair schedule (tcgen05 cta_group::2, clusters of two CTAs per 256x256 tile);
    auto only picks it for M > 256 * (SMs / 2)."""
    _lib.check(lib.arb_set_gemm_mode(2))
    yield
    _lib.check(lib.arb_set_gemm_mode(0))


@pytest.mark.parametrize("dt", ["bf16", "fp16"])
@pytest.mark.parametrize("shape", [(256, 256, 64), (128, 256, 128), (300, 512, 768), (4099, 2304, 768), (513, 768, 3072),
                                   (1, 32, 8), (777, 264, 200), (40000, 3072, 64)])
def test_gemm_pair_schedule_mainloop(lib, cuda, pair_schedule, dt, shape):
    """Same bar as test_gemm_mainloop_fp32_out through the CTA-pair kernel: ragged M (the second
    CTA's rows past the end), N not a multiple of the half tile, K tail, more tiles than pairs."""
    M, N, K = shape
    tdt, code, _ = DT[dt]
    torch.manual_seed(0)
    A = (torch.randn(M, K, device=cuda) * 0.5).to(tdt)
    B = (torch.randn(N, K, device=cuda) * 0.5).to(tdt)
    C = torch.full((M, N), float("nan"), device=cuda)
    _lib.check(lib.arb_gemm16_f32out(A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), N, M, N, K, code, _stream()))
    assert _rel_err(C, A.float() @ B.float().T) < 2e-5


@pytest.mark.parametrize("dt", ["bf16", "fp16"])
@pytest.mark.parametrize("epi", [0, 1, 2])
@pytest.mark.parametrize("shape", [(1777, 768, 768), (40000, 264, 128)])
def test_gemm_pair_schedule_epilogues(lib, cuda, pair_schedule, dt, epi, shape):
    """Epilogues of the CTA-pair kernel (double-buffered staging, residual chunks prefetched a tile
    ahead); N = 264 leaves staging chunks unused in the last column block."""
    tdt, code, tol = DT[dt]
    torch.manual_seed(1)
    M, N, K = shape
    A = (torch.randn(M, K, device=cuda) * 0.3).to(tdt)
    B = (torch.randn(N, K, device=cuda) * 0.05).to(tdt)
    bias = torch.randn(N, device=cuda)
    R = torch.randn(M, N, device=cuda).to(tdt)
    ref = A.float() @ B.float().T + bias
    ref = [ref, torch.nn.functional.gelu(ref), ref + R.float()][epi]
    C = torch.zeros(M, N, device=cuda, dtype=tdt)
    _lib.check(lib.arb_gemm16(A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), N, bias.data_ptr(),
                              R.data_ptr() if epi == 2 else 0, N, M, N, K, epi, code, _stream()))
    assert _rel_err(C, ref) < tol


def _row_partials(x16, parts):
    """(sum, sum of squares) over `parts` column blocks of 128, layout [parts][M] float2."""
    xf = x16.float()
    M = xf.shape[0]
    blocks = xf.view(M, parts, 128)
    return torch.stack([blocks.sum(2), (blocks * blocks).sum(2)], dim=2).permute(1, 0, 2).contiguous()  # [parts, M, 2]


@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("dt", ["bf16", "fp16"])
@pytest.mark.parametrize("shape", [(777, 768, 768), (3000, 384, 384), (70, 768, 3072)])
def test_gemm_layernorm_fold(lib, cuda, dt, shape, mode):
    """The LayerNorm-folding epilogues the encoder uses instead of LayerNorm passes
    (modeling_mpnet.py:210 / :242 post-LN): a producer writes pre-LN rows + row partials
    (epilogue 6, then 5 with the normalised residual), consumers apply LN through gamma-scaled
    weights, column sums and the partials (epilogues 3, 4). Reference: torch layer_norm + linear."""
    import torch.nn.functional as F

    tdt, code, tol = DT[dt]
    M, H, K = shape
    torch.manual_seed(5)
    parts = H // 128
    eps = 1e-5
    x = (torch.randn(M, H, device=cuda) * 1.7 + 0.4).to(tdt)  # pre-LN rows with a non-zero mean
    g = torch.randn(H, device=cuda) * 0.2 + 1.0
    b = torch.randn(H, device=cuda) * 0.1
    st_x = _row_partials(x, parts)
    ln_x = F.layer_norm(x.float(), (H,), g, b, eps)
    _lib.check(lib.arb_set_gemm_mode(mode))
    try:
        # consumer: C = [gelu](LN(x) W^T + bias) through folded weights
        N = 3 * H
        W = torch.randn(N, H, device=cuda) * 0.05
        bias = torch.randn(N, device=cuda)
        Wf = (W * g[None, :]).to(tdt)
        colsum = Wf.float().sum(1).contiguous()
        bias_f = (bias + W @ b).contiguous()
        ref = ln_x @ W.T + bias
        for epi, r in ((3, ref), (4, F.gelu(ref))):
            C = torch.zeros(M, N, device=cuda, dtype=tdt)
            _lib.check(lib.arb_gemm16_lnfold(x.data_ptr(), H, Wf.data_ptr(), H, C.data_ptr(), N, bias_f.data_ptr(), 0, 0,
                                             colsum.data_ptr(), 0, 0, st_x.data_ptr(), parts, H, 0, eps, M, N, H, epi, code, _stream()))
            assert _rel_err(C, r) < 2 * tol, (epi, _rel_err(C, r))
        # producer: C = A Wo^T + bo + LN(x) (5) / + x (6), plus the row partials of the rounded C
        A = (torch.randn(M, K, device=cuda) * 0.3).to(tdt)
        Wo = (torch.randn(H, K, device=cuda) * 0.05).to(tdt)
        bo = torch.randn(H, device=cuda)
        base = A.float() @ Wo.float().T + bo
        for epi, r in ((5, base + ln_x), (6, base + x.float())):
            C = torch.zeros(M, H, device=cuda, dtype=tdt)
            st_out = torch.full((parts, M, 2), float("nan"), device=cuda)
            _lib.check(lib.arb_gemm16_lnfold(A.data_ptr(), K, Wo.data_ptr(), K, C.data_ptr(), H, bo.data_ptr(), x.data_ptr(), H,
                                             0, g.data_ptr(), b.data_ptr(), st_x.data_ptr(), parts, H, st_out.data_ptr(), eps,
                                             M, H, K, epi, code, _stream()))
            assert _rel_err(C, r) < tol, (epi, _rel_err(C, r))
            want = _row_partials(r, parts)  # statistics of the fp32 values before the 16-bit rounding
            assert torch.allclose(st_out, want, rtol=2e-3, atol=0.05), (epi, (st_out - want).abs().max())
    finally:
        _lib.check(lib.arb_set_gemm_mode(0))


def test_gemm_schedules_agree_bitwise(lib, cuda):
    """All three schedules (128x256 single CTA, 256x256 CTA pairs, 128x128 narrow tiles) accumulate
    each output element over K in the same order, so they must agree bit for bit — the encoder's
    result does not depend on which one `auto` picks for a batch size."""
    torch.manual_seed(2)
    M, N, K = 3000, 768, 768
    A = (torch.randn(M, K, device=cuda) * 0.3).to(torch.bfloat16)
    B = (torch.randn(N, K, device=cuda) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=cuda)
    out = []
    for mode in (1, 2, 3):
        _lib.check(lib.arb_set_gemm_mode(mode))
        C = torch.zeros(M, N, device=cuda, dtype=torch.bfloat16)
        _lib.check(lib.arb_gemm16(A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), N, bias.data_ptr(), 0, N, M, N, K, 1,
                                  _lib.ARB_DTYPE_BF16, _stream()))
        out.append(C)
    _lib.check(lib.arb_set_gemm_mode(0))
    assert torch.equal(out[0], out[1]) and torch.equal(out[0], out[2])


def test_gemm_strided_operands(lib, cuda):
    """The encoder reads q|k|v column blocks and writes into wider buffers: lda/ldc > K/N."""
    torch.manual_seed(2)
    M, N, K = 500, 256, 128
    Abig = (torch.randn(M, 3 * K, device=cuda)).to(torch.bfloat16)
    B = (torch.randn(N, K, device=cuda) * 0.1).to(torch.bfloat16)
    Cbig = torch.zeros(M, 2 * N, device=cuda)
    A = Abig[:, K:2 * K]
    _lib.check(lib.arb_gemm16_f32out(A.data_ptr(), 3 * K, B.data_ptr(), K, Cbig[:, N:].data_ptr(), 2 * N, M, N, K,
                                     _lib.ARB_DTYPE_BF16, _stream()))
    assert _rel_err(Cbig[:, N:], A.float() @ B.float().T) < 2e-5
    assert (Cbig[:, :N] == 0).all()


def test_gemm_rejects_bad_arguments(lib, cuda):
    A = torch.zeros(8, 8, device=cuda, dtype=torch.bfloat16)
    C = torch.zeros(8, 40, device=cuda)
    assert lib.arb_gemm16_f32out(A.data_ptr(), 8, A.data_ptr(), 8, C.data_ptr(), 36, 8, 36, 8, _lib.ARB_DTYPE_BF16, _stream()) == -1
    assert b"multiple of 8" in lib.arb_last_error()
    assert lib.arb_gemm16_f32out(A.data_ptr(), 8, A.data_ptr(), 8, C.data_ptr(), 32, 8, 32, 8, 9, _stream()) == -1


@pytest.mark.parametrize("dt", ["bf16", "fp16"])
@pytest.mark.parametrize("H,eps", [(768, 1e-5), (768, 1e-12), (384, 1e-5)])
def test_layernorm(lib, cuda, dt, H, eps):
    tdt, code, tol = DT[dt]
    torch.manual_seed(3)
    rows = 1003
    x = (torch.randn(rows, H, device=cuda) * 3 + 0.5).to(tdt)
    g = torch.randn(H, device=cuda) * 0.1 + 1
    b = torch.randn(H, device=cuda) * 0.1
    out = torch.zeros_like(x)
    _lib.check(lib.arb_layernorm16(x.data_ptr(), g.data_ptr(), b.data_ptr(), out.data_ptr(), rows, H, eps, code, _stream()))
    assert _rel_err(out, torch.nn.functional.layer_norm(x.float(), (H,), g, b, eps)) < tol


@pytest.mark.parametrize("dt", ["bf16", "fp16"])
def test_embed_layernorm_and_position_ids(lib, cuda, dt):
    """MPNetEmbeddings (modeling_mpnet.py:72-96) with create_position_ids_from_input_ids
    (:889-897): pads keep position `padding_idx`, tokens count from padding_idx+1."""
    tdt, code, tol = DT[dt]
    torch.manual_seed(4)
    B, S, V, P, H = 6, 70, 1000, 514, 768
    ids = torch.randint(4, V, (B, S), device=cuda, dtype=torch.int32)
    for r, ln in enumerate([70, 1, 20, 0, 69, 33]):
        ids[r, ln:] = 1
    ids[5, 10] = 1  # a pad id in the middle of a row: position counting must skip it
    we = torch.randn(V, H, device=cuda) * 0.02
    pe = torch.randn(P, H, device=cuda) * 0.02
    g = torch.randn(H, device=cuda) * 0.1 + 1
    b = torch.randn(H, device=cuda) * 0.1
    out = torch.zeros(B * S, H, device=cuda, dtype=tdt)
    _lib.check(lib.arb_embed_layernorm(ids.data_ptr(), we.data_ptr(), pe.data_ptr(), g.data_ptr(), b.data_ptr(),
                                       out.data_ptr(), B, S, H, V, P, 1, 0, 1e-5, code, _stream()))
    m = (ids != 1).int()
    pos = (torch.cumsum(m, 1) * m).long() + 1
    ref = torch.nn.functional.layer_norm(we[ids.long()] + pe[pos], (H,), g, b, 1e-5).reshape(B * S, H)
    assert _rel_err(out, ref) < tol


@pytest.mark.parametrize("dt", ["bf16", "fp16"])
def test_pool_normalize(lib, cuda, dt):
    """Pooling(mean) + Normalize: sum(mask*x)/clamp(sum mask,1e-9), x/max(||x||,1e-12); an
    all-pad row gives the zero vector (not NaN); arbitrary (non-prefix) masks are honoured."""
    tdt, code, _ = DT[dt]
    torch.manual_seed(5)
    B, S, H = 6, 50, 768
    hid = torch.randn(B, S, H, device=cuda).to(tdt)
    lens = torch.tensor([50, 1, 20, 0, 49, 7], device=cuda)
    mask = (torch.arange(S, device=cuda)[None, :] < lens[:, None]).int().contiguous()
    mask[5, 30] = 1
    out = torch.full((B, H), float("nan"), device=cuda)
    _lib.check(lib.arb_pool_normalize(hid.data_ptr(), mask.data_ptr(), out.data_ptr(), B, S, H, code, _stream()))
    mm = mask.unsqueeze(-1).float()
    e = (hid.float() * mm).sum(1) / torch.clamp(mm.sum(1), min=1e-9)
    ref = torch.nn.functional.normalize(e, p=2, dim=1)
    assert (out - ref).abs().max().item() < 1e-6
    assert (out[3] == 0).all()


@pytest.mark.parametrize("dt", ["bf16", "fp16"])
@pytest.mark.parametrize("case", [(3, 64, [64, 1, 17]), (2, 200, [200, 77]), (2, 9, [9, 4])])
def test_attention_head_dim_32_no_bias(lib, cuda, dt, case):
    """BertSelfAttention shape of all-MiniLM-L6-v2: 12 heads of 32, no relative-position bias
    (NULL table), additive mask only."""
    tdt, code, _ = DT[dt]
    tol = 1.5e-2 if dt == "bf16" else 2e-3
    B, S, lens = case
    nH, dh = 12, 32
    H = nH * dh
    torch.manual_seed(8)
    qkv = torch.randn(B * S, 3 * H, device=cuda).to(tdt)
    mask = (torch.arange(S, device=cuda)[None, :] < torch.tensor(lens, device=cuda)[:, None]).int().contiguous()
    ctx = torch.zeros(B * S, H, device=cuda, dtype=tdt)
    _lib.check(lib.arb_attention16(qkv.data_ptr(), 0, 0, mask.data_ptr(), ctx.data_ptr(), B, S, nH, dh, code, 0, _stream()))
    q, k, v = [t.float().view(B, S, nH, dh).transpose(1, 2) for t in qkv.split(H, dim=1)]
    ext = (1.0 - mask[:, None, None, :].float()) * torch.finfo(torch.float32).min
    ref = (torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh) + ext, -1) @ v).transpose(1, 2).reshape(B * S, H)
    lens_t = torch.tensor(lens, device=cuda)
    live = (torch.arange(S, device=cuda)[None, :] < lens_t[:, None]).reshape(B * S)
    assert torch.isfinite(ctx.float()).all()
    assert _rel_err(ctx[live], ref[live]) < tol


def test_embed_layernorm_bert_positions(lib, cuda):
    """position_mode 1: absolute positions (BertEmbeddings), pads included."""
    torch.manual_seed(9)
    B, S, V, P, H = 3, 20, 500, 64, 384
    ids = torch.randint(1, V, (B, S), device=cuda, dtype=torch.int32)
    ids[1, 5:] = 0
    we = torch.randn(V, H, device=cuda) * 0.02
    pe = torch.randn(P, H, device=cuda) * 0.02
    g = torch.randn(H, device=cuda) * 0.1 + 1
    b = torch.randn(H, device=cuda) * 0.1
    out = torch.zeros(B * S, H, device=cuda, dtype=torch.float16)
    _lib.check(lib.arb_embed_layernorm(ids.data_ptr(), we.data_ptr(), pe.data_ptr(), g.data_ptr(), b.data_ptr(),
                                       out.data_ptr(), B, S, H, V, P, 0, 1, 1e-12, _lib.ARB_DTYPE_F16, _stream()))
    pos = torch.arange(S, device=cuda)[None, :].expand(B, S)
    ref = torch.nn.functional.layer_norm(we[ids.long()] + pe[pos], (H,), g, b, 1e-12).reshape(B * S, H)
    assert _rel_err(out, ref) < 8e-4


def _attention_reference(qkv, relb, mask, B, S, nH, dh, P):
    H = nH * dh
    q, k, v = [t.float().view(B, S, nH, dh).transpose(1, 2) for t in qkv.split(H, dim=1)]
    idx = torch.arange(S, device=qkv.device)
    bias = relb[:, idx[None, :] - idx[:, None] + (P - 1)]
    ext = (1.0 - mask[:, None, None, :].float()) * torch.finfo(torch.float32).min
    sc = q @ k.transpose(-1, -2) / math.sqrt(dh) + bias[None] + ext
    return (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(B * S, H)


@pytest.mark.parametrize("impl", [2, 3, 4])
@pytest.mark.parametrize("dt", ["bf16", "fp16"])
def test_attention_wide_bias_and_moving_maximum(lib, cuda, dt, impl):
    """What a trained relative-position table and peaked attention do to the softmax: biases spread
    over +-10 (29 log2 units between the most and the least favoured offset — the round-1 kernel's
    shift is only a bound of the row maximum, which pushed fp16 probabilities into subnormals), and
    keys whose scores grow along the sequence by far more than the lazy-rescale threshold, so that
    the sub-block pipelined kernel must take its accumulator-rescale path on most rows."""
    tdt, code, _ = DT[dt]
    tol = 1.5e-2 if dt == "bf16" else 3e-3
    B, S, nH, dh, P = 3, 384, 12, 64, 512
    H = nH * dh
    torch.manual_seed(12)
    qkv = torch.randn(B * S, 3 * H, device=cuda)
    ramp = torch.linspace(0.2, 3.0, S, device=cuda).repeat(B)[:, None]  # later keys score much higher
    qkv[:, H:2 * H] *= ramp
    qkv = qkv.to(tdt)
    relb = (torch.rand(nH, 2 * P - 1, device=cuda) * 20.0 - 10.0)
    lens = [384, 300, 97]
    mask = (torch.arange(S, device=cuda)[None, :] < torch.tensor(lens, device=cuda)[:, None]).int().contiguous()
    ctx = torch.zeros(B * S, H, device=cuda, dtype=tdt)
    rc = lib.arb_attention16(qkv.data_ptr(), relb.data_ptr(), P, mask.data_ptr(), ctx.data_ptr(), B, S, nH, dh, code, impl, _stream())
    if impl == 3 and rc == -4:
        pytest.skip("attention impl 3 is an experiment compiled only with -DARB_WITH_ATTENTION_TC2")
    _lib.check(rc)
    ref = _attention_reference(qkv, relb, mask, B, S, nH, dh, P)
    live = (torch.arange(S, device=cuda)[None, :] < torch.tensor(lens, device=cuda)[:, None]).reshape(B * S)
    assert torch.isfinite(ctx.float()).all()
    assert _rel_err(ctx[live], ref[live]) < tol


@pytest.mark.parametrize("impl", [1, 2, 4])
@pytest.mark.parametrize("dt", ["bf16", "fp16"])
@pytest.mark.parametrize("case", [(2, 384, [384, 250]), (3, 320, [320, 319, 40]), (2, 200, [200, 7])])
def test_attention_bucketed_bias(lib, cuda, dt, case, impl):
    """The relative-position table as MPNet really builds it: 32 buckets, constant beyond |j - i| = 91
    (modeling_mpnet.py:343-360). The 16-warp kernel reads one table entry per 32-key chunk where the
    chunk lies wholly in a constant tail; the result must not change."""
    tdt, code, _ = DT[dt]
    tol = 1.5e-2 if dt == "bf16" else 2e-3
    B, S, lens = case
    nH, dh, P = 12, 64, 512
    H = nH * dh
    torch.manual_seed(9)
    qkv = torch.randn(B * S, 3 * H, device=cuda).to(tdt)
    emb = torch.randn(32, nH, device=cuda) * 0.7
    rel = torch.arange(-(P - 1), P)
    buckets = torch.tensor([lib.arb_mpnet_relative_bucket(int(r), 32, 128) for r in rel], device=cuda)
    relb = emb[buckets].t().contiguous()  # [nH, 2P-1], entry (j - i) + P - 1
    mask = (torch.arange(S, device=cuda)[None, :] < torch.tensor(lens, device=cuda)[:, None]).int().contiguous()
    ctx = torch.zeros(B * S, H, device=cuda, dtype=tdt)
    _lib.check(lib.arb_attention16(qkv.data_ptr(), relb.data_ptr(), P, mask.data_ptr(), ctx.data_ptr(), B, S, nH, dh, code, impl, _stream()))
    q, k, v = [t.float().view(B, S, nH, dh).transpose(1, 2) for t in qkv.split(H, dim=1)]
    idx = torch.arange(S, device=cuda)
    bias = relb[:, idx[None, :] - idx[:, None] + (P - 1)]
    ext = (1.0 - mask[:, None, None, :].float()) * torch.finfo(torch.float32).min
    ref = (torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh) + bias[None] + ext, -1) @ v).transpose(1, 2).reshape(B * S, H)
    live = (torch.arange(S, device=cuda)[None, :] < torch.tensor(lens, device=cuda)[:, None]).reshape(B * S)
    assert torch.isfinite(ctx.float()).all()
    assert _rel_err(ctx[live], ref[live]) < tol


@pytest.mark.parametrize("impl", [1, 2, 3, 4])
@pytest.mark.parametrize("dt", ["bf16", "fp16"])
@pytest.mark.parametrize("case", [(3, 64, [64, 1, 17]), (4, 100, [100, 37, 0, 99]), (2, 384, [384, 200]), (1, 5, [3]),
                                  (5, 256, [256, 255, 130, 3, 0]), (3, 200, [200, 101, 100]), (2, 33, [33, 20]),
                                  (150, 320, [320] * 149 + [11])])
def test_attention(lib, cuda, dt, case, impl):
    """softmax(qk^T/8 + position_bias + (1-m)*finfo.min) v (modeling_mpnet.py:162-177); S not a
    multiple of the 64-key block, 1-token rows and an all-masked row (uniform attention, as the
    reference's fp32 arithmetic yields)."""
    tdt, code, _ = DT[dt]
    tol = 1.5e-2 if dt == "bf16" else 2e-3
    B, S, lens = case
    nH, dh, P = 12, 64, 512
    H = nH * dh
    torch.manual_seed(6)
    qkv = torch.randn(B * S, 3 * H, device=cuda).to(tdt)
    relb = torch.randn(nH, 2 * P - 1, device=cuda) * 0.5
    mask = (torch.arange(S, device=cuda)[None, :] < torch.tensor(lens, device=cuda)[:, None]).int().contiguous()
    ctx = torch.zeros(B * S, H, device=cuda, dtype=tdt)
    rc = lib.arb_attention16(qkv.data_ptr(), relb.data_ptr(), P, mask.data_ptr(), ctx.data_ptr(), B, S, nH, dh, code, impl, _stream())
    if impl == 3 and rc == -4:
        pytest.skip("attention impl 3 is an experiment compiled only with -DARB_WITH_ATTENTION_TC2")
    _lib.check(rc)
    q, k, v = [t.float().view(B, S, nH, dh).transpose(1, 2) for t in qkv.split(H, dim=1)]
    idx = torch.arange(S, device=cuda)
    bias = relb[:, idx[None, :] - idx[:, None] + (P - 1)]
    ext = (1.0 - mask[:, None, None, :].float()) * torch.finfo(torch.float32).min
    sc = q @ k.transpose(-1, -2) / math.sqrt(dh) + bias[None] + ext
    ref = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(B * S, H)
    assert torch.isfinite(ctx.float()).all()
    # query rows past a sequence's last real token are never pooled: the kernel only keeps them
    # finite (whole 16-row blocks past it are zero-filled, not computed); rows of real tokens
    # (and every row of an all-masked sequence) must match the reference
    lens_t = torch.tensor(lens, device=cuda)
    live = ((torch.arange(S, device=cuda)[None, :] < lens_t[:, None]) | (lens_t[:, None] == 0)).reshape(B * S)
    assert _rel_err(ctx[live], ref[live]) < tol
