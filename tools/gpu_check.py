"""Staged on-GPU bring-up checks with diagnostics (developer tool, not a test suite).

    python tools/gpu_check.py <stage> [...]

Every stage compares one kernel (through the C ABI) with a plain torch fp32 computation of the
same op and prints error statistics; on a mismatch it prints where the error lives (rows/cols)
so a single gpurun round-trip gives enough to debug. Stages are run as separate processes under
`timeout` by tools/gpu_check.sh so a hang cannot take the rest down.
"""
from __future__ import annotations

import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arxiv_rag_b200 import _lib  # noqa: E402

DEV = "cuda:0"
L = None


def lib():
    global L
    if L is None:
        L = _lib.lib()
    return L


def stream():
    return torch.cuda.current_stream().cuda_stream


def tdtype(d):
    return torch.float16 if d == "fp16" else torch.bfloat16


def dcode(d):
    return _lib.ARB_DTYPE_F16 if d == "fp16" else _lib.ARB_DTYPE_BF16


def report(name, got, ref, tol):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    rel = err.max().item() / max(ref.abs().max().item(), 1e-30)
    ok = bool(torch.isfinite(got).all()) and rel <= tol
    print(f"[{'OK ' if ok else 'BAD'}] {name}: max_abs_err={err.max().item():.3e} rel_to_max={rel:.3e} "
          f"ref_absmax={ref.abs().max().item():.3e} finite={bool(torch.isfinite(got).all())}", flush=True)
    if not ok and err.dim() == 2:
        bad = err > tol * ref.abs().max()
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print(f"      bad elems {int(bad.sum())}/{bad.numel()} rows[{rows[:8].tolist()}..{rows[-3:].tolist()}] n={len(rows)} "
              f"cols[{cols[:8].tolist()}..{cols[-3:].tolist()}] n={len(cols)}")
        r0 = int(rows[0]) if len(rows) else 0
        c0 = int(cols[0]) if len(cols) else 0
        print("      got", got[r0, c0:c0 + 8].tolist())
        print("      ref", ref[r0, c0:c0 + 8].tolist())
    return ok


def stage_gemm():
    ok = True
    torch.manual_seed(0)
    for d in ("bf16", "fp16"):
        for (M, N, K) in [(128, 256, 64), (128, 256, 128), (300, 256, 768), (1000, 768, 768), (4096, 2304, 768),
                          (513, 768, 3072), (20000, 3072, 768)]:
            A = (torch.randn(M, K, device=DEV) * 0.5).to(tdtype(d))
            B = (torch.randn(N, K, device=DEV) * 0.5).to(tdtype(d))
            C = torch.full((M, N), float("nan"), device=DEV, dtype=torch.float32)
            _lib.check(lib().arb_gemm16_f32out(A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), N, M, N, K, dcode(d), stream()))
            torch.cuda.synchronize()
            ref = A.float() @ B.float().T
            ok &= report(f"gemm_f32out {d} M{M} N{N} K{K}", C, ref, 2e-5)
    # epilogues
    for d in ("bf16", "fp16"):
        M, N, K = 777, 768, 768
        A = (torch.randn(M, K, device=DEV) * 0.3).to(tdtype(d))
        B = (torch.randn(N, K, device=DEV) * 0.05).to(tdtype(d))
        bias = torch.randn(N, device=DEV)
        R = torch.randn(M, N, device=DEV).to(tdtype(d))
        base = A.float() @ B.float().T + bias
        for epi, name, ref in [(0, "bias", base), (1, "bias_gelu", torch.nn.functional.gelu(base)),
                               (2, "bias_residual", base + R.float())]:
            C = torch.zeros(M, N, device=DEV, dtype=tdtype(d))
            _lib.check(lib().arb_gemm16(A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), N, bias.data_ptr(),
                                        R.data_ptr() if epi == 2 else 0, N, M, N, K, epi, dcode(d), stream()))
            torch.cuda.synchronize()
            ok &= report(f"gemm_{name} {d}", C, ref, 6e-3 if d == "bf16" else 8e-4)
    return ok


def stage_rowops():
    ok = True
    torch.manual_seed(1)
    for d in ("bf16", "fp16"):
        H = 768
        rows = 1000
        x = (torch.randn(rows, H, device=DEV) * 3 + 0.5).to(tdtype(d))
        g = torch.randn(H, device=DEV) * 0.1 + 1
        b = torch.randn(H, device=DEV) * 0.1
        out = torch.zeros_like(x)
        _lib.check(lib().arb_layernorm16(x.data_ptr(), g.data_ptr(), b.data_ptr(), out.data_ptr(), rows, H, 1e-5, dcode(d), stream()))
        ref = torch.nn.functional.layer_norm(x.float(), (H,), g, b, 1e-5)
        ok &= report(f"layernorm {d}", out, ref, 6e-3 if d == "bf16" else 8e-4)
        # embed + LN
        B, S, V, P = 5, 37, 1000, 514
        ids = torch.randint(4, V, (B, S), device=DEV, dtype=torch.int32)
        lens = [37, 1, 20, 0, 36]
        for r, ln in enumerate(lens):
            ids[r, ln:] = 1
        we = torch.randn(V, H, device=DEV) * 0.02
        pe = torch.randn(P, H, device=DEV) * 0.02
        out = torch.zeros(B * S, H, device=DEV, dtype=tdtype(d))
        _lib.check(lib().arb_embed_layernorm(ids.data_ptr(), we.data_ptr(), pe.data_ptr(), g.data_ptr(), b.data_ptr(),
                                             out.data_ptr(), B, S, H, V, P, 1, 0, 1e-5, dcode(d), stream()))
        m = (ids != 1).int()
        pos = (torch.cumsum(m, 1) * m).long() + 1
        ref = torch.nn.functional.layer_norm(we[ids.long()] + pe[pos], (H,), g, b, 1e-5).reshape(B * S, H)
        ok &= report(f"embed_ln {d}", out, ref, 6e-3 if d == "bf16" else 8e-4)
        # pool + normalize
        hid = torch.randn(B, S, H, device=DEV).to(tdtype(d))
        mask = (torch.arange(S, device=DEV)[None, :] < torch.tensor(lens, device=DEV)[:, None]).int().contiguous()
        out = torch.full((B, H), float("nan"), device=DEV)
        _lib.check(lib().arb_pool_normalize(hid.data_ptr(), mask.data_ptr(), out.data_ptr(), B, S, H, dcode(d), stream()))
        mm = mask.unsqueeze(-1).float()
        e = (hid.float() * mm).sum(1) / torch.clamp(mm.sum(1), min=1e-9)
        ref = torch.nn.functional.normalize(e, p=2, dim=1)
        ok &= report(f"pool_normalize {d}", out, ref, 1e-5)
    return ok


def stage_attention():
    ok = True
    torch.manual_seed(2)
    impls = [int(x) for x in os.environ.get("CHECK_ATTN_IMPLS", "1,2").split(",")]
    for d, impl in [(d, i) for i in impls for d in ("bf16", "fp16")]:
        for (B, S, lens) in [(3, 64, [64, 1, 17]), (4, 100, [100, 37, 0, 99]), (2, 384, [384, 200]), (5, 256, [256, 255, 130, 3, 0]),
                             (150, 320, [320] * 149 + [11]), (2, 33, [33, 20]), (1, 5, [3])]:
            if impl == 2 and S < 64:
                continue
            nH, dh = 12, 64
            H = nH * dh
            P = 512
            qkv = (torch.randn(B * S, 3 * H, device=DEV) * 1.0).to(tdtype(d))
            relb = torch.randn(nH, 2 * P - 1, device=DEV) * 0.5
            mask = (torch.arange(S, device=DEV)[None, :] < torch.tensor(lens, device=DEV)[:, None]).int().contiguous()
            ctx = torch.zeros(B * S, H, device=DEV, dtype=tdtype(d))
            _lib.check(lib().arb_attention16(qkv.data_ptr(), relb.data_ptr(), P, mask.data_ptr(), ctx.data_ptr(), B, S, nH, dh,
                                             dcode(d), impl, stream()))
            torch.cuda.synchronize()
            q, k, v = [t.float().view(B, S, nH, dh).transpose(1, 2) for t in qkv.split(H, dim=1)]
            idx = torch.arange(S, device=DEV)
            rel = idx[None, :] - idx[:, None] + (P - 1)
            bias = relb[:, rel]  # [nH,S,S]
            ext = (1.0 - mask[:, None, None, :].float()) * torch.finfo(torch.float32).min
            sc = q @ k.transpose(-1, -2) / math.sqrt(dh) + bias[None] + ext
            ref = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(B * S, H)
            lens_t = torch.tensor(lens, device=DEV)
            live = ((torch.arange(S, device=DEV)[None, :] < lens_t[:, None]) | (lens_t[:, None] == 0)).reshape(B * S)
            # rows past the last real token are unspecified (finite): compare live rows only
            ok &= bool(torch.isfinite(ctx.float()).all())
            ok &= report(f"attention impl{impl} {d} B{B} S{S}", ctx[live], ref[live], 1.5e-2 if d == "bf16" else 2e-3)
    return ok


def stage_search():
    from oracle import search_oracle as so

    ok = True
    for (Q, N, k, D) in [(5, 300, 10, 768), (130, 5000, 10, 768), (300, 20000, 100, 768), (64, 1000, 128, 128), (3, 7, 10, 768)]:
        c = so.synthetic_unit_rows(N, D, seed=0, bf16=True, plant_ties=True)
        q = so.synthetic_unit_rows(Q, D, seed=1, bf16=True)
        cd = torch.from_numpy(c).to(DEV).to(torch.bfloat16)
        qd = torch.from_numpy(q).to(DEV).to(torch.bfloat16)
        need = lib().arb_topk_search_workspace_bytes(_lib.ARB_DTYPE_BF16, Q, N, D, k)
        ws = torch.empty(max(need, 256), dtype=torch.uint8, device=DEV)
        os_ = torch.zeros(Q, k, device=DEV)
        oi = torch.zeros(Q, k, device=DEV, dtype=torch.int64)
        _lib.check(lib().arb_topk_search(qd.data_ptr(), cd.data_ptr(), _lib.ARB_DTYPE_BF16, Q, N, D, k, os_.data_ptr(), oi.data_ptr(),
                                         1000, ws.data_ptr(), ws.numel(), stream()))
        torch.cuda.synchronize()
        rep = so.check_topk(os_.cpu().numpy(), oi.cpu().numpy(), q, c, k, id_offset=1000)
        print(f"[{'OK ' if rep['ok'] else 'BAD'}] search bf16 Q{Q} N{N} k{k} D{D}: {rep}", flush=True)
        ok &= rep["ok"]
    for (Q, N, k, D) in [(7, 500, 10, 768), (200, 30000, 10, 768)]:
        c = so.synthetic_unit_rows(N, D, seed=0, plant_ties=True)
        q = so.synthetic_unit_rows(Q, D, seed=1)
        cd = torch.from_numpy(c).to(DEV)
        qd = torch.from_numpy(q).to(DEV)
        need = lib().arb_topk_search_workspace_bytes(_lib.ARB_DTYPE_F32, Q, N, D, k)
        ws = torch.empty(max(need, 256), dtype=torch.uint8, device=DEV)
        os_ = torch.zeros(Q, k, device=DEV)
        oi = torch.zeros(Q, k, device=DEV, dtype=torch.int64)
        _lib.check(lib().arb_topk_search(qd.data_ptr(), cd.data_ptr(), _lib.ARB_DTYPE_F32, Q, N, D, k, os_.data_ptr(), oi.data_ptr(),
                                         0, ws.data_ptr(), ws.numel(), stream()))
        torch.cuda.synchronize()
        rep = so.check_topk(os_.cpu().numpy(), oi.cpu().numpy(), q, c, k)
        print(f"[{'OK ' if rep['ok'] else 'BAD'}] search f32 Q{Q} N{N} k{k}: {rep}", flush=True)
        ok &= rep["ok"]
    return ok


def stage_encode():
    from arxiv_rag_b200.encoder import B200SentenceEncoder
    from arxiv_rag_b200.weights import ALL_MPNET_BASE_V2, MPNetArch, synthetic_state_dict
    from oracle import encode_oracle as eo

    ok = True
    small = MPNetArch(vocab_size=1000, num_layers=2, max_position_embeddings=514)
    for arch, name, n, S in [(small, "2-layer", 6, 40), (ALL_MPNET_BASE_V2, "12-layer", 8, 64)]:
        sd = synthetic_state_dict(arch, 0)
        model = eo.reference_model(arch, sd)
        ids, mask = eo.synthetic_tokens(n, S, vocab_size=arch.vocab_size, seed=1)
        ref = eo.oracle_encode(model, ids, mask)
        for d in ("bf16", "fp16", "bf16_pure"):
            enc = B200SentenceEncoder(sd, arch=arch, max_batch=16, max_seq=128, dtype=d)
            got = enc.encode((ids, mask), batch_size=16, normalize_embeddings=True)
            cos = (got * ref).sum(1)
            good = bool(np.isfinite(got).all()) and cos.min() >= (0.9995 if d == "bf16_pure" else 0.9999)
            print(f"[{'OK ' if good else 'BAD'}] encode {name} {d}: cos min {cos.min():.6f} mean {cos.mean():.6f} "
                  f"lens {mask.sum(1).tolist()} cos {np.round(cos, 6).tolist()}", flush=True)
            ok &= good
            enc.close()
    return ok


def _time(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def stage_perf():
    # GEMM shapes of one encoder layer at batch 1024 x 384 tokens, plus cuBLAS for scale
    M = 1024 * 384
    for (N, K, epi) in [(2304, 768, 0), (768, 768, 2), (3072, 768, 1), (768, 3072, 2)]:
        A = torch.randn(M, K, device=DEV).to(torch.bfloat16)
        B = (torch.randn(N, K, device=DEV) * 0.05).to(torch.bfloat16)
        C = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
        R = torch.randn(M, N, device=DEV).to(torch.bfloat16) if epi == 2 else None
        bias = torch.randn(N, device=DEV)
        ms = _time(lambda: _lib.check(lib().arb_gemm16(A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), N, bias.data_ptr(),
                                                       R.data_ptr() if R is not None else 0, N, M, N, K, epi, _lib.ARB_DTYPE_BF16, stream())))
        ms_cublas = _time(lambda: torch.matmul(A, B.T, out=C))
        fl = 2.0 * M * N * K
        print(f"gemm M{M} N{N} K{K} epi{epi}: {ms:.3f} ms {fl / ms / 1e9:.1f} TFLOP/s | cuBLAS(no epi) {ms_cublas:.3f} ms {fl / ms_cublas / 1e9:.1f} TFLOP/s", flush=True)
        del A, B, C, R
    # attention + row ops at the same size
    B_, S, H = 1024, 384, 768
    qkv = torch.randn(B_ * S, 3 * H, device=DEV).to(torch.bfloat16)
    relb = torch.randn(12, 1023, device=DEV)
    mask = torch.ones(B_, S, device=DEV, dtype=torch.int32)
    ctx = torch.empty(B_ * S, H, device=DEV, dtype=torch.bfloat16)
    fl = 4.0 * B_ * 12 * S * S * 64
    for impl in (1, 2):
        ms = _time(lambda: _lib.check(lib().arb_attention16(qkv.data_ptr(), relb.data_ptr(), 512, mask.data_ptr(), ctx.data_ptr(), B_, S, 12, 64,
                                                            _lib.ARB_DTYPE_BF16, impl, stream())), iters=10, warm=3)
        print(f"attention impl{impl} B{B_} S{S}: {ms:.3f} ms {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
    g = torch.ones(H, device=DEV)
    b = torch.zeros(H, device=DEV)
    out = torch.empty_like(ctx)
    ms = _time(lambda: _lib.check(lib().arb_layernorm16(ctx.data_ptr(), g.data_ptr(), b.data_ptr(), out.data_ptr(), B_ * S, H, 1e-5,
                                                        _lib.ARB_DTYPE_BF16, stream())))
    print(f"layernorm rows {B_ * S}: {ms:.3f} ms {2 * B_ * S * H * 2 / ms / 1e6:.0f} GB/s", flush=True)
    del qkv, ctx, out
    # full encode
    from arxiv_rag_b200.encoder import B200SentenceEncoder

    enc = B200SentenceEncoder(None, max_batch=1024, max_seq=384)
    for (bb, ss) in [(1024, 384), (1024, 256), (256, 384), (32, 128)]:
        ids = torch.randint(4, 30000, (bb, ss), device=DEV, dtype=torch.int32)
        m = torch.ones(bb, ss, device=DEV, dtype=torch.int32)
        ms = _time(lambda: enc.encode_tokens(ids, m), iters=3, warm=1)
        gf = 12 * ss * (14155776 + 3072 * ss) / 1e9
        print(f"encode B{bb} S{ss}: {ms:.2f} ms {bb / ms * 1e3:.0f} chunks/s {bb * gf / ms:.1f} TFLOP/s", flush=True)
    enc.close()
    del enc
    torch.cuda.empty_cache()
    # search
    for (Q, N, k) in [(128, 1_000_000, 10), (1024, 1_000_000, 10), (10000, 1_000_000, 10), (4096, 5_000_000, 10), (1024, 1_000_000, 100)]:
        c = torch.nn.functional.normalize(torch.randn(N, 768, device=DEV), dim=1).to(torch.bfloat16)
        q = torch.nn.functional.normalize(torch.randn(Q, 768, device=DEV), dim=1).to(torch.bfloat16)
        need = lib().arb_topk_search_workspace_bytes(_lib.ARB_DTYPE_BF16, Q, N, 768, k)
        ws = torch.empty(max(need, 256), dtype=torch.uint8, device=DEV)
        os_ = torch.zeros(Q, k, device=DEV)
        oi = torch.zeros(Q, k, device=DEV, dtype=torch.int64)
        ms = _time(lambda: _lib.check(lib().arb_topk_search(q.data_ptr(), c.data_ptr(), _lib.ARB_DTYPE_BF16, Q, N, 768, k, os_.data_ptr(),
                                                            oi.data_ptr(), 0, ws.data_ptr(), ws.numel(), stream())), iters=3, warm=1)
        fl = 2.0 * Q * N * 768
        print(f"search Q{Q} N{N} k{k}: {ms:.3f} ms {Q / ms * 1e3:.0f} q/s {fl / ms / 1e9:.1f} TFLOP/s corpus {N * 768 * 2 / ms / 1e6:.0f} GB/s", flush=True)
        del c, q, ws
    return True


def stage_gemm2():
    """CTA-pair GEMM (cta_group::2): parity against torch for every epilogue incl. ragged edges,
    then the encoder's four shapes timed in both schedules."""
    ok = True
    torch.manual_seed(3)
    _lib.check(lib().arb_set_gemm_mode(2))
    for d in ("bf16", "fp16"):
        for (M, N, K) in [(256, 256, 64), (256, 256, 768), (128, 256, 128), (300, 512, 768), (1000, 768, 768), (4099, 2304, 768),
                          (513, 768, 3072), (40000, 3072, 768), (777, 264, 64)]:
            A = (torch.randn(M, K, device=DEV) * 0.5).to(tdtype(d))
            B = (torch.randn(N, K, device=DEV) * 0.5).to(tdtype(d))
            C = torch.full((M, N), float("nan"), device=DEV, dtype=torch.float32)
            _lib.check(lib().arb_gemm16_f32out(A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), N, M, N, K, dcode(d), stream()))
            torch.cuda.synchronize()
            ok &= report(f"pair gemm_f32out {d} M{M} N{N} K{K}", C, A.float() @ B.float().T, 2e-5)
        for (M, N, K) in [(1777, 768, 768), (40000, 264, 128), (70000, 3072, 64)]:
            A = (torch.randn(M, K, device=DEV) * 0.3).to(tdtype(d))
            B = (torch.randn(N, K, device=DEV) * 0.05).to(tdtype(d))
            bias = torch.randn(N, device=DEV)
            R = torch.randn(M, N, device=DEV).to(tdtype(d))
            base = A.float() @ B.float().T + bias
            for epi, name, ref in [(0, "bias", base), (1, "bias_gelu", torch.nn.functional.gelu(base)), (2, "bias_residual", base + R.float())]:
                C = torch.zeros(M, N, device=DEV, dtype=tdtype(d))
                _lib.check(lib().arb_gemm16(A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), N, bias.data_ptr(),
                                            R.data_ptr() if epi == 2 else 0, N, M, N, K, epi, dcode(d), stream()))
                torch.cuda.synchronize()
                ok &= report(f"pair gemm_{name} {d} M{M} N{N} K{K}", C, ref, 6e-3 if d == "bf16" else 8e-4)
            del A, B, R, base
    M = 1024 * 384
    for (N, K, epi) in [(2304, 768, 0), (768, 768, 2), (3072, 768, 1), (768, 3072, 2)]:
        A = torch.randn(M, K, device=DEV).to(torch.bfloat16)
        B = (torch.randn(N, K, device=DEV) * 0.05).to(torch.bfloat16)
        C = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
        R = torch.randn(M, N, device=DEV).to(torch.bfloat16) if epi == 2 else None
        bias = torch.randn(N, device=DEV)
        res = []
        for mode in (1, 2):
            _lib.check(lib().arb_set_gemm_mode(mode))
            ms = _time(lambda: _lib.check(lib().arb_gemm16(A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), N, bias.data_ptr(),
                                                           R.data_ptr() if R is not None else 0, N, M, N, K, epi, _lib.ARB_DTYPE_BF16, stream())),
                       iters=10, warm=3)
            res.append(ms)
        fl = 2.0 * M * N * K
        print(f"gemm M{M} N{N} K{K} epi{epi}: single {res[0]:.3f} ms {fl / res[0] / 1e9:.1f} TF | pair {res[1]:.3f} ms {fl / res[1] / 1e9:.1f} TF", flush=True)
        del A, B, C, R
    _lib.check(lib().arb_set_gemm_mode(0))
    return ok


def stage_foldperf():
    """The encoder's four GEMM shapes with plain epilogues (+ LayerNorm pass) and with the folded-LN epilogues."""
    M, H, I = 1024 * 384, 768, 3072
    parts = H // 128
    x = torch.randn(M, H, device=DEV).to(torch.bfloat16)
    st = torch.rand(parts, M, 2, device=DEV) * 100 + 200
    g = torch.ones(H, device=DEV)
    b = torch.zeros(H, device=DEV)
    for (N, K, plain, fold) in [(2304, 768, 0, 3), (768, 768, 2, 5), (3072, 768, 1, 4), (768, 3072, 2, 5)]:
        A = torch.randn(M, K, device=DEV).to(torch.bfloat16)
        B = (torch.randn(N, K, device=DEV) * 0.05).to(torch.bfloat16)
        C = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
        bias = torch.randn(N, device=DEV)
        cs = torch.randn(N, device=DEV)
        so = torch.empty(N // 128, M, 2, device=DEV)
        R = x if plain == 2 else None
        t0 = _time(lambda: _lib.check(lib().arb_gemm16(A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), N, bias.data_ptr(),
                                                       R.data_ptr() if R is not None else 0, N, M, N, K, plain, _lib.ARB_DTYPE_BF16, stream())),
                   iters=10, warm=3)
        t1 = _time(lambda: _lib.check(lib().arb_gemm16_lnfold(A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), N, bias.data_ptr(),
                                                              x.data_ptr() if fold == 5 else 0, H, cs.data_ptr(), g.data_ptr(), b.data_ptr(),
                                                              st.data_ptr(), parts, H, so.data_ptr() if fold == 5 else 0, 1e-5, M, N, K, fold,
                                                              _lib.ARB_DTYPE_BF16, stream())), iters=10, warm=3)
        print(f"gemm N{N} K{K}: plain epi{plain} {t0:.3f} ms | folded epi{fold} {t1:.3f} ms", flush=True)
        del A, B, C
    return True


def stage_searchperf():
    """Search only: small-shard / small-Q (HBM-bound, the 8-GPU regime) and large-k cases, plus a
    per-kernel breakdown of one call from the CUPTI activity records (torch.profiler)."""
    shapes = [(1, 625_000, 10), (64, 625_000, 10), (64, 5_000_000, 10), (512, 625_000, 10), (1024, 1_000_000, 32),
              (1024, 1_000_000, 64), (1024, 1_000_000, 100), (8192, 2_500_000, 100), (1024, 1_000_000, 128)]
    for (Q, N, k) in shapes:
        c = torch.nn.functional.normalize(torch.randn(N, 768, device=DEV), dim=1).to(torch.bfloat16)
        q = torch.nn.functional.normalize(torch.randn(Q, 768, device=DEV), dim=1).to(torch.bfloat16)
        need = lib().arb_topk_search_workspace_bytes(_lib.ARB_DTYPE_BF16, Q, N, 768, k)
        ws = torch.empty(max(need, 256), dtype=torch.uint8, device=DEV)
        os_ = torch.zeros(Q, k, device=DEV)
        oi = torch.zeros(Q, k, device=DEV, dtype=torch.int64)
        call = lambda: _lib.check(lib().arb_topk_search(q.data_ptr(), c.data_ptr(), _lib.ARB_DTYPE_BF16, Q, N, 768, k, os_.data_ptr(),
                                                        oi.data_ptr(), 0, ws.data_ptr(), ws.numel(), stream()))
        ms = _time(call, iters=10, warm=3)
        fl = 2.0 * Q * N * 768
        print(f"search Q{Q} N{N} k{k}: {ms:.3f} ms {Q / ms * 1e3:.0f} q/s {fl / ms / 1e9:.1f} TFLOP/s corpus {N * 768 * 2 / ms / 1e6:.0f} GB/s",
              flush=True)
        if (Q, N) in ((1, 625_000), (64, 625_000), (1024, 1_000_000)):
            from torch.profiler import ProfilerActivity, profile

            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for _ in range(5):
                    call()
                torch.cuda.synchronize()
            for ev in prof.key_averages():
                print(f"    {ev.key[:70]:70s} n={ev.count} avg {ev.device_time_total / max(ev.count, 1):.1f} us", flush=True)
        del c, q, ws
    return True


STAGES = {"foldperf": stage_foldperf, "gemm2": stage_gemm2, "searchperf": stage_searchperf, "gemm": stage_gemm, "rowops": stage_rowops, "attention": stage_attention, "search": stage_search,
          "encode": stage_encode, "perf": stage_perf}

if __name__ == "__main__":
    names = sys.argv[1:] or list(STAGES)
    print("device:", torch.cuda.get_device_name(0), "lib:", _lib.LIB_PATH, flush=True)
    allok = True
    for n in names:
        t = time.time()
        try:
            r = STAGES[n]()
        except Exception as e:  # noqa: BLE001
            import traceback

            traceback.print_exc()
            r = False
        print(f"== stage {n}: {'PASS' if r else 'FAIL'} ({time.time() - t:.1f}s)", flush=True)
        allok &= bool(r)
    sys.exit(0 if allok else 1)
