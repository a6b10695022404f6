"""CPU study of the 16-bit rounding budget of the encode path (DESIGN.md 'Numerics').

Simulates the CUDA path's storage points (folded-LayerNorm schedule: pre-LN rows are stored in
16 bit, the GEMMs read the stored rows, statistics come from the fp32 values) with independent
formats for weights / activations / softmax probabilities, and prints the cosine against the fp32
oracle for short rows, where nothing averages the per-token noise.

    python tools/rounding_budget.py [--heavy] [--rows 64] [--seq 8]
"""
from __future__ import annotations

import argparse
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arxiv_rag_b200.weights import MPNetArch, synthetic_state_dict  # noqa: E402
from oracle import encode_oracle as eo  # noqa: E402

F = {
    "f32": lambda x: x,
    "bf16": lambda x: x.to(torch.bfloat16).float(),
    "fp16": lambda x: x.to(torch.float16).float(),
}


@torch.no_grad()
def forward(arch, sd, ids, mask, w="f32", a="f32", p="f32", resid="same", gelu="erf", w_ffn=None, f_fmt=None, y_fmt=None,
            w_up=None):
    """w / a / p: formats of the weights, the stored activations and the stored probabilities;
    resid: format of the stored pre-LN rows ('same' = activations' format)."""
    rw, ra, rp = F[w], F[a], F[p]
    rr = ra if resid == "same" else F[resid]
    rw_dn = F[w_ffn] if w_ffn else rw   # FFN-down weights
    rw_up = F[w_up] if w_up else rw     # FFN-up weights
    rf = F[f_fmt] if f_fmt else ra      # stored GELU output (A operand of FFN-down)
    ry = F[y_fmt] if y_fmt else rr      # stored pre-LN y rows (A operand of FFN-up and residual of FFN-down)
    t = lambda k: torch.as_tensor(np.asarray(sd[k])).float()
    ids = torch.as_tensor(np.asarray(ids)).long()
    mask = torch.as_tensor(np.asarray(mask)).long()
    B, S = ids.shape
    H, nH = arch.hidden_size, arch.num_heads
    dh = H // nH
    eps = arch.layer_norm_eps
    pos = eo.position_ids_from_input_ids(ids, arch.pad_token_id)
    x = t("embeddings.word_embeddings.weight")[ids] + t("embeddings.position_embeddings.weight")[pos]
    x = ra(eo._ln(x, t("embeddings.LayerNorm.weight"), t("embeddings.LayerNorm.bias"), eps))
    ctxpos = torch.arange(S)[:, None]
    mempos = torch.arange(S)[None, :]
    bucket = eo.relative_position_bucket(mempos - ctxpos, arch.relative_attention_num_buckets)
    pbias = t("encoder.relative_attention_bias.weight")[bucket].permute(2, 0, 1).unsqueeze(0)
    ext = (1.0 - mask[:, None, None, :].float()) * torch.finfo(torch.float32).min

    def ln_fold(stored, exact, g, b):  # statistics from the fp32 values, applied to the stored rows
        mu = exact.mean(-1, keepdim=True)
        var = (exact * exact).mean(-1, keepdim=True) - mu * mu
        rstd = torch.rsqrt(var.clamp_min(0) + eps)
        return (stored - mu) * rstd * g + b

    # x_norm(A operand) semantics: the GEMM reads the STORED pre-LN row times gamma-folded weights;
    # numerically that is LN(stored) W^T up to the weight rounding of W*gamma
    xs, xe, g_prev, b_prev = x, x, None, None  # stored rows, fp32 rows, LN params to apply (None = already normalised)
    for l in range(arch.num_layers):
        pfx = f"encoder.layer.{l}."
        W = lambda n: t(pfx + n + ".weight")
        bia = lambda n: t(pfx + n + ".bias")

        def lin_in(name, stored, exact, g, b, rw=rw):
            if g is None:
                return stored @ rw(W(name)).T + bia(name)
            mu = exact.mean(-1, keepdim=True)
            var = (exact * exact).mean(-1, keepdim=True) - mu * mu
            rstd = torch.rsqrt(var.clamp_min(0) + eps)
            Wg = rw(W(name) * g[None, :])
            return rstd * (stored @ Wg.T - mu * Wg.sum(1)[None, :]) + (bia(name) + W(name) @ b)

        q = ra(lin_in("attention.attn.q", xs, xe, g_prev, b_prev)).view(B, S, nH, dh).transpose(1, 2)
        k = ra(lin_in("attention.attn.k", xs, xe, g_prev, b_prev)).view(B, S, nH, dh).transpose(1, 2)
        v = ra(lin_in("attention.attn.v", xs, xe, g_prev, b_prev)).view(B, S, nH, dh).transpose(1, 2)
        scores = q @ k.transpose(-1, -2) / math.sqrt(dh) + pbias + ext
        m = scores.max(-1, keepdim=True).values
        e = torch.exp(scores - m)
        c = (rp(e) @ v) / e.sum(-1, keepdim=True)
        c = ra(c).transpose(1, 2).reshape(B, S, H)
        res = xs if g_prev is None else ln_fold(xs, xe, g_prev, b_prev)
        ye = c @ rw(W("attention.attn.o")).T + bia("attention.attn.o") + res
        ys = ry(ye)
        g1, b1 = t(pfx + "attention.LayerNorm.weight"), t(pfx + "attention.LayerNorm.bias")
        pre = lin_in("intermediate.dense", ys, ye, g1, b1, rw_up)
        f = torch.nn.functional.gelu(pre) if gelu == "erf" else 0.5 * pre * (1 + torch.tanh(0.8 * pre + 0.03475 * pre ** 3))
        f = rf(f)
        xe = f @ rw_dn(W("output.dense")).T + bia("output.dense") + ln_fold(ys, ye, g1, b1)
        xs = rr(xe)
        g_prev, b_prev = t(pfx + "output.LayerNorm.weight"), t(pfx + "output.LayerNorm.bias")
    h = ra(ln_fold(xs, xe, g_prev, b_prev))
    return eo.pool_normalize(h, mask).numpy()


def heavy_tail(sd, arch, seed=7):
    """Outlier channels x20 in two hidden dims, rel-pos spread +-8, LN gamma up to 5."""
    rng = np.random.default_rng(seed)
    sd = {k: np.array(v, copy=True) for k, v in sd.items()}
    H = arch.hidden_size
    out = rng.choice(H, 2, replace=False)
    sd["encoder.relative_attention_bias.weight"] = rng.uniform(-8, 8, sd["encoder.relative_attention_bias.weight"].shape).astype(np.float32)
    for l in range(arch.num_layers):
        p = f"encoder.layer.{l}."
        for nm in ("attention.LayerNorm", "output.LayerNorm"):
            g = sd[p + nm + ".weight"]
            g[rng.choice(H, 8, replace=False)] = rng.uniform(2, 5, 8)
        sd[p + "output.dense.weight"][out, :] *= 20.0
        sd[p + "output.dense.bias"][out] *= 20.0
    return sd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--heavy", action="store_true")
    ap.add_argument("--rows", type=int, default=48)
    ap.add_argument("--seq", type=int, default=8)
    ap.add_argument("--layers", type=int, default=12)
    args = ap.parse_args()
    arch = MPNetArch(num_layers=args.layers)
    sd = synthetic_state_dict(arch, 0)
    if args.heavy:
        sd = heavy_tail(sd, arch)
    rng = np.random.default_rng(5)
    ids, mask = eo.synthetic_tokens(args.rows, args.seq, seed=3)
    lengths = np.minimum(1 + np.arange(args.rows) % 4, args.seq)  # 1..4-token rows
    mask = (np.arange(args.seq)[None, :] < lengths[:, None]).astype(np.int32)
    ids = np.where(mask == 1, ids, 1)
    ids[:, 0] = rng.integers(4, 30000, args.rows)  # distinct single tokens
    ref = forward(arch, sd, ids, mask)
    ref2 = eo.oracle_encode(eo.reference_model(arch, sd), ids, mask)
    print("restatement vs transformers: min cos", float((ref * ref2).sum(1).min()))
    for name, kw in [
        ("all bf16", dict(w="bf16", a="bf16", p="bf16")),
        ("all bf16, tanh-fit gelu", dict(w="bf16", a="bf16", p="bf16", gelu="fit")),
        ("w fp16, a bf16, p bf16", dict(w="fp16", a="bf16", p="bf16")),
        ("w fp16, a bf16, p fp16", dict(w="fp16", a="bf16", p="fp16")),
        ("w f32,  a bf16, p bf16", dict(w="f32", a="bf16", p="bf16")),
        ("w bf16, a f32", dict(w="bf16")),
        ("w fp16, a bf16, resid fp32", dict(w="fp16", a="bf16", p="fp16", resid="f32")),
        ("w fp16, a fp16, resid bf16", dict(w="fp16", a="fp16", p="fp16", resid="bf16")),
        ("all fp16", dict(w="fp16", a="fp16", p="fp16")),
        ("fp16, FFN-down bf16 (W2, gelu out)", dict(w="fp16", a="fp16", p="fp16", w_ffn="bf16", f_fmt="bf16")),
        ("fp16, FFN up+down bf16 (y bf16)", dict(w="fp16", a="fp16", p="fp16", w_ffn="bf16", w_up="bf16", f_fmt="bf16", y_fmt="bf16")),
        ("fp16, FFN up+down W bf16 only", dict(w="fp16", a="fp16", p="fp16", w_ffn="bf16", w_up="bf16")),
    ]:
        out = forward(arch, sd, ids, mask, **kw)
        cos = (out * ref2).sum(1)
        by_len = [float(cos[lengths == n].min()) for n in (1, 2, 3, 4)]
        print(f"{name:32s} min {cos.min():.6f} mean {cos.mean():.6f}  min by length 1..4: "
              + " ".join(f"{c:.6f}" for c in by_len))


if __name__ == "__main__":
    main()
