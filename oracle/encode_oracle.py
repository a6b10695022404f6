"""CPU oracle for stage A (encode). TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import this package; the product (`arxiv_rag_b200/`) never does.

PARITY STATUS: the reference repository holds no golden vectors, tests or fixtures for this path
 and its own entry file cannot be imported (SyntaxError at
4-embed/generation/generate_embeddings_parallel.py:239; sentence-transformers not installed; no
weights offline). The arithmetic lives in third-party dependencies, unpinned in
3-chunks/pipeline/requirements.txt:10-13 (sentence-transformers>=2.2.2, transformers>=4.35.0,
torch>=2.1.0). The oracle is therefore pinned like this:
  * the encoder math IS the reference's own dependency: `transformers.MPNetModel` (installed
    5.5.0, modeling_mpnet.py:403-455) is imported and run unmodified (`reference_model`);
  * sentence-transformers' Pooling(mean) + Normalize are restated (3 lines, `pool_normalize`);
  * an independent plain-torch restatement (`restated_forward`) is checked against
    `MPNetModel` in the CPU tests and against committed golden fixtures generated from
    `MPNetModel` by tools/make_golden.py.
For sentence-transformers itself (absent) parity is "unpinned": its published algorithm is
restated from SURVEY.md §3.2.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def mpnet_config(arch):
    """transformers.MPNetConfig for an `arxiv_rag_b200.weights.MPNetArch`-like object."""
    from transformers import MPNetConfig

    return MPNetConfig(
        vocab_size=arch.vocab_size, hidden_size=arch.hidden_size, num_hidden_layers=arch.num_layers,
        num_attention_heads=arch.num_heads, intermediate_size=arch.intermediate_size,
        max_position_embeddings=arch.max_position_embeddings, layer_norm_eps=arch.layer_norm_eps,
        relative_attention_num_buckets=arch.relative_attention_num_buckets, hidden_act="gelu",
        hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, pad_token_id=arch.pad_token_id)


def reference_model(arch, state_dict: dict):
    """The reference dependency itself: transformers.MPNetModel(add_pooling_layer=False), fp32,
    eval, loaded with `state_dict` (numpy or torch values under HF names)."""
    if getattr(arch, "kind", "mpnet") == "bert":
        # all-MiniLM-L6-v2 is a BertModel (the reference's second encode site,
        # 3-chunks/pipeline/src/processors/text_processor.py:853-885, :1379-1396)
        from transformers import BertConfig, BertModel

        cfg = BertConfig(vocab_size=arch.vocab_size, hidden_size=arch.hidden_size, num_hidden_layers=arch.num_layers,
                         num_attention_heads=arch.num_heads, intermediate_size=arch.intermediate_size,
                         max_position_embeddings=arch.max_position_embeddings, layer_norm_eps=arch.layer_norm_eps,
                         hidden_act="gelu", hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                         pad_token_id=arch.pad_token_id, type_vocab_size=2)
        model = BertModel(cfg, add_pooling_layer=False).eval()
    else:
        from transformers import MPNetModel

        model = MPNetModel(mpnet_config(arch), add_pooling_layer=False).eval()
    sd = {k: torch.as_tensor(np.asarray(v)).float() if not torch.is_tensor(v) else v.float()
          for k, v in state_dict.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [m for m in missing if "position_ids" not in m]
    if missing or unexpected:
        raise RuntimeError(f"state dict mismatch: missing={missing} unexpected={unexpected}")
    return model


def pool_normalize(token_embeddings: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """sentence-transformers Pooling(mean) then Normalize (restated):
    e = sum_s(mask*tok) / clamp(sum_s mask, min=1e-9);  e / max(||e||_2, 1e-12)."""
    m = attention_mask.unsqueeze(-1).to(token_embeddings.dtype)
    e = (token_embeddings * m).sum(1) / torch.clamp(m.sum(1), min=1e-9)
    return torch.nn.functional.normalize(e, p=2, dim=1)


@torch.no_grad()
def oracle_encode(model, input_ids, attention_mask, batch_size: int = 16) -> np.ndarray:
    """What `SentenceTransformer.encode(..., normalize_embeddings=True, convert_to_numpy=True)`
    (generate_embeddings_parallel.py:146-153) returns for pre-tokenised input: float32 [n, H]."""
    ids = torch.as_tensor(np.asarray(input_ids)).long()
    mask = torch.as_tensor(np.asarray(attention_mask)).long()
    outs = []
    for i in range(0, ids.shape[0], batch_size):
        tok = model(input_ids=ids[i:i + batch_size], attention_mask=mask[i:i + batch_size])[0]
        e = pool_normalize(tok, mask[i:i + batch_size])
        e = torch.nn.functional.normalize(e, p=2, dim=1)  # encode(normalize_embeddings=True): 2nd pass
        outs.append(e.float().numpy())
    return np.concatenate(outs, 0) if outs else np.zeros((0, model.config.hidden_size), np.float32)


# ---------------------------------------------------------------------------------------------
# Independent restatement of the MPNet forward in plain torch (no transformers import).
# ---------------------------------------------------------------------------------------------
def relative_position_bucket(relative_position: torch.Tensor, num_buckets: int = 32, max_distance: int = 128):
    """modeling_mpnet.py:343-360."""
    n = -relative_position
    num_buckets //= 2
    ret = (n < 0).to(torch.long) * num_buckets
    n = torch.abs(n)
    max_exact = num_buckets // 2
    is_small = n < max_exact
    val_if_large = max_exact + (
        torch.log(n.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)
    ).to(torch.long)
    val_if_large = torch.min(val_if_large, torch.full_like(val_if_large, num_buckets - 1))
    return ret + torch.where(is_small, n, val_if_large)


def position_ids_from_input_ids(input_ids: torch.Tensor, padding_idx: int) -> torch.Tensor:
    """modeling_mpnet.py:889-897."""
    mask = input_ids.ne(padding_idx).int()
    return (torch.cumsum(mask, dim=1).type_as(mask) * mask).long() + padding_idx


def _ln(x, g, b, eps):
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), g, b, eps)


@torch.no_grad()
def restated_forward(arch, sd: dict, input_ids, attention_mask, round_fn=None, return_hidden=False):
    """MPNetModel.forward + pooling, written out op by op (modeling_mpnet.py:72-96, 145-183,
    210, 225-243, 293, 324-341; mask: modeling_utils.py:936-947).

    `round_fn`, if given, is applied at the points where the CUDA path stores bf16 (to study the
    rounding budget on CPU); None = pure fp32.
    """
    t = lambda k: torch.as_tensor(np.asarray(sd[k])).float() if not torch.is_tensor(sd[k]) else sd[k].float()
    r = round_fn or (lambda x: x)
    rw = (lambda x: x) if round_fn is None else round_fn  # weights are bf16 on the device
    ids = torch.as_tensor(np.asarray(input_ids)).long()
    mask = torch.as_tensor(np.asarray(attention_mask)).long()
    B, S = ids.shape
    H, nH = arch.hidden_size, arch.num_heads
    dh = H // nH
    eps = arch.layer_norm_eps
    pos = position_ids_from_input_ids(ids, arch.pad_token_id)
    x = t("embeddings.word_embeddings.weight")[ids] + t("embeddings.position_embeddings.weight")[pos]
    x = r(_ln(x, t("embeddings.LayerNorm.weight"), t("embeddings.LayerNorm.bias"), eps))
    ctxpos = torch.arange(S)[:, None]
    mempos = torch.arange(S)[None, :]
    bucket = relative_position_bucket(mempos - ctxpos, arch.relative_attention_num_buckets)
    pbias = t("encoder.relative_attention_bias.weight")[bucket].permute(2, 0, 1).unsqueeze(0)  # [1,nH,S,S]
    ext = (1.0 - mask[:, None, None, :].float()) * torch.finfo(torch.float32).min
    for l in range(arch.num_layers):
        p = f"encoder.layer.{l}."
        lin = lambda name, inp: inp @ rw(t(p + name + ".weight")).T + t(p + name + ".bias")
        q = r(lin("attention.attn.q", x)).view(B, S, nH, dh).transpose(1, 2)
        k = r(lin("attention.attn.k", x)).view(B, S, nH, dh).transpose(1, 2)
        v = r(lin("attention.attn.v", x)).view(B, S, nH, dh).transpose(1, 2)
        scores = q @ k.transpose(-1, -2) / math.sqrt(dh) + pbias + ext
        probs = torch.softmax(scores, dim=-1)
        c = r((r(probs) @ v)).transpose(1, 2).reshape(B, S, H)
        a = r(lin("attention.attn.o", c) + x)
        x1 = r(_ln(a, t(p + "attention.LayerNorm.weight"), t(p + "attention.LayerNorm.bias"), eps))
        f = r(torch.nn.functional.gelu(lin("intermediate.dense", x1)))
        o = r(lin("output.dense", f) + x1)
        x = r(_ln(o, t(p + "output.LayerNorm.weight"), t(p + "output.LayerNorm.bias"), eps))
    emb = pool_normalize(x, mask)
    if return_hidden:
        return emb.numpy(), x.numpy()
    return emb.numpy()


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


# ---------------------------------------------------------------------------------------------
# Synthetic inputs
# ---------------------------------------------------------------------------------------------
def synthetic_tokens(n: int, seq_len: int, vocab_size: int = 30527, seed: int = 1, full_length: bool = False,
                     pad_id: int = 1):
    """ids uniform in [4, vocab-2), position 0 = <s>=0, last valid = </s>=2, pad = 1.
    full_length=False draws lengths ~ U[1, S] and forces one 1-token row and one full row."""
    rng = np.random.default_rng(seed)
    ids = rng.integers(4, vocab_size - 1, size=(n, seq_len), dtype=np.int64)
    if full_length:
        lengths = np.full(n, seq_len)
    else:
        lengths = rng.integers(1, seq_len + 1, size=n)
        if n >= 1:
            lengths[0] = seq_len
        if n >= 2:
            lengths[1] = 1
    mask = (np.arange(seq_len)[None, :] < lengths[:, None]).astype(np.int64)
    ids[:, 0] = 0
    for r in range(n):
        if lengths[r] >= 2:
            ids[r, lengths[r] - 1] = 2
    ids = np.where(mask == 1, ids, pad_id)
    return ids.astype(np.int32), mask.astype(np.int32)
