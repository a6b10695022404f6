"""Generate the committed golden fixtures under tests/golden/ (run in the dev container).

The reference holds no golden vectors for this path, so these are minted from
the reference's own dependency: `transformers.MPNetModel` (installed 5.5.0) run in fp32 on
seeded synthetic weights/inputs, plus the restated pooling/normalise; and, for search, from
NumPy fp32 `Q @ C.T` cross-checked against `torch.topk`. The generated arrays are small; the
weights are NOT stored — they are re-derived from the seed (numpy PCG64 is stable) and pinned by
a checksum.

    python tools/make_golden.py
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from arxiv_rag_b200.weights import ALL_MPNET_BASE_V2, MPNetArch, heavy_tail_state_dict, synthetic_state_dict  # noqa: E402
from oracle import encode_oracle as eo  # noqa: E402
from oracle import search_oracle as so  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def weights_digest(sd: dict) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k]).tobytes())
    return h.hexdigest()


def make_encode(name: str, arch: MPNetArch, seed: int, n: int, S: int, tok_seed: int, heavy: bool = False,
                lengths=None):
    sd = heavy_tail_state_dict(arch, seed) if heavy else synthetic_state_dict(arch, seed)
    model = eo.reference_model(arch, sd)  # transformers.MPNetModel, fp32
    ids, mask = eo.synthetic_tokens(n, S, vocab_size=arch.vocab_size, seed=tok_seed)
    if lengths is not None:  # explicit row lengths (short rows are where 16-bit arithmetic is weakest)
        lengths = np.asarray(lengths)
        mask = (np.arange(S)[None, :] < lengths[:, None]).astype(np.int32)
        ids = np.where(mask == 1, ids, arch.pad_token_id).astype(np.int32)
    emb = eo.oracle_encode(model, ids, mask)
    with torch.no_grad():
        hidden = model(input_ids=torch.from_numpy(ids).long(), attention_mask=torch.from_numpy(mask).long())[0].numpy()
    np.savez_compressed(
        os.path.join(OUT, name), ids=ids, mask=mask, embeddings=emb.astype(np.float32),
        hidden_row0=hidden[0].astype(np.float32), weight_seed=seed, token_seed=tok_seed,
        weights_sha256=weights_digest(sd), heavy_tail=heavy,
        arch=np.array([arch.vocab_size, arch.max_position_embeddings, arch.hidden_size, arch.num_layers,
                       arch.num_heads, arch.intermediate_size, arch.relative_attention_num_buckets,
                       arch.pad_token_id]),
        layer_norm_eps=arch.layer_norm_eps)
    print(name, emb.shape, "norms", np.linalg.norm(emb, axis=1))


def make_search(name: str, Q: int, N: int, D: int, k: int, bf16: bool, store_data: bool):
    c = so.synthetic_unit_rows(N, D, seed=0, bf16=bf16, plant_ties=True)
    q = so.synthetic_unit_rows(Q, D, seed=1, bf16=bf16)
    s, i = so.oracle_search(q, c, k)
    # cross-check the id sets against torch.topk (scores only; torch's tie order is unspecified)
    ts, ti = torch.topk(torch.from_numpy(q) @ torch.from_numpy(c).T, k, dim=1)
    assert np.allclose(np.sort(ts.numpy(), 1), np.sort(s, 1), atol=1e-6), "oracle vs torch.topk scores"
    payload = dict(scores=s, ids=i, Q=Q, N=N, D=D, k=k, bf16=bf16, corpus_seed=0, query_seed=1)
    if store_data:
        payload.update(corpus=c, queries=q)
    np.savez_compressed(os.path.join(OUT, name), **payload)
    print(name, s.shape, "first ids", i[0, :5])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    tiny = MPNetArch(vocab_size=1000, num_layers=2)
    make_encode("encode_tiny_2layer.npz", tiny, seed=0, n=3, S=16, tok_seed=5)
    make_encode("encode_mpnet_base_b4_s32.npz", ALL_MPNET_BASE_V2, seed=0, n=4, S=32, tok_seed=7)
    # heavy-tailed stand-in for trained weights (outlier channels, wide relative-position table,
    # large LayerNorm gains), rows of 1..96 tokens
    make_encode("encode_heavy_tail_b10_s96.npz", ALL_MPNET_BASE_V2, seed=0, n=10, S=96, tok_seed=11, heavy=True,
                lengths=[96, 1, 2, 3, 5, 9, 17, 33, 64, 80])
    make_search("search_64x32_k5.npz", Q=8, N=64, D=32, k=5, bf16=False, store_data=True)
    make_search("search_2000x768_k10_bf16.npz", Q=16, N=2000, D=768, k=10, bf16=True, store_data=False)
    make_search("search_2000x768_k10_f32.npz", Q=16, N=2000, D=768, k=10, bf16=False, store_data=False)
