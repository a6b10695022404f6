"""CPU oracle for stage B (exact cosine top-k). TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import this; the product never does.

PARITY UNPINNED IN THE REFERENCE: no search routine, test or golden vector exists there
(SURVEY.md F3, F4, §8c). The oracle generalises the reference's one cosine,
`TextChunker._cosine_similarity` (3-chunks/pipeline/src/processors/text_processor.py:1601-1605:
`np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b))`), to `[Q,D] x [N,D]^T`: on the unit
rows produced with `normalize_embeddings=True` (generate_embeddings_parallel.py:149) the
denominator is 1, so the score is the fp32 dot product of the STORED values (bf16 corpora are
upcast, so they are judged on their bf16-rounded values). Top-k order: score descending, ties by
ascending row id. It is pinned by (a) `cosine_pairwise` == the reference formula, (b) agreement
with `torch.topk`, (c) committed golden fixtures (tests/golden/search_*.npz).
"""
from __future__ import annotations

import numpy as np


def cosine_pairwise(a: np.ndarray, b: np.ndarray) -> float:
    """The reference formula verbatim in meaning (text_processor.py:1605)."""
    return float(np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b)))


def to_f32(x) -> np.ndarray:
    """Upcast stored values (numpy fp32/fp64, torch fp32/bf16) to float32 without re-rounding."""
    if hasattr(x, "detach"):
        x = x.detach().cpu().float().numpy()
    return np.ascontiguousarray(x, dtype=np.float32)


def scores_f32(queries, corpus) -> np.ndarray:
    return to_f32(queries) @ to_f32(corpus).T


def topk_from_scores(scores: np.ndarray, k: int, id_offset: int = 0):
    """Row-wise top-k ordered by (score desc, id asc); slots beyond N hold (-inf, -1)."""
    Q, N = scores.shape
    kk = min(k, N)
    out_s = np.full((Q, k), -np.inf, np.float32)
    out_i = np.full((Q, k), -1, np.int64)
    if kk == 0:
        return out_s, out_i
    ids = np.arange(N)
    for q in range(Q):
        row = scores[q]
        if kk < N:
            kth = np.partition(row, N - kk)[N - kk]
            cand = ids[row >= kth]  # keeps every tie at the boundary
        else:
            cand = ids
        order = np.lexsort((cand, -row[cand]))[:kk]  # primary -score, secondary id
        sel = cand[order]
        out_s[q, :kk] = row[sel]
        out_i[q, :kk] = sel + id_offset
    return out_s, out_i


def oracle_search(queries, corpus, k: int, id_offset: int = 0, block: int = 4096):
    """fp32 `Q @ C.T` + top-k, blocked over queries to bound memory."""
    q32, c32 = to_f32(queries), to_f32(corpus)
    outs_s, outs_i = [], []
    for i in range(0, q32.shape[0], block):
        s = q32[i:i + block] @ c32.T
        a, b = topk_from_scores(s, k, id_offset)
        outs_s.append(a)
        outs_i.append(b)
    if not outs_s:
        return np.zeros((0, k), np.float32), np.zeros((0, k), np.int64)
    return np.concatenate(outs_s), np.concatenate(outs_i)


def merge_topk(scores: np.ndarray, ids: np.ndarray, k: int | None = None):
    """Merge [G,Q,k] lists into [Q,k] by (score desc, id asc); id -1 marks an empty slot."""
    G, Q, kk = scores.shape
    k = k or kk
    out_s = np.full((Q, k), -np.inf, np.float32)
    out_i = np.full((Q, k), -1, np.int64)
    for q in range(Q):
        s = scores[:, q, :].reshape(-1)
        i = ids[:, q, :].reshape(-1)
        valid = i >= 0
        s, i = s[valid], i[valid]
        order = np.lexsort((i, -s))[:k]
        out_s[q, :len(order)] = s[order]
        out_i[q, :len(order)] = i[order]
    return out_s, out_i


def check_topk(got_scores, got_ids, queries, corpus, k: int, tol: float = 1e-5, id_offset: int = 0,
               ref=None) -> dict:
    """North-star acceptance: ids bit-exact against the oracle except where the competing oracle
    scores lie within `tol` (1e-5); returned scores within `tol` of the oracle's.
    Returns a report dict; `ok` is the verdict."""
    got_scores = np.asarray(got_scores, np.float32)
    got_ids = np.asarray(got_ids, np.int64)
    ref_s, ref_i = ref if ref is not None else oracle_search(queries, corpus, k, id_offset)
    q32, c32 = to_f32(queries), to_f32(corpus)
    exact = int((got_ids == ref_i).sum())
    total = int(ref_i.size)
    bad = []
    tie_swaps = 0
    for q, r in zip(*np.nonzero(got_ids != ref_i)):
        gid = got_ids[q, r]
        if gid < 0 or gid - id_offset >= c32.shape[0]:
            bad.append((int(q), int(r), int(gid), int(ref_i[q, r]), "invalid id"))
            continue
        true_s = float(q32[q] @ c32[gid - id_offset])
        if abs(true_s - float(ref_s[q, r])) <= tol:
            tie_swaps += 1
        else:
            bad.append((int(q), int(r), int(gid), int(ref_i[q, r]), true_s - float(ref_s[q, r])))
    # no duplicate ids within a row
    dup = 0
    for q in range(got_ids.shape[0]):
        row = got_ids[q][got_ids[q] >= 0]
        dup += len(row) - len(np.unique(row))
    finite = np.isfinite(ref_s)
    score_err = float(np.max(np.abs(got_scores[finite] - ref_s[finite]))) if finite.any() else 0.0
    ok = (not bad) and dup == 0 and score_err <= tol
    return {"ok": ok, "exact": exact, "total": total, "tie_swaps": tie_swaps, "bad": bad[:10],
            "n_bad": len(bad), "duplicates": dup, "max_score_err": score_err}


# ---------------------------------------------------------------------------------------------
# Synthetic inputs
# ---------------------------------------------------------------------------------------------
def synthetic_unit_rows(n: int, d: int = 768, seed: int = 0, bf16: bool = False, plant_ties: bool = False):
    """rows ~ N(0, I_d), L2-normalised in fp32; optionally rounded to bf16 (returned as the fp32
    image of the bf16 values) and with 1% exact duplicates + 1% near-duplicates planted."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    if plant_ties and n >= 200:
        m = max(n // 100, 1)
        src = rng.integers(0, n, size=m)
        dst = rng.integers(0, n, size=m)
        x[dst] = x[src]  # exact duplicates -> exact score ties, resolved by ascending id
        src2 = rng.integers(0, n, size=m)
        dst2 = rng.integers(0, n, size=m)
        x[dst2] = x[src2] * (1.0 + 1e-6)  # near-duplicates: score gap ~1e-6 < tolerance
    if bf16:
        import torch

        x = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    return x
