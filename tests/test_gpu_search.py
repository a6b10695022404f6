"""GPU: stage B parity. Fused score+top-k search (through the C ABI) against the oracle on the
same seeded inputs, the committed golden fixtures, and size-independent properties at sizes the
oracle cannot brute-force quickly.

Tolerance (north_star): ids bit-exact except for ties whose oracle scores lie within 1e-5;
scores within 1e-5 (`oracle.search_oracle.check_topk`)."""
import os

import numpy as np
import pytest
import torch

from arxiv_rag_b200 import search as S
from oracle import search_oracle as so
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _run(q, c, k, bf16, id_offset=0):
    dt = torch.bfloat16 if bf16 else torch.float32
    idx = S.CorpusIndex(torch.from_numpy(c).to(dt), id_offset=id_offset)
    s, i = idx.search(torch.from_numpy(q).to(dt), k)
    return s.cpu().numpy(), i.cpu().numpy()


@pytest.mark.parametrize("name", ["search_2000x768_k10_bf16.npz", "search_2000x768_k10_f32.npz"])
def test_golden_from_seed(cuda, name):
    fx = np.load(os.path.join(GOLDEN, name))
    bf16 = bool(fx["bf16"])
    c = so.synthetic_unit_rows(int(fx["N"]), int(fx["D"]), seed=int(fx["corpus_seed"]), bf16=bf16, plant_ties=True)
    q = so.synthetic_unit_rows(int(fx["Q"]), int(fx["D"]), seed=int(fx["query_seed"]), bf16=bf16)
    s, i = _run(q, c, int(fx["k"]), bf16)
    rep = so.check_topk(s, i, q, c, int(fx["k"]), tol=TOL, ref=(fx["scores"], fx["ids"]))
    assert rep["ok"], rep


def test_golden_small_d32(cuda):
    fx = np.load(os.path.join(GOLDEN, "search_64x32_k5.npz"))
    s, i = _run(fx["queries"], fx["corpus"], 5, bf16=False)
    rep = so.check_topk(s, i, fx["queries"], fx["corpus"], 5, tol=TOL, ref=(fx["scores"], fx["ids"]))
    assert rep["ok"], rep


@pytest.mark.parametrize("bf16", [True, False])
@pytest.mark.parametrize("shape", [(5, 300, 10), (130, 5000, 10), (257, 20000, 100), (3, 7, 10), (1, 1, 1),
                                   (1000, 10000, 10), (64, 3000, 128), (70, 4000, 17), (200, 30000, 64),
                                   (129, 9000, 65), (40, 50000, 100), (16, 2500, 127)])
def test_oracle_parity(cuda, bf16, shape):
    """Ragged Q (not a multiple of 128), N not a multiple of 256, k > N, planted exact duplicates
    (tie rule: ascending id) and near-duplicates (inside the 1e-5 tolerance)."""
    Q, N, k = shape
    c = so.synthetic_unit_rows(N, 768, seed=0, bf16=bf16, plant_ties=True)
    q = so.synthetic_unit_rows(Q, 768, seed=1, bf16=bf16)
    s, i = _run(q, c, k, bf16, id_offset=12345)
    rep = so.check_topk(s, i, q, c, k, tol=TOL, id_offset=12345)
    assert rep["ok"], rep
    kk = min(k, N)
    assert (np.diff(s[:, :kk], axis=1) <= 0).all()  # sorted descending
    if k > N:
        assert (i[:, N:] == -1).all() and np.isinf(s[:, N:]).all()


def test_exact_duplicates_order_by_id(cuda):
    c = so.synthetic_unit_rows(600, 768, seed=4, bf16=True)
    c[[17, 300, 301, 599]] = c[5]
    q = c[5:6].copy()
    s, i = _run(q, c, 6, bf16=True)
    assert i[0, :5].tolist() == [5, 17, 300, 301, 599]
    assert np.ptp(s[0, :5]) == 0.0  # identical rows -> bit-identical scores


@pytest.mark.parametrize("k", [24, 40, 100])
def test_many_duplicates_buffered_lists(cuda, k):
    """k > 16 keeps candidates in a per-row buffer that is merged into the sorted list in
    batches: blocks of identical rows (more of them than k, scattered over several 256-row
    chunks) must still come back in ascending-id order, and the cut at k must keep the lowest ids."""
    rng = np.random.default_rng(k)
    c = so.synthetic_unit_rows(3000, 768, seed=6, bf16=True)
    dup = np.sort(rng.choice(3000, size=k + 37, replace=False))
    c[dup] = c[dup[0]]
    q = np.concatenate([c[dup[0]:dup[0] + 1], so.synthetic_unit_rows(4, 768, seed=7, bf16=True)])
    s, i = _run(q, c, k, bf16=True)
    assert i[0].tolist() == dup[:k].tolist()
    assert np.ptp(s[0]) == 0.0
    rep = so.check_topk(s, i, q, c, k, tol=TOL)
    assert rep["ok"], rep


def test_cfg1_reference_case(cuda):
    """BASELINE configs[0] search half: 1k queries over 10k rows, top-10, fp32."""
    c = so.synthetic_unit_rows(10_000, 768, seed=0)
    q = so.synthetic_unit_rows(1_000, 768, seed=1)
    s, i = _run(q, c, 10, bf16=False)
    rep = so.check_topk(s, i, q, c, 10, tol=TOL)
    assert rep["ok"] and rep["exact"] >= rep["total"] - 20, rep


def test_merge_kernel_matches_oracle(cuda):
    rng = np.random.default_rng(0)
    G, Q, k = 5, 37, 9
    sc = np.sort(rng.standard_normal((G, Q, k)).astype(np.float32), axis=2)[:, :, ::-1].copy()
    sc[:, :, 3] = sc[:, :, 2]  # ties inside and across lists
    sc[1] = sc[0]
    ids = np.stack([np.sort(rng.choice(10_000, size=(Q, k), replace=False), axis=1) + g * 10_000 for g in range(G)])
    sc[2, :, 6:] = -np.inf
    ids[2, :, 6:] = -1  # a short shard
    ms, mi = S.merge_topk(torch.from_numpy(sc).cuda(), torch.from_numpy(ids).cuda())
    rs, ri = so.merge_topk(sc, ids)
    assert np.array_equal(mi.cpu().numpy(), ri) and np.array_equal(ms.cpu().numpy(), rs)


# ---------------------------------------------------------------- properties at larger sizes
def test_shard_invariance_large(cuda):
    """top-k over G row shards (with id offsets) + merge == top-k over the unsharded matrix
    (SURVEY.md §8e), at a size where every CTA handles several splits."""
    N, Q, k = 300_000, 700, 10
    g = torch.Generator(device="cuda").manual_seed(0)
    c = torch.nn.functional.normalize(torch.randn(N, 768, device=cuda, generator=g), dim=1).to(torch.bfloat16)
    q = torch.nn.functional.normalize(torch.randn(Q, 768, device=cuda, generator=g), dim=1).to(torch.bfloat16)
    fs, fi = S.CorpusIndex(c).search(q, k)
    parts = []
    for r in range(4):
        lo, hi = S.shard_bounds(N, 4, r)
        parts.append(S.CorpusIndex(c[lo:hi], id_offset=lo).search(q, k))
    ms, mi = S.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    assert torch.equal(mi, fi) and torch.equal(ms, fs)
    # the returned scores are the true fp32 dot products of the returned rows
    chk = (q[:64].float()[:, None, :] * c[fi[:64]].float()).sum(-1)
    assert (chk - fs[:64]).abs().max().item() < TOL
    # spot-check against the oracle on a query subset
    rep = so.check_topk(fs[:16].cpu().numpy(), fi[:16].cpu().numpy(), q[:16].float().cpu().numpy(), c.float().cpu().numpy(), k)
    assert rep["ok"], rep


def test_row_permutation_property(cuda):
    """Permuting corpus rows permutes the returned ids and leaves the scores unchanged."""
    N, Q, k = 50_000, 200, 10
    g = torch.Generator(device="cuda").manual_seed(1)
    c = torch.nn.functional.normalize(torch.randn(N, 768, device=cuda, generator=g), dim=1).to(torch.bfloat16)
    q = torch.nn.functional.normalize(torch.randn(Q, 768, device=cuda, generator=g), dim=1).to(torch.bfloat16)
    s0, i0 = S.CorpusIndex(c).search(q, k)
    perm = torch.randperm(N, device=cuda, generator=g)
    s1, i1 = S.CorpusIndex(c[perm]).search(q, k)
    assert torch.equal(s0, s1)
    assert torch.equal(perm[i1], i0)  # no exact ties in a continuous random draw


def test_argument_errors(cuda):
    from arxiv_rag_b200._lib import ArbError

    idx = S.CorpusIndex(torch.zeros(10, 768, dtype=torch.bfloat16))
    with pytest.raises(ArbError):
        idx.search(torch.zeros(2, 768, dtype=torch.bfloat16), k=1000)  # k > kMaxK
    with pytest.raises(ValueError):
        idx.search(torch.zeros(2, 64, dtype=torch.bfloat16), k=5)
    s, i = idx.search(torch.zeros(0, 768, dtype=torch.bfloat16), k=5)
    assert tuple(s.shape) == (0, 5)
