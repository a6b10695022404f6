"""GPU: stage B parity. Fused score+top-k search (through the C ABI) against the oracle on the
same seeded inputs, the committed golden fixtures, and size-independent properties at sizes the
oracle cannot brute-force quickly.

Tolerance (north_star): ids bit-exact except for ties whose oracle scores lie within 1e-5;
scores within 1e-5 (`oracle.search_oracle.check_topk`)."""
import os

import numpy as np
import pytest
import torch

from arxiv_rag_b200 import _lib
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


@pytest.mark.parametrize("bf16", [True, False])
@pytest.mark.parametrize("shape", [(129, 5000, 10), (300, 20000, 16), (257, 20000, 100), (1000, 30000, 10), (640, 9000, 64),
                                   (130, 3000, 128), (5, 300, 10), (200, 50, 10), (300, 129, 5), (257, 7, 10)])
def test_pair_schedule_equals_single_cta(cuda, bf16, shape):
    """The CTA-pair schedule (tcgen05 cta_group::2, two query tiles per corpus chunk; auto for
    Q > 128) and the single-CTA schedule must return identical ids and scores — odd tile counts
    (the pair's second CTA has no queries), every list type, ragged N — and meet the oracle."""
    from arxiv_rag_b200 import _lib

    Q, N, k = shape
    c = so.synthetic_unit_rows(N, 768, seed=2, bf16=bf16, plant_ties=True)
    q = so.synthetic_unit_rows(Q, 768, seed=3, bf16=bf16)
    out = []
    try:
        for mode in (1, 2):
            _lib.check(_lib.lib().arb_set_search_mode(mode))
            out.append(_run(q, c, k, bf16, id_offset=7))
    finally:
        _lib.check(_lib.lib().arb_set_search_mode(0))
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][0], out[1][0])
    rep = so.check_topk(out[1][0], out[1][1], q, c, k, tol=TOL, id_offset=7)
    assert rep["ok"], rep


@pytest.mark.parametrize("shape", [(4096, 120_000, 10), (700, 90_000, 10), (2100, 60_000, 100)])
def test_split_pacing_does_not_change_results(cuda, shape):
    """Work units that walk the same corpus split keep within a few chunks of each other through
    progress slots in the workspace (arb_set_search_pace). That only changes WHEN a unit loads a chunk:
    ids and scores are identical with the pacing off, in both schedules, call after call on the same
    (uninitialised, then stale) workspace."""
    from arxiv_rag_b200 import _lib

    Q, N, k = shape
    c = so.synthetic_unit_rows(N, 768, seed=12, bf16=True, plant_ties=True)
    q = so.synthetic_unit_rows(Q, 768, seed=13, bf16=True)
    lib = _lib.lib()
    out = {}
    try:
        for pace in (0, 1):
            for mode in (1, 2):
                _lib.check(lib.arb_set_search_pace(pace))
                _lib.check(lib.arb_set_search_mode(mode))
                for rep in range(2):
                    out[(pace, mode, rep)] = _run(q, c, k, True, id_offset=3)
    finally:
        _lib.check(lib.arb_set_search_pace(1))
        _lib.check(lib.arb_set_search_mode(0))
    first = out[(0, 1, 0)]
    for key, o in out.items():
        assert np.array_equal(o[1], first[1]) and np.array_equal(o[0], first[0]), key
    rep = so.check_topk(first[0][:256], first[1][:256], q[:256], c, k, tol=TOL, id_offset=3)  # and they are right
    assert rep["ok"], rep


def test_cfg1_reference_case(cuda):
    """BASELINE configs[0] search half: 1k queries over 10k rows, top-10, fp32."""
    c = so.synthetic_unit_rows(10_000, 768, seed=0)
    q = so.synthetic_unit_rows(1_000, 768, seed=1)
    s, i = _run(q, c, 10, bf16=False)
    rep = so.check_topk(s, i, q, c, 10, tol=TOL)
    assert rep["ok"] and rep["exact"] >= rep["total"] - 20, rep


@pytest.mark.parametrize("shape", [(5, 37, 9), (37, 20, 100), (144, 64, 10), (1, 5, 4), (2, 1, 1), (7, 300, 33),
                                   (300, 3, 128)])
def test_merge_kernel_matches_oracle(cuda, shape):
    """k-way merge of sorted lists: the shared-memory tree merge, and (last shape: G*k entries do
    not fit in shared memory) the one-warp-per-query fallback. Ties inside and across lists, a
    short list, an empty list."""
    rng = np.random.default_rng(0)
    G, Q, k = shape
    sc = np.sort(rng.standard_normal((G, Q, k)).astype(np.float32), axis=2)[:, :, ::-1].copy()
    if k > 3:
        sc[:, :, 3] = sc[:, :, 2]  # ties inside a list
    if G > 1:
        sc[1] = sc[0]  # and across lists
    ids = np.stack([np.sort(rng.choice(10_000, size=(Q, k), replace=False), axis=1) + g * 10_000 for g in range(G)])
    if G > 2 and k > 6:
        sc[2, :, 6:] = -np.inf
        ids[2, :, 6:] = -1  # a short shard
    if G > 4:
        sc[4] = -np.inf
        ids[4] = -1  # an empty shard
    ms, mi = S.merge_topk(torch.from_numpy(sc).cuda(), torch.from_numpy(ids).cuda())
    rs, ri = so.merge_topk(sc, ids)
    assert np.array_equal(mi.cpu().numpy(), ri) and np.array_equal(ms.cpu().numpy(), rs)


def test_merge_skewed_lists_take_the_tree_path(cuda):
    """One weak list drags the pruning bound down so that (almost) every entry of the other lists
    is a candidate: more than the pruned path accepts, so the pairwise tree merge must finish."""
    rng = np.random.default_rng(1)
    G, Q, k = 64, 6, 128
    sc = np.sort(rng.uniform(1.0, 2.0, size=(G, Q, k)).astype(np.float32), axis=2)[:, :, ::-1].copy()
    sc[9] = np.sort(rng.uniform(0.1, 0.5, size=(Q, k)).astype(np.float32), axis=1)[:, ::-1]
    ids = np.stack([np.sort(rng.choice(100_000, size=(Q, k), replace=False), axis=1) + g * 100_000 for g in range(G)])
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


@pytest.mark.parametrize("Qk", [(700, 10), (33, 7), (64, 100)])
def test_record_merge_equals_unsharded(cuda, Qk):
    """The multi-GPU exchange format: every rank's result written as one record (scores then ids in
    a single buffer), G records laid end to end as one all-gather leaves them, merged by
    arb_topk_merge_records == the unsharded search. Odd Q*k exercises the 8-byte id alignment."""
    from arxiv_rag_b200 import _lib

    Q, k = Qk
    N, G = 120_000, 5
    g = torch.Generator(device="cuda").manual_seed(3)
    c = torch.nn.functional.normalize(torch.randn(N, 768, device=cuda, generator=g), dim=1).to(torch.bfloat16)
    c[N // 2 + 5] = c[7]  # an exact tie across two shards
    q = torch.nn.functional.normalize(torch.randn(Q, 768, device=cuda, generator=g), dim=1).to(torch.bfloat16)
    q[0] = c[7]
    fs, fi = S.CorpusIndex(c).search(q, k)
    rec, off = int(_lib.lib().arb_topk_record_bytes(Q, k)), int(_lib.lib().arb_topk_record_ids_offset(Q, k))
    assert off % 8 == 0 and off >= Q * k * 4 and rec == off + Q * k * 8
    gathered = torch.zeros(G * rec, dtype=torch.uint8, device=cuda)
    for r in range(G):
        lo, hi = S.shard_bounds(N, G, r)
        ls = gathered[r * rec:r * rec + Q * k * 4].view(torch.float32).view(Q, k)
        li = gathered[r * rec + off:r * rec + off + Q * k * 8].view(torch.int64).view(Q, k)
        S.CorpusIndex(c[lo:hi], id_offset=lo).search(q, k, out_scores=ls, out_ids=li)
    ms = torch.empty((Q, k), dtype=torch.float32, device=cuda)
    mi = torch.empty((Q, k), dtype=torch.int64, device=cuda)
    _lib.check(_lib.lib().arb_topk_merge_records(_lib.ptr(gathered), G, Q, k, _lib.ptr(ms), _lib.ptr(mi), _lib.current_stream()))
    assert torch.equal(mi, fi) and torch.equal(ms, fs)
    assert fi[0, :2].tolist() == [7, N // 2 + 5]


def test_sharded_index_graph_replay(cuda):
    """ShardedCorpusIndex.search_graphed (world 1 here; the collective joins the same graph when
    world > 1) replays bit-identical results, also for new queries of the captured shape."""
    import torch.distributed as dist

    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29577", rank=0, world_size=1)
    g = torch.Generator(device="cuda").manual_seed(5)
    c = torch.nn.functional.normalize(torch.randn(40_000, 768, device=cuda, generator=g), dim=1).to(torch.bfloat16)
    idx = S.ShardedCorpusIndex(c, 40_000)
    for seed in (1, 2):
        q = torch.nn.functional.normalize(torch.randn(16, 768, device=cuda, generator=g), dim=1).to(torch.bfloat16)
        es, ei = idx.search(q, 10)
        gs_, gi_ = idx.search_graphed(q, 10)
        assert torch.equal(es, gs_) and torch.equal(ei, gi_)
    assert len(idx._graphs) == 1
    # A captured graph owns its workspace: an eager call with a much larger batch (which makes the
    # shared workspace grow and the allocator recycle the old block) must not disturb a later replay.
    q_small = torch.nn.functional.normalize(torch.randn(16, 768, device=cuda, generator=g), dim=1).to(torch.bfloat16)
    want_s, want_i = (t.clone() for t in idx.index.search(q_small, 10))
    q_big = torch.nn.functional.normalize(torch.randn(5000, 768, device=cuda, generator=g), dim=1).to(torch.bfloat16)
    idx.search(q_big, 100)
    junk = torch.full((64 << 20,), 0x7F, dtype=torch.uint8, device=cuda)  # reuse freed blocks with garbage
    gs_, gi_ = idx.search_graphed(q_small, 10)
    assert torch.equal(want_s, gs_) and torch.equal(want_i, gi_)
    del junk
    dist.destroy_process_group()


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


# ---------------------------------------------------------------- BASELINE full size (configs[1]/[3] corpus): properties
def test_full_size_5m_corpus_properties(cuda):
    """5 M x 768 bf16 (the BASELINE search corpus, 7.7 GB): the oracle cannot brute-force it in test
    time, so check size-independent properties — (a) 8 row shards + record merge == unsharded,
    (b) planted rows are found at rank 0 with score 1 and an exact duplicate pair comes back in
    ascending-id order, (c) returned scores are the fp32 dot products of the returned rows,
    (d) both schedules (single CTA / CTA pair) agree bit for bit."""
    from arxiv_rag_b200 import _lib

    N, Q, k, G = 5_000_000, 300, 10, 8
    g = torch.Generator(device="cuda").manual_seed(11)
    c = torch.empty((N, 768), device=cuda, dtype=torch.bfloat16)
    for s0 in range(0, N, 500_000):
        c[s0:s0 + 500_000] = torch.nn.functional.normalize(torch.randn(500_000, 768, device=cuda, generator=g), dim=1).to(torch.bfloat16)
    q = torch.nn.functional.normalize(torch.randn(Q, 768, device=cuda, generator=g), dim=1).to(torch.bfloat16)
    planted = torch.tensor([0, 1_234_567, 2_500_000, 4_999_999], device=cuda)
    q[:4] = c[planted]
    c[3_333_333] = c[1_234_567]  # exact duplicate of a planted row, in another shard
    index = S.CorpusIndex(c)
    fs, fi = index.search(q, k)
    assert fi[0, 0].item() == 0 and fi[2, 0].item() == 2_500_000 and fi[3, 0].item() == 4_999_999
    assert fi[1, :2].tolist() == [1_234_567, 3_333_333] and fs[1, 0].item() == fs[1, 1].item()
    assert (fs[:4, 0] - 1.0).abs().max().item() < 1e-2  # bf16 unit rows: |row|^2 within bf16 rounding of 1
    assert (fs[:, :-1] >= fs[:, 1:]).all()
    chk = (q[:32].float()[:, None, :] * c[fi[:32]].float()).sum(-1)
    assert (chk - fs[:32]).abs().max().item() < TOL
    # (a) shards
    rec, off = int(_lib.lib().arb_topk_record_bytes(Q, k)), int(_lib.lib().arb_topk_record_ids_offset(Q, k))
    gathered = torch.zeros(G * rec, dtype=torch.uint8, device=cuda)
    for r in range(G):
        lo, hi = S.shard_bounds(N, G, r)
        ls = gathered[r * rec:r * rec + Q * k * 4].view(torch.float32).view(Q, k)
        li = gathered[r * rec + off:r * rec + off + Q * k * 8].view(torch.int64).view(Q, k)
        S.CorpusIndex(c[lo:hi], id_offset=lo).search(q, k, out_scores=ls, out_ids=li)
    ms = torch.empty((Q, k), dtype=torch.float32, device=cuda)
    mi = torch.empty((Q, k), dtype=torch.int64, device=cuda)
    _lib.check(_lib.lib().arb_topk_merge_records(_lib.ptr(gathered), G, Q, k, _lib.ptr(ms), _lib.ptr(mi), _lib.current_stream()))
    assert torch.equal(mi, fi) and torch.equal(ms, fs)
    # (d) schedules
    try:
        _lib.check(_lib.lib().arb_set_search_mode(1))
        s1, i1 = index.search(q, k)
        assert torch.equal(i1, fi) and torch.equal(s1, fs)
    finally:
        _lib.check(_lib.lib().arb_set_search_mode(0))


def test_fp32_verdict_and_exact_fallback(cuda):
    """fp32 corpus: one tf32 pass + exact re-score carries a per-query proof of exactness; queries
    whose k-th best score does not clear the worst candidate by the tf32 error bound are re-run
    through the split-bf16 path. Planted: 60 near-copies of one query (scores within 2e-4 of each
    other, far more of them than the 32 candidates kept) — those queries cannot be verified and
    must come back exact through the fallback; ordinary queries verify."""
    rng = np.random.default_rng(3)
    N, D, k = 20_000, 768, 10
    c = so.synthetic_unit_rows(N, D, seed=0)
    q = so.synthetic_unit_rows(8, D, seed=1)
    for j in range(60):  # a dense cluster around query 0 and query 5
        for qi, base in ((0, 1000), (5, 9000)):
            v = q[qi] + 2e-3 * rng.standard_normal(D).astype(np.float32)
            c[base + 7 * j] = v / np.linalg.norm(v)
    idx = S.CorpusIndex(torch.from_numpy(c))
    s, i = idx.search(torch.from_numpy(q), k)
    rep = so.check_topk(s.cpu().numpy(), i.cpu().numpy(), q, c, k, tol=TOL)
    assert rep["ok"], rep
    assert 2 <= idx.fallback_queries <= 4, idx.fallback_queries  # the two clustered queries (a few more at most)
    # the two modes of the C entry point, side by side
    qd, cd = torch.from_numpy(q).cuda(), torch.from_numpy(c).cuda()
    lib = _lib.lib()
    outs = []
    for mode in (0, 1):
        need = lib.arb_topk_search_f32_workspace_bytes(8, N, D, k, mode)
        ws = torch.empty(need, dtype=torch.uint8, device=cuda)
        sc = torch.empty(8, k, device=cuda)
        ids = torch.empty(8, k, dtype=torch.int64, device=cuda)
        fl = torch.full((8,), -1, dtype=torch.int32, device=cuda)
        _lib.check(lib.arb_topk_search_f32(qd.data_ptr(), cd.data_ptr(), 8, N, D, k, 1.0 + 1e-6, sc.data_ptr(), ids.data_ptr(), 0,
                                           fl.data_ptr(), mode, ws.data_ptr(), ws.numel(), _lib.current_stream()))
        outs.append((sc.cpu().numpy(), ids.cpu().numpy(), fl.cpu().numpy()))
    assert outs[0][2][0] == 1 and outs[0][2][5] == 1 and outs[0][2][[1, 2, 3, 4, 6, 7]].sum() == 0  # verdicts of the tf32 pass
    assert (outs[1][2] == 0).all()
    assert so.check_topk(outs[1][0], outs[1][1], q, c, k, tol=TOL)["ok"]
    ok_rows = [1, 2, 3, 4, 6, 7]  # verified rows of the tf32 pass are exact on their own
    assert so.check_topk(outs[0][0][ok_rows], outs[0][1][ok_rows], q[ok_rows], c, k, tol=TOL)["ok"]


def test_fp32_non_unit_rows(cuda):
    """The verdict scales with |q| and the largest corpus row norm, so un-normalised rows work too."""
    rng = np.random.default_rng(5)
    c = (rng.standard_normal((5000, 256)) * rng.uniform(0.1, 7.0, (5000, 1))).astype(np.float32)
    q = (rng.standard_normal((33, 256)) * 3.0).astype(np.float32)
    idx = S.CorpusIndex(torch.from_numpy(c))
    assert 6.0 < idx.max_norm < 200.0
    s, i = idx.search(torch.from_numpy(q), 10)
    ref_s, ref_i = so.oracle_search(q, c, 10)
    assert (i.cpu().numpy() == ref_i).all()
    assert np.abs(s.cpu().numpy() - ref_s).max() < 1e-3 * np.abs(ref_s).max()


def test_gpu_collection_query(cuda, tmp_path):
    """The persisted GPU index answers Chroma-shaped queries with the exact top-k of its rows."""
    from arxiv_rag_b200 import vector_store as vs

    c = so.synthetic_unit_rows(3000, 768, seed=0, bf16=True)
    chunks = [{"chunk_id": f"c{i}", "text": f"text {i}", "metadata": {"paper_id": i // 10, "quality_score": 0.9}} for i in range(3000)]
    vs.store_in_gpu_index_batched(chunks, list(c), str(tmp_path), "papers", batch_size=700)
    col = vs.GpuCollection(str(tmp_path), "papers")
    q = so.synthetic_unit_rows(5, 768, seed=1, bf16=True)
    res = col.query(q, n_results=4)
    ref_s, ref_i = so.oracle_search(q, c, 4)
    assert res["ids"] == [[f"c{j}" for j in row] for row in ref_i]
    assert np.allclose(np.array(res["scores"]), ref_s, atol=1e-6) and np.allclose(np.array(res["distances"]), 1 - ref_s, atol=1e-6)
    assert res["documents"][2][0] == f"text {ref_i[2, 0]}" and res["metadatas"][0][1]["paper_id"] == str(ref_i[0, 1] // 10)
