"""CPU: pin the search oracle — reference cosine formula, torch.topk agreement, tie rule,
golden fixtures, merge, and the acceptance checker itself."""
import os

import numpy as np
import torch

from oracle import search_oracle as so
from tests.conftest import GOLDEN


def test_cosine_is_the_reference_formula():
    """text_processor.py:1605: np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b)); on unit
    rows it equals the plain dot product the search uses."""
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal(768), rng.standard_normal(768)
    assert abs(so.cosine_pairwise(a, b) - np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b))) < 1e-15
    rows = so.synthetic_unit_rows(4, 768, seed=3)
    s = so.scores_f32(rows[:2], rows[2:])
    for i in range(2):
        for j in range(2):
            assert abs(s[i, j] - so.cosine_pairwise(rows[i], rows[2 + j])) < 1e-6


def test_oracle_matches_torch_topk():
    c = so.synthetic_unit_rows(3000, 64, seed=0)
    q = so.synthetic_unit_rows(17, 64, seed=1)
    s, i = so.oracle_search(q, c, 10)
    ts, ti = torch.topk(torch.from_numpy(q) @ torch.from_numpy(c).T, 10, dim=1)
    assert np.allclose(s, ts.numpy(), atol=1e-6)
    assert (i == ti.numpy()).all()  # no ties in this draw
    assert (np.diff(s, axis=1) <= 0).all()


def test_tie_rule_ascending_id():
    c = np.zeros((6, 4), np.float32)
    c[:, 0] = 1.0  # six identical rows -> identical scores
    c[3, 1] = 0.0
    q = np.array([[1, 0, 0, 0]], np.float32)
    s, i = so.oracle_search(q, c, 4)
    assert i.tolist() == [[0, 1, 2, 3]]
    s, i = so.oracle_search(q, c, 4, id_offset=100)
    assert i.tolist() == [[100, 101, 102, 103]]


def test_k_larger_than_n_pads():
    c = so.synthetic_unit_rows(3, 8, seed=0)
    q = so.synthetic_unit_rows(2, 8, seed=1)
    s, i = so.oracle_search(q, c, 5)
    assert (i[:, 3:] == -1).all() and np.isinf(s[:, 3:]).all()
    assert sorted(i[0, :3].tolist()) == [0, 1, 2]


def test_golden_small_with_data():
    fx = np.load(os.path.join(GOLDEN, "search_64x32_k5.npz"))
    s, i = so.oracle_search(fx["queries"], fx["corpus"], int(fx["k"]))
    assert (i == fx["ids"]).all()
    assert np.allclose(s, fx["scores"], atol=1e-7)


def test_golden_from_seed():
    for name in ("search_2000x768_k10_bf16.npz", "search_2000x768_k10_f32.npz"):
        fx = np.load(os.path.join(GOLDEN, name))
        c = so.synthetic_unit_rows(int(fx["N"]), int(fx["D"]), seed=int(fx["corpus_seed"]), bf16=bool(fx["bf16"]), plant_ties=True)
        q = so.synthetic_unit_rows(int(fx["Q"]), int(fx["D"]), seed=int(fx["query_seed"]), bf16=bool(fx["bf16"]))
        rep = so.check_topk(fx["scores"], fx["ids"], q, c, int(fx["k"]))
        assert rep["ok"], rep


def test_shard_merge_equals_unsharded():
    c = so.synthetic_unit_rows(1000, 32, seed=0, plant_ties=True)
    q = so.synthetic_unit_rows(9, 32, seed=1)
    full_s, full_i = so.oracle_search(q, c, 7)
    parts_s, parts_i = [], []
    bounds = [0, 130, 131, 600, 1000]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        s, i = so.oracle_search(q, c[lo:hi], 7, id_offset=lo)
        parts_s.append(s)
        parts_i.append(i)
    ms, mi = so.merge_topk(np.stack(parts_s), np.stack(parts_i))
    assert (mi == full_i).all() and np.allclose(ms, full_s, atol=1e-6)  # BLAS blocking differs by shape


def test_checker_flags_wrong_results():
    c = so.synthetic_unit_rows(500, 32, seed=0)
    q = so.synthetic_unit_rows(4, 32, seed=1)
    s, i = so.oracle_search(q, c, 5)
    assert so.check_topk(s, i, q, c, 5)["ok"]
    bad = i.copy()
    bad[0, 0] = (bad[0, 0] + 7) % 500
    assert not so.check_topk(s, bad, q, c, 5)["ok"]
    dup = i.copy()
    dup[1, 1] = dup[1, 0]
    assert not so.check_topk(s, dup, q, c, 5)["ok"]
    # a swap of two entries whose scores differ by < 1e-5 is accepted
    c2 = c.copy()
    c2[11] = c2[10] * (1 + 1e-6)
    s2, i2 = so.oracle_search(c2[10:11], c2, 3)
    sw = i2.copy()
    sw[0, [0, 1]] = sw[0, [1, 0]]
    rep = so.check_topk(s2, sw, c2[10:11], c2, 3)
    assert rep["ok"] and rep["tie_swaps"] == 2
