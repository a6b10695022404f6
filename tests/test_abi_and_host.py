"""CPU: the C-ABI library loads and exports every symbol include/*.h declares; error behaviour
without a GPU; host-side logic (saved layouts, task split/reorder, sharding, gloo world_size 2)."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from arxiv_rag_b200 import _lib, generation, storage
from arxiv_rag_b200.search import shard_bounds
from tests.conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "arxiv_rag_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(arb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.arb_abi_version() == 2


def test_struct_layout_matches_header():
    # ArbMpnetConfig: 8 x int32, float, 2 x int32; layer weights: 16 pointers; weights: 6 pointers
    assert C.sizeof(_lib.MpnetConfig) == 44
    assert C.sizeof(_lib.MpnetLayerWeights) == 16 * C.sizeof(C.c_void_p)
    assert C.sizeof(_lib.MpnetWeights) == 6 * C.sizeof(C.c_void_p)


def integration_stub(which: str) -> str:
    """The fenced python block that follows `<!-- stub:<which> -->` in INTEGRATION.md, verbatim."""
    text = open(os.path.join(ROOT, "INTEGRATION.md"), encoding="utf-8").read()
    m = re.search(r"<!-- stub:%s -->\s*```python\n(.*?)```" % which, text, flags=re.S)
    assert m, f"INTEGRATION.md has no stub block '{which}'"
    return m.group(1)


def test_integration_stub_binding(lib):
    """Execute INTEGRATION.md's ctypes binding as a maintainer would paste it: it must load the
    library and declare structs of exactly the header's sizes (the round-1 stub had lost
    `position_mode`: 40 bytes against the header's 44)."""
    ns: dict = {}
    cwd = os.getcwd()
    os.chdir(ROOT)  # the stub names the library by its path relative to the repository root
    try:
        exec(compile(integration_stub("binding"), "INTEGRATION.md:stub:binding", "exec"), ns)
    finally:
        os.chdir(cwd)
    assert C.sizeof(ns["ArbMpnetConfig"]) == C.sizeof(_lib.MpnetConfig) == 44
    assert [f[0] for f in ns["ArbMpnetConfig"]._fields_] == [f[0] for f in _lib.MpnetConfig._fields_]
    assert C.sizeof(ns["ArbMpnetLayerWeights"]) == C.sizeof(_lib.MpnetLayerWeights)
    assert C.sizeof(ns["ArbMpnetWeights"]) == C.sizeof(_lib.MpnetWeights)
    assert (ns["ARB_DTYPE_F32"], ns["ARB_DTYPE_BF16"], ns["ARB_DTYPE_F16"]) == (_lib.ARB_DTYPE_F32, _lib.ARB_DTYPE_BF16, _lib.ARB_DTYPE_F16)
    with pytest.raises(RuntimeError):
        ns["check"](ns["lib"].arb_mpnet_encode(0, 0, 0, 1, 1, 0, 0))


def test_integration_stub_tokenizer(lib):
    """INTEGRATION.md's tokenizer block, verbatim, on top of its binding block: same ids as the
    in-tree Python tokenizer (itself checked against transformers' MPNetTokenizer)."""
    from arxiv_rag_b200.tokenizer import WordPieceTokenizer

    vocab = ["<s>", "<pad>", "</s>", "[UNK]", "<mask>"] + "the quick brown fox jump ##s ##ed over lazy dog . , cafe 你 好".split()
    texts = ["The quick brown foxes jumped over the lazy dog.", "Café 你好, dogs", "", "zzz " * 500]
    ns: dict = {"vocab": vocab, "texts": texts}
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        exec(compile(integration_stub("binding"), "INTEGRATION.md:stub:binding", "exec"), ns)
        exec(compile(integration_stub("tokenizer"), "INTEGRATION.md:stub:tokenizer", "exec"), ns)
    finally:
        os.chdir(cwd)
    py = WordPieceTokenizer({t: i for i, t in enumerate(vocab)}, kind="mpnet", max_length=384)
    assert not ns["redo"].any()
    for row, n, text in zip(ns["input_ids"], ns["lengths"], texts):
        want = py.encode(text)
        assert row[:n].tolist() == want and (row[n:] == vocab.index("<pad>")).all()
    assert ns["attention_mask"].sum(1).tolist() == ns["lengths"].tolist() and int(ns["lengths"][3]) == 384


def test_argument_errors_return_codes_not_crashes(lib):
    assert lib.arb_topk_search_workspace_bytes(_lib.ARB_DTYPE_BF16, 0, 10, 768, 10) == 0
    assert lib.arb_topk_search_workspace_bytes(_lib.ARB_DTYPE_BF16, 128, 1_000_000, 768, 10) > 0
    assert lib.arb_topk_search_workspace_bytes(7, 128, 1000, 768, 10) == 0
    rc = lib.arb_topk_search(0, 0, _lib.ARB_DTYPE_BF16, 4, 4, 768, 10, 0, 0, 0, 0, 0, 0)
    assert rc == -1 and b"null" in lib.arb_last_error()
    rc = lib.arb_gemm16(0, 8, 0, 8, 0, 8, 0, 0, 0, 4, 32, 8, 0, _lib.ARB_DTYPE_BF16, 0)
    assert rc == -1
    rc = lib.arb_mpnet_encode(0, 0, 0, 1, 1, 0, 0)
    assert rc == -1
    with pytest.raises(_lib.ArbError):
        _lib.check(rc)


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from arxiv_rag_b200.encoder import B200SentenceEncoder
    from arxiv_rag_b200.search import CorpusIndex

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B200SentenceEncoder(None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CorpusIndex(np.zeros((4, 8), np.float32))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "arxiv_rag_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


# ------------------------------------------------------------------ saved layouts
def _chunks(n):
    return [{"chunk_id": f"2101.{i:05d}_chunk_{i % 3}", "text": f"text {i} é", "metadata":
             {"paper_id": f"2101.{i:05d}", "section": "intro", "quality_score": 0.9 + 0.001 * i}} for i in range(n)]


def _dir_bytes(d):
    return {f.name: f.read_bytes() for f in sorted(__import__("pathlib").Path(d).iterdir())}


def test_layouts_byte_identical_to_reference_writers(tmp_path):
    """The saved embedding/ID layout is half of the drop-in contract. tests/golden/layout_* were
    written by the reference's OWN writers (tools/make_layout_golden.py: save_embeddings_to_disk.py
    imported, generate_embeddings_parallel.py:271-321 executed from the mounted reference); every
    writer of this repo must reproduce them byte for byte — and, where the reference is mounted, the
    fixtures are re-minted and must not have drifted."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_layout_golden", os.path.join(ROOT, "tools", "make_layout_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    golden = os.path.join(ROOT, "tests", "golden")
    chunks, rows = mk.fixture_inputs()
    want_b, want_s = _dir_bytes(os.path.join(golden, "layout_batched")), _dir_bytes(os.path.join(golden, "layout_single"))
    assert set(want_b) == {"index.json"} | {f"{k}_batch_{i:04d}.{e}" for i in range(3) for k, e in (("embeddings", "npy"), ("metadata", "json"))}
    storage.save_embeddings_disk(chunks, rows, str(tmp_path / "b"), batch_size=3)
    assert _dir_bytes(tmp_path / "b") == want_b
    storage.save_embeddings_to_disk_fallback(chunks, rows, str(tmp_path / "s"))
    assert _dir_bytes(tmp_path / "s") == want_s
    w = storage.StreamingShardWriter(str(tmp_path / "w"), batch_size=3, float32_sidecar=False)
    w.append(chunks[:5], rows[:5])
    w.append(chunks[5:], rows[5:])
    w.close()
    assert _dir_bytes(tmp_path / "w") == want_b
    generation._worker_model, generation._worker_model_name = type("M", (), {"encode": lambda self, b, **kw: np.stack([rows[int(r[0])] for r in b[0]])})(), "all-mpnet-base-v2"
    try:
        tok = [dict(c, input_ids=[i]) for i, c in enumerate(chunks)]
        generation.generate_embeddings_to_disk(tok, str(tmp_path / "g"), batch_size=2, shard_rows=3)
    finally:
        generation.configure_worker_model()
    assert _dir_bytes(tmp_path / "g") == want_b
    if os.path.isdir(mk.REF):  # dev container: pin the fixtures to the reference itself
        mk.main(tmp_path / "ref")
        assert _dir_bytes(tmp_path / "ref" / "layout_batched") == want_b
        assert _dir_bytes(tmp_path / "ref" / "layout_single") == want_s


def test_single_file_layout_matches_reference(tmp_path):
    """generate_embeddings_parallel.py:271-321: float64 C-order matrix, metadata rows in order,
    index.json keys."""
    n = 7
    rows = [np.random.default_rng(i).standard_normal(768).astype(np.float32) for i in range(n)]
    storage.save_embeddings_to_disk_fallback(_chunks(n), rows, str(tmp_path))
    arr = np.load(tmp_path / "embeddings.npy")
    assert arr.dtype == np.float64 and arr.shape == (n, 768) and arr.flags.c_contiguous
    assert np.array_equal(arr.astype(np.float32), np.stack(rows))  # values stay fp32-representable
    meta = json.load(open(tmp_path / "metadata.json", encoding="utf-8"))
    assert [m["chunk_id"] for m in meta] == [c["chunk_id"] for c in _chunks(n)]
    assert set(meta[0]) == {"chunk_id", "paper_id", "section", "quality_score", "text", "text_length"}
    assert meta[3]["text_length"] == len(_chunks(n)[3]["text"])
    idx = json.load(open(tmp_path / "index.json"))
    assert set(idx) == {"total_embeddings", "embedding_dimension", "total_size_gb"}
    assert idx["total_embeddings"] == n and idx["embedding_dimension"] == 768
    emb, meta2 = storage.load_embeddings_from_disk(str(tmp_path))
    assert np.array_equal(emb, arr) and meta2 == meta


def test_batched_layout_round_trip(tmp_path):
    """save_embeddings_to_disk.py:15-117: batch files of `batch_size` rows, np.vstack on load."""
    n = 25
    rows = np.random.default_rng(0).standard_normal((n, 16)).astype(np.float32)
    storage.save_embeddings_disk(_chunks(n), list(rows), str(tmp_path), batch_size=10)
    assert sorted(p.name for p in tmp_path.glob("embeddings_batch_*.npy")) == [f"embeddings_batch_{i:04d}.npy" for i in range(3)]
    idx = json.load(open(tmp_path / "index.json"))
    assert idx["num_batches"] == 3 and idx["batch_size"] == 10 and len(idx["chunks"]) == n
    emb, meta = storage.load_embeddings_from_disk(str(tmp_path))
    assert emb.shape == (n, 16) and np.array_equal(emb.astype(np.float32), rows)
    assert meta[12]["batch_index"] == 1 and meta[12]["batch_position"] == 2
    e1, m1 = storage.load_embeddings_from_disk(str(tmp_path), batch_index=2)
    assert e1.shape == (5, 16) and len(m1) == 5
    assert storage.chunk_ids_of(meta)[24] == _chunks(n)[24]["chunk_id"]
    storage.save_search_matrix(rows, str(tmp_path))
    assert np.array_equal(np.asarray(storage.load_search_matrix(str(tmp_path))), rows)


def test_streaming_shard_writer_matches_batched_layout_and_resumes(tmp_path):
    """Shards written while encoding == what save_embeddings_disk writes at the end; an
    interrupted run leaves a loadable prefix and resumes at rows_persisted."""
    n = 27
    rows = np.random.default_rng(1).standard_normal((n, 16)).astype(np.float32)
    chunks = _chunks(n)
    ref_dir, out_dir = tmp_path / "ref", tmp_path / "stream"
    storage.save_embeddings_disk(chunks, list(rows), str(ref_dir), batch_size=10)
    w = storage.StreamingShardWriter(str(out_dir), batch_size=10)
    w.append(chunks[:13], list(rows[:13]))  # one full shard flushed, 3 rows pending ... then "crash"
    assert w.rows_persisted == 10
    emb, meta = storage.load_embeddings_from_disk(str(out_dir))
    assert emb.shape == (10, 16) and len(meta) == 10
    w2 = storage.StreamingShardWriter(str(out_dir), batch_size=10)  # resume
    assert w2.rows_persisted == 10 and w2.num_batches == 1
    w2.append(chunks[10:], list(rows[10:]))
    w2.close()
    for name in sorted(p.name for p in ref_dir.iterdir()):
        a, b = (ref_dir / name).read_bytes(), (out_dir / name).read_bytes()
        assert a == b, f"{name} differs from the reference writer's output"
    assert np.array_equal(np.asarray(storage.load_search_matrix(str(out_dir))), rows)
    with pytest.raises(ValueError):
        storage.StreamingShardWriter(str(out_dir), batch_size=10)  # ends with a partial shard


# ------------------------------------------------------------------ task split / sharding
def test_chunk_ingest_quality_filter_and_order(tmp_path):
    """load_chunks_* (reference :76-129): quality filter, '._' files skipped, broken files
    ignored, deterministic order (the reference's imap_unordered is not)."""
    for p, qs in (("b/2.json", [0.95, 0.5]), ("a/1.json", [0.9, 0.85, 0.99]), ("a/._1.json", [1.0])):
        f = tmp_path / p
        f.parent.mkdir(exist_ok=True)
        f.write_text(json.dumps({"chunks": [{"chunk_id": f"{p}:{i}", "text": "t", "metadata": {"quality_score": q}}
                                            for i, q in enumerate(qs)]}))
    (tmp_path / "a" / "broken.json").write_text("{not json")
    got = generation.load_chunks_parallel(tmp_path, min_quality=0.9)
    assert [c["chunk_id"] for c in got] == ["a/1.json:0", "a/1.json:2", "b/2.json:0"]
    assert generation.load_chunks_parallel(tmp_path, min_quality=0.9, num_workers=4) == got
    assert len(generation.load_chunks_from_file(tmp_path / "a" / "1.json")) == 3  # default 0.8


def test_task_split_and_reorder():
    tasks = generation.split_tasks(1234, 500)
    assert tasks == [(0, 0, 500), (1, 500, 1000), (2, 1000, 1234)]
    assert generation.tasks_of_rank(tasks, 1, 2) == [(1, 500, 1000)]
    res = {2: [np.full(2, 2.0)], 0: [np.full(2, 0.0), np.full(2, 0.5)], 1: [np.full(2, 1.0)]}
    out = generation.reorder(res, 3)
    assert [float(r[0]) for r in out] == [0.0, 0.5, 1.0, 2.0]
    with pytest.raises(RuntimeError):
        generation.reorder({0: [], 2: []}, 3)  # a missing task must not shift rows silently


def test_shard_bounds_cover_exactly():
    for n, g in [(10, 3), (5_000_000, 8), (7, 8), (0, 2)]:
        spans = [shard_bounds(n, g, r) for r in range(g)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))


_GLOO_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["ARB_ROOT"])
from arxiv_rag_b200.search import shard_bounds
from arxiv_rag_b200 import generation
from oracle import search_oracle as so
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# --- search protocol: shard rows, local top-k with id offsets, all_gather [G,Q,k], merge
N, Q, D, k = 1001, 13, 32, 6
c = so.synthetic_unit_rows(N, D, seed=0, plant_ties=True); q = so.synthetic_unit_rows(Q, D, seed=1)
lo, hi = shard_bounds(N, world, rank)
ls, li = so.oracle_search(q, c[lo:hi], k, id_offset=lo)
gs = [torch.empty(Q, k) for _ in range(world)]; gi = [torch.empty(Q, k, dtype=torch.int64) for _ in range(world)]
dist.all_gather(gs, torch.from_numpy(ls)); dist.all_gather(gi, torch.from_numpy(li))
ms, mi = so.merge_topk(torch.stack(gs).numpy(), torch.stack(gi).numpy())
fs, fi = so.oracle_search(q, c, k)
assert (mi == fi).all() and np.allclose(ms, fs, atol=1e-6), "sharded != unsharded"
# --- encode protocol: round-robin tasks, fixed-shape tensor gather, reorder (with a stub model)
class Stub:
    def encode(self, batch, **kw):
        ids = batch[0]
        return np.stack([np.full(4, float(r[0]), np.float32) for r in ids])
generation._worker_model, generation._worker_model_name = Stub(), "all-mpnet-base-v2"
chunks = [{"input_ids": [i, 5, 2]} for i in range(23)]
rows = generation.generate_embeddings_parallel(chunks, batch_size=4, chunks_per_worker=5)
assert len(rows) == 23 and [int(r[0]) for r in rows] == list(range(23)), "row order"
assert rows[7].dtype == np.float32 and rows[7].shape == (4,)
# --- data-parallel output path: every rank writes its own shards of the reference's batched layout
from arxiv_rag_b200 import storage
out_dir = os.environ["ARB_OUT"]
chunks = [{"chunk_id": f"c{i}", "input_ids": [i, 5, 2], "text": f"t{i}", "metadata": {"paper_id": f"p{i}", "section": "s", "quality_score": 0.95}} for i in range(23)]
idx = generation.generate_embeddings_to_disk(chunks, out_dir, batch_size=4, shard_rows=5)
assert idx["num_batches"] == 5 and idx["total_embeddings"] == 23
emb, meta = storage.load_embeddings_from_disk(out_dir)   # the reference loader's layout
assert emb.shape == (23, 4) and emb.dtype == np.float64 and [int(r[0]) for r in emb] == list(range(23))
assert [m["chunk_id"] for m in meta] == [f"c{i}" for i in range(23)] and meta[12]["batch_index"] == 2 and meta[12]["batch_position"] == 2
mine = [i for i in range(5) if i % world == rank]
dist.barrier()
if rank == 0:
    # resume: drop one shard, rerun -> only that shard is encoded again
    os.remove(os.path.join(out_dir, "embeddings_batch_0003.npy"))
dist.barrier()
calls = []
orig = generation.generate_embeddings_worker
generation.generate_embeddings_worker = lambda a: (calls.append(a[3]), orig(a))[1]
generation.generate_embeddings_to_disk(chunks, out_dir, batch_size=4, shard_rows=5)
assert calls == ([3] if 3 % world == rank else []), calls
emb2, _ = storage.load_embeddings_from_disk(out_dir)
assert np.array_equal(emb, emb2)
dist.barrier()
if rank == 0: print("GLOO_OK")
"""


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, ARB_ROOT=ROOT, MASTER_ADDR="127.0.0.1", ARB_OUT=str(tmp_path / "shards"))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert res.returncode == 0 and "GLOO_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


def test_reference_process_pool_restatement_matches_in_process():
    """oracle/refpath.ReferencePool (the reference's spawn Pool with one model per worker,
    generate_embeddings_parallel.py:190,205,213-226,614) returns what the in-process restatement
    returns, in chunk order."""
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from oracle import refpath, encode_oracle as eo\n"
        "from arxiv_rag_b200.weights import MPNetArch, synthetic_state_dict\n"
        "if __name__ == '__main__':\n"
        "    arch = MPNetArch(vocab_size=1000, num_layers=2)\n"
        "    ids, mask = eo.synthetic_tokens(12, 24, vocab_size=1000, seed=1)\n"
        "    pool = refpath.ReferencePool(arch, 0, num_workers=3)\n"
        "    a = pool.generate_embeddings_parallel(ids, mask, batch_size=2, chunks_per_worker=4)\n"
        "    pool.close()\n"
        "    b = refpath.generate_embeddings_parallel(ids, mask, refpath.OracleSentenceTransformer(arch, synthetic_state_dict(arch, 0)), batch_size=2, chunks_per_worker=4)\n"
        "    assert len(a) == len(b) == 12\n"
        "    print('MAXDIFF', float(np.abs(np.stack(a) - np.stack(b)).max()))\n"
    )
    import tempfile

    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(code)
        path = f.name
    try:
        env = dict(os.environ, OMP_NUM_THREADS="2")
        out = subprocess.run([sys.executable, path], capture_output=True, text=True, timeout=240, env=env)
        assert out.returncode == 0, out.stderr[-2000:]
        assert float(out.stdout.split("MAXDIFF")[1]) < 1e-6
    finally:
        os.unlink(path)


def test_gpu_collection_persist_and_reload(tmp_path):
    """vector_store.GpuCollection: the store `store_in_chroma_batched` (:323-468) fills, as plain
    files. Same id / document / metadata rules, shards of `shard_rows`, manifest lists only complete
    shards, bf16 rows round to nearest even, a reopened collection sees the same records."""
    import torch

    from arxiv_rag_b200 import vector_store as vs

    rng = np.random.default_rng(1)
    emb = rng.standard_normal((23, 16)).astype(np.float32)
    chunks = _chunks(23)
    del chunks[7]["chunk_id"]
    chunks[9]["metadata"].pop("section")
    col = vs.store_in_gpu_index_batched(chunks, list(emb) + [emb[0]], str(tmp_path), "papers", batch_size=5)  # one embedding too many
    assert col.count() == 23
    col2 = vs.GpuCollection(str(tmp_path), "papers")
    assert col2.count() == 23 and col2.manifest["dim"] == 16 and col2.manifest["dtype"] == "bf16"
    assert col2.ids[7] == "chunk_7" and col2.ids[8] == chunks[8]["chunk_id"]
    assert col2.documents[3] == chunks[3]["text"]
    assert col2.metadatas[9] == {"paper_id": "2101.00009", "section": "unknown", "quality_score": pytest.approx(0.909), "chunk_index": "9"}
    bits = col2.load_rows(5, 12)
    assert bits.dtype == np.uint16 and bits.shape == (7, 16)
    want = torch.from_numpy(emb[5:12]).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(bits, want)
    # small shards + append after reopen
    col3 = vs.GpuCollection(str(tmp_path), "small", dtype="float32", shard_rows=10)
    col3.add(ids=[f"a{i}" for i in range(23)], embeddings=emb)
    assert len(col3.manifest["shards"]) == 2 and col3.manifest["total"] == 20 and col3.count() == 23
    col3.persist()
    col4 = vs.GpuCollection(str(tmp_path), "small")
    assert [s["rows"] for s in col4.manifest["shards"]] == [10, 10, 3]
    assert np.array_equal(col4.load_rows(8, 22), emb[8:22])
    col4.add(ids=["z"], embeddings=emb[:1], documents=["doc"], metadatas=[{"k": 1}])
    col4.persist()
    assert vs.GpuCollection(str(tmp_path), "small").ids[-1] == "z"
    with pytest.raises(ValueError):
        col4.add(ids=["bad"], embeddings=np.zeros((1, 8), np.float32))
