"""Host-side mirror of 4-embed/generation/generate_embeddings_parallel.py for the encode path.

Same function names, argument meaning and return shapes as the reference:

* `init_worker_model` / `get_worker_model`  (:40-74)  — per-process model singleton, here a
  `B200SentenceEncoder` bound to this process's GPU instead of a CPU SentenceTransformer.
* `generate_embeddings_worker((texts, model_name, batch_size, batch_idx))` (:131-177) ->
  `(batch_idx, List[np.ndarray], error_or_None)`.
* `generate_embeddings_parallel(chunks, model_name, batch_size, num_workers, chunks_per_worker)`
  (:179-269) -> `List[np.ndarray]` in chunk order. `num_workers` is the number of GPUs.

Deliberate differences: errors raise instead of being swallowed and
zero-filled (:155-169 — there is no CPU fallback to retry on), and a failed task can therefore
never shift later rows against the metadata (the reference's latent bug at :240-265).
Data-parallel layout: tasks of `chunks_per_worker` chunks tagged with their index, rank r of G
takes tasks r, r+G, ... (no collective on the encode path); the parent re-orders by task index
exactly like :236-244.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

MODEL_NAMES = ("all-mpnet-base-v2", "all-MiniLM-L6-v2")  # the reference's --model choices (:473-475)

_worker_model = None
_worker_model_name = None
_worker_model_kwargs: dict = {}


def configure_worker_model(**kwargs) -> None:
    """Extra constructor arguments (state_dict=..., tokenizer=..., device=...) for the singleton."""
    global _worker_model_kwargs, _worker_model
    _worker_model_kwargs = dict(kwargs)
    _worker_model = None


def init_worker_model(model_name: str):
    """Initialise the model of this process (reference :40-65)."""
    global _worker_model, _worker_model_name
    if _worker_model is None or _worker_model_name != model_name:
        if model_name not in MODEL_NAMES:
            raise ValueError(f"model '{model_name}' not supported by the B200 path (have {MODEL_NAMES})")
        from .encoder import B200SentenceEncoder

        kwargs = dict(_worker_model_kwargs)
        kwargs.setdefault("model_name", model_name)
        if "arch" in kwargs:  # an explicit architecture (tests) wins over the name lookup
            kwargs.pop("model_name")
        _worker_model = B200SentenceEncoder(**kwargs)
        _worker_model_name = model_name


def get_worker_model(model_name: str):
    """Model of the current worker, created on first use (reference :67-74)."""
    if _worker_model is None or _worker_model_name != model_name:
        init_worker_model(model_name)
    return _worker_model


def load_chunks_from_file(file_path, min_quality: float = 0.8) -> List[Dict]:
    """Chunks of one `{paper_id}.json` whose `metadata.quality_score >= min_quality`
    (reference :76-92; unreadable files yield [] exactly like the reference)."""
    import json

    chunks: List[Dict] = []
    try:
        with open(file_path, "r", encoding="utf-8") as f:
            data = json.load(f)
        for chunk in data.get("chunks", []):
            if chunk.get("metadata", {}).get("quality_score", 0) >= min_quality:
                chunks.append(chunk)
    except Exception:
        pass
    return chunks


def load_chunks_parallel(output_dir, min_quality: float = 0.8, num_workers: int | None = None) -> List[Dict]:
    """All high-quality chunks under `output_dir` (reference :94-129: rglob('*.json'), skip '._*').

    Unlike the reference's `Pool.imap_unordered`, files are visited in sorted path order and the
    result order is deterministic, so row i of the saved matrix means the same chunk on every
    run. `num_workers` > 1 reads files on a thread pool (JSON decoding
    releases no GIL, but the I/O overlaps); order is preserved."""
    from concurrent.futures import ThreadPoolExecutor
    from pathlib import Path

    files = sorted(f for f in Path(output_dir).rglob("*.json") if not f.name.startswith("._"))
    if num_workers is None or num_workers <= 1:
        per_file = [load_chunks_from_file(f, min_quality) for f in files]
    else:
        with ThreadPoolExecutor(max_workers=num_workers) as ex:
            per_file = list(ex.map(lambda f: load_chunks_from_file(f, min_quality), files))
    return [c for part in per_file for c in part]


def generate_embeddings_worker(args: Tuple[Sequence, str, int, int]) -> Tuple[int, List[np.ndarray], Optional[str]]:
    """One task: encode `texts_batch` in sub-batches of `batch_size` (reference :131-177).

    `texts_batch` is a list of strings (needs a tokenizer) or a pre-tokenised
    `(input_ids[n,S], attention_mask[n,S])` pair."""
    texts_batch, model_name, batch_size, batch_idx = args
    model = get_worker_model(model_name)
    embeddings: List[np.ndarray] = []
    pretok = isinstance(texts_batch, tuple)
    n = texts_batch[0].shape[0] if pretok else len(texts_batch)
    for i in range(0, n, batch_size):
        batch = (texts_batch[0][i:i + batch_size], texts_batch[1][i:i + batch_size]) if pretok \
            else texts_batch[i:i + batch_size]
        m = batch[0].shape[0] if pretok else len(batch)
        batch_embeddings = model.encode(
            batch,
            batch_size=min(batch_size, m),
            normalize_embeddings=True,
            show_progress_bar=False,
            convert_to_numpy=True,
            convert_to_tensor=False,
        )
        embeddings.extend(batch_embeddings)  # rows: np.ndarray (768,) float32, as at :154
    return (batch_idx, embeddings, None)


def split_tasks(n_items: int, chunks_per_worker: int) -> List[Tuple[int, int, int]]:
    """(task_idx, start, stop) triples — the reference's task split at :197-200."""
    return [(t, s, min(s + chunks_per_worker, n_items))
            for t, s in enumerate(range(0, n_items, chunks_per_worker))]


def tasks_of_rank(tasks: Sequence, rank: int, world_size: int) -> List:
    """Static round-robin deal of tasks to GPUs (replaces Pool.imap_unordered, :213-226)."""
    return [t for i, t in enumerate(tasks) if i % world_size == rank]


def reorder(results: Dict[int, List[np.ndarray]], n_tasks: int) -> List[np.ndarray]:
    """Concatenate per-task rows in task order (reference :236-244); a missing task is an error."""
    out: List[np.ndarray] = []
    for i in range(n_tasks):
        if i not in results:
            raise RuntimeError(f"task {i} produced no embeddings")
        out.extend(results[i])
    return out


def generate_embeddings_parallel(chunks: List[Dict], model_name: str = "all-mpnet-base-v2",
                                 batch_size: int = 200, num_workers: int | None = None,
                                 chunks_per_worker: int = 500) -> List[np.ndarray]:
    """Encode `chunks` (dicts with 'text', or with 'input_ids'/'attention_mask' rows) and return one
    float32 row per chunk in input order (reference :179-269).

    Under torch.distributed (one process per GPU) every rank encodes its share of the tasks — no
    collective on the encode path — and, because this function must return every row like the
    reference does, the rows are then gathered as ONE fixed-shape float32 tensor per rank
    (`all_gather_into_tensor` of zero-padded `[max_rows, H]` blocks plus their global row numbers),
    not as pickled Python lists. Corpus-scale runs should not gather at all: see
    `generate_embeddings_to_disk`, where every rank writes its own shards. Single-process runs encode
    all tasks on the current GPU. `num_workers` is accepted for signature parity: the worker count is
    the number of GPUs (world size)."""
    import torch.distributed as dist

    n = len(chunks)
    if n == 0:
        return []
    pretok = "input_ids" in chunks[0]
    tasks = split_tasks(n, chunks_per_worker)
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0

    def payload(start, stop):
        part = chunks[start:stop]
        if pretok:
            S = max(len(c["input_ids"]) for c in part)
            ids = np.ones((len(part), S), np.int32)
            mask = np.zeros((len(part), S), np.int32)
            for r, c in enumerate(part):
                L = len(c["input_ids"])
                ids[r, :L] = c["input_ids"]
                mask[r, :L] = c.get("attention_mask", np.ones(L, np.int32))
            return (ids, mask)
        return [c["text"] for c in part]

    mine: Dict[int, List[np.ndarray]] = {}
    for t, start, stop in tasks_of_rank(tasks, rank, world):
        idx, rows, err = generate_embeddings_worker((payload(start, stop), model_name, batch_size, t))
        if err:
            raise RuntimeError(err)
        mine[idx] = rows
    if world > 1:
        return _gather_rows(mine, tasks, n, dist)
    embeddings = reorder(mine, len(tasks))
    if len(embeddings) != n:
        raise RuntimeError(f"embedding count {len(embeddings)} != chunk count {n}")
    return embeddings


def _gather_rows(mine: Dict[int, List[np.ndarray]], tasks, n: int, dist) -> List[np.ndarray]:
    """All ranks' rows -> the full list in chunk order on every rank, through two fixed-shape
    collectives (row numbers, rows)."""
    import torch

    world = dist.get_world_size()
    starts = {t: s for t, s, _ in tasks}
    row_ids = [starts[t] + j for t in sorted(mine) for j in range(len(mine[t]))]
    rows = [r for t in sorted(mine) for r in mine[t]]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    dim = torch.tensor([rows[0].shape[0] if rows else 0, len(rows)], dtype=torch.int64, device=dev)
    dist.all_reduce(dim, op=dist.ReduceOp.MAX)
    H, max_rows = int(dim[0]), int(dim[1])
    block = torch.zeros((max_rows, H), dtype=torch.float32, device=dev)
    ids = torch.full((max_rows,), -1, dtype=torch.int64, device=dev)
    if rows:
        block[:len(rows)] = torch.from_numpy(np.stack(rows).astype(np.float32, copy=False)).to(dev)
        ids[:len(rows)] = torch.tensor(row_ids, dtype=torch.int64, device=dev)
    all_blocks = torch.empty((world * max_rows, H), dtype=torch.float32, device=dev)
    all_ids = torch.empty((world * max_rows,), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_blocks, block)
    dist.all_gather_into_tensor(all_ids, ids)
    valid = all_ids >= 0
    if int(valid.sum()) != n:
        raise RuntimeError(f"embedding count {int(valid.sum())} != chunk count {n}")
    full = torch.empty((n, H), dtype=torch.float32, device=dev)
    full[all_ids[valid]] = all_blocks[valid]
    return list(full.cpu().numpy())


def generate_embeddings_to_disk(chunks: List[Dict], output_dir: str, model_name: str = "all-mpnet-base-v2",
                                batch_size: int = 200, shard_rows: int = 10000) -> Dict:
    """The data-parallel output path for corpus-scale runs: the saved layout of
    4-embed/utils/save_embeddings_to_disk.py:15-80 (`embeddings_batch_XXXX.npy` float64,
    `metadata_batch_XXXX.json`, `index.json`), written shard by shard by the rank that encoded it.
    Shard i holds rows [i*shard_rows, (i+1)*shard_rows); rank r of G takes shards r, r+G, ...; nothing
    is gathered, ranks only meet at a barrier before rank 0 writes `index.json`. Shards whose files
    already exist are skipped, so an interrupted run resumes (the reference keeps everything in
    RAM until the end, :257, :555). Returns the index dict."""
    import json
    import os
    from pathlib import Path

    import torch.distributed as dist

    from .storage import _as_matrix, _meta_row

    out = Path(output_dir)
    out.mkdir(parents=True, exist_ok=True)
    n = len(chunks)
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    shards = split_tasks(n, shard_rows)
    pretok = n > 0 and "input_ids" in chunks[0]
    dim = None
    for i, start, stop in tasks_of_rank(shards, rank, world):
        emb_file, meta_file = out / f"embeddings_batch_{i:04d}.npy", out / f"metadata_batch_{i:04d}.json"
        if emb_file.exists() and meta_file.exists():
            continue  # resume: this shard was completed by an earlier run
        part = chunks[start:stop]
        if pretok:
            S = max(len(c["input_ids"]) for c in part)
            ids = np.ones((len(part), S), np.int32)
            mask = np.zeros((len(part), S), np.int32)
            for r_, c in enumerate(part):
                L = len(c["input_ids"])
                ids[r_, :L] = c["input_ids"]
                mask[r_, :L] = c.get("attention_mask", np.ones(L, np.int32))
            payload = (ids, mask)
        else:
            payload = [c["text"] for c in part]
        _, rows, err = generate_embeddings_worker((payload, model_name, batch_size, i))
        if err:
            raise RuntimeError(err)
        arr = _as_matrix(rows)
        dim = int(arr.shape[1])
        metadata = []
        for j, chunk in enumerate(part):
            row = _meta_row(chunk, start + j)
            row["batch_index"] = i
            row["batch_position"] = j
            metadata.append(row)
        tmp_e, tmp_m = emb_file.with_suffix(".npy.tmp"), meta_file.with_suffix(".json.tmp")
        with open(tmp_e, "wb") as f:
            np.save(f, arr)
        with open(tmp_m, "w", encoding="utf-8") as f:
            json.dump(metadata, f, indent=2, ensure_ascii=False)
        os.replace(tmp_m, meta_file)
        os.replace(tmp_e, emb_file)  # the .npy appears last: its presence marks the shard complete
    if world > 1:
        dist.barrier()
    index = {"total_embeddings": n, "embedding_dimension": dim, "num_batches": len(shards), "batch_size": shard_rows,
             "chunks": [c.get("chunk_id") for c in chunks]}
    if rank == 0:
        if index["embedding_dimension"] is None and shards:
            index["embedding_dimension"] = int(np.load(out / "embeddings_batch_0000.npy", mmap_mode="r").shape[1])
        tmp = out / "index.json.tmp"
        with open(tmp, "w", encoding="utf-8") as f:
            json.dump(index, f, indent=2)
        os.replace(tmp, out / "index.json")
    if world > 1:
        dist.barrier()
    return index
