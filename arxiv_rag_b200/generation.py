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

    Under torch.distributed (one process per GPU) every rank encodes its share of the tasks and
    the rows are gathered to every rank with `all_gather_object`; single-process runs encode all
    tasks on the current GPU. `num_workers` is accepted for signature parity: the worker count is
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
        gathered: List[Optional[dict]] = [None] * world
        dist.all_gather_object(gathered, mine)
        merged: Dict[int, List[np.ndarray]] = {}
        for g in gathered:
            merged.update(g)
        mine = merged
    embeddings = reorder(mine, len(tasks))
    if len(embeddings) != n:
        raise RuntimeError(f"embedding count {len(embeddings)} != chunk count {n}")
    return embeddings
