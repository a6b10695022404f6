"""CPU restatement of the reference's encode control flow. TEST INFRASTRUCTURE ONLY
(imported by tests/, smoke() and bench.py's cpu_baseline / --impl reference legs; never by the
product).

Follows 4-embed/generation/generate_embeddings_parallel.py:
  * `OracleSentenceTransformer` — a SentenceTransformer-shaped shim over the oracle
    (`transformers.MPNetModel` + pooling + normalise), standing where `SentenceTransformer(name)`
    does at :47. `encode` restates sentence-transformers' length-sort / batch / pad-to-longest /
    un-sort.
  * `generate_embeddings_worker` — :131-177 (sub-batch loop, `.encode` flags of :146-153).
  * `generate_embeddings_parallel` — :179-269 with the mis-indented `for` at :239 re-indented;
    task split :197-200, unordered collect :213-226, reorder by batch_idx :236-244,
    count fix-up :259-267. Tasks run in-process (one model, all torch threads);
  * `generate_embeddings_pool` — the same with the reference's own process pool: `Pool(num_workers,
    initializer=init_worker_model)` (:205), `num_workers` = 75 % of the cores (:190), `spawn` start
    (:614), one model per process (:40-65), unordered collect + reorder.
Input is pre-tokenised (ids, mask): no tokenizer vocabulary exists offline.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import encode_oracle


class OracleSentenceTransformer:
    def __init__(self, arch, state_dict: dict):
        self.arch = arch
        self.model = encode_oracle.reference_model(arch, state_dict)

    def get_sentence_embedding_dimension(self) -> int:
        return self.arch.hidden_size

    @torch.no_grad()
    def encode(self, sentences, batch_size: int = 32, normalize_embeddings: bool = False,
               show_progress_bar: bool = False, convert_to_numpy: bool = True,
               convert_to_tensor: bool = False):
        ids, mask = (np.asarray(sentences[0]), np.asarray(sentences[1]))
        n = ids.shape[0]
        lengths = mask.sum(1)
        order = np.argsort(-lengths, kind="stable")
        out = np.zeros((n, self.arch.hidden_size), np.float32)
        for s in range(0, n, batch_size):
            sel = order[s:s + batch_size]
            S = max(int(lengths[sel].max()), 1)
            tok = self.model(input_ids=torch.from_numpy(ids[sel, :S]).long(),
                             attention_mask=torch.from_numpy(mask[sel, :S]).long())[0]
            e = encode_oracle.pool_normalize(tok, torch.from_numpy(mask[sel, :S]).long())
            if normalize_embeddings:
                e = torch.nn.functional.normalize(e, p=2, dim=1)
            out[sel] = e.float().numpy()
        return out


def generate_embeddings_worker(args, model) -> Tuple[int, List[np.ndarray], Optional[str]]:
    """Reference :131-177 on pre-tokenised input."""
    (ids, mask), model_name, batch_size, batch_idx = args
    embeddings: List[np.ndarray] = []
    for i in range(0, ids.shape[0], batch_size):
        b_ids, b_mask = ids[i:i + batch_size], mask[i:i + batch_size]
        batch_embeddings = model.encode((b_ids, b_mask), batch_size=min(batch_size, b_ids.shape[0]),
                                        normalize_embeddings=True, show_progress_bar=False,
                                        convert_to_numpy=True, convert_to_tensor=False)
        embeddings.extend(batch_embeddings)
    return (batch_idx, embeddings, None)


def generate_embeddings_parallel(ids: np.ndarray, mask: np.ndarray, model, model_name: str = "all-mpnet-base-v2",
                                 batch_size: int = 200, chunks_per_worker: int = 500) -> List[np.ndarray]:
    """Reference :179-269 (line 239 re-indented), tasks executed in-process."""
    n = ids.shape[0]
    text_batches = []
    for i in range(0, n, chunks_per_worker):
        text_batches.append(((ids[i:i + chunks_per_worker], mask[i:i + chunks_per_worker]), model_name,
                             batch_size, len(text_batches)))
    embeddings_dict: Dict[int, List[np.ndarray]] = {}
    for task in text_batches:
        batch_idx, batch_embeddings, error = generate_embeddings_worker(task, model)
        if batch_embeddings:
            embeddings_dict[batch_idx] = batch_embeddings
    embeddings: List[np.ndarray] = []
    for i in range(len(text_batches)):
        batch_embeds = embeddings_dict.get(i)
        if batch_embeds is not None:
            embeddings.extend(batch_embeds)
    if len(embeddings) != n:
        if len(embeddings) < n:
            dim = len(embeddings[0]) if embeddings else 768
            embeddings.extend([np.zeros(dim)] * (n - len(embeddings)))
        else:
            embeddings = embeddings[:n]
    return embeddings


# ---------------------------------------------------------------------------------------------
# The reference's process-pool variant (:190, :205, :213-226, :614)
# ---------------------------------------------------------------------------------------------
_pool_model = None


def _pool_init(arch, weight_seed: int):
    """init_worker_model (:40-65): one CPU model per worker process."""
    global _pool_model
    from arxiv_rag_b200.weights import synthetic_state_dict

    _pool_model = OracleSentenceTransformer(arch, synthetic_state_dict(arch, weight_seed))


def _pool_task(args):
    return generate_embeddings_worker(args, _pool_model)


def default_pool_workers() -> int:
    """`max(1, int(cpu_count * 0.75))` (:190)."""
    import os

    return max(1, int((os.cpu_count() or 1) * 0.75))


class ReferencePool:
    """The reference's worker pool, kept open across calls so that a benchmark can time the
    encode work without the per-run model construction."""

    def __init__(self, arch, weight_seed: int = 0, num_workers: Optional[int] = None):
        import multiprocessing as mp

        self.num_workers = num_workers or default_pool_workers()
        self.pool = mp.get_context("spawn").Pool(self.num_workers, initializer=_pool_init, initargs=(arch, weight_seed))

    def generate_embeddings_parallel(self, ids: np.ndarray, mask: np.ndarray, model_name: str = "all-mpnet-base-v2",
                                     batch_size: int = 200, chunks_per_worker: int = 500) -> List[np.ndarray]:
        n = ids.shape[0]
        tasks = []
        for i in range(0, n, chunks_per_worker):
            tasks.append(((ids[i:i + chunks_per_worker], mask[i:i + chunks_per_worker]), model_name, batch_size, len(tasks)))
        results: Dict[int, List[np.ndarray]] = {}
        for batch_idx, rows, err in self.pool.imap_unordered(_pool_task, tasks):  # :213-226
            if rows:
                results[batch_idx] = rows
        out: List[np.ndarray] = []
        for i in range(len(tasks)):  # :236-244
            if i in results:
                out.extend(results[i])
        return out

    def close(self):
        self.pool.close()
        self.pool.join()
