"""Saved embedding / ID layouts, byte-compatible with the reference's writers and loader.

* single-file layout  — `save_embeddings_to_disk_fallback`
  (4-embed/generation/generate_embeddings_parallel.py:271-321): `embeddings.npy` float64 C-order
  `[N,D]` (an accident of `e.tolist()` at :281-284, kept for compatibility), `metadata.json`
  (list of {chunk_id, paper_id, section, quality_score, text, text_length} in row order,
  :292-306) and `index.json` ({total_embeddings, embedding_dimension, total_size_gb}, :310-318).
* batched layout — `save_embeddings_disk` / `load_embeddings_from_disk`
  (4-embed/utils/save_embeddings_to_disk.py:15-80, :82-117): `embeddings_batch_{i:04d}.npy`,
  `metadata_batch_{i:04d}.json` per `batch_size` rows, `index.json` with num_batches, batch_size
  and the chunk-id list.

Row index <-> `metadata[i]['chunk_id']` is the ID mapping the search stage returns ids into.
`save_search_matrix` / `load_search_matrix` add the float32 side file the GPU search wants
(the float64 .npy is 2x the bytes).
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import numpy as np


def _as_matrix(embeddings) -> np.ndarray:
    """Reference conversion (:278-284): list of rows -> 2-D array via python floats => float64."""
    if isinstance(embeddings, np.ndarray) and embeddings.ndim == 2:
        return embeddings.astype(np.float64)
    return np.array([np.asarray(e, dtype=np.float64) for e in embeddings], dtype=np.float64)


def _meta_row(chunk: Dict, i: int) -> Dict:
    meta = chunk.get("metadata", {})
    return {
        "chunk_id": chunk.get("chunk_id", f"chunk_{i}"),
        "paper_id": meta.get("paper_id"),
        "section": meta.get("section"),
        "quality_score": meta.get("quality_score"),
        "text": chunk["text"],
        "text_length": len(chunk["text"]),
    }


def save_embeddings_to_disk_fallback(chunks: List[Dict], embeddings, output_dir: str = "./embeddings_saved") -> None:
    """Single-file layout (reference :271-321)."""
    out = Path(output_dir)
    out.mkdir(parents=True, exist_ok=True)
    arr = _as_matrix(embeddings)
    np.save(out / "embeddings.npy", arr)
    metadata = [_meta_row(c, i) for i, c in enumerate(chunks)]
    with open(out / "metadata.json", "w", encoding="utf-8") as f:
        json.dump(metadata, f, indent=2, ensure_ascii=False)
    index = {
        "total_embeddings": len(embeddings),
        "embedding_dimension": arr.shape[1],
        "total_size_gb": arr.nbytes / 1024 / 1024 / 1024,
    }
    with open(out / "index.json", "w", encoding="utf-8") as f:
        json.dump(index, f, indent=2)


def save_embeddings_disk(chunks: List[Dict], embeddings, output_dir: str = "./embeddings_saved",
                         batch_size: int = 10000) -> None:
    """Batched layout (reference save_embeddings_to_disk.py:15-80)."""
    out = Path(output_dir)
    out.mkdir(parents=True, exist_ok=True)
    arr = _as_matrix(embeddings)
    n = len(embeddings)
    num_batches = (n + batch_size - 1) // batch_size
    for i in range(num_batches):
        lo, hi = i * batch_size, min((i + 1) * batch_size, n)
        np.save(out / f"embeddings_batch_{i:04d}.npy", arr[lo:hi])
        metadata = []
        for j, chunk in enumerate(chunks[lo:hi]):
            row = _meta_row(chunk, lo + j)
            row["batch_index"] = i
            row["batch_position"] = j
            metadata.append(row)
        with open(out / f"metadata_batch_{i:04d}.json", "w", encoding="utf-8") as f:
            json.dump(metadata, f, indent=2, ensure_ascii=False)
    index = {
        "total_embeddings": n,
        "embedding_dimension": arr.shape[1],
        "num_batches": num_batches,
        "batch_size": batch_size,
        "chunks": [c.get("chunk_id") for c in chunks],
    }
    with open(out / "index.json", "w", encoding="utf-8") as f:
        json.dump(index, f, indent=2)


def load_embeddings_from_disk(input_dir: str, batch_index: Optional[int] = None) -> Tuple[np.ndarray, list]:
    """Loader with the reference's semantics (save_embeddings_to_disk.py:82-117); additionally
    understands the single-file layout when `index.json` has no `num_batches`."""
    path = Path(input_dir)
    if batch_index is not None:
        emb = np.load(path / f"embeddings_batch_{batch_index:04d}.npy")
        with open(path / f"metadata_batch_{batch_index:04d}.json", "r", encoding="utf-8") as f:
            return emb, json.load(f)
    with open(path / "index.json", "r", encoding="utf-8") as f:
        index = json.load(f)
    if "num_batches" not in index:
        emb = np.load(path / "embeddings.npy")
        with open(path / "metadata.json", "r", encoding="utf-8") as f:
            return emb, json.load(f)
    embs, metas = [], []
    for i in range(index["num_batches"]):
        embs.append(np.load(path / f"embeddings_batch_{i:04d}.npy"))
        with open(path / f"metadata_batch_{i:04d}.json", "r", encoding="utf-8") as f:
            metas.extend(json.load(f))
    return np.vstack(embs), metas


def save_search_matrix(embeddings, output_dir: str) -> Path:
    """float32 `[N,D]` side file for the GPU search stage (mmap-able, half the float64 bytes)."""
    out = Path(output_dir)
    out.mkdir(parents=True, exist_ok=True)
    arr = np.ascontiguousarray(np.asarray(embeddings, dtype=np.float32))
    np.save(out / "embeddings_f32.npy", arr)
    return out / "embeddings_f32.npy"


def load_search_matrix(input_dir: str, mmap: bool = True) -> np.ndarray:
    """float32 search matrix; falls back to converting the reference layouts."""
    p = Path(input_dir) / "embeddings_f32.npy"
    if p.exists():
        return np.load(p, mmap_mode="r" if mmap else None)
    emb, _ = load_embeddings_from_disk(input_dir)
    return np.ascontiguousarray(emb, dtype=np.float32)


def chunk_ids_of(metadata: list) -> List[str]:
    """Row index -> chunk_id mapping used to translate search ids back to chunks."""
    return [m["chunk_id"] for m in metadata]
