"""The GPU index as the store of record — the place `store_in_chroma_batched`
(4-embed/generation/generate_embeddings_parallel.py:323-468) holds in the reference.

The reference hands its chunks and embeddings to a ChromaDB collection
(`collection.add(ids=, embeddings=, documents=, metadatas=)`, :415-422) and never queries it
(SURVEY.md F3/F4). `GpuCollection` keeps that calling surface, persists to plain files, and answers
`query` with the exact fused score+top-k search of `search.py`, row-sharded over the ranks of a
process group when there is one:

    <db_path>/<collection>/manifest.json      dim, dtype, total, [{file, rows}], format version
    <db_path>/<collection>/shard_XXXX.npy     [rows, dim] bf16 (stored as uint16) or float32
    <db_path>/<collection>/records_XXXX.json  ids, documents, metadatas of the same rows

`store_in_gpu_index_batched` mirrors the reference function's id / document / metadata rules.
Writing and reading need no GPU; `query` does (there is no CPU search path).
"""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np

from .search import CorpusIndex, ShardedCorpusIndex, shard_bounds

_FORMAT = 1


def _to_bf16_bits(a: np.ndarray) -> np.ndarray:
    """float32 -> bf16 bit patterns (round to nearest even), as uint16."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)


class GpuCollection:
    def __init__(self, db_path: str, name: str = "scientific_papers", dtype: str = "bf16",
                 shard_rows: int = 1_000_000, metadata: Optional[dict] = None):
        if dtype not in ("bf16", "float32"):
            raise ValueError("dtype must be 'bf16' or 'float32'")
        self.dir = Path(db_path) / name
        self.name = name
        self.shard_rows = int(shard_rows)
        self._pending_vec: List[np.ndarray] = []
        self._pending_rec: List[tuple] = []
        self._index = None
        self._loaded_total = -1
        mf = self.dir / "manifest.json"
        if mf.exists():
            with open(mf, "r", encoding="utf-8") as f:
                self.manifest = json.load(f)
            if self.manifest.get("format") != _FORMAT:
                raise ValueError(f"{mf}: unknown format {self.manifest.get('format')}")
        else:
            self.manifest = {"format": _FORMAT, "name": name, "dim": None, "dtype": dtype, "total": 0, "shards": [],
                             "metadata": metadata or {"description": "Scientific paper chunks for RAG"}}
        self.ids: List[str] = []
        self.documents: List[str] = []
        self.metadatas: List[dict] = []
        for sh in self.manifest["shards"]:
            with open(self.dir / sh["records"], "r", encoding="utf-8") as f:
                rec = json.load(f)
            self.ids.extend(rec["ids"])
            self.documents.extend(rec["documents"])
            self.metadatas.extend(rec["metadatas"])

    # ------------------------------------------------------------------ Chroma-shaped surface
    def count(self) -> int:
        return self.manifest["total"] + sum(v.shape[0] for v in self._pending_vec)

    def add(self, ids: Sequence[str], embeddings, documents: Optional[Sequence[str]] = None,
            metadatas: Optional[Sequence[dict]] = None) -> None:
        vec = np.asarray(embeddings, dtype=np.float32)
        if vec.ndim != 2 or vec.shape[0] != len(ids):
            raise ValueError("embeddings must be [len(ids), dim]")
        if self.manifest["dim"] is None:
            self.manifest["dim"] = int(vec.shape[1])
        elif vec.shape[1] != self.manifest["dim"]:
            raise ValueError(f"embedding dimension {vec.shape[1]} != {self.manifest['dim']}")
        documents = list(documents) if documents is not None else [""] * len(ids)
        metadatas = list(metadatas) if metadatas is not None else [{} for _ in ids]
        if len(documents) != len(ids) or len(metadatas) != len(ids):
            raise ValueError("ids, documents and metadatas differ in length")
        self._pending_vec.append(vec)
        self._pending_rec.append((list(ids), documents, metadatas))
        while sum(v.shape[0] for v in self._pending_vec) >= self.shard_rows:
            self._flush(self.shard_rows)

    def persist(self) -> None:
        """Write whatever `add` has buffered (the last shard may be short) and the manifest."""
        n = sum(v.shape[0] for v in self._pending_vec)
        if n:
            self._flush(n)
        self._write_manifest()

    def _flush(self, rows: int) -> None:
        vec = np.concatenate(self._pending_vec, 0)
        ids = [x for r in self._pending_rec for x in r[0]]
        docs = [x for r in self._pending_rec for x in r[1]]
        metas = [x for r in self._pending_rec for x in r[2]]
        take, rest = vec[:rows], vec[rows:]
        self.dir.mkdir(parents=True, exist_ok=True)
        i = len(self.manifest["shards"])
        shard, records = f"shard_{i:04d}.npy", f"records_{i:04d}.json"
        data = _to_bf16_bits(take) if self.manifest["dtype"] == "bf16" else np.ascontiguousarray(take, np.float32)
        tmp = self.dir / (shard + ".tmp")
        with open(tmp, "wb") as f:
            np.save(f, data)
        os.replace(tmp, self.dir / shard)
        with open(self.dir / (records + ".tmp"), "w", encoding="utf-8") as f:
            json.dump({"ids": ids[:rows], "documents": docs[:rows], "metadatas": metas[:rows]}, f, ensure_ascii=False)
        os.replace(self.dir / (records + ".tmp"), self.dir / records)
        self.manifest["shards"].append({"file": shard, "records": records, "rows": int(rows)})
        self.manifest["total"] += int(rows)
        self.ids.extend(ids[:rows])
        self.documents.extend(docs[:rows])
        self.metadatas.extend(metas[:rows])
        self._pending_vec = [rest] if rest.shape[0] else []
        self._pending_rec = [(ids[rows:], docs[rows:], metas[rows:])] if rest.shape[0] else []
        self._write_manifest()

    def _write_manifest(self) -> None:
        self.dir.mkdir(parents=True, exist_ok=True)
        tmp = self.dir / "manifest.json.tmp"
        with open(tmp, "w", encoding="utf-8") as f:
            json.dump(self.manifest, f, indent=2)
        os.replace(tmp, self.dir / "manifest.json")  # the manifest lists only complete shards

    # ------------------------------------------------------------------ rows -> HBM
    def load_rows(self, lo: int, hi: int) -> np.ndarray:
        """Persisted rows [lo, hi) in the stored dtype (uint16 bf16 bits or float32), read through
        memory maps so that a rank only touches its own range."""
        out, base = [], 0
        for sh in self.manifest["shards"]:
            a, b = max(lo, base), min(hi, base + sh["rows"])
            if a < b:
                out.append(np.asarray(np.load(self.dir / sh["file"], mmap_mode="r")[a - base:b - base]))
            base += sh["rows"]
        if not out:
            return np.zeros((0, self.manifest["dim"] or 0), np.uint16 if self.manifest["dtype"] == "bf16" else np.float32)
        return np.concatenate(out, 0)

    def _device_index(self, group=None, device: Optional[int] = None):
        import torch
        import torch.distributed as dist

        total = self.manifest["total"]
        if self._index is not None and self._loaded_total == total:
            return self._index
        sharded = group is not None or (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)
        world = dist.get_world_size(group) if sharded else 1
        rank = dist.get_rank(group) if sharded else 0
        lo, hi = shard_bounds(total, world, rank)
        rows = self.load_rows(lo, hi)
        t = torch.from_numpy(rows)
        t = t.view(torch.bfloat16) if self.manifest["dtype"] == "bf16" else t
        self._index = ShardedCorpusIndex(t, total, group=group, device=device) if sharded else CorpusIndex(t, device=device)
        self._loaded_total = total
        return self._index

    def query(self, query_embeddings, n_results: int = 10, group=None, device: Optional[int] = None,
              include: Sequence[str] = ("documents", "metadatas", "distances")) -> Dict[str, list]:
        """Exact cosine top-`n_results` over the PERSISTED rows (call `persist` after `add`).
        Chroma-shaped result: lists per query of ids / distances (1 - cosine) / documents / metadatas,
        plus 'scores' (the cosines). Under a process group every rank holds a row shard and returns
        the same global answer."""
        import torch

        if self.manifest["total"] == 0:
            raise ValueError("the collection holds no persisted rows")
        index = self._device_index(group, device)
        q = torch.as_tensor(np.asarray(query_embeddings, dtype=np.float32))
        if q.dim() == 1:
            q = q[None, :]
        k = min(int(n_results), self.manifest["total"])
        scores, ids = index.search(q.to(torch.bfloat16) if self.manifest["dtype"] == "bf16" else q, k)
        scores, ids = scores.cpu().numpy(), ids.cpu().numpy()
        out: Dict[str, list] = {"ids": [[self.ids[j] for j in row] for row in ids], "scores": scores.tolist()}
        if "distances" in include:
            out["distances"] = (1.0 - scores).tolist()
        if "documents" in include:
            out["documents"] = [[self.documents[j] for j in row] for row in ids]
        if "metadatas" in include:
            out["metadatas"] = [[self.metadatas[j] for j in row] for row in ids]
        return out


def store_in_gpu_index_batched(chunks: List[Dict], embeddings: List, db_path: str = "./gpu_index",
                               collection_name: str = "scientific_papers", batch_size: int = 2000,
                               dtype: str = "bf16") -> GpuCollection:
    """`store_in_chroma_batched` (:323-468) with the GPU index as the store: same truncation to the
    shorter of chunks / embeddings (:335-341), same ids (`chunk_id` or `chunk_{i}`), documents
    (`text`) and metadata fields (:393-408). No retry ladder: writes are plain files and raise."""
    n = min(len(chunks), len(embeddings))
    col = GpuCollection(db_path, collection_name, dtype=dtype)
    for i in range(0, n, batch_size):
        batch = chunks[i:i + batch_size]
        ids, docs, metas = [], [], []
        for j, chunk in enumerate(batch):
            md = chunk.get("metadata", {})
            ids.append(chunk.get("chunk_id", f"chunk_{i + j}"))
            docs.append(chunk["text"])
            metas.append({"paper_id": str(md.get("paper_id", "unknown")), "section": str(md.get("section", "unknown")),
                          "quality_score": float(md.get("quality_score", 0.0)),
                          "chunk_index": str(md.get("chunk_index", i + j))})
        col.add(ids=ids, embeddings=np.stack([np.asarray(e, dtype=np.float32) for e in embeddings[i:i + len(batch)]]),
                documents=docs, metadatas=metas)
    col.persist()
    return col
