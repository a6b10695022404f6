"""Mint the saved-layout fixtures under tests/golden/layout_* from the REFERENCE's own writers
(run in the dev container, where /root/reference is mounted; the fixtures travel, the reference
does not).

  layout_batched/  <- 4-embed/utils/save_embeddings_to_disk.py::save_embeddings_disk (imported as a module)
  layout_single/   <- generate_embeddings_parallel.py:271-321 save_embeddings_to_disk_fallback. That
                      file cannot be imported (SyntaxError at :239), so the function's
                      source lines are read from the reference at run time and executed as they are;
                      nothing of it is copied into this repository.

    python tools/make_layout_golden.py [out_root]
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("ARB_REFERENCE", "/root/reference"))


def fixture_inputs():
    """7 chunks (one without chunk_id, non-ASCII text) x 8-d float32 rows, as `encode` returns them."""
    rng = np.random.default_rng(42)
    rows = [rng.standard_normal(8).astype(np.float32) for _ in range(7)]
    chunks = [{"chunk_id": f"2101.{i:05d}_chunk_{i % 3}", "text": f"text {i} é — 你好", "metadata":
               {"paper_id": f"2101.{i:05d}", "section": "intro" if i % 2 else None, "quality_score": 0.9 + 0.01 * i}} for i in range(7)]
    del chunks[4]["chunk_id"]
    return chunks, rows


def reference_writers():
    spec = importlib.util.spec_from_file_location("ref_save", REF / "4-embed/utils/save_embeddings_to_disk.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lines = (REF / "4-embed/generation/generate_embeddings_parallel.py").read_text(encoding="utf-8").splitlines()
    src = "\n".join(lines[270:321])  # :271-321, the whole function
    ns = {"List": list, "Dict": dict, "Path": Path, "np": np, "json": __import__("json")}
    exec(compile("from typing import List, Dict\n" + src, "generate_embeddings_parallel.py:271-321", "exec"), ns)
    return mod.save_embeddings_disk, ns["save_embeddings_to_disk_fallback"]


def main(out_root: Path):
    batched, single = reference_writers()
    chunks, rows = fixture_inputs()
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        batched(chunks, rows, str(out_root / "layout_batched"), batch_size=3)
        single(chunks, rows, str(out_root / "layout_single"))
    for p in sorted(out_root.glob("layout_*/*")):
        print(p.relative_to(out_root), p.stat().st_size)


if __name__ == "__main__":
    main(Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "tests" / "golden")
