"""Semantic-boundary test of the reference's chunker on the GPU.

`TextChunker._chunk_semantic` (3-chunks/pipeline/src/processors/text_processor.py:1547-1561)
encodes the sentences of a paper with all-MiniLM-L6-v2 (:1379-1396) and starts a new chunk when
`_cosine_similarity(embedding[i], embedding[i-1]) < 0.7` (:1557-1561, formula :1601-1605) or the
size cap is hit. This module provides the two numeric pieces: sentence embeddings come from
`B200SentenceEncoder(model_name='all-MiniLM-L6-v2')`, the adjacent-pair cosines from one kernel
(`arb_adjacent_cosine`), and `semantic_breaks` applies the reference's threshold rule. The string
handling around it (sentence split, overlap, size caps) is host-side text processing and stays out
of scope.
"""
from __future__ import annotations

import numpy as np

from . import _lib

SEMANTIC_BREAK_THRESHOLD = 0.7  # text_processor.py:1560


def adjacent_cosine(embeddings):
    """cos(e[i], e[i-1]) for consecutive rows of a float32 `[n, D]` CUDA tensor (or numpy array,
    copied to the current device); element 0 is 1. Returns a CUDA float32 `[n]` tensor."""
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("arxiv_rag_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
    if isinstance(embeddings, np.ndarray):
        embeddings = torch.from_numpy(np.ascontiguousarray(embeddings, dtype=np.float32)).cuda()
    e = embeddings.to(torch.float32).contiguous()
    if e.dim() != 2:
        raise ValueError("embeddings must be [n, D]")
    out = torch.empty(e.shape[0], dtype=torch.float32, device=e.device)
    if e.shape[0]:
        with torch.cuda.device(e.device):
            _lib.check(_lib.lib().arb_adjacent_cosine(_lib.ptr(e), e.shape[0], e.shape[1], _lib.ptr(out),
                                                      _lib.current_stream()))
    return out


def semantic_breaks(embeddings, threshold: float = SEMANTIC_BREAK_THRESHOLD) -> np.ndarray:
    """Boolean `[n]`: sentence i starts a new chunk because its similarity to sentence i-1 dropped
    below `threshold` (the semantic half of the rule at text_processor.py:1555-1561; index 0 is
    never a break)."""
    sim = adjacent_cosine(embeddings).cpu().numpy()
    out = sim < threshold
    if len(out):
        out[0] = False
    return out
