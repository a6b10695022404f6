"""B200SentenceEncoder — drop-in for the `SentenceTransformer('all-mpnet-base-v2')` object that
`generate_embeddings_parallel.py` keeps in `_worker_model` (:37, :47) and calls as
`model.encode(batch, batch_size=..., normalize_embeddings=True, show_progress_bar=False,
convert_to_numpy=True, convert_to_tensor=False)` (:146-153, :160-165) and
`model.get_sentence_embedding_dimension()` (:169).

Semantics restated from sentence-transformers `encode`: sort inputs by length
descending, batch, pad each batch to its longest row, MPNet forward, masked mean-pool,
L2-normalise, undo the sort, return float32 `[n, 768]` rows in input order.

All arithmetic runs in the CUDA library (`_lib`); torch is used for device memory, pinned
staging buffers and streams only. There is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .weights import ALL_MPNET_BASE_V2, MPNetArch, PackedWeights, synthetic_state_dict


def _require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("arxiv_rag_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
    return torch


class B200SentenceEncoder:
    """MPNet sentence encoder resident on one B200.

    Parameters
    ----------
    state_dict : HF `MPNetModel.state_dict()` (or numpy arrays under the same names).  None ->
        seeded synthetic weights (`weights.synthetic_state_dict(seed)`), since no checkpoint
        exists offline.
    tokenizer : optional callable `tokenizer(list[str], padding=True, truncation=True,
        max_length=..., return_tensors='np') -> {'input_ids', 'attention_mask'}` — a HF tokenizer or
        the in-tree `tokenizer.WordPieceTokenizer`. `vocab_file=...` builds the in-tree native one
        (`tokenizer.NativeWordPieceTokenizer`, C-ABI `arb_tokenizer_*`) from the model's `vocab.txt`. Without either, `encode` accepts pre-tokenised `(input_ids, attention_mask)`
        only (no vocabulary ships offline).
    max_batch / max_seq : capacity of the activation workspace (tokens = max_batch * max_seq).
    dtype : 16-bit format of weights, activations and tensor-core operands (fp32 accumulation and
        statistics always; both operands of a tcgen05 MMA must share one format):
        'fp16' (default) — cosine vs the fp32 reference >= 0.9999 on every row (~0.999997), also on
        heavy-tailed weights; the format the reference's own config asks for (`fp16: true`,
        3-chunks/pipeline/config.yaml:49).
        'bf16' — what BASELINE configs[1] names. Same speed, 8-bit mantissa: ~0.99995 on full rows;
        rows shorter than `short_seq` (32) tokens are batched apart and run in fp16 (an fp16 copy
        of the weights), because nothing averages a short row's rounding noise. Meets 0.9999 on
        well-behaved weights, not on heavy-tailed ones (DESIGN.md 'Numerics').
        'bf16_pure' — bf16 for every batch: the A/B baseline (~0.9998 on 1-token rows).
    """

    DTYPES = {"fp16": _lib.ARB_DTYPE_F16, "bf16": _lib.ARB_DTYPE_BF16, "bf16_pure": _lib.ARB_DTYPE_BF16_PURE}

    def __init__(self, state_dict: dict | None = None, arch: MPNetArch = ALL_MPNET_BASE_V2,
                 device: int | None = None, max_batch: int = 1024, max_seq: int | None = None,
                 tokenizer=None, seed: int = 0, dtype: str = "fp16", model_name: str | None = None,
                 vocab_file: str | None = None):
        torch = _require_cuda()
        self._torch = torch
        if model_name is not None:  # 'all-mpnet-base-v2' | 'all-MiniLM-L6-v2' (reference CLI choices, :473-475)
            from .weights import ARCH_BY_MODEL_NAME

            if model_name not in ARCH_BY_MODEL_NAME:
                raise ValueError(f"model '{model_name}' not supported (have {sorted(ARCH_BY_MODEL_NAME)})")
            arch = ARCH_BY_MODEL_NAME[model_name]
        self.arch = arch
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.max_seq = int(max_seq or arch.max_seq_length)
        self.max_batch = int(max_batch)
        self.max_seq_length = min(arch.max_seq_length, self.max_seq)
        if tokenizer is None and vocab_file is not None:
            from .tokenizer import NativeWordPieceTokenizer

            tokenizer = NativeWordPieceTokenizer(vocab_file, kind="bert" if arch.kind == "bert" else "mpnet",
                                                 max_length=self.max_seq_length)
        self.tokenizer = tokenizer
        if state_dict is None:
            state_dict = synthetic_state_dict(arch, seed)
        packed = PackedWeights(arch, state_dict)
        if dtype not in self.DTYPES:
            raise ValueError(f"dtype must be one of {sorted(self.DTYPES)}")
        self.dtype = dtype
        cfg = arch.c_struct(self.DTYPES[dtype])
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().arb_mpnet_create(C.byref(cfg), C.byref(packed.struct),
                                                   self.max_batch * self.max_seq, self.max_seq,
                                                   self.device, C.byref(handle)))
        self._h = handle
        # rows shorter than this are batched among themselves (a bf16 handle runs such batches in
        # fp16); 0 = no split
        self.short_seq = int(_lib.lib().arb_mpnet_short_seq(handle))
        self._staging = None

    # ------------------------------------------------------------------ reference surface
    def get_sentence_embedding_dimension(self) -> int:
        return self.arch.hidden_size

    def encode(self, sentences, batch_size: int = 32, show_progress_bar: bool | None = None,
               convert_to_numpy: bool = True, convert_to_tensor: bool = False,
               normalize_embeddings: bool = False, **_ignored):
        """Same call shape as SentenceTransformer.encode. all-mpnet-base-v2 ends with a Normalize
        module, so rows are unit-norm whether or not `normalize_embeddings` is set."""
        torch = self._torch
        single = isinstance(sentences, str)
        if single:
            sentences = [sentences]
        if self._is_text(sentences):
            out = self._encode_texts(list(sentences), batch_size)
            if convert_to_tensor:
                self.check_status(synchronize=True)
                return out[0] if single else out
            res = self._to_host(out)
            self.check_status()
            return res[0] if single else res
        ids, mask = self._tokenize(sentences)
        n = ids.shape[0]
        out = torch.empty((n, self.arch.hidden_size), dtype=torch.float32, device=f"cuda:{self.device}")
        if n:
            lengths = mask.sum(axis=1)
            order = np.argsort(-lengths, kind="stable")  # longest first, like sentence-transformers
            bs = max(1, min(int(batch_size), self.max_batch))
            with torch.cuda.device(self.device):
                for sel in self._batches(order, lengths, bs):
                    S = max(int(lengths[sel].max()), 1)  # pad to the longest row of the batch
                    d_ids, d_mask, d_rows = self._to_device(ids[sel, :S], mask[sel, :S], sel)
                    out.index_copy_(0, d_rows, self.encode_tokens(d_ids, d_mask))
        if convert_to_tensor:
            self.check_status(synchronize=True)
            return out[0] if single else out
        res = self._to_host(out)  # synchronises the stream
        self.check_status()
        return res[0] if single else res

    def _to_host(self, out) -> np.ndarray:
        """Device rows -> numpy through page-locked memory (torch's caching host allocator recycles the
        block once the returned array is dropped): a pageable `.cpu()` of 1 M x 768 floats runs at a
        few GB/s and sits, un-overlapped, at the end of every `encode` call."""
        torch = self._torch
        if out.numel() * out.element_size() > (1 << 30):  # do not page-lock gigabytes for one result
            return out.cpu().numpy()
        host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
        host.copy_(out, non_blocking=True)
        torch.cuda.current_stream(out.device).synchronize()
        return host.numpy()

    @staticmethod
    def _is_text(sentences) -> bool:
        if isinstance(sentences, (dict, tuple)):
            return False
        try:
            return len(sentences) > 0 and isinstance(sentences[0], str)
        except TypeError:
            return False

    def _encode_texts(self, sentences: list, batch_size: int):
        """Strings -> embeddings the way sentence-transformers feeds its model: sort by text length
        (longest first), cut into batches, tokenise each batch and pad it to its longest row. The
        tokeniser runs on a background thread two batches ahead of the GPU, and each tokenised batch
        goes through the two pinned staging buffers, so host work overlaps the forward passes."""
        from concurrent.futures import ThreadPoolExecutor

        torch = self._torch
        if self.tokenizer is None:
            raise RuntimeError(
                "no tokenizer: the all-mpnet-base-v2 vocabulary is not available offline; pass "
                "tokenizer=... / vocab_file=... or pre-tokenised (input_ids, attention_mask)")
        n = len(sentences)
        out = torch.empty((n, self.arch.hidden_size), dtype=torch.float32, device=f"cuda:{self.device}")
        order = np.argsort(-np.fromiter((len(t) for t in sentences), dtype=np.int64, count=n), kind="stable")
        bs = max(1, min(int(batch_size), self.max_batch))
        groups = [order[i:i + bs] for i in range(0, n, bs)]

        def tokenize(sel):
            enc = self.tokenizer([sentences[j] for j in sel], padding=True, truncation=True,
                                 max_length=self.max_seq_length, return_tensors="np")
            ids = np.asarray(enc["input_ids"]).astype(np.int32, copy=False)
            mask = np.asarray(enc["attention_mask"]).astype(np.int32, copy=False)
            return ids[:, :self.max_seq], mask[:, :self.max_seq]

        with ThreadPoolExecutor(max_workers=1) as pool, torch.cuda.device(self.device):
            ahead = [pool.submit(tokenize, g) for g in groups[:2]]
            for k, sel in enumerate(groups):
                ids, mask = ahead.pop(0).result()
                if k + 2 < len(groups):
                    ahead.append(pool.submit(tokenize, groups[k + 2]))
                lengths = mask.sum(axis=1)
                rows = np.argsort(-lengths, kind="stable")
                for part in self._batches(rows, lengths, bs):  # a bf16 handle takes its short rows apart
                    S = max(int(lengths[part].max()), 1)
                    d_ids, d_mask, d_rows = self._to_device(ids[part, :S], mask[part, :S], sel[part])
                    out.index_copy_(0, d_rows, self.encode_tokens(d_ids, d_mask))
        return out

    def _batches(self, order: np.ndarray, lengths: np.ndarray, bs: int):
        """Length-sorted rows cut into batches of `bs`; rows shorter than `short_seq` never share a
        batch with longer ones (a batch's padded length decides its activation format)."""
        n_long = int((lengths[order] >= self.short_seq).sum()) if self.short_seq else len(order)
        for lo, hi in ((0, n_long), (n_long, len(order))):
            for start in range(lo, hi, bs):
                yield order[start:min(start + bs, hi)]

    def check_status(self, synchronize: bool = False) -> None:
        """Raise if a kernel of an earlier call met a token id outside the vocabulary (torch's
        embedding raises there; the kernels clamp and flag). The stream must be idle."""
        if synchronize:
            self._torch.cuda.synchronize(self.device)
        _lib.check(_lib.lib().arb_mpnet_status(self._h))

    # ------------------------------------------------------------------ device fast path
    def encode_tokens(self, ids_dev, mask_dev, out=None):
        """ids/mask: CUDA int32 [B,S] tensors -> CUDA float32 [B,H] (enqueued on the current stream)."""
        torch = self._torch
        if ids_dev.dtype != torch.int32 or mask_dev.dtype != torch.int32:
            raise TypeError("encode_tokens expects int32 CUDA tensors")
        if not (ids_dev.is_cuda and mask_dev.is_cuda and ids_dev.is_contiguous() and mask_dev.is_contiguous()):
            raise ValueError("encode_tokens expects contiguous CUDA tensors")
        B, S = ids_dev.shape
        if out is None:
            out = torch.empty((B, self.arch.hidden_size), dtype=torch.float32, device=ids_dev.device)
        # one handle = one device and one stream at a time (its activation buffers are shared)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().arb_mpnet_encode(self._h, _lib.ptr(ids_dev), _lib.ptr(mask_dev), B, S,
                                                   _lib.ptr(out), _lib.current_stream()))
        return out

    def encode_tokens_graphed(self, ids_dev, mask_dev):
        """Same as `encode_tokens` through a CUDA graph cached per (B, S): the 86 launches of a
        forward become one graph launch, which is what bounds small-batch (query-time) latency.
        Inputs are copied into the graph's static buffers; the returned tensor is the graph's
        static output (valid until the next call with the same shape)."""
        torch = self._torch
        B, S = ids_dev.shape
        key = (B, S)
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        if key not in self._graphs:
            s_ids, s_mask = torch.empty_like(ids_dev), torch.empty_like(mask_dev)
            s_out = torch.empty((B, self.arch.hidden_size), dtype=torch.float32, device=ids_dev.device)
            s_ids.copy_(ids_dev)
            s_mask.copy_(mask_dev)
            self.encode_tokens(s_ids, s_mask, s_out)  # warm-up outside capture (module load, attributes)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.encode_tokens(s_ids, s_mask, s_out)
            self._graphs[key] = (graph, s_ids, s_mask, s_out)
        graph, s_ids, s_mask, s_out = self._graphs[key]
        s_ids.copy_(ids_dev, non_blocking=True)
        s_mask.copy_(mask_dev, non_blocking=True)
        graph.replay()
        return s_out

    @property
    def launches_per_encode(self) -> int:
        return int(_lib.lib().arb_mpnet_launches_per_encode(self._h))

    @property
    def device_bytes(self) -> int:
        return int(_lib.lib().arb_mpnet_device_bytes(self._h))

    # ------------------------------------------------------------------ helpers
    def _tokenize(self, sentences) -> tuple[np.ndarray, np.ndarray]:
        if isinstance(sentences, dict):
            sentences = (sentences["input_ids"], sentences["attention_mask"])
        if isinstance(sentences, tuple) and len(sentences) == 2 and not isinstance(sentences[0], str):
            ids = np.asarray(_to_numpy(sentences[0]))
            mask = np.asarray(_to_numpy(sentences[1]))
        else:
            sentences = list(sentences)
            if len(sentences) == 0:
                return (np.zeros((0, 1), np.int32), np.zeros((0, 1), np.int32))
            if not isinstance(sentences[0], str):
                raise TypeError("encode expects list[str] or a (input_ids, attention_mask) pair")
            if self.tokenizer is None:
                raise RuntimeError(
                    "no tokenizer: the all-mpnet-base-v2 vocabulary is not available offline; pass "
                    "tokenizer=... or pre-tokenised (input_ids, attention_mask)")
            enc = self.tokenizer(sentences, padding=True, truncation=True,
                                 max_length=self.max_seq_length, return_tensors="np")
            ids, mask = np.asarray(enc["input_ids"]), np.asarray(enc["attention_mask"])
        if ids.ndim != 2 or ids.shape != mask.shape:
            raise ValueError(f"input_ids {ids.shape} / attention_mask {mask.shape} must be equal 2-D shapes")
        if ids.shape[1] > self.max_seq:
            ids, mask = ids[:, :self.max_seq], mask[:, :self.max_seq]  # tokenizer-style truncation
        return ids.astype(np.int32, copy=False), mask.astype(np.int32, copy=False)

    def _to_device(self, ids: np.ndarray, mask: np.ndarray, rows: np.ndarray | None = None):
        """Pinned staging + async H2D on the current stream. Two pinned buffers of the handle's
        capacity, used alternately and sliced per batch: filling one overlaps the copy out of the
        other, and no shape ever needs a new pinned allocation. `rows` (the batch's positions in the
        caller's order) ride in the same copy: a separate pageable H2D would make the driver wait for
        the stream to drain first, i.e. serialise the host with every batch's forward."""
        torch = self._torch
        B, S = ids.shape
        if self._staging is None:
            cap = 2 * self.max_batch * self.max_seq + self.max_batch
            self._staging = [[torch.empty(cap, dtype=torch.int32, pin_memory=True), torch.cuda.Event()]
                             for _ in range(2)]
            self._staging_next = 0
        st, ev = self._staging[self._staging_next]
        self._staging_next ^= 1
        ev.synchronize()  # the previous H2D out of this pinned buffer must have drained
        n = 2 * B * S
        view = st[:n + (B if rows is not None else 0)]
        host = view.numpy()
        host[:B * S].reshape(B, S)[...] = ids
        host[B * S:n].reshape(B, S)[...] = mask
        if rows is not None:
            host[n:] = rows
        dev = view.to(f"cuda:{self.device}", non_blocking=True)
        ev.record()
        d_ids, d_mask = dev[:B * S].view(B, S), dev[B * S:n].view(B, S)
        return (d_ids, d_mask) if rows is None else (d_ids, d_mask, dev[n:].long())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().arb_mpnet_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _to_numpy(x):
    if hasattr(x, "detach"):
        return x.detach().cpu().numpy()
    return x
