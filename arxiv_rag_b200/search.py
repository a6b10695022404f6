"""Exact cosine top-k search over a (row-sharded) chunk-embedding matrix.

The reference implies but never wrote this stage (`retrieval.top_k: 10` at
3-chunks/pipeline/config.yaml:62-64 is never read); its only cosine is
`TextChunker._cosine_similarity` (text_processor.py:1601-1605). On the unit-norm rows that
`encode(..., normalize_embeddings=True)` (generate_embeddings_parallel.py:149) produces, cosine
is the dot product, so search is `top_k(Q @ C.T)` with ties ordered by ascending row id.

Multi-GPU (one process per GPU, torch.distributed): the corpus is row-sharded, every rank runs
the fused score+top-k kernel on its shard, the per-rank `[Q,k]` lists are all-gathered over
NCCL/NVLink and merged by the k-way merge kernel (`arb_topk_merge`).
"""
from __future__ import annotations

import os

import numpy as np

from . import _lib


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("arxiv_rag_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
    return torch


def shard_bounds(n_rows: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous row range of `rank` (rows r*N/G .. (r+1)*N/G)."""
    lo = (n_rows * rank) // world_size
    hi = (n_rows * (rank + 1)) // world_size
    return lo, hi


class CorpusIndex:
    """A corpus shard resident in HBM plus the reusable search workspace.

    corpus : `[N, D]` float32 or bfloat16 (numpy / torch, host or device). bf16 is searched as
        stored. float32 stays float32: one tf32 tensor-core pass over the stored rows, the k + 22 best
        re-scored in exact fp32, and a per-query proof that no other row can belong to the top-k
        (tf32 scoring is off by at most 2^-9 |q| |c|); the few queries without that proof (dense
        near-ties around rank k) are re-run through the 3-term split-bf16 path, so the result is
        the exact fp32 top-k either way. No converted copy of the corpus is kept.
    id_offset : global id of local row 0 (for row-sharded corpora).
    """

    def __init__(self, corpus, id_offset: int = 0, device: int | None = None, dtype=None):
        torch = _torch()
        self._torch = torch
        self.device = torch.cuda.current_device() if device is None else int(device)
        dev = f"cuda:{self.device}"
        if isinstance(corpus, np.ndarray):
            corpus = torch.from_numpy(np.ascontiguousarray(corpus))
        if dtype is not None:
            corpus = corpus.to(dtype)
        if corpus.dtype == torch.float64:  # the reference's embeddings.npy is float64 (:281-284)
            corpus = corpus.to(torch.float32)
        if corpus.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError(f"corpus dtype {corpus.dtype} unsupported (float32 / bfloat16)")
        if corpus.dim() != 2:
            raise ValueError("corpus must be [N, D]")
        self.corpus = corpus.to(dev).contiguous()
        self.n, self.d = self.corpus.shape
        self.id_offset = int(id_offset)
        self.dtype_code = _lib.ARB_DTYPE_F32 if self.corpus.dtype == torch.float32 else _lib.ARB_DTYPE_BF16
        self._ws = None
        self.max_norm = 1.0
        self.fallback_queries = 0  # fp32 corpora: queries that needed the exact split-bf16 re-run so far
        if self.dtype_code == _lib.ARB_DTYPE_F32 and self.n:
            mx = 0.0
            for s0 in range(0, self.n, 1 << 20):  # slabs: no [N] temporary next to a large corpus
                mx = max(mx, float(torch.linalg.vector_norm(self.corpus[s0:s0 + (1 << 20)], dim=1).max()))
            self.max_norm = mx * (1.0 + 1e-6) + 1e-30

    def workspace_bytes(self, Q: int, k: int, mode: int = 0) -> int:
        if self.dtype_code == _lib.ARB_DTYPE_F32:
            return int(_lib.lib().arb_topk_search_f32_workspace_bytes(Q, self.n, self.d, k, mode))
        return int(_lib.lib().arb_topk_search_workspace_bytes(self.dtype_code, Q, self.n, self.d, k))

    def new_workspace(self, Q: int, k: int):
        """A workspace tensor of its own for one (Q, k) — what a captured CUDA graph must use: the
        shared eager workspace below is replaced when a later call needs more bytes, and a graph
        that kept its address would then write into freed memory."""
        return self._torch.empty(max(self.workspace_bytes(Q, k), 256), dtype=self._torch.uint8, device=self.corpus.device)

    def _workspace(self, Q: int, k: int):
        need = self.workspace_bytes(Q, k)
        if self._ws is None or self._ws.numel() < need:
            self._ws = self._torch.empty(max(need, 256), dtype=self._torch.uint8, device=self.corpus.device)
        return self._ws, need

    def search(self, queries, k: int = 10, out_scores=None, out_ids=None, workspace=None):
        """queries `[Q, D]` (same dtype family as the corpus; converted if not) ->
        (scores float32 `[Q,k]`, ids int64 `[Q,k]`) as CUDA tensors, enqueued on the current stream.
        `workspace`: a tensor from `new_workspace(Q, k)` to use instead of the shared one."""
        torch = self._torch
        if isinstance(queries, np.ndarray):
            queries = torch.from_numpy(np.ascontiguousarray(queries))
        q = queries.to(self.corpus.device, non_blocking=True).to(self.corpus.dtype).contiguous()
        if q.dim() != 2 or q.shape[1] != self.d:
            raise ValueError(f"queries must be [Q, {self.d}]")
        Q = q.shape[0]
        if out_scores is None:
            out_scores = torch.empty((Q, k), dtype=torch.float32, device=q.device)
        if out_ids is None:
            out_ids = torch.empty((Q, k), dtype=torch.int64, device=q.device)
        if Q == 0:
            return out_scores, out_ids
        if self.n == 0:
            out_scores.fill_(float("-inf"))
            out_ids.fill_(-1)
            return out_scores, out_ids
        if workspace is not None:
            ws = workspace
            if ws.numel() < self.workspace_bytes(Q, k):
                raise ValueError("workspace too small for this (Q, k)")
        else:
            ws, _ = self._workspace(Q, k)
        with torch.cuda.device(self.device):
            if self.dtype_code == _lib.ARB_DTYPE_F32:
                self._search_f32(q, k, out_scores, out_ids, ws)
            else:
                _lib.check(_lib.lib().arb_topk_search(_lib.ptr(q), _lib.ptr(self.corpus), self.dtype_code, Q,
                                                      self.n, self.d, k, _lib.ptr(out_scores), _lib.ptr(out_ids),
                                                      self.id_offset, _lib.ptr(ws), ws.numel(),
                                                      _lib.current_stream()))
        return out_scores, out_ids

    def _search_f32(self, q, k, out_scores, out_ids, ws):
        """tf32 pass + exact re-score + verdict; unverified queries go through the split-bf16 path.
        Reads one int back from the device (the number of unverified queries)."""
        torch = self._torch
        lib = _lib.lib()
        Q = q.shape[0]
        flags = torch.empty(Q, dtype=torch.int32, device=q.device)
        _lib.check(lib.arb_topk_search_f32(_lib.ptr(q), _lib.ptr(self.corpus), Q, self.n, self.d, k, self.max_norm,
                                           _lib.ptr(out_scores), _lib.ptr(out_ids), self.id_offset, _lib.ptr(flags), 0,
                                           _lib.ptr(ws), ws.numel(), _lib.current_stream()))
        bad = torch.nonzero(flags, as_tuple=False).flatten()  # synchronises
        if bad.numel() == 0:
            return
        self.fallback_queries += int(bad.numel())
        qb = q[bad].contiguous()
        need = self.workspace_bytes(int(bad.numel()), k, mode=1)
        ws1 = torch.empty(max(need, 256), dtype=torch.uint8, device=q.device)
        s1 = torch.empty((bad.numel(), k), dtype=torch.float32, device=q.device)
        i1 = torch.empty((bad.numel(), k), dtype=torch.int64, device=q.device)
        _lib.check(lib.arb_topk_search_f32(_lib.ptr(qb), _lib.ptr(self.corpus), int(bad.numel()), self.n, self.d, k,
                                           self.max_norm, _lib.ptr(s1), _lib.ptr(i1), self.id_offset, 0, 1,
                                           _lib.ptr(ws1), ws1.numel(), _lib.current_stream()))
        out_scores[bad] = s1
        out_ids[bad] = i1

    @property
    def launches_per_search(self) -> int:
        return int(_lib.lib().arb_topk_search_launches(self.dtype_code))


def merge_topk(scores, ids, out_scores=None, out_ids=None):
    """Merge `[G,Q,k]` sorted lists (score desc, id asc) into `[Q,k]` on the GPU."""
    torch = _torch()
    G, Q, k = scores.shape
    scores = scores.contiguous()
    ids = ids.contiguous()
    if out_scores is None:
        out_scores = torch.empty((Q, k), dtype=torch.float32, device=scores.device)
    if out_ids is None:
        out_ids = torch.empty((Q, k), dtype=torch.int64, device=scores.device)
    if Q:
        with torch.cuda.device(scores.device):
            _lib.check(_lib.lib().arb_topk_merge(_lib.ptr(scores), _lib.ptr(ids), G, Q, k,
                                                 _lib.ptr(out_scores), _lib.ptr(out_ids),
                                                 _lib.current_stream()))
    return out_scores, out_ids


def search(queries, corpus, k: int = 10, return_numpy: bool = True):
    """One-shot API: `search(queries[Q,D], corpus[N,D], k) -> (scores[Q,k] f32, ids[Q,k] i64)`."""
    idx = CorpusIndex(corpus)
    s, i = idx.search(queries, k)
    if return_numpy:
        return s.cpu().numpy(), i.cpu().numpy()
    return s, i


class ShardedCorpusIndex:
    """Row-sharded corpus across the ranks of a torch.distributed process group.

    Every rank holds rows `shard_bounds(N, world, rank)` and the full query batch; `search`
    returns the identical global top-k on every rank: local fused top-k written as one record
    (`[Q,k]` float32 scores + int64 ids in a single buffer) -> exchange -> k-way merge.

    Exchange, ranks of one node (the 8xB200 NVSwitch box): ONE kernel per rank stores the record
    into every peer's exchange buffer (mapped through CUDA IPC), raises a flag, waits for the G
    records of the round and merges them (`arb_topk_exchange_merge`) — no library collective on
    the path. Records larger than a slot (`peer_slot_bytes`), ranks on different hosts, or
    `peer_exchange=False`: one NCCL all_gather of the records + `arb_topk_merge_records`.
    `search_graphed` replays search + exchange + merge from a CUDA graph per (Q, k), for the
    latency-bound small-batch regime.
    """

    def __init__(self, local_corpus, global_rows: int, group=None, device: int | None = None,
                 peer_exchange: bool | None = None, peer_slot_bytes: int = 1 << 20):
        import torch.distributed as dist

        self._dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        lo, hi = shard_bounds(global_rows, self.world, self.rank)
        n_local = local_corpus.shape[0]
        if n_local != hi - lo:
            raise ValueError(f"rank {self.rank}: local shard has {n_local} rows, expected {hi - lo}")
        self.index = CorpusIndex(local_corpus, id_offset=lo, device=device)
        self._bufs = {}    # (Q, k) -> (local record, gathered records, out scores, out ids)
        self._graphs = {}  # (Q, k) -> (graph, static queries)
        self._exch = None  # (own buffer address, imported peer addresses, device array of the G addresses)
        self.peer_slot_bytes = int(peer_slot_bytes) // 8 * 8
        if peer_exchange is None:
            peer_exchange = os.environ.get("ARB_PEER_EXCHANGE", "1") != "0"
        if self.world > 1 and peer_exchange:
            self._setup_peer_exchange()

    def _setup_peer_exchange(self):
        """Allocate this rank's exchange buffer, swap IPC handles, map the peers' buffers. Every
        rank must end up with the same decision, so failures are agreed on with an all_reduce."""
        import ctypes
        import socket

        torch = self.index._torch
        dist = self._dist
        lib = _lib.lib()
        dev = self.index.corpus.device
        ok, own, handle = True, ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        with torch.cuda.device(dev):
            try:
                if self.world > 16:
                    raise RuntimeError("more than 16 ranks")
                nbytes = int(lib.arb_topk_exchange_bytes(self.world, self.peer_slot_bytes))
                _lib.check(lib.arb_exchange_alloc(nbytes, ctypes.byref(own)))
                _lib.check(lib.arb_ipc_export(own, handle))
            except Exception:  # noqa: BLE001 - any failure means "use the all_gather path"
                ok = False
            infos = [None] * self.world
            dist.all_gather_object(infos, (socket.gethostname(), bytes(handle), ok), group=self.group)
            ok = ok and all(i[2] for i in infos) and len({i[0] for i in infos}) == 1
            peers, addrs = [], []
            if ok:
                for r, (_, h, _) in enumerate(infos):
                    if r == self.rank:
                        addrs.append(own.value)
                        continue
                    p = ctypes.c_void_p()
                    try:
                        _lib.check(lib.arb_ipc_import((ctypes.c_ubyte * 64).from_buffer_copy(h), ctypes.byref(p)))
                    except Exception:  # noqa: BLE001
                        ok = False
                        break
                    peers.append(p.value)
                    addrs.append(p.value)
            flag = torch.tensor([1 if ok else 0], device=dev, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if int(flag.item()) == 1:
                table = torch.tensor(addrs, dtype=torch.int64).to(dev)
                self._exch = (own.value, peers, table)
            else:
                for p in peers:
                    lib.arb_ipc_close(p)
                if own.value:
                    lib.arb_exchange_free(own)
            dist.barrier(group=self.group)

    def check_exchange(self) -> None:
        """Raise if the peer-memory exchange lost a rank (its bounded wait timed out) since the index
        was built. Synchronises the device; call it after a search whose result looks empty (-1 ids)."""
        if self._exch is not None:
            with self.index._torch.cuda.device(self.index.corpus.device):
                _lib.check(_lib.lib().arb_topk_exchange_status(self._exch[0]))

    @property
    def exchange(self) -> str:
        return "peer-memory kernel (CUDA IPC over NVLink)" if self._exch is not None else "nccl all_gather"

    def _new_buffers(self, Q: int, k: int):
        torch = self.index._torch
        dev = self.index.corpus.device
        rec = int(_lib.lib().arb_topk_record_bytes(Q, k))
        return (torch.empty(rec, dtype=torch.uint8, device=dev),
                torch.empty(self.world * rec, dtype=torch.uint8, device=dev),
                torch.empty((Q, k), dtype=torch.float32, device=dev),
                torch.empty((Q, k), dtype=torch.int64, device=dev))

    def _buffers(self, Q: int, k: int):
        key = (Q, k)
        if key not in self._bufs:
            self._bufs[key] = self._new_buffers(Q, k)
        return self._bufs[key]

    def _search_into(self, queries, k: int, bufs, workspace=None):
        torch = self.index._torch
        local, gathered, out_s, out_i = bufs
        Q = queries.shape[0]
        ids_off = int(_lib.lib().arb_topk_record_ids_offset(Q, k))
        ls = local[:Q * k * 4].view(torch.float32).view(Q, k)
        li = local[ids_off:ids_off + Q * k * 8].view(torch.int64).view(Q, k)
        self.index.search(queries, k, out_scores=ls, out_ids=li, workspace=workspace)
        if self._exch is not None and local.numel() <= self.peer_slot_bytes:
            with torch.cuda.device(local.device):
                _lib.check(_lib.lib().arb_topk_exchange_merge(_lib.ptr(local), _lib.ptr(self._exch[2]), self.rank, self.world,
                                                              Q, k, self.peer_slot_bytes, _lib.ptr(out_s), _lib.ptr(out_i),
                                                              _lib.current_stream()))
            return out_s, out_i
        self._dist.all_gather_into_tensor(gathered, local, group=self.group)
        with torch.cuda.device(gathered.device):
            _lib.check(_lib.lib().arb_topk_merge_records(_lib.ptr(gathered), self.world, Q, k, _lib.ptr(out_s),
                                                         _lib.ptr(out_i), _lib.current_stream()))
        return out_s, out_i

    def search(self, queries, k: int = 10):
        """-> (scores `[Q,k]` float32, ids `[Q,k]` int64), identical on every rank. The returned
        tensors are reused by the next call with the same (Q, k); clone them to keep them."""
        if self.world == 1:
            return self.index.search(queries, k)
        Q = queries.shape[0]
        if Q == 0:
            return self.index.search(queries, k)
        return self._search_into(queries, k, self._buffers(Q, k))

    def search_graphed(self, queries, k: int = 10):
        """`search` replayed from a CUDA graph (captured on first use per (Q, k); the collective is
        captured with it). `queries` must be a CUDA tensor of the corpus dtype; every rank must call
        it with the same shapes in the same order."""
        torch = self.index._torch
        Q = queries.shape[0]
        key = (Q, k)
        if key not in self._graphs:
            static_q = torch.empty_like(queries, dtype=self.index.corpus.dtype, device=self.index.corpus.device)
            static_q.copy_(queries)
            # every captured graph owns its workspace and result buffers: the eager path's shared
            # workspace may be reallocated by a later, larger call while this graph is still replayed
            ws = self.index.new_workspace(Q, k)
            bufs = self._new_buffers(Q, k) if self.world > 1 else None
            outs = None if self.world > 1 else (torch.empty((Q, k), dtype=torch.float32, device=static_q.device),
                                                torch.empty((Q, k), dtype=torch.int64, device=static_q.device))

            def run():
                if self.world > 1:
                    return self._search_into(static_q, k, bufs, workspace=ws)
                return self.index.search(static_q, k, out_scores=outs[0], out_ids=outs[1], workspace=ws)

            side = torch.cuda.Stream(device=static_q.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                run()  # warm-up outside the capture (lazy NCCL init, function attributes)
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                res = run()
            self._graphs[key] = (g, static_q, res, ws, bufs)  # ws / bufs: kept alive for the graph
        g, static_q, res = self._graphs[key][:3]
        static_q.copy_(queries, non_blocking=True)
        g.replay()
        return res

    def close(self):
        """Drop the captured graphs and exchange buffers. Call it (on every rank) before
        `destroy_process_group`: a live graph that holds a captured collective keeps the
        communicator busy at teardown."""
        torch = self.index._torch
        torch.cuda.synchronize(self.index.corpus.device)
        self._graphs.clear()
        self._bufs.clear()
        if self._exch is not None:
            own, peers, _ = self._exch
            self._exch = None
            self._dist.barrier(group=self.group)  # nobody is still storing into a buffer about to go away
            with torch.cuda.device(self.index.corpus.device):
                for p in peers:
                    _lib.lib().arb_ipc_close(p)
                _lib.lib().arb_exchange_free(own)
