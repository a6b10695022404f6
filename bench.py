#!/usr/bin/env python
"""Headline benchmark of the B200 retrieval hot path (contract: task brief §④ / "How to work").

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): all-mpnet-base-v2 encode of synthetic chunks, seq 384,
batch 1024 per GPU per step, 16-bit tensor-core operands / fp32 accumulate, seeded synthetic
weights (no checkpoint offline). One step = one batch of 1024 chunks through the whole hot path
(embedding -> 12 layers -> masked mean-pool -> L2 norm). Data parallel: every rank encodes its own
batches, no collective on the data path ("scaling": "weak").

The operands are fp16, the shipped default: BASELINE names bf16, fp16 has the same width and
tensor-core rate and is the format that meets the north-star's cosine >= 0.9999 on every row
(DESIGN.md 'Numerics'); the bf16 mode is timed beside it (`encode_bf16`).

Rank 0 prints ONE JSON line. `value` = chunks/s with token ids already resident in HBM; `e2e` =
the same metric through the public API (`B200SentenceEncoder.encode`) from HOST numpy ids to HOST
numpy embeddings, copies inside the timed region. `roofline` is the tensor-pipe roofline of the
dominant kernel (the tcgen05 GEMM), timed live with CUDA events. Further records in the same line:
`e2e_strings` (List[str] of ragged U[16,384]-token chunks through the native tokenizer, beside the
same batches resident in HBM), `encode_s256` (the metric's own 256-token shape), `encode_bf16`, `search` (queries/s exact top-10
over a 5M x 768 bf16 corpus, row-sharded over the ranks, with its roofline and a NumPy CPU
baseline), `configs` (BASELINE configs[2], [3] and a 3-point configs[4] latency sweep), and
`cpu_baseline` (the reference's encode loop on the host cores).

`--impl reference` times the reference's CPU implementation of the path — the restatement of
generate_embeddings_parallel.py:131-269 over transformers.MPNetModel (oracle/refpath.py; the
reference file itself cannot be imported) — on a bounded sample per step, both
single-process with every torch thread and through the reference's own process pool (75 % of the
cores, one model per worker).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEQ = 384
BATCH = 1024
HEADLINE_DTYPE = "fp16"
GEMM_SHAPES = [(2304, 768, 0), (768, 768, 2), (3072, 768, 1), (768, 3072, 2)]  # (N, K, epilogue) per layer
SEARCH_N, SEARCH_D, SEARCH_K, SEARCH_Q = 5_000_000, 768, 10, 4096
SEARCH_Q_SMALL = 64


def gflop_per_chunk(seq: int) -> float:
    """12 S (14 155 776 + 3 072 S) FLOP — 70.67 GFLOP at S=384, 45.90 at S=256."""
    return 12 * seq * (14155776 + 3072 * seq) / 1e9


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update({k: m[k] for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in m})
        p["source"] = "measured"
    except Exception:
        pass
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.stop, self.t = index, [], threading.Event(), None

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) > 3 + j and r[3 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "power_w_max": max((float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()), default=None),
                "reasons": reasons, "samples": len(self.rows)}


def ncu_traffic(kernel: str, count: int, skip: int = 0):
    """DRAM bytes (read + write) of `count` consecutive launches of `kernel` from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json, written by tools/ncu_summary.py); None if absent."""
    try:
        rows = [r for r in json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["launches"] if kernel in r["kernel"]]
        rows = rows[skip:skip + count]
        return sum(r["dram_bytes"] for r in rows) if len(rows) == count else None
    except (OSError, KeyError, ValueError):
        return None


# ====================================================================================== CPU arms
def cpu_encode_sample(n_target_s: float = 12.0, seq: int = SEQ):
    """Oracle encode (transformers.MPNetModel fp32 + pooling) on the host cores, single process with
    every torch thread; bounded sample."""
    import torch

    from arxiv_rag_b200.weights import ALL_MPNET_BASE_V2, synthetic_state_dict
    from oracle import encode_oracle as eo
    from oracle import refpath

    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is entitled to every host core
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    model = refpath.OracleSentenceTransformer(ALL_MPNET_BASE_V2, synthetic_state_dict(ALL_MPNET_BASE_V2, 0))
    ids, mask = eo.synthetic_tokens(64, seq, seed=1, full_length=True)
    t0 = time.perf_counter()
    refpath.generate_embeddings_parallel(ids[:4], mask[:4], model, batch_size=4, chunks_per_worker=500)
    per = (time.perf_counter() - t0) / 4
    n = int(max(4, min(64, n_target_s / max(per, 1e-3))))
    t0 = time.perf_counter()
    refpath.generate_embeddings_parallel(ids[:n], mask[:n], model, batch_size=200, chunks_per_worker=500)
    dt = time.perf_counter() - t0
    return n / dt, n, torch.get_num_threads(), model, (ids, mask)


def cpu_search_baseline():
    """NumPy fp32 `Q @ C.T` + argpartition + stable (score desc, id asc) sort,
    1 k queries x 10 k rows x 768, top-10, in full (configs[0]'s search half), on the host cores."""
    from oracle import search_oracle as so

    Q, N, D, k = 1000, 10_000, 768, 10
    c = so.synthetic_unit_rows(N, D, seed=0)
    q = so.synthetic_unit_rows(Q, D, seed=1)
    so.oracle_search(q[:8], c, k)  # warm the BLAS threads
    reps, t0 = 0, time.perf_counter()
    while reps < 3 or time.perf_counter() - t0 < 1.0:
        so.oracle_search(q, c, k)
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    try:
        from threadpoolctl import threadpool_info

        threads = max((p.get("num_threads", 1) for p in threadpool_info()), default=1)
    except Exception:
        threads = os.cpu_count() or 1
    return {"value": Q / dt, "unit": "queries/s", "cores": int(threads), "kind": "port",
            "sample": f"oracle/search_oracle.py (NumPy fp32 Q@C.T + argpartition + stable sort), {Q} queries x {N} rows x {D}, "
                      f"top-{k}, whole config, {reps} repetitions, {dt * 1e3:.1f} ms each; the GPU lines are over a 5M-row corpus "
                      f"(500x the rows per query)"}


def run_reference(args):
    """--impl reference: the reference's CPU path (restated), bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    from arxiv_rag_b200.weights import ALL_MPNET_BASE_V2
    from oracle import refpath

    rate0, n0, threads, model, (ids, mask) = cpu_encode_sample(3.0)
    n = int(max(2, min(64, 3.0 * rate0)))  # ~3 s of CPU work per step
    for _ in range(args.warmup):
        refpath.generate_embeddings_parallel(ids[:2], mask[:2], model, batch_size=200, chunks_per_worker=500)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        refpath.generate_embeddings_parallel(ids[:n], mask[:n], model, batch_size=200, chunks_per_worker=500)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    sample = (f"{n} synthetic {SEQ}-token chunks per step through the restated generate_embeddings_worker/"
              f"_parallel loop (batch_size 200, chunks_per_worker 500) over transformers.MPNetModel fp32, "
              f"{threads} torch threads")
    # The reference's own arrangement (:190, :205): a spawn Pool of 75 % of the cores, one model per
    # worker. chunks_per_worker is cut so that a bounded sample gives every worker one task.
    pool_rec = None
    try:
        workers = refpath.default_pool_workers()
        per_worker = max(2, min(8, int(round(6.0 * rate0 / workers)) or 2))
        n_pool = workers * per_worker
        import numpy as _np

        reps = -(-n_pool // ids.shape[0])
        p_ids, p_mask = _np.tile(ids, (reps, 1))[:n_pool], _np.tile(mask, (reps, 1))[:n_pool]
        pool = refpath.ReferencePool(ALL_MPNET_BASE_V2, 0, workers)
        pool.generate_embeddings_parallel(p_ids[:workers], p_mask[:workers], batch_size=200, chunks_per_worker=1)  # models built
        t1 = time.perf_counter()
        pool.generate_embeddings_parallel(p_ids, p_mask, batch_size=200, chunks_per_worker=per_worker)
        dtp = time.perf_counter() - t1
        pool.close()
        pool_rec = {"value": n_pool / dtp, "unit": "chunks/s", "workers": workers, "cores": os.cpu_count(), "kind": "port",
                    "sample": f"{n_pool} chunks, {workers} spawn workers (75 % of {os.cpu_count()} cores, "
                              f"generate_embeddings_parallel.py:190,205), one model per worker, {per_worker} chunks per task, "
                              f"torch threads per worker left at the default like the reference"}
    except Exception as e:  # noqa: BLE001 - the pool variant is a second opinion, never fatal
        pool_rec = {"unavailable": f"{type(e).__name__}: {e}"}
    return ({
        "impl": "reference", "metric": "chunks/sec encoded (MPNet)", "value": val, "unit": "chunks/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"all-mpnet-base-v2 encode, seq {SEQ}, batch {BATCH} (configs[1]); CPU sample of {n} chunks/step"},
        "cpu_baseline": {"value": val, "unit": "chunks/s", "cores": threads, "kind": "port", "sample": sample},
        "process_pool": pool_rec,
        "search_cpu_baseline": cpu_search_baseline(),
        "e2e": {"value": val, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ====================================================================================== B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    from arxiv_rag_b200 import _lib
    from arxiv_rag_b200.encoder import B200SentenceEncoder
    from arxiv_rag_b200.search import CorpusIndex, ShardedCorpusIndex, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()  # fails loudly if the CUDA library is missing
    pk = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def unit_rows(n, seed, dtype):
        g = torch.Generator(device=dev).manual_seed(seed)
        out = torch.empty((n, SEARCH_D), device=dev, dtype=dtype)
        for s in range(0, n, 500_000):  # generate in slabs: no fp32 copy of a whole shard
            e = min(s + 500_000, n)
            out[s:e] = torch.nn.functional.normalize(torch.randn(e - s, SEARCH_D, device=dev, generator=g), dim=1).to(dtype)
        return out

    def timed_encode(enc, seq, steps, warmup, sample_clocks=False):
        g = torch.Generator(device=dev).manual_seed(1 + rank + seq)
        nbuf = 4  # rotate input batches; a step's activations (GBs) are far larger than the 126 MB L2
        ids = torch.randint(4, 30525, (nbuf, BATCH, seq), device=dev, dtype=torch.int32, generator=g)
        ids[:, :, 0] = 0
        ids[:, :, -1] = 2
        mask = torch.ones((BATCH, seq), device=dev, dtype=torch.int32)
        out = torch.empty((BATCH, 768), device=dev, dtype=torch.float32)
        for i in range(warmup):
            enc.encode_tokens(ids[i % nbuf], mask, out)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clk = ClockSampler(local) if sample_clocks else None
        if clk:
            clk.__enter__()
        e0.record()
        for i in range(steps):
            enc.encode_tokens(ids[i % nbuf], mask, out)
        e1.record()
        barrier()
        if clk:
            clk.__exit__()
        ms = max_over_ranks(e0.elapsed_time(e1)) / steps
        enc.check_status()
        return ms, ids, out, (clk.summary() if clk else None)

    # ------------------------------------------------------------------ encode (headline)
    enc = B200SentenceEncoder(None, max_batch=BATCH, max_seq=SEQ, dtype=HEADLINE_DTYPE, seed=0)
    ms_step, ids, out, clocks = timed_encode(enc, SEQ, args.steps, args.warmup, sample_clocks=True)
    value = world * BATCH / (ms_step / 1e3)
    checksum = float(out.float().norm(dim=1).mean().item())  # ~1.0: unit-norm rows came out
    launches = args.steps * enc.launches_per_encode
    side_steps = max(3, min(args.steps, 10))

    # the metric's own shape: 256-token chunks (BASELINE.json `metric`, configs[0])
    ms256, _, _, _ = timed_encode(enc, 256, side_steps, 3)
    v256 = world * BATCH / (ms256 / 1e3)
    encode_s256 = {"metric": "chunks/sec encoded (MPNet, 256 tok)", "value": v256, "unit": "chunks/s", "ms_per_step": ms256,
                   "steps": side_steps, "seq_len": 256, "batch_per_gpu": BATCH, "dtype": HEADLINE_DTYPE,
                   "gflop_per_chunk": gflop_per_chunk(256),
                   "step_achieved_tflops": v256 / world * gflop_per_chunk(256) / 1e3,
                   "step_frac": v256 / world * gflop_per_chunk(256) / 1e3 / pk["bf16_tflops_sustained"]}

    # ------------------------------------------------------------------ e2e through the public API (host -> host)
    h_ids = ids[0].cpu().numpy()
    h_mask = np.ones((BATCH, SEQ), np.int32)
    enc.encode((h_ids, h_mask), batch_size=BATCH, normalize_embeddings=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(side_steps):
        enc.encode((h_ids, h_mask), batch_size=BATCH, normalize_embeddings=True, convert_to_numpy=True)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / side_steps
    e2e = {"value": world * BATCH / (e2e_ms / 1e3), "unit": "chunks/s", "h2d_bytes_per_step": int(2 * BATCH * SEQ * 4),
           "d2h_bytes_per_step": int(BATCH * 768 * 4), "ms_per_step": e2e_ms,
           "api": "B200SentenceEncoder.encode((ids, mask) numpy, batch_size=1024) -> numpy float32 [1024,768]"}
    # ------------------------------------------------------------------ e2e from List[str], ragged lengths
    # The call the reference makes (generate_embeddings_parallel.py:146-153): strings in, float32 rows
    # out. Chunks of U[16,384] tokens over a synthetic vocabulary (none ships offline), tokenised by the
    # in-tree native tokenizer on a background thread, sorted by length, padded per batch. Beside it the
    # same batches pre-tokenised and resident in HBM (what the GPU could do if the host cost nothing).
    from arxiv_rag_b200.tokenizer import NativeWordPieceTokenizer
    rs = np.random.RandomState(11 + rank)
    letters = np.array(list("abcdefghijklmnopqrstuvwxyz"))
    wset = {"".join(rs.choice(letters, rs.randint(2, 10))) for _ in range(30000)}
    words = sorted(wset)
    vocab = {t: i for i, t in enumerate(["<s>", "<pad>", "</s>", "[UNK]"] + words + list(".,;:()") +
                                        ["##" + c for c in letters] + list(letters))}
    enc.tokenizer = NativeWordPieceTokenizer(vocab, kind="mpnet", max_length=SEQ)
    n_text = 8 * BATCH
    wi = rs.randint(0, len(words), size=(n_text, SEQ))
    n_tok = rs.randint(16, SEQ + 1, size=n_text)
    texts = [" ".join(words[j] for j in wi[r, :n_tok[r] - 2]) for r in range(n_text)]
    enc.encode(texts[:2 * BATCH], batch_size=BATCH)
    barrier()
    str_reps = 3
    t0 = time.perf_counter()
    for _ in range(str_reps):
        enc.encode(texts, batch_size=BATCH, convert_to_numpy=True)
    torch.cuda.synchronize()
    str_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / str_reps
    t_ids, t_mask = enc.tokenizer.tokenize_batch(texts)
    t_len = t_mask.sum(1)
    order = np.argsort(-np.array([len(t) for t in texts]), kind="stable")
    dev_batches = []
    for b0 in range(0, n_text, BATCH):
        sel = order[b0:b0 + BATCH]
        S = int(t_len[sel].max())
        dev_batches.append((torch.from_numpy(np.ascontiguousarray(t_ids[sel, :S])).to(dev),
                            torch.from_numpy(np.ascontiguousarray(t_mask[sel, :S])).to(dev)))
    for d_i, d_m in dev_batches[:2]:
        enc.encode_tokens(d_i, d_m, out)
    barrier()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for _ in range(str_reps):
        for d_i, d_m in dev_batches:
            enc.encode_tokens(d_i, d_m, out)
    r1.record()
    barrier()
    res_ms = max_over_ranks(r0.elapsed_time(r1)) / str_reps
    t0 = time.perf_counter()
    enc.tokenizer.tokenize_batch(texts)
    tok_s = time.perf_counter() - t0
    e2e_strings = {"value": world * n_text / (str_ms / 1e3), "unit": "chunks/s", "ms_per_pass": str_ms,
                   "device_resident": world * n_text / (res_ms / 1e3), "ratio": res_ms / str_ms,
                   "chunks_per_pass_per_gpu": n_text, "tokens_mean": float(t_len.mean()),
                   "tokenizer_chunks_per_s": n_text / tok_s, "tokenizer_threads": enc.tokenizer.num_threads,
                   "h2d_bytes_per_pass": int(sum(2 * a.numel() * 4 for a, _ in dev_batches)),
                   "d2h_bytes_per_pass": int(n_text * 768 * 4),
                   "api": "B200SentenceEncoder.encode(List[str] of U[16,384]-token chunks, batch_size=1024) -> numpy float32; "
                          "native WordPiece tokenizer (arb_tokenizer_encode), synthetic 30k vocabulary"}
    del dev_batches, texts
    enc.close()
    del enc, ids
    torch.cuda.empty_cache()

    # what BASELINE configs[1] names literally: bf16 operands (same kernels, 8-bit mantissa)
    enc_b = B200SentenceEncoder(None, max_batch=BATCH, max_seq=SEQ, dtype="bf16", seed=0)
    ms_b, _, _, _ = timed_encode(enc_b, SEQ, side_steps, 3)
    encode_bf16 = {"value": world * BATCH / (ms_b / 1e3), "unit": "chunks/s", "ms_per_step": ms_b, "steps": side_steps,
                   "note": "dtype='bf16' handle, same shape; parity budget in DESIGN.md 'Numerics'"}
    enc_b.close()
    del enc_b
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ dominant kernel: tcgen05 GEMM, timed alone
    M = BATCH * SEQ
    gemm = []
    A768 = torch.randn(M, 768, device=dev).to(torch.float16)
    A3072 = torch.randn(M, 3072, device=dev).to(torch.float16)
    Cbuf = torch.empty(M, 3072, device=dev, dtype=torch.float16)
    Rbuf = torch.randn(M, 768, device=dev).to(torch.float16)
    # the epilogues the encode path launches: LayerNorm folded in (kernels.h EPI_LNIN_* / EPI_*_STATS)
    parts = 768 // 128
    stats_in = torch.rand(parts, M, 2, device=dev) * 50 + 100
    stats_out = torch.empty(parts, M, 2, device=dev)
    ln_g, ln_b = torch.ones(768, device=dev), torch.zeros(768, device=dev)
    for (N, K, epi) in GEMM_SHAPES:
        A = A768 if K == 768 else A3072
        W = (torch.randn(N, K, device=dev) * 0.04).to(torch.float16)
        bias = torch.randn(N, device=dev)
        colsum = torch.randn(N, device=dev)
        fold_epi = {0: 3, 1: 4, 2: 5}[epi]  # bias -> LN-in bias; gelu -> LN-in gelu; residual -> LN(residual) + row stats
        call = lambda: _lib.check(lib.arb_gemm16_lnfold(A.data_ptr(), K, W.data_ptr(), K, Cbuf.data_ptr(), N, bias.data_ptr(),
                                                        Rbuf.data_ptr() if epi == 2 else 0, 768, colsum.data_ptr(), ln_g.data_ptr(),
                                                        ln_b.data_ptr(), stats_in.data_ptr(), parts, 768,
                                                        stats_out.data_ptr() if epi == 2 else 0, 1e-5, M, N, K, fold_epi,
                                                        _lib.ARB_DTYPE_F16, torch.cuda.current_stream().cuda_stream))
        # "timed alone" against the BURST peak (best of 10 launches from an idle GPU, MEASURED_PEAKS.json
        # `how`): let the clocks recover from the power-capped encode phase first, then time each of 10
        # launches with its own event pair; `ms` = their mean (what `achieved` uses), `ms_best` the best
        torch.cuda.synchronize()
        time.sleep(0.5)
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        reps = 10
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in evs:
            a.record()
            call()
            b.record()
        torch.cuda.synchronize()
        each = [a.elapsed_time(b) for a, b in evs]
        ms = sum(each) / reps
        # the library GEMM on the SAME shape, no epilogue at all (torch -> cuBLASLt), timed the same way:
        # the burst peak is an 8192^3 figure; at K = 768 a tile's main loop is short for any kernel
        lib_out = Cbuf.view(-1)[:M * N].view(M, N)
        torch.cuda.synchronize()
        time.sleep(0.5)
        for _ in range(3):
            torch.mm(A, W.t(), out=lib_out)
        torch.cuda.synchronize()
        for a, b in evs:
            a.record()
            torch.mm(A, W.t(), out=lib_out)
            b.record()
        torch.cuda.synchronize()
        lib_ms = sum(a.elapsed_time(b) for a, b in evs) / reps
        gemm.append({"N": N, "K": K, "epilogue": fold_epi, "ms": ms, "ms_best": min(each), "tflops": 2.0 * M * N * K / ms / 1e9,
                     "cublas_no_epilogue_ms": lib_ms, "cublas_no_epilogue_tflops": 2.0 * M * N * K / lib_ms / 1e9})
    del A768, A3072, Cbuf, Rbuf, stats_in, stats_out
    gemm_flops = sum(2.0 * M * s["N"] * s["K"] for s in gemm)
    gemm_ms = sum(s["ms"] for s in gemm)
    gemm_ach = gemm_flops / gemm_ms / 1e9
    gfc = gflop_per_chunk(SEQ)
    roofline = {
        "bound": "tensor", "kernel": "gemm16_kernel (tcgen05.mma cta_group::2, LayerNorm-folding epilogues; 4 launches per layer)",
        "achieved": gemm_ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": gemm_ach / pk["bf16_tflops"],
        "peak_source": f"{pk['source']} burst (kernel timed alone: 0.5 s idle, 3 warm-ups, mean of 10 launches timed one by one); fp16 and bf16 share the tensor-core rate",
        "traffic": ncu_traffic("gemm16_kernel", 4), "traffic_unit": "DRAM bytes for the 4 launches of one layer (ncu --set full, profiles/ncu_traffic.json)",
        "algorithmic_bytes": sum(2.0 * M * (s["K"] + s["N"] * (2 if s["epilogue"] == 5 else 1)) + 2.0 * s["N"] * s["K"] for s in gemm),
        "per_shape": gemm,
        "schedule": "CTA pairs: tcgen05 cta_group::2, 256x256 tiles, clusters of 2",
        "step_achieved": value / world * gfc / 1e3, "step_peak": pk["bf16_tflops_sustained"],
        "step_frac": value / world * gfc / 1e3 / pk["bf16_tflops_sustained"],
        "gemm_share_of_step": 12 * gemm_ms / ms_step,
        "library_same_shapes": {"achieved": gemm_flops / sum(x["cublas_no_epilogue_ms"] for x in gemm) / 1e9, "unit": "TFLOP/s",
                                "what": "torch.mm (cuBLASLt) fp16 on the same four shapes with NO epilogue, same timing protocol"},
    }

    # ------------------------------------------------------------------ search (second headline)
    lo, hi = shard_bounds(SEARCH_N, world, rank)
    corpus = unit_rows(hi - lo, 100 + rank, torch.bfloat16)
    sharded = ShardedCorpusIndex(corpus, SEARCH_N) if world > 1 else None
    index = sharded.index if sharded is not None else CorpusIndex(corpus, id_offset=lo)
    gq = torch.Generator(device=dev).manual_seed(7)  # same queries on every rank
    search = {}

    def timed(fn, reps, warm_ms=250.0):
        # warm up for >= 3 calls and ~warm_ms: the clocks need that long to settle after the
        # power-capped encode phase, and the millisecond-scale HBM-bound step is clock sensitive.
        # The count comes from an all-reduced probe so every rank makes the same number of calls.
        fn()  # first call may allocate / capture a graph
        torch.cuda.synchronize()
        t_w = time.perf_counter()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        per_call_ms = max_over_ranks((time.perf_counter() - t_w) * 1e3 / 3)
        for _ in range(int(min(20000, max(0, warm_ms / max(per_call_ms, 1e-3))))):
            fn()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(reps):
            out = fn()
        s1.record()
        barrier()
        return max_over_ranks(s0.elapsed_time(s1)) / reps, out

    def search_roofline(Q, n_rows, k, ms, elt_bytes=2):
        nbytes = n_rows * SEARCH_D * elt_bytes + Q * SEARCH_D * elt_bytes + Q * k * 12
        flops = 2.0 * Q * n_rows * SEARCH_D
        t_hbm = nbytes / (pk["hbm_gbs"] * 1e9)
        t_mma = flops / (pk["bf16_tflops_sustained"] * 1e12)
        bound = "hbm" if t_hbm >= t_mma else "tensor"
        return {"bound": bound, "achieved": (nbytes / ms / 1e6) if bound == "hbm" else (flops / ms / 1e9),
                "peak": pk["hbm_gbs"] if bound == "hbm" else pk["bf16_tflops_sustained"],
                "unit": "GB/s" if bound == "hbm" else "TFLOP/s", "frac": max(t_hbm, t_mma) / (ms / 1e3),
                "algorithmic_bytes": nbytes, "algorithmic_flops": flops}

    for label, Q in (("large_batch", SEARCH_Q), ("small_batch", SEARCH_Q_SMALL)):
        q = torch.nn.functional.normalize(torch.randn(Q, SEARCH_D, device=dev, generator=gq), dim=1).to(torch.bfloat16)
        graphed = sharded is not None and Q <= 256  # latency-bound regime: replay search + exchange + merge from a CUDA graph
        if sharded is None:
            step = lambda: index.search(q, SEARCH_K)
        elif graphed:
            step = lambda: sharded.search_graphed(q, SEARCH_K)
        else:
            step = lambda: sharded.search(q, SEARCH_K)
        reps = max(5, min(args.steps, 20)) if Q > 256 else 100
        # the sub-millisecond HBM-bound step follows the SM/L2 clock, which takes ~1 s to recover from
        # the power-capped encode phase; the tensor-bound large batch is itself power-capped
        warm = 1500.0 if Q <= 256 else 250.0
        with ClockSampler(local) as sclk:
            ms, (fs, fi) = timed(step, reps if Q > 256 else 400, warm)
        # the rank-local part alone (fused score+top-k kernel and its split merge; no collective)
        local_ms, _ = timed(lambda: index.search(q, SEARCH_K), reps, warm) if sharded is not None else (ms, None)
        rl = search_roofline(Q, hi - lo, SEARCH_K, ms)
        rl["frac_local"] = rl["frac"] * ms / local_ms
        # ncu capture is of one launch over the full 5M-row corpus on one GPU
        rl["traffic"] = ncu_traffic("search_topk_kernel", 1, skip=0 if Q > 256 else 1) if world == 1 else None
        search[label] = {
            "metric": f"queries/sec exact top-{SEARCH_K} @ {SEARCH_N}x{SEARCH_D} bf16", "Q": Q, "value": Q / (ms / 1e3),
            "unit": "queries/s", "ms_per_batch": ms, "local_ms_per_batch": local_ms,
            "path": ("cuda graph: " if graphed else "") + (f"local search -> [Q,k] records exchanged + merged by {sharded.exchange}" if sharded is not None else "local search"),
            "roofline": rl, "top1_score_mean": float(fs[:, 0].mean().item()), "clocks": sclk.summary(),
        }
    search["oracle_note"] = ("results are checked against oracle/search_oracle.py (NumPy fp32) in tests/; the reference has no "
                             "search routine to pin that oracle to (SURVEY.md F3/F4): parity unpinned by the reference")

    # ------------------------------------------------------------------ BASELINE configs[2], [3], [4]
    configs = {}
    # configs[3]: exact top-100, 100 k queries over the same 5M x 768 bf16 corpus, row-sharded over the ranks
    Q3, K3, B3 = 100_000, 100, 32768
    q3 = unit_rows(Q3, 11, torch.bfloat16)

    def run_cfg3():
        last = None
        for s in range(0, Q3, B3):
            qb = q3[s:s + B3]
            last = sharded.search(qb, K3) if sharded is not None else index.search(qb, K3)
        return last

    run_cfg3()
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    run_cfg3()
    c1.record()
    barrier()
    ms3 = max_over_ranks(c0.elapsed_time(c1))
    rl3 = search_roofline(Q3, hi - lo, K3, ms3)
    configs["cfg3_top100_100k_x_5M_bf16"] = {"value": Q3 / (ms3 / 1e3), "unit": "queries/s", "ms": ms3, "Q": Q3, "N": SEARCH_N, "k": K3,
                                             "query_batch": B3, "n_gpus": world, "roofline": rl3,
                                             "note": "BASELINE names 2/4/8 GPUs; at N=1 the whole corpus sits on one GPU"}
    del q3
    if sharded is not None:
        sharded.close()
    del index, sharded, corpus
    torch.cuda.empty_cache()

    # configs[2]: exact top-10, 10 k queries over a 1M x 768 fp32 corpus on ONE GPU (every rank runs its own replica)
    Q2, N2 = 10_000, 1_000_000
    idx2 = CorpusIndex(unit_rows(N2, 100, torch.float32))
    q2 = unit_rows(Q2, 7, torch.float32)
    ms2, _ = timed(lambda: idx2.search(q2, 10), 5, 100.0)
    rl2 = search_roofline(Q2, N2, 10, ms2, elt_bytes=4)
    configs["cfg2_top10_10k_x_1M_fp32"] = {"value": Q2 / (ms2 / 1e3), "unit": "queries/s", "ms": ms2, "Q": Q2, "N": N2, "k": 10,
                                           "n_gpus": 1, "replicas": world, "roofline": rl2,
                                           "note": "roofline counts the algorithmic 2*Q*N*768 FLOP against the bf16 sustained peak"}
    del idx2, q2
    torch.cuda.empty_cache()

    # configs[4]: encode-then-search latency, pinned host token ids in -> pinned host top-k out, over a
    # 50M x 768 bf16 corpus on 8 GPUs = 6.25M rows per GPU (fewer ranks keep the 6.25M-row shard)
    rows4, seq4 = 6_250_000, 64
    if world == 1 and not dist.is_initialized():
        import socket

        with socket.socket() as sck:
            sck.bind(("127.0.0.1", 0))
            port = sck.getsockname()[1]
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    enc4 = B200SentenceEncoder(None, max_batch=4096, max_seq=seq4, dtype=HEADLINE_DTYPE, seed=0)
    idx4 = ShardedCorpusIndex(unit_rows(rows4, 200 + rank, torch.bfloat16), rows4 * world)
    sweep = []
    for Q in (1, 64, 4096):
        gq4 = torch.Generator().manual_seed(Q)
        h_ids4 = torch.randint(4, 30000, (Q, seq4), dtype=torch.int32, generator=gq4).pin_memory()
        h_mask4 = torch.ones((Q, seq4), dtype=torch.int32).pin_memory()
        d_ids4 = torch.empty((Q, seq4), device=dev, dtype=torch.int32)
        d_mask4 = torch.empty((Q, seq4), device=dev, dtype=torch.int32)
        h_s4 = torch.empty((Q, 10), dtype=torch.float32).pin_memory()
        h_i4 = torch.empty((Q, 10), dtype=torch.int64).pin_memory()
        q16 = torch.empty((Q, 768), device=dev, dtype=torch.bfloat16)

        def step4():
            d_ids4.copy_(h_ids4, non_blocking=True)
            d_mask4.copy_(h_mask4, non_blocking=True)
            q16.copy_(enc4.encode_tokens_graphed(d_ids4, d_mask4))
            s, i = idx4.search_graphed(q16, 10)
            h_s4.copy_(s, non_blocking=True)
            h_i4.copy_(i, non_blocking=True)

        for _ in range(5):
            step4()
        torch.cuda.synchronize()
        times = []
        for _ in range(15):
            if world > 1:
                dist.barrier()
            t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0e.record()
            step4()
            t1e.record()
            torch.cuda.synchronize()
            times.append(t0e.elapsed_time(t1e))
        med = max_over_ranks(sorted(times)[len(times) // 2])
        sweep.append({"Q": Q, "ms": med, "queries_per_s": Q / med * 1e3})
    configs["cfg4_encode_then_search_latency"] = {
        "corpus_rows": rows4 * world, "rows_per_gpu": rows4, "seq": seq4, "k": 10, "n_gpus": world, "sweep": sweep,
        "shard_hbm_floor_ms": rows4 * 768 * 2 / (pk["hbm_gbs"] * 1e9) * 1e3,
        "path": "pinned host ids -> H2D -> encode (CUDA graph) -> sharded search + exchange + merge (CUDA graph) -> D2H pinned top-k; median of 15"}
    idx4.close()
    enc4.close()
    del idx4, enc4
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ CPU baselines (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, n, threads, _, _ = cpu_encode_sample(12.0)
        cpu = {"value": rate, "unit": "chunks/s", "cores": threads, "kind": "port",
               "sample": f"{n} synthetic {SEQ}-token chunks, restated reference loop (oracle/refpath.py) over "
                         f"transformers.MPNetModel fp32, {threads} torch threads of {os.cpu_count()} host cores"}
        search["cpu_baseline"] = cpu_search_baseline()

    result = None
    if rank == 0:
        result = ({
            "metric": "chunks/sec encoded (MPNet)", "value": value, "unit": "chunks/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": HEADLINE_DTYPE, "data": "synthetic",
            "config": {"workload": "all-mpnet-base-v2 encode, 1024 synthetic 384-token chunks per GPU per step (BASELINE configs[1])",
                       "seq_len": SEQ, "batch_per_gpu": BATCH, "global_batch": BATCH * world, "weights": "seeded synthetic (no checkpoint offline)",
                       "operands": "fp16 x fp16 -> fp32 (BASELINE names bf16: same width and tensor-core rate; fp16 is what holds cosine >= 0.9999 on every row — `encode_bf16` times the bf16 mode)",
                       "parallelism": f"dp{world} (chunk batches sharded, no collective)",
                       "l2": "inputs rotate over 4 batches; per-step activations 6.6 GB >> 126 MB L2"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "e2e_strings": e2e_strings, "encode_s256": encode_s256, "encode_bf16": encode_bf16, "search": search, "configs": configs,
            "unit_norm_check": checksum, "gflop_per_chunk": gfc,
        })
    if dist.is_initialized():
        dist.destroy_process_group()
    return result


class _StdoutGuard:
    """Keep stdout for the ONE JSON line: while the benchmark runs, file descriptor 1 points at
    stderr, so native libraries that print to stdout (NCCL's 'NCCL version ...' banner on rank 0)
    cannot get in front of it."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *a):
        import ctypes

        sys.stdout.flush()
        try:
            ctypes.CDLL(None).fflush(None)  # drain C stdio buffers (NCCL uses printf) into stderr
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "b200":
        args.warmup = max(args.warmup, 3)  # timing rule: at least 3 untimed warm-up steps
    with _StdoutGuard():
        result = run_reference(args) if args.impl == "reference" else run_b200(args)
    if result is not None:  # rank 0 only: the one JSON line, alone on stdout
        print(json.dumps(result), flush=True)


if __name__ == "__main__":
    main()
