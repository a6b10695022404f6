#!/usr/bin/env python
"""Headline benchmark of the B200 retrieval hot path (contract: task brief §④ / "How to work").

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): all-mpnet-base-v2 encode of synthetic chunks, seq 384,
batch 1024 per GPU per step, bf16 tensor-core operands / fp32 accumulate, seeded synthetic
weights (no checkpoint offline). One step = one batch of 1024 chunks through the whole hot path
(embedding -> 12 layers -> masked mean-pool -> L2 norm). Data parallel: every rank encodes its own
batches, no collective on the data path ("scaling": "weak").

Rank 0 prints ONE JSON line. `value` = chunks/s with token ids already resident in HBM; `e2e` =
the same metric through the public API (`B200SentenceEncoder.encode`) from HOST numpy ids to HOST
numpy embeddings, copies inside the timed region. `roofline` is the tensor-pipe roofline of the
dominant kernel (the tcgen05 GEMM), timed live with CUDA events; `search` carries the second
headline (queries/s exact top-10 over a 5M x 768 bf16 corpus, row-sharded over the ranks with an
NCCL all-gather + merge) with its own roofline. `cpu_baseline` is the oracle (the reference's
own dependency, transformers.MPNetModel fp32, + pooling) on the box's host cores.

`--impl reference` times the reference's CPU implementation of the path — the restatement of
generate_embeddings_parallel.py:131-269 over transformers.MPNetModel (oracle/refpath.py; the
reference file itself cannot be imported) — on a bounded sample per step.
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
GFLOP_PER_CHUNK = 12 * SEQ * (14155776 + 3072 * SEQ) / 1e9  # 70.67 @ S=384
GEMM_SHAPES = [(2304, 768, 0), (768, 768, 2), (3072, 768, 1), (768, 3072, 2)]  # (N, K, epilogue) per layer
SEARCH_N, SEARCH_D, SEARCH_K, SEARCH_Q = 5_000_000, 768, 10, 4096
SEARCH_Q_SMALL = 64


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


def cpu_encode_sample(n_target_s: float = 12.0):
    """Oracle encode (transformers.MPNetModel fp32 + pooling) on the host cores; bounded sample."""
    import torch

    from arxiv_rag_b200.weights import ALL_MPNET_BASE_V2, synthetic_state_dict
    from oracle import encode_oracle as eo
    from oracle import refpath

    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is entitled to every host core
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    model = refpath.OracleSentenceTransformer(ALL_MPNET_BASE_V2, synthetic_state_dict(ALL_MPNET_BASE_V2, 0))
    ids, mask = eo.synthetic_tokens(64, SEQ, seed=1, full_length=True)
    t0 = time.perf_counter()
    refpath.generate_embeddings_parallel(ids[:4], mask[:4], model, batch_size=4, chunks_per_worker=500)
    per = (time.perf_counter() - t0) / 4
    n = int(max(4, min(64, n_target_s / max(per, 1e-3))))
    t0 = time.perf_counter()
    refpath.generate_embeddings_parallel(ids[:n], mask[:n], model, batch_size=200, chunks_per_worker=500)
    dt = time.perf_counter() - t0
    return n / dt, n, torch.get_num_threads(), model, (ids, mask)


def run_reference(args):
    """--impl reference: the reference's CPU path (restated), bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    import torch

    rate0, n0, threads, model, (ids, mask) = cpu_encode_sample(3.0)
    from oracle import refpath

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
    return ({
        "impl": "reference", "metric": "chunks/sec encoded (MPNet)", "value": val, "unit": "chunks/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"all-mpnet-base-v2 encode, seq {SEQ}, batch {BATCH} (configs[1]); CPU sample of {n} chunks/step"},
        "cpu_baseline": {"value": val, "unit": "chunks/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def run_b200(args):
    import torch
    import torch.distributed as dist

    from arxiv_rag_b200 import _lib
    from arxiv_rag_b200.encoder import B200SentenceEncoder
    from arxiv_rag_b200.search import CorpusIndex, merge_topk, shard_bounds

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

    # ------------------------------------------------------------------ encode (headline)
    enc = B200SentenceEncoder(None, max_batch=BATCH, max_seq=SEQ, dtype="bf16", seed=0)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    nbuf = 4  # rotate input batches; activations (6.6 GB/step) are far larger than the 126 MB L2
    ids = torch.randint(4, 30525, (nbuf, BATCH, SEQ), device=dev, dtype=torch.int32, generator=g)
    ids[:, :, 0] = 0
    ids[:, :, -1] = 2
    mask = torch.ones((BATCH, SEQ), device=dev, dtype=torch.int32)
    out = torch.empty((BATCH, 768), device=dev, dtype=torch.float32)
    for i in range(args.warmup):
        enc.encode_tokens(ids[i % nbuf], mask, out)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for i in range(args.steps):
            enc.encode_tokens(ids[i % nbuf], mask, out)
        e1.record()
        barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = world * BATCH / (ms_step / 1e3)
    clocks = clk.summary()
    checksum = float(out.float().norm(dim=1).mean().item())  # ~1.0: unit-norm rows came out

    # ------------------------------------------------------------------ dominant kernel: tcgen05 GEMM, timed alone
    M = BATCH * SEQ
    gemm = []
    A768 = torch.randn(M, 768, device=dev).to(torch.bfloat16)
    A3072 = torch.randn(M, 3072, device=dev).to(torch.bfloat16)
    Cbuf = torch.empty(M, 3072, device=dev, dtype=torch.bfloat16)
    Rbuf = torch.randn(M, 768, device=dev).to(torch.bfloat16)
    # the epilogues the encode path launches: LayerNorm folded in (kernels.h EPI_LNIN_* / EPI_*_STATS)
    parts = 768 // 128
    stats_in = torch.rand(parts, M, 2, device=dev) * 50 + 100
    stats_out = torch.empty(parts, M, 2, device=dev)
    ln_g, ln_b = torch.ones(768, device=dev), torch.zeros(768, device=dev)
    for (N, K, epi) in GEMM_SHAPES:
        A = A768 if K == 768 else A3072
        W = (torch.randn(N, K, device=dev) * 0.04).to(torch.float16)  # the encode path's operands: bf16 activations x fp16 weights
        bias = torch.randn(N, device=dev)
        colsum = torch.randn(N, device=dev)
        fold_epi = {0: 3, 1: 4, 2: 5}[epi]  # bias -> LN-in bias; gelu -> LN-in gelu; residual -> LN(residual) + row stats
        call = lambda: _lib.check(lib.arb_gemm16_lnfold(A.data_ptr(), K, W.data_ptr(), K, Cbuf.data_ptr(), N, bias.data_ptr(),
                                                        Rbuf.data_ptr() if epi == 2 else 0, 768, colsum.data_ptr(), ln_g.data_ptr(),
                                                        ln_b.data_ptr(), stats_in.data_ptr(), parts, 768,
                                                        stats_out.data_ptr() if epi == 2 else 0, 1e-5, M, N, K, fold_epi,
                                                        _lib.ARB_DTYPE_BF16_WF16, torch.cuda.current_stream().cuda_stream))
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        g0.record()
        for _ in range(reps):
            call()
        g1.record()
        torch.cuda.synchronize()
        ms = g0.elapsed_time(g1) / reps
        gemm.append({"N": N, "K": K, "epilogue": fold_epi, "ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9})
    del A768, A3072, Cbuf, Rbuf, stats_in, stats_out
    gemm_flops = sum(2.0 * M * s["N"] * s["K"] for s in gemm)
    gemm_ms = sum(s["ms"] for s in gemm)
    gemm_ach = gemm_flops / gemm_ms / 1e9
    roofline = {
        "bound": "tensor", "kernel": "gemm16_kernel (tcgen05.mma cta_group::2, LayerNorm-folding epilogues; 4 launches per layer)",
        "achieved": gemm_ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": gemm_ach / pk["bf16_tflops"],
        "peak_source": f"{pk['source']} burst (kernel timed alone)",
        "traffic": ncu_traffic("gemm16_kernel", 4), "traffic_unit": "DRAM bytes for the 4 launches of one layer (ncu --set full, profiles/ncu_traffic.json)",
        "algorithmic_bytes": sum(2.0 * M * (s["K"] + s["N"] * (2 if s["epilogue"] == 5 else 1)) + 2.0 * s["N"] * s["K"] for s in gemm),
        "per_shape": gemm,
        "schedule": "CTA pairs: tcgen05 cta_group::2, 256x256 tiles, clusters of 2",
        "step_achieved": value / world * GFLOP_PER_CHUNK / 1e3, "step_peak": pk["bf16_tflops_sustained"],
        "step_frac": value / world * GFLOP_PER_CHUNK / 1e3 / pk["bf16_tflops_sustained"],
        "gemm_share_of_step": 12 * gemm_ms / ms_step,
    }

    # ------------------------------------------------------------------ e2e through the public API (host -> host)
    h_ids = ids[0].cpu().numpy()
    h_mask = np.ones((BATCH, SEQ), np.int32)
    enc.encode((h_ids, h_mask), batch_size=BATCH, normalize_embeddings=True)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        emb = enc.encode((h_ids, h_mask), batch_size=BATCH, normalize_embeddings=True, convert_to_numpy=True)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    e2e = {"value": world * BATCH / (e2e_ms / 1e3), "unit": "chunks/s", "h2d_bytes_per_step": int(2 * BATCH * SEQ * 4),
           "d2h_bytes_per_step": int(BATCH * 768 * 4), "ms_per_step": e2e_ms,
           "api": "B200SentenceEncoder.encode((ids, mask) numpy, batch_size=1024) -> numpy float32 [1024,768]"}
    launches = args.steps * enc.launches_per_encode
    enc.close()
    del enc, ids
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ search (second headline)
    lo, hi = shard_bounds(SEARCH_N, world, rank)
    gs = torch.Generator(device=dev).manual_seed(100 + rank)
    corpus = torch.empty((hi - lo, SEARCH_D), device=dev, dtype=torch.bfloat16)
    for s in range(0, hi - lo, 500_000):  # generate in slabs: no fp32 copy of the whole shard
        e = min(s + 500_000, hi - lo)
        corpus[s:e] = torch.nn.functional.normalize(torch.randn(e - s, SEARCH_D, device=dev, generator=gs), dim=1).to(torch.bfloat16)
    from arxiv_rag_b200.search import ShardedCorpusIndex

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
        shard_bytes = (hi - lo) * SEARCH_D * 2 + Q * SEARCH_D * 2 + Q * SEARCH_K * 12
        flops = 2.0 * Q * (hi - lo) * SEARCH_D
        t_hbm = shard_bytes / (pk["hbm_gbs"] * 1e9)
        t_mma = flops / (pk["bf16_tflops_sustained"] * 1e12)
        bound = "hbm" if t_hbm >= t_mma else "tensor"
        search[label] = {
            "metric": f"queries/sec exact top-{SEARCH_K} @ {SEARCH_N}x{SEARCH_D} bf16", "Q": Q, "value": Q / (ms / 1e3),
            "unit": "queries/s", "ms_per_batch": ms, "local_ms_per_batch": local_ms,
            "path": ("cuda graph: " if graphed else "") + (f"local search -> [Q,k] records exchanged + merged by {sharded.exchange}" if sharded is not None else "local search"),
            "roofline": {"bound": bound,
                         "achieved": (shard_bytes / ms / 1e6) if bound == "hbm" else (flops / ms / 1e9),
                         "peak": pk["hbm_gbs"] if bound == "hbm" else pk["bf16_tflops_sustained"],
                         "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
                         "frac": max(t_hbm, t_mma) / (ms / 1e3), "frac_local": max(t_hbm, t_mma) / (local_ms / 1e3),
                         "algorithmic_bytes": shard_bytes,
                         # ncu capture is of one launch over the full 5M-row corpus on one GPU
                         "traffic": ncu_traffic("search_topk_kernel", 1, skip=0 if Q > 256 else 1) if world == 1 else None},
            "top1_score_mean": float(fs[:, 0].mean().item()), "clocks": sclk.summary(),
        }
    if sharded is not None:
        sharded.close()
    del index, sharded, corpus
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, n, threads, _, _ = cpu_encode_sample(12.0)
        cpu = {"value": rate, "unit": "chunks/s", "cores": threads, "kind": "port",
               "sample": f"{n} synthetic {SEQ}-token chunks, restated reference loop (oracle/refpath.py) over "
                         f"transformers.MPNetModel fp32, {threads} torch threads of {os.cpu_count()} host cores"}

    result = None
    if rank == 0:
        result = ({
            "metric": "chunks/sec encoded (MPNet)", "value": value, "unit": "chunks/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "all-mpnet-base-v2 encode, 1024 synthetic 384-token chunks per GPU per step (BASELINE configs[1])",
                       "seq_len": SEQ, "batch_per_gpu": BATCH, "global_batch": BATCH * world, "weights": "seeded synthetic (no checkpoint offline)",
                       "parallelism": f"dp{world} (chunk batches sharded, no collective)",
                       "l2": "inputs rotate over 4 batches; per-step activations 6.6 GB >> 126 MB L2"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "search": search, "unit_norm_check": checksum, "gflop_per_chunk": GFLOP_PER_CHUNK,
        })
    if world > 1:
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
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "b200":
        args.warmup = max(args.warmup, 3)  # timing rule: at least 3 untimed warm-up steps
    with _StdoutGuard():
        result = run_reference(args) if args.impl == "reference" else run_b200(args)
    if result is not None:  # rank 0 only: the one JSON line, alone on stdout
        print(json.dumps(result), flush=True)


if __name__ == "__main__":
    main()
