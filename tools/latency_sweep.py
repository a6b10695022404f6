"""BASELINE configs[4]: end-to-end encode-then-search latency, query batch 1..4096, against a bf16
corpus row-sharded over the ranks (50M x 768 over 8 GPUs = 6.25M rows per GPU; on fewer GPUs the
per-GPU shard is kept at 6.25M rows so a single rank measures one shard's share of the work).

    python tools/latency_sweep.py [--rows-per-gpu 6250000] [--seq 64] [--k 10]
    torchrun --nproc-per-node 8 tools/latency_sweep.py

Per batch size, every step: query token ids copied from pinned host memory -> MPNet encode (one
CUDA-graph launch) -> ShardedCorpusIndex.search_graphed (fused score+top-k over the local shard,
ONE NCCL all-gather of the [Q,k] records, merge; one CUDA-graph launch) -> result ids/scores
copied back to pinned host memory. Timed with CUDA events around the whole step (median of 20
after 5 warm-ups), max over ranks. Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arxiv_rag_b200.encoder import B200SentenceEncoder  # noqa: E402
from arxiv_rag_b200.search import ShardedCorpusIndex  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-per-gpu", type=int, default=6_250_000)
    ap.add_argument("--seq", type=int, default=64)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--max-batch", type=int, default=4096)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29578", rank=0, world_size=1)
    enc = B200SentenceEncoder(None, max_batch=args.max_batch, max_seq=args.seq, dtype="bf16")
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    n = args.rows_per_gpu
    corpus = torch.empty((n, 768), device=dev, dtype=torch.bfloat16)
    for s in range(0, n, 500_000):
        e = min(s + 500_000, n)
        corpus[s:e] = torch.nn.functional.normalize(torch.randn(e - s, 768, device=dev, generator=g), dim=1).to(torch.bfloat16)
    index = ShardedCorpusIndex(corpus, n * world)
    rows = []
    Q = 1
    while Q <= args.max_batch:
        gq = torch.Generator().manual_seed(Q)
        h_ids = torch.randint(4, 30000, (Q, args.seq), dtype=torch.int32, generator=gq).pin_memory()
        h_mask = torch.ones((Q, args.seq), dtype=torch.int32).pin_memory()
        d_ids = torch.empty((Q, args.seq), device=dev, dtype=torch.int32)
        d_mask = torch.empty((Q, args.seq), device=dev, dtype=torch.int32)
        h_scores = torch.empty((Q, args.k), dtype=torch.float32).pin_memory()
        h_out = torch.empty((Q, args.k), dtype=torch.int64).pin_memory()
        q16 = torch.empty((Q, 768), device=dev, dtype=torch.bfloat16)

        def step():
            d_ids.copy_(h_ids, non_blocking=True)
            d_mask.copy_(h_mask, non_blocking=True)
            q16.copy_(enc.encode_tokens_graphed(d_ids, d_mask))
            s, i = index.search_graphed(q16, args.k)
            h_scores.copy_(s, non_blocking=True)
            h_out.copy_(i, non_blocking=True)

        for _ in range(5):
            step()
        torch.cuda.synchronize()
        times = []
        for _ in range(20):
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        med = sorted(times)[len(times) // 2]
        if world > 1:
            t = torch.tensor([med], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            med = float(t.item())
        hbm_ms = n * 768 * 2 / 6552.6e9 * 1e3
        rows.append({"Q": Q, "ms": med, "queries_per_s": Q / med * 1e3, "shard_hbm_floor_ms": hbm_ms})
        Q *= 2
    if rank == 0:
        print(json.dumps({"metric": "encode+search latency (pinned host token ids in, pinned host top-k out)", "n_gpus": world,
                          "corpus_rows": n * world, "rows_per_gpu": n, "seq": args.seq, "k": args.k, "dtype": "bf16",
                          "sweep": rows}), flush=True)
    index.close()
    enc.close()
    del index, enc, corpus
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sys.stdout.flush()
    os._exit(0)  # measurement tool: skip communicator teardown


if __name__ == "__main__":
    main()
