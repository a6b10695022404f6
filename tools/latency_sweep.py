"""BASELINE configs[4]: end-to-end encode-then-search latency, query batch 1..4096, against a bf16
corpus row-sharded over the ranks (50M x 768 over 8 GPUs = 6.25M rows per GPU; on fewer GPUs the
per-GPU shard is kept at 6.25M rows so a single rank measures one shard's share of the work).

    python tools/latency_sweep.py [--rows-per-gpu 6250000] [--seq 64] [--k 10]
    torchrun --nproc-per-node 8 tools/latency_sweep.py

Per batch size: query tokens resident on the device -> MPNet encode (one CUDA-graph launch) ->
fused score+top-k over the local shard -> NCCL all-gather of the [Q,k] lists + k-way merge.
Timed with CUDA events (median of 20 after 5 warm-ups), max over ranks. Prints one JSON line.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arxiv_rag_b200.encoder import B200SentenceEncoder  # noqa: E402
from arxiv_rag_b200.search import CorpusIndex, merge_topk  # noqa: E402


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
    enc = B200SentenceEncoder(None, max_batch=args.max_batch, max_seq=args.seq, dtype="bf16")
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    n = args.rows_per_gpu
    corpus = torch.empty((n, 768), device=dev, dtype=torch.bfloat16)
    for s in range(0, n, 500_000):
        e = min(s + 500_000, n)
        corpus[s:e] = torch.nn.functional.normalize(torch.randn(e - s, 768, device=dev, generator=g), dim=1).to(torch.bfloat16)
    index = CorpusIndex(corpus, id_offset=rank * n)
    rows = []
    Q = 1
    while Q <= args.max_batch:
        gq = torch.Generator(device=dev).manual_seed(Q)
        ids = torch.randint(4, 30000, (Q, args.seq), device=dev, dtype=torch.int32, generator=gq)
        mask = torch.ones((Q, args.seq), device=dev, dtype=torch.int32)

        def step():
            emb = enc.encode_tokens_graphed(ids, mask)
            ls, li = index.search(emb.to(torch.bfloat16), args.k)
            if world > 1:
                ga = torch.empty((world, Q, args.k), device=dev, dtype=torch.float32)
                gi = torch.empty((world, Q, args.k), device=dev, dtype=torch.int64)
                dist.all_gather_into_tensor(ga, ls)
                dist.all_gather_into_tensor(gi, li)
                return merge_topk(ga, gi)
            return ls, li

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
        rows.append({"Q": Q, "ms": med, "queries_per_s": Q / med * 1e3})
        Q *= 2
    if rank == 0:
        print(json.dumps({"metric": "encode+search latency", "n_gpus": world, "corpus_rows": n * world, "rows_per_gpu": n,
                          "seq": args.seq, "k": args.k, "dtype": "bf16", "sweep": rows}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
