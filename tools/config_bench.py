"""BASELINE configs[2] and configs[3], one JSON line each on rank 0.

    python tools/config_bench.py cfg2                  # 10k queries x 1M x 768 fp32, top-10, 1 GPU
    torchrun --nproc-per-node N tools/config_bench.py cfg3   # 100k queries x 5M x 768 bf16, top-100, N GPUs

cfg3 shards the corpus row-wise (ShardedCorpusIndex: local fused top-100 -> one all-gather of the
[Q,k] records -> exchange -> merge) and feeds the 100k queries in batches of --batch (default 32768). Inputs are
resident; timed with CUDA events, max over ranks. The roofline denominator is
max(bytes/HBM peak, flops/sustained bf16 peak) with MEASURED_PEAKS.json numbers.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arxiv_rag_b200.search import CorpusIndex, ShardedCorpusIndex, shard_bounds  # noqa: E402

HBM_GBS, BF16_TF = 6552.6, 1391.8


def unit_rows(n, dev, seed, dtype):
    g = torch.Generator(device=dev).manual_seed(seed)
    out = torch.empty((n, 768), device=dev, dtype=dtype)
    for s in range(0, n, 500_000):
        e = min(s + 500_000, n)
        out[s:e] = torch.nn.functional.normalize(torch.randn(e - s, 768, device=dev, generator=g), dim=1).to(dtype)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["cfg2", "cfg3"])
    ap.add_argument("--batch", type=int, default=32768)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def sync_max(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.config == "cfg2":
        Q, N, k, dt, esz = 10_000, 1_000_000, 10, torch.float32, 4
        index = CorpusIndex(unit_rows(N, dev, 100, dt))
        q = unit_rows(Q, dev, 7, dt)
        run = lambda: index.search(q, k)
        n_local = N
    else:
        Q, N, k, dt, esz = 100_000, 5_000_000, 100, torch.bfloat16, 2
        lo, hi = shard_bounds(N, world, rank)
        n_local = hi - lo
        corpus = unit_rows(n_local, dev, 100 + rank, dt)
        index = ShardedCorpusIndex(corpus, N) if world > 1 else CorpusIndex(corpus)
        q = unit_rows(Q, dev, 7, dt)
        outs = torch.empty((Q, k), device=dev, dtype=torch.float32)
        outi = torch.empty((Q, k), device=dev, dtype=torch.int64)

        def run():
            for b in range(0, Q, args.batch):
                s, i = index.search(q[b:b + args.batch], k)
                outs[b:b + args.batch].copy_(s)
                outi[b:b + args.batch].copy_(i)
            return outs, outi

    run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        s, i = run()
    e1.record()
    torch.cuda.synchronize()
    ms = sync_max(e0.elapsed_time(e1) / args.reps)
    flops = 2.0 * Q * n_local * 768 * (3 if args.config == "cfg2" else 1)  # fp32 path: 3 bf16 products per element
    algo_flops = 2.0 * Q * n_local * 768
    bytes_ = n_local * 768 * esz + Q * 768 * esz + Q * k * 12
    t_floor = max(bytes_ / (HBM_GBS * 1e9), algo_flops / (BF16_TF * 1e12))
    if rank == 0:
        print(json.dumps({
            "config": args.config, "Q": Q, "N": N, "k": k, "dtype": str(dt).split(".")[-1], "n_gpus": world, "ms": ms,
            "queries_per_s": Q / ms * 1e3, "algorithmic_tflops_per_gpu": algo_flops / ms / 1e9,
            "tensor_tflops_per_gpu": flops / ms / 1e9, "roofline_frac": t_floor / (ms / 1e3),
            "bound": "tensor", "batch": args.batch if args.config == "cfg3" else Q,
            "top1_mean": float(s[:, 0].mean().item())}), flush=True)
    if world > 1:
        if hasattr(index, "close"):
            index.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
