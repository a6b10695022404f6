"""A/B helper: one configuration of the encode step, timed and checked.

    [ARB_LIB_PATH=...] [ARB_ATTN_DEFER=0|1] python tools/ab_step.py [--dtype fp16] [--batch 1024] [--seq 384] [--kernels]

Prints the step time (CUDA events, 8 steps after 3 warm-ups), optionally the per-kernel averages of
one step (CUPTI activity records through torch.profiler), and the cosine of a small ragged batch
against the fp32 oracle — so a variant that is fast but wrong shows up in the same line.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arxiv_rag_b200.encoder import B200SentenceEncoder  # noqa: E402
from arxiv_rag_b200.weights import ALL_MPNET_BASE_V2, synthetic_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="fp16")
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--seq", type=int, default=384)
ap.add_argument("--kernels", action="store_true")
ap.add_argument("--no-check", action="store_true")
ap.add_argument("--graph", action="store_true", help="also time the step replayed from one CUDA graph")
args = ap.parse_args()

tag = f"lib={os.path.basename(os.environ.get('ARB_LIB_PATH', 'default'))} dtype={args.dtype} B{args.batch} S{args.seq}"
sd = synthetic_state_dict(ALL_MPNET_BASE_V2, 0)
enc = B200SentenceEncoder(sd, max_batch=args.batch, max_seq=args.seq, dtype=args.dtype)
cos_txt = ""
if not args.no_check:
    from oracle import encode_oracle as eo

    ids, mask = eo.synthetic_tokens(8, min(96, args.seq), seed=1)
    ref = eo.oracle_encode(eo.reference_model(ALL_MPNET_BASE_V2, sd), ids, mask)
    got = enc.encode((ids, mask), batch_size=8)
    cos = (got * ref).sum(1)
    cos_txt = f" | cos vs fp32 oracle min {cos.min():.6f} mean {cos.mean():.6f}"
ids = torch.randint(4, 30000, (args.batch, args.seq), device="cuda", dtype=torch.int32)
m = torch.ones(args.batch, args.seq, device="cuda", dtype=torch.int32)
for _ in range(3):
    enc.encode_tokens(ids, m)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(8):
    enc.encode_tokens(ids, m)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 8
graph_txt = ""
if args.graph:
    enc.encode_tokens_graphed(ids, m)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(8):
        enc.encode_tokens_graphed(ids, m)
    e1.record()
    torch.cuda.synchronize()
    graph_txt = f" | one CUDA graph: {e0.elapsed_time(e1) / 8:.2f} ms/step"
print(f"{tag}: {ms:.2f} ms/step {args.batch / ms * 1e3:.0f} chunks/s{graph_txt}{cos_txt}", flush=True)
if args.kernels:
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        e0.record()
        enc.encode_tokens(ids, m)
        e1.record()
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    ksum = sum(e.device_time_total for e in rows if "arb::" in e.key)
    print(f"    profiled step: {e0.elapsed_time(e1):.2f} ms between events, {ksum / 1e3:.2f} ms inside kernels", flush=True)
    for ev in rows[:10]:
        print(f"    {ev.key[:100]:100s} n={ev.count:3d} avg {ev.device_time_total / max(ev.count, 1):9.1f} us total {ev.device_time_total / 1e3:7.2f} ms", flush=True)
enc.close()
