"""Attention kernels alone at the bench shape: impl 1 (mma.sync), 2 (tcgen05, round 1), 3 (tcgen05, sub-block pipelined)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arxiv_rag_b200 import _lib
lib = _lib.lib()
shapes = [tuple(int(x) for x in t.split("x")) for t in os.environ["ATTN_SHAPES"].split(",")] if os.environ.get("ATTN_SHAPES") \
    else [(1024, 384), (1024, 256), (1024, 128), (4096, 64)]
for (B, S) in shapes:
    H = 768
    for dt, code in ((torch.float16, _lib.ARB_DTYPE_F16), (torch.bfloat16, _lib.ARB_DTYPE_BF16)):
        qkv = torch.randn(B * S, 3 * H, device="cuda").to(dt)
        if os.environ.get("ATTN_RANDOM_BIAS") == "1":
            relb = torch.randn(12, 1023, device="cuda")
        else:  # the table as MPNet builds it: 32 buckets, constant beyond |j - i| = 91
            bk = torch.tensor([lib.arb_mpnet_relative_bucket(int(r), 32, 128) for r in range(-511, 512)], device="cuda")
            relb = (torch.randn(32, 12, device="cuda") * 0.7)[bk].t().contiguous()
        mask = torch.ones(B, S, device="cuda", dtype=torch.int32)
        if os.environ.get("ATTN_RAGGED"):  # lengths S - (b mod R): what a length-sorted batch of real chunks looks like
            lens = S - (torch.arange(B, device="cuda") % int(os.environ["ATTN_RAGGED"]))
            mask = (torch.arange(S, device="cuda")[None, :] < lens[:, None]).int().contiguous()
        ctx = torch.empty(B * S, H, device="cuda", dtype=dt)
        fl = 4.0 * B * 12 * S * S * 64
        line = f"attention B{B} S{S} {str(dt)[6:]}:"
        for impl in [int(x) for x in os.environ.get("ATTN_IMPLS", "2").split(",")]:
            call = lambda: _lib.check(lib.arb_attention16(qkv.data_ptr(), relb.data_ptr(), 512, mask.data_ptr(), ctx.data_ptr(), B, S, 12, 64, code, impl, torch.cuda.current_stream().cuda_stream))
            for _ in range(3): call()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): call()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            line += f"  impl{impl} {ms:.3f} ms ({fl / ms / 1e9:.0f} TFLOP/s)"
        print(line, flush=True)
        del qkv, ctx
