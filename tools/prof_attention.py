"""Launch the tcgen05 attention kernel at the bench shape (warm-up + 1 profiled launch)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arxiv_rag_b200 import _lib
lib = _lib.lib()
B, S, H = int(os.environ.get("PROF_B", 1024)), int(os.environ.get("PROF_S", 384)), 768
qkv = torch.randn(B * S, 3 * H, device="cuda").to(torch.bfloat16)
bk = torch.tensor([lib.arb_mpnet_relative_bucket(int(r), 32, 128) for r in range(-511, 512)], device="cuda")
relb = (torch.randn(32, 12, device="cuda") * 0.7)[bk].t().contiguous()  # the table as MPNet builds it (32 buckets)
mask = torch.ones(B, S, device="cuda", dtype=torch.int32)
ctx = torch.empty(B * S, H, device="cuda", dtype=torch.bfloat16)
for impl in [int(x) for x in os.environ.get("PROF_IMPLS", "2,3").split(",")]:
    for _ in range(2):
        _lib.check(lib.arb_attention16(qkv.data_ptr(), relb.data_ptr(), 512, mask.data_ptr(), ctx.data_ptr(), B, S, 12, 64,
                                       _lib.ARB_DTYPE_BF16, impl, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("ok")
