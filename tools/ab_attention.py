"""A/B helper: time the tcgen05 attention kernel (B=1024, S from argv or 384) with the library named by ARB_LIB_PATH."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arxiv_rag_b200 import _lib
lib = _lib.lib()
B, S, H = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 384, 768
qkv = torch.randn(B * S, 3 * H, device="cuda").to(torch.bfloat16)
relb = torch.randn(12, 1023, device="cuda")
mask = torch.ones(B, S, device="cuda", dtype=torch.int32)
ctx = torch.empty(B * S, H, device="cuda", dtype=torch.bfloat16)
call = lambda: _lib.check(lib.arb_attention16(qkv.data_ptr(), relb.data_ptr(), 512, mask.data_ptr(), ctx.data_ptr(), B, S, 12, 64,
                                               _lib.ARB_DTYPE_BF16, 2, torch.cuda.current_stream().cuda_stream))
for _ in range(5): call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): call()
e1.record(); torch.cuda.synchronize()
print(os.environ.get("ARB_LIB_PATH", "current"), f"S={S} {e0.elapsed_time(e1)/20:.3f} ms")
