"""Launch the fused residual+LN cluster GEMM at the bench shape (warm-up + 1 profiled launch)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arxiv_rag_b200 import _lib
lib = _lib.lib()
M, N, K = 1024 * 384, 768, int(os.environ.get("PROF_K", 768))
A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
B = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
R = torch.randn(M, N, device="cuda").to(torch.bfloat16)
bias = torch.randn(N, device="cuda"); g = torch.ones(N, device="cuda")
for _ in range(2):
    _lib.check(lib.arb_gemm16_residual_ln(A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), N, bias.data_ptr(), R.data_ptr(), N,
                                          g.data_ptr(), bias.data_ptr(), 1e-5, M, N, K, _lib.ARB_DTYPE_BF16,
                                          torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("ok")
