"""Time the four per-layer GEMMs of the encode step (folded-LayerNorm epilogues) at the bench M, one by one
(0.5 s idle, 3 warm-ups, mean of 10 launches), like bench.py's roofline section."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arxiv_rag_b200 import _lib
lib = _lib.lib()
dev = "cuda"
M = int(os.environ.get("GEMM_M", 1024 * 384))
A768 = torch.randn(M, 768, device=dev).to(torch.float16)
A3072 = torch.randn(M, 3072, device=dev).to(torch.float16)
C = torch.empty(M, 3072, device=dev, dtype=torch.float16)
R = torch.randn(M, 768, device=dev).to(torch.float16)
parts = 6
st_in = torch.rand(parts, M, 2, device=dev) * 50 + 100
st_out = torch.empty(parts, M, 2, device=dev)
g, b = torch.ones(768, device=dev), torch.zeros(768, device=dev)
for (N, K, epi) in [(2304, 768, 3), (768, 768, 5), (3072, 768, 4), (768, 3072, 5)]:
    A = A768 if K == 768 else A3072
    W = (torch.randn(N, K, device=dev) * 0.04).to(torch.float16)
    bias, colsum = torch.randn(N, device=dev), torch.randn(N, device=dev)
    call = lambda: _lib.check(lib.arb_gemm16_lnfold(A.data_ptr(), K, W.data_ptr(), K, C.data_ptr(), N, bias.data_ptr(),
                                                    R.data_ptr() if epi == 5 else 0, 768, colsum.data_ptr(), g.data_ptr(), b.data_ptr(),
                                                    st_in.data_ptr(), parts, 768, st_out.data_ptr() if epi == 5 else 0, 1e-5, M, N, K, epi,
                                                    _lib.ARB_DTYPE_F16, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize(); time.sleep(0.5)
    for _ in range(3): call()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): call()
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 10
    print(f"N={N} K={K} epi={epi}: {ms:.3f} ms  {2.0*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
