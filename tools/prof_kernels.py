"""Launch each hot kernel at its bench shape: once to warm up, once to be profiled.

    ncu --set full --clock-control none --import-source on \
        -k regex:'gemm16|search_topk|attention_tc' -s 8 -c 8 -o gpurun_out/prof python tools/prof_kernels.py

Order of the matching launches (x2): gemm QKV, gemm O+LN(residual), gemm FFN-up+GELU, gemm FFN-down+LN(residual)
(fp16 operands, the shipped default), attention, search Q=4096 (bf16, tensor-bound), search Q=64 (bf16,
HBM-bound), search 10k x 1M fp32 (kind::tf32)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arxiv_rag_b200 import _lib  # noqa: E402

dev = "cuda:0"
lib = _lib.lib()
st = lambda: torch.cuda.current_stream().cuda_stream
B, S, H = 1024, 384, 768
M = B * S
A768 = torch.randn(M, 768, device=dev).to(torch.float16)
A3072 = torch.randn(M, 3072, device=dev).to(torch.float16)
C = torch.empty(M, 3072, device=dev, dtype=torch.float16)
R = torch.randn(M, 768, device=dev).to(torch.float16)
Ws = {(N, K): (torch.randn(N, K, device=dev) * 0.04).to(torch.float16) for (N, K) in [(2304, 768), (768, 768), (3072, 768), (768, 3072)]}
bias = torch.randn(3072, device=dev)
colsum = torch.randn(3072, device=dev)
ln_g, ln_b = torch.ones(768, device=dev), torch.zeros(768, device=dev)
stats_in = torch.rand(6, M, 2, device=dev) * 50 + 100
stats_out = torch.empty(6, M, 2, device=dev)
qkv = torch.randn(M, 3 * H, device=dev).to(torch.float16)
relb = torch.randn(12, 1023, device=dev)
mask = torch.ones(B, S, device=dev, dtype=torch.int32)
ctx = torch.empty(M, H, device=dev, dtype=torch.float16)
N_C = int(os.environ.get("PROF_CORPUS", 2_000_000))
corpus = torch.nn.functional.normalize(torch.randn(N_C, 768, device=dev), dim=1).to(torch.bfloat16)
queries = {Q: torch.nn.functional.normalize(torch.randn(Q, 768, device=dev), dim=1).to(torch.bfloat16) for Q in (4096, 64)}
c32 = torch.nn.functional.normalize(torch.randn(1_000_000, 768, device=dev), dim=1)
q32 = torch.nn.functional.normalize(torch.randn(10_000, 768, device=dev), dim=1)
ws32 = torch.empty(lib.arb_topk_search_f32_workspace_bytes(10_000, 1_000_000, 768, 10, 0), dtype=torch.uint8, device=dev)
os32 = torch.empty(10_000, 10, device=dev)
oi32 = torch.empty(10_000, 10, device=dev, dtype=torch.int64)
ws = torch.empty(max(lib.arb_topk_search_workspace_bytes(1, 4096, N_C, 768, 10), lib.arb_topk_search_workspace_bytes(1, 64, N_C, 768, 10)),
                 dtype=torch.uint8, device=dev)
os_ = torch.empty(4096, 10, device=dev)
oi = torch.empty(4096, 10, device=dev, dtype=torch.int64)

for rep in range(2):
    for (N, K, epi) in [(2304, 768, 3), (768, 768, 5), (3072, 768, 4), (768, 3072, 5)]:  # the encode path's folded-LN epilogues
        A = A768 if K == 768 else A3072
        _lib.check(lib.arb_gemm16_lnfold(A.data_ptr(), K, Ws[(N, K)].data_ptr(), K, C.data_ptr(), N, bias.data_ptr(),
                                         R.data_ptr() if epi == 5 else 0, 768, colsum.data_ptr(), ln_g.data_ptr(), ln_b.data_ptr(),
                                         stats_in.data_ptr(), 6, 768, stats_out.data_ptr() if epi == 5 else 0, 1e-5, M, N, K, epi,
                                         _lib.ARB_DTYPE_F16, st()))
    _lib.check(lib.arb_attention16(qkv.data_ptr(), relb.data_ptr(), 512, mask.data_ptr(), ctx.data_ptr(), B, S, 12, 64,
                                   _lib.ARB_DTYPE_F16, 0, st()))
    for Q in (4096, 64):
        _lib.check(lib.arb_topk_search(queries[Q].data_ptr(), corpus.data_ptr(), 1, Q, N_C, 768, 10, os_.data_ptr(), oi.data_ptr(), 0,
                                       ws.data_ptr(), ws.numel(), st()))
    _lib.check(lib.arb_topk_search_f32(q32.data_ptr(), c32.data_ptr(), 10_000, 1_000_000, 768, 10, 1.001, os32.data_ptr(), oi32.data_ptr(), 0,
                                       0, 0, ws32.data_ptr(), ws32.numel(), st()))
    torch.cuda.synchronize()
print("prof_kernels done")
