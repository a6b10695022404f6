"""One search shape, launched a few times, for an ncu capture of the search kernel:

    PROF_Q=1024 PROF_N=1000000 PROF_K=32 ncu --set full --clock-control none --import-source on \
        -k regex:search_topk -s 2 -c 1 -o gpurun_out/search_k32 python tools/prof_search.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arxiv_rag_b200 import _lib  # noqa: E402

Q, N, k = int(os.environ.get("PROF_Q", 1024)), int(os.environ.get("PROF_N", 1_000_000)), int(os.environ.get("PROF_K", 32))
dev = "cuda:0"
lib = _lib.lib()
c = torch.nn.functional.normalize(torch.randn(N, 768, device=dev), dim=1).to(torch.bfloat16)
q = torch.nn.functional.normalize(torch.randn(Q, 768, device=dev), dim=1).to(torch.bfloat16)
ws = torch.empty(max(lib.arb_topk_search_workspace_bytes(_lib.ARB_DTYPE_BF16, Q, N, 768, k), 256), dtype=torch.uint8, device=dev)
os_ = torch.zeros(Q, k, device=dev)
oi = torch.zeros(Q, k, device=dev, dtype=torch.int64)
for _ in range(4):
    _lib.check(lib.arb_topk_search(q.data_ptr(), c.data_ptr(), _lib.ARB_DTYPE_BF16, Q, N, 768, k, os_.data_ptr(), oi.data_ptr(), 0,
                                   ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("prof_search done", Q, N, k)
