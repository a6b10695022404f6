"""One large-Q search in both tile schedules, for an ncu DRAM-traffic comparison:

    ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:search_topk \
        python tools/prof_search_modes.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arxiv_rag_b200 import _lib  # noqa: E402

Q, N, k = int(os.environ.get("PROF_Q", 4096)), int(os.environ.get("PROF_N", 5_000_000)), int(os.environ.get("PROF_K", 10))
dev = "cuda:0"
lib = _lib.lib()
c = torch.empty((N, 768), device=dev, dtype=torch.bfloat16)
for s in range(0, N, 500_000):
    c[s:s + 500_000] = torch.nn.functional.normalize(torch.randn(min(500_000, N - s), 768, device=dev), dim=1).to(torch.bfloat16)
q = torch.nn.functional.normalize(torch.randn(Q, 768, device=dev), dim=1).to(torch.bfloat16)
os_ = torch.zeros(Q, k, device=dev)
oi = torch.zeros(Q, k, device=dev, dtype=torch.int64)
for mode in (1, 2):
    _lib.check(lib.arb_set_search_mode(mode))
    ws = torch.empty(max(lib.arb_topk_search_workspace_bytes(_lib.ARB_DTYPE_BF16, Q, N, 768, k), 256), dtype=torch.uint8, device=dev)
    for _ in range(2):
        _lib.check(lib.arb_topk_search(q.data_ptr(), c.data_ptr(), _lib.ARB_DTYPE_BF16, Q, N, 768, k, os_.data_ptr(), oi.data_ptr(), 0,
                                       ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
print("done")
