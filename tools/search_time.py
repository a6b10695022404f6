"""Time one search shape (CUDA events, mean of REPS after warm-up): PROF_Q / PROF_N / PROF_K / PROF_REPS."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arxiv_rag_b200.search import CorpusIndex
Q, N, k = int(os.environ.get("PROF_Q", 4096)), int(os.environ.get("PROF_N", 5_000_000)), int(os.environ.get("PROF_K", 10))
reps = int(os.environ.get("PROF_REPS", 20))
dev = "cuda:0"
c = torch.empty((N, 768), device=dev, dtype=torch.bfloat16)
g = torch.Generator(device=dev).manual_seed(0)
for s in range(0, N, 500_000):
    c[s:s + 500_000] = torch.nn.functional.normalize(torch.randn(min(500_000, N - s), 768, device=dev, generator=g), dim=1).to(torch.bfloat16)
q = torch.nn.functional.normalize(torch.randn(Q, 768, device=dev, generator=g), dim=1).to(torch.bfloat16)
index = CorpusIndex(c)
for _ in range(5):
    s_, i_ = index.search(q, k)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    s_, i_ = index.search(q, k)
b.record()
torch.cuda.synchronize()
print(f"Q={Q} N={N} k={k} MAX_CPS={os.environ.get('ARB_SEARCH_MAX_CPS', 'default')}: {a.elapsed_time(b)/reps:.3f} ms  checksum {float(s_.sum()):.4f} {int(i_.sum())}")
