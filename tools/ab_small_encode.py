"""A/B helper: CUDA-graph encode latency at query-time batch shapes with the library named by ARB_LIB_PATH."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arxiv_rag_b200.encoder import B200SentenceEncoder
enc = B200SentenceEncoder(None, max_batch=64, max_seq=128)
out = []
for (B, S) in [(1, 64), (3, 40), (8, 100), (16, 64)]:
    ids = torch.randint(4, 30000, (B, S), device="cuda", dtype=torch.int32)
    m = torch.ones(B, S, device="cuda", dtype=torch.int32)
    for _ in range(5): enc.encode_tokens_graphed(ids, m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): enc.encode_tokens_graphed(ids, m)
    e1.record(); torch.cuda.synchronize()
    out.append(f"{B}x{S}: {e0.elapsed_time(e1)/50*1e3:.0f} us")
print(os.environ.get("ARB_LIB_PATH", "current"), " | ".join(out))
