"""A/B helper: time the 1024x384 encode step with the library named by ARB_LIB_PATH (default: the in-tree build)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arxiv_rag_b200.encoder import B200SentenceEncoder
enc = B200SentenceEncoder(None, max_batch=1024, max_seq=384)
ids = torch.randint(4, 30000, (1024, 384), device="cuda", dtype=torch.int32)
m = torch.ones(1024, 384, device="cuda", dtype=torch.int32)
for _ in range(3): enc.encode_tokens(ids, m)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(8): enc.encode_tokens(ids, m)
e1.record(); torch.cuda.synchronize()
print(os.environ.get("ARB_LIB_PATH", "current"), f"{e0.elapsed_time(e1)/8:.2f} ms")
