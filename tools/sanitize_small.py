"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
Encode (query-time path: programmatic launches, narrow tiles; and a 4 k-token batch on CTA pairs),
bf16 search in both schedules (incl. the paced pair schedule) and the fp32 (tf32 + re-score) search."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arxiv_rag_b200.encoder import B200SentenceEncoder
from arxiv_rag_b200.search import CorpusIndex
from arxiv_rag_b200.weights import MPNetArch

arch = MPNetArch(num_layers=2)
enc = B200SentenceEncoder(None, arch=arch, max_batch=64, max_seq=128, dtype="fp16", seed=0)
rs = np.random.RandomState(0)
for B, S in ((1, 16), (3, 64), (40, 96), (64, 128)):
    ids = rs.randint(4, 30000, (B, S)).astype(np.int32)
    mask = (np.arange(S)[None, :] < rs.randint(1, S + 1, (B, 1))).astype(np.int32)
    out = enc.encode((ids, mask), batch_size=64)
    assert np.isfinite(out).all()
print("encode ok")
g = torch.Generator(device="cuda").manual_seed(0)
c = torch.nn.functional.normalize(torch.randn(40_000, 768, device="cuda", generator=g), dim=1)
q = torch.nn.functional.normalize(torch.randn(700, 768, device="cuda", generator=g), dim=1)
ib = CorpusIndex(c.to(torch.bfloat16))
for Q, k in ((1, 10), (64, 10), (300, 10), (700, 100)):
    s, i = ib.search(q[:Q].to(torch.bfloat16), k)
    assert torch.isfinite(s).all() and int(i.min()) >= 0
print("bf16 search ok")
i32 = CorpusIndex(c)
s, i = i32.search(q[:300], 10)
assert torch.isfinite(s).all()
torch.cuda.synchronize()
print("fp32 search ok")
