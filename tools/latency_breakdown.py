"""Where the query-time latency (BASELINE configs[4], Q <= 64) goes: the encode graph alone, the
search graph alone, an empty-kernel graph of the same node count (launch floor), and the whole step.
    python tools/latency_breakdown.py [rows]      # rows per GPU, default 6.25M
LAT_NCU=1 runs one un-graphed forward per Q only (for an ncu launch list)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arxiv_rag_b200.encoder import B200SentenceEncoder
from arxiv_rag_b200.search import CorpusIndex

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 6_250_000
S = int(os.environ.get("LAT_S", 64))
dev = torch.device("cuda:0")
enc = B200SentenceEncoder(None, max_batch=int(os.environ.get("LAT_MAXB", 64)), max_seq=S, dtype=os.environ.get("LAT_DTYPE", "fp16"), seed=0)
ncu = os.environ.get("LAT_NCU") == "1"
if os.environ.get("LAT_GEMM_MODE"):
    from arxiv_rag_b200 import _lib
    _lib.check(_lib.lib().arb_set_gemm_mode(int(os.environ["LAT_GEMM_MODE"])))


def timeit(fn, reps=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


if ncu:
    for Q in (1, 64):
        ids = torch.randint(4, 30000, (Q, S), device=dev, dtype=torch.int32)
        mask = torch.ones((Q, S), device=dev, dtype=torch.int32)
        for _ in range(2):
            enc.encode_tokens(ids, mask)
    torch.cuda.synchronize()
    print("ok")
    sys.exit(0)

if os.environ.get("LAT_ENCODE_ONLY") == "1":
    for Q in [int(x) for x in os.environ.get("LAT_QS", "1,2,8,16,32,64").split(",")]:
        ids = torch.randint(4, 30000, (Q, S), device=dev, dtype=torch.int32)
        mask = torch.ones((Q, S), device=dev, dtype=torch.int32)
        print(f"Q={Q:3d} S={S}: encode plain {timeit(lambda: enc.encode_tokens(ids, mask)):7.1f} us  "
              f"graph {timeit(lambda: enc.encode_tokens_graphed(ids, mask)):7.1f} us")
    sys.exit(0)
g = torch.Generator(device=dev).manual_seed(0)
corpus = torch.empty((rows, 768), device=dev, dtype=torch.bfloat16)
for s in range(0, rows, 500_000):
    e = min(rows, s + 500_000)
    corpus[s:e] = torch.nn.functional.normalize(torch.randn(e - s, 768, device=dev, generator=g), dim=1).to(torch.bfloat16)
index = CorpusIndex(corpus)
print(f"launches per encode: {enc.launches_per_encode}; rows {rows}; HBM floor {rows*768*2/6.55e12*1e6:.0f} us")
for Q in (1, 8, 64):
    ids = torch.randint(4, 30000, (Q, S), device=dev, dtype=torch.int32)
    mask = torch.ones((Q, S), device=dev, dtype=torch.int32)
    q16 = torch.empty((Q, 768), device=dev, dtype=torch.bfloat16)
    t_plain = timeit(lambda: enc.encode_tokens(ids, mask))
    t_graph = timeit(lambda: enc.encode_tokens_graphed(ids, mask))
    q16.copy_(enc.encode_tokens_graphed(ids, mask))
    t_search = timeit(lambda: index.search(q16, 10), 100)
    t_search_g = timeit(lambda: index.search_graphed(q16, 10), 100) if hasattr(index, "search_graphed") else float("nan")

    def both():
        q16.copy_(enc.encode_tokens_graphed(ids, mask))
        return index.search(q16, 10)
    t_both = timeit(both, 100)
    print(f"Q={Q:3d} S={S}: encode plain {t_plain:7.1f} us  graph {t_graph:7.1f} us | search {t_search:7.1f} us  graphed {t_search_g:7.1f} us | encode+search {t_both:7.1f} us")
