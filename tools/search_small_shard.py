"""Local search time for the shard sizes an 8/4/2/1-GPU split of the 5M-row corpus leaves per rank
(Q <= 64: the HBM-bound regime), eager and through a CUDA graph, against the HBM floor."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arxiv_rag_b200.search import CorpusIndex

dev = torch.device("cuda:0")
PEAK = float(os.environ.get("HBM_GBS", 6552.6))


def timeit(fn, reps=300):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


g = torch.Generator(device=dev).manual_seed(0)
for rows in (625_000, 1_250_000, 2_500_000, 5_000_000):
    corpus = torch.nn.functional.normalize(torch.randn(rows, 768, device=dev, generator=g), dim=1).to(torch.bfloat16)
    index = CorpusIndex(corpus)
    floor = rows * 768 * 2 / (PEAK * 1e9) * 1e6
    for Q in (1, 64):
        q = torch.nn.functional.normalize(torch.randn(Q, 768, device=dev, generator=g), dim=1).to(torch.bfloat16)
        t = timeit(lambda: index.search(q, 10))
        gr = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            index.search(q, 10)
            ws = index.new_workspace(Q, 10) if hasattr(index, "new_workspace") else None
            with torch.cuda.graph(gr, stream=s):
                index.search(q, 10, workspace=ws) if ws is not None else index.search(q, 10)
        tg = timeit(gr.replay)
        print(f"rows {rows:8d} Q={Q:3d}: eager {t:7.1f} us  graph {tg:7.1f} us  floor {floor:7.1f} us  frac {floor/tg:.3f}")
    del index, corpus
    torch.cuda.empty_cache()
