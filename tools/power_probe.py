"""Sustained small-Q search: time per pass, SM clock and power for several Q (is the HBM-bound pass power-capped,
and do zero query rows of the 128-row MMA tile cost power?).   python tools/power_probe.py"""
import os, subprocess, sys, threading
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arxiv_rag_b200.search import CorpusIndex

N = int(os.environ.get("PROBE_ROWS", 5_000_000))
c = torch.empty((N, 768), device="cuda", dtype=torch.bfloat16)
for s in range(0, N, 500_000):
    c[s:s + 500_000] = torch.nn.functional.normalize(torch.randn(min(500_000, N - s), 768, device="cuda"), dim=1).to(torch.bfloat16)
index = CorpusIndex(c)
for Q in [int(x) for x in os.environ.get("PROBE_Q", "1,16,64,128").split(",")]:
    q = torch.nn.functional.normalize(torch.randn(Q, 768, device="cuda"), dim=1).to(torch.bfloat16)
    if os.environ.get("PROBE_PAD"):  # zero rows up to a full 128-row tile: no out-of-bounds TMA boxes
        qp = torch.zeros((128, 768), device="cuda", dtype=torch.bfloat16)
        qp[:Q] = q
        q, Q = qp, 128
    os_ = torch.empty((Q, 10), device="cuda"); oi = torch.empty((Q, 10), device="cuda", dtype=torch.int64)
    for _ in range(200): index.search(q, 10, out_scores=os_, out_ids=oi)
    torch.cuda.synchronize()
    rows, stop = [], threading.Event()
    def poll():
        while not stop.is_set():
            out = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True).stdout.strip()
            if out: rows.append([float(x) for x in out.split(",")])
            stop.wait(0.1)
    t = threading.Thread(target=poll); t.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(600): index.search(q, 10, out_scores=os_, out_ids=oi)
    e1.record(); torch.cuda.synchronize(); stop.set(); t.join()
    ms = e0.elapsed_time(e1) / 600
    clk = sorted(r[0] for r in rows)[len(rows) // 2] if rows else 0
    pw = max(r[1] for r in rows) if rows else 0
    print(f"Q={Q:4d}: {ms:.3f} ms/pass  {N * 768 * 2 / ms / 1e6:.0f} GB/s  SM {clk:.0f} MHz  power max {pw:.0f} W", flush=True)
