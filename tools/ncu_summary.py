"""Summarise an `ncu --set full` report for profiles/: one CSV row per captured launch, plus a
JSON of per-kernel DRAM traffic that bench.py reads for its `roofline.traffic` fields.

    python tools/ncu_summary.py gpurun_out/prof_full.ncu-rep profiles/r1_ncu_full_v6_summary.csv profiles/ncu_traffic.json
"""
import csv
import io
import json
import subprocess
import sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def main():
    rep, out_csv, out_json = sys.argv[1:4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in COLS if c in idx]
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["ID", "Kernel Name"] + cols)
        w.writerow(["", ""] + [units[idx[c]] for c in cols])
        for r in data:
            w.writerow([r[idx["ID"]], r[idx["Kernel Name"]][:110]] + [r[idx[c]] for c in cols])
    # DRAM traffic per launch, in bytes, in capture order
    def to_bytes(v, unit):
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
        return float(v) * scale

    traffic = []
    for r in data:
        rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        traffic.append({"kernel": r[idx["Kernel Name"]][:80], "dram_bytes": rd + wr,
                        "ms": float(r[idx["gpu__time_duration.sum"]]) * {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6}[units[idx["gpu__time_duration.sum"]]]})
    json.dump({"source": rep.split("/")[-1], "launches": traffic}, open(out_json, "w"), indent=1)
    for t in traffic:
        print(f"{t['kernel'][:60]:60s} {t['ms']:8.3f} ms  dram {t['dram_bytes'] / 1e9:7.3f} GB")


if __name__ == "__main__":
    main()
