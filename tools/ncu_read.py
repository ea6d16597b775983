"""Summarise an .ncu-rep: per-kernel headline metrics + stall samples by CUDA source line / SASS.
    python tools/ncu_read.py gpurun_out/x.ncu-rep [launch_index] [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 22

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("---", d.get("ID"), d.get("Kernel Name", "")[:60])
    for k in KEYS[1:]:
        if k in d:
            print("   %-90s %s %s" % (k, d[k], rows[1][hdr.index(k)]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--launch-skip", str(idx),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if len(r) > 5][0]
hdr = rows[hi]
S = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
cuda_rows, sass_rows = [], []
for r in rows[hi + 1:]:
    if len(r) <= S:
        continue
    try:
        n = int(r[S])
    except ValueError:
        continue
    st = sorted(((int(r[c]) if r[c].isdigit() else 0, hdr[c][6:]) for c in stall_cols), reverse=True)[:3]
    if r[0]:
        cuda_rows.append((n, r[0], r[1].strip()[:100], st))
    else:
        sass_rows.append((n, r[2][-5:], r[3][:90], st))
tot = sum(n for n, *_ in sass_rows) or 1
print("total samples", tot)
print("== by CUDA line")
for n, ln, text, st in sorted(cuda_rows, reverse=True)[:topn]:
    print("%6d %5.1f%% L%-5s %s  %s" % (n, 100.0 * n / tot, ln, text, st))
print("== by SASS")
for n, addr, text, st in sorted(sass_rows, reverse=True)[:topn]:
    print("%6d %5.1f%% %s %s  %s" % (n, 100.0 * n / tot, addr, text, st))
