"""Summarise an .ncu-rep: headline metrics (raw page) + hottest SASS lines by stall samples (source page)."""
import csv, io, subprocess, sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "smsp__cycles_active.avg", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("== kernel:", d.get("Kernel Name", "?")[:80])
    for w in WANT:
        if w in d:
            print(f"  {w:75s} {d[w]:>16s} {units[hdr.index(w)]}")
    stalls = {h: float(d[h]) for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and d[h] not in ("", "n/a")}
    for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]:
        print(f"  stall {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:30s} {v:8.3f}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next((i for i, r in enumerate(rows) if r and r[0] == "Address"), None)
if hi is not None:
    hdr = rows[hi]
    cs_name = "# Samples" if "# Samples" in hdr else next((h for h in hdr if "Samples" in h), None)
    if cs_name is None:
        raise SystemExit(0)
    ci, cs, cx = hdr.index("Source"), hdr.index(cs_name), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[hi + 1:]:
        try:
            st = sorted(((float(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
            data.append((float(r[cs] or 0), r[ci].strip(), r[cx], st))
        except (ValueError, IndexError):
            pass
    tot = sum(d[0] for d in data) or 1
    print(f"total samples {tot:.0f}")
    for s_, ins, ex, st in sorted(data, key=lambda t: -t[0])[:top_n]:
        why = ", ".join(f"{n}:{v:.0f}" for v, n in st if v > 0)
        print(f"  {100*s_/tot:5.1f}%  {ins[:90]:90s} exec={ex:>9s}  {why}")
