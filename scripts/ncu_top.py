"""Top stalled SASS lines + headline metrics of an ncu report: python scripts/ncu_top.py gpurun_out/prof.ncu-rep [N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for r in rows[2:]:
    for w in want:
        for i, h in enumerate(hdr):
            if h == w:
                print(f"  {h} [{units[i]}] = {r[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
out = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            s = int(d["# Samples"])
        except Exception:
            continue
        stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
        out.append((s, d["Source"], stalls))
# the source page lists each SASS line twice (all samples / not-issued views share the row) -> dedupe by position
tot = sum(o[0] for o in out)
print("total samples", tot)
seen = set()
for s, srcline, stalls in sorted(out, key=lambda x: -x[0]):
    key = (s, srcline)
    if key in seen:
        continue
    seen.add(key)
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
    print(f"{s:8d} {100.0 * s / max(tot, 1):5.1f}%  {srcline[:70]:70s} {top}")
    if len(seen) >= topn:
        break
