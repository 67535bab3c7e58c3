#!/usr/bin/env python
"""Read an .ncu-rep here (no GPU): headline metrics of the kernel and its warp-state samples grouped into SASS
regions (runs of instructions with the same execution count).  usage: tools/ncu_regions.py file.ncu-rep [show a b]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__warps_eligible.avg.per_cycle_active"]
for h, u, v in zip(hdr, rows[1], vals):
    if h in want:
        print(f"{h} = {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
S = [int(r[ix["# Samples"]] or 0) for r in data]
E = [int(r[ix["Instructions Executed"]] or 0) for r in data]
tot = sum(S)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
if len(sys.argv) > 2 and sys.argv[2] == "show":
    for k in range(int(sys.argv[3]), int(sys.argv[4])):
        r = data[k]
        top = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
        print(k, f"{100 * S[k] / tot:5.2f}%", E[k], r[ix["Source"]][:84], [t for t in top if t[0]])
    sys.exit(0)
b = [0]
for k in range(1, len(E)):
    if E[k] != E[k - 1] and (E[k - 1] == 0 or abs(E[k] - E[k - 1]) / max(E[k], E[k - 1]) > 0.05):
        b.append(k)
b.append(len(E))
print("total samples", tot)
for a, c in zip(b[:-1], b[1:]):
    s = sum(S[a:c])
    if s * 200 > tot:
        st = {h[6:]: sum(int(data[k][ix[h]] or 0) for k in range(a, c)) for h in stalls}
        top = sorted(st.items(), key=lambda x: -x[1])[:5]
        nf = sum("FFMA2" in data[k][ix["Source"]] for k in range(a, c))
        print(f"[{a},{c}) n={c - a} ffma2={nf} exec={E[a]} samples={100 * s / tot:.1f}%", [(k, f"{100 * v / tot:.1f}") for k, v in top])
