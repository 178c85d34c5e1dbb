#!/usr/bin/env python
"""Per-SASS-instruction view of an .ncu-rep source page: prints the instructions sorted by
address with executed counts, avg active threads, stall samples; marks the hottest ones."""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_samp = sum(int(r[ix["# Samples"]]) for r in data)
print(f"total warp instructions {tot_inst}, samples {tot_samp}")
recs = []
for n, r in enumerate(data):
    recs.append((n, r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]]), float(r[ix["Avg. Threads Executed"]] or 0),
                 int(r[ix["# Samples"]]), int(r[ix["L1 Wavefronts Shared"]] or 0),
                 {k: int(r[ix[k]]) for k in ("stall_wait", "stall_short_sb", "stall_barrier", "stall_long_sb", "stall_math", "stall_branch_resolving", "stall_not_selected", "stall_selected", "stall_mio", "stall_no_inst") if int(r[ix[k]] or 0) > 0}))
if top:
    sel = sorted(recs, key=lambda x: -x[4])[:top]
    sel = sorted(sel)
else:
    sel = recs
for n, src, ie, thr, samp, wf, st in sel:
    print(f"{n:5d} {100*ie/tot_inst:5.2f}%i {100*samp/max(tot_samp,1):5.2f}%s thr{thr:5.1f} wf{wf:10d}  {src[:90]:90s} {st}")
