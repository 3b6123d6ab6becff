#!/usr/bin/env python
"""Top stall locations of a kernel from an .ncu-rep (SASS view of `ncu --page source --csv`): which instructions the
warp-stall samples land on, with the dominant stall reasons per instruction.
    python tools/ncu_hotspots.py gpurun_out/prof.ncu-rep [N]"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = None
data = []
for r in rows:
    if len(r) > 3 and r[0] == "Address":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        if len(r) >= 2 and r[0] == "Kernel Name":
            print("kernel:", r[1])
        continue
    data.append(r)
ci = {h: i for i, h in enumerate(hdr)}
s_all = ci["Warp Stall Sampling (All Samples)"]
stall_cols = [(h, i) for h, i in ci.items() if h.startswith("stall_") or h.lower().startswith("stall")]
tot = sum(float(r[s_all] or 0) for r in data)
order = sorted(range(len(data)), key=lambda k: -float(data[k][s_all] or 0))
print("total samples", tot, "| stall columns:", [h for h, _ in stall_cols][:30])
for k in order[:top]:
    r = data[k]
    reasons = sorted(((float(r[i] or 0), h) for h, i in stall_cols if r[i] not in ("", "-")), reverse=True)[:3]
    print("%5.1f%%  #%-5d %-70s exec %-8s %s" % (100 * float(r[s_all] or 0) / tot, k, r[ci["Source"]].strip()[:70],
                                              r[ci["Instructions Executed"]],
                                              ", ".join("%s %.0f" % (h.replace("stall_", ""), v) for v, h in reasons if v > 0)))
