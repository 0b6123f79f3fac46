#!/usr/bin/env python
"""Per-step kernel shares from an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file ...` of bench.py):
steps are delimited by the Stage-2 launch (histogram_flux_kernel).  Times under ncu are cold-cache and serialised - compare
the SHARES with bench.py's `roofline.step_shares` / `roofline_production.step_shares`, not the absolutes."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
ik, iv = H.index("Kernel Name"), H.index("Metric Value")
steps, cur = [], collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= iv:
        continue
    name = r[ik].split("(")[0].replace("com::", "").split("<")[0]
    cur[name] = cur.get(name, 0.0) + float(r[iv].replace(",", "")) / 1e6
    if name.startswith("histogram_flux_kernel"):
        steps.append(cur)
        cur = collections.OrderedDict()
for k, st in enumerate(steps):
    tot = sum(st.values())
    kind = "every-ray (headline) step" if "filter_kernel_lattice" in st else "shipped step"
    print(f"step {k}: {kind}, {tot:.1f} ms of kernels: " +
          ", ".join(f"{n} {v:.1f} ms ({100 * v / tot:.1f} %)" for n, v in st.items() if v / tot > 0.002))
