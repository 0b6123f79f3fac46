#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel from an .ncu-rep taken with --import-source on (compiled -lineinfo).

    python scripts/ncu_source_lines.py prof.ncu-rep <kernel regex> [top N]
"""
import csv
import subprocess
import sys


def main(path, kernel, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                          f"regex:{kernel}"], capture_output=True, text=True).stdout
    fname, lines, total, hdr = "", {}, 0, None
    for r in csv.reader(out.splitlines()):
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
            ii, it, ism = r.index("Instructions Executed"), r.index("Thread Instructions Executed"), r.index("# Samples")
        elif hdr and len(r) > 10 and r[0].isdigit():  # a source line with its aggregated counters
            try:
                n, t, sm = int(r[ii]), int(r[it]), int(r[ism])
            except ValueError:
                continue
            key = (fname, int(r[0]))
            a = lines.setdefault(key, [0, 0, 0, r[1].strip()[:100]])
            a[0] += n
            a[1] += t
            a[2] += sm
            total += n
    print(f"# {kernel}: {total} warp instructions with line info")
    for (f, ln), (n, t, sm, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * n / max(total, 1):5.1f}%  lanes {t / max(n, 1):4.1f}  samples {sm:6d}  {f}:{ln}  {src}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
