#!/usr/bin/env python
"""Condense `large_grid.py --kernel-times` output: per channel total, filter-kernel and fp64-kernel milliseconds."""
import json
import sys

for line in open(sys.argv[1]):
    d = json.loads(line)
    k = d["chunk_filter_ms_exact_ms_candidates_passing"]
    print(d["element"], "total ms", round(d["ms"], 1), "filter", round(sum(x[0] for x in k), 1), "fp64", round(sum(x[1] for x in k), 1),
          "candidates", sum(x[2] for x in k), "passing", d["passing_rays"], "tests/s %.3g" % d["tests_per_s"])
