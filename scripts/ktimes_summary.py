#!/usr/bin/env python
"""Condense `large_grid.py --kernel-times` output: per channel total, filter-, weight-only- and exact-kernel milliseconds."""
import json
import sys

for line in open(sys.argv[1]):
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    k = d.get("chunk_filter_ms_weight_ms_exact_ms_candidates_passing")
    if k is None:  # files written before the fp64 stage was split in two
        k = [[x[0], 0.0, x[1], x[2], x[3]] for x in d["chunk_filter_ms_exact_ms_candidates_passing"]]
    print(d["element"], d.get("mode", "production"), "total ms", round(d["ms"], 1), "filter", round(sum(x[0] for x in k), 1),
          "weight", round(sum(x[1] for x in k), 1), "exact", round(sum(x[2] for x in k), 1),
          "candidates", sum(x[3] for x in k), "passing", d["passing_rays"], "tests/s %.3g" % d["tests_per_s"])
