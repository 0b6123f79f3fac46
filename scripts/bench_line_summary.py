#!/usr/bin/env python
"""One-screen summary of a bench.py JSON line."""
import json
import sys

for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    c = d["config"]
    print(f"== {path}: N={d['n_gpus']} steps={d['steps']}")
    print(f"  value (every ray) {d['value']:.4g} tests/s, {d['ms_per_step']:.1f} ms/step; production {d['production']['value']:.4g} tests/s, "
          f"{d['production']['ms_per_step']:.1f} ms/step")
    e = d["e2e"]
    print(f"  e2e {e['value']:.4g} ({e['ms_per_step_over_device_ms_per_step']:.3f}x device), production {e['production']['value']:.4g} "
          f"({e['production']['ms_per_step_over_device_ms_per_step']:.3f}x); flux diff {e['flux_rel_diff_vs_device']:.1e}")
    print(f"  exchange: {c['exchange'][:90]}...; exchange_ms {c.get('exchange_ms') and {k: v for k, v in c['exchange_ms'].items() if k != 'what'}}")
    r = d["roofline"]
    print(f"  roofline frac {r.get('frac')}, issue {r.get('warp_issue_frac')}, instr/test {r.get('executed_lane_instr_per_test')}, "
          f"shares {r.get('step_shares')}")
    print(f"  production shares {d['roofline_production']['step_shares']}")
    print(f"  passing rays {c['passing_rays']}; total emissivity {c['total_emissivity']}; clocks {d['clocks']}")
