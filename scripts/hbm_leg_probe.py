#!/usr/bin/env python
"""The Stage-2 voxel pass of bench.py's `roofline_stage2` leg in isolation (same synthetic 5e8-voxel channel), so that
ncu can capture com::weight_histogram_mm_kernel alone:
    ncu --set full --clock-control none -k regex:weight_histogram_mm -c 1 -o hist python scripts/hbm_leg_probe.py
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from co_monitor_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda")
n = 500_000_000
gen = torch.Generator(device=dev).manual_seed(0)
r = torch.randint(0, 600, (n,), dtype=torch.int32, device=dev, generator=gen).to(torch.int16)
w = torch.rand(n, dtype=torch.float64, device=dev, generator=gen)
w[torch.rand(n, device=dev, generator=gen) < 0.89] = 0.0
h = torch.zeros(_lib.COM_MM_BINS, dtype=torch.float64, device=dev)
for _ in range(3):
    h.zero_()
    _lib.check(lib.com_weight_histogram_mm(C.c_void_p(r.data_ptr()), C.c_void_p(w.data_ptr()), n, None,
                                           C.c_void_p(h.data_ptr()), None), "com_weight_histogram_mm")
torch.cuda.synchronize()
print(float(h.sum().item()))
