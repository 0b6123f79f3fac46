"""co_monitor_b200 - B200-native (sm_100a) implementation of co_monitor's two data-parallel stages.

Stage 1: visibility + geometric weight of every (plasma voxel, crystal point) ray
         (``co_monitor_b200.geometry.simulation.Simulation``).
Stage 2: line-emissivity integration (``co_monitor_b200.emissivity.reader.Emissivity``).

The compute path is hand-written CUDA behind a C-ABI (``include/comonitor_sm100.h``,
``co_monitor_b200/csrc``); there is no CPU fallback.
"""
__version__ = "0.1.0"
