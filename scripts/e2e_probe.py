import sys, time, ctypes as C, contextlib, io
sys.path.insert(0, '.')
import numpy as np, torch
from co_monitor_b200 import _lib
from co_monitor_b200.emissivity.engine import LineTables
from co_monitor_b200.emissivity.kinetic_profiles import TwoGaussSumProfile
from co_monitor_b200.geometry.engine import lattice_struct, make_channel
from co_monitor_b200.geometry.mesh_calculation import CuboidMesh
from co_monitor_b200.geometry.reff import interpolate_reff
lib = _lib.load()
with contextlib.redirect_stdout(io.StringIO()):
    lat = CuboidMesh(10).lattice
    sc = make_channel("C", 10, 30, 20, lattice=lat)
tables = LineTables("C", "Z5", 33.7, ["EXCIT", "RECOM"])
reff = interpolate_reff(lat, 15)
prof = np.ascontiguousarray(TwoGaussSumProfile([7e13,0,0.37,9.8e12,0.5,0.11],[1870,0,0.155,210,0.38,0.07]).profiles_df[["Reff [m]","T_e [eV]","n_e [m-3]"]].to_numpy())
n = lat.n_points
w_h = torch.empty(n, dtype=torch.float64).pin_memory(); n_h = torch.empty(n, dtype=torch.int32).pin_memory()
st = _lib.ComStats(); ls = lattice_struct(lat); flux = np.zeros(3)
def stage1():
    rc = lib.com_visibility_weights_host(C.byref(sc.desc), C.byref(ls), None, 0, n, C.c_void_p(w_h.data_ptr()), C.c_void_p(n_h.data_ptr()), None, 0, None, C.byref(st)); assert rc == 0, lib.com_last_error()
def gather():
    vis = np.flatnonzero(n_h.numpy()); return np.ascontiguousarray(reff[vis]), np.ascontiguousarray(w_h.numpy()[vis])
def stage2(r, w):
    rc = lib.com_emissivity_integrate_host(C.byref(tables.desc), _lib.as_double_ptr(prof), len(prof), 0.02, _lib.as_double_ptr(r), _lib.as_double_ptr(w), len(r), _lib.as_double_ptr(flux), None, None, None); assert rc == 0
for _ in range(3):
    stage1(); r, w = gather(); stage2(r, w)
for name, fn in (("stage1", stage1), ("gather", gather), ("stage2", lambda: stage2(r, w))):
    t = time.perf_counter()
    for _ in range(10): fn()
    print(name, (time.perf_counter() - t) / 10 * 1e3, "ms")
print(flux)
