"""Small end-to-end run for compute-sanitizer (memcheck): Stage 1 (lattice + explicit voxels), compaction, Stage 2."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from co_monitor_b200.emissivity.engine import LineTables
from co_monitor_b200.emissivity.kinetic_profiles import TwoGaussSumProfile
from co_monitor_b200.geometry.engine import make_channel
from co_monitor_b200.geometry.mesh_calculation import CuboidMesh
from co_monitor_b200.geometry.reff import interpolate_reff
from co_monitor_b200.pipeline import ChannelPipeline
with contextlib.redirect_stdout(io.StringIO()):
    mesh = CuboidMesh(30)
    lat = mesh.lattice
    scene = make_channel("C", 10, 7, 13, lattice=lat)
    reff = interpolate_reff(lat, 30)
res = scene.visibility(lattice=lat, want_rays=True)
pts = torch.from_numpy(mesh.outer_cube_mesh).cuda()
res2 = scene.visibility(voxels=pts, begin=100, end=lat.n_points - 7)
assert torch.equal(res.nrays[100:lat.n_points - 7], res2.nrays)
sh = lat.shard(3, 1, 5)
res3 = scene.visibility(lattice=sh)
tables = LineTables("C", "Z5", 33.7, ["EXCIT", "RECOM"])
pipe = ChannelPipeline(scene, tables, lat, torch.from_numpy(reff).cuda(), 0, lat.n_points)
pipe.set_profile(TwoGaussSumProfile([7e13, 0, 0.37, 9.8e12, 0.5, 0.11], [1870, 0, 0.155, 210, 0.38, 0.07]).profiles_df[["Reff [m]", "T_e [eV]", "n_e [m-3]"]].to_numpy())
flux = pipe.launch().cpu().numpy()
hist, b = tables.weight_histogram(np.arange(0, 0.539, 0.001), reff, res.weight)
sw = tables.sweep(np.stack([pipe._profile.cpu().numpy()] * 3), hist, 0.02).cpu().numpy()
torch.cuda.synchronize()
print("sanitizer smoke ok", res.stats, flux, sw[0])
