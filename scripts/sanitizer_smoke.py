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
# whole-volume pipeline (interval records, expand, record-based weight kernel, histogram pass, fused Stage 2) in ragged
# launches, and the exchange-fused-into-Stage-2 kernel on three "rank" buffers
from co_monitor_b200 import _lib
from co_monitor_b200.geometry.reff import source_volume
from co_monitor_b200.volume import VolumePipeline, flux_from_peer_buffers
prof = pipe._profile.cpu().numpy()
bufs = []
for r in range(3):
    vp = VolumePipeline(lat.shard(3, r, cols_per_block=4), [scene], [tables], prof, 2.0, chunk=1001)
    vp.volume.set_reff_resampled(source_volume(30))
    vp.step()
    bufs.append(vp.packed.clone())
    vp.close()
fl = torch.zeros((1, 1, 3), dtype=torch.float64, device="cuda")
bd = torch.zeros(1, dtype=torch.float64, device="cuda")
hs = torch.zeros((1, _lib.COM_MM_BINS), dtype=torch.float64, device="cuda")
cs = torch.zeros((1, _lib.N_STATS), dtype=torch.float64, device="cuda")
flux_from_peer_buffers([tables], torch.from_numpy(prof[None].copy()).cuda(), [b.data_ptr() for b in bufs], _lib.N_STATS, 0.02,
                       fl, bd, hs, cs)
# per-ray lattice kernels (run-level cull on / off) and packed explicit voxels
for kw in ({"intervals": False}, {"intervals": False, "run_cull": False}):
    with contextlib.redirect_stdout(io.StringIO()):
        sc2 = make_channel("C", 10, 7, 13, lattice=lat, **kw)
    r4 = sc2.visibility(lattice=lat)
    assert torch.equal(r4.nrays, res.nrays)
    sc2.close()
torch.cuda.synchronize()
print("sanitizer smoke ok", res.stats, flux, sw[0], fl.cpu().numpy().ravel())
