"""Optional physics extensions of Stage 1, each off by default: the ECRH shields as
circular stops and the segment-direction check on aperture hits.  With the switches off nothing changes (the golden
parity tests cover that); with them on the CUDA path is compared ray-count-exactly with the NumPy statements in
oracle/extensions.py."""
import contextlib
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(el, spacing=10, **kw):
    from co_monitor_b200.geometry.engine import make_channel
    from co_monitor_b200.geometry.mesh_calculation import CuboidMesh

    with contextlib.redirect_stdout(io.StringIO()):
        mesh = CuboidMesh(spacing)
        scene = make_channel(el, 10, 30, 20, lattice=mesh.lattice, **kw)
    return mesh, scene


@pytest.mark.parametrize("el,lo", [("C", 131000), ("N", 140000)])
def test_shipped_ecrh_shields_are_clear_of_the_beam(el, lo):
    """The four shields of the database do not vignette the channels: as stops they change nothing, on the GPU and in
    the oracle alike (every accepted ray passes within 35 mm of the axis of a 67.5 mm stop)."""
    import torch

    from co_monitor_b200.geometry.radiation_shield import ecrh_shields
    from oracle import extensions as ext
    from oracle import geometry_setup as gs
    from oracle import stage1 as o1

    mesh, plain = _setup(el)
    _, shielded = _setup(el, shields=ecrh_shields(el))
    lat = mesh.lattice
    a = plain.visibility(lattice=lat)
    b = shielded.visibility(lattice=lat)
    assert torch.equal(a.nrays, b.nrays) and b.stats["shield_blocked_rays"] == 0
    assert b.stats["certain_candidates"] == 0 < a.stats["certain_candidates"]  # stops are tested in fp64 for every ray
    idx = np.arange(lo, lo + 600)
    db = gs.load_db()
    ref = ext.run_chunk_physical(o1.build_scene(el, 10, 30, 20), gs.voxel_mesh(db, 10, flat_indices=idx), ext.channel_shields(db, el))
    assert ref.shield_blocked == 0 and ref.nrays.sum() > 100
    assert np.array_equal(b.nrays.cpu().numpy()[idx], ref.nrays)
    plain.close()
    shielded.close()


@pytest.mark.parametrize("n_sides", [360, 0])
def test_narrow_stops_against_the_oracle(n_sides):
    """Stops narrowed to 18 mm do cut into the beam: per-voxel ray counts, weights and the number of stopped rays equal the
    oracle's (regular 360-gon as the reference builds the shield; n_sides = 0 is a true circle and has no oracle hull)."""
    from co_monitor_b200.geometry.radiation_shield import ecrh_shields
    from oracle import extensions as ext
    from oracle import geometry_setup as gs
    from oracle import stage1 as o1

    el, lo, hi = "O", 131000, 131900
    stops = [dict(s, radius=18.0, n_sides=n_sides) for s in ecrh_shields(el)]
    mesh, scene = _setup(el, shields=stops)
    res = scene.visibility(lattice=mesh.lattice, begin=lo, end=hi)
    db = gs.load_db()
    osc = o1.build_scene(el, 10, 30, 20)
    vox = gs.voxel_mesh(db, 10, flat_indices=np.arange(lo, hi))
    if n_sides:
        verts = []
        for ch_sh in (("upper chamber", "1st shield"), ("upper chamber", "2nd shield")):
            db2 = {"ECRH shield": {ch_sh[0]: {ch_sh[1]: dict(db["ECRH shield"][ch_sh[0]][ch_sh[1]], radius=18.0)}}}
            verts.append(ext.shield_vertices(db2, *ch_sh))
        ref = ext.run_chunk_physical(osc, vox, verts)
        assert ref.shield_blocked > 50 and ref.nrays.sum() > 200
        assert np.array_equal(res.nrays.cpu().numpy(), ref.nrays)
        np.testing.assert_allclose(res.weight.cpu().numpy(), ref.weight, rtol=1e-12, atol=1e-300)
        assert res.stats["shield_blocked_rays"] == ref.shield_blocked
    else:
        plain = o1.run_chunk(osc, vox)
        assert 0 < res.stats["passing_rays"] < plain.nrays.sum()
        assert res.stats["passing_rays"] + res.stats["shield_blocked_rays"] == plain.nrays.sum()
    scene.close()


def test_segment_direction_check_against_the_oracle():
    """Voxels placed between the crystal and the collimator: their lines to the crystal points meet slits, port and stops
    only beyond the voxel, so the reference's infinite-line test accepts rays the physical ray cannot make.  With
    COM_PHYS_SEGMENT_CHECK those are rejected; without it the reference's (unphysical) answer is reproduced."""
    import torch

    from oracle import extensions as ext
    from oracle import geometry_setup as gs
    from oracle import stage1 as o1

    el = "C"
    mesh, plain = _setup(el)
    _, checked = _setup(el, segment_check=True)
    db = gs.load_db()
    c0 = np.array(db["dispersive element"]["element"][el]["crystal central point"])
    far = gs.voxel_mesh(db, 10, flat_indices=np.arange(131000, 131800))
    near = c0 + 0.03 * (far - c0)  # ~100 mm in front of the crystal, the collimator sits at ~160 mm
    vox = np.ascontiguousarray(np.concatenate((near, far[:300])))
    vt = torch.from_numpy(vox).cuda()
    osc = o1.build_scene(el, 10, 30, 20)
    want_ref = o1.run_chunk(osc, vox)
    want_chk = ext.run_chunk_physical(osc, vox, segment_check=True)
    assert want_chk.segment_rejected > 100 and want_chk.nrays[800:].sum() == want_ref.nrays[800:].sum() > 0
    assert want_chk.nrays[:800].sum() == 0 < want_ref.nrays[:800].sum()
    got_ref = plain.visibility(voxels=vt)
    got_chk = checked.visibility(voxels=vt)
    assert np.array_equal(got_ref.nrays.cpu().numpy(), want_ref.nrays)
    assert np.array_equal(got_chk.nrays.cpu().numpy(), want_chk.nrays)
    np.testing.assert_allclose(got_chk.weight.cpu().numpy(), want_chk.weight, rtol=1e-12, atol=1e-300)
    assert got_chk.stats["segment_rejected_rays"] == want_chk.segment_rejected
    assert got_ref.stats["segment_rejected_rays"] == 0
    # on the lattice itself every aperture lies between voxel and crystal: the check changes nothing
    a = plain.visibility(lattice=mesh.lattice, begin=120000, end=150000)
    b = checked.visibility(lattice=mesh.lattice, begin=120000, end=150000)
    assert torch.equal(a.nrays, b.nrays) and b.stats["segment_rejected_rays"] == 0
    plain.close()
    checked.close()


def test_simulation_switches():
    from co_monitor_b200.geometry.simulation import Simulation

    base = Simulation("B", 10, 10, 30, 20, plot=False, verbose=False)
    ext_sim = Simulation("B", 10, 10, 30, 20, plot=False, verbose=False, ecrh_shields=True, segment_check=True)
    assert ext_sim.stats["passing_rays"] == base.stats["passing_rays"] == 652069
    assert ext_sim.stats["shield_blocked_rays"] == 0 == ext_sim.stats["segment_rejected_rays"]
    assert ext_sim.ddf.equals(base.ddf) or np.allclose(ext_sim.ddf.to_numpy(), base.ddf.to_numpy(), rtol=1e-13)
