"""The JSON line bench.py printed on a B200 (committed under profiles/ by the run that measured it) against the contract the
driver reads: required keys, arithmetic relations between them, and the identities the path guarantees (same passing-ray
counts whatever the filter and the rank count; flux equal across transports)."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = {1: "r2y_bench_n1_default_run_final.json", 2: "r2w_bench_n2_default_run.json", 4: "r2y_bench_n4_default_run_final.json",
         8: "r2w_bench_n8_default_run.json"}


def load(n):
    with open(os.path.join(ROOT, "profiles", LINES[n])) as fh:
        return json.loads(fh.read().strip().splitlines()[-1])


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_contract_keys_and_relations(n):
    d = load(n)
    assert d["metric"] == "voxel-detector visibility tests/s" and d["unit"] == "tests/s" and d["n_gpus"] == n
    assert d["higher_is_better"] is True and d["scaling"] == "strong" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["warmup"] >= 3 and d["steps"] >= 1
    cfg = d["config"]
    assert cfg["workload"].startswith("configs[4]") and cfg["voxels"] == 2076 * 1230 * 384 and "model" not in cfg
    tests = cfg["tests_per_step"]
    assert tests == cfg["voxels"] * 600 * 4
    assert d["value"] == pytest.approx(tests / (d["ms_per_step"] * 1e-3), rel=1e-9)
    assert d["production"]["value"] == pytest.approx(tests / (d["production"]["ms_per_step"] * 1e-3), rel=1e-9)
    assert d["flux_time_s"]["full_step_s"] == pytest.approx(d["production"]["ms_per_step"] * 1e-3)
    e = d["e2e"]
    assert e["unit"] == "tests/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"] and e["production"]["value"] < d["production"]["value"]  # copies and calls cost something
    assert e["value"] > 0.9 * d["value"] and e["flux_rel_diff_vs_device"] < 1e-12
    assert d["gpu_launches"] == d["steps"] * (4 * cfg["launches_per_channel_per_gpu"] * 4 + 1)
    c = d["clocks"]
    assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"].startswith("issue") and r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9)
    assert 0.5 < r["frac"] < 1.0 and r["frac"] <= r["warp_issue_frac"] <= 1.0 and r["traffic"] > 0
    assert r["kernel_share_of_step"] == pytest.approx(r["step_shares"]["filter"]) and sum(r["step_shares"].values()) == pytest.approx(1.0)
    # what the path guarantees, whatever the filter and the number of ranks
    assert cfg["passing_rays"] == {"B": 2481985576, "C": 2341144738, "N": 4011227602, "O": 3439576266}
    assert cfg["visible_voxels"] == {"B": 123972425, "C": 106866897, "N": 186082303, "O": 192223253}
    assert cfg["flux_rel_diff_value_vs_production"] < 1e-12
    assert cfg["total_emissivity"] == {"B": "3.31e+16", "C": "3.80e+16", "N": "1.98e+16", "O": "1.67e+16"}
    if n == 1:
        cb = d["cpu_baseline"]
        assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["unit"] == "tests/s" and 1e6 < cb["value"] < 1e8
        s2 = d["roofline_stage2"]
        assert s2["bound"] == "hbm" and s2["frac"] == pytest.approx(s2["achieved"] / s2["peak"]) and 0.8 < s2["frac"] < 1.05
        assert 0.9 < s2["traffic"] / (10 * cfg["voxels_per_launch"]) < 1.1  # DRAM bytes = algorithmic bytes
        assert d["explicit_voxels"]["equal_to_the_lattice_path"] is True
        assert d["sweep_1000"]["max_rel_err_vs_reference_slices_0_499_999"] < 1e-12
    else:
        assert d["cpu_baseline"] is None and cfg["exchange"].startswith("peer memory")
        x = cfg["exchange_ms"]
        assert x["peer_barrier_then_fused_launch_ms"] < x["nccl_all_reduce_then_stage2_ms"] and x["flux_rel_diff_peer_vs_nccl"] < 1e-12


def test_strong_scaling_of_the_committed_runs():
    one = load(1)
    for n in (2, 4, 8):
        d = load(n)
        assert d["value"] / (n * one["value"]) > 0.95
        flux1, fluxn = one["config"]["flux_excit_recom_total"], d["config"]["flux_excit_recom_total"]
        for el in "BCNO":
            assert fluxn[el] == pytest.approx(flux1[el], rel=1e-12)
