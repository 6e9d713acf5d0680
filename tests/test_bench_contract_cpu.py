"""CPU test of the bench contract's reference arm: `bench.py --impl reference` needs no GPU, must print ONE JSON line
with the contract's keys, and under torchrun only rank 0 may do work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "3"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "particle-updates/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("particle-updates/sec") and d["value"] > 1e4 and d["steps"] == 2 and d["warmup"] == 3
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_without_work():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_roofline_block_assembly():
    """The roofline object is assembled by a pure function: exercise the timing layouts without a GPU."""
    sys.path.insert(0, ROOT)
    import bench
    n, a = 67108864.0, 9184376.0
    # default: the F-update runs inside the P2G kernel (no F-update time of its own)
    r = bench.roofline_block(7.7, n, a, 1, [0.22, 0.03, 5.08, 0.11, 2.1, 0.003, 7.6, -1.0], "config5")
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert r["dominant_kernel"] == "p2g+fupdate" and r["kernel_ms_last_substep"]["g2p_gather"] == 2.1
    assert abs(r["achieved"] - (272 * n + 80 * a) / 7.7e-3 / 1e9) < 1.0
    assert r["dominant_kernel_achieved_gbs"] == round((280 * n + 16 * a) / 5.08e-3 / 1e9, 1)
    assert abs(r["achieved_fp32_tflops"] - 4500 * n / 7.7e-3 / 1e12) < 0.01
    # p2g_variant 2: F-update and gather as kernels of their own, timed apart
    r1 = bench.roofline_block(9.92, n, a, 1, [0.48, 0.03, 4.21, 0.11, 5.01, 0.003, 9.85, 2.44], "config5", fupd_in_p2g=False)
    assert r1["dominant_kernel"] == "p2g" and r1["kernel_ms_last_substep"]["g2p_gather"] == round(5.01 - 2.44, 4)
    assert r1["dominant_kernel_achieved_gbs"] == round((88 * n + 16 * a) / 4.21e-3 / 1e9, 1)
    # side-stream overlap: only the combined G2P time exists
    r2 = bench.roofline_block(9.92, n, a, 1, [0.48, 0.03, 4.21, 0.11, 5.01, 0.003, 9.85, -1.0], "config5", fupd_in_p2g=False)
    assert r2["dominant_kernel"] == "g2p(fupdate+gather)" and "fupdate" not in r2["kernel_ms_last_substep"]
    # 8 GPUs: per-GPU achieved against a per-GPU peak
    r8 = bench.roofline_block(2.0, n, a, 8, [0.1, 0.01, 0.55, 0.4, 0.65, 0.3, 1.7, -1.0], "config5")
    assert abs(r8["achieved"] - (272 * n + 80 * a) / 2.0e-3 / 1e9 / 8) < 1.0
    # the profiler capture is only quoted for the kernel sources it was taken from
    assert r["traffic"] is None or isinstance(r["traffic"], (int, float))


def test_expected_id_checksums_and_partition():
    """bench.py's invariants: the id checksums of an intact particle set, and the particle-count-balanced slab partition of a
    two-ball scene keeping every particle exactly once."""
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    import mpm_b200
    from importlib import import_module
    multi = import_module("realtime-deformations_b200.multi")
    s, h = bench.expected_id_sums(1000)
    assert s == 999 * 1000 // 2 and 0 <= h < 1 << 64
    sc = mpm_b200.scenes.snowball_collision(grid=64, n=1 << 15)
    full, layers, ranges = multi.partition_scene(sc, 2)
    assert ranges[0][0] == 0 and ranges[0][1] == ranges[1][0] and ranges[1][1] == sc["n"]
    assert layers[0][1] == layers[1][0] and abs((ranges[0][1] - ranges[0][0]) - sc["n"] // 2) < 0.1 * sc["n"]
    lay = multi.particle_block_layers(full["pos"], full["h"])
    assert (np.diff(lay) >= 0).all() and lay[ranges[0][1] - 1] < layers[0][1] <= lay[ranges[1][0]]
    assert np.array_equal(np.sort(full["pos"].view([("", np.float32)] * 3).ravel()), np.sort(sc["pos"].view([("", np.float32)] * 3).ravel()))


def test_port_openmp_sample_block():
    """cpu_baseline.port: the restatement single-threaded and with OpenMP on all host cores (SURVEY 8d ii), here on a miniature."""
    sys.path.insert(0, ROOT)
    import bench
    import mpm_b200
    out = bench.port_openmp_sample(mpm_b200, steps=2, grid=32, n=8192)
    assert out["unit"] == bench.UNIT and "8192 particles" in out["sample"]
    assert out["single_thread"]["threads"] == 1 and out["single_thread"]["value"] > 0
    assert out["openmp_all_cores"]["threads"] >= 1 and out["openmp_all_cores"]["value"] > 0
