"""Host emulation of the CUDA tile kernels (tests/emu): the kernel sources under csrc/ are compiled by g++ against a
shim (threads = fibers resumed in seeded random order, real barriers, emulated bulk copies with byte accounting) and the kernel VARIANTS are
checked against each other -- in particular the experimental ones (linear gather tile, packed fp32 pairs, F-update
inside P2G) against the default kernels that the GPU tests validate against the oracle and the reference.
Test infrastructure only: nothing in the package can reach this code."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
BUILD = os.path.join(EMU, "_build")
CSRC = os.path.join(ROOT, "realtime-deformations_b200", "csrc")


def _build(name):
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, name)
    srcs = [os.path.join(EMU, name + ".cpp"), os.path.join(EMU, "cuda_emu.h")] + \
           [os.path.join(CSRC, f) for f in ("mpm_math.cuh", "mpm_kernels.cuh", "mpm_tile_kernels.cuh")]
    if not os.path.exists(exe) or any(os.path.getmtime(s) > os.path.getmtime(exe) for s in srcs):
        cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-ffp-contract=off", "-pthread", "-I/usr/local/cuda/include", "-I" + CSRC,
               "-I" + os.path.join(ROOT, "include"), "-include", os.path.join(EMU, "cuda_emu.h"), "-x", "c++",
               os.path.join(EMU, name + ".cpp"), "-o", exe]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, "host build of the kernel sources failed:\n" + r.stdout[-4000:]
    return exe


@pytest.fixture(scope="module")
def harness():
    return _build("emu_harness")


def test_device_code_reproduces_reference_kats_on_the_host(tmp_path):
    """The reference's known-answer vectors through the device routines compiled for the host: weights, bodyCollision
    (static and moving colliders, as a function and through k_grid_update), updateDeformationGradient (k_fupdate)."""
    import numpy as np
    k = np.load(os.path.join(ROOT, "tests", "golden", "kat_functions.npz"))
    for n in ("weights_x", "weights_w", "collide_pos", "collide_vel", "collide_out", "collide_moving_out", "fupdate_in", "fupdate_out"):
        np.ascontiguousarray(k[n], np.float32).tofile(str(tmp_path / (n + ".f32")))
    for n in ("colliders", "colliders_moving"):       # reference dump rows -> world_to_local[16], half[3], velocity[3]
        raw = k[n]
        np.ascontiguousarray(np.concatenate([raw[:, 13:29], raw[:, 0:3], raw[:, 10:13]], 1), np.float32).tofile(str(tmp_path / (n + ".f32")))
    r = subprocess.run([_build("emu_kat"), str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "all emulated known-answer tests passed" in r.stdout, r.stdout[-4000:]
    assert r.stdout.count("ok  ") == 7


@pytest.mark.parametrize("seed,fast_div", [(1, 1), (7, 0)])
def test_kernel_variants_agree_under_host_emulation(harness, seed, fast_div):
    r = subprocess.run([harness, str(seed), str(fast_div)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "all emulation checks passed" in r.stdout, r.stdout[-4000:]
    assert r.stdout.count("\nok  ") >= 14          # every check ran


def test_p2g_shared_memory_wavefront_model(harness):
    """Bank model of the emulator on the benchmark layout (8 particles in every cell): the aligned record walk measured
    in round 1 reproduces most of ncu's bank-conflict count, the rotated walk (the default) is conflict-free in phase 1."""
    r = subprocess.run([harness, "1", "1", "smem"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and "all emulation checks passed" in r.stdout, r.stdout[-4000:]
    assert r.stdout.count("\nok  ") == 4


def test_product_cannot_reach_the_emulator():
    """The shim and harness live under tests/ only; the package and the C ABI never mention them."""
    for base in (os.path.join(ROOT, "realtime-deformations_b200"), os.path.join(ROOT, "include"), os.path.join(ROOT, "adapter")):
        for dirpath, _, files in os.walk(base):
            for f in files:
                if f.endswith((".py", ".cu", ".cpp", ".h")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert "cuda_emu.h" not in txt and "emu_harness" not in txt, os.path.join(dirpath, f)
    # the kernel sources only carry the MPM_HOST_EMU switch, which nvcc builds never define
    assert "MPM_HOST_EMU" not in open(os.path.join(ROOT, "realtime-deformations_b200", "build.py")).read()
