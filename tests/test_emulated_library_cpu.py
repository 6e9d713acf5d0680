"""The GPU parity tests' logic, run without a GPU: tests/emu builds the WHOLE library (C ABI, host orchestration, kernels)
for the host -- kernel launches become emulated launches where every CUDA thread is a fiber -- and this file points
tests/test_gpu_parity.py at that build in a subprocess. It pre-validates what the `-m gpu` run then confirms on hardware:
index arithmetic, barrier structure, bulk-copy byte counts, the host-side launch sequences of the staged and fused paths,
and the experimental kernel variants. It is test infrastructure: capi refuses the emulated build unless
MPM_B200_ALLOW_EMULATION=1 (set only here), and nothing in the package, bench.py or smoke() can select it."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

# the full-size configs and the 2-GPU test need the real device; graphs are refused by the fake runtime (the C++ adapter
# binary runs against the emulated build through LD_LIBRARY_PATH)
NEEDS_DEVICE = "not full_size and not two_gpus and not graph"
QUICK = "binning or stage_level or fupdate_kat or moving or volumes or staged_and_fused or parked or render or error_paths or ragged or empty_handle or rotated or adapter or peer_memory or stiff_sweep or deterministic"


@pytest.fixture(scope="module")
def emu_lib():
    import emu_build
    return emu_build.build()


def _run_gpu_suite(emu_lib, select, experimental=False, timeout=1500):
    env = dict(os.environ, MPM_B200_LIB=emu_lib, MPM_B200_ALLOW_EMULATION="1")
    env.pop("MPM_TEST_EXPERIMENTAL", None)
    if experimental:
        env["MPM_TEST_EXPERIMENTAL"] = "1"
    try:                      # the emulated kernels are single-threaded: spread the selected tests over a few worker processes
        import xdist          # noqa: F401
        workers = ["-n", str(max(1, min(4, (os.cpu_count() or 2) // 2)))]
    except ImportError:
        workers = []
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q", "-x",
                        "-p", "no:cacheprovider", "-k", f"({select}) and {NEEDS_DEVICE}"] + workers,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout, env=env, cwd=ROOT)
    return r


def test_loader_refuses_the_emulated_build(emu_lib):
    env = dict(os.environ, MPM_B200_LIB=emu_lib)
    env.pop("MPM_B200_ALLOW_EMULATION", None)
    code = "import sys; sys.path.insert(0, %r); import mpm_b200\ntry:\n    mpm_b200.Sim(20, 20, 20, 1)\nexcept mpm_b200.capi.MpmError as e:\n    print('REFUSED', e)" % ROOT
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=300)
    assert "REFUSED" in r.stdout and "no CPU fallback" in r.stdout, r.stdout[-2000:]


def test_gpu_parity_logic_on_emulated_library(emu_lib):
    """Stage-level parity with the reference at substeps 1 and 120, F-update and collision KATs, staged vs fused, parked
    particles, render buffers, error paths, ragged block occupancy, empty / interleaved handles."""
    r = _run_gpu_suite(emu_lib, QUICK)
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-4000:]


def test_implicit_time_integration_on_emulated_library(emu_lib):
    """tests/test_gpu_implicit.py (Energy vs the reference's golden values, gradient and minimiser vs the oracle, the
    semi-implicit substep sequence, error paths) against the emulated build."""
    env = dict(os.environ, MPM_B200_LIB=emu_lib, MPM_B200_ALLOW_EMULATION="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_implicit.py"), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500, env=env, cwd=ROOT)
    assert r.returncode == 0 and "13 passed" in r.stdout, r.stdout[-4000:]


def test_default_scene_trajectory_on_emulated_library(emu_lib):
    """400 substeps of the reference's default scene through the fused path, against the reference's own trajectory."""
    r = _run_gpu_suite(emu_lib, "default_scene_trajectory and variants0")
    assert r.returncode == 0 and "1 passed" in r.stdout, r.stdout[-4000:]


def test_experimental_variants_on_emulated_library(emu_lib):
    """The A/B kernel pairings (F-update as its own kernel, tile / baseline mixes) through the C ABI: stage-level parity,
    KATs, edge cases; and the synthetic-ball trajectory against the oracle with the F-update as its own kernel."""
    r = _run_gpu_suite(emu_lib, "binning or stage_level or fupdate_kat or parked or ragged", experimental=True)
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-4000:]
    r = _run_gpu_suite(emu_lib, "synthetic_ball and variants2", experimental=True)
    assert r.returncode == 0 and "1 passed" in r.stdout, r.stdout[-4000:]


@pytest.mark.parametrize("world,balanced,port", [(8, 0, 29731), (8, 1, 29732)])
def test_slab_protocol_with_the_real_library_over_gloo(emu_lib, world, balanced, port):
    """realtime-deformations_b200/multi.py (halo exchange, fixed-size migration messages, collective count checks) with
    the library's own slab entry points on 8 ranks: equal layers with particles driven across the slab boundaries
    (ranks that start empty receive particles), and the bench's balanced partition. Against the same scene in one domain."""
    env = dict(os.environ, MPM_B200_LIB=emu_lib, MPM_B200_ALLOW_EMULATION="1", OMP_NUM_THREADS="1", MPM_B200_PEER_HALO="0")      # the NCCL-message fallback
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "emu", "multi_check_emulated.py"), "64", "16384", "12", str(balanced)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0 and "MULTI_CHECK_OK" in r.stdout, r.stdout[-4000:]
    assert "unique ids 16384" in r.stdout


@pytest.mark.parametrize("world,balanced,port", [(3, 0, 29734), (8, 1, 29735)])
def test_peer_memory_halo_protocol_across_processes(emu_lib, world, balanced, port):
    """The peer-memory halo end to end across processes: handle exchange and mapping in multi.py
    (all_gather_object, unequal slab sizes, ranks that start empty), remote reds from P2G into the neighbours' grids and the
    device-side flag protocol under real concurrency -- the fake runtime backs "device" memory with POSIX shared memory so
    that cudaIpcOpenMemHandle maps another rank's grid (EMU_SHM_IPC=1). Against the same scene in one domain."""
    env = dict(os.environ, MPM_B200_LIB=emu_lib, MPM_B200_ALLOW_EMULATION="1", OMP_NUM_THREADS="1", EMU_SHM_IPC="1", MPM_B200_PEER_HALO="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "emu", "multi_check_emulated.py"), "64", "16384", "12", str(balanced)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0 and "MULTI_CHECK_OK" in r.stdout and "peer-memory halo: True" in r.stdout, r.stdout[-4000:]


def test_migration_overflow_stops_every_rank_together(emu_lib):
    """Failure path of the slab protocol with the real library: migration messages too small -> the library flags the
    overflow, and the collective check raises on all ranks in the same substep (no rank is left waiting)."""
    env = dict(os.environ, MPM_B200_LIB=emu_lib, MPM_B200_ALLOW_EMULATION="1", OMP_NUM_THREADS="1", MPM_B200_PEER_HALO="0")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nproc-per-node", "3", "--master-addr", "127.0.0.1",
                        "--master-port", "29733", os.path.join(ROOT, "tests", "emu", "multi_check_emulated.py"), "64", "16384", "12", "2"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0 and "OVERFLOW_HANDLED" in r.stdout, r.stdout[-4000:]


@pytest.mark.skipif(os.environ.get("MPM_EMU_FULL") != "1", reason="the whole emulated suite takes ~5 min: MPM_EMU_FULL=1")
def test_whole_gpu_suite_on_emulated_library(emu_lib):
    r = _run_gpu_suite(emu_lib, "test_", experimental=True, timeout=3400)
    assert r.returncode == 0, r.stdout[-4000:]


def test_randomized_scenes_against_the_oracle_on_emulated_library(emu_lib):
    """tests/emu/fuzz_vs_oracle.py: non-cubic grids, particles on cell faces / at the clamp / outside the grid, crowded
    cells, moving rotated colliders, random FE / FP, every kernel variant -- a dozen seeded cases per run."""
    env = dict(os.environ, MPM_B200_LIB=emu_lib, MPM_B200_ALLOW_EMULATION="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "fuzz_vs_oracle.py"), "14", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1200, env=env, cwd=ROOT)
    assert r.returncode == 0 and "FUZZ_OK 14 cases" in r.stdout, r.stdout[-4000:]
