"""Code-generation contract of the hot kernels, checked on the built library without a GPU (cuobjdump -sass / -res-usage):
the properties DESIGN.md section 4 relies on -- bulk-copy (TMA) tile loads completing on an mbarrier, one vector
red.global per tile node, no shared-memory float atomics (CAS loops on sm_100a), 128-bit shared/global accesses, register
budgets that give the designed occupancy, no spills -- so that a source change that silently loses one of them fails here
instead of showing up as a slower bench line."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
LIB = os.path.join(ROOT, "realtime-deformations_b200", "libmpm_b200.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None and not os.path.exists("/usr/local/cuda/bin/cuobjdump"),
                                reason="cuobjdump not available")


@pytest.fixture(scope="module")
def kernels():
    import mpm_b200
    mpm_b200.build.build()
    os.environ["PATH"] = os.environ.get("PATH", "") + ":/usr/local/cuda/bin"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    usage = {m.group(1): (int(m.group(2)), int(m.group(3)), int(m.group(4)))
             for m in re.finditer(r"Function (\S+):\n\s*REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", res)}
    out = {}
    for part in re.split(r"\n\s*Function : ", sass)[1:]:
        name = part.split("\n", 1)[0].strip()
        ops = []
        for ins in re.findall(r"/\*[0-9a-f]{4}\*/\s+(.*?);", part):
            t = re.sub(r"^@!?U?P\d+\s+", "", ins).split()
            if t:
                ops.append(t[0])
        out[name] = {"ops": ops, "regs": usage.get(name, (None, None, None))[0], "local": usage.get(name, (None, None, None))[2]}
    return out


def _one(kernels, *needles):
    hits = [k for k in kernels if all(n in k for n in needles)]
    assert len(hits) == 1, (needles, hits)
    return kernels[hits[0]]


def _count(k, prefix):
    return sum(o.startswith(prefix) for o in k["ops"])


def test_built_for_sm_100a_only():
    r = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", LIB], stdout=subprocess.PIPE, text=True)
    elfs = [l for l in r.stdout.splitlines() if "ELF file" in l]
    assert elfs and all("sm_100a" in l for l in elfs), r.stdout


def test_no_kernel_spills_to_local_memory(kernels):
    spilled = {k: v["local"] for k, v in kernels.items() if v["local"]}
    assert not spilled, spilled


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_p2g_tile_kernel_contract(kernels, mode):
    k = _one(kernels, f"k_p2g_tileILi{mode}ELi0ELb0ELi4E")
    assert k["regs"] <= 128, "2 CTAs of 256 threads per SM need <= 128 registers"
    assert _count(k, "REDG.E.ADD.F32x4") == 1, "one vector red.global.add.v4.f32 per tile node"
    assert not any("CAST" in o for o in k["ops"]), "no shared-memory float atomics (ATOMS.CAST.SPIN loops)"
    assert _count(k, "ATOMS") <= 2, "the counting sort uses one round of integer shared atomics (two unrolled particles)"
    assert _count(k, "LDS.128") >= 5 and _count(k, "STS.128") >= 6, "records and fold buffers move as 128-bit accesses"
    assert _count(k, "LDG.E.128") >= (6, 6, 12)[mode], "particle planes are read as float4 (2 particles x 4 / 3 / 6 planes)"
    assert _count(k, "SHFL.IDX") >= 48, "z-fold: 3 rounds x 16 values by warp shuffles (+ the warp-local scan look-ups)"
    assert _count(k, "BAR.SYNC") <= 7
    assert _count(k, "FFMA2") >= 40, "the accumulation loop runs on packed fp32 pairs"


def test_gather_kernel_contract(kernels):
    for flags in (2, 14, 30):           # staged gather; fused gather + advect + re-sort; + next substep's keys and histogram
        k = _one(kernels, f"k_g2p_tileILi{flags}ELi4EE")
        assert k["regs"] <= 128
        assert _count(k, "UBLKCP") == 4, "the tile arrives as 128 row-wise 64-byte bulk copies (TMA), 4 per lane"
        assert _count(k, "SYNCS.ARRIVE.TRANS64") == 1 and any("TRYWAIT" in o for o in k["ops"]), "mbarrier expect_tx + try_wait"
        assert _count(k, "LDS.128") == 64, "64 stencil nodes, one LDS.128 [base + immediate] each"
        assert _count(k, "BAR.SYNC") == 0, "warp-per-block: no CTA barrier"
        assert _count(k, "FFMA2") >= 200, "separable gather on packed fp32 pairs (252 FFMA2 + 84 FFMA per particle)"
    assert _count(_one(kernels, "k_g2p_tileILi30ELi4EE"), "MATCH.ANY") == 1, "warp-aggregated histogram of next substep's keys"


def test_fupdate_kernel_contract(kernels):
    exact, fast = _one(kernels, "k_fupdateILb1ELb0E"), _one(kernels, "k_fupdateILb1ELb1E")
    for k in (exact, fast):
        assert k["regs"] <= 64, "4 CTAs of 256 threads per SM"
        assert _count(k, "LDG.E.128") >= 7 and _count(k, "STG.E.128") == 7, "eight planes in, seven planes out, all float4"
        assert not any(o.startswith(("LDS", "STS", "BAR")) for o in k["ops"])
    assert len(fast["ops"]) < 0.5 * len(exact["ops"]), "the tolerance form is less than half the bit-faithful one"
    assert _count(fast, "MUFU") >= 4, "MUFU reciprocals / square roots instead of IEEE division sequences"


def test_fused_substep_p2g_carries_the_f_update(kernels):
    k = _one(kernels, "k_p2g_tileILi2ELi2ELb0ELi4E")
    assert k["regs"] <= 128 and _count(k, "STG.E.128") >= 7 and _count(k, "REDG.E.ADD.F32x4") == 1


def test_p2g_issues_l2_prefetches(kernels):
    """prefetch.global.L2 (CCTL.E.PF2) per thread and block: the planes the next block's derive phase reads (2 particles x the
    planes of the mode) and, with the in-kernel F-update, the five planes it reads (2 particles x 5). 5.09 -> 4.84 ms at 64 Mi."""
    assert _count(_one(kernels, "k_p2g_tileILi2ELi2ELb0ELi4E"), "CCTL.E.PF2") == 2 * 6 + 2 * 5
    assert _count(_one(kernels, "k_p2g_tileILi2ELi0ELb0ELi4E"), "CCTL.E.PF2") == 2 * 6
    assert _count(_one(kernels, "k_p2g_tileILi0ELi0ELb0ELi4E"), "CCTL.E.PF2") == 2 * 4      # momentum mode: planes 0..3
    assert _count(_one(kernels, "k_p2g_tileILi1ELi0ELb0ELi4E"), "CCTL.E.PF2") == 2 * 3      # force mode: planes 0, 4, 5


def test_peer_halo_p2g_issues_remote_vector_reds(kernels):
    for fu in (0, 1, 2):
        k = _one(kernels, f"k_p2g_tileILi2ELi{fu}ELb1ELi4E")
        assert _count(k, "REDG.E.ADD.F32x4") == 3, "local copy + the upper / lower neighbour's copy of a shared layer"
        assert k["regs"] <= 128 and not any("CAST" in o for o in k["ops"])


def test_quadratic_stencil_instantiations_skip_the_zero_column(kernels):
    """MpmParams.stencil = 1: the fused substep's tile kernels have W = 3 instantiations -- 27 instead of 64 tile reads per
    particle in the gather, 3 instead of 4 y-rows per visit in P2G's accumulation loop."""
    g3, g4 = _one(kernels, "k_g2p_tileILi30ELi3EE"), _one(kernels, "k_g2p_tileILi30ELi4EE")
    assert _count(g3, "LDS.128") == 27 and _count(g4, "LDS.128") == 64
    assert _count(g3, "FFMA2") < 0.5 * _count(g4, "FFMA2")
    p3, p4 = _one(kernels, "k_p2g_tileILi2ELi2ELb0ELi3E"), _one(kernels, "k_p2g_tileILi2ELi2ELb0ELi4E")
    assert _count(p3, "FFMA2") < _count(p4, "FFMA2") and p3["regs"] <= 128


def test_binning_uses_warp_aggregated_atomics(kernels):
    for name in ("k_bin_count", "k_bin_scatter"):
        k = _one(kernels, name)
        assert _count(k, "MATCH.ANY") == 4, "4 elements per thread, each warp-aggregated with match.any"


def test_implicit_integration_tile_kernels_contract(kernels):
    """The objective evaluation of mpm_time_integration in block-tile form (mpm_implicit.cuh): the trial field's gradient comes
    from the same TMA-staged separable gather as G2P (other pair weights, result to the aux array instead of particle planes),
    the gradient scatter accumulates in registers on packed pairs and issues one vector red per tile node."""
    g = _one(kernels, "k_g2p_tileILi34ELi4EE")
    assert g["regs"] <= 128 and _count(g, "UBLKCP") == 4 and _count(g, "LDS.128") == 64 and _count(g, "FFMA2") >= 200
    assert _count(g, "STG.E.128") == 3, "I + dt grad v: three float4 per particle, no particle plane written"
    s = _one(kernels, "k_imp_scatter_tile")
    assert s["regs"] <= 128 and _count(s, "REDG.E.ADD.F32x4") == 1 and _count(s, "FFMA2") == 48
    assert not any("CAST" in o for o in s["ops"]) and _count(s, "SHFL.IDX") >= 36
    for name in ("k_imp_stressILb0E", "k_imp_stressILb1E"):
        k = _one(kernels, name)
        assert _count(k, "DFMA") >= 40, "polar factor and energy density in double"
        assert k["regs"] <= 128
    b = _one(kernels, "k_imp_particlesILb1E")
    assert _count(b, "REDG.E.ADD.F32x4") >= 1 and _count(b, "REDG.E.ADD.F32x4") <= 64      # baseline: a vector red per (particle, node)
