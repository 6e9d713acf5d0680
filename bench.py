#!/usr/bin/env python
"""Headline benchmark: particle-updates/s per MPM substep on the 64 Mi-particle / 512^3 snow-slab scene
(BASELINE.json config 5), as absolute numbers and as a fraction of the measured HBM roofline.

    python bench.py --gpus 1 --steps 20 --warmup 5                 # this framework (CUDA, sm_100a)
    torchrun --nproc-per-node N bench.py --gpus N ...              # slab-decomposed over N GPUs (strong scaling)
    python bench.py --impl reference --gpus 1 --steps K --warmup W # the reference's own CPU path on the host cores

One JSON line on stdout (rank 0). See DESIGN.md "Measurement" for how each field is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-updates/sec per substep"
UNIT = "particle-updates/s"
ALG_BYTES_PER_PARTICLE = 272      # read 35 + write 33 fp32 of particle state once per substep
ALG_BYTES_PER_NODE = 80           # 16 B node x 5 touches, active nodes only


def measured_peak():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json: hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while a timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower() == "active":
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_sample_scene(mpm_b200):
    """Bounded sample of the slab workload for the CPU arms: same generator, 8 ppc slab, 32^3 grid (the largest the
    reference's own class can allocate comfortably: its WeightStorage holds I*J*K*N floats)."""
    sc = mpm_b200.scenes.snow_slab(grid=32, n=8192)
    return sc


def run_reference_binary(sc, steps, nproc):
    """Times the UNMODIFIED reference (oracle/_ref/ref_mpm) on `nproc` host cores: one independent replica of the
    sample per core (the reference is single-threaded by construction). Returns aggregate particle-updates/s."""
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_mpm")
    d = tempfile.mkdtemp()
    np.concatenate([sc["pos"], sc["vel"], sc["mass"][:, None]], 1).astype(np.float32).tofile(d + "/p.f32")
    t = -sc["w2l"][0][12:15]
    np.array([t[0], t[1], t[2], 0.0, *sc["half"][0]], np.float32).tofile(d + "/c.f32")
    I, J, K = sc["dims"]
    cmd = [ref, "--grid", str(I), str(J), str(K), "--n", str(sc["n"]), "--load", d + "/p.f32", "--colliders", d + "/c.f32",
           "--steps", str(steps), "--quiet", "--bench"]
    t0 = time.time()
    # the class's constructor also cudaMallocs its (dead) weight table (hpp:149-156; failure is non-fatal there): hide the GPU
    # so that a replica per host core does not open a CUDA context per core -- this is the CPU path that is being timed
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    procs = [subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env) for _ in range(nproc)]
    outs = [p.communicate()[0] for p in procs]
    wall = time.time() - t0
    secs = []
    for o in outs:
        for line in o.splitlines():
            if line.startswith("{"):
                secs.append(json.loads(line)["seconds"])
    if len(secs) != nproc:
        raise RuntimeError("reference binary failed: " + outs[0][-300:])
    return nproc * sc["n"] * steps / max(secs), wall, max(secs)


def run_oracle_port(sc, steps, threads=1):
    """The CPU restatement (oracle/mpm_oracle.c); threads > 1 uses OpenMP with per-thread grids."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as op
    I, J, K = sc["dims"]
    o = op.Oracle(I, J, K, sc["n"], op.default_params(h=float(sc["h"])), threads=threads)
    o.set_state(op.initial_state(sc["pos"], sc["vel"], sc["mass"]))
    o.rasterize(); o.volumes()
    cols, nc = op.make_colliders(sc["w2l"], sc["half"], sc["cvel"])
    t0 = time.time()
    o.substep(float(sc["dt"]), cols, nc, steps)
    sec = time.time() - t0
    return sc["n"] * steps / sec, sec


def port_openmp_sample(mpm_b200, steps=3, grid=128, n=1 << 20):
    """the CPU restatement (oracle/mpm_oracle.c) single-threaded and with OpenMP on all host cores, on a
    1 Mi-particle / 128^3 sample of the slab workload (the reference's own class cannot allocate this size). Reported next to
    the reference's number, never used as the baseline value."""
    sc = mpm_b200.scenes.snow_slab(grid=grid, n=n)
    out = {"sample": f"snow_slab sample: {sc['n']} particles, {grid}^3 grid, {steps} substeps", "unit": UNIT}
    # (every OpenMP thread scatters into a private dense grid that is reduced afterwards: beyond a few dozen threads the
    # reduction, not the scatter, is what is timed, so the thread count is capped)
    for name, th in (("single_thread", 1), ("openmp_all_cores", min(os.cpu_count() or 1, 32))):
        v, sec = run_oracle_port(sc, steps, th)
        out[name] = {"value": v, "threads": th, "seconds": sec}
    return out


def cpu_baseline(mpm_b200, budget_steps=150):
    out = cpu_baseline_reference(mpm_b200, budget_steps)
    try:
        out["port"] = port_openmp_sample(mpm_b200)
    except Exception as e:
        out["port"] = {"error": str(e)}
    return out


def cpu_baseline_reference(mpm_b200, budget_steps=150):
    sc = cpu_sample_scene(mpm_b200)
    sample = f"snow_slab sample: {sc['n']} particles, 32^3 grid, 8 ppc, {budget_steps} substeps, default gravity"
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_mpm")
    if os.path.exists(ref):
        try:
            v, wall, sec = run_reference_binary(sc, budget_steps, 1)
            return {"value": v, "unit": UNIT, "cores": 1, "kind": "reference", "sample": sample, "seconds": sec,
                    "cpu_model": cpu_model(), "host_cores": os.cpu_count()}
        except Exception as e:   # fall through to the port
            sample += f" (reference binary unusable: {e})"
    v, sec = run_oracle_port(sc, budget_steps)
    return {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample, "seconds": sec,
            "cpu_model": cpu_model(), "host_cores": os.cpu_count()}


def reference_arm(args):
    import mpm_b200
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sc = cpu_sample_scene(mpm_b200)
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_mpm")
    kind = "reference" if os.path.exists(ref) else "port"
    if kind == "reference":
        # every replica of the real class holds its I*J*K*N-float weight table (1.07 GB for the sample): stay inside host memory
        try:
            avail_kb = next(int(l.split()[1]) for l in open("/proc/meminfo") if l.startswith("MemAvailable"))
            cores = max(1, min(cores, int(0.5 * avail_kb * 1024 / (4.0 * sc["dims"][0] * sc["dims"][1] * sc["dims"][2] * sc["n"] + 2e8))))
        except Exception:
            pass
    per_step = []
    # one "step" = one substep of the bounded sample on every core; W warm-up + K timed, as one run of W+K substeps
    # per process is what the binary exposes, the warm-up run is a separate short launch
    if kind == "reference":
        run_reference_binary(sc, max(args.warmup, 1), cores)
        v, wall, sec = run_reference_binary(sc, args.steps, cores)
    else:
        run_oracle_port(sc, max(args.warmup, 1), cores)
        v, sec = run_oracle_port(sc, args.steps, cores)
    ms = sec * 1e3 / args.steps
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "cpu_sample": f"{cores} independent replicas of a {sc['n']}-particle 32^3 slab sample"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{sc['n']} particles x {args.steps} substeps per core, {cores} cores", "cpu_model": cpu_model()},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def roofline_block(ms_per_step, n_total, n_active, world, phase_ms, grid, particles):
    """The `roofline` object of the JSON line. phase_ms = MpmStats.last_ms of the last substep (max over ranks):
    [bin, clear, p2g, grid, g2p (F-update + gather), between begin/end, total, F-update alone or -1]."""
    peak, peak_src = measured_peak()
    alg_bytes = ALG_BYTES_PER_PARTICLE * n_total + ALG_BYTES_PER_NODE * n_active
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9 / world      # per-GPU GB/s against a per-GPU peak
    names = ["bin_sort", "grid_clear", "p2g", "halo_wait", "grid_update", "g2p(fupdate+gather)", "substep_total"]
    order = [0, 1, 2, 5, 3, 4, 6]
    kern = {names[i]: round(phase_ms[order[i]], 4) for i in range(7)}
    if len(phase_ms) > 7 and phase_ms[7] > 0:   # the two G2P kernels timed apart (default: back to back on one stream)
        kern["fupdate"] = round(phase_ms[7], 4)
        kern["g2p_gather"] = round(phase_ms[4] - phase_ms[7], 4)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{grid}:{particles}")
    except Exception:
        pass
    cands = ("p2g", "fupdate", "g2p_gather", "bin_sort", "grid_update") if "fupdate" in kern else ("p2g", "g2p(fupdate+gather)", "bin_sort", "grid_update")
    dom = max(cands, key=lambda k: kern[k])
    # per-kernel algorithmic bytes, per GPU
    npg, apg = n_total / world, n_active / world
    kern_alg = {"p2g": 88 * npg + 16 * apg, "g2p(fupdate+gather)": (128 + 112) * npg + (16 + 64) * npg + 16 * apg,
                "fupdate": (128 + 112) * npg, "g2p_gather": (16 + 64) * npg + 16 * apg,
                "bin_sort": 24 * npg, "grid_update": 32 * apg}
    dom_gbs = kern_alg[dom] / max(kern[dom] * 1e-3, 1e-12) / 1e9
    return {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
            "traffic": traffic, "peak_source": peak_src, "scope": "one whole substep (all kernels), per GPU",
            "algorithmic_bytes_per_substep": alg_bytes, "active_nodes": n_active,
            "kernel_ms_last_substep": kern, "dominant_kernel": dom,
            "dominant_kernel_algorithmic_bytes": kern_alg[dom], "dominant_kernel_achieved_gbs": round(dom_gbs, 1),
            "dominant_kernel_frac": round(dom_gbs / peak, 4),
            "dominant_share_of_substep": round(kern[dom] / max(kern["substep_total"], 1e-9), 3),
            # the cubic stencil makes the substep fp32-issue-bound before it is HBM-bound: the same run
            # expressed against the non-tensor fp32 peak, with the survey's model of ~4.5 kflop per particle-update
            "fp32_model_flop_per_particle": FP32_MODEL_FLOP_PER_PARTICLE,
            "achieved_fp32_tflops": round(FP32_MODEL_FLOP_PER_PARTICLE * n_total / world / (ms_per_step * 1e-3) / 1e12, 2),
            "fp32_peak_tflops_nominal": 74.4}


FP32_MODEL_FLOP_PER_PARTICLE = 4500     # parity-faithful substep, FMA = 2 flop
WORKLOAD_NAME = "snow_slab_512: 64Mi-particle snow slab avalanche, 512^3 grid (BASELINE config 5)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)           # >= 50 timed substeps after >= 10 warm-up
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=512)            # overrides are for development runs only
    ap.add_argument("--particles", type=int, default=1 << 26)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import mpm_b200
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libmpm_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        from datetime import timedelta
        # a rank that dies must not leave the others waiting for NCCL's default 10-minute watchdog (but slow first-time CUDA/NCCL start-up on a fresh box must fit)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=timedelta(seconds=240))
    workload = WORKLOAD_NAME if (args.grid == 512 and args.particles == 1 << 26) else f"snow_slab_{args.grid}: {args.particles} particles (development override)"

    from importlib import import_module
    multi = import_module("realtime-deformations_b200.multi")
    # development only: MPM_B200_VARIANTS="p2g:g2p" selects experimental kernel variants for A/B runs (default 0:0)
    variants = tuple(int(x) for x in os.environ.get("MPM_B200_VARIANTS", "0:0").split(":"))
    runner = multi.SlabRunner(args.grid, args.particles, rank, world, torch, variants=variants)
    stream = runner.stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for _ in range(args.warmup):
        runner.substep()
    barrier()

    # ---- timed region 1: K substeps, state resident in HBM (inputs = 22 GB of particle state >> 126 MB L2) ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = runner.sim.stats().kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        runner.substep()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    st = runner.sim.stats()
    launches = st.kernel_launches - launches0
    n_local, n_active_local = st.n_particles, st.n_active_nodes
    phase_ms = list(st.last_ms)

    # ---- timed region 2 (e2e): the viewer-frame contract through the C ABI with HOST buffers. Every step copies that
    # step's host inputs (collider transforms, main.cpp:187-190 moves them each frame) in, runs one substep, and copies
    # the render buffers the reference's drawParticles() consumes (xyz+size, main.cpp:257-271) out to pinned memory.
    n_up = runner.sim.n
    n_rows = max(int(n_up * 1.5) + (1 << 16), 1) if runner.migrates else max(n_up, 1)
    xyzs = torch.empty((n_rows, 4), dtype=torch.float32, pin_memory=True)
    xyzs_ptr = xyzs.numpy().ctypes.data
    e2e_steps = max(3, min(args.steps, 10))

    # rows copied out per step: all particles; in slab mode the live count drifts slowly, so one read-back before the
    # loop (+2 % head-room) fixes it without a per-step host sync
    n_dl = n_up if not runner.migrates else min(n_rows, int(runner.sim.capacity_rows() * 1.02) + 4096)

    def e2e_loop(pipelined):
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            runner.substep(host_colliders=True)
            if pipelined:
                runner.sim.wait_render_buffers()               # frame t-1 is complete (and consumed) before its buffer is reused
                runner.sim.render_buffers_async(xyzs_ptr, n_dl)
            else:
                runner.sim.L.mpm_download_render_buffers(runner.sim.h, n_dl, xyzs_ptr, None, 0.02)
        if pipelined:
            runner.sim.wait_render_buffers()
        barrier()
        return (time.perf_counter() - t0) * 1e3 / e2e_steps

    e2e_mode = "pipelined (copy of frame t overlaps the substep of frame t+1)"
    try:
        e2e_ms = e2e_loop(True)
    except Exception as exc:     # keep an end-to-end number even if the pipelined path is unavailable
        e2e_mode = f"synchronous (pipelined path failed: {exc})"
        e2e_ms = e2e_loop(False)

    # ---- reduce over ranks: max time, summed particles ----
    vals = torch.tensor([ms_total, e2e_ms, float(n_local), float(n_active_local), float(launches), float(n_dl)] + [float(x) for x in phase_ms],
                        dtype=torch.float64, device="cuda")
    if world > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, e2e_ms = mx[0].item(), mx[1].item()
        n_total, n_active, launches, n_dl_total = sm[2].item(), sm[3].item(), sm[4].item(), sm[5].item()
        phase_ms = mx[6:].tolist()
    else:
        n_total, n_active, n_dl_total = float(n_local), float(n_active_local), float(n_dl)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = ms_total / args.steps
    value = n_total / (ms_per_step * 1e-3)
    roof = roofline_block(ms_per_step, n_total, n_active, world, phase_ms, args.grid, args.particles)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline(mpm_b200)
        except Exception as exc:      # the GPU measurement above must not be lost to a failing CPU arm
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": f"cpu baseline failed: {exc}"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload, "particles": int(n_total), "grid": [args.grid] * 3, "dt": 1e-5, "stencil": "cubic (reference)",
                       "decomposition": f"{world} slab(s) along i" + ("" if world == 1 else (", ghost layer reduced by P2G over peer memory (experimental)"
                                                                                   if getattr(runner, "peer_halo", False) else ", halo over NCCL send/recv")), "l2_policy": "inputs (22 GB particle state) >> 126 MB L2, no flush needed",
                       "timing": "CUDA events on the library stream, max over ranks",
                       "kernel_variants": {"p2g": variants[0], "g2p": variants[1]},
                       "p2g_record_walk": "aligned (MPM_B200_P2G_ROTATE=0)" if os.environ.get("MPM_B200_P2G_ROTATE") == "0" else "rotated (default)"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": n_total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": runner.h2d_bytes_per_step,
                    "d2h_bytes_per_step": int(16 * n_dl_total), "ms_per_step": e2e_ms,
                    "what": "C-ABI substep with host collider structs in + render buffers (xyz,size) out to pinned host memory, every step",
                    "mode": e2e_mode},
            "roofline": roof, "cpu_baseline": cpu}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
