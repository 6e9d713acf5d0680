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


def cpu_baseline(mpm_b200, budget_steps=150, config=5):
    out = cpu_baseline_reference(mpm_b200, budget_steps)
    out["for_config"] = config
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
            "config": {"workload": CONFIGS[args.config]["name"], "baseline_config": args.config, "cpu_sample": f"{cores} independent replicas of a {sc['n']}-particle 32^3 slab sample"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{sc['n']} particles x {args.steps} substeps per core, {cores} cores", "cpu_model": cpu_model()},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def adapter_e2e(mpm_b200, config, steps):
    """End-to-end time of the ZERO-CHANGE drop-in: the reference's own headless driver loop (oracle/ref_driver.cpp = main.cpp:192-218,
    one call per stage) linked against adapter/lagrange_euler_view_b200.cpp instead of material_point_method.cpp, with host
    std::vector<Particle> mirrored after every substep -- all 35 floats (default) or only what the viewer reads
    (MPM_B200_ADAPTER_MIRROR=render) -- next to the unmodified reference binary on the same scene where its class can allocate it."""
    exe = os.path.join(ROOT, "oracle", "_ref", "adapter_mpm")
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_mpm")
    if not os.path.exists(exe):
        return None
    d = tempfile.mkdtemp()
    if config == 1:
        scene_args, n, what = [], 2147, "reference default scene (rand() fill, 2147 particles, 20^3)"
    else:
        sc = mpm_b200.scenes.snowball_drop(grid=128, n=1 << 20)
        np.concatenate([sc["pos"], sc["vel"], sc["mass"][:, None]], 1).astype(np.float32).tofile(d + "/p.f32")
        t = -sc["w2l"][0][12:15]
        np.array([t[0], t[1], t[2], 0.0, *sc["half"][0]], np.float32).tofile(d + "/c.f32")
        scene_args = ["--grid", "128", "128", "128", "--n", str(sc["n"]), "--load", d + "/p.f32", "--colliders", d + "/c.f32"]
        n, what = sc["n"], "snowball_drop_128 (1 Mi particles, 128^3)"

    def run(binary, env_extra):
        r = subprocess.run([binary] + scene_args + ["--steps", str(steps), "--quiet", "--bench"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                           text=True, env=dict(os.environ, **env_extra), timeout=600)
        for line in r.stdout.splitlines():
            if line.startswith("{"):
                return json.loads(line)["seconds"]
        raise RuntimeError(r.stdout[-300:])
    out = {"scene": what, "substeps": steps, "unit": UNIT, "what": "reference driver loop + adapter (staged C-ABI calls) + host Particle mirror per substep"}
    for mode in ("full", "render"):
        try:
            sec = run(exe, {"MPM_B200_ADAPTER_MIRROR": mode})
            out[f"mirror_{mode}"] = {"ms_per_substep": sec * 1e3 / steps, "value": n * steps / sec}
        except Exception as exc:
            out[f"mirror_{mode}"] = {"error": str(exc)}
    if config == 1 and os.path.exists(ref):
        try:
            sec = run(ref, {"CUDA_VISIBLE_DEVICES": ""})
            out["unmodified_reference"] = {"ms_per_substep": sec * 1e3 / steps, "value": n * steps / sec, "cores": 1}
        except Exception as exc:
            out["unmodified_reference"] = {"error": str(exc)}
    return out


def bind_to_gpu_numa_node(gpu_index):
    """Pin this rank's process to the CPUs NVML reports as local to its GPU, BEFORE any pinned host buffer is allocated: the
    per-frame render-buffer copies (e2e) then land in NUMA-local memory instead of all ranks sharing node 0's memory channels
    and the inter-socket link. Returns the CPU list, or None if NVML / affinity is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def csrc_sha16():
    """Identity of the kernel sources a profiler capture belongs to (profiles/traffic.json carries the same hash)."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "realtime-deformations_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def profiler_capture(key):
    """DRAM bytes / L2 reduction sectors per substep from the committed ncu capture of THIS kernel build (tools/traffic_from_ncu.py
    writes profiles/traffic.json with the hash of the kernel sources); None when the sources changed since the capture."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if d.get("csrc_sha16") == csrc_sha16():
            return d.get(key)
    except Exception:
        pass
    return None


def roofline_block(ms_per_step, n_total, n_active, world, phase_ms, cfg_key, fupd_in_p2g=True):
    """The `roofline` object of the JSON line. phase_ms = MpmStats.last_ms of the last substep (max over ranks):
    [bin, clear, p2g, grid, g2p (F-update + gather), between begin/end, total, F-update alone or -1]."""
    peak, peak_src = measured_peak()
    alg_bytes = ALG_BYTES_PER_PARTICLE * n_total + ALG_BYTES_PER_NODE * n_active
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9 / world      # per-GPU GB/s against a per-GPU peak
    timed_apart = len(phase_ms) > 7 and phase_ms[7] > 0       # the F-update as a kernel of its own, on the main stream
    names = ["bin_sort", "grid_clear", "p2g+fupdate" if fupd_in_p2g else "p2g", "halo_wait", "grid_update",
             "g2p_gather" if fupd_in_p2g else "g2p(fupdate+gather)", "substep_total"]
    order = [0, 1, 2, 5, 3, 4, 6]
    kern = {names[i]: round(phase_ms[order[i]], 4) for i in range(7)}
    if not fupd_in_p2g and timed_apart:     # the F-update as a kernel of its own (p2g_variant 2): timed apart from the gather
        kern["fupdate"] = round(phase_ms[7], 4)
        kern["g2p_gather"] = round(phase_ms[4] - phase_ms[7], 4)
    cap = profiler_capture(cfg_key) or {}
    cands = [k for k in ("p2g+fupdate", "p2g", "fupdate", "g2p_gather", "bin_sort", "grid_update") if k in kern]
    if not fupd_in_p2g and not timed_apart:
        cands = ["p2g", "g2p(fupdate+gather)", "bin_sort", "grid_update"]
    dom = max(cands, key=lambda k: kern[k])
    # per-kernel algorithmic bytes, per GPU
    npg, apg = n_total / world, n_active / world
    kern_alg = {"p2g": 88 * npg + 16 * apg, "p2g+fupdate": (88 + 80 + 112) * npg + 16 * apg, "g2p(fupdate+gather)": (128 + 112) * npg + (16 + 64) * npg + 16 * apg,
                "fupdate": (128 + 112) * npg, "g2p_gather": (16 + 64) * npg + 16 * apg,
                "bin_sort": 12 * npg, "grid_update": 32 * apg}
    dom_gbs = kern_alg[dom] / max(kern[dom] * 1e-3, 1e-12) / 1e9
    out = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
           "traffic": cap.get("dram_bytes_per_substep"), "peak_source": peak_src, "scope": "one whole substep (all kernels), per GPU",
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
    if cap:       # atomic throughput of P2G (north_star): L2 reduction sectors from the same ncu capture, per substep
        out["atomics"] = {"p2g_red_sectors_per_substep": cap.get("p2g_red_sectors"), "p2g_ms_in_capture": cap.get("p2g_ms"),
                          "p2g_red_sectors_per_s": cap.get("p2g_red_sectors_per_s"), "source": cap.get("source")}
    return out


FP32_MODEL_FLOP_PER_PARTICLE = 4500     # parity-faithful substep, FMA = 2 flop
CONFIGS = {
    1: dict(name="reference_default: the reference's own start-up scene (constants.hpp / main.cpp:48-52), 2147 particles, 20^3 grid (BASELINE config 1)", grid=20, n=2147, dt=1e-5),
    2: dict(name="snowball_drop_128: 1Mi-particle snowball dropped on a ground plane, 128^3 grid (BASELINE config 2)", grid=128, n=1 << 20, dt=1e-5, scene="snowball_drop"),
    3: dict(name="snowball_collision_256: two-snowball collision, 8Mi particles, 256^3 grid (BASELINE config 3)", grid=256, n=1 << 23, dt=1e-5, scene="snowball_collision"),
    4: dict(name="stiff_snowball_256: stiff-snow sweep point (xi=20, theta_c=1.5e-2, theta_s=7.5e-3), 4Mi particles, 256^3 grid, dt=2.5e-6 (BASELINE config 4)",
            grid=256, n=1 << 22, dt=2.5e-6, scene="stiff_snowball", params=dict(hardening_xi=20.0, theta_c=1.5e-2, theta_s=7.5e-3)),
    5: dict(name="snow_slab_512: 64Mi-particle snow slab avalanche, 512^3 grid (BASELINE config 5)", grid=512, n=1 << 26, dt=1e-5, scene="snow_slab"),
}
WORKLOAD_NAME = CONFIGS[5]["name"]


class DefaultSceneRunner:
    """BASELINE config 1: the reference's own start-up scene (the golden run's initial state, tests/golden/c1_default.npz, dumped
    from the unmodified reference) on one GPU. 13 launches per substep dominate at 2147 particles, so the fused substeps are
    replayed as a CUDA graph of substep pairs (MPM_B200_GRAPH, validated by test_graph_substeps_match_plain_path)."""
    migrates = False

    def __init__(self, torch, mpm_b200):
        os.environ["MPM_B200_GRAPH"] = "1"
        g = np.load(os.path.join(ROOT, "tests", "golden", "c1_default.npz"))
        s0 = g["state0"]
        self.dt = float(g["dt"])
        self.sim = mpm_b200.Sim(int(g["I"]), int(g["J"]), int(g["K"]), s0.shape[0], mpm_b200.capi.default_params(h=float(g["h"])))
        self.stream = torch.cuda.Stream()
        self.sim.set_stream(self.stream.cuda_stream)
        self.sim.upload_state35(s0)
        raw = np.asarray(g["colliders"], np.float32).reshape(-1, 29)
        self.cols, self.nc = mpm_b200.capi.make_colliders(raw[:, 13:29], raw[:, 0:3], raw[:, 10:13])
        self.h2d_bytes_per_step = 88 * self.nc + 4
        self.sim.rasterizeParticlesToGrid()
        self.pair = 2

    def substep(self, host_colliders=False, n=1):
        self.sim.substep(self.dt, self.cols, self.nc, n)


def build_runner(args, torch, mpm_b200, multi, rank, world, variants):
    cfg = CONFIGS[args.config]
    if args.config == 1:
        if world > 1:
            raise SystemExit("config 1 (2147 particles) is a single-GPU configuration")
        return DefaultSceneRunner(torch, mpm_b200), cfg["name"], None
    grid, n = (args.grid or cfg["grid"]), (args.particles or cfg["n"])
    name = cfg["name"] if (grid == cfg["grid"] and n == cfg["n"]) else f"{cfg['scene']}_{grid}: {n} particles (development override)"
    if cfg["scene"] == "snow_slab":       # every rank generates only its own slab (counter-based RNG keyed by cell id)
        return multi.SlabRunner(grid, n, rank, world, torch, dt=cfg["dt"], variants=variants), name, None
    full = getattr(mpm_b200.scenes, cfg["scene"])(grid=grid, n=n, dt=cfg["dt"])
    full, layers, ranges = multi.partition_scene(full, world)
    mine = multi.slice_scene(full, *ranges[rank])
    r = multi.SlabRunner(grid, mine["n"], rank, world, torch, scene=mine, dt=cfg["dt"], variants=variants,
                         layers=layers[rank] if world > 1 else None, params=cfg.get("params"))
    part = None if world == 1 else {"block_layers": [list(l) for l in layers], "particles": [b - a for a, b in ranges]}
    return r, name, part


def expected_id_sums(n):
    """sum id and sum splitmix64(id) over ids 0..n-1, mod 2^64 (what mpm_reduce_invariants must return for an intact particle set)."""
    tot, h = 0, 0
    step = 1 << 22
    with np.errstate(over="ignore"):
        for a in range(0, n, step):
            i = np.arange(a, min(n, a + step), dtype=np.uint64)
            z = i + np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            h = (h + int(z.sum(dtype=np.uint64))) & 0xFFFFFFFFFFFFFFFF
    tot = (n * (n - 1) // 2) & 0xFFFFFFFFFFFFFFFF
    return tot, h


def implicit_line(args):
    """SURVEY 8 row f4, opt-in (`--workload implicit`; the default contract line is the substep): one JSON line for the implicit
    (optimisation-based) time integration, mpm_time_integration = LagrangeEulerView::timeIntegration (material_point_method.cpp:
    211-233). A step is one solve (the vendored optimiser's up to 50 L-BFGS iterations) on a snowball mid-impact; the metric counts
    objective evaluations (one pass over all particles each, most with the gradient scatter) per second. Timed on the host around
    the C-ABI call, which ends with a device read-back (the library runs on its own stream, which torch events do not see)."""
    import mpm_b200
    import numpy as np
    grid, n = args.grid or 256, args.particles or (1 << 22)
    sc = mpm_b200.scenes.snowball_drop(grid=grid, n=n)
    sim = mpm_b200.Sim(grid, grid, grid, sc["n"], mpm_b200.capi.default_params())
    sim.upload(sc["pos"], sc["vel"], sc["mass"])
    sim.rasterizeParticlesToGrid(); sim.computeParticleVolumesAndDensities()
    cols, nc = mpm_b200.capi.make_colliders(sc["w2l"], sc["half"], sc["cvel"])
    sim.substep(float(sc["dt"]), cols, nc, 400)                  # explicit substeps until the ball is well into the impact
    dt, E, nu = 1e-4, 1.4e5, 0.2
    q = mpm_b200.capi.default_implicit_params(mu0=E / (2 * (1 + nu)), lambda0=E * nu / ((1 + nu) * (1 - 2 * nu)), hardening=1)
    steps, warm = max(1, min(args.steps, 10)), max(1, min(args.warmup, 3))
    launches0 = None
    evals, iters, times, e_drop = 0, 0, [], []
    sampler = ClockSampler(0)
    for k in range(warm + steps):
        sim.rasterizeParticlesToGrid(); sim.gridVelocitiesUpdate(dt); sim.synchronize()
        if k == warm:
            launches0 = sim.stats().kernel_launches
            sampler.start()
        t0 = time.perf_counter()
        st = sim.timeIntegration(dt, q)
        t1 = time.perf_counter() - t0
        if k >= warm:
            times.append(t1); evals += st.evaluations; iters += st.iterations; e_drop.append(st.energy_end / st.energy_start)
        sim.gridBasedCollisions(dt, cols, nc); sim.updateDeformationGradient(dt); sim.updateParticleVelocities(); sim.updateParticlePositions(dt)
    clocks = sampler.stop()
    stt = sim.stats()
    launches = stt.kernel_launches - launches0
    total = sum(times)
    n_p, n_active = sc["n"], stt.n_active_nodes
    value = n_p * evals / total
    peak, peak_src = measured_peak()
    # algorithmic bytes of one evaluation: read x, V0, FE, FP (88 B/particle); per active node read the trial velocity and (m, v*),
    # write the gradient (48 B)
    bytes_eval = 88.0 * n_p + 48.0 * n_active
    achieved = bytes_eval * evals / total / 1e9
    line = {"metric": "particle_evaluations_per_s", "value": value, "unit": "particle-evaluations/s", "n_gpus": 1, "steps": steps, "warmup": warm,
            "ms_per_step": total / steps * 1e3, "higher_is_better": True, "scaling": "replicas only", "vs_baseline": None, "dtype": "f32 (polar factor and sums in f64)",
            "data": "synthetic",
            "config": {"workload": f"implicit_snowball_{grid}: one implicit time-integration solve per step on a {n_p}-particle snowball mid-impact, {grid}^3 grid, dt = 1e-4 "
                                   "(SURVEY 8 row f4; not a BASELINE.json config: the reference never runs this path)",
                       "particles": n_p, "grid": [grid] * 3, "dt": dt, "iterations_per_solve": iters / steps, "evaluations_per_solve": evals / steps,
                       "energy_end_over_start": float(np.mean(e_drop)), "timing": "host clock around mpm_time_integration (ends with a device read-back)",
                       "l2_policy": "particle state (88 B/particle read per evaluation) >> 126 MB L2 at 4 Mi particles"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_evaluation": bytes_eval, "n_active_nodes": n_active,
                         "note": "whole solve, optimiser algebra and host waits included; the evaluation kernels alone are in profiles/ (ncu)"},
            "e2e": {"value": value, "unit": "particle-evaluations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(32 * (3 * iters + evals) / steps),
                    "note": "same call: the grid the solve starts from is produced on the device by the preceding stages; the read-backs are the optimiser's scalars"},
            "gpu_launches": int(launches), "clocks": clocks}
    if not args.no_cpu_baseline:
        # the oracle's evaluation of the same objective (value + analytic gradient) on one host core, bounded sample
        sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_py as op
        small = mpm_b200.scenes.small_ball(grid=48, radius_cells=12.0)
        o = op.Oracle(48, 48, 48, small["n"], op.default_params(h=float(small["h"])))
        o.set_state(op.initial_state(small["pos"], small["vel"], small["mass"])); o.rasterize(); o.volumes(); o.rasterize()
        used = o.used_cells(); v = o.grid()[used][:, 4:7]
        qo = op.default_implicit_params(mu0=q.mu0, lambda0=q.lambda0, hardening=1)
        t0 = time.perf_counter(); reps = 0
        while time.perf_counter() - t0 < 10.0:
            o.energy(v, dt, qo); o.energy_gradient(v, dt, qo); reps += 1
        tc = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": small["n"] * reps / tc, "unit": "particle-evaluations/s", "cores": 1, "kind": "port",
                                "sample": f"{reps} value + gradient evaluations of the oracle (oracle_energy, oracle_energy_gradient) on a {small['n']}-particle ball, 48^3 grid",
                                "cpu": cpu_model()}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)           # >= 50 timed substeps after >= 10 warm-up
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=5, choices=[1, 2, 3, 4, 5])   # BASELINE.json configs; 5 is the headline
    ap.add_argument("--grid", type=int, default=0)              # overrides are for development runs only
    ap.add_argument("--particles", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-multi-check", action="store_true")
    ap.add_argument("--workload", default="substep", choices=["substep", "implicit"])   # implicit: the f4 line (opt-in; one GPU)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return reference_arm(args)
    if args.workload == "implicit":
        return implicit_line(args)

    import torch
    import mpm_b200
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libmpm_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if os.environ.get("MPM_B200_NUMA_BIND", "1") != "0" else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        from datetime import timedelta
        # a rank that dies must not leave the others waiting for NCCL's default 10-minute watchdog (but slow first-time CUDA/NCCL start-up on a fresh box must fit)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=timedelta(seconds=240))

    from importlib import import_module
    multi = import_module("realtime-deformations_b200.multi")
    # development only: MPM_B200_VARIANTS="p2g:g2p" selects the A/B kernel pairings (default 0:0)
    variants = tuple(int(x) for x in os.environ.get("MPM_B200_VARIANTS", "0:0").split(":"))
    runner, workload, partition = build_runner(args, torch, mpm_b200, multi, rank, world, variants)
    stream = runner.stream
    cfg = CONFIGS[args.config]
    graph_pairs = args.config == 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    inv0 = runner.sim.invariants()
    # ---- warm-up ----
    if graph_pairs:
        runner.substep(n=2 * ((args.warmup + 1) // 2))
    else:
        for _ in range(args.warmup):
            runner.substep()
    barrier()

    # ---- timed region 1: K substeps, state resident in HBM (inputs = 22 GB of particle state >> 126 MB L2) ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = runner.sim.stats().kernel_launches
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    evs[0].record(stream)
    if graph_pairs:                 # one call: the library replays K/2 captured substep pairs (+ one plain substep if K is odd)
        runner.substep(n=args.steps)
        evs[-1].record(stream)
    else:
        for i in range(args.steps):
            runner.substep()
            evs[i + 1].record(stream)
    barrier()
    ms_total = evs[0].elapsed_time(evs[-1])
    per_step = [ms_total / args.steps] * args.steps if graph_pairs else [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop()
    st = runner.sim.stats()
    launches = st.kernel_launches - launches0
    n_local, n_active_local = st.n_particles, st.n_active_nodes
    phase_ms = list(st.last_ms)
    inv1 = runner.sim.invariants()

    # ---- timed region 2 (e2e): the viewer-frame contract through the C ABI with HOST buffers. Every step hands that
    # step's host inputs (collider structs, main.cpp:187-190 moves them each frame) in, runs one substep, and copies
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
                mpm_b200.capi._ck(runner.sim.L.mpm_download_render_buffers(runner.sim.h, n_dl, xyzs_ptr, None, 0.02))
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

    # ---- reduce over ranks: max time (per substep and in total), summed particles and conserved quantities ----
    def u2i(v):      # uint64 checksum -> the int64 with the same bits (torch has no uint64 reductions); sums wrap identically
        return v - (1 << 64) if v >= (1 << 63) else v
    vals = torch.tensor([ms_total, e2e_ms, float(n_local), float(n_active_local), float(launches), float(n_dl)] + [float(x) for x in phase_ms] + per_step,
                        dtype=torch.float64, device="cuda")
    fsum = torch.tensor([inv0["mass"]] + inv0["momentum"] + [inv1["mass"]] + inv1["momentum"], dtype=torch.float64, device="cuda")
    isum = torch.tensor([u2i(inv1["count"]), u2i(inv1["id_sum"]), u2i(inv1["id_hash"])], dtype=torch.int64, device="cuda")
    if world > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dist.all_reduce(fsum); dist.all_reduce(isum)
        ms_total, e2e_ms = mx[0].item(), mx[1].item()
        n_total, n_active, launches, n_dl_total = sm[2].item(), sm[3].item(), sm[4].item(), sm[5].item()
        phase_ms = mx[6:14].tolist()
        per_step = mx[14:].tolist()
    else:
        n_total, n_active, n_dl_total = float(n_local), float(n_active_local), float(n_dl)
    fsum = fsum.tolist()
    isum = [int(x) & 0xFFFFFFFFFFFFFFFF for x in isum.tolist()]

    # ---- multi-GPU correctness inside the same run: a small driven slab on all ranks against one domain on rank 0 ----
    mcheck = None
    if world > 1 and not args.no_multi_check:
        try:
            mcheck = multi.multi_vs_single_check(torch, rank, world)
        except Exception as exc:
            mcheck = {"ok": False, "error": str(exc)} if rank == 0 else None
        barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # the metric is the MEDIAN substep of >= 50; shorter runs report the mean (the driver's K / W)
    ms_mean = ms_total / args.steps
    ms_median = float(np.median(per_step))
    use_median = args.steps >= 50 and not graph_pairs
    ms_per_step = ms_median if use_median else ms_mean
    # (n_total summed the ranks' SLOT counts, which include the slots of just-migrated particles until the next re-sort:
    # the scene's particle count is what the metric and the conserved-quantity check are about)
    n_int = int(args.particles or cfg["n"])
    n_total = float(n_int)
    value = n_total / (ms_per_step * 1e-3)
    roof = roofline_block(ms_per_step, n_total, n_active, world, phase_ms, f"config{args.config}",
                          fupd_in_p2g=(variants == (0, 0) and not os.environ.get("MPM_B200_OVERLAP")))
    # conserved quantities after all substeps of this run (mass and ids exactly; momentum against p0 + M g dt k where nothing collides)
    exp_id_sum, exp_id_hash = expected_id_sums(n_int)
    mass_expected = n_int * float(np.float32(mpm_b200.scenes.PARTICLE_MASS))
    inv = {"particles": isum[0], "particles_expected": n_int, "id_sum": isum[1], "id_sum_expected": exp_id_sum,
           "id_hash": isum[2], "id_hash_expected": exp_id_hash, "mass": fsum[4], "mass_expected": mass_expected,
           "momentum_start": fsum[1:4], "momentum_after_timed_steps": fsum[5:8],
           "checked_after_substeps": args.warmup + args.steps}
    inv["ok"] = bool(isum[0] == n_int and isum[1] == exp_id_sum and isum[2] == exp_id_hash and abs(fsum[4] - mass_expected) <= 1e-9 * mass_expected)
    if cfg.get("scene") == "snowball_collision":      # no collider: momentum changes by gravity alone (cpp:256-262)
        k = args.warmup + args.steps
        g = [0.0, -9.8, 0.0]
        exp_p = [fsum[1 + a] + fsum[0] * g[a] * cfg["dt"] * k for a in range(3)]
        scale = max(abs(fsum[0] * 100.0), 1e-30)          # |m v| of one ball (v0 = 100 m/s): the two balls' momenta cancel
        inv["momentum_expected"] = exp_p
        inv["momentum_rel_err"] = max(abs(fsum[5 + a] - exp_p[a]) for a in range(3)) / scale
        inv["ok"] = bool(inv["ok"] and inv["momentum_rel_err"] <= 1e-4)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline(mpm_b200, config=args.config)
        except Exception as exc:      # the GPU measurement above must not be lost to a failing CPU arm
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": f"cpu baseline failed: {exc}"}
    e2e_adapter = None
    if world == 1 and args.config in (1, 2) and not args.no_cpu_baseline:
        try:
            runner.sim.close()           # the adapter's own process creates its handle on the same GPU
            e2e_adapter = adapter_e2e(mpm_b200, args.config, 400 if args.config == 1 else 40)
        except Exception as exc:
            e2e_adapter = {"error": str(exc)}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "ms_per_step_mean": ms_mean, "ms_per_step_median": ms_median,
            "ms_per_step_min_max": [float(min(per_step)), float(max(per_step))],
            "value_statistic": "median of the timed substeps (SURVEY 8d)" if use_median else "mean of the timed substeps",
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload, "baseline_config": args.config, "particles": n_int, "grid": [args.grid or cfg["grid"]] * 3, "dt": cfg["dt"], "stencil": "cubic (reference)",
                       "decomposition": f"{world} slab(s) along i" + ("" if world == 1 else (", ghost layers reduced by P2G over peer memory (NVLink), migration pulled through peer mappings"
                                                                                   if getattr(runner, "peer_halo", False) else ", halo over NCCL send/recv")),
                       "partition": partition,
                       "l2_policy": "inputs (particle state, 176 B/particle per buffer) >> 126 MB L2 at configs 2-5, no flush needed; config 1 fits in L2 (launch-bound)",
                       "timing": "CUDA events on the library stream, one per substep, max over ranks",
                       "kernel_variants": {"p2g": variants[0], "g2p": variants[1]},
                       "fupdate": "tolerance form inside the P2G kernel (fupdate_exact=0)" if variants[0] == 0 else "tolerance form, own kernel"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": n_total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": runner.h2d_bytes_per_step,
                    "d2h_bytes_per_step": int(16 * n_dl_total), "ms_per_step": e2e_ms,
                    "what": "C-ABI substep with host collider structs in + render buffers (xyz,size) out to pinned host memory, every step",
                    "mode": e2e_mode, "host_numa_binding": (f"{len(numa)} CPUs local to the GPU (NVML)" if numa else "none")},
            "e2e_adapter": e2e_adapter,
            "roofline": roof, "invariants": inv, "multi_gpu_check": mcheck, "cpu_baseline": cpu}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if not inv["ok"] or (mcheck is not None and not mcheck.get("ok", False)):
        sys.stderr.write("bench.py: conserved quantities or the multi-GPU cross-check do not hold, see `invariants` / `multi_gpu_check`\n")
        sys.exit(3)


if __name__ == "__main__":
    main()
