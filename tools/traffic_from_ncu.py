"""profiles/traffic.json from an ncu CSV of one run of tools/profile_step.py (per-kernel dram__bytes_read/write.sum,
lts__t_sectors_op_red.sum, gpu__time_duration.sum): sums over ONE substep (the launches between two k_scan_reduce launches) and
records the hash of the kernel sources, so that bench.py quotes the capture only for the build it was taken from.
  python tools/traffic_from_ncu.py gpurun_out/ncu_substep.csv config5 [profiles/traffic.json]"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

src, key = sys.argv[1], sys.argv[2]
out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "traffic.json")
rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
hdr = rows[0]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
launches = {}
for r in rows[1:]:
    launches.setdefault(int(r[ii]), {"name": r[ki]})[r[mi]] = float(r[vi].replace(",", ""))
ids = sorted(launches)
starts = [i for i in ids if launches[i]["name"].startswith("k_scan_reduce")]
assert len(starts) >= 2, "the capture window must hold one whole substep (two k_scan_reduce launches)"
cycle = [launches[i] for i in ids if starts[0] <= i < starts[1]]
dram = sum(l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0) for l in cycle)
p2g = [l for l in cycle if "k_p2g_tile" in l["name"]]
red = sum(l.get("lts__t_sectors_op_red.sum", 0) for l in p2g)
p2g_ms = sum(l.get("gpu__time_duration.sum", 0) for l in p2g) / 1e6
per_kernel = [{"kernel": l["name"].split("(")[0][:60], "ms": round(l.get("gpu__time_duration.sum", 0) / 1e6, 4),
               "dram_gb": round((l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0)) / 1e9, 3),
               "red_sectors": int(l.get("lts__t_sectors_op_red.sum", 0)), "atom_sectors": int(l.get("lts__t_sectors_op_atom.sum", 0))} for l in cycle]
d = {}
if os.path.exists(out_path):
    try:
        d = json.load(open(out_path))
    except Exception:
        d = {}
sha = bench.csrc_sha16()
if d.get("csrc_sha16") != sha:
    d = {"csrc_sha16": sha}
d[key] = {"dram_bytes_per_substep": dram, "p2g_red_sectors": red, "p2g_ms": round(p2g_ms, 4),
          "p2g_red_sectors_per_s": red / max(p2g_ms * 1e-3, 1e-12), "source": os.path.basename(src) + " (ncu, cold-cache serialised launches)",
          "kernels": per_kernel}
json.dump(d, open(out_path, "w"), indent=1)
print(json.dumps({k: v for k, v in d[key].items() if k != "kernels"}))
for k in per_kernel:
    print(k)
