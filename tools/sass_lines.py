"""Static SASS instruction count per SOURCE LINE of one kernel of the built library (no GPU needed):
extracts the cubin, runs `nvdisasm -g`, attributes every instruction to the innermost `//## File ..., line N` marker
before it and prints the lines of the tile-kernel source with their counts. Used to see where a kernel's
instructions go (derive / sort / accumulation / fold) before spending GPU time.
  python tools/sass_lines.py <mangled-name-substring> [source-file-substring] [min-count]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("MPM_B200_LIB", os.path.join(ROOT, "realtime-deformations_b200", "libmpm_b200.so"))


def line_histogram(kernel_substr, lib=LIB):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, stdout=subprocess.DEVNULL, check=True)
        cubin = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
        sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    hist = collections.Counter()
    ops = collections.defaultdict(collections.Counter)
    name, cur, inside = None, None, False
    for ln in sass.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            inside = kernel_substr in m.group(1)
            name = m.group(1) if inside else name
            cur = None
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", ln)
        if m and cur:
            op = re.sub(r"^@!?U?P\d+\s+", "", m.group(1)).split()[0].split(".")[0]
            hist[cur] += 1
            ops[cur][op] += 1
    return name, hist, ops


if __name__ == "__main__":
    kname = sys.argv[1]
    fsub = sys.argv[2] if len(sys.argv) > 2 else ""
    minc = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    name, hist, ops = line_histogram(kname)
    print("kernel:", name, " total attributed:", sum(hist.values()))
    by_file = collections.Counter()
    for (f, l), c in hist.items():
        by_file[f] += c
    print("by file:", dict(by_file))
    for (f, l), c in sorted(hist.items()):
        if fsub in f and c >= minc:
            top = ", ".join(f"{o}:{n}" for o, n in ops[(f, l)].most_common(4))
            print(f"{f}:{l:4d} {c:5d}   {top}")
