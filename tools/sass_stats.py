"""Per-kernel SASS statistics of the built library (no GPU needed): instruction count, packed / scalar fp32 FMA counts,
shared-memory loads, registers and spills. Used to compare kernel variants before spending GPU time.
  python tools/sass_stats.py [filter-substring]"""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("MPM_B200_LIB", os.path.join(ROOT, "realtime-deformations_b200", "libmpm_b200.so"))


def kernels(lib=LIB):
    sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", lib], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    regs = {}
    for m in re.finditer(r"Function (\S+):\n\s*REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", res):
        regs[m.group(1)] = (int(m.group(2)), int(m.group(4)))
    out = []
    for part in re.split(r"\n\s*Function : ", sass)[1:]:
        name = part.split("\n", 1)[0].strip()
        ins = re.findall(r"/\*[0-9a-f]{4}\*/\s+(.*?);", part)
        ops = [re.sub(r"^@!?U?P\d+\s+", "", i).split()[0] if i.split() else "" for i in ins]
        body = "\n".join(ins)
        out.append({"name": name, "n": len(ins), "ffma2": sum(o.startswith("FFMA2") for o in ops),
                    "ffma": sum(o == "FFMA" or o.startswith("FFMA.") for o in ops),
                    "fmul_fadd": sum(o.split(".")[0] in ("FMUL", "FADD", "FMUL2", "FADD2") for o in ops),
                    "lds": sum(o.startswith("LDS") for o in ops), "sts": sum(o.startswith("STS") for o in ops),
                    "shfl": sum(o.startswith("SHFL") for o in ops), "bar": sum(o.startswith("BAR") for o in ops),
                    "regs": regs.get(name, (None, None))[0], "local": regs.get(name, (None, None))[1],
                    "md5": hashlib.md5(body.encode()).hexdigest()[:10]})
    return out


if __name__ == "__main__":
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    print(f"{'instr':>6} {'FFMA2':>6} {'FFMA':>6} {'FMUL/ADD':>8} {'LDS':>5} {'STS':>5} {'SHFL':>5} {'BAR':>4} {'regs':>5} {'local':>6}  md5         kernel")
    for k in kernels():
        if flt in k["name"]:
            short = subprocess.run(["c++filt", k["name"]], stdout=subprocess.PIPE, text=True).stdout.strip().split("(")[0]
            print(f"{k['n']:6d} {k['ffma2']:6d} {k['ffma']:6d} {k['fmul_fadd']:8d} {k['lds']:5d} {k['sts']:5d} {k['shfl']:5d} {k['bar']:4d} {str(k['regs']):>5} {str(k['local']):>6}  {k['md5']}  {short}")
