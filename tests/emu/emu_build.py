"""TEST INFRASTRUCTURE ONLY. Builds tests/emu/_build/libmpm_b200_emu.so: the WHOLE library (C ABI + host orchestration +
kernels) compiled by g++ for the host, from copies of the product sources in which every kernel launch
    kernel<<<grid, block, smem, stream>>>(args);
is rewritten to the emulator's synchronous launch
    emu::launch(grid, block, smem, [&] { kernel(args); });
Nothing else in the sources is touched (they carry `#ifdef MPM_HOST_EMU` switches for inline PTX only). The emulated
library exports `mpm_emulated_build`; capi refuses to load it unless MPM_B200_ALLOW_EMULATION=1, which only
tests/test_emulated_library_cpu.py sets. It exists to run the GPU parity tests' logic on small scenes before GPU time is
spent; it is orders of magnitude slower than the real thing and is never a fallback."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "realtime-deformations_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
SRC_OUT = os.path.join(BUILD, "src")
LIB = os.path.join(BUILD, "libmpm_b200_emu.so")
FILES = ("mpm_api.cu", "mpm_tile_kernels.cuh", "mpm_implicit.cuh", "mpm_kernels.cuh", "mpm_math.cuh")


def _kernel_start(text, i):
    """index where the kernel expression ending at text[:i] starts (identifier with optional template arguments)"""
    j = i
    while j > 0 and text[j - 1].isspace():
        j -= 1
    if text[j - 1] == ">":
        angle, paren = 0, 0
        while True:
            j -= 1
            c = text[j]
            if c == ")":
                paren += 1
            elif c == "(":
                paren -= 1
            elif paren == 0 and c == ">":
                angle += 1
            elif paren == 0 and c == "<":
                angle -= 1
                if angle == 0:
                    break
    while j > 0 and (text[j - 1].isalnum() or text[j - 1] in "_:"):
        j -= 1
    return j


def _split_top(s):
    out, depth, cur = [], 0, ""
    for c in s:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += c
    out.append(cur.strip())
    return out


def rewrite_launches(text):
    out, pos, n = "", 0, 0
    while True:
        i = text.find("<<<", pos)
        if i < 0:
            break
        k0 = _kernel_start(text, i)
        j = text.index(">>>", i)
        cfg = _split_top(text[i + 3:j])
        assert len(cfg) == 4, f"launch configuration with {len(cfg)} fields: {text[i:j + 3]}"
        a = j + 3
        while text[a].isspace():
            a += 1
        assert text[a] == "(", text[i:a + 20]
        depth, b = 0, a
        while True:
            if text[b] == "(":
                depth += 1
            elif text[b] == ")":
                depth -= 1
                if depth == 0:
                    break
            b += 1
        kernel, args = text[k0:i].strip(), text[a + 1:b]
        name = kernel.split("<")[0]
        out += text[pos:k0] + f"emu::launch({cfg[0]}, {cfg[1]}, {cfg[2]}, [&] {{ {kernel}({args}); }}, \"{name}\")"
        pos = b + 1
        n += 1
    return out + text[pos:], n


def build(force=False):
    os.makedirs(SRC_OUT, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in FILES] + [os.path.join(HERE, f) for f in ("cuda_emu.h", "fake_cudart.cpp", "emu_build.py")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    total = 0
    for f in FILES:
        text, n = rewrite_launches(open(os.path.join(CSRC, f)).read())
        text = text.replace('#include "../../include/mpm_b200.h"', '#include "mpm_b200.h"')      # found through -I include/
        text = text.replace("cudaFuncSetAttribute(", "emu_func_set_attribute(")                   # g++ has no template wrapper for it
        total += n
        open(os.path.join(SRC_OUT, f.replace(".cu", ".cpp") if f.endswith(".cu") else f), "w").write(text)
    assert total >= 40, f"only {total} launches rewritten"
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-ffp-contract=off", "-pthread", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-DMPM_HOST_EMU_API=1",   # -Bsymbolic: the fake runtime wins even if torch has loaded libcudart
           "-I/usr/local/cuda/include", "-I" + SRC_OUT, "-I" + os.path.join(ROOT, "include"), "-include", os.path.join(HERE, "cuda_emu.h"),
           os.path.join(SRC_OUT, "mpm_api.cpp"), os.path.join(HERE, "fake_cudart.cpp"), "-o", LIB]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("host build of the library failed:\n" + r.stdout[-6000:])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
