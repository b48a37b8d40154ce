"""TEST INFRASTRUCTURE: build oracle/_ref/libftc_emu.so -- csrc/train_ops.cu and csrc/loss_ops.cu compiled for HOST threads
through the CUDA-on-CPU shim (cuda_emu.h), so tests/test_emu_kernels.py can execute the real kernel source (indexing, shared
memory reductions, warp shuffles, atomics) in a container without a GPU.  The only source transformation is syntactic:
``kernel<<<grid, block, smem, stream>>>(args);`` becomes ``emu::launch(grid, block, smem, [&]{ kernel(args); });`` and
``extern __shared__ T name[];`` becomes a pointer to the emulated dynamic shared memory.

    python oracle/emu/build_emu.py        ->  oracle/_ref/libftc_emu.so (git-ignored)
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "findtextcenternet_b200", "csrc")
OUT = os.path.join(ROOT, "oracle", "_ref")
SOURCES = ["train_ops.cu", "loss_ops.cu", "page_ops.cu", "data_ops.cu"]


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{<" and not (ch == "<" and False):
            depth += ch in "([{"
        if ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def transform(src: str) -> str:
    out, pos = [], 0
    for m in re.finditer(r"<<<", src):
        start = m.start()
        if start < pos:
            continue
        # kernel name (with template arguments) = the token run before <<<
        i = start
        depth = 0
        while i > 0:
            c = src[i - 1]
            if c == ">":
                depth += 1
            elif c == "<":
                depth -= 1
            elif depth == 0 and not (c.isalnum() or c in "_:"):
                break
            i -= 1
        name = src[i:start]
        cfg_end = src.index(">>>", start)
        cfg = _split_top(src[start + 3:cfg_end])
        assert src[cfg_end + 3] == "(", src[cfg_end:cfg_end + 40]
        j, depth = cfg_end + 3, 0
        while True:
            depth += src[j] == "("
            depth -= src[j] == ")"
            j += 1
            if depth == 0:
                break
        args = src[cfg_end + 4:j - 1]
        assert src[j] == ";", (name, src[j:j + 20])
        grid, block = cfg[0], cfg[1]
        smem = cfg[2] if len(cfg) > 2 else "0"
        out.append(src[pos:i])
        out.append(f"emu::launch(dim3({grid}), dim3({block}), (size_t)({smem}), [&] {{ {name}({args}); }});")
        pos = j + 1
    out.append(src[pos:])
    res = "".join(out)
    res = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)\[\];", r"\1* \2 = reinterpret_cast<\1*>(emu::dyn_smem);", res)
    return res


def build() -> str:
    os.makedirs(OUT, exist_ok=True)
    gen = os.path.join(OUT, "emu_src")
    os.makedirs(gen, exist_ok=True)
    cpps = [os.path.join(HERE, "cuda_emu.cpp")]
    for name in SOURCES:
        text = transform(open(os.path.join(CSRC, name)).read())
        dst = os.path.join(gen, name[:-3] + "_emu.cpp")
        with open(dst, "w") as f:
            f.write(text)
        cpps.append(dst)
    lib = os.path.join(OUT, "libftc_emu.so")
    cmd = ["g++", "-std=c++20", "-O1", "-g", "-pthread", "-shared", "-fPIC", "-DFTC_EMU", "-fpermissive", "-w",
           "-I", HERE, "-I", CSRC, "-I", os.path.join(ROOT, "include"), *cpps, "-o", lib]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emu build failed:\n" + r.stderr[-6000:])
    return lib


if __name__ == "__main__":
    print(build())
