"""TEST INFRASTRUCTURE -- builds tests/emu/_build/libslr_splat_emu.so: the CUDA sources of
slr-sfs_b200/csrc compiled for the CPU against tests/emu/include/cuda_runtime.h, with the
identical C ABI (include/slr_splat.h) taking HOST pointers.

The only source transformation is the launch syntax, which is not C++:
    kernel<<<grid, block, smem, stream>>>(args)  ->  emu::launch(dim3(grid), dim3(block), [&]() { (kernel)(args); })
Everything else (kernels, device helpers, host-side C ABI) is compiled as it stands.
Used by tests/test_emu_*.py only; nothing in the package knows about it.
"""
import glob
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "slr-sfs_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libslr_splat_emu.so")
_LAUNCH = re.compile(r"([A-Za-z_][\w:]*(?:<[^<>;(){}]*>)?)\s*<<<")


def _match_paren(text, i):
    """Index just past the parenthesis group opening at text[i] == '('."""
    depth = 0
    for j in range(i, len(text)):
        if text[j] == "(":
            depth += 1
        elif text[j] == ")":
            depth -= 1
            if depth == 0:
                return j + 1
    raise ValueError("unbalanced parentheses")


def _split_top(text):
    parts, depth, cur = [], 0, ""
    for ch in text:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return [p.strip() for p in parts]


def convert(text):
    out, pos = "", 0
    while True:
        m = _LAUNCH.search(text, pos)
        if not m:
            return out + text[pos:]
        end_cfg = text.index(">>>", m.end())
        cfg = _split_top(text[m.end():end_cfg])
        assert len(cfg) >= 2, cfg
        a0 = end_cfg + 3
        while text[a0].isspace():
            a0 += 1
        assert text[a0] == "(", text[a0:a0 + 40]
        a1 = _match_paren(text, a0)
        out += text[pos:m.start()]
        out += "emu::launch(dim3(%s), dim3(%s), [&]() { (%s)%s; })" % (cfg[0], cfg[1], m.group(1), text[a0:a1])
        pos = a1


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = glob.glob(os.path.join(CSRC, "*.cu*")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(HERE, "*.cpp")) + glob.glob(os.path.join(HERE, "include", "*.h")) + \
        [os.path.join(ROOT, "include", "slr_splat.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, asan=False):
    """asan=True: a second library instrumented with AddressSanitizer (load it in a process started
    with LD_PRELOAD=libasan): the CPU stand-in for compute-sanitizer's memcheck."""
    if asan:
        build(force)
        lib = LIB[:-3] + "_asan.so"
        if force or not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(LIB):
            _compile(lib, ["-fsanitize=address", "-fno-omit-frame-pointer"])
        return lib
    if not force and not stale():
        return LIB
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(HERE, "emu_runtime.cpp")]
    for cu in sorted(glob.glob(os.path.join(CSRC, "*.cu"))):
        dst = os.path.join(OUT, os.path.basename(cu)[:-3] + ".emu.cpp")
        with open(cu) as fh:
            text = convert(fh.read())
        assert "<<<" not in text
        with open(dst, "w") as fh:
            fh.write('#line 1 "%s"\n' % cu + text)
        srcs.append(dst)
    _compile(LIB, [])
    return LIB


def build_variant(tag, defines):
    """The same sources with extra -D defines (compile-time variants that are prepared but not yet
    enabled in the product build), as _build/libslr_splat_emu_<tag>.so."""
    build()
    lib = LIB[:-3] + "_" + tag + ".so"
    if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(LIB):
        _compile(lib, list(defines))
    return lib


def _compile(lib, extra):
    srcs = [os.path.join(HERE, "emu_runtime.cpp")] + sorted(glob.glob(os.path.join(OUT, "*.emu.cpp")))
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-fno-strict-aliasing",
           "-Wno-attributes", "-Wno-unknown-pragmas", "-I", os.path.join(HERE, "include"), "-I", CSRC] + extra + \
          os.environ.get("SLR_DEFINES", "").split() + \
          ["-o", lib] + srcs
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("emulation build failed:\n" + proc.stdout[-6000:])


def libasan():
    """Path of the ASan runtime to LD_PRELOAD, or None."""
    try:
        path = subprocess.run(["gcc", "-print-file-name=libasan.so"], stdout=subprocess.PIPE, text=True).stdout.strip()
    except OSError:
        return None
    return path if os.path.isabs(path) and os.path.exists(path) else None


if __name__ == "__main__":
    print(build(force=True))
