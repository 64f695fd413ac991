"""Host check of the CUDA sources of the reference-order path (TEST INFRASTRUCTURE).

build() compiles engine.cu, kernel_reference.cu, mcpe.cu, stepgen.cu, tabulate.cu and tables.cpp of clsim_b200/csrc/ with g++
for the HOST, unmodified except for one syntax rewrite C++ has no spelling for -- the three kernel launches
`kernel<<<grid, block, shmem, stream>>>(args)` become `HOSTCHECK_LAUNCH(kernel, grid, block, args)` -- under
tests/hostcheck/cuda_runtime.h, a stand-in for the CUDA runtime (memory is memory, streams are program order, a kernel is a
loop over its threads).  The result, tests/hostcheck/_build/libclsimcuda_hostcheck.so, has the C ABI of include/clsimcuda.h and
runs the reference-order kernel, the photon -> MCPE converter, the step generator and the reference-order table maker on the CPU.

What it is for: the SOURCE of those units can be held against the oracle bit for bit on a machine without a GPU (same libm),
including the engine's threads, staging and result assembly.  What it is not: a way to run the product.  Nothing under
clsim_b200/ loads it, clsim_b200.capi never falls back to it, and the fast kernel is not in it (its launcher returns an
error): the product path has no CPU fallback."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "clsim_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libclsimcuda_hostcheck.so")
UNITS = ("engine.cu", "kernel_reference.cu", "mcpe.cu", "stepgen.cu", "tabulate.cu")
_LAUNCH = re.compile(r"(\w+)<<<(.*?)>>>\(", re.S)


def _split_top(text):
    parts, depth, cur = [], 0, ""
    text = text.replace("->", "\x00")
    for ch in text:
        if ch in "(<[":
            depth += 1
        elif ch in ")>]":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return [p.replace("\x00", "->") for p in parts]


def rewrite_launches(text):
    """-> (text with every <<<>>> launch spelled as HOSTCHECK_LAUNCH, the rewritten lines)."""
    changed = []

    def rep(m):
        cfg = _split_top(m.group(2))
        new = "HOSTCHECK_LAUNCH(%s, %s, %s, " % (m.group(1), cfg[0], cfg[1])
        changed.append((m.group(0), new))
        return new
    return _LAUNCH.sub(rep, text), changed


def build(force=False, sanitize=False):
    """sanitize=True: a second library, built with -fsanitize=address,undefined (tests/test_hostcheck.py runs the byte-identity
    check and the engine tests under it)."""
    global LIB
    if sanitize == "thread":
        return _build(os.path.join(BUILD, "libclsimcuda_hostcheck_tsan.so"), force, ["-fsanitize=thread", "-fno-omit-frame-pointer", "-g", "-O1"])
    lib = os.path.join(BUILD, "libclsimcuda_hostcheck_asan.so") if sanitize else LIB
    return _build(lib, force, ["-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-g", "-O1"] if sanitize else ["-O2"])


def _build(LIB, force, opt_flags):
    sources = [os.path.join(CSRC, u) for u in UNITS] + [os.path.join(CSRC, "tables.cpp"), os.path.join(HERE, "cuda_runtime.h"),
                                                        os.path.join(HERE, "fast_kernel_stub.cpp"), os.path.abspath(__file__)]
    sources += [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith(".h")]
    if not force and os.path.isfile(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in sources):
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    log, cpps = [], []
    for u in UNITS:
        with open(os.path.join(CSRC, u)) as f:
            text, changed = rewrite_launches(f.read())
        log += ["%s: %s  ->  %s" % (u, a, b) for a, b in changed]
        cpp = os.path.join(BUILD, ("san_" if "-g" in opt_flags else "") + u.replace(".cu", ".cpp"))
        with open(cpp, "w") as f:
            f.write(text)
        cpps.append(cpp)
    with open(os.path.join(BUILD, "rewrite.log"), "w") as f:
        f.write("\n".join(log) + "\n%d launches rewritten\n" % len(log))
    # -ffp-contract=off: kernel_reference.cu is built with --fmad=false for the device as well (the exact twin)
    cmd = ["g++"] + opt_flags + ["-std=c++17", "-fPIC", "-w", "-ffp-contract=off", "-fno-fast-math", "-I" + HERE, "-I" + CSRC, "-I" + os.path.join(ROOT, "include"),
           "-shared", "-o", LIB] + cpps + [os.path.join(CSRC, "tables.cpp"), os.path.join(HERE, "fast_kernel_stub.cpp"), "-lpthread", "-ldl"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    for cpp in cpps:
        os.remove(cpp)
    if res.returncode != 0:
        raise RuntimeError("host check build failed:\n" + res.stdout[-6000:])
    return LIB
