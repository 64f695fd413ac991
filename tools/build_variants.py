"""Build libclsimcuda variants that differ only in compile-time knobs of kernel_fast.cu, for
A/B runs in ONE gpurun call:  python tools/build_variants.py "256 4" "256 3" ...
Each variant lands in clsim_b200/variants/libclsimcuda_T<threads>_B<blocks>[_tag].so; select it
with CLSIMCU_LIB=<path> (clsim_b200/capi.py)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def main():
    g.build_product()
    out_dir = os.path.join(ROOT, "clsim_b200", "variants")
    os.makedirs(out_dir, exist_ok=True)
    for spec in sys.argv[1:]:
        parts = spec.split()
        threads, blocks = parts[0], parts[1]
        extra = parts[2:]
        tag = "T%s_B%s" % (threads, blocks) + "".join("_" + e.replace("-D", "").replace("=", "") for e in extra)
        obj = os.path.join(g.BUILD, "kernel_fast_%s.o" % tag)
        cmd = [g._nvcc(), "-ccbin", g._host_cxx()] + g.ARCH + g.NVCC_COMMON + ["-DCLSIMCU_THREADS=" + threads, "-DCLSIMCU_BLOCKS_PER_SM=" + blocks] + extra + \
            ["-c", os.path.join(g.CSRC, "kernel_fast.cu"), "-o", obj]
        log = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if log.returncode != 0:
            sys.stderr.write(log.stdout)
            raise SystemExit(1)
        for line in log.stdout.splitlines():
            if "ILb0ELb0ELb0" in line:
                idx = log.stdout.splitlines().index(line)
                print(tag, " | ".join(l.strip() for l in log.stdout.splitlines()[idx + 1: idx + 3]))
        objs = [obj] + [os.path.join(g.BUILD, o) for o in ("kernel_reference.o", "engine.o", "mcpe.o", "stepgen.o", "tabulate.o", "tables.o")]
        lib = os.path.join(out_dir, "libclsimcuda_%s.so" % tag)
        subprocess.check_call([g._nvcc(), "-ccbin", g._host_cxx()] + g.ARCH + ["-shared", "-o", lib] + objs + ["-lpthread", "-ldl"])
        print("built", lib)


if __name__ == "__main__":
    main()
