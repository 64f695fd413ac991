#!/usr/bin/env python
"""Build-time syntax translation of the reference's OpenCL C kernel text to something g++ parses.

TEST INFRASTRUCTURE (oracle/_ref recipe).  Reads the reference's kernel files where they lie
(/root/reference/resources/kernels/*.cl), writes the translated text to the output directory
(oracle/_ref/, git-ignored, removed again after the compile).  Nothing of the reference is stored in the
repository.

Exactly ONE rewrite is applied, because C++ has no spelling for it: the OpenCL C vector literal

    (float4)(a, b, c, d)      (const floating4_t)(a, b, c, d)      (float2)(a, b)

(in C++ a C-style cast of the comma expression `a, b, c, d`, i.e. of `d`) becomes the constructor call

    float4(a, b, c, d)        floating4_t(a, b, c, d)              float2(a, b)

Everything else -- address-space qualifiers, built-in functions, swizzles, atomics, work-item functions -- is
supplied by oracle/ref_shim/opencl_c_shim.inc without touching the text.  The script prints every line it
changed, so the build log shows the whole difference between what the reference ships and what is compiled.
"""
import os
import re
import sys

VECTOR_LITERAL = re.compile(r"\(\s*(?:const\s+)?(floating4_t|floating2_t|float4|float2|double4|double2)\s*\)(\s*)\(")

FILES = ["mwcrng_kernel.cl", "propagation_kernel.h.cl", "propagation_kernel.c.cl", "sparse_collision_kernel.h.cl",
         "sparse_collision_kernel.c.cl"]


def translate(text):
    return VECTOR_LITERAL.sub(lambda m: m.group(1) + m.group(2) + "(", text)


def main():
    src_dir, out_dir = sys.argv[1], sys.argv[2]
    os.makedirs(out_dir, exist_ok=True)
    changed = 0
    for name in FILES:
        with open(os.path.join(src_dir, name)) as f:
            text = f.read()
        new = translate(text)
        for lineno, (a, b) in enumerate(zip(text.split("\n"), new.split("\n")), 1):
            if a != b:
                changed += 1
                sys.stdout.write("%s:%d\n  - %s\n  + %s\n" % (name, lineno, a.strip(), b.strip()))
        with open(os.path.join(out_dir, name + ".inc"), "w") as f:
            f.write(new)
    sys.stdout.write("translate.py: %d lines rewritten in %d files\n" % (changed, len(FILES)))


if __name__ == "__main__":
    main()
