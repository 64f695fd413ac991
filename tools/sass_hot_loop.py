"""Static look at the hot loop of the fast kernel: dump the SASS of one instantiation of propagate_persistent from an
object file and print, per basic region between the loop head (first VOTE of the kernel body) and the loop's back
branch, the opcode histogram.  Usage: python tools/sass_hot_loop.py <kernel_fast.o> [mangled-substring]"""
import collections
import re
import subprocess
import sys


def main():
    obj = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else "propagate_persistentILb0ELb0ELb0ELb1E"
    out = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout.splitlines()
    lines, on = [], False
    for l in out:
        if "Function :" in l:
            on = want in l
            continue
        if on:
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
            if m:
                lines.append((int(m.group(1), 16), m.group(2).strip()))
    print("instructions in function:", len(lines))
    # the hot loop: the largest backward branch whose body contains MUFU.SIN and no CALL to the slow phase before it
    best = None
    for i, (addr, ins) in enumerate(lines):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", ins)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr:
                body = [x for x in lines if tgt <= x[0] <= addr]
                if all(any(k in b[1] for b in body) for k in ("MUFU.SIN", "MUFU.EX2", "MUFU.LG2", "VOTE")) and (best is None or len(body) < len(best)):
                    best = body
    if best is None:
        print("no loop found")
        return
    print("hot loop: 0x%x .. 0x%x, %d instructions" % (best[0][0], best[-1][0], len(best)))
    hist = collections.Counter()
    for _, ins in best:
        op = ins.split()[0]
        if op.startswith("@"):
            op = ins.split()[1]
        hist[op.split(".")[0] if not op.startswith(("FFMA2", "FMUL2", "FADD2")) else op] += 1
    for k, v in hist.most_common():
        print("  %-12s %d" % (k, v))
    if "-v" in sys.argv:
        for a, ins in best:
            print("/*%04x*/ %s" % (a, ins))


if __name__ == "__main__":
    main()
