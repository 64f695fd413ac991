"""Group the SASS instructions of an .ncu-rep by execution count (loop nest levels), to read the
cost per hot-loop iteration.  usage: python tools/ncu_sass_groups.py rep [min_share]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = rows[1]; ci = {n: i for i, n in enumerate(h)}
body = [r for r in rows[2:] if len(r) >= len(h)]
g = collections.OrderedDict()
for r in body:
    c = int(r[ci["Instructions Executed"]] or 0)
    g.setdefault(c, []).append(r)
tot = sum(c * len(v) for c, v in g.items())
# the hot loop count = the count with the largest c*len among groups with >= 40 instructions
N = max((c for c, v in g.items() if len(v) >= 40), key=lambda c: c * len(g[c]))
print("total warp-instr %d; hot-loop iterations N=%d; instr per iteration %.1f" % (tot, N, tot / N))
for c, v in sorted(g.items(), key=lambda kv: -kv[0] * len(kv[1]))[:22]:
    te = sum(int(r[ci["Thread Instructions Executed"]] or 0) for r in v)
    ops = collections.Counter((r[1].split()[1] if r[1].strip().startswith('@') else r[1].split()[0]) for r in v)
    print("x%.3f n=%4d -> %6.1f/iter eff %4.1f  %s" % (c / N, len(v), c * len(v) / N, te / max(1, c * len(v)), dict(ops.most_common(7))))
if len(sys.argv) > 2:
    lo = float(sys.argv[2])
    for r in body:
        c = int(r[ci["Instructions Executed"]] or 0)
        if c >= lo * N:
            te = int(r[ci["Thread Instructions Executed"]] or 0)
            print("%6s x%.3f eff %4.1f smp %6s  %s" % (r[0][-5:], c / N, te / max(c, 1), r[ci["# Samples"]], r[1].strip()))
