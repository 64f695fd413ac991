"""Per-source-line / per-region instruction shares of an .ncu-rep (needs --import-source on, -lineinfo).
usage: python tools/ncu_regions.py rep lane_segments [top_n]"""
import csv, io, subprocess, sys, re
rep = sys.argv[1]
lane_segments = float(sys.argv[2])
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hidx = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
h = rows[hidx[0]]
ci = {n: i for i, n in enumerate(h)}
sec = rows[hidx[0] + 1: (hidx[1] - 2 if len(hidx) > 1 else len(rows))]
src_rows = [r for r in sec if len(r) >= len(h) and r[0] != '']
def f(r, k):
    try: return int(r[ci[k]] or 0)
    except Exception: return 0
tot_i = sum(f(r, "Instructions Executed") for r in src_rows); tot_s = sum(f(r, "# Samples") for r in src_rows)
print("warp-instructions %d  samples %d  warp-instr per 32 lane-segments %.1f" % (tot_i, tot_s, 32 * tot_i / lane_segments))
# regions from marker comments in the source: lines containing '// ----' start a region
src = open("clsim_b200/csrc/kernel_fast.cu").read().split("\n")
marks = [(i + 1, l.strip()) for i, l in enumerate(src) if re.search(r"// -{4,}|^__device__|^template|^__global__|^struct ", l)]
marks.append((len(src) + 1, "end"))
for (a, name), (b, _) in zip(marks, marks[1:]):
    rs = [r for r in src_rows if a <= int(r[0]) < b]
    ie = sum(f(r, "Instructions Executed") for r in rs); te = sum(f(r, "Thread Instructions Executed") for r in rs); s = sum(f(r, "# Samples") for r in rs)
    if ie * 200 > tot_i:
        print("%4d-%4d inst %5.1f%% smp %5.1f%% eff %4.1f  per32seg %6.1f  %s" % (a, b - 1, 100 * ie / tot_i, 100 * s / max(1, tot_s), te / max(ie, 1), 32 * ie / lane_segments, name[:70]))
print("top lines:")
for r in sorted(src_rows, key=lambda r: -f(r, "Instructions Executed"))[:topn]:
    ie = f(r, "Instructions Executed"); te = f(r, "Thread Instructions Executed")
    print("%4s inst %5.2f%% smp %5.2f%% eff %4.1f per32seg %5.1f  %s" % (r[0], 100 * ie / tot_i, 100 * f(r, "# Samples") / max(1, tot_s), te / max(1, ie), 32 * ie / lane_segments, r[1].strip()[:90]))
