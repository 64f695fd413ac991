"""Repeat the fast-kernel vs reference-order-kernel comparison with different seeds and print the p-values:
a check that a borderline p-value in one run of a statistical test is a fluctuation and not a bias.
usage (on the GPU box): python tests/tools/pvalue_scan.py [scene] [n_steps] [repeats]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from clsim_b200 import steps
from clsim_b200.description import KERNEL_FAST, KERNEL_REFERENCE
from tests.scenes import make_scene, dom_near
from tests.test_gpu_fast_kernel import _run_resident, _compare_distributions

scene = sys.argv[1] if len(sys.argv) > 1 else "homogeneous"
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 17
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
sc = make_scene(scene)
src = dom_near(sc.geo, (0.0, 0.0, 0.0)) + np.array([10.0, 5.0, 3.0])
allp = {}
sum_fast = sum_ref = 0
for r in range(reps):
    if scene == "homogeneous":
        bunch = steps.point_source_steps(n_steps, 200, pos=tuple(src), seed=100 + r)
    else:
        bunch = steps.muon_track_steps(n_steps, seed=100 + r)
    fast, tf = _run_resident(sc, bunch, KERNEL_FAST, seed=1000 + r, repeat=2)
    ref, tr = _run_resident(sc, bunch, KERNEL_REFERENCE, seed=2000 + r, repeat=2)
    p = _compare_distributions(fast, ref, tf, tr)
    sum_fast += len(fast); sum_ref += len(ref)
    print(r, len(fast), len(ref), {k: float("%.3g" % v) for k, v in p.items()}, flush=True)
    for k, v in p.items():
        allp.setdefault(k, []).append(v)
print("min p per statistic:", {k: float("%.3g" % min(v)) for k, v in allp.items()})
print("hits fast %d ref %d: ratio %.5f, z = %.2f" % (sum_fast, sum_ref, sum_fast / float(sum_ref), (sum_fast - sum_ref) / np.sqrt(sum_fast + sum_ref)))
