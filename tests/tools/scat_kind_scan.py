import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from clsim_b200 import steps
from clsim_b200.description import KERNEL_FAST, KERNEL_REFERENCE
from tests.scenes import make_scene, dom_near
from tests.test_gpu_fast_kernel import _run_resident, _compare_distributions
kind = int(sys.argv[1])
sc = make_scene("homogeneous"); sc.medium.scat_kind = kind
src = dom_near(sc.geo, (0.0, 0.0, 0.0)) + np.array([10.0, 5.0, 3.0])
allp = {}; sf = sr = 0
for k in range(8):
    bunch = steps.point_source_steps(1 << 18, 200, pos=tuple(src), seed=300 + k)
    fast, tf = _run_resident(sc, bunch, KERNEL_FAST, seed=31 + 1000 * k)
    ref, tr = _run_resident(sc, bunch, KERNEL_REFERENCE, seed=32 + 1000 * k)
    p = _compare_distributions(fast, ref, tf, tr)
    sf += len(fast); sr += len(ref)
    print(k, len(fast), len(ref), {a: float("%.3g" % b) for a, b in p.items()}, flush=True)
    for a, b in p.items(): allp.setdefault(a, []).append(b)
print("min p:", {a: float("%.3g" % min(b)) for a, b in allp.items()})
print("hits fast %d ref %d ratio %.5f z %.2f" % (sf, sr, sf / sr, (sf - sr) / np.sqrt(sf + sr)))
