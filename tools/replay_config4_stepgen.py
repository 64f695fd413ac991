#!/usr/bin/env python
"""Which rank of bench.py's config-4 leg (device-made cascade steps) is handed a step at infinity?

The step generator is deterministic: the host draws the photon count of the cascade (numpy Generator seeded per rank),
the kernel's thread `me` makes steps me, me + threads, ... of every bunch from its own MWC stream (csrc/stepgen.cu).
This script replays exactly that on the CPU for the seeds bench.py uses (SetRandomService(40 + rank), multiplier rows
stepgen_row_offset(rank)), and reports the steps whose position along the shower axis is not finite
(gammaDistributedNumber at ry == 1, I3CLSimLightSourceToStepConverterUtils.h:100-108) and the draws with a zero low
word (the only way to get there).  Diagnostic for the hang recorded in profiles/bench_r02_v38_n8_config4_hang.err.

    python tools/replay_config4_stepgen.py [--ranks 8 | --rank R] [--launches K] [--sms 148]
"""
import argparse
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranks", type=int, default=8)
    ap.add_argument("--rank", type=int, default=None, help="this rank only")
    ap.add_argument("--launches", type=int, default=None, help="replay only the first K launches of a rank")
    ap.add_argument("--sms", type=int, default=148)
    ap.add_argument("--bunch", type=int, default=1 << 20)
    ap.add_argument("--photons", type=float, default=1.25e10)
    args = ap.parse_args()
    from clsim_b200 import capi, geometry, ice, mcpe, stepgen
    from clsim_b200.sharding import stepgen_row_offset
    from oracle import pyoracle

    so = os.path.join(tempfile.mkdtemp(), "replay.so")
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tools", "replay_config4_stepgen.c"), "-lm"])
    lib = C.CDLL(so)
    lib.replay_launch.restype = C.c_long
    lib.replay_launch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_double, C.c_double, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                  C.POINTER(C.c_longlong), C.POINTER(C.c_uint32)]

    lea = ice.MakeIceCubeMediumProperties(iceDataDirectory="spice_lea", useTiltIfAvailable=True)
    acc = ice.GetIceCubeDOMAcceptance(domRadius=geometry.DOM_RADIUS * 5.0, efficiency=0.9 * mcpe.GetHoleIcePeak())
    threads = 2 * args.sms * 256

    class NoDevice(object):   # stands in for the CUDA object: only the host side of the converter is run here
        def __init__(self, **kw):
            pass

    stepgen.StepGenerator = NoDevice
    for rank in (range(args.ranks) if args.rank is None else [args.rank]):
        conv = stepgen.I3CLSimLightSourceToStepConverterPPC(photonsPerStep=200, device=0)
        conv.SetMediumProperties(lea)
        conv.SetWlenBias(acc)
        conv.SetRandomService(40 + rank)
        conv.Initialize(rngFirstMultiplierRow=stepgen_row_offset(rank))
        a = capi.safeprime_multipliers(stepgen_row_offset(rank), threads)
        x = pyoracle.seed_states(40 + rank, a)
        a = np.ascontiguousarray(a, dtype=np.uint32)
        launches = []

        class Engine(object):
            def max_num_workitems(self):
                return args.bunch

        def enqueue_into(engine, sources, identifier, launches=launches):
            assert len(sources) == 1 and int(sources[0]["kind"]) == stepgen.CASCADE
            s = sources[0]
            launches.append((int(s["num_steps"]) + (1 if int(s["photons_in_last_step"]) > 0 else 0), float(s["pa"]), float(s["pb"]),
                             int(s["num_steps"]) * int(s["photons_per_step"]) + int(s["photons_in_last_step"])))

        conv.generator.enqueue_into = enqueue_into
        vertex, axis = (20.0, -30.0, -250.0), (0.3, 0.2, -0.93)
        conv.EnqueueLightSource(stepgen.Particle("EMinus", 1e3, vertex, axis), 0)
        while conv.EnqueueInto(Engine(), 0):
            pass
        small = sum(l[3] for l in launches)
        energy = 1e3 * args.photons / max(1.0, small)
        conv.EnqueueLightSource(stepgen.Particle("EMinus", energy, vertex, axis), 1)
        conv.EnqueueBarrier()
        while conv.EnqueueInto(Engine(), 100):
            pass
        draws, zeros = C.c_uint64(0), C.c_uint64(0)
        report = []
        for i, (total, pa, pb, _) in enumerate(launches[:args.launches]):
            step, thread = C.c_longlong(-1), C.c_uint32(0)
            bad = lib.replay_launch(x.ctypes.data, a.ctypes.data, threads, total, pa, pb, C.byref(draws), C.byref(zeros), C.byref(step), C.byref(thread))
            if bad:
                report.append("launch %d: %d step(s) at infinity, the first is step %d (thread %d)" % (i, bad, step.value, thread.value))
        print("rank %d: %d launches, %.4g steps, %.4g photons, a %.4g GeV e-, %.4g draws, %d with a zero low word; %s"
              % (rank, len(launches), sum(l[0] for l in launches), sum(l[3] for l in launches), energy, draws.value, zeros.value,
                 "; ".join(report) if report else "no step at infinity"))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
