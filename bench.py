#!/usr/bin/env python
"""Headline benchmark of the step->photon path (BASELINE.json metric: photons propagated/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one bunch of synthetic I3CLSimSteps: BASELINE
config 2 (muon-track steps as the ppc parameterisation emits them, SpiceMie layered ice with
171 layers, IC86-like 5160-DOM detector, DOM oversize 5, stop on detection), 2^20 steps x 200
photons = 2.1e8 photons per bunch and GPU.

* value    : photons/s, whole job, bunch resident in HBM, timed with CUDA events on the
             engine's compute stream (one kernel launch per step); max over ranks.
* e2e      : photons/s through the reference-facing API (EnqueueSteps / GetConversionResult on
             the C ABI) from HOST buffers: pinned staging + H2D of every bunch and D2H of its
             hits are inside the timed region.
* roofline : compute (FP32 issue) roofline of SURVEY.md 8(d): algorithmic lane-ops
             (155 per segment + 120 per photon, the reference formulation's count) per second
             against SMs x 128 lanes x clock.  HBM figures for the hit stream are attached to
             show that memory is not the bound.
* cpu_baseline : the reference's own kernel text compiled for the host (oracle/_ref, kind
             "reference"; POCL/OpenCL does not exist in this image) -- or, where that library was not
             built, the oracle restatement (kind "port") -- on all host cores, on a bounded sample of
             the same workload.  --impl reference times exactly that as its own arm.

Multi-GPU (--gpus N under torchrun): weak scaling, every rank propagates its own bunch on
its own GPU with its own slice of the MWC multiplier table; there is no collective on the
data path (steps shard by bunch), NCCL is used only for the barrier and the max-over-ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

STEPS_PER_BUNCH = 1 << 20
PHOTONS_PER_STEP = 200
WORKLOAD = ("config2: muon-track steps (ppc parameterisation shape), SpiceMie 171 layers tilt off, "
            "IC86-like 5160 DOMs, oversize 5, stop on detection, %d steps x %d photons per bunch" % (STEPS_PER_BUNCH, PHOTONS_PER_STEP))
OPS_PER_SEGMENT = 155.0  # SURVEY.md 8(d), reference formulation
OPS_PER_PHOTON = 120.0


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "sm_max_mhz": d.get("sm_max_mhz", 1965.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, cell in zip(names, r[5:9]):
                if cell.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_traffic(bunch_steps):
    """DRAM bytes (read + write) of one launch of the fast kernel at this bunch size, from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json, written by tools/ncu_traffic.py); None if there is
    no capture of this launch shape."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.isfile(path):
        return None
    with open(path) as f:
        d = json.load(f)
    if int(d.get("bunch_steps", -1)) != int(bunch_steps):
        return None
    return d.get("dram_bytes_total")


def build_scene(ice_model="spice_mie", tilt=False):
    from clsim_b200 import geometry, ice
    medium = ice.MakeIceCubeMediumProperties(iceDataDirectory=ice_model, useTiltIfAvailable=tilt)
    geo = geometry.make_ic86_like_geometry(oversize=5.0)
    bias = ice.GetIceCubeDOMAcceptance(domRadius=geometry.DOM_RADIUS * 5.0)
    gen = ice.makeCherenkovWavelengthGenerator(bias, False, medium)
    return medium, geo, [gen], bias


def make_bunch(n, seed):
    from clsim_b200 import steps
    return steps.muon_track_steps(n, photons_per_step=PHOTONS_PER_STEP, seed=seed)


class CpuArm(object):
    """The reference's implementation of the path on the host cores.  kind "reference": the reference's own kernel
    text (resources/kernels/*.cl) compiled for the host under oracle/ref_shim (oracle/_ref/libclsim_ref.so, built
    where /root/reference exists and shipped prebuilt); kind "port": the oracle restatement, when that library is
    absent.  The scene and the MWC multipliers are built ONCE; each run() times one launch over `sample_steps`
    work-items (one per step, like an OpenCL CPU device) on all host threads."""

    def __init__(self, scene, max_steps, threads):
        from clsim_b200.description import ConverterOptions
        from oracle import pyoracle
        self.threads = threads
        medium, geo, gens, bias = scene
        opt = ConverterOptions(stop_detected_photons=True, pancake_factor=5.0)
        if pyoracle.ref_available():
            self.kind = "reference"
            self.scene = pyoracle.RefScene(medium, geo, gens, bias, opt)
            self.flags = "g++ -O3 -ffp-contract=off -fopenmp (no -march=native: the .so is built off-box), glibc libm; kernel text of the reference under a C++ shim for OpenCL C, variant %s" % self.scene.variant()
        else:
            self.kind = "port"
            self.scene = pyoracle.Scene(medium, geo, gens, bias, opt)
            self.flags = "g++ -O3 -ffp-contract=off -fopenmp (no -march=native), glibc libm; oracle restatement"
        self.a, _, _ = pyoracle.safeprimes(0, max_steps)
        self.max_steps = max_steps
        self._seed_states = pyoracle.seed_states

    def run(self, sample_steps, seed):
        """-> (photons/s, seconds, photons, hits)"""
        n = min(sample_steps, self.max_steps)
        bunch = make_bunch(n, seed)
        x = self._seed_states(seed, self.a[:n])
        photons = int(bunch["num_photons"].sum())
        t0 = time.perf_counter()
        out = self.scene.propagate(bunch, x, self.a[:n], cap=max(1000, 10 * n), num_threads=self.threads)
        dt = time.perf_counter() - t0
        return photons / dt, dt, photons, int(out[1])

    def describe(self, value, sample):
        return {"value": value, "unit": "photons/s", "cores": self.threads, "kind": self.kind, "sample": sample, "flags": self.flags}


def run_variants(args, scene, bunch, opt, local, rank, barrier, max_over_ranks, sum_over_ranks):
    """The same end-to-end loop with (a) photons thinned to MCPEs on the device (only photo-electrons come back) and
    (b) additionally the bunch generated on the device from step-generation queue entries (nothing but a few hundred
    bytes goes up).  Informational: the headline `e2e` stays the plain EnqueueSteps / GetConversionResult path."""
    from clsim_b200 import capi, mcpe, stepgen, steps
    from clsim_b200.sharding import mcpe_row_offset, stepgen_row_offset
    medium, geo, gens, bias = scene
    ang = mcpe.GetIceCubeDOMAngularSensitivity()
    # acceptance = the generation bias itself (UnshadowedFraction and hole-ice peak folded in would only scale both)
    acc_of = {(int(s), int(o)): bias for s, o in zip(geo.stringIDs, geo.domIDs)}
    out = {}
    photons_per_bunch = float(bunch["num_photons"].sum())
    for name in ("host_steps_mcpe_out", "device_steps_mcpe_out"):
        conv = mcpe.I3CLSimPhotonToMCPEConverterForDOMs(900 + rank, acc_of, ang, device=local, rngFirstMultiplierRow=mcpe_row_offset(rank))
        gen = stepgen.StepGenerator(device=local, rng_seed=950 + rank, rng_first_multiplier=stepgen_row_offset(rank))
        src = steps.muon_track_sources(len(bunch), photons_per_step=PHOTONS_PER_STEP)
        eng = capi.Engine(medium, geo, gens, bias, opt)
        conv.attach_to(eng)
        send = (lambda i: eng.enqueue(bunch, i)) if name == "host_steps_mcpe_out" else (lambda i: gen.enqueue_into(eng, src, i))
        for i in range(args.warmup):
            send(10 + i)
        for i in range(args.warmup):
            eng.get_result()
        barrier()
        t0 = time.perf_counter()
        pes, hits, pending = 0, 0, 0
        for i in range(args.steps):
            send(100 + i)
            pending += 1
            while eng.more_photons_available():
                r = eng.get_result()
                pes += len(r.mcpes); hits += r.num_hits_counted; pending -= 1
        while pending:
            r = eng.get_result()
            pes += len(r.mcpes); hits += r.num_hits_counted; pending -= 1
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        out[name] = {"value": sum_over_ranks(photons_per_bunch * args.steps) / dt, "unit": "photons/s",
                     "h2d_bytes_per_step": int(len(bunch) * 48) if name == "host_steps_mcpe_out" else int(src.nbytes + 24),
                     "d2h_bytes_per_step": int(pes / max(1, args.steps) * 16 + 40),
                     "mcpe_per_hit": pes / float(max(1, hits))}
        eng.close()
        conv.close()
        gen.close()
    return out


def run_reference_arm(args):
    """The reference's own implementation of the path on the host cores (CpuArm: oracle/_ref when it was built, else
    the oracle port; OpenCL itself cannot run in this image)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    t_start = time.perf_counter()
    max_steps = 1 << 16
    arm = CpuArm(build_scene(), max_steps, threads)
    rate, _, _, _ = arm.run(1 << 12, 100)   # calibration (and first-touch of the thread pool)
    # every timed step about 2.5 s of CPU work, bounded; warm-up steps an eighth of that
    per_step = int(min(max_steps, max(1 << 11, rate * 2.5 / PHOTONS_PER_STEP)))
    for i in range(args.warmup):
        arm.run(max(1 << 10, per_step // 8), 200 + i)
    t_total, photons = 0.0, 0
    for i in range(args.steps):
        _, dt, n_photons, _ = arm.run(per_step, 300 + i)
        t_total += dt
        photons += n_photons
    value = photons / t_total
    sample = "%d steps x %d photons per timed step (bounded sample of the bunch), %d timed steps" % (per_step, PHOTONS_PER_STEP, args.steps)
    line = {
        "impl": "reference", "metric": "photons propagated/sec", "value": value, "unit": "photons/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample, "wall_s": time.perf_counter() - t_start},
        "cpu_baseline": arm.describe(value, sample),
        "e2e": {"value": value, "unit": "photons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bunch", type=int, default=STEPS_PER_BUNCH, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-variants", action="store_true", help=argparse.SUPPRESS)
    # side measurements (not the headline): BASELINE config 3's ice, "spice_lea" = SpiceLea with tilt and anisotropy
    ap.add_argument("--ice", default="spice_mie", help=argparse.SUPPRESS)
    ap.add_argument("--tilt", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from clsim_b200 import capi
    from clsim_b200.description import KERNEL_FAST, ConverterOptions
    from clsim_b200.sharding import rng_row_offset

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    scene = build_scene(args.ice, args.tilt)
    medium, geo, gens, bias = scene
    n = args.bunch
    bunch = make_bunch(n, seed=1000 + rank)  # every rank its own bunch: weak scaling
    opt = ConverterOptions(device=local, stop_detected_photons=True, pancake_factor=5.0, kernel_mode=KERNEL_FAST,
                           enable_double_buffering=True, max_num_workitems=n, rng_seed=4242 + rank,
                           rng_first_multiplier=rng_row_offset(rank))
    eng = capi.Engine(medium, geo, gens, bias, opt)
    sampler = ClockSampler(local)

    # ---------------- value: bunch resident in HBM, kernel-only timing ----------------------
    eng.upload_resident(bunch)
    eng.run_resident(args.warmup)
    barrier()
    sampler.start()
    res = eng.run_resident(args.steps)  # CUDA events on the compute stream around each launch; L2 flushed between launches
    barrier()
    kernel_ms = max_over_ranks(res["kernel_ms"])
    photons_all = sum_over_ranks(float(res["photons"]))
    segments_all = sum_over_ranks(float(res["segments"]))
    hits_all = sum_over_ranks(float(res["hits"]))
    value = photons_all / (kernel_ms * 1e-3)

    # ---------------- e2e: host buffers through the C ABI -----------------------------------
    for i in range(args.warmup):  # warm the pipeline (first-touch of staging and result buffers)
        eng.enqueue(bunch, 10 + i)
    for i in range(args.warmup):
        eng.get_result()
    barrier()
    stats0 = eng.statistics()
    t0 = time.perf_counter()
    e2e_hits = 0
    pending = 0
    timeline = []  # (what, ms since t0): where the host side spends the end-to-end time
    for i in range(args.steps):
        eng.enqueue(bunch, 100 + i)  # blocks when 5 bunches are queued, like the reference
        timeline.append(("enq", round(1e3 * (time.perf_counter() - t0), 1)))
        pending += 1
        while eng.more_photons_available():
            e2e_hits += len(eng.get_result().photons)
            timeline.append(("res", round(1e3 * (time.perf_counter() - t0), 1)))
            pending -= 1
    while pending:
        e2e_hits += len(eng.get_result().photons)
        timeline.append(("res", round(1e3 * (time.perf_counter() - t0), 1)))
        pending -= 1
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_kernel_ms = (eng.statistics()["TotalDeviceTime"] - stats0["TotalDeviceTime"]) * 1e-6
    barrier()
    clocks = sampler.stop()
    e2e_value = sum_over_ranks(float(bunch["num_photons"].sum()) * args.steps) / e2e_s
    stats = eng.statistics()
    eng.close()

    # ---------------- e2e with the neighbours on the device (rows f2/f3): informational -------------
    variants = None
    if not args.no_variants:
        variants = run_variants(args, scene, bunch, opt, local, rank, barrier, max_over_ranks, sum_over_ranks)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    props = torch.cuda.get_device_properties(local)
    sms = props.multi_processor_count
    peak_ops = sms * 128 * peaks["sm_max_mhz"] * 1e6 * world
    achieved_ops = (segments_all * OPS_PER_SEGMENT + photons_all * OPS_PER_PHOTON) / (kernel_ms * 1e-3)
    hit_bytes = hits_all * 80.0
    roofline = {
        "bound": "compute-fp32-issue", "achieved": achieved_ops / 1e12, "peak": peak_ops / 1e12, "unit": "Tlane-op/s",
        "frac": achieved_ops / peak_ops, "traffic": measured_traffic(n),
        "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum), algorithmic: %d" % int(
            n * 48 + hits_all / max(1, args.steps * world) * 80),
        "peak_source": "%d SMs x 128 lanes x %.0f MHz (sm_max_mhz %s)" % (sms, peaks["sm_max_mhz"], peaks["source"]),
        "ops_per_segment": OPS_PER_SEGMENT, "ops_per_photon": OPS_PER_PHOTON,
        "segments_per_photon": segments_all / photons_all,
        "segments_per_s": segments_all / (kernel_ms * 1e-3),
        "frac_at_sampled_clock": (achieved_ops / (sms * 128 * clocks["sm_mhz"] * 1e6 * world)) if clocks.get("sm_mhz") else None,
        "hbm": {"achieved_gbs": (hit_bytes + photons_all / PHOTONS_PER_STEP * 48.0) / (kernel_ms * 1e-3) / 1e9,
                "peak_gbs": peaks["hbm_gbs"] * world, "note": "hit records out + step records in; negligible by design"},
    }
    line = {
        "metric": "photons propagated/sec", "value": value, "unit": "photons/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": kernel_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD if (args.ice == "spice_mie" and not args.tilt) else WORKLOAD.replace("SpiceMie 171 layers tilt off", "%s tilt %s (side measurement)" % (args.ice, "on" if args.tilt else "off")),
                   "l2": "256 MiB memset between timed launches (L2 flush)", "kernel": "fast persistent",
                   "ns_per_photon": 1e9 / value, "hit_fraction": hits_all / photons_all,
                   "parallelism": "steps sharded by bunch, %d independent GPU(s), no collective" % world},
        "e2e": {"value": e2e_value, "unit": "photons/s", "h2d_bytes_per_step": int(n * 48),
                "d2h_bytes_per_step": int(e2e_hits / max(1, args.steps) * 80 + 8),
                "api": "clsimcu_enqueue/clsimcu_get_result (EnqueueSteps/GetConversionResult), double buffering on",
                "kernel_ms_per_step": e2e_kernel_ms / max(1, args.steps), "wall_ms": e2e_s * 1e3, "timeline_ms": timeline},
        "gpu_launches": args.steps,
        "e2e_variants": variants,
        "roofline": roofline,
        "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        arm = CpuArm(scene, 1 << 17, threads)
        rate0, _, _, _ = arm.run(1 << 12, 77)
        sample = int(min(1 << 17, max(1 << 12, rate0 * 12.0 / PHOTONS_PER_STEP)))   # about 12 s of CPU work
        rate, dt, _, _ = arm.run(sample, 78)
        line["cpu_baseline"] = arm.describe(rate, "%d steps x %d photons of the same workload, %.1f s" % (sample, PHOTONS_PER_STEP, dt))
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
