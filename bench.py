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


def roofline_frac(photons, segments, kernel_ms, sms, sm_mhz, gpus=1):
    return ((segments * OPS_PER_SEGMENT + photons * OPS_PER_PHOTON) / (kernel_ms * 1e-3)) / (sms * 128 * sm_mhz * 1e6 * gpus)


def run_other_configs(args, local, sms, sm_mhz):
    """The single-GPU BASELINE configurations besides the headline one (resident bunch, kernel-only timing, the same
    launch path and L2 flush as `value`): C1 point source in homogeneous ice at its stated size (1e6 photons: a
    fraction of one wave of the persistent kernel) and filled up to 2^20 steps, C3's ice (SpiceLea + tilt +
    anisotropy) under a muon bundle, C5 flasher at oversize 1.  N = 1 only."""
    from clsim_b200 import capi, geometry, ice, steps
    from clsim_b200.description import KERNEL_FAST, ConverterOptions
    k = max(3, args.steps // 4)
    out = {}

    def one(name, workload, medium, geo, gens, bias, bunch, pancake):
        opt = ConverterOptions(device=local, stop_detected_photons=True, pancake_factor=pancake, kernel_mode=KERNEL_FAST,
                               max_num_workitems=len(bunch), rng_seed=777, rng_first_multiplier=0)
        with capi.Engine(medium, geo, gens, bias, opt) as eng:
            eng.upload_resident(bunch)
            eng.run_resident(args.warmup)
            r = eng.run_resident(k)
        out[name] = {"workload": workload, "value": r["photons"] / (r["kernel_ms"] * 1e-3), "unit": "photons/s", "steps": k,
                     "ms_per_step": r["kernel_ms"] / k, "hit_fraction": r["hits"] / float(max(1, r["photons"])),
                     "segments_per_photon": r["segments"] / float(max(1, r["photons"])),
                     "roofline_frac": roofline_frac(r["photons"], r["segments"], r["kernel_ms"], sms, sm_mhz)}

    geo5 = geometry.make_ic86_like_geometry(oversize=5.0)
    bias5 = ice.GetIceCubeDOMAcceptance(domRadius=geometry.DOM_RADIUS * 5.0)
    hom = ice.MakeHomogeneousIceMediumProperties("spice_mie")
    gen = [ice.makeCherenkovWavelengthGenerator(bias5, False, hom)]
    one("config1", "point source at the origin, 5000 steps x 200 photons = 1e6 photons (the stated size), homogeneous bulk ice, IC86-like, oversize 5",
        hom, geo5, gen, bias5, steps.point_source_steps(5000, 200, seed=1), 5.0)
    one("config1_filled", "the same source, 2^20 steps x 200 photons (fills the device)", hom, geo5, gen, bias5,
        steps.point_source_steps(1 << 20, 200, seed=1), 5.0)
    lea = ice.MakeIceCubeMediumProperties(iceDataDirectory="spice_lea", useTiltIfAvailable=True)
    gen = [ice.makeCherenkovWavelengthGenerator(bias5, False, lea)]
    one("config3", "muon bundle (100 muons, 20 m spread), SpiceLea + tilt + anisotropy, IC86-like, oversize 5, 2^20 steps x 200 photons",
        lea, geo5, gen, bias5, steps.muon_bundle_steps(1 << 20, num_muons=100, seed=3), 5.0)
    geo1 = geometry.make_ic86_like_geometry(oversize=1.0)
    bias1 = ice.GetIceCubeDOMAcceptance(domRadius=geometry.DOM_RADIUS)
    wl, val = ice.GetFlasherLED405Spectrum()
    gens = [ice.makeCherenkovWavelengthGenerator(bias1, False, lea), ice.makeWavelengthGenerator(wl, val, bias1, lea)]
    d2 = (geo1.posX - 0.0) ** 2 + (geo1.posY - 0.0) ** 2 + (geo1.posZ + 200.0) ** 2
    i = int(np.argmin(d2))
    dom = np.array([geo1.posX[i], geo1.posY[i], geo1.posZ[i]])
    one("config5", "flasher LED (405 nm spectrum, sourceType 1) inside a DOM, SpiceLea + tilt + anisotropy, oversize 1 (no pancake), 2^20 steps x 200 photons",
        lea, geo1, gens, bias1, steps.flasher_steps(1 << 20, dom, seed=5), 1.0)
    return out


def e2e_through_engine(eng, bunches, first_id=100):
    """EnqueueSteps / GetConversionResult over a list of host bunches; -> (seconds, hits, results as (id, photons))."""
    t0 = time.perf_counter()
    pending, got = 0, []
    for i, b in enumerate(bunches):
        eng.enqueue(b, first_id + i)
        pending += 1
        while eng.more_photons_available():
            r = eng.get_result()
            got.append((r.identifier, r.photons))
            pending -= 1
    while pending:
        r = eng.get_result()
        got.append((r.identifier, r.photons))
        pending -= 1
    return time.perf_counter() - t0, got


def run_guarded(fn, deadline_s, out):
    """Run `fn(out)` in a worker thread and give up after `deadline_s`: the side legs must not be able to take the
    headline line down with them (a rank that never reaches a collective would hold every other rank for NCCL's ten
    minutes).  -> None, or the reason the legs did not finish; `out` keeps what finished."""
    import threading
    box = {}

    def work():
        try:
            fn(out)
        except BaseException as ex:   # noqa: B902 -- reported, not swallowed
            box["error"] = "%s: %s" % (type(ex).__name__, ex)

    t = threading.Thread(target=work, daemon=True)
    t.start()
    t.join(deadline_s)
    if t.is_alive():
        return "no result within %d s (legs finished before that are kept)" % deadline_s
    return box.get("error")


def first_error(dist, group, world, err):
    """-> `err` (this rank's error text or None), or what another rank said.  Every rank learns, over the CPU-side group,
    whether all of them got through their own part BEFORE anyone enters the collectives that follow: a rank that failed
    would be missing there, and the others would wait for it until the watchdog."""
    if world == 1:
        return err
    said = [None] * world
    dist.all_gather_object(said, err, group=group)
    for r, e in enumerate(said):
        if e:
            return err or ("rank %d: %s" % (r, e))
    return None


def run_multi_gpu_legs(args, rank, world, local, dist, gloo, barrier, max_over_ranks, sum_over_ranks, out):
    """What the weak-scaling headline cannot show (every rank its own feeder, nothing shared):
    (a) STRONG scaling: ONE fixed step series of config 3 (muon bundle, SpiceLea + tilt + anisotropy) split over the
        ranks (sharding.split_steps), propagated end to end from host buffers, hit lists gathered on rank 0 and merged
        by identifier (sharding.merge_results); time from the barrier to the merged result;
    (b) ONE PROCESS, N devices: rank 0 alone drives one converter per GPU behind the in-process server with a single
        feeder -- the reference's own topology (I3CLSimModule.cxx:611-638, I3CLSimServer.cxx:126-135) -- while the other
        ranks wait on a CPU barrier (an NCCL barrier would sit on the SMs the persistent kernel needs);
    (c) config 4 at its stated size: cascade steps made on the device, photo-electrons out, 1.25e10 propagated photons
        per GPU (= the config's ~1e11 on eight)."""
    import torch
    from clsim_b200 import capi, geometry, ice, mcpe, stepgen, steps
    from clsim_b200.converter import configureCUDADevices, initializeCUDA
    from clsim_b200.description import KERNEL_FAST, ConverterOptions
    from clsim_b200.server import I3CLSimServerInProcess
    from clsim_b200.sharding import RNG_ROWS_PER_DEVICE, mcpe_row_offset, merge_results, rng_row_offset, split_steps, stepgen_row_offset
    torch.cuda.set_device(local)   # (called from a worker thread, see run_guarded)
    lea = ice.MakeIceCubeMediumProperties(iceDataDirectory="spice_lea", useTiltIfAvailable=True)
    geo = geometry.make_ic86_like_geometry(oversize=5.0)
    bias = ice.GetIceCubeDOMAcceptance(domRadius=geometry.DOM_RADIUS * 5.0)
    gens = [ice.makeCherenkovWavelengthGenerator(bias, False, lea)]
    bunch_steps = args.bunch

    def cpu_barrier():
        if world > 1:
            dist.barrier(group=gloo)

    def any_failed(err):
        return first_error(dist, gloo, world, err)

    def text(ex):
        sys.stderr.write("bench.py rank %d: %s: %s\n" % (rank, type(ex).__name__, ex))
        return "%s: %s" % (type(ex).__name__, ex)

    # ---------------- (a) strong scaling -------------------------------------------------------------------------
    total_steps = 16 * bunch_steps   # (16 pieces at most per rank: the size of the meta record below)
    series = steps.muon_bundle_steps(total_steps, num_muons=100, seed=3)   # the same series on every rank (seeded)
    shard = split_steps(series, world)[rank]
    pieces = [shard[i:i + bunch_steps] for i in range(0, len(shard), bunch_steps)]
    opt = ConverterOptions(device=local, stop_detected_photons=True, pancake_factor=5.0, kernel_mode=KERNEL_FAST, enable_double_buffering=True,
                           max_num_workitems=bunch_steps, rng_seed=5150 + rank, rng_first_multiplier=rng_row_offset(rank))
    err, eng, got = None, None, []
    try:
        eng = capi.Engine(lea, geo, gens, bias, opt)
        e2e_through_engine(eng, pieces[:1] * 2)   # warm the staging pool
    except Exception as ex:   # noqa: BLE001 -- reported in the line, the other legs still run
        err = text(ex)
    err = any_failed(err)
    if not err:
        barrier()
        t0 = time.perf_counter()
        try:
            # bunch identifiers name (rank, piece): the caller's key for putting results back together (I3CLSimClientModule.cxx:359-439)
            _, got = e2e_through_engine(eng, pieces, first_id=1000 * rank)
        except Exception as ex:   # noqa: BLE001
            err = text(ex)
        t_prop = time.perf_counter() - t0
        err = any_failed(err)
    if not err:
        # hand the hit lists to rank 0: lengths first, then the records as bytes over NCCL (padded to the longest)
        ids = np.array([i for i, _ in got], dtype=np.int64)
        lens = np.array([len(p) for _, p in got], dtype=np.int64)
        mine = np.concatenate([p for _, p in got]) if got else np.zeros(0, dtype=capi.PHOTON_DTYPE)
        if world > 1:
            meta = torch.zeros(2 * 16 + 1, dtype=torch.int64, device="cuda")
            meta[0] = len(got)
            meta[1:1 + len(got)] = torch.from_numpy(ids).cuda()
            meta[17:17 + len(got)] = torch.from_numpy(lens).cuda()
            metas = torch.empty(world * meta.numel(), dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(metas, meta)
            metas = metas.view(world, -1).cpu().numpy()
            longest = int(max(metas[r, 17:17 + metas[r, 0]].sum() for r in range(world))) * mine.dtype.itemsize
            payload = torch.zeros(max(16, longest), dtype=torch.uint8, device="cuda")
            payload[:mine.nbytes] = torch.from_numpy(mine.view(np.uint8)).cuda()
            everything = torch.empty(world * payload.numel(), dtype=torch.uint8, device="cuda")
            dist.all_gather_into_tensor(everything, payload)
            results = []
            if rank == 0:
                host = everything.view(world, -1).cpu().numpy()
                for r in range(world):
                    k = int(metas[r, 0])
                    recs = host[r, :int(metas[r, 17:17 + k].sum()) * mine.dtype.itemsize].view(mine.dtype)
                    at = 0
                    for j in range(k):
                        results.append((int(metas[r, 1 + j]), recs[at:at + int(metas[r, 17 + j])]))
                        at += int(metas[r, 17 + j])
        else:
            results = got
        merged_bunches, merged_hits = 0, 0
        if rank == 0:
            merged = merge_results(results)
            merged_bunches, merged_hits = len(merged), int(sum(len(v) for v in merged.values()))
        t_all = time.perf_counter() - t0
        t_prop = max_over_ranks(t_prop)
        t_all = max_over_ranks(t_all)
        photons_total = float(series["num_photons"].sum())
        out["strong_scaling_config3"] = {
            "workload": "ONE muon-bundle step series (config 3: SpiceLea + tilt + anisotropy), %d steps x %d photons, split over the ranks; host buffers in, hit lists handed to rank 0 and merged by bunch identifier" % (total_steps, PHOTONS_PER_STEP),
            "value": photons_total / t_all, "unit": "photons/s", "scaling": "strong", "seconds_to_merged_result": t_all,
            "seconds_to_last_rank_result": t_prop, "bunches_merged": merged_bunches, "hits_merged": merged_hits, "photons": photons_total}
    else:
        out["strong_scaling_config3"] = {"error": err}
    if eng is not None:
        eng.close()

    # ---------------- (b) one process, N converters behind the server seam ------------------------------------------
    cpu_barrier()
    if rank == 0:
        try:
            c2 = build_scene()
            devices = configureCUDADevices(UseGPUs=True, OverrideApproximateNumberOfWorkItems=bunch_steps, numDevices=world)
            converters = [initializeCUDA(dev, 6000 + i, c2[1], c2[0], c2[3], c2[2], enableDoubleBuffering=True, stopDetectedPhotons=True, pancakeFactor=5.0,
                                         kernelMode=KERNEL_FAST, rngFirstMultiplierRow=i * RNG_ROWS_PER_DEVICE) for i, dev in enumerate(devices)]
            server = I3CLSimServerInProcess(converters)
            client = server.Connect()
            bunch = make_bunch(bunch_steps, seed=4000)
            per_gpu = 6
            for i in range(2 * world):   # warm every converter's staging pool
                client.EnqueueSteps(bunch, i)
            for i in range(2 * world):
                client.GetConversionResult()
            t0 = time.perf_counter()
            hits, sent, received, window = 0, 0, 0, 4 * world   # one feeder: at most `window` bunches in flight
            total = per_gpu * world
            while received < total:
                while sent < total and sent - received < window:
                    client.EnqueueSteps(bunch, sent)
                    sent += 1
                hits += len(client.GetConversionResult().photons)
                received += 1
            dt = time.perf_counter() - t0
            st = server.GetStatistics()
            calls = [st.get("NumKernelCalls" + ("" if world == 1 else "_%d" % i), 0.0) for i in range(world)]
            server.Close()
            for c in converters:
                c.Close()
            out["one_process_n_converters"] = {
                "workload": "config 2 bunches (2^20 steps x 200 photons) from ONE feeder thread through I3CLSimServerInProcess to %d converters, one per GPU, host buffers in and hit lists out" % world,
                "value": total * float(bunch["num_photons"].sum()) / dt, "unit": "photons/s", "bunches": total, "seconds": dt, "hits": hits,
                "kernel_calls_per_converter": calls}
        except Exception as ex:   # noqa: BLE001
            out["one_process_n_converters"] = {"error": text(ex)}
    cpu_barrier()

    # ---------------- (c) config 4 at its stated size -----------------------------------------------------------------
    want_photons = 1.25e10
    err, conv, pe, eng = None, None, None, None
    sent = photons = hits = pes = 0
    energy, dt = 0.0, 0.0
    try:
        ang = mcpe.GetIceCubeDOMAngularSensitivity()
        acc = ice.GetIceCubeDOMAcceptance(domRadius=geometry.DOM_RADIUS * 5.0, efficiency=0.9 * mcpe.GetHoleIcePeak())
        gen4 = ice.makeCherenkovWavelengthGenerator(acc, False, lea)
        opt4 = ConverterOptions(device=local, stop_detected_photons=True, pancake_factor=5.0, kernel_mode=KERNEL_FAST, enable_double_buffering=True,
                                max_num_workitems=bunch_steps, rng_seed=7000 + rank, output_photons_per_workitem=2, rng_first_multiplier=rng_row_offset(rank))
        conv = stepgen.I3CLSimLightSourceToStepConverterPPC(photonsPerStep=PHOTONS_PER_STEP, device=local)
        conv.SetMediumProperties(lea)
        conv.SetWlenBias(acc)
        conv.SetRandomService(40 + rank)
        conv.Initialize(rngFirstMultiplierRow=stepgen_row_offset(rank))
        pe = mcpe.I3CLSimPhotonToMCPEConverterForDOMs(50 + rank, {(int(s), int(o)): acc for s, o in zip(geo.stringIDs, geo.domIDs)}, ang,
                                                     device=local, rngFirstMultiplierRow=mcpe_row_offset(rank))
        eng = capi.Engine(lea, geo, [gen4], acc, opt4)
        pe.attach_to(eng)
        vertex, axis = (20.0, -30.0, -250.0), (0.3, 0.2, -0.93)
        conv.EnqueueLightSource(stepgen.Particle("EMinus", 1e3, vertex, axis), 0)   # warm-up, and the yield per GeV
        small = 0
        while conv.EnqueueInto(eng, 0):
            small += eng.get_result().num_photons_generated
        energy = 1e3 * want_photons / max(1.0, small)
        conv.EnqueueLightSource(stepgen.Particle("EMinus", energy, vertex, axis), 1)
        conv.EnqueueBarrier()
    except Exception as ex:   # noqa: BLE001
        err = text(ex)
    err = any_failed(err)
    if not err:
        barrier()
        t0 = time.perf_counter()
        try:
            pending = 0
            while True:
                if conv.EnqueueInto(eng, 100 + sent) == 0:
                    break
                sent += 1
                pending += 1
                while eng.more_photons_available():
                    r = eng.get_result(); pending -= 1
                    photons += r.num_photons_generated; hits += r.num_hits_counted; pes += len(r.mcpes)
            while pending:
                r = eng.get_result(); pending -= 1
                photons += r.num_photons_generated; hits += r.num_hits_counted; pes += len(r.mcpes)
        except Exception as ex:   # noqa: BLE001
            err = text(ex)
        dt = time.perf_counter() - t0
        err = any_failed(err)
    for thing in (eng, pe):
        try:
            if thing is not None:
                thing.close()
        except Exception as ex:   # noqa: BLE001 -- (an engine that died of a device error cannot free its buffers either)
            text(ex)
    if err:
        out["config4_cascade"] = {"error": err}
        return out
    dt = max_over_ranks(dt)
    out["config4_cascade"] = {
        "workload": "config 4: e- cascade in SpiceLea + tilt + anisotropy, steps made on the device from the step-generation queue entry, photo-electrons out; %.3g propagated photons per GPU (a %.3g GeV e- with the DOM-acceptance bias; the config's ~1e11 photons on eight GPUs)" % (want_photons, energy),
        "value": sum_over_ranks(float(photons)) / dt, "unit": "photons/s", "scaling": "weak", "seconds": dt, "photons": sum_over_ranks(float(photons)),
        "hits": sum_over_ranks(float(hits)), "mcpes": sum_over_ranks(float(pes)), "bunches_per_gpu": sent}
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
    ap.add_argument("--multi-gpu-legs", action="store_true", help=argparse.SUPPRESS)   # run the N > 1 legs at N = 1 too (the series' first point)
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
    gloo = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        gloo = dist.new_group(backend="gloo")   # CPU-side barriers and the host-side merge of hit lists

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

    peaks = measured_peaks()
    props = torch.cuda.get_device_properties(local)
    sms = props.multi_processor_count
    other = None
    if world == 1 and not args.no_variants and args.ice == "spice_mie" and not args.tilt:
        other = run_other_configs(args, local, sms, peaks["sm_max_mhz"])
    multi, legs_failed = None, None
    if not args.no_variants and args.ice == "spice_mie" and not args.tilt and (world > 1 or args.multi_gpu_legs):
        multi = {}
        legs_failed = run_guarded(lambda out: run_multi_gpu_legs(args, rank, world, local, dist, gloo, barrier, max_over_ranks, sum_over_ranks, out),
                                  240.0, multi)
        if legs_failed:
            multi["error"] = legs_failed

    def leave():
        sys.stdout.flush()
        if legs_failed:
            os._exit(0)   # a leg is still stuck in a worker thread (or in a collective): no teardown that could wait for it
        if world > 1:
            dist.destroy_process_group()

    if rank != 0:
        leave()
        return

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
        "other_configs": other,
        "multi_gpu": multi,
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
    leave()


if __name__ == "__main__":
    main()
