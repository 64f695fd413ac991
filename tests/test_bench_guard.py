"""bench.py's watchdog around the multi-GPU side legs: a leg that never returns (a rank stuck before a collective) must not
take the headline line down; what finished before is kept."""
import importlib.util
import os
import time

HERE = os.path.dirname(os.path.abspath(__file__))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(os.path.dirname(HERE), "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_guard_keeps_finished_legs_and_reports_the_rest():
    bench = _bench()
    out = {}

    def legs(o):
        o["first"] = {"value": 1.0}
        time.sleep(30.0)          # the second leg hangs
        o["second"] = {"value": 2.0}

    t0 = time.perf_counter()
    why = bench.run_guarded(legs, 0.5, out)
    assert time.perf_counter() - t0 < 5.0
    assert why is not None and "no result within" in why
    assert out == {"first": {"value": 1.0}}

    out = {}
    assert bench.run_guarded(lambda o: o.update(done=True), 5.0, out) is None and out == {"done": True}

    def broken(o):
        raise RuntimeError("converter failed")

    assert "RuntimeError: converter failed" in bench.run_guarded(broken, 5.0, {})


def _two_rank_worker(rank, world, port, out_dir):
    import json
    import sys

    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    group = dist.new_group(backend="gloo")
    bench = _bench()
    seen = {}
    # nobody failed; then rank 1 fails its own part of a leg: both ranks must know before the leg's collectives
    seen["clean"] = bench.first_error(dist, group, world, None)
    seen["one_failed"] = bench.first_error(dist, group, world, "ClsimCudaError: boom" if rank == 1 else None)
    # a rank that never arrives: the others sit in a collective behind it, the watchdog lets every rank go with what it has
    out = {}

    def legs(o):
        o["first"] = rank
        if rank == 1:
            time.sleep(60.0)
        dist.barrier(group=group)
        o["second"] = rank

    t0 = time.perf_counter()
    seen["why"] = bench.run_guarded(legs, 2.0, out)
    seen["waited"] = time.perf_counter() - t0
    seen["out"] = out
    with open(os.path.join(out_dir, "rank%d.json" % rank), "w") as f:
        json.dump(seen, f)
    sys.stdout.flush()
    os._exit(0)   # what bench.py's leave() does after a failed leg: no teardown that could wait for the stuck thread


def test_ranks_agree_on_a_failed_leg_and_leave_a_stuck_one(tmp_path):
    import json
    import socket

    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_two_rank_worker, args=(r, 2, port, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(90.0)
        assert p.exitcode == 0
    for r in range(2):
        with open(os.path.join(str(tmp_path), "rank%d.json" % r)) as f:
            seen = json.load(f)
        assert seen["clean"] is None
        assert "boom" in seen["one_failed"] and (r == 1 or seen["one_failed"].startswith("rank 1:"))
        assert "no result within" in seen["why"] and seen["waited"] < 30.0
        assert seen["out"] == {"first": r}


def test_the_step_at_infinity_that_stopped_the_eight_gpu_run():
    """profiles/bench_r02_v38_n8_config4_hang.err: rank 6 of 8 never came back from the config-4 leg.  The step generator is
    deterministic, so the leg's steps can be replayed on the CPU (tools/replay_config4_stepgen.py): with rank 6's seeds the
    reference's gamma sampler draws ry == 1 in the fifteenth launch and puts one cascade step at infinity -- on no other
    rank (profiles/replay_config4_stepgen_r02.txt).  The fast kernel now ends such steps on the spot
    (tests/test_zz_gpu_steps_at_infinity.py); this test keeps the diagnosis reproducible."""
    import subprocess
    import sys
    tool = os.path.join(os.path.dirname(HERE), "tools", "replay_config4_stepgen.py")
    text = subprocess.run([sys.executable, tool, "--rank", "6", "--launches", "15"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                          timeout=300).stdout
    assert "launch 14: 1 step(s) at infinity, the first is step 497015 (thread 42359)" in text, text
    text = subprocess.run([sys.executable, tool, "--rank", "5", "--launches", "15"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                          timeout=300).stdout
    assert "no step at infinity" in text, text
