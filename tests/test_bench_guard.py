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
