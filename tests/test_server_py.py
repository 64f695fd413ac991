"""Python twin of the in-process server seam (clsim_b200/server.py) against a DummyConverter, restating the reference's only
interface test (resources/tests/testCLSimServer.py:26-78): ragged bunches in, every result back under the sender's
identifier with one photon per step; plus the multi-converter handshake and the failure path."""
import threading

import numpy as np
import pytest

from clsim_b200.converter import ConversionResult_t
from clsim_b200.description import PHOTON_DTYPE, STEP_DTYPE
from clsim_b200.server import I3CLSimServerInProcess, ServerFailure


class DummyConverter(object):
    def __init__(self, workgroup=1, max_items=64, good=10 ** 9):
        self.workgroup, self.max_items, self.good = workgroup, max_items, good
        self.lock = threading.Lock()
        self.queue = []
        self.calls = 0

    def IsInitialized(self):
        return True

    def GetWorkgroupSize(self):
        return self.workgroup

    def GetMaxNumWorkitems(self):
        return self.max_items

    def EnqueueSteps(self, steps, identifier):
        with self.lock:
            self.good -= 1
            if self.good < 0:
                raise RuntimeError("device lost")
            self.queue.append((steps, identifier))

    def GetConversionResult(self):
        with self.lock:
            steps, identifier = self.queue.pop(0)
            self.calls += 1
        photons = np.zeros(len(steps), dtype=PHOTON_DTYPE)
        photons["identifier"] = steps["identifier"]
        photons["num_scatters"] = steps["num_photons"]
        return ConversionResult_t(identifier, photons, None)

    def GetStatistics(self):
        return {"NumKernelCalls": float(self.calls)}


def make_steps(n, tag):
    s = np.zeros(n, dtype=STEP_DTYPE)
    s["identifier"] = tag
    s["num_photons"] = np.arange(n) % 7
    return s


def exercise(client, n_bunches, seed):
    rng = np.random.default_rng(seed)
    sizes = {}
    for i in range(n_bunches):
        sizes[i] = int(rng.integers(1, client.GetMaxNumWorkitems() + 1))
        client.EnqueueSteps(make_steps(sizes[i], 1000 * seed + i), i)
    for _ in range(n_bunches):
        r = client.GetConversionResult()
        assert len(r.photons) == sizes.pop(r.identifier)
        assert np.all(r.photons["identifier"] == 1000 * seed + r.identifier)
    assert not sizes and client.GetConversionResult() is None


def test_one_and_many_clients():
    server = I3CLSimServerInProcess([DummyConverter()])
    exercise(server.Connect(), 10, 1)
    threads = [threading.Thread(target=exercise, args=(server.Connect(), 25, 2 + k)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert server.GetStatistics()["NumKernelCalls"] == 10 + 4 * 25
    server.Close()


def test_bunch_size_handshake_over_converters():
    server = I3CLSimServerInProcess([DummyConverter(4, 100), DummyConverter(6, 64)])
    assert server.GetWorkgroupSize() == 12 and server.GetMaxNumWorkitems() == 60     # I3CLSimServer.cxx:95-113
    st = server.GetStatistics()
    assert "NumKernelCalls_0" in st and "NumKernelCalls_1" in st
    server.Close()
    with pytest.raises(RuntimeError, match="incompatible"):
        I3CLSimServerInProcess([DummyConverter(64, 64), DummyConverter(48, 100)])
    with pytest.raises(RuntimeError):
        I3CLSimServerInProcess([])


def test_converter_failure_fails_the_clients():
    server = I3CLSimServerInProcess([DummyConverter(good=2)])
    client = server.Connect()
    got = 0
    with pytest.raises(ServerFailure, match="device lost"):
        for i in range(6):
            client.EnqueueSteps(make_steps(8, 1), i)
        for i in range(6):
            r = client.GetConversionResult()
            assert len(r.photons) == 8     # a result that does arrive is a real one
            got += 1
    assert got <= 2 and "device lost" in server.Failure()
    with pytest.raises(ServerFailure):
        client.EnqueueSteps(make_steps(8, 1), 99)
    server.Close()
