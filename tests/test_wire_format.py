"""Row f1, wire format: the body of a serialized step / photon series (clsim_b200/wire.py) against a byte-level fixture
assembled by hand from the reference's description (tests/golden/make_wire_fixture.py), round trips, and the reader's
refusals (version, truncation, size mismatch)."""
import os

import numpy as np
import pytest

from clsim_b200 import steps, wire
from clsim_b200.description import PHOTON_DTYPE, STEP_DTYPE

HERE = os.path.dirname(os.path.abspath(__file__))


def _hex(name):
    with open(os.path.join(HERE, "golden", name)) as f:
        return bytes.fromhex(f.read().strip())


def test_step_series_against_the_hand_made_fixture():
    body = _hex("wire_step_series.hex")
    got = wire.unpack_step_series(body)
    assert len(got) == 2 and got.dtype == STEP_DTYPE
    assert got["num_photons"].tolist() == [200, 65539] and got["identifier"].tolist() == [7, 0xDEADBEEF]
    assert got["source_type"].tolist() == [0, 1] and got["z"].tolist() == [3.5, -400.0]
    assert wire.pack_step_series(got) == body
    # a count that needs two bytes
    many = np.zeros(300, dtype=STEP_DTYPE)
    assert wire.pack_step_series(many)[:6] == _hex("wire_step_series_300_header.hex")


def test_round_trips_and_integer_rule():
    bunch = steps.muon_track_steps(1237, seed=5)
    back = wire.unpack_step_series(wire.pack_step_series(bunch))
    assert back.tobytes() == np.ascontiguousarray(bunch, dtype=STEP_DTYPE).tobytes()
    rng = np.random.default_rng(3)
    photons = np.frombuffer(rng.integers(0, 256, 80 * 513, dtype=np.uint8).tobytes(), dtype=PHOTON_DTYPE)
    blob = wire.pack_photon_series(photons)
    assert len(blob) == 2 + 1 + 3 + 80 * 513            # base, version 0, count 513 = size byte + 2 bytes, records
    assert wire.unpack_photon_series(blob).tobytes() == photons.tobytes()
    empty = wire.pack_photon_series(np.zeros(0, dtype=PHOTON_DTYPE))
    assert empty == b"\x00\x00\x00\x00" and len(wire.unpack_photon_series(empty)) == 0
    for v in (0, 1, 255, 256, 65535, 65536, 2 ** 32 - 1, 2 ** 32, 2 ** 64 - 1):
        enc = wire.put_uint(v)
        assert wire.get_uint(enc, 0) == (v, len(enc)) and len(enc) == (1 if v == 0 else 1 + (v.bit_length() + 7) // 8)


def test_reader_refuses_what_the_reference_refuses():
    body = bytearray(_hex("wire_step_series.hex"))
    newer = bytes(body[:2]) + b"\x01\x01" + bytes(body[3:])       # version 1
    with pytest.raises(wire.WireError, match="can only read I3Vector<I3CLSimStep> version 0, but 1 was provided"):
        wire.unpack_step_series(newer)
    with pytest.raises(wire.WireError, match="blob of"):
        wire.unpack_step_series(bytes(body[:-1]))
    with pytest.raises(wire.WireError, match="blob of"):
        wire.unpack_photon_series(bytes(body))                    # a step series is not a photon series
    with pytest.raises(wire.WireError):
        wire.unpack_step_series(b"\x00\x00\x00")


# ----------------------------------------------------------------------------------------------------------------
# ... and against the reference's own records and serialization code, compiled unmodified (oracle/_ref/libclsim_ref_wire.so:
# public/clsim/I3CLSimStep.h, I3CLSimPhoton.h, private/clsim/I3CLSimStep.cxx, I3CLSimPhoton.cxx; the archive classes they
# write to are stand-ins that restate the portable archive's encoding of ONE value -- what is written, of which type and in
# which order, is the reference's code)
# ----------------------------------------------------------------------------------------------------------------
from oracle import pyoracle  # noqa: E402

needs_ref = pytest.mark.skipif(not pyoracle.ref_wire_available(), reason="oracle/_ref not built (no /root/reference at build time)")


def _random_records(dtype, n, seed):
    rng = np.random.default_rng(seed)
    rec = np.zeros(n, dtype=dtype)
    for name in dtype.names:
        kind = dtype[name].kind
        if kind == "f":
            rec[name] = rng.normal(size=n).astype(np.float32) * 100
        else:
            info = np.iinfo(dtype[name])
            rec[name] = rng.integers(info.min, int(info.max) + 1, n, dtype=np.int64).astype(dtype[name])
    return rec


@needs_ref
def test_record_layouts_are_the_references():
    """T1 / T2: a record filled through the reference's own SETTERS is, byte for byte, the product's clsimcu_step / clsimcu_photon."""
    import ctypes as C
    L = pyoracle.ref_wire_lib()
    assert L.ref_wire_step_size() == STEP_DTYPE.itemsize == wire.STEP_BLOB == 48
    assert L.ref_wire_photon_size() == PHOTON_DTYPE.itemsize == wire.PHOTON_BLOB == 80
    assert L.ref_wire_step_version() == wire.STEP_VERSION and L.ref_wire_photon_version() == wire.PHOTON_VERSION
    for s in _random_records(STEP_DTYPE, 50, 1):
        f = np.array([s[k] for k in ("x", "y", "z", "t", "theta", "phi", "length", "beta", "weight")], dtype=np.float32)
        u = np.array([s[k] for k in ("num_photons", "identifier", "source_type", "dummy1", "dummy2")], dtype=np.uint32)
        out = C.create_string_buffer(48)
        L.ref_wire_make_step(f.ctypes.data, u.ctypes.data, out)
        assert out.raw == s.tobytes()
        f2, u2 = np.zeros(9, np.float32), np.zeros(5, np.uint32)
        L.ref_wire_read_step_fields(s.tobytes(), f2.ctypes.data, u2.ctypes.data)      # ... and back through its getters
        assert f2.tobytes() == f.tobytes() and np.array_equal(u2, u)
    for p in _random_records(PHOTON_DTYPE, 50, 2):
        f = np.array([p[k] for k in ("x", "y", "z", "t", "theta", "phi", "wavelength", "cherenkov_dist", "weight", "start_x", "start_y",
                                     "start_z", "start_t", "start_theta", "start_phi", "group_velocity", "dist_in_abs_lens")], dtype=np.float32)
        u = np.array([p[k] for k in ("num_scatters", "identifier", "string_id", "om_id")], dtype=np.int64)
        out = C.create_string_buffer(80)
        L.ref_wire_make_photon(f.ctypes.data, u.ctypes.data, out)
        assert out.raw == p.tobytes()


@needs_ref
@pytest.mark.parametrize("n", [0, 1, 2, 255, 256, 300, 70000])
def test_series_body_is_what_the_references_serialize_writes(n):
    """I3Vector<I3CLSimStep>::serialize / I3Vector<I3CLSimPhoton>::serialize (portable_binary_oarchive) == wire.pack_*; and
    the reference's reader takes what wire.pack_* writes, to the last byte."""
    steps_ = _random_records(STEP_DTYPE, n, 3)
    body = pyoracle.ref_wire_write(steps_)
    assert body == wire.pack_step_series(steps_)
    got, left = pyoracle.ref_wire_read(wire.pack_step_series(steps_), STEP_DTYPE)
    assert left == 0 and got.tobytes() == steps_.tobytes()
    assert wire.unpack_step_series(body).tobytes() == steps_.tobytes()
    photons = _random_records(PHOTON_DTYPE, n, 4)
    body = pyoracle.ref_wire_write(photons)
    assert body == wire.pack_photon_series(photons)
    got, left = pyoracle.ref_wire_read(wire.pack_photon_series(photons), PHOTON_DTYPE)
    assert left == 0 and got.tobytes() == photons.tobytes()


@needs_ref
def test_the_hand_made_fixture_is_what_the_reference_writes():
    body = _hex("wire_step_series.hex")
    steps_, left = pyoracle.ref_wire_read(body, STEP_DTYPE)
    assert left == 0 and len(steps_) == 2 and steps_["identifier"][1] == 0xDEADBEEF
    assert pyoracle.ref_wire_write(steps_) == body
    assert pyoracle.ref_wire_write(np.zeros(300, dtype=STEP_DTYPE))[:6] == _hex("wire_step_series_300_header.hex")


@needs_ref
def test_the_reference_refuses_what_the_reader_refuses():
    body = bytearray(_hex("wire_step_series.hex"))
    newer = bytes(body[:2]) + b"\x01\x01" + bytes(body[3:])       # version 1
    with pytest.raises(RuntimeError, match="can only read I3Vector<I3CLSimStep> version 0, but 1 was provided"):
        pyoracle.ref_wire_read(newer, STEP_DTYPE)
    with pytest.raises(RuntimeError):
        pyoracle.ref_wire_read(bytes(body[:-1]), STEP_DTYPE)      # short blob
