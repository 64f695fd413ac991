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
