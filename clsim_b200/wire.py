"""SURVEY 8(f) row f1, the wire format: the BODY of a serialized step / photon series as clsim itself defines it.

``I3Vector<I3CLSimStep>::serialize(portable_binary_oarchive)`` (private/clsim/I3CLSimStep.cxx:126-144) and its
photon twin (private/clsim/I3CLSimPhoton.cxx:139-170) write, after the ``I3FrameObject`` base,

    unsigned  I3CLSimStep_version / I3CLSimPhoton_version   (0; a reader refuses any other, :131-134)
    uint64    num
    binary    blob: num records of blobSizeV0 = 48 / 80 bytes, the in-memory little-endian structs themselves
              (I3CLSimStep.cxx:96-104 "just save the binary blob", BOOST_STATIC_ASSERT on the size)

which is what the server and its clients exchange over ZeroMQ (I3CLSimServer.cxx:46-70, 320, 339, 386, 408).  The
records are byte for byte ``clsimcu_step`` / ``clsimcu_photon`` (include/clsimcuda.h), so a series goes from the wire
into ``clsimcu_enqueue`` -- and from ``clsimcu_get_result`` onto the wire -- without a per-record conversion.

Integers are written the way the portable binary archive does it (eos portable archive, the base of
``icecube::archive::portable_binary_oarchive``): one signed size byte, then that many bytes of the value,
least-significant first; zero is the single byte 0x00.

Pinned (tests/test_wire_format.py): the reference's own records and serialize() members (public/clsim/I3CLSimStep.h,
I3CLSimPhoton.h, private/clsim/I3CLSimStep.cxx, I3CLSimPhoton.cxx), compiled unmodified into oracle/_ref/libclsim_ref_wire.so,
write the bytes ``pack_*`` writes and read what ``pack_*`` wrote; a record filled through the reference's setters is byte for
byte ``clsimcu_step`` / ``clsimcu_photon``.  (The archive those members write to is a stand-in that restates the portable
archive's encoding of one value; which values, of which C++ type, in which order is the reference's code.)

NOT reproduced -- parity unpinned, and stated as such: the archive preamble and the object framing around the body
(archive signature and version, class id / class name / tracking / object id of the shared pointer, the
``I3FrameObject`` base's class-info bytes).  They belong to IceTray's ``serialization`` project, which is not vendored
under /root/reference; a maintainer wiring this into the real server lets the archive write them and calls
``pack_body`` / ``unpack_body`` for what lies in between.  ``FRAME_OBJECT_BASE`` is the two zero bytes (tracking flag,
class version 0) such a base writes under boost's default traits -- a stated assumption, kept in one place.
"""
import numpy as np

from .description import PHOTON_DTYPE, STEP_DTYPE

STEP_VERSION = 0        # i3clsimstep_version_   (public/clsim/I3CLSimStep.h:66)
PHOTON_VERSION = 0      # i3clsimphoton_version_ (public/clsim/I3CLSimPhoton.h:65)
STEP_BLOB = 48          # blobSizeV0 (private/clsim/I3CLSimStep.cxx:35)
PHOTON_BLOB = 80        # blobSizeV0 (private/clsim/I3CLSimPhoton.cxx:36)
FRAME_OBJECT_BASE = b"\x00\x00"


class WireError(ValueError):
    pass


def put_uint(value):
    """Portable-archive integer: size byte + little-endian bytes, zero as a single 0x00."""
    value = int(value)
    if value < 0:
        raise WireError("unsigned value expected")
    if value == 0:
        return b"\x00"
    n = (value.bit_length() + 7) // 8
    return bytes([n]) + value.to_bytes(n, "little")


def get_uint(buf, at, max_bytes=8):
    if at >= len(buf):
        raise WireError("truncated: integer expected at byte %d" % at)
    n = buf[at]
    if n > 127:
        raise WireError("negative integer where an unsigned one is expected (size byte %d)" % (n - 256))
    if n > max_bytes:
        raise WireError("integer of %d bytes does not fit the %d-byte field" % (n, max_bytes))
    if at + 1 + n > len(buf):
        raise WireError("truncated: %d-byte integer at byte %d" % (n, at))
    return int.from_bytes(buf[at + 1: at + 1 + n], "little"), at + 1 + n


def _pack(series, dtype, blob, version):
    series = np.ascontiguousarray(series, dtype=dtype)
    assert series.dtype.itemsize == blob
    return FRAME_OBJECT_BASE + put_uint(version) + put_uint(len(series)) + series.tobytes()


def _unpack(buf, dtype, blob, version, what):
    buf = bytes(buf)
    if buf[:len(FRAME_OBJECT_BASE)] != FRAME_OBJECT_BASE:
        raise WireError("unexpected I3FrameObject base bytes")
    got, at = get_uint(buf, len(FRAME_OBJECT_BASE), 4)
    if got != version:
        # the reference's message (I3CLSimStep.cxx:133-134)
        raise WireError("This reader can only read I3Vector<%s> version %u, but %u was provided." % (what, version, got))
    num, at = get_uint(buf, at, 8)
    if len(buf) - at != num * blob:
        raise WireError("blob of %d bytes where %d records of %d bytes were announced" % (len(buf) - at, num, blob))
    return np.frombuffer(buf, dtype=dtype, count=num, offset=at).copy()


def pack_step_series(steps):
    """I3CLSimStepSeries -> body bytes (I3CLSimStep.cxx:136-144)."""
    return _pack(steps, STEP_DTYPE, STEP_BLOB, STEP_VERSION)


def unpack_step_series(buf):
    """Body bytes -> I3CLSimStepSeries (I3CLSimStep.cxx:119-134)."""
    return _unpack(buf, STEP_DTYPE, STEP_BLOB, STEP_VERSION, "I3CLSimStep")


def pack_photon_series(photons):
    return _pack(photons, PHOTON_DTYPE, PHOTON_BLOB, PHOTON_VERSION)


def unpack_photon_series(buf):
    return _unpack(buf, PHOTON_DTYPE, PHOTON_BLOB, PHOTON_VERSION, "I3CLSimPhoton")
