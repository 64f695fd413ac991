"""Writes tests/golden/wire_step_series.hex: the body of a two-step I3CLSimStepSeries as the format description in
private/clsim/I3CLSimStep.cxx:96-144 + the portable archive's integer rule gives it, assembled here BY HAND (struct.pack,
not clsim_b200.wire), so that the codec is checked against an independent reading of the same lines.
usage: python tests/golden/make_wire_fixture.py"""
import os
import struct

steps = [
    # x, y, z, t, theta, phi, length, beta, num, weight, id, sourceType, dummy1, dummy2
    (1.0, -2.0, 3.5, 10.0, 0.5, 1.5, 0.25, 1.0, 200, 1.0, 7, 0, 0, 0),
    (-100.0, 50.0, -400.0, 2500.0, 2.0, 4.0, 1.5, 0.99, 65536 + 3, 0.5, 0xDEADBEEF, 1, 0, 0),
]
body = b"\x00\x00"                      # I3FrameObject base: tracking flag, class version (see clsim_b200/wire.py)
body += b"\x00"                          # I3CLSimStep_version = 0: the single byte 0
body += b"\x01\x02"                      # num = 2: one byte of value
for s in steps:
    body += struct.pack("<8fIfIBBH", *s)
assert len(body) == 5 + 2 * 48
here = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(here, "wire_step_series.hex"), "w") as f:
    f.write(body.hex() + "\n")
# a series long enough for a two-byte count: 300 identical dummy steps -> size byte 2, then 0x2c 0x01
with open(os.path.join(here, "wire_step_series_300_header.hex"), "w") as f:
    f.write((b"\x00\x00" + b"\x00" + b"\x02\x2c\x01").hex() + "\n")
