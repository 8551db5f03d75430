"""The applicator checks of the reference's own SuiteRamp (Media/Tests/TestMsg.cpp:1445-1591), restated: 792 bytes of
0x7f (and of 0xff) as stereo 8/16/24/32-bit audio under ramps [Max..Min], [Min..Max], [Max..50 %], [Min..50 %] and
[50 %..25 %].  The reference only pins tolerances here -- first value close to the start, never rising (falling),
channels equal, last value 0 / within 2 of (0x7f * kRampArray[k]) >> 15 -- so these are secondary invariants next to
the bit-exact parity tests; they run against the oracle port on the CPU and against the CUDA path on the GPU."""
import numpy as np
import pytest

from ohpipeline_b200 import abi
from util import make_desc

K = abi.RAMP_MAX
N = 792


def read(process, fill, bits, start, end):
    d = make_desc(bytes=N, bit_depth=bits, channels=2, flags=abi.F_RAMP_ENABLED, ramp_start=start, ramp_end=end)
    inp = np.full(N + 64, fill, dtype=np.uint8)
    out = process(d, inp, N + 64)
    b = bits // 8
    frames = out[:N].reshape(-1, 2, b).astype(np.uint32)
    vals = np.zeros(frames.shape[:2], dtype=np.uint64)
    for k in range(b):
        vals = (vals << np.uint64(8)) | frames[:, :, k].astype(np.uint64)
    return vals  # [frame, channel]


def check_suite(process, ramp_array):
    mid = lambda k: (0x7f * int(ramp_array[k])) >> 15
    # [Max..Min] on 0x7f, 8-bit: starts close to the sample value, never rises, both channels equal, ends at 0
    v = read(process, 0x7f, 8, K, 0)
    assert v[0, 0] >= 0x7d and (v[:, 0] == v[:, 1]).all() and (np.diff(v[:, 0].astype(np.int64)) <= 0).all() and v[-1, 0] == 0
    # ... on 0xff (negative): stays <= 0, never rises as an unsigned byte, ends at 0
    v = read(process, 0xff, 8, K, 0)
    assert v[0, 0] >= 0xfd and (v[:, 0] == v[:, 1]).all() and (np.diff(v[:, 0].astype(np.int64)) <= 0).all() and v[-1, 0] == 0
    assert (((v[:, 0] & 0x80) != 0) | (v[:, 0] == 0)).all()
    # 16/24/32-bit: channels equal, never rises
    for bits in (16, 24, 32):
        v = read(process, 0x7f, bits, K, 0)
        assert (v[:, 0] == v[:, 1]).all() and (np.diff(v[:, 0].astype(np.int64)) <= 0).all()
    # [Min..Max]: starts close to zero, never falls, ends close to the sample value
    v = read(process, 0x7f, 8, 0, K)
    assert v[0, 0] <= 0x02 and (v[:, 0] == v[:, 1]).all() and (np.diff(v[:, 0].astype(np.int64)) >= 0).all() and v[-1, 0] >= 0x7d
    # [Max..50 %], [Min..50 %], [50 %..25 %]: end (start) values within 2 of the curve's
    v = read(process, 0x7f, 8, K, K // 2)
    assert v[0, 0] >= 0x7d and mid(256) - int(v[-1, 0]) <= 2
    v = read(process, 0x7f, 8, 0, K // 2)
    assert v[0, 0] <= 0x02 and mid(256) - int(v[-1, 0]) <= 2
    v = read(process, 0x7f, 8, K // 2, K // 4)
    assert mid(256) - int(v[0, 0]) < 2 and mid(384) - int(v[-1, 0]) <= 2
    # the one exact constant of the reference's tests: 0x7f7f7f at full level reads 0x7f7e00 (TestMuter.cpp:332)
    v = read(process, 0x7f, 24, K, K)
    assert (v == 0x7f7e00).all()


def test_suite_ramp_applicator_properties_on_the_oracle(port):
    def process(d, inp, out_bytes):
        rc, out = port.process_chunks(d, inp, out_bytes)
        assert rc == 0
        return out
    check_suite(process, port.ramp_array)


@pytest.mark.gpu
def test_suite_ramp_applicator_properties_on_the_gpu(ctx, port):
    def process(d, inp, out_bytes):
        out = np.zeros(out_bytes, dtype=np.uint8)
        ctx.process_host(d, inp, out)
        return out
    check_suite(process, port.ramp_array)
