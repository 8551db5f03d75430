#!/usr/bin/env python
"""Generate tests/golden/flywheel.npz from the REFERENCE ITSELF: oracle/_ref/libohref.so links the reference's own
FlywheelRamper.cpp and StarvationRamper.cpp, and ref_flywheel() drives the real RampGenerator (Start -> its
FlywheelRamperManager::Ramp thread -> ProcessFragment / EndBlock -> TryGetAudio).  Run in the build container:

    python tests/golden/make_golden_flywheel.py

Per case: the job, the training block (FlywheelInput's layout), the generated audio, the per-block ramp descriptors,
the final ramp value, and the bytes Read(ProcessorPcmBufTest) delivers for the ramped messages."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from ohpipeline_b200 import abi, capi  # noqa: E402
from oracle import pyoracle  # noqa: E402
from flywheel_util import training_block  # noqa: E402

CASES = [(44100, 2, 16, "tone", abi.RAMP_MAX), (48000, 2, 24, "tone", 9000), (192000, 2, 24, "noise", abi.RAMP_MAX),
         (96000, 6, 32, "tone", abi.RAMP_MAX), (384000, 2, 8, "step", 1), (176400, 1, 16, "max", abi.RAMP_MAX),
         (88200, 8, 24, "noise", 0), (48000, 1, 32, "dc", 12345)]


def main():
    ref = pyoracle.Ref()
    out = {"jobs": np.concatenate([capi.flywheel_job(r, c, b) for r, c, b, _, _ in CASES]),
           "starts": np.array([s for *_, s in CASES], dtype=np.uint32), "finals": np.zeros(len(CASES), dtype=np.uint32)}
    for k, (rate, ch, bits, kind, start) in enumerate(CASES):
        training = training_block(rate, ch, kind, seed=900 + k)
        rc, raw, ramped, descs, info, final = ref.flywheel(rate, ch, bits, start, training)
        assert rc == 0, (k, rc)
        out["training_%d" % k] = training
        out["raw_%d" % k] = raw
        out["ramped_%d" % k] = ramped
        out["descs_%d" % k] = descs
        out["info_%d" % k] = info
        out["finals"][k] = final
    path = os.path.join(HERE, "flywheel.npz")
    np.savez_compressed(path, **out)
    print("%d cases -> %s (%d B)" % (len(CASES), path, os.path.getsize(path)))


if __name__ == "__main__":
    main()
