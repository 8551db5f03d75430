#!/usr/bin/env python
"""Generate tests/golden/starvation_flywheel.npz from the REFERENCE ITSELF: the streams of flywheel_util.starved_streams()
pulled through the real StarvationRamper element (oracle/ref_elements.cpp, ref_elements_generated_audio), starved at the
positions their OHP_EV_STARVATION events name.  Per stream: the bytes a driver reads from the flywheel messages the element
plays (every starvation, one after the other) and the ramp value the element had when each began.  Run in the build
container:

    python tests/golden/make_golden_starvation.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyoracle  # noqa: E402
from flywheel_util import starved_streams  # noqa: E402


def main():
    ref, port = pyoracle.Ref(), pyoracle.Port()
    out = {}
    for name, w, seed in starved_streams():
        inp = port.fill_pcm(w.in_bytes, seed)
        rc, audio, ramps = ref.elements_generated_audio(w.streams, w.events, inp)
        assert rc == 0 and len(ramps) >= 1, (name, rc)
        out["audio_" + name] = audio
        out["ramps_" + name] = ramps
    path = os.path.join(HERE, "starvation_flywheel.npz")
    np.savez_compressed(path, **out)
    print("%d streams -> %s (%d B)" % (len(out) // 2, path, os.path.getsize(path)))


if __name__ == "__main__":
    main()
