#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libohref.so = the reference's own Msg.cpp
compiled unmodified).  Run in the build container, where /root/reference is mounted:

    python tests/golden/make_golden.py

Each fixture holds a small workload (stream specs, ramp events, PCM seed) and what the reference produced for it:
the playable descriptors (MsgPlayable state) and the bytes Read(ProcessorPcmBufTest) delivered.  The fixtures are
committed; the GPU box (which has no /root/reference) checks the oracle port and the CUDA path against them.
"""
import hashlib
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from ohpipeline_b200 import abi, workloads as W  # noqa: E402
from oracle import pyoracle  # noqa: E402


def cases():
    yield "config1_2s", W.config1(2.0)
    c1 = W.config1(6.2)                       # the 20 ms ramp down at 3 s, muted, 20 ms ramp up at 5 s
    yield "config1_ramps", c1
    yield "config2_x2", W.config2(n_streams=2, seconds=0.05)
    yield "config3_x3", W.config3(n_streams=3, seconds=0.4)
    yield "config5_x2", W.config5(n_streams=2, seconds=0.05)
    c4 = W.config4(n_streams=24, seconds=0.1, seed=4)
    c4.streams["out_fmt"] = abi.OUT_PACKED_BE     # the linked reference only has the packed-BE sink
    yield "config4_x24", c4
    yield "all_rates", W.all_rates(seconds=0.2)               # 18 rates x 4 depths through a starvation
    yield "steady_edges_3", W.steady_edges(3, n_streams=48)   # the schedule walk's bulk step: events on / next to message boundaries
    for seed in (101, 102, 103, 104):
        w = W.mixed(n_streams=12, seed=seed, max_frames=1500)
        w.streams["out_fmt"] = abi.OUT_PACKED_BE  # the linked reference only has the packed-BE sink
        yield "mixed_%d" % seed, w
    # recorded from the reference's ELEMENT OBJECTS (Ramper, StarvationRamper, Muter; oracle/ref_elements.cpp), not from the
    # stage model on the reference's message classes like the ones above
    yield "elements_11", W.elements(11, n_streams=20, seconds=0.4)


def main():
    ref = pyoracle.Ref()
    port = pyoracle.Port()
    total = 0
    only = sys.argv[1:]
    for name, w in cases():
        if only and name not in only:
            continue
        inp = port.fill_pcm(w.in_bytes, w.seed)
        if name.startswith("elements"):
            rc, out, chunks, info, _, _, _ = ref.elements_run(w.streams, w.events, inp, w.out_bytes)
        else:
            rc, out, chunks, info = ref.run(w.streams, w.events, inp, w.out_bytes, threads=1)
        assert rc == 0, (name, rc)
        path = os.path.join(HERE, name + ".npz")
        # every chunk's bytes are pinned by a CRC-32; the bytes themselves are kept for small outputs only
        ob = abi.chunk_out_bytes(chunks)
        crc = np.array([zlib.crc32(out[int(d["dst_off"]):int(d["dst_off"]) + int(n)].tobytes()) for d, n in zip(chunks, ob)],
                       dtype=np.uint32)
        extra = {"out": out} if w.out_bytes <= 131072 else {}
        np.savez_compressed(path, streams=w.streams, events=w.events, in_bytes=w.in_bytes, out_bytes=w.out_bytes,
                            seed=np.uint64(w.seed), chunks=chunks, info=info, chunk_crc=crc,
                            out_sha256=np.frombuffer(hashlib.sha256(out.tobytes()).digest(), dtype=np.uint8), **extra)
        total += os.path.getsize(path)
        print("%-16s %6d chunks %8d bytes out -> %s (%d B)" % (name, len(chunks), w.out_bytes, os.path.basename(path),
                                                                os.path.getsize(path)))
    if only:
        return
    # function-level vectors: Ramp::Set / Ramp::Split on random and edge arguments
    rng = np.random.default_rng(7)
    K = abi.RAMP_MAX
    set_rows, split_rows = [], []
    edge = [0, 1, 2, 31, 32, 47, 48, K // 4, K // 2, 3 * K // 4, K - 1, K]
    for i in range(3000):
        if i < 1000:
            cs, ce = int(rng.choice(edge)), int(rng.choice(edge))
            start = int(rng.choice(edge))
        else:
            cs, ce, start = (int(x) for x in rng.integers(0, K + 1, 3))
        enabled = int(rng.integers(0, 2))
        if not enabled:
            cur = (K, K, abi.DIR_NONE, 0)
        elif rng.random() < 0.1:
            cur = (0, 0, abi.DIR_MUTE, 1)
        else:
            cur = (cs, ce, abi.DIR_NONE if cs == ce else (abi.DIR_UP if cs < ce else abi.DIR_DOWN), 1)
        frag = int(rng.integers(1, 300000))
        dur = frag + int(rng.integers(0, 3000000)) if rng.random() < 0.9 else frag
        direction = int(rng.choice((abi.DIR_UP, abi.DIR_DOWN)))
        rc, r, s, pos = ref.ramp_set(cur, start, frag, dur, direction)
        set_rows.append(cur + (start, frag, dur, direction, rc) + r + s + (pos,))
        if cur[3]:
            cursize = int(rng.integers(2, 300000))
            new = int(rng.integers(1, cursize))
            rc2, a, b = ref.ramp_split(cur, new, cursize)
            split_rows.append(cur + (new, cursize, rc2) + a + b)
    path = os.path.join(HERE, "ramp_algebra.npz")
    np.savez_compressed(path, set=np.array(set_rows, dtype=np.int64), split=np.array(split_rows, dtype=np.int64))
    total += os.path.getsize(path)
    print("ramp_algebra: %d Set + %d Split vectors (%d B)" % (len(set_rows), len(split_rows), os.path.getsize(path)))
    print("total %.1f KB" % (total / 1024.0))


if __name__ == "__main__":
    main()
