"""Device schedule build for Aiff-born streams (block-reading codec + DecodedAudioAggregator: ragged message sizes)
next to the same streams Wav-born (uniform 5 ms messages).  python profiles/sweeps/aiff_schedule.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ohpipeline_b200 import capi, workloads

ctx = capi.Context(0)
for rate, bits in ((44100, 16), (192000, 24), (96000, 24)):
    w = workloads.config2(n_streams=1024, seconds=10.0)
    w.streams["sample_rate"] = rate; w.streams["bit_depth"] = bits
    fb = 2 * bits // 8
    w.streams["chunk_frames"] = workloads.max_chunk_frames(rate, bits, 2)
    w.streams["total_frames"] = rate * 10
    for mode in (0, 9216 // fb):
        w.streams["codec_read_frames"] = mode
        workloads.layout(w.streams)
        host = capi.schedule_build(w.streams, w.events)
        dev = ctx.schedule_build_device(w.streams, w.events)
        d_specs = torch.from_numpy(w.streams.view(np.uint8).copy()).cuda()
        d_events = torch.from_numpy(w.events.view(np.uint8).copy()).cuda()
        d_begin = torch.zeros(len(w.streams) + 1, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        tc = []
        for _ in range(4):
            t0 = time.perf_counter()
            ctx.schedule_count_device(d_specs.data_ptr(), len(w.streams), d_events.data_ptr(), len(w.events), d_begin.data_ptr())
            tc.append(time.perf_counter() - t0)
        print("%6d Hz %2d bit %-4s chunks %8d identical to host model %s count pass %.2f ms"
              % (rate, bits, "aiff" if mode else "wav", len(host.chunks), bool(np.array_equal(dev.chunks, host.chunks)), min(tc) * 1e3))
