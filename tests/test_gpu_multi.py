"""ohp_multi (include/ohp_multi.h): the whole stage over several devices from one process -- a contiguous block of streams, a
context, a host thread and its own CUDA streams per device, no collective; per-stream checksums gathered on the host
(SURVEY 8e).  On a box with one GPU the "devices" are several contexts on that GPU: the sharding, the re-basing of every
block's arenas and events, the gather and the error paths are the same code."""
import numpy as np
import pytest

from ohpipeline_b200 import abi, capi, workloads

pytestmark = pytest.mark.gpu


def device_lists():
    n = capi.device_count()
    lists = [[0], [0, 0], [0, 0, 0]]
    if n > 1:
        lists += [list(range(n)), list(range(n)) * 2]
    return lists


def stream_checksums(port, out, streams, outb):
    return np.array([port.checksum(out[int(s["dst_base"]):int(s["dst_base"]) + int(n)]) for s, n in zip(streams, outb)], dtype=np.uint64)


@pytest.mark.parametrize("make", [
    lambda: workloads.config4(n_streams=90, seconds=0.12, seed=19),
    lambda: workloads.mixed(n_streams=67, seed=41, max_frames=4000),
    lambda: workloads.elements(5, n_streams=24),
    lambda: workloads.config5(n_streams=701, seconds=0.25),      # every block several slices
    lambda: workloads.config1(seconds=0.5),                       # one stream: all blocks but one are empty
], ids=["config4", "mixed", "elements", "config5_sliced", "one_stream"])
def test_blocks_of_streams_over_devices_give_the_bytes_one_device_gives(ctx, port, make):
    w = make()
    inp = port.fill_pcm(w.in_bytes, w.seed)
    want = np.full(w.out_bytes, 0xA5, dtype=np.uint8)             # gaps between streams stay as the caller left them
    want_outb, want_total = ctx.run_streams_host(w.streams, w.events, inp, want)
    want_sums = stream_checksums(port, want, w.streams, want_outb)
    for devices in device_lists():
        m = capi.MultiContext(devices)
        try:
            for _ in range(2):                                     # the second call reuses every device's buffers
                got = np.full(w.out_bytes, 0xA5, dtype=np.uint8)
                outb, sums, total = m.run_streams_host(w.streams, w.events, inp, got)
                assert total == want_total and np.array_equal(outb, want_outb), devices
                assert np.array_equal(got, want), devices
                assert np.array_equal(sums, want_sums), devices
        finally:
            m.close()


def test_streams_may_lie_anywhere_in_the_arenas(ctx, port):
    """Blocks are blocks of the specs ARRAY; where their streams sit in the arenas is the caller's business (here: reversed,
    so that block 0's bytes are the arenas' last)."""
    w = workloads.mixed(n_streams=48, seed=43, max_frames=3000)
    inp = port.fill_pcm(w.in_bytes, w.seed)
    streams = w.streams[::-1].copy()
    want = np.zeros(w.out_bytes, dtype=np.uint8)
    want_outb, _ = ctx.run_streams_host(streams, w.events, inp, want)
    m = capi.MultiContext([0, 0, 0])
    try:
        got = np.zeros(w.out_bytes, dtype=np.uint8)
        outb, sums, _ = m.run_streams_host(streams, w.events, inp, got)
        assert np.array_equal(got, want) and np.array_equal(outb, want_outb)
        assert np.array_equal(sums, stream_checksums(port, want, streams, want_outb))
    finally:
        m.close()


def test_pinned_buffers_and_errors(ctx, port):
    w = workloads.config5(n_streams=12, seconds=0.05)
    m = capi.MultiContext([0, 0])
    try:
        h_in, p_in = m.host_alloc(w.in_bytes)
        h_out, p_out = m.host_alloc(w.out_bytes)
        h_in[:] = port.fill_pcm(w.in_bytes, w.seed)
        h_out[:] = 0
        outb, sums, total = m.run_streams_host(w.streams, w.events, h_in, h_out)
        rc, want, chunks, _ = port.run(w.streams, w.events, h_in, w.out_bytes)
        assert rc == 0 and total == len(chunks) and np.array_equal(h_out, want)
        # nothing to do
        outb, sums, total = m.run_streams_host(w.streams[:0], w.events[:0], h_in, h_out)
        assert total == 0 and len(outb) == 0
        # a spec the message model cannot represent, in the second device's block: its status, its index, the stream's number
        bad = w.streams.copy()
        bad["sample_rate"][9] = 12345
        with pytest.raises(capi.OhpError) as e:
            m.run_streams_host(bad, w.events, h_in, h_out)
        assert e.value.status == abi.E_INVALID_ARG and "device index 1" in str(e.value) and "stream 3" in str(e.value) \
            and "counted from 6" in str(e.value)
        # a stream outside the arenas
        with pytest.raises(capi.OhpError) as e:
            m.run_streams_host(w.streams, w.events, h_in[:w.in_bytes // 2], h_out)
        assert e.value.status == abi.E_OUT_OF_RANGE and "device index 1" in str(e.value)
        # ... and everything still works
        h_out[:] = 0
        m.run_streams_host(w.streams, w.events, h_in, h_out)
        assert np.array_equal(h_out, want)
        m.host_free(p_in)
        m.host_free(p_out)
    finally:
        m.close()
