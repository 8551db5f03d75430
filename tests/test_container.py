"""PCM container front end (include/ohp_container.h), CPU side.

* header parsing against files written by an independent implementation (Python's own `wave` / `aifc` modules) and
  against hand-built headers for every decision in CodecWav::ProcessHeader (Media/Codec/Wav.cpp:225-353) and
  CodecAiffBase::ProcessHeader (Media/Codec/AiffBase.cpp:149-281, Aiff.cpp:44-52, Aifc.cpp:44-69).  The codecs
  themselves do not compile stand-alone (CodecController, Container, MimeTypeList ...), so the parser's parity is
  "restated from the source, checked against independent writers", not linked;
* message sizes (codec reads -> CodecController pieces -> DecodedAudioAggregator) against the REAL
  DecodedAudioAggregator linked into oracle/_ref, and through the schedule: host model, class-free walk and C oracle
  must produce the same playables for container-born streams."""
import io
import struct
import warnings
import wave

import numpy as np
import pytest

from ohpipeline_b200 import abi, capi, workloads

with warnings.catch_warnings():
    warnings.simplefilter("ignore", DeprecationWarning)
    import aifc


def wav_bytes(rate, bits, ch, frames, seed=0):
    rng = np.random.default_rng(seed)
    pcm = rng.integers(0, 256, frames * ch * bits // 8, dtype=np.uint8).tobytes()
    f = io.BytesIO()
    with wave.open(f, "wb") as w:
        w.setnchannels(ch); w.setsampwidth(bits // 8); w.setframerate(rate); w.writeframes(pcm)
    return f.getvalue(), pcm


class _KeepOpen(io.BytesIO):
    def close(self):  # aifc closes the file it was given
        pass


def aiff_bytes(rate, bits, ch, frames, seed=0, sowt=False):
    rng = np.random.default_rng(seed)
    pcm = rng.integers(0, 256, frames * ch * bits // 8, dtype=np.uint8).tobytes()
    f = _KeepOpen()
    w = aifc.open(f, "wb")
    if sowt:
        w.aifc()
        w.setcomptype(b"NONE", b"not compressed")
    else:
        w.aiff()
    w.setnchannels(ch); w.setsampwidth(bits // 8); w.setframerate(rate); w.writeframes(pcm)
    w.close()
    data = bytearray(f.getvalue())
    if sowt:
        i = data.index(b"NONE")
        data[i:i + 4] = b"sowt"
    return bytes(data), pcm


def chunk(cid, payload, little=True):
    return cid + struct.pack("<I" if little else ">I", len(payload)) + payload + (b"\0" if len(payload) % 2 else b"")


def riff(body, size=None):
    return b"RIFF" + struct.pack("<I", len(body) + 4 if size is None else size) + b"WAVE" + body


def fmt(tag=1, ch=2, rate=44100, bits=16, extra=b""):
    return chunk(b"fmt ", struct.pack("<HHIIHH", tag, ch, rate, rate * ch * bits // 8, ch * bits // 8, bits) + extra)


@pytest.mark.parametrize("rate,bits,ch", [(44100, 16, 2), (48000, 24, 2), (192000, 24, 2), (96000, 32, 6), (8000, 8, 1), (384000, 16, 8)])
def test_wav_written_by_the_standard_library(rate, bits, ch):
    data, pcm = wav_bytes(rate, bits, ch, 1234)
    rc, info = capi.container_parse(data)
    assert rc == abi.CONTAINER_OK
    assert (info["kind"], info["sample_rate"], info["bit_depth"], info["channels"], info["little_endian"]) == (abi.CONTAINER_WAV, rate, bits, ch, 1)
    assert int(info["total_frames"]) == 1234 and int(info["audio_bytes"]) == len(pcm)
    assert data[int(info["data_offset"]):int(info["data_offset"]) + len(pcm)] == pcm
    assert int(info["track_length_jiffies"]) == 1234 * abi.JIFFIES_PER_SECOND // rate
    assert int(info["bit_rate"]) == rate * ch * bits
    rc, spec = capi.container_stream_spec(info, len(data), arena_offset=4096, dst_base=77)
    assert rc == abi.CONTAINER_OK
    assert int(spec["src_base"]) == 4096 + int(info["data_offset"]) and int(spec["dst_base"]) == 77
    assert int(spec["chunk_frames"]) == workloads.max_chunk_frames(rate, bits, ch) and int(spec["codec_read_frames"]) == 0
    assert int(spec["in_little_endian"]) == (1 if bits > 8 else 0)


@pytest.mark.parametrize("rate,bits,ch,sowt", [(44100, 16, 2, False), (48000, 24, 2, False), (96000, 24, 6, False), (22050, 8, 1, False),
                                               (44100, 16, 2, True), (192000, 24, 2, True)])
def test_aiff_written_by_the_standard_library(rate, bits, ch, sowt):
    data, pcm = aiff_bytes(rate, bits, ch, 777, sowt=sowt)
    rc, info = capi.container_parse(data)
    assert rc == abi.CONTAINER_OK
    assert int(info["kind"]) == (abi.CONTAINER_AIFC if sowt else abi.CONTAINER_AIFF)
    assert (info["sample_rate"], info["bit_depth"], info["channels"], info["little_endian"]) == (rate, bits, ch, int(sowt))
    assert int(info["total_frames"]) == 777
    assert data[int(info["data_offset"]):int(info["data_offset"]) + len(pcm)] == pcm
    rc, spec = capi.container_stream_spec(info, len(data))
    assert rc == abi.CONTAINER_OK
    fb = ch * bits // 8
    assert int(spec["codec_read_frames"]) == 9216 // fb       # AiffBase.cpp:66
    assert int(spec["in_little_endian"]) == (1 if sowt and bits > 8 else 0)


def test_wav_header_decisions():
    pcm = bytes(range(200))
    ok = riff(fmt() + chunk(b"data", pcm))
    rc, info = capi.container_parse(ok)
    assert rc == abi.CONTAINER_OK and int(info["data_offset"]) == 44 and int(info["total_frames"]) == 50
    # a LIST chunk (odd size: one pad byte) between fmt and data is skipped (FindChunk, Wav.cpp:319-353)
    rc, info = capi.container_parse(riff(fmt() + chunk(b"LIST", b"INFOabc") + chunk(b"data", pcm)))
    assert rc == abi.CONTAINER_OK and int(info["data_offset"]) == 44 + 8 + 8
    # fmt sizes 18 and 40 (WAVE_FORMAT_EXTENSIBLE, tag 0xfffe) are accepted, others are corrupt (Wav.cpp:275-277)
    for extra, tag, want in ((b"\0\0", 1, abi.CONTAINER_OK), (b"\0" * 24, 0xfffe, abi.CONTAINER_OK), (b"\0" * 4, 1, abi.CONTAINER_E_CORRUPT)):
        assert capi.container_parse(riff(fmt(tag=tag, extra=extra) + chunk(b"data", pcm)))[0] == want
    # compressed formats are unsupported (Wav.cpp:289-292); zero channels / rate / odd depth are corrupt (:304-307)
    assert capi.container_parse(riff(fmt(tag=0x55) + chunk(b"data", pcm)))[0] == abi.CONTAINER_E_UNSUPPORTED
    assert capi.container_parse(riff(fmt(ch=0) + chunk(b"data", pcm)))[0] == abi.CONTAINER_E_CORRUPT
    assert capi.container_parse(riff(fmt(rate=0) + chunk(b"data", pcm)))[0] == abi.CONTAINER_E_CORRUPT
    assert capi.container_parse(riff(fmt(bits=12) + chunk(b"data", pcm)))[0] == abi.CONTAINER_E_CORRUPT
    # data bytes are truncated to whole frames (Wav.cpp:327-330)
    rc, info = capi.container_parse(riff(fmt() + chunk(b"data", pcm[:198])))
    assert int(info["audio_bytes"]) == 196
    # ... after FindChunk has rounded an odd chunk size up to its pad byte (Wav.cpp:330-331): 199 -> 200, like the reference
    rc, info = capi.container_parse(riff(fmt() + chunk(b"data", pcm[:199])))
    assert int(info["audio_bytes"]) == 200
    # a RIFF size of zero is a continuous stream: length unknown (Wav.cpp:258, 320-322); the spec takes what is there
    live = riff(fmt() + chunk(b"data", pcm), size=0)
    rc, info = capi.container_parse(live)
    assert rc == abi.CONTAINER_OK and int(info["streaming"]) == 1 and int(info["audio_bytes"]) == 0
    assert int(capi.container_stream_spec(info, len(live))[1]["total_frames"]) == 50
    # the animator's depth limit (Wav.cpp:299): 32-bit audio on a 24-bit pipeline re-quantises -> not a plain stream
    rc, info = capi.container_parse(riff(fmt(bits=32) + chunk(b"data", pcm)), max_bit_depth=24)
    assert rc == abi.CONTAINER_OK and (int(info["bit_depth_src"]), int(info["bit_depth"])) == (32, 24)
    assert capi.container_stream_spec(info, 1000)[0] == abi.CONTAINER_E_UNSUPPORTED
    # the stream ends inside the header
    for cut in (11, 12, 20, 30, 43):
        rc = capi.container_parse(ok[:cut])[0]
        assert rc == (abi.CONTAINER_E_UNRECOGNISED if cut < 12 else abi.CONTAINER_E_ENDED), cut
    assert capi.container_parse(b"RIFX" + ok[4:])[0] == abi.CONTAINER_E_UNRECOGNISED
    # a truncated file shortens the stream
    rc, info = capi.container_parse(ok)
    assert int(capi.container_stream_spec(info, len(ok) - 10)[1]["total_frames"]) == 47


def ext80(rate):
    """80-bit IEEE extended big-endian, as AIFF's COMM chunk stores the sample rate."""
    e = rate.bit_length() - 1
    mant = rate << (63 - e)
    return struct.pack(">HQ", 16383 + e, mant)


def form(kind, body):
    return b"FORM" + struct.pack(">I", len(body) + 4) + kind + body


def comm(ch, frames, bits, rate, comp=None):
    payload = struct.pack(">HIH", ch, frames, bits) + ext80(rate)
    if comp is not None:
        payload += comp + b"\x00\x00"
    return chunk(b"COMM", payload, little=False)


def test_aiff_header_decisions():
    pcm = bytes(range(240))
    ssnd = chunk(b"SSND", struct.pack(">II", 0, 0) + pcm, little=False)
    for rate in (8000, 22050, 44100, 48000, 88200, 96000, 192000, 384000):   # both branches of DetermineRate
        rc, info = capi.container_parse(form(b"AIFF", comm(2, 60, 16, rate) + ssnd))
        assert rc == abi.CONTAINER_OK and int(info["sample_rate"]) == rate, rate
    for mac, want in ((22255, 22050), (11127, 11025)):                         # AiffBase.cpp:176-183
        assert int(capi.container_parse(form(b"AIFF", comm(2, 60, 16, mac) + ssnd))[1]["sample_rate"]) == want
    # metadata chunks before COMM are skipped; COMM must be exactly 18 bytes in AIFF, at least 22 in AIFC
    rc, info = capi.container_parse(form(b"AIFF", chunk(b"NAME", b"abc", little=False) + comm(2, 60, 16, 44100) + ssnd))
    assert rc == abi.CONTAINER_OK and int(info["data_offset"]) == 12 + 12 + 26 + 16
    assert capi.container_parse(form(b"AIFF", comm(2, 60, 16, 44100, comp=b"NONE") + ssnd))[0] == abi.CONTAINER_E_CORRUPT
    assert capi.container_parse(form(b"AIFC", comm(2, 60, 16, 44100) + ssnd))[0] == abi.CONTAINER_E_CORRUPT
    for comp, rc_want, le in ((b"NONE", abi.CONTAINER_OK, 0), (b"sowt", abi.CONTAINER_OK, 1), (b"SOWT", abi.CONTAINER_OK, 1),
                              (b"ulaw", abi.CONTAINER_E_UNSUPPORTED, 0)):
        rc, info = capi.container_parse(form(b"AIFC", comm(2, 60, 16, 44100, comp=comp) + ssnd))
        assert rc == rc_want and (rc != abi.CONTAINER_OK or int(info["little_endian"]) == le), comp
    # depths: 8/16/24 as they are, 20 is played as 24, anything else unsupported (AiffBase.cpp:241-251)
    assert capi.container_parse(form(b"AIFF", comm(2, 30, 32, 44100) + ssnd))[0] == abi.CONTAINER_E_UNSUPPORTED
    assert int(capi.container_parse(form(b"AIFF", comm(2, 60, 20, 44100) + ssnd))[1]["bit_depth"]) == 24
    # more audio promised than the SSND chunk holds is corrupt (AiffBase.cpp:266-268; the comparison includes the
    # chunk's 8 bytes of offset / block size, so two frames too many still pass, like in the reference)
    assert capi.container_parse(form(b"AIFF", comm(2, 62, 16, 44100) + ssnd))[0] == abi.CONTAINER_OK
    assert capi.container_parse(form(b"AIFF", comm(2, 63, 16, 44100) + ssnd))[0] == abi.CONTAINER_E_CORRUPT
    assert capi.container_parse(form(b"AIFF", comm(2, 60, 16, 44100)))[0] == abi.CONTAINER_E_ENDED


@pytest.mark.parametrize("rate,bits,ch", [(44100, 16, 2), (48000, 24, 2), (96000, 24, 6), (192000, 24, 2), (8000, 8, 1),
                                          (384000, 16, 8), (44100, 24, 5), (176400, 8, 3), (32000, 16, 7)])
def test_message_sizes_match_the_real_aggregator(ref, rate, bits, ch):
    """Aiff reads of 9216 B, cut into 5 ms pieces by CodecController::OutputAudioPcm, through the REAL
    DecodedAudioAggregator -- against ohp_codec_message_frames; and Wav's uniform messages, which it must not change."""
    fb = ch * bits // 8
    jps = abi.jiffies_per_sample(rate)
    piece = workloads.max_chunk_frames(rate, bits, ch)
    assert piece == min((5 * abi.JIFFIES_PER_MS) // jps, 9216 // fb)
    for total in (1, piece - 1, piece, 9216 // fb, 3 * (9216 // fb) + 17, 50 * piece + 3):
        for read in (9216 // fb, 0):
            pieces = []
            left = total
            while left > 0:
                r = min(read, left) if read else min(piece, left)
                left -= r
                while r > 0:
                    p = min(piece, r)
                    pieces.append(p)
                    r -= p
            want = ref.aggregate(rate, ch, bits, pieces)
            assert want is not None and int(want.sum()) == total
            spec = np.zeros(1, dtype=abi.STREAM_SPEC)[0]
            spec["sample_rate"] = rate; spec["bit_depth"] = bits; spec["channels"] = ch
            spec["chunk_frames"] = piece; spec["codec_read_frames"] = read; spec["total_frames"] = total
            got = capi.codec_message_frames(spec)
            assert list(got) == list(want), (rate, bits, ch, total, read)
            if read == 0:
                assert list(got) == pieces  # 5 ms (or cell-sized) messages pass through unchanged


def test_container_born_streams_through_every_schedule_builder(port):
    """AIFF-style reads change the message boundaries a ramp is cut at: host model, class-free walk and C oracle agree."""
    rng = np.random.default_rng(9)
    specs, evs = [], []
    for n, (rate, bits, ch) in enumerate([(44100, 16, 2), (48000, 24, 2), (96000, 24, 6), (44100, 24, 5), (192000, 24, 2)]):
        data, _ = aiff_bytes(rate, bits, ch, 20000 + 37 * n, seed=n)
        rc, info = capi.container_parse(data)
        assert rc == abi.CONTAINER_OK
        rc, spec = capi.container_stream_spec(info, len(data))
        assert rc == abi.CONTAINER_OK and int(spec["codec_read_frames"]) > 0
        specs.append(spec)
        jps = abi.jiffies_per_sample(rate)
        total_j = int(spec["total_frames"]) * jps
        evs.append([(int(rng.integers(0, total_j // 2)), 0, abi.EV_RAMP_DOWN, 20 * abi.JIFFIES_PER_MS),
                    (int(rng.integers(total_j // 2, total_j)), 0, abi.EV_RAMP_UP, 50 * abi.JIFFIES_PER_MS)])
    w = workloads._finish("aiff-born", specs, evs, seed=1)
    host = capi.schedule_build(w.streams, w.events)
    walk = capi.schedule_build(w.streams, w.events, walk=True)
    rc, chunks, info, begin, outb = port.schedule_run(w.streams, w.events)
    assert rc == 0
    assert np.array_equal(host.chunks, walk.chunks) and np.array_equal(host.info, walk.info)
    assert np.array_equal(host.chunks, chunks) and np.array_equal(host.info, info)
    # and the boundaries are the aggregator's: the unramped head of each stream is cut exactly at its message sizes
    for k in range(len(w.streams)):
        msgs = capi.codec_message_frames(w.streams[k])
        fb = int(w.streams[k]["channels"]) * int(w.streams[k]["bit_depth"]) // 8
        first = host.chunks[int(host.stream_chunk_begin[k])]
        assert int(first["bytes"]) in (int(msgs[0]) * fb,) or int(first["flags"]) & abi.F_RAMP_ENABLED


@pytest.mark.gpu
def test_batch_born_from_container_bytes_on_the_gpu(ctx, port):
    """The input arena holds whole WAV / AIFF / AIFC files back to back; only headers are read on the host.  The GPU
    builds the playables (device walk, aggregator-shaped message sizes) and ramps + converts straight out of the
    containers' data chunks (odd src offsets, LE subsamples for WAV and "sowt").  Bytes vs the oracle, and -- for the
    unramped part of every stream -- vs the PCM the independent writer was given."""
    rng = np.random.default_rng(5)
    arena = bytearray()
    specs, evs, pcms = [], [], []
    dst = 0
    cases = [("wav", 44100, 16, 2), ("wav", 192000, 24, 2), ("wav", 96000, 32, 6), ("wav", 8000, 8, 1),
             ("aiff", 44100, 16, 2), ("aiff", 96000, 24, 6), ("aiff", 44100, 24, 5), ("sowt", 48000, 24, 2),
             ("sowt", 44100, 16, 2), ("wav", 384000, 16, 8), ("aiff", 22050, 8, 1), ("wav", 48000, 24, 7)]
    for n, (kind, rate, bits, ch) in enumerate(cases):
        frames = 9000 + 531 * n
        if kind == "wav":
            data, pcm = wav_bytes(rate, bits, ch, frames, seed=n)
        else:
            data, pcm = aiff_bytes(rate, bits, ch, frames, seed=n, sowt=(kind == "sowt"))
        arena += b"\xa5" * int(rng.integers(0, 7))  # files land wherever the previous one ended
        at = len(arena)
        arena += data
        rc, info = capi.container_parse(data)
        assert rc == abi.CONTAINER_OK
        rc, spec = capi.container_stream_spec(info, len(data), arena_offset=at, dst_base=dst)
        assert rc == abi.CONTAINER_OK
        assert bytes(arena[int(spec["src_base"]):int(spec["src_base"]) + len(pcm)]) == pcm
        dst += len(pcm) + int(rng.integers(0, 5))
        specs.append(spec)
        pcms.append((pcm, bits, bool(info["little_endian"])))
        total_j = frames * abi.jiffies_per_sample(rate)
        evs.append([(int(rng.integers(total_j // 4, total_j // 2)), 0, abi.EV_RAMP_DOWN, 20 * abi.JIFFIES_PER_MS),
                    (int(rng.integers(total_j // 2, 3 * total_j // 4)), 0, abi.EV_RAMP_UP, 50 * abi.JIFFIES_PER_MS)])
    w = workloads._finish("containers", specs, evs, seed=1)
    for k, spec in enumerate(specs):  # _finish lays streams out afresh; these live where their files lie
        w.streams[k]["src_base"] = spec["src_base"]
        w.streams[k]["dst_base"] = spec["dst_base"]
    inp = np.frombuffer(bytes(arena) + b"\0" * 64, dtype=np.uint8).copy()
    out_bytes = dst + 64
    rc, want, chunks, _ = port.run(w.streams, w.events, inp, out_bytes)
    assert rc == 0
    host = capi.schedule_build(w.streams, w.events)
    dev = ctx.schedule_build_device(w.streams, w.events)
    assert np.array_equal(dev.chunks, host.chunks) and np.array_equal(dev.chunks, chunks)
    got = np.zeros(out_bytes, dtype=np.uint8)
    ctx.process_host(dev.chunks, inp, got)
    from util import covered_mask
    mask = covered_mask(chunks, out_bytes)
    assert np.array_equal(got[mask], want[mask])
    # independent of oracle and schedule: the unramped head of each stream is the writer's PCM, big-endian
    for k, (pcm, bits, little) in enumerate(pcms):
        b = bits // 8
        first_ramped = next(i for i in range(int(host.stream_chunk_begin[k]), int(host.stream_chunk_begin[k + 1]))
                            if int(host.chunks[i]["flags"]) & abi.F_RAMP_ENABLED)
        n = int(host.chunks[first_ramped]["dst_off"]) - int(w.streams[k]["dst_base"])
        assert n > 0
        raw = np.frombuffer(pcm, dtype=np.uint8)[:n]
        be = raw.reshape(-1, b)[:, ::-1].reshape(-1) if (little and b > 1) else raw
        lo = int(w.streams[k]["dst_base"])
        assert np.array_equal(got[lo:lo + n], be), cases[k]


def test_parser_survives_mutated_and_truncated_headers():
    """Hostile bytes: every prefix of valid files and thousands of random mutations parse to a status, never past the
    buffer (the reader is bounds-checked; run under the default allocator this would crash on an over-read of the
    exact-size numpy buffer often enough to notice), and whatever parses OK describes audio that lies inside the file
    or is clamped by ohp_container_stream_spec."""
    rng = np.random.default_rng(11)
    seeds = [wav_bytes(44100, 16, 2, 50)[0], wav_bytes(96000, 24, 6, 20)[0], aiff_bytes(48000, 24, 2, 40)[0],
             aiff_bytes(44100, 16, 2, 40, sowt=True)[0]]
    statuses = set()
    for data in seeds:
        for cut in range(0, min(len(data), 120)):
            rc, _ = capi.container_parse(data[:cut])
            statuses.add(rc)
            assert rc != abi.CONTAINER_OK or cut >= 44
        for _ in range(1500):
            b = bytearray(data[:int(rng.integers(12, len(data) + 1))])
            for _ in range(int(rng.integers(1, 6))):
                i = int(rng.integers(0, min(len(b), 64)))
                b[i] = int(rng.integers(0, 256))
            rc, info = capi.container_parse(bytes(b))
            statuses.add(rc)
            assert 0 <= rc <= abi.CONTAINER_E_ARG
            if rc == abi.CONTAINER_OK:
                rc2, spec = capi.container_stream_spec(info, len(b))
                if rc2 == abi.CONTAINER_OK:
                    fb = int(spec["channels"]) * int(spec["bit_depth"]) // 8
                    assert int(spec["src_base"]) + int(spec["total_frames"]) * fb <= len(b)
    assert {abi.CONTAINER_OK, abi.CONTAINER_E_UNRECOGNISED, abi.CONTAINER_E_ENDED, abi.CONTAINER_E_CORRUPT} <= statuses
