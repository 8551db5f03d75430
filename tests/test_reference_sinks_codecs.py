"""The oracle port and the host-side container parser against reference code that round 1 could not link:

* ProcessorPcmSwpEndianPacked (Media/Tests/TestCodecInteractiveMain.cpp:114-124, 540-593) and the Songcast Sender's
  IPcmProcessor half (Av/Songcast/Sender.cpp:350-398): their text is cut out of /root/reference at build time
  (oracle/extract_ref.py) and compiled into oracle/_ref -- real MsgPlayable::Read calls into the real sinks;
* CodecWav / CodecAiff / CodecAifc (Media/Codec/Wav.cpp, AiffBase.cpp, Aiff.cpp, Aifc.cpp) compiled UNMODIFIED, fed the
  container bytes through a fake CodecController: every header decision, every truncation, and a few thousand random
  mutations of valid headers must come out of ohp_container_parse the way they come out of the codecs.

CPU only; skipped where oracle/_ref did not travel.  The GPU suites check the kernel against the port on the same sinks."""
import numpy as np
import pytest

from ohpipeline_b200 import abi, capi
from util import make_desc, pack_chunks
from test_container import aiff_bytes, chunk, comm, fmt, form, riff, wav_bytes

import struct


# ----------------------------------------------------------------------------------------------------------------
# sinks

def _chunks(rng, n, depths, fmt_of, aux_of, channels=(1, 9)):
    specs = []
    for k in range(n):
        bits = int(rng.choice(depths))
        ch = int(rng.integers(channels[0], channels[1]))
        fb = ch * bits // 8
        frames = int(rng.choice((1, 2, 3, 256 // fb, 256 // fb + 1, 2 * (256 // fb), int(rng.integers(1, 9216 // fb + 1)))))
        ramped = k % 3 != 0
        specs.append(dict(bytes=frames * fb, bit_depth=bits, channels=ch,
                          flags=(abi.F_RAMP_ENABLED if ramped else 0) | (abi.F_IN_LITTLE_ENDIAN if k % 5 == 0 else 0),
                          ramp_start=int(rng.integers(0, 16385)), ramp_end=int(rng.integers(0, 16385)),
                          attenuation=int(rng.choice((256, 256, 100, 511))) if bits == 16 else 256,
                          out_fmt=fmt_of(k), aux=aux_of(k, ch), src_pad=int(rng.integers(0, 4)), dst_pad=int(rng.integers(0, 4))))
    return pack_chunks(specs)


def test_packed_le_sink_to_the_letter(ref, port):
    """aux = 0: the chunk writes what ProcessorPcmSwpEndianPacked is left holding -- everything for 8-bit and for unramped
    audio, the last <= 256-byte fragment of a ramped 16/24-bit playable (its SwapEndianness16/24 overwrite)."""
    rng = np.random.default_rng(20)
    descs, in_bytes, out_bytes = _chunks(rng, 400, (8, 16, 24), lambda k: abi.OUT_PACKED_LE, lambda k, ch: 0)
    inp = port.fill_pcm(in_bytes, 5)
    rc, want, sizes = ref.process_chunks_sinks(descs, inp, out_bytes)
    assert rc == 0
    assert np.array_equal(sizes, abi.chunk_out_bytes(descs)), "bytes the sink holds after the read"
    assert np.array_equal(sizes, np.array([port.lib.ohpo_chunk_out_bytes(descs[k:k + 1].ctypes.data) for k in range(len(descs))]))
    ramped16 = ((descs["flags"] & abi.F_RAMP_ENABLED) != 0) & (descs["bit_depth"] > 8) & (descs["bytes"] > 256)
    assert ramped16.any() and (sizes[ramped16] < descs["bytes"][ramped16]).all() and (sizes[~ramped16] == descs["bytes"][~ramped16]).all()
    rc, got = port.process_chunks(descs, inp, out_bytes)
    assert rc == 0
    assert np.array_equal(got, want)


def test_packed_le_append_mode_is_the_byte_swapped_big_endian_read(ref, port):
    """aux = OHP_LE_APPEND: every fragment in order = the linked ProcessorPcmBufTest's bytes with each subsample reversed."""
    rng = np.random.default_rng(21)
    descs, in_bytes, out_bytes = _chunks(rng, 300, (8, 16, 24), lambda k: abi.OUT_PACKED_BE, lambda k, ch: 0)
    inp = port.fill_pcm(in_bytes, 6)
    rc, be, sizes = ref.process_chunks_sinks(descs, inp, out_bytes)
    assert rc == 0 and np.array_equal(sizes, descs["bytes"])
    le = descs.copy()
    le["out_fmt"] = abi.OUT_PACKED_LE
    le["aux"] = abi.LE_APPEND
    rc, got = port.process_chunks(le, inp, out_bytes)
    assert rc == 0
    for d in descs:
        b = int(d["bit_depth"]) // 8
        lo, n = int(d["dst_off"]), int(d["bytes"])
        assert np.array_equal(got[lo:lo + n].reshape(-1, b)[:, ::-1], be[lo:lo + n].reshape(-1, b))


def test_packed_le_sink_asserts_like_the_reference(ref, port):
    inp = np.zeros(64, dtype=np.uint8)
    for d in (make_desc(bytes=32, bit_depth=32, channels=2, out_fmt=abi.OUT_PACKED_LE),
              make_desc(bytes=24, bit_depth=24, channels=2, out_fmt=abi.OUT_PACKED_LE, flags=abi.F_SILENCE)):
        assert ref.process_chunks_sinks(d, inp, 64)[0] == -1      # ASSERTS(), TestCodecInteractiveMain.cpp:561-568
        assert port.process_chunks(d, inp, 64)[0] == -1
        assert capi.validate(d, 64, 64)[0] == abi.E_INVALID_DESC


def test_songcast_sink_matches_the_reference_sender(ref, port):
    """Sender::DoProcessFragment (Av/Songcast/Sender.cpp:356-377): two channels from FirstChannelToSend (0, or 8 from ten
    channels up), at most three bytes per subsample, mono read as its own two 'channels' -- ramped, unramped and silence."""
    rng = np.random.default_rng(22)
    specs = []
    for k in range(400):
        bits = int(rng.choice((8, 16, 24, 32)))
        ch = int(rng.choice((1, 2, 2, 3, 6, 8, 10, 12)))
        fb = ch * bits // 8
        frames = int(rng.integers(1, min(9216 // fb, 700) + 1))
        kind = k % 4
        specs.append(dict(bytes=frames * fb, bit_depth=bits, channels=ch,
                          flags=(abi.F_RAMP_ENABLED if kind in (1, 2) else 0) | (abi.F_SILENCE if kind == 3 else 0)
                                | (abi.F_IN_LITTLE_ENDIAN if k % 7 == 0 and kind != 3 else 0),
                          ramp_start=int(rng.integers(0, 16385)), ramp_end=int(rng.integers(0, 16385)),
                          out_fmt=abi.OUT_SONGCAST, aux=0 if ch < 10 else 8, src_pad=int(rng.integers(0, 4)), dst_pad=int(rng.integers(0, 4))))
    descs, in_bytes, out_bytes = pack_chunks(specs)
    inp = port.fill_pcm(in_bytes, 7)
    rc, want, sizes = ref.process_chunks_sinks(descs, inp, out_bytes)
    assert rc == 0
    assert np.array_equal(sizes, abi.chunk_out_bytes(descs))
    rc, got = port.process_chunks(descs, inp, out_bytes)
    assert rc == 0
    for k, d in enumerate(descs):
        lo, n = int(d["dst_off"]), int(sizes[k])
        assert np.array_equal(got[lo:lo + n], want[lo:lo + n]), (k, d)


# ----------------------------------------------------------------------------------------------------------------
# containers

FIELDS = ("kind", "sample_rate", "bit_depth", "channels", "little_endian", "bit_rate", "data_offset", "track_length_jiffies", "total_frames")


def same_verdict(ref, data, max_bit_depth=32):
    rc_ref, want, reads = ref.container_decode(data, max_bit_depth)
    rc, got = capi.container_parse(data, max_bit_depth)
    if rc_ref == 100:
        # the reference's codec divides by zero on this header (a crash, not a verdict): the parser must not accept it
        assert rc in (abi.CONTAINER_E_CORRUPT, abi.CONTAINER_E_UNSUPPORTED, abi.CONTAINER_E_ENDED), (rc, bytes(data[:64]))
        return rc
    assert rc == rc_ref, "status %d, the reference's codecs say %d for %r..." % (rc, rc_ref, bytes(data[:64]))
    if rc == abi.CONTAINER_OK:
        if int(got["streaming"]):
            # a continuous WAV stream: the codec reports no length; the parser leaves the sizes to the caller
            assert int(want["track_length_jiffies"]) == 0 and int(got["track_length_jiffies"]) == 0
            fields = FIELDS[:7]
        else:
            fields = FIELDS
        if int(got["bit_depth"]) != int(got["bit_depth_src"]):
            # 20-bit AIFF: the codec sizes frames by 20 / 8 = 2 bytes but plays them as 24-bit audio (AiffBase.cpp:233-251), so
            # the pipeline sees two thirds of the frames COMM names; WAV above the animator's depth is re-quantised.  Neither is
            # a plain stream: ohp_container_stream_spec refuses both.
            assert capi.container_stream_spec(got, len(data))[0] == abi.CONTAINER_E_UNSUPPORTED
            fields = tuple(f for f in fields if f != "total_frames")
        if abi.jiffies_per_sample(int(got["sample_rate"])) == 0:
            # a rate the pipeline does not play: the codec's first message throws SampleRateInvalid, so the harness has no
            # audio to count frames from (the stream spec refuses such a rate: ohp_container_stream_spec)
            fields = tuple(f for f in fields if f != "total_frames")
        if int(want["kind"]) != abi.CONTAINER_WAV and "total_frames" in fields and int(got["channels"]) != 0:
            # the harness counts the frames the codec DECODED: a file that ends before the audio COMM promises plays what is there
            # (CodecAiffBase::Process throws CodecStreamEnded after its last, short read) -- what ohp_container_stream_spec takes
            fields = tuple(f for f in fields if f != "total_frames")
            fb = int(got["channels"]) * int(got["bit_depth"]) // 8
            present = max(0, len(data) - int(got["data_offset"])) // fb
            assert int(want["total_frames"]) == min(int(got["total_frames"]), present)
            rc_spec, spec = capi.container_stream_spec(got, len(data))
            assert rc_spec == abi.CONTAINER_OK and int(spec["total_frames"]) == int(want["total_frames"])
        if int(want["kind"]) != abi.CONTAINER_WAV and int(want["audio_bytes"]) == 0:
            # the byte order is only observable on the audio the codec hands on (OutputAudioPcm's aEndian)
            fields = tuple(f for f in fields if f != "little_endian")
        for f in fields:
            assert int(got[f]) == int(want[f]), (f, int(got[f]), int(want[f]), bytes(data[:64]))
        if (int(want["kind"]) != abi.CONTAINER_WAV and abi.jiffies_per_sample(int(got["sample_rate"])) != 0
                and int(got["bit_depth"]) == int(got["bit_depth_src"]) and int(got["channels"]) != 0):
            # what the codec actually handed to the pipeline: as many bytes as the file holds of the promised audio, in reads of
            # 9216 bytes rounded down to frames (AiffBase.cpp:66-76)
            fb = int(got["channels"]) * int(got["bit_depth"]) // 8
            present = max(0, min(int(got["audio_bytes"]), len(data) - int(got["data_offset"])))
            assert int(want["audio_bytes"]) == present - present % fb or int(want["audio_bytes"]) == present
            if len(reads) > 1:
                assert (reads[:-1] == 9216 - 9216 % fb).all()
    return rc


def corpus():
    out = []
    for rate, bits, ch in ((44100, 16, 2), (48000, 24, 2), (192000, 24, 2), (96000, 32, 6), (8000, 8, 1), (384000, 16, 8), (44100, 24, 5)):
        out.append(wav_bytes(rate, bits, ch, 1234)[0])
        if bits <= 24:
            out.append(aiff_bytes(rate, bits, ch, 777)[0])
            out.append(aiff_bytes(rate, bits, ch, 777, sowt=True)[0])
    pcm = bytes(range(200))
    out += [riff(fmt() + chunk(b"data", pcm)),
            riff(fmt() + chunk(b"LIST", b"INFOabc") + chunk(b"data", pcm)),
            riff(fmt(extra=b"\0\0") + chunk(b"data", pcm)), riff(fmt(tag=0xfffe, extra=b"\0" * 24) + chunk(b"data", pcm)),
            riff(fmt(extra=b"\0" * 4) + chunk(b"data", pcm)), riff(fmt(tag=0x55) + chunk(b"data", pcm)),
            riff(fmt(ch=0) + chunk(b"data", pcm)), riff(fmt(rate=0) + chunk(b"data", pcm)), riff(fmt(bits=12) + chunk(b"data", pcm)),
            riff(fmt(bits=0) + chunk(b"data", pcm)), riff(fmt(bits=4) + chunk(b"data", pcm)),
            riff(fmt() + chunk(b"data", pcm[:198])), riff(fmt() + chunk(b"data", pcm[:199])),
            riff(fmt() + chunk(b"data", pcm), size=0), riff(chunk(b"data", pcm) + fmt()), riff(fmt() + fmt(ch=6) + chunk(b"data", pcm)),
            riff(fmt(bits=32) + chunk(b"data", pcm)), b"RIFX" + riff(fmt() + chunk(b"data", pcm))[4:], b"", b"RIFF"]
    pcm = bytes(range(240))
    ssnd = chunk(b"SSND", struct.pack(">II", 0, 0) + pcm, little=False)
    for rate in (8000, 11127, 22050, 22255, 44100, 48000, 88200, 96000, 192000, 384000, 7350, 12345, 1, 0):
        out.append(form(b"AIFF", comm(2, 60, 16, rate) + ssnd))
    out += [form(b"AIFF", chunk(b"NAME", b"abc", little=False) + comm(2, 60, 16, 44100) + ssnd),
            form(b"AIFF", comm(2, 60, 16, 44100, comp=b"NONE") + ssnd), form(b"AIFC", comm(2, 60, 16, 44100) + ssnd),
            form(b"AIFC", comm(2, 60, 16, 44100, comp=b"NONE") + ssnd), form(b"AIFC", comm(2, 60, 16, 44100, comp=b"sowt") + ssnd),
            form(b"AIFC", comm(2, 60, 16, 44100, comp=b"SOWT") + ssnd), form(b"AIFC", comm(2, 60, 16, 44100, comp=b"ulaw") + ssnd),
            form(b"AIFF", comm(2, 30, 32, 44100) + ssnd), form(b"AIFF", comm(2, 60, 20, 44100) + ssnd), form(b"AIFF", comm(2, 80, 12, 44100) + ssnd),
            form(b"AIFF", comm(2, 62, 16, 44100) + ssnd), form(b"AIFF", comm(2, 63, 16, 44100) + ssnd), form(b"AIFF", comm(2, 60, 16, 44100)),
            form(b"AIFF", comm(0, 60, 16, 44100) + ssnd), form(b"AIFF", comm(2, 0, 16, 44100) + ssnd), form(b"AIFF", ssnd + comm(2, 60, 16, 44100)),
            form(b"AIFF", comm(2, 60, 16, 44100) + chunk(b"SSND", struct.pack(">II", 4, 0) + pcm, little=False)),
            form(b"AIFX", comm(2, 60, 16, 44100) + ssnd)]
    return out


def test_container_parser_matches_the_reference_codecs(ref):
    """Every header decision of Wav.cpp:225-353 / AiffBase.cpp:149-281 / Aiff.cpp / Aifc.cpp, decided by those files themselves."""
    ok = 0
    for data in corpus():
        for depth in (32, 24, 16):
            ok += same_verdict(ref, data, depth) == abi.CONTAINER_OK
    assert ok > 60


def test_container_parser_matches_the_reference_codecs_on_truncated_files(ref):
    for data in corpus()[:9] + corpus()[21:24]:
        for cut in list(range(0, min(len(data), 120))) + [len(data) - 1, len(data) - 7]:
            if cut >= 12:
                same_verdict(ref, data[:cut])
            elif cut >= 0:
                # CodecAiffBase::Recognise compares all 12 bytes of its stack buffer however few the stream delivered
                # (AiffBase.cpp:25-32; CodecWav checks the count, Wav.cpp:93): below 12 bytes the reference's answer depends on
                # what the stack held before.  Here such a stream is simply not recognised.
                assert capi.container_parse(data[:cut])[0] == abi.CONTAINER_E_UNRECOGNISED


def test_container_parser_matches_the_reference_codecs_on_mutated_headers(ref):
    """Random damage to the first bytes of valid files: sizes, tags, depths, rates, chunk ids."""
    rng = np.random.default_rng(77)
    bases = [wav_bytes(48000, 24, 2, 300)[0], wav_bytes(44100, 16, 6, 300)[0], aiff_bytes(44100, 16, 2, 300)[0],
             aiff_bytes(96000, 24, 6, 300, sowt=True)[0], riff(fmt() + chunk(b"LIST", b"INFOabc") + chunk(b"data", bytes(200)))]
    verdicts = {}
    for it in range(4000):
        data = bytearray(bases[it % len(bases)])
        head = min(len(data), 72)
        for _ in range(int(rng.integers(1, 4))):
            i = int(rng.integers(0, head))
            data[i] = int(rng.choice((0, 1, 2, 8, 16, 24, 32, 0xff, 0xfe, data[i] ^ (1 << int(rng.integers(0, 8))), int(rng.integers(0, 256)))))
        rc = same_verdict(ref, bytes(data))
        verdicts[rc] = verdicts.get(rc, 0) + 1
    assert len(verdicts) >= 4, verdicts   # the mutations reach accept, ended, corrupt and unsupported
