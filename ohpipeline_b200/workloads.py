"""Synthetic workloads: stream specs + ramp events for BASELINE.json's configs and for randomized parity tests.

Pure host logic (numpy only).  Everything is a function of explicit seeds so that every arm -- the CUDA path,
the C oracle and the linked reference -- sees identical inputs.
"""
import numpy as np

from . import abi

MS = abi.JIFFIES_PER_MS


class Workload:
    """streams: STREAM_SPEC array; events: RAMP_EVENT array; in_bytes/out_bytes: arena sizes; name; seed."""

    def __init__(self, name, streams, events, in_bytes, out_bytes, seed):
        self.name = name
        self.streams = streams
        self.events = events
        self.in_bytes = int(in_bytes)
        self.out_bytes = int(out_bytes)
        self.seed = int(seed)

    @property
    def total_frames(self):
        return int(self.streams["total_frames"].sum())

    @property
    def total_subsamples(self):
        return int((self.streams["total_frames"] * self.streams["channels"]).sum())


def _align(x, a):
    return (x + a - 1) // a * a


def layout(streams, align=16, slack_bytes=None):
    """Assign src_base/dst_base so streams tile the arenas (each stream `align`-byte aligned).

    Output can exceed input only through inserted silence, which callers account for via slack_bytes
    (per-stream extra output bytes)."""
    src = 0
    dst = 0
    for i in range(len(streams)):
        s = streams[i]
        fb = int(s["channels"]) * int(s["bit_depth"]) // 8
        nbytes = int(s["total_frames"]) * fb
        streams[i]["src_base"] = src
        streams[i]["dst_base"] = dst
        src = _align(src + nbytes, align)
        extra = 0 if slack_bytes is None else int(slack_bytes[i])
        dst = _align(dst + nbytes + extra, align)
    return src, dst


def _spec(rate, bits, ch, le, chunk_frames, total_frames, out_fmt=abi.OUT_PACKED_BE, driver_block_frames=0):
    s = np.zeros(1, dtype=abi.STREAM_SPEC)[0]
    s["sample_rate"] = rate
    s["bit_depth"] = bits
    s["channels"] = ch
    s["in_little_endian"] = int(le)
    s["chunk_frames"] = chunk_frames
    s["out_fmt"] = out_fmt
    s["total_frames"] = total_frames
    s["driver_block_frames"] = driver_block_frames
    return s


def _events(lst):
    ev = np.zeros(len(lst), dtype=abi.RAMP_EVENT)
    for i, (at, stage, op, arg) in enumerate(lst):
        ev[i] = (at, stage, op, arg, 0)
    return ev


def _finish(name, specs, per_stream_events, seed, slack=None):
    streams = np.array(specs, dtype=abi.STREAM_SPEC)
    evs = []
    first = 0
    for i, lst in enumerate(per_stream_events):
        lst = sorted(lst, key=lambda e: e[0])
        streams[i]["first_event"] = first
        streams[i]["num_events"] = len(lst)
        first += len(lst)
        evs.extend(lst)
    events = _events(evs)
    in_bytes, out_bytes = layout(streams, slack_bytes=slack)
    return Workload(name, streams, events, in_bytes, out_bytes, seed)


def max_chunk_frames(rate, bits, ch, ms=5):
    """min(5 ms, 9216 B rounded down to whole frames) -- SURVEY 8d config 4."""
    fb = ch * bits // 8
    return max(1, min(rate * ms // 1000, abi.MAX_PCM_CHUNK_BYTES // fb))


def config1(seconds=10.0):
    """BASELINE configs[0]: one stereo 16-bit 44.1 kHz stream, 220-frame chunks; at 3 s a 20 ms ramp down
    (then muted), at 5 s a 20 ms ramp up."""
    rate = 44100
    jps = abi.jiffies_per_sample(rate)
    total = int(rate * seconds)
    spec = _spec(rate, 16, 2, False, 220, total)
    ev = []
    if seconds >= 6:
        ev = [(3 * rate * jps, 0, abi.EV_RAMP_DOWN, 20 * MS), (5 * rate * jps, 0, abi.EV_RAMP_UP, 20 * MS)]
    return _finish("config1: 1 x 2ch/16/44.1k %.3gs, 20 ms down + up" % seconds, [spec], [ev], seed=1 << 32)


def config2(n_streams=1024, seconds=10.0):
    """BASELINE configs[1]: n streams stereo 24-bit 192 kHz, 960-frame (5 ms) chunks, EVERY chunk ramped:
    a ramp down over the first half of the stream and a ramp up over the second half; packed 24-bit BE out."""
    rate = 192000
    jps = abi.jiffies_per_sample(rate)
    total = int(rate * seconds)
    total -= total % 2
    half_j = (total // 2) * jps
    specs, evs = [], []
    for _ in range(n_streams):
        specs.append(_spec(rate, 24, 2, False, 960, total))
        evs.append([(0, 0, abi.EV_RAMP_DOWN, half_j), (half_j, 0, abi.EV_RAMP_UP, half_j)])
    return _finish("config2: %d x 2ch/24/192k %.3gs, full-length ramps, packed 24-bit BE" % (n_streams, seconds),
                   specs, evs, seed=2 << 32)


def config3(n_streams=4096, seconds=1.0, rate=48000, n_events=8, seed=3):
    """BASELINE configs[2]: n streams 8-channel 32-bit LE input, StarvationRamper pattern: K events at random
    jiffy positions (not aligned to chunks or samples): 20 ms down, mute, 50 ms up; a following event may land
    inside the ramp up.  Messages are capped at 5 ms like StarvationRamper's output."""
    jps = abi.jiffies_per_sample(rate)
    total = int(rate * seconds)
    chunk = max_chunk_frames(rate, 32, 8)
    rng = np.random.default_rng(seed)
    specs, evs = [], []
    total_j = total * jps
    for _ in range(n_streams):
        specs.append(_spec(rate, 32, 8, True, chunk, total))
        lst = [(0, 1, abi.EV_MAX_MSG_JIFFIES, 5 * MS)]
        pos = rng.integers(0, max(1, total_j // (n_events + 1)), n_events, dtype=np.int64)
        t = 0
        for k in range(n_events):
            t += int(pos[k]) + 1
            if t >= total_j:
                break
            lst.append((t, 1, abi.EV_RAMP_DOWN, 20 * MS))
            # next audio after the Halt: ramp up 50 ms; sometimes the next starvation lands mid ramp-up
            up_at = t + 20 * MS + int(rng.integers(0, 10 * MS))
            lst.append((up_at, 1, abi.EV_RAMP_UP, 50 * MS))
            t = up_at + (int(rng.integers(1, 50 * MS)) if rng.random() < 0.4 else 50 * MS)
        evs.append(lst)
    return _finish("config3: %d x 8ch/32LE/%dk %.3gs, random starvation ramps" % (n_streams, rate // 1000, seconds),
                   specs, evs, seed=3 << 32)


def config5(n_streams=65536, seconds=1.0):
    """BASELINE configs[4]: n streams stereo 24-bit 96 kHz, 480-frame chunks; same full-length ramps as config 2."""
    rate = 96000
    jps = abi.jiffies_per_sample(rate)
    total = int(rate * seconds)
    total -= total % 2
    half_j = (total // 2) * jps
    specs, evs = [], []
    for _ in range(n_streams):
        specs.append(_spec(rate, 24, 2, False, 480, total))
        evs.append([(0, 0, abi.EV_RAMP_DOWN, half_j), (half_j, 0, abi.EV_RAMP_UP, half_j)])
    return _finish("config5: %d x 2ch/24/96k %.3gs, full-length ramps" % (n_streams, seconds), specs, evs, seed=5 << 32)


def config4(n_streams=16384, seconds=0.25, seed=4):
    """BASELINE configs[3] as SURVEY 8d spells it out: per stream the PRNG picks bits in {8,16,24,32}, 1-8 channels
    (6-ch/32-bit included), one of the eight rates 44.1-384 kHz, wire endianness and sink (P1, or P2 where the depth
    allows); messages are what a codec delivers, min(5 ms, 9216 B).  Ramps as the pipeline elements make them:
    a Ramper-style ramp up at the start (stage 0), StarvationRamper events -- 20 ms down, 50 ms up -- at jiffy
    positions aligned to nothing, some of them cut short by the next one (stage 1: partial ramps, messages split at
    the remaining ramp size), a Muter-style 500 ms ramp down / up on top (stage 2: ramps running against each other,
    Ramp::Set's intersect + split-fragment path), a driver pulling 1-10 ms blocks on a quarter of the streams
    (MsgPlayable::Split at arbitrary byte counts), muted stretches, MsgSilence, one-sample chunks (events one sample
    apart) and, on 16-bit streams, an attenuation window over a tenth of the stream.  `mixed` below is the stress
    version of the same idea (tiny messages, caps of one sample); this one keeps the sizes a pipeline produces."""
    rng = np.random.default_rng(seed)
    rates = (44100, 48000, 88200, 96000, 176400, 192000, 352800, 384000)
    specs, evs, slack = [], [], []
    for i in range(n_streams):
        bits = int(rng.choice((8, 16, 24, 32)))
        ch = int(rng.integers(1, 9))
        rate = int(rng.choice(rates))
        le = bool(rng.integers(0, 2))
        jps = abi.jiffies_per_sample(rate)
        fb = ch * bits // 8
        chunk = max_chunk_frames(rate, bits, ch)
        total = max(chunk, int(rate * seconds))
        total_j = total * jps
        use_silence = rng.random() < 0.1
        # the packed-LE sink ASSERTs on ProcessSilence (TestCodecInteractiveMain.cpp:564-567): P2 streams get no
        # MsgSilence and nothing that mutes
        want_p2 = bits != 32 and not use_silence and rng.random() < 0.25
        block = int(rng.integers(rate // 1000, rate // 100 + 1)) if rng.random() < 0.25 else 0
        spec = _spec(rate, bits, ch, le, chunk, total, abi.OUT_PACKED_LE if want_p2 else abi.OUT_PACKED_BE, block)
        q = jps if use_silence else 1   # a MsgSilence cannot be split inside a sample: sample-aligned event grid
        lst = []
        extra = 0

        def at(lo, hi):
            return int(rng.integers(lo, max(lo + 1, hi))) // q * q

        if rng.random() < 0.5:
            lst.append((0, 0, abi.EV_MUTE, 0))
            lst.append((0, 0, abi.EV_RAMP_UP, 50 * MS))          # Ramper: short ramp up when a stream starts
        t = at(0, total_j // 4)
        for _ in range(int(rng.integers(0, 4))):                 # StarvationRamper: down 20 ms, (halt,) up 50 ms
            if t >= total_j:
                break
            if not want_p2:
                lst.append((t, 1, abi.EV_RAMP_DOWN, 20 * MS))
                up_at = t + (at(1, 20 * MS) if rng.random() < 0.3 else 20 * MS + at(0, 10 * MS))
            else:
                lst.append((t, 1, abi.EV_RAMP_DOWN, 4 * total_j + 20 * MS))   # stays audible
                up_at = t + at(1, 20 * MS)
            lst.append((up_at, 1, abi.EV_RAMP_UP, 50 * MS))
            t = up_at + (at(1, 50 * MS) if rng.random() < 0.3 else 50 * MS + at(0, total_j // 4))
        if rng.random() < 0.3:                                   # Muter on top
            m0 = at(0, total_j)
            lst.append((m0, 2, abi.EV_RAMP_DOWN, (4 * total_j + 500 * MS) if want_p2 else 500 * MS))
            if rng.random() < 0.7:
                lst.append((m0 + at(1, 600 * MS), 2, abi.EV_RAMP_UP, 500 * MS))
        if rng.random() < 0.02:                                  # one-sample chunks
            e0 = at(0, total_j) // jps * jps
            lst.append((e0, 1, abi.EV_RAMP_UP, 50 * MS))
            lst.append((e0 + jps, 1, abi.EV_RAMP_UP, 50 * MS))
        if bits == 16:
            a0 = at(0, total_j - total_j // 10)
            lst.append((a0, 3, abi.EV_SET_ATTENUATION, int(rng.choice((64, 128, 192, 255, 257, 384, 511)))))
            lst.append((a0 + total_j // 10 // q * q, 3, abi.EV_SET_ATTENUATION, abi.UNITY_ATTENUATION))
        if use_silence:
            for _ in range(int(rng.integers(1, 3))):
                sj = int(rng.integers(1, 20 * (rate // 1000) + 1)) * jps
                lst.append((at(0, total_j) // jps * jps, 0, abi.EV_INSERT_SILENCE, sj))
                extra += (sj // jps) * fb
        specs.append(spec)
        evs.append(lst)
        slack.append(extra)
    return _finish("config4: %d streams mixed formats %.3gs, pipeline-shaped ramps" % (n_streams, seconds), specs, evs,
                   seed=(4 << 32) + seed, slack=slack)


def all_rates(channels=2, seconds=0.3):
    """Every PCM rate Jiffies::PerSample accepts (Msg.cpp:424-470: 7350 Hz ... 384 kHz) at every bit depth, the way
    SuiteStarvationRamper sweeps them (TestStarvationRamper.cpp:861-915): one starvation per stream -- ramp down
    kRampDownJiffies = 20 ms from a position aligned to nothing, halt, ramp up 50 ms -- messages of 5 ms."""
    specs, evs = [], []
    k = 0
    for rate in abi.PCM_SAMPLE_RATES:
        jps = abi.jiffies_per_sample(rate)
        for bits in (8, 16, 24, 32):
            chunk = max_chunk_frames(rate, bits, channels)
            total = int(rate * seconds)
            specs.append(_spec(rate, bits, channels, (k % 2) == 1, chunk, total))
            down_at = total * jps // 4 + 977 * k + 1          # mid-message, mid-sample
            up_at = down_at + 20 * MS + 5 * MS
            evs.append([(0, 1, abi.EV_MAX_MSG_JIFFIES, 5 * MS), (down_at, 1, abi.EV_RAMP_DOWN, 20 * MS),
                        (up_at, 1, abi.EV_RAMP_UP, 50 * MS)])
            k += 1
    return _finish("all rates x depths: %d streams, 20 ms down / 50 ms up" % len(specs), specs, evs, seed=7 << 32)


def steady_edges(seed, n_streams=160):
    """Streams made for the walk's bulk step (32 uniform messages at a time): events exactly on, one jiffy and one
    sample either side of message boundaries; ramps whose length is a whole number of messages, give or take one
    jiffy; ramps that finish early; silence insertions on and off message boundaries; driver blocks that divide,
    equal, exceed and are coprime to the message size; a last message that is short; events past the end."""
    rng = np.random.default_rng(seed)
    specs, evs, slack = [], [], []
    for _ in range(n_streams):
        rate = int(rng.choice((44100, 48000, 96000, 192000)))
        bits = int(rng.choice((8, 16, 24, 32)))
        ch = int(rng.choice((1, 2, 6)))
        jps = abi.jiffies_per_sample(rate)
        fb = ch * bits // 8
        chunk = int(rng.choice((max_chunk_frames(rate, bits, ch), 64, 100)))
        chunk = min(chunk, max_chunk_frames(rate, bits, ch))
        msgs = int(rng.integers(1, 150))
        total = msgs * chunk + int(rng.choice((0, 0, 1, chunk // 2)))
        msg_j = chunk * jps
        block = int(rng.choice((0, 0, chunk, chunk // 2, 2 * chunk, chunk + 1, 7, 3 * chunk - 1)))
        use_silence = rng.random() < 0.3
        q = jps if use_silence else 1
        spec = _spec(rate, bits, ch, bool(rng.integers(0, 2)), chunk, total, abi.OUT_PACKED_BE, max(block, 0))
        if chunk == max_chunk_frames(rate, bits, ch) and rng.random() < 0.3:
            spec["codec_read_frames"] = abi.MAX_PCM_CHUNK_BYTES // fb     # born from an Aiff file (CodecAiffBase reads blocks)
        lst, extra = [], 0

        def near_boundary():
            k = int(rng.integers(0, msgs + 2))
            return max(0, k * msg_j + int(rng.choice((0, 0, 1, -1, jps, -jps, msg_j // 2)))) // q * q

        for _ in range(int(rng.integers(0, 5))):
            stage = int(rng.integers(0, 3))
            op = int(rng.choice((abi.EV_RAMP_DOWN, abi.EV_RAMP_UP, abi.EV_RAMP_DOWN, abi.EV_RAMP_UP, abi.EV_MUTE, abi.EV_UNMUTE)))
            dur = max(q, (int(rng.integers(1, 40)) * msg_j + int(rng.choice((0, 0, 1, -1, jps)))) // q * q)
            lst.append((near_boundary(), stage, op, dur if op in (abi.EV_RAMP_DOWN, abi.EV_RAMP_UP) else 0))
        if rng.random() < 0.3:
            lst.append((0, int(rng.integers(0, 3)), abi.EV_MAX_MSG_JIFFIES, int(rng.choice((msg_j, msg_j + 1, 2 * msg_j, max(jps, msg_j // 3))))))
        if bits == 16 and rng.random() < 0.5:
            lst.append((near_boundary(), 3, abi.EV_SET_ATTENUATION, int(rng.integers(0, 512))))
        if use_silence:
            for _ in range(int(rng.integers(1, 3))):
                sj = int(rng.integers(1, 300)) * jps
                lst.append((near_boundary() // jps * jps, 0, abi.EV_INSERT_SILENCE, sj))
                extra += (sj // jps) * fb
        specs.append(spec)
        evs.append(lst)
        slack.append(extra)
    return _finish("steady-edges %d" % seed, specs, evs, seed=(6 << 32) + seed, slack=slack)


def mixed(n_streams=64, seed=4, max_frames=6000, with_silence=True, p2=True):
    """BASELINE configs[3] in miniature: random formats (8/16/24/32-bit, 1-8 ch, 44.1-384 kHz, BE/LE, P1/P2),
    partial ramps, stacked ramps on several stages (-> Ramp::Set merge / intersect / split), muted stretches,
    MsgSilence, attenuation on 16-bit streams, driver-block MsgPlayable::Split, tiny chunks."""
    rng = np.random.default_rng(seed)
    rates = (44100, 48000, 88200, 96000, 176400, 192000, 352800, 384000)
    specs, evs, slack = [], [], []
    for i in range(n_streams):
        bits = int(rng.choice((8, 16, 24, 32)))
        ch = int(rng.integers(1, 9))
        rate = int(rng.choice(rates))
        le = bool(rng.integers(0, 2))
        jps = abi.jiffies_per_sample(rate)
        fb = ch * bits // 8
        cmax = max_chunk_frames(rate, bits, ch)
        chunk = int(rng.integers(1, cmax + 1)) if rng.random() < 0.3 else cmax
        total = int(rng.integers(1, max_frames + 1))
        use_silence = with_silence and rng.random() < 0.3
        # the packed-LE sink ASSERTs on ProcessSilence (TestCodecInteractiveMain.cpp:564-567): such streams get
        # no MsgSilence and nothing that mutes (a completed ramp down mutes what follows)
        want_p2 = p2 and bits != 32 and not use_silence and rng.random() < 0.3
        out_fmt = abi.OUT_PACKED_LE if want_p2 else abi.OUT_PACKED_BE
        block = int(rng.integers(1, 2 * cmax)) if rng.random() < 0.4 else 0
        spec = _spec(rate, bits, ch, le, chunk, total, out_fmt, block)
        total_j = total * jps
        lst = []
        extra = 0
        # sample-aligned event grid when silence is in play (a MsgSilence cannot be split inside a sample)
        q = jps if use_silence else 1
        n_ev = int(rng.integers(0, 7))
        for _ in range(n_ev):
            stage = int(rng.integers(0, 3))
            at = int(rng.integers(0, total_j + 1)) // q * q
            dur = int(rng.integers(1, max(2, 2 * total_j))) // q * q + q
            op = int(rng.choice((abi.EV_RAMP_DOWN, abi.EV_RAMP_UP, abi.EV_RAMP_DOWN, abi.EV_RAMP_UP, abi.EV_MUTE, abi.EV_UNMUTE)))
            if want_p2 and op in (abi.EV_RAMP_DOWN, abi.EV_MUTE):
                # keep it audible: ramp down only part of the way by giving it far more time than the stream has
                op, dur = abi.EV_RAMP_DOWN, (4 * total_j + jps) // q * q + q
            lst.append((at, stage, op, dur if op in (abi.EV_RAMP_DOWN, abi.EV_RAMP_UP) else 0))
        msg_frames = chunk
        if rng.random() < 0.3:
            cap = max(jps, int(rng.integers(jps, 5 * MS)) // q * q)
            lst.append((0, int(rng.integers(0, 3)), abi.EV_MAX_MSG_JIFFIES, cap))
            msg_frames = max(1, min(chunk, cap // jps))
        if want_p2 and -(-total // msg_frames) + n_ev > 6000:
            # Ramp::Set rounds every message's share of the ramp UP (Msg.cpp:606-611), so a long stream of tiny
            # messages reaches kMin however slow the ramp is -- and then mutes.  Such P2 streams only ramp up.
            lst = [(at, st, abi.EV_RAMP_UP if op == abi.EV_RAMP_DOWN else op, arg) for (at, st, op, arg) in lst]
        if bits == 16 and rng.random() < 0.5:
            lst.append((int(rng.integers(0, total_j + 1)) // q * q, 3, abi.EV_SET_ATTENUATION, int(rng.integers(0, 512))))
        if use_silence:
            for _ in range(int(rng.integers(1, 3))):
                sj = int(rng.integers(1, 200)) * jps
                lst.append((int(rng.integers(0, total_j)) // jps * jps, 0, abi.EV_INSERT_SILENCE, sj))
                extra += (sj // jps) * fb
        specs.append(spec)
        evs.append(lst)
        slack.append(extra)
    return _finish("mixed: %d streams, seed %d" % (n_streams, seed), specs, evs, seed=(4 << 32) + seed, slack=slack)


def elements(seed, n_streams=24, seconds=0.6, illegal=False, attenuation=False):
    """Streams whose stages ARE the reference's elements (include/ohp_schedule.h, ops 8-12): stage 0 a Ramper (a stream that
    starts with its 50 ms ramp up, sometimes a second MsgDecodedStream mid-stream, MsgHalt), stage 1 a StarvationRamper
    (the reservoir runs dry at positions aligned to nothing: ramp up from silence over 50 ms afterwards, messages held to
    5 ms), stage 2 a Muter (Mute / Unmute in turn, 30 ms ramps, some calls landing inside the ramp the previous one
    started, MsgHalt while ramping down and while muted), MsgSilence entering at the top now and then (every element reacts
    to it differently), a driver pulling fixed blocks on some streams.  tests/test_elements_vs_reference.py runs them
    through the element objects themselves.  illegal: Mute() twice in a row on some streams -- the reference ASSERTS.
    attenuation: on the 16-bit streams an Attenuator ahead of one of the stages changes its mind a few times
    (OHP_EV_SET_ATTENUATION; drawn from a generator of their own, so the streams are otherwise those of the same seed)."""
    rng = np.random.default_rng(seed)
    rng_att = np.random.default_rng((seed << 8) | 0xA7)
    specs, evs, slack = [], [], []
    for i in range(n_streams):
        rate = int(rng.choice((44100, 48000, 96000, 192000)))
        bits = int(rng.choice((8, 16, 24, 32)))
        ch = int(rng.choice((1, 2, 2, 6, 8)))
        jps = abi.jiffies_per_sample(rate)
        fb = ch * bits // 8
        chunk = max_chunk_frames(rate, bits, ch)
        total = int(rate * seconds * rng.uniform(0.3, 1.0))
        total_j = total * jps
        use_silence = rng.random() < 0.4
        block = int(rng.integers(rate // 1000, rate // 100 + 1)) if rng.random() < 0.3 else 0
        spec = _spec(rate, bits, ch, bool(rng.integers(0, 2)), chunk, total, abi.OUT_PACKED_BE, block)
        if rng.random() < 0.25:
            spec["codec_read_frames"] = abi.MAX_PCM_CHUNK_BYTES // fb
        q = jps if use_silence else 1   # a MsgSilence cannot be split inside a sample: sample-aligned event grid
        lst, extra = [], 0

        def at(lo, hi):
            return int(rng.integers(lo, max(lo + 1, hi))) // q * q

        which = rng.random(3) < 0.7
        if which[0]:   # Ramper
            lst.append((0, 0, abi.EV_RAMPER_STREAM, 50 * MS if rng.random() < 0.8 else 0))
            if rng.random() < 0.4:
                lst.append((at(0, total_j), 0, abi.EV_RAMPER_STREAM, 50 * MS if rng.random() < 0.7 else 0))
            if rng.random() < 0.3:
                lst.append((at(0, 60 * MS), 0, abi.EV_HALT, 0))
        if which[1]:   # StarvationRamper
            t = at(0, total_j // 2)
            for _ in range(int(rng.integers(1, 4))):
                lst.append((t, 1, abi.EV_STARVATION, 50 * MS))
                t += at(1, 120 * MS)
            if rng.random() < 0.3:
                lst.append((at(0, total_j), 1, abi.EV_HALT, 0))
        if which[2]:   # Muter
            t = at(0, total_j // 3)
            op = abi.EV_MUTER_MUTE
            for k in range(int(rng.integers(1, 6))):
                lst.append((t, 2, op, 30 * MS))
                if illegal and rng.random() < 0.3:
                    lst.append((t + q, 2, op, 30 * MS))
                op = abi.EV_MUTER_UNMUTE if op == abi.EV_MUTER_MUTE else abi.EV_MUTER_MUTE
                t += at(1, 25 * MS) if rng.random() < 0.4 else 30 * MS + at(0, 60 * MS)
            if rng.random() < 0.4:
                lst.append((at(0, total_j), 2, abi.EV_HALT, 0))
        if attenuation and bits == 16:
            stage = int(rng_att.integers(0, 3))
            for _ in range(int(rng_att.integers(1, 4))):
                lst.append((int(rng_att.integers(0, total_j)) // q * q, stage, abi.EV_SET_ATTENUATION, int(rng_att.choice((1, 64, 100, 255, 256)))))
        if use_silence:
            for _ in range(int(rng.integers(1, 4))):
                sj = int(rng.integers(1, 12 * (rate // 1000) + 1)) * jps
                lst.append((at(0, total_j) // jps * jps, 0, abi.EV_INSERT_SILENCE, sj))
                extra += (sj // jps) * fb
        specs.append(spec)
        evs.append(lst)
        slack.append(extra)
    return _finish("elements %d" % seed, specs, evs, seed=(8 << 32) + seed, slack=slack)
