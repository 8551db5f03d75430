"""The ramp-setting ELEMENTS of the reference against this repo's statements of them.

oracle/_ref links Ramper.cpp, Muter.cpp, StarvationRamper.cpp and VolumeRamper.cpp unmodified; oracle/ref_elements.cpp pulls
audio through the element OBJECTS the way the reference's own suites do (Media/Tests/TestRamper.cpp:164-170,
TestMuter.cpp:295-338): a fake upstream, IMute::Mute()/Unmute() called from another thread, a MsgDecodedStream or MsgHalt
handed in, the StarvationRamper's reservoir played dry -- each at the stream position an element-level event names
(include/ohp_schedule.h, OHP_EV_RAMPER_STREAM ... OHP_EV_STARVATION).  The playables those objects produce must be the ones

  * the stage model produces on the reference's own message classes (stage_chain.h instantiated in ref_harness.cpp),
  * the product's host mirror (ohp_schedule_build), its class-free walk (what the GPU compiles) and the C port produce,

descriptor for descriptor, and the bytes read out of them must be the port's.  CPU only (tests/test_gpu_schedule.py holds
the device walk against the golden file recorded here); skipped where oracle/_ref did not travel."""
import os

import numpy as np
import pytest

from ohpipeline_b200 import abi, capi, workloads

MS = abi.JIFFIES_PER_MS
STARVED_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "starvation_flywheel.npz")


def one_stream(w, s):
    st = w.streams[s:s + 1].copy()
    ev = w.events[int(st[0]["first_event"]):int(st[0]["first_event"]) + int(st[0]["num_events"])].copy()
    st[0]["first_event"] = 0
    return st, ev


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_stage_model_is_the_element_objects(ref, port, seed):
    w = workloads.elements(seed, n_streams=24)
    inp = port.fill_pcm(w.in_bytes, w.seed)
    rc, out, chunks, info, begin, outb, generated = ref.elements_run(w.streams, w.events, inp, w.out_bytes)
    assert rc == 0
    ops = set(int(o) for o in w.events["op"])
    assert {abi.EV_RAMPER_STREAM, abi.EV_MUTER_MUTE, abi.EV_MUTER_UNMUTE, abi.EV_HALT, abi.EV_STARVATION, abi.EV_INSERT_SILENCE} <= ops
    assert generated.sum() >= 20, "no StarvationRamper played its flywheel ramp"
    dirs = set(int(d) for d in info["direction"])
    assert {abi.DIR_UP, abi.DIR_DOWN} <= dirs and ((chunks["flags"] & abi.F_SILENCE) != 0).any()
    # the stage model on the reference's message classes
    rc2, out2, chunks2, info2 = ref.run(w.streams, w.events, inp, w.out_bytes, threads=2)
    assert rc2 == 0
    assert np.array_equal(chunks2, chunks) and np.array_equal(info2, info) and np.array_equal(out2, out)
    # the product's host mirror, the walk the GPU compiles, the C port
    for sched in (capi.schedule_build(w.streams, w.events), capi.schedule_build(w.streams, w.events, walk=True)):
        assert np.array_equal(sched.chunks, chunks), "descriptors differ from the element objects' playables"
        assert np.array_equal(sched.info, info)
        assert np.array_equal(sched.stream_out_bytes, outb) and np.array_equal(sched.stream_chunk_begin, begin)
    rc3, out3, chunks3, info3 = port.run(w.streams, w.events, inp, w.out_bytes)
    assert rc3 == 0
    assert np.array_equal(chunks3, chunks) and np.array_equal(info3, info) and np.array_equal(out3, out)


def without_endless_cuts(w):
    """The streams of w without a MsgSilence inside the last millisecond before a starvation that plays: there the real
    StarvationRamper's cut to kTrainingJiffies may not terminate (ohp_schedule.h, recent_jiffies; the harness answers -3 for
    such a stream, and a whole-batch call would stop at it)."""
    specs, evs = [], []
    for s in range(len(w.streams)):
        st, ev = one_stream(w, s)
        try:
            sv = capi.schedule_build(st, ev).starvations
        except capi.OhpError:
            sv = np.zeros(0, dtype=abi.STARVATION)
        if (sv["recent_jiffies"][sv["plays"] == 1] < abi.FLYWHEEL_TRAINING_JIFFIES).any():
            continue
        specs.append(st[0])
        evs.append([tuple(int(x) for x in e)[:4] for e in ev])
    streams = np.array(specs, dtype=abi.STREAM_SPEC)
    first = 0
    for i, lst in enumerate(evs):
        streams[i]["first_event"] = first
        first += len(lst)
    return workloads.Workload(w.name, streams, workloads._events([e for lst in evs for e in lst]), w.in_bytes, w.out_bytes, w.seed)


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_attenuated_streams_through_the_element_objects(ref, port, seed):
    """The same with an Attenuator's SetAttenuation calls (Attenuator.cpp:55-58) landing between the elements of the 16-bit
    streams: the playables carry the attenuation the messages were given, the bytes are attenuated after the ramp."""
    w = without_endless_cuts(workloads.elements(seed, n_streams=60, attenuation=True))
    assert len(w.streams) >= 50 and (w.events["op"] == abi.EV_SET_ATTENUATION).sum() >= 10
    inp = port.fill_pcm(w.in_bytes, w.seed)
    rc, out, chunks, info, begin, outb, generated = ref.elements_run(w.streams, w.events, inp, w.out_bytes)
    assert rc == 0
    assert (chunks["attenuation"] != abi.UNITY_ATTENUATION).sum() >= 50
    for sched in (capi.schedule_build(w.streams, w.events), capi.schedule_build(w.streams, w.events, walk=True)):
        assert np.array_equal(sched.chunks, chunks) and np.array_equal(sched.info, info)
    rc3, out3, chunks3, info3 = port.run(w.streams, w.events, inp, w.out_bytes)
    assert rc3 == 0 and np.array_equal(chunks3, chunks) and np.array_equal(out3, out)


def test_every_stage_on_its_own(ref, port):
    """One element per stream, so that a difference names its element."""
    w = workloads.elements(9, n_streams=30)
    inp = port.fill_pcm(w.in_bytes, w.seed)
    seen = set()
    for s in range(len(w.streams)):
        st, ev = one_stream(w, s)
        for stage in (0, 1, 2):
            keep = ev[(ev["stage"] == stage) | (ev["op"] == abi.EV_INSERT_SILENCE)].copy()
            if not (keep["op"] != abi.EV_INSERT_SILENCE).any():
                continue
            st1 = st.copy()
            st1[0]["num_events"] = len(keep)
            rc, out, chunks, info, _, _, _ = ref.elements_run(st1, keep, inp, w.out_bytes, want_audio=False)
            assert rc == 0
            got = capi.schedule_build(st1, keep, walk=True)
            assert np.array_equal(got.chunks, chunks) and np.array_equal(got.info, info), (s, stage, [tuple(e) for e in keep])
            seen.add(stage)
    assert seen == {0, 1, 2}


def test_what_the_muter_asserts_on_is_refused(ref, port):
    """Mute() while already muting, Unmute() while running: ASSERTS() in the reference (Muter.cpp:88-90, 107-111)."""
    w = workloads.config5(n_streams=1, seconds=0.1)
    st = w.streams.copy()
    for ops in ((abi.EV_MUTER_MUTE, abi.EV_MUTER_MUTE), (abi.EV_MUTER_UNMUTE,), (abi.EV_MUTER_MUTE, abi.EV_MUTER_UNMUTE, abi.EV_MUTER_UNMUTE)):
        ev = np.zeros(len(ops), dtype=abi.RAMP_EVENT)
        for i, op in enumerate(ops):
            ev[i] = (1000 + 5 * MS * i, 2, op, 30 * MS, 0)
        st[0]["first_event"], st[0]["num_events"] = 0, len(ev)
        inp = np.zeros(w.in_bytes, dtype=np.uint8)
        assert ref.elements_run(st, ev, inp, w.out_bytes, want_audio=False)[0] == -1
        assert ref.run(st, ev, inp, w.out_bytes, want_audio=False)[0] == -1
        assert port.schedule_run(st, ev)[0] == -1
        for walk in (False, True):
            with pytest.raises(capi.OhpError) as e:
                capi.schedule_build(st, ev, walk=walk)
            assert e.value.status == abi.E_INVALID_DESC
    # an element's ops do not mix with the bare ramp ops on one stage, nor with another element's
    for ops in ((abi.EV_MUTER_MUTE, abi.EV_RAMP_UP), (abi.EV_RAMPER_STREAM, abi.EV_STARVATION)):
        ev = np.zeros(2, dtype=abi.RAMP_EVENT)
        for i, op in enumerate(ops):
            ev[i] = (1000 + 5 * MS * i, 2, op, 30 * MS, 0)
        st[0]["first_event"], st[0]["num_events"] = 0, 2
        assert port.schedule_run(st, ev)[0] == -2
        for walk in (False, True):
            with pytest.raises(capi.OhpError) as e:
                capi.schedule_build(st, ev, walk=walk)
            assert e.value.status == abi.E_INVALID_ARG


def test_known_element_behaviours(ref, port):
    """The differences VERDICT r1 asked about, each pinned on the element object: a Muter called during its own ramp turns the
    ramp round where it is; a MsgSilence ends a Ramper's and a Muter's ramp on the spot; a StarvationRamper ramps up from
    silence after it starved and holds messages to 5 ms; a halted Muter mutes without a ramp."""
    w = workloads.config5(n_streams=1, seconds=0.2)       # 96 kHz stereo 24-bit, 480-frame (5 ms) messages
    st = w.streams.copy()
    jps = abi.jiffies_per_sample(96000)
    inp = port.fill_pcm(w.in_bytes, 3)

    def run(events):
        ev = np.zeros(len(events), dtype=abi.RAMP_EVENT)
        for i, e in enumerate(sorted(events)):
            ev[i] = e + (0,)
        st[0]["first_event"], st[0]["num_events"] = 0, len(ev)
        rc, out, chunks, info, _, _, gen = ref.elements_run(st, ev, inp, w.out_bytes + (1 << 16), want_audio=False)
        assert rc == 0
        got = capi.schedule_build(st, ev, walk=True)
        assert np.array_equal(got.chunks, chunks) and np.array_equal(got.info, info)
        return chunks, info, gen

    # Muter: Mute, and Unmute 10 ms into the 30 ms ramp down: the ramp up takes the 10 ms back, not 30 (Muter.cpp:112-121)
    chunks, info, _ = run([(20 * MS, 2, abi.EV_MUTER_MUTE, 30 * MS), (30 * MS, 2, abi.EV_MUTER_UNMUTE, 30 * MS)])
    down = np.nonzero(info["direction"] == abi.DIR_DOWN)[0]
    up = np.nonzero(info["direction"] == abi.DIR_UP)[0]
    assert int(info["jiffies"][down].sum()) == 10 * MS and int(info["jiffies"][up].sum()) == 10 * MS
    assert int(chunks["ramp_start"][down[0]]) == abi.RAMP_MAX and int(chunks["ramp_end"][up[-1]]) == abi.RAMP_MAX
    assert int(chunks["ramp_end"][down[-1]]) == int(chunks["ramp_start"][up[0]])
    # Muter: muted before any audio has passed (halted) -> no ramp at all, silence from the first message (Muter.cpp:62-64)
    chunks, info, _ = run([(0, 2, abi.EV_MUTER_MUTE, 30 * MS)])
    assert (info["direction"] == abi.DIR_NONE).all() and ((chunks["flags"] & abi.F_SILENCE) != 0).all()
    # Ramper: a MsgSilence 10 ms into the 50 ms ramp up ends it; the audio after it plays at full level (Ramper.cpp:106-112)
    chunks, info, _ = run([(0, 0, abi.EV_RAMPER_STREAM, 50 * MS), (10 * MS, 0, abi.EV_INSERT_SILENCE, 480 * jps)])
    up = np.nonzero(info["direction"] == abi.DIR_UP)[0]
    assert int(info["jiffies"][up].sum()) == 10 * MS
    assert ((chunks["flags"][up[-1] + 1:] & abi.F_RAMP_ENABLED) == 0).all()
    # StarvationRamper: starves 12.5 ms in: 20 ms of flywheel audio play (not part of the stream), then 50 ms up from silence
    chunks, info, gen = run([(12 * MS + MS // 2, 1, abi.EV_STARVATION, 50 * MS)])
    assert int(gen[0]) == 20
    up = np.nonzero(info["direction"] == abi.DIR_UP)[0]
    assert int(chunks["ramp_start"][up[0]]) == 0 and int(chunks["ramp_end"][up[-1]]) == abi.RAMP_MAX
    assert int(info["jiffies"][up].sum()) == 50 * MS and int(info["jiffies"].max()) <= 5 * MS
    # ... and a second starvation before any audio has passed the first one's ramp start changes nothing (:628-629)
    again, info2, gen2 = run([(12 * MS + MS // 2, 1, abi.EV_STARVATION, 50 * MS), (12 * MS + MS // 2, 1, abi.EV_STARVATION, 50 * MS)])
    assert np.array_equal(again, chunks) and int(gen2[0]) == 20


def test_volume_ramper_hands_on_the_median_multiplier_and_clears_the_ramp(ref):
    """VolumeRamper::ProcessAudio (VolumeRamper.cpp:111-122): when the stream is volume-ramped every message's
    MedianRampMultiplier goes to IVolumeRamper and the message leaves unramped (muted ones stay muted); a MsgSilence sends
    zero; a sample-ramped stream is left alone.  ohp_median_multiplier is that multiplier."""
    rng = np.random.default_rng(12)
    ramps, kinds = [], []
    for i in range(200):
        a, b = int(rng.integers(64, abi.RAMP_MAX + 1)), int(rng.integers(64, abi.RAMP_MAX + 1))
        r = rng.random()
        if r < 0.15:
            ramps.append((abi.RAMP_MAX, abi.RAMP_MAX, abi.DIR_NONE, 0))
        elif r < 0.25:
            ramps.append((0, 0, abi.DIR_MUTE, 1))
        else:
            ramps.append((a, b, abi.DIR_NONE if a == b else (abi.DIR_UP if a < b else abi.DIR_DOWN), 1))
        kinds.append(int(rng.random() < 0.1))
    mult, after = ref.volume_ramper(True, ramps, kinds)
    assert len(mult) == len(ramps)
    for i, (r, k) in enumerate(zip(ramps, kinds)):
        if k == 1:
            assert int(mult[i]) == 0 and tuple(int(x) for x in after[i]) == r     # ProcessMsg(MsgSilence): kMultiplierZero, untouched
            continue
        assert int(mult[i]) == capi.median_multiplier(r[0], r[1], r[2], r[3]), (i, r)
        if r[3] and r[2] != abi.DIR_MUTE:
            assert tuple(int(x) for x in after[i]) == (abi.RAMP_MAX, abi.RAMP_MAX, abi.DIR_NONE, 0)   # MsgAudio::MedianRampMultiplier clears it
        else:
            assert tuple(int(x) for x in after[i]) == r
    mult, after = ref.volume_ramper(False, ramps, kinds)
    assert len(mult) == 0 and all(tuple(int(x) for x in after[i]) == ramps[i] for i in range(len(ramps)))


def _flywheel_on_cpu(port, st, starvation, inp):
    """ohp_flywheel_plan's three launches, run by the C port: FlywheelInput::Prepare, FlywheelRamperManager::Ramp,
    RampGenerator's ramped 1 ms blocks -> the bytes a driver reads."""
    prep, job, blocks = capi.flywheel_plan(st, starvation)
    assert 1 <= len(prep) <= abi.FLYWHEEL_MAX_PREP
    rc, training = port.process_chunks(prep, inp, int(job["train_frames"][0]) * 4 * int(st[0]["channels"]))
    assert rc == 0
    fb = int(st[0]["channels"]) * int(st[0]["bit_depth"]) // 8
    rc, raw = port.flywheel(job, training, int(job["out_frames"][0]) * fb)
    assert rc == 0
    rc, played = port.process_chunks(blocks, raw, raw.size)
    assert rc == 0
    return played


@pytest.mark.parametrize("rate,ch,bits,le", [(44100, 2, 16, True), (48000, 2, 24, False), (96000, 2, 24, False), (192000, 2, 24, True),
                                             (96000, 6, 32, False), (176400, 1, 16, False), (88200, 8, 24, True), (384000, 2, 8, False),
                                             (32000, 3, 16, True)])
def test_planned_flywheel_audio_is_what_the_starved_element_plays(ref, port, rate, ch, bits, le):
    """The real StarvationRamper object, starved at positions aligned to nothing (twice: the second time while it is still
    ramping up from the first), against ohp_schedule_build's starvation records + ohp_flywheel_plan: the training block is the
    last millisecond of PCM that passed the element, ramps cleared (FlywheelPlayableCreator, StarvationRamper.cpp:23-88), and
    RampGenerator starts from the element's ramp value (:526-531)."""
    jps = abi.jiffies_per_sample(rate)
    total = rate * 3 // 10
    spec = workloads._spec(rate, bits, ch, le, workloads.max_chunk_frames(rate, bits, ch), total)
    first = 37 * MS + 12345
    second = first + 17 * MS + 777          # 17 ms into the 50 ms ramp up
    events = [(0, 0, abi.EV_RAMPER_STREAM, 40 * MS),        # a Ramper upstream: its ramp is on the messages the element keeps
              (first, 1, abi.EV_STARVATION, 50 * MS), (second, 1, abi.EV_STARVATION, 50 * MS)]
    w = workloads._finish("starved", [spec], [events], seed=77)
    inp = port.fill_pcm(w.in_bytes, 1000 + rate + ch)
    rc, audio, ramps = ref.elements_generated_audio(w.streams, w.events, inp)
    assert rc == 0 and len(ramps) == 2
    sched = capi.schedule_build(w.streams, w.events)
    sv = sched.starvations
    assert len(sv) == 2 and list(sv["plays"]) == [1, 1] and list(sv["event"]) == [1, 2]
    assert [int(r) for r in sv["ramp"]] == [int(r) for r in ramps]
    assert int(sv["ramp"][0]) == abi.RAMP_MAX and 0 < int(sv["ramp"][1]) < abi.RAMP_MAX
    assert [int(f) for f in sv["pcm_jiffies"]] == [first, second]
    fb = ch * bits // 8
    per = abi.FLYWHEEL_RAMP_JIFFIES // jps * fb
    assert audio.size == 2 * per
    for k in range(2):
        played = _flywheel_on_cpu(port, w.streams, sv[k:k + 1], inp)
        assert np.array_equal(played, audio[k * per:(k + 1) * per]), (rate, ch, bits, le, k)


def test_starvations_that_play_nothing_are_recorded_as_such(ref, port):
    """Starved before any audio has passed (Starting / Halted) or again before the ramp up has moved: no flywheel ramp
    (StarvationRamper.cpp:628-629, 640-650); ohp_flywheel_plan refuses those, and one with under 1 ms of PCM behind it."""
    w = workloads.config5(n_streams=1, seconds=0.2)
    st = w.streams.copy()
    inp = port.fill_pcm(w.in_bytes, 3)
    at = 12 * MS + MS // 2

    def run(events):
        ev = workloads._events(sorted(events))
        st[0]["first_event"], st[0]["num_events"] = 0, len(ev)
        rc, audio, ramps = ref.elements_generated_audio(st, ev, inp)
        assert rc == 0
        return capi.schedule_build(st, ev).starvations, audio, ramps

    sv, audio, ramps = run([(0, 1, abi.EV_STARVATION, 50 * MS)])
    assert len(sv) == 1 and int(sv["plays"][0]) == 0 and audio.size == 0 and len(ramps) == 0
    with pytest.raises(capi.OhpError) as e:
        capi.flywheel_plan(st, sv[0:1])
    assert e.value.status == abi.E_INVALID_ARG
    sv, audio, ramps = run([(at, 1, abi.EV_STARVATION, 50 * MS), (at, 1, abi.EV_STARVATION, 50 * MS)])
    assert [int(p) for p in sv["plays"]] == [1, 0] and len(ramps) == 1
    assert np.array_equal(_flywheel_on_cpu(port, st, sv[0:1], inp), audio)
    # half a millisecond after a MsgHalt: the halt did not empty the element's recent audio, the block reaches back across it
    sv, audio, ramps = run([(at, 1, abi.EV_HALT, 0), (at + MS // 2, 1, abi.EV_STARVATION, 50 * MS)])
    assert len(sv) == 1 and int(sv["plays"][0]) == 1 and int(sv["recent_jiffies"][0]) == at + MS // 2 and len(ramps) == 1
    assert np.array_equal(_flywheel_on_cpu(port, st, sv[0:1], inp), audio)
    # ... but starved before any audio has followed the halt, it plays nothing
    sv, audio, ramps = run([(at, 1, abi.EV_HALT, 0), (at, 1, abi.EV_STARVATION, 50 * MS)])
    assert len(sv) == 1 and int(sv["plays"][0]) == 0 and audio.size == 0
    # half a millisecond after a MsgSilence: the reference's training block holds the silence; not planned
    jps = abi.jiffies_per_sample(96000)
    # (the MsgSilence enters ahead of the message that begins at 15 ms; event positions count it)
    sv, audio, ramps = run([(15 * MS, 0, abi.EV_INSERT_SILENCE, 96 * jps), (16 * MS + MS // 2, 1, abi.EV_STARVATION, 50 * MS)])
    assert len(sv) == 1 and int(sv["plays"][0]) == 1 and int(sv["recent_jiffies"][0]) < MS and len(ramps) == 1 and audio.size > 0
    with pytest.raises(capi.OhpError) as e:
        capi.flywheel_plan(st, sv[0:1])
    assert e.value.status == abi.E_INVALID_ARG
    # ... and 3 ms after it the block is PCM again
    sv, audio, ramps = run([(15 * MS, 0, abi.EV_INSERT_SILENCE, 96 * jps), (19 * MS, 1, abi.EV_STARVATION, 50 * MS)])
    assert int(sv["recent_jiffies"][0]) == 3 * MS and int(sv["pcm_jiffies"][0]) == 18 * MS
    assert np.array_equal(_flywheel_on_cpu(port, st, sv[0:1], inp), audio)


def test_planned_flywheel_audio_keeps_the_messages_attenuation(ref, port):
    """FlywheelPlayableCreator clears a message's ramp, not its attenuation (StarvationRamper.cpp:61-74; MsgPlayablePcm::Read
    applies it, Msg.cpp ApplyAttenuation): a 16-bit stream attenuated ahead of the element trains the flywheel on the
    attenuated samples.  A change of attenuation inside the last millisecond is not planned."""
    rate, ch, bits = 44100, 2, 16
    total = rate * 2 // 10
    spec = workloads._spec(rate, bits, ch, False, workloads.max_chunk_frames(rate, bits, ch), total)
    at = 60 * MS + 4321
    for att_stage in (0, 1):
        events = [(10 * MS, att_stage, abi.EV_SET_ATTENUATION, 100), (at, 1, abi.EV_STARVATION, 50 * MS)]
        w = workloads._finish("starved", [spec], [events], seed=78)
        inp = port.fill_pcm(w.in_bytes, 555)
        rc, audio, ramps = ref.elements_generated_audio(w.streams, w.events, inp)
        assert rc == 0 and len(ramps) == 1
        sv = capi.schedule_build(w.streams, w.events).starvations
        assert int(sv["attenuation"][0]) == 100 and int(sv["recent_jiffies"][0]) == at - 10 * MS
        played = _flywheel_on_cpu(port, w.streams, sv[0:1], inp)
        assert np.array_equal(played, audio)
        plain = sv.copy()
        plain["attenuation"] = abi.UNITY_ATTENUATION
        assert not np.array_equal(_flywheel_on_cpu(port, w.streams, plain[0:1], inp), audio)
    events = [(at - MS // 3, 0, abi.EV_SET_ATTENUATION, 100), (at, 1, abi.EV_STARVATION, 50 * MS)]
    w = workloads._finish("starved", [spec], [events], seed=78)
    sv = capi.schedule_build(w.streams, w.events).starvations
    assert int(sv["plays"][0]) == 1 and int(sv["recent_jiffies"][0]) == MS // 3
    with pytest.raises(capi.OhpError) as e:
        capi.flywheel_plan(w.streams, sv[0:1])
    assert e.value.status == abi.E_INVALID_ARG


def test_golden_starvation_audio_is_current(ref, port):
    """tests/golden/starvation_flywheel.npz is what the element object plays today (tests/golden/make_golden_starvation.py)."""
    from flywheel_util import starved_streams
    g = np.load(STARVED_GOLDEN)
    quirk = 0
    for name, w, seed in starved_streams():
        inp = port.fill_pcm(w.in_bytes, seed)
        rc, audio, ramps = ref.elements_generated_audio(w.streams, w.events, inp)
        assert rc == 0 and np.array_equal(audio, g["audio_" + name]) and np.array_equal(ramps, g["ramps_" + name]), name
        for sv in capi.schedule_build(w.streams, w.events).starvations:
            quirk += len(capi.flywheel_plan(w.streams, sv)[0]) > 1
    assert quirk >= 3, "no stream left a frame too many in its training block"


def test_planned_flywheel_audio_reproduces_the_golden_file(port):
    """No reference needed: the records of ohp_schedule_build, ohp_flywheel_plan and the C port against the recorded audio."""
    from flywheel_util import starved_streams
    g = np.load(STARVED_GOLDEN)
    for name, w, seed in starved_streams():
        inp = port.fill_pcm(w.in_bytes, seed)
        sv = capi.schedule_build(w.streams, w.events).starvations
        assert [int(r) for r in sv["ramp"][sv["plays"] == 1]] == [int(r) for r in g["ramps_" + name]]
        played = np.concatenate([_flywheel_on_cpu(port, w.streams, sv[k:k + 1], inp) for k in range(len(sv)) if sv["plays"][k]])
        assert np.array_equal(played, g["audio_" + name]), name


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_planned_flywheel_audio_on_random_element_schedules(ref, port, seed):
    """Random element schedules (Ramper, StarvationRamper and Muter stages, halts, inserted silence, starvations wherever the
    PRNG put them): every starvation the real StarvationRamper plays through, planned and played by the C port, byte for byte;
    what the plan refuses is a block with silence in it (or a shape FlywheelRamper does not take), never one it gets wrong."""
    w = workloads.elements(seed, n_streams=24)
    inp = port.fill_pcm(w.in_bytes, w.seed)
    compared = refused = 0
    for s in range(len(w.streams)):
        st, ev = one_stream(w, s)
        if not (ev["op"] == abi.EV_STARVATION).any():
            continue
        sv = capi.schedule_build(st, ev).starvations
        playing = sv[sv["plays"] == 1]
        if (playing["recent_jiffies"] < abi.FLYWHEEL_TRAINING_JIFFIES).any():
            # Silence inside the last millisecond.  The reference's cut to kTrainingJiffies (StarvationRamper.cpp:495-507) splits
            # the MsgSilence at a jiffy count that need not be a whole sample; MsgSilence::SplitCompleted rounds the front part
            # down (Msg.cpp:2530-2535), the loop comes round with less than a sample of excess, splits off a message of ZERO
            # jiffies, subtracts nothing and never ends.  Such a stream cannot be put through the real element.
            refused += len(playing)
            continue
        rc, audio, ramps = ref.elements_generated_audio(st, ev, inp)
        assert rc == 0
        assert [int(r) for r in playing["ramp"]] == [int(r) for r in ramps], (seed, s)
        jps = abi.jiffies_per_sample(int(st[0]["sample_rate"]))
        per = abi.FLYWHEEL_RAMP_JIFFIES // jps * int(st[0]["channels"]) * int(st[0]["bit_depth"]) // 8
        assert audio.size == per * len(playing), (seed, s)
        for k in range(len(playing)):
            try:
                played = _flywheel_on_cpu(port, st, playing[k:k + 1], inp)
            except capi.OhpError as e:
                assert e.status in (abi.E_INVALID_ARG, abi.E_INVALID_DESC)
                refused += 1
                continue
            assert np.array_equal(played, audio[k * per:(k + 1) * per]), (seed, s, k, playing[k])
            compared += 1
    assert compared >= 8, (compared, refused)


@pytest.mark.parametrize("seed", [31, 32])
def test_a_batch_of_starvations_in_three_launches(ref, port, seed):
    """ohp_flywheel_plan_batch: every starvation of a batch of random element schedules that plays, laid out in three arenas and
    run as THREE calls over the whole batch (the C port standing in for ohp_process_device / ohp_flywheel_device /
    ohp_process_device); what lands at out_off[k] is what the k-th starving element of the reference played."""
    w = without_endless_cuts(workloads.elements(seed, n_streams=40))
    inp = port.fill_pcm(w.in_bytes, w.seed)
    sched = capi.schedule_build(w.streams, w.events)
    sv = sched.starvations
    b = capi.flywheel_plan_batch(w.streams, sv, training_base=5, generated_base=40, out_base=100)
    playing = np.nonzero(sv["plays"] == 1)[0]
    assert len(b.planned) >= 10 and set(int(k) for k in b.planned) <= set(int(k) for k in playing)
    assert len(b.jobs) == len(b.planned) and (b.out_off % 16 == 0).all() and int(b.out_off[0]) == 112
    # what the device calls check before they launch (ohp_validate / ohp_flywheel_validate: host functions of the CUDA library)
    assert capi.validate(b.prep, w.in_bytes, b.training_bytes) == (abi.OK, 0)
    assert capi.flywheel_validate(b.jobs, b.training_bytes, b.generated_bytes) == (abi.OK, 0)
    assert capi.validate(b.blocks, b.generated_bytes, b.out_bytes) == (abi.OK, 0)
    rc, training = port.process_chunks(b.prep, inp, b.training_bytes)
    assert rc == 0
    rc, generated = port.flywheel(b.jobs, training, b.generated_bytes)
    assert rc == 0
    rc, out = port.process_chunks(b.blocks, generated, b.out_bytes)
    assert rc == 0
    # the reference, stream by stream: its starvations that play, in order
    at = {int(k): i for i, k in enumerate(b.planned)}
    compared = 0
    for s in range(len(w.streams)):
        mine = [int(k) for k in playing if int(sv["stream"][k]) == s]
        if not mine:
            continue
        st, ev = one_stream(w, s)
        rc, audio, ramps = ref.elements_generated_audio(st, ev, inp)
        assert rc == 0 and len(ramps) == len(mine)
        per = audio.size // len(mine)
        for j, k in enumerate(mine):
            if k not in at:
                continue
            i = at[k]
            assert int(b.out_len[i]) == per
            assert np.array_equal(out[int(b.out_off[i]):int(b.out_off[i]) + per], audio[j * per:(j + 1) * per]), (s, k)
            compared += 1
    assert compared == len(b.planned)
    # a record that names no stream of the batch
    bad = sv[:1].copy()
    bad["stream"] = len(w.streams)
    with pytest.raises(capi.OhpError) as e:
        capi.flywheel_plan_batch(w.streams, bad)
    assert e.value.status == abi.E_INVALID_ARG
    assert len(capi.flywheel_plan_batch(w.streams, sv[:0]).planned) == 0


def test_plan_refuses_what_the_flywheel_kernel_refuses():
    """Shapes the reference's fixed flywheel buffers do not hold (ohp_flywheel_validate): the plan says so, as the launch would."""
    for rate, ch, bits, ok in [(384000, 8, 32, False), (384000, 6, 8, False), (352800, 5, 24, True), (8000, 2, 16, True),
                               (7350, 2, 16, True), (192000, 8, 32, True)]:
        total = rate // 5
        spec = workloads._spec(rate, bits, ch, False, workloads.max_chunk_frames(rate, bits, ch), total)
        w = workloads._finish("starved", [spec], [[(30 * MS + 999, 1, abi.EV_STARVATION, 50 * MS)]], seed=80)
        sv = capi.schedule_build(w.streams, w.events).starvations
        assert len(sv) == 1 and int(sv["plays"][0]) == 1
        job = capi.flywheel_job(rate, ch, bits)
        assert (capi.flywheel_validate(job, 1 << 20, 1 << 24)[0] == abi.OK) == ok, (rate, ch, bits)
        if ok:
            prep, planned_job, blocks = capi.flywheel_plan(w.streams, sv[0:1])
            assert capi.flywheel_validate(planned_job, 1 << 20, 1 << 24) == (abi.OK, 0)
        else:
            with pytest.raises(capi.OhpError) as e:
                capi.flywheel_plan(w.streams, sv[0:1])
            assert e.value.status == abi.E_INVALID_DESC
            assert len(capi.flywheel_plan_batch(w.streams, sv).planned) == 0


def test_the_cut_that_never_ends_is_recognised_not_run(ref, port):
    """44.1 kHz: 1 ms is 44 samples and 128 jiffies.  A MsgSilence, half a millisecond of audio, then the reservoir runs dry:
    the reference's cut to kTrainingJiffies (StarvationRamper.cpp:495-507) falls inside the silence at a jiffy count that is
    not a whole sample, MsgSilence::SplitCompleted rounds the front part down (Msg.cpp:2530-2535) and the loop goes on
    splitting off messages of zero jiffies for ever.  The harness sees it coming (-3, oracle/ref_elements.cpp CutNeverEnds)
    instead of hanging; the model records the starvation as playing with less than 1 ms of PCM behind it, and the plan
    refuses it.  The same silence at 48 kHz (1 ms = 48 samples exactly) cuts cleanly and plays."""
    for rate, never_ends in ((44100, True), (48000, False)):
        jps = abi.jiffies_per_sample(rate)
        spec = workloads._spec(rate, 16, 2, False, rate // 200, rate // 5)
        msg = (rate // 200) * jps                                 # one 5 ms message
        silence = (rate // 250) * jps                             # 4 ms, enters ahead of the fourth message
        starve = 3 * msg + silence + (rate // 2000) * jps         # half a millisecond of that message has passed
        events = [(3 * msg, 0, abi.EV_INSERT_SILENCE, silence), (starve, 1, abi.EV_STARVATION, 50 * MS)]
        w = workloads._finish("cut", [spec], [events], seed=81)
        inp = port.fill_pcm(w.in_bytes, 7)
        sv = capi.schedule_build(w.streams, w.events).starvations
        assert len(sv) == 1 and int(sv["plays"][0]) == 1 and int(sv["recent_jiffies"][0]) == (rate // 2000) * jps
        with pytest.raises(capi.OhpError) as e:
            capi.flywheel_plan(w.streams, sv[0:1])
        assert e.value.status == abi.E_INVALID_ARG
        rc, audio, ramps = ref.elements_generated_audio(w.streams, w.events, inp)
        if never_ends:
            assert rc == -3
            assert ref.elements_run(w.streams, w.events, inp, w.out_bytes + (1 << 16), want_audio=False)[0] == -3
            # ... and it is the reference that does not come back, not the harness being cautious: let in (in a process of
            # its own), the element is still cutting after five seconds; the 48 kHz stream takes milliseconds
            import subprocess
            import sys
            code = ("import sys, numpy as np; sys.path.insert(0, %r); from oracle import pyoracle; from ohpipeline_b200 import abi;"
                    "st = np.frombuffer(bytes.fromhex(%r), dtype=abi.STREAM_SPEC); ev = np.frombuffer(bytes.fromhex(%r), dtype=abi.RAMP_EVENT);"
                    "print(pyoracle.Ref().elements_generated_audio(st, ev, np.zeros(%d, dtype=np.uint8))[0])"
                    % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), w.streams.tobytes().hex(), w.events.tobytes().hex(), w.in_bytes))
            with pytest.raises(subprocess.TimeoutExpired):
                subprocess.run([sys.executable, "-c", code], env=dict(os.environ, OHP_REF_LET_IT_RUN="1"), timeout=5,
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            done = subprocess.run([sys.executable, "-c", code], timeout=60, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            assert done.returncode == 0 and done.stdout.strip() == "-3", done.stderr[-2000:]
        else:
            assert rc == 0 and len(ramps) == 1 and audio.size == abi.FLYWHEEL_RAMP_JIFFIES // jps * 4


def _flywheel_from_recent_on_cpu(port, st, starvation, recent, inp):
    prep, job, blocks = capi.flywheel_plan_recent(st, starvation, recent)
    assert capi.validate(prep, len(inp), int(job["train_frames"][0]) * 4 * int(st[0]["channels"])) == (abi.OK, 0)
    rc, training = port.process_chunks(prep, inp, int(job["train_frames"][0]) * 4 * int(st[0]["channels"]))
    assert rc == 0
    rc, raw = port.flywheel(job, training, int(job["out_frames"][0]) * int(st[0]["channels"]) * int(st[0]["bit_depth"]) // 8)
    assert rc == 0
    rc, played = port.process_chunks(blocks, raw, raw.size)
    assert rc == 0
    return played, prep


def test_silence_inside_the_last_millisecond_is_planned_from_the_recent_audio(ref, port):
    """ohp_flywheel_plan_recent: the element's recent audio piece by piece.  96 kHz stereo 24-bit, a 1 ms MsgSilence ahead of
    the message at 15 ms, the reservoir dry half a millisecond later: the training block is half silence, half audio, and
    what the plan plays is what the real element plays.  The PCM-only plan refuses that starvation; where there is no
    silence in the block both plans give the same descriptors."""
    w = workloads.config5(n_streams=1, seconds=0.2)
    st = w.streams.copy()
    inp = port.fill_pcm(w.in_bytes, 3)
    jps = abi.jiffies_per_sample(96000)
    for at, with_silence in ((16 * MS + MS // 2, True), (16 * MS + MS // 4 + 3 * jps, True), (19 * MS, False)):
        ev = workloads._events(sorted([(15 * MS, 0, abi.EV_INSERT_SILENCE, 96 * jps), (at, 1, abi.EV_STARVATION, 50 * MS)]))
        st[0]["first_event"], st[0]["num_events"] = 0, len(ev)
        sched = capi.schedule_build(st, ev)
        sv = sched.starvations
        assert len(sv) == 1 and int(sv["plays"][0]) == 1
        pieces = sched.recent_of(0)
        assert bool((pieces["silence"] != 0).any()) == with_silence or not with_silence
        rc, audio, ramps = ref.elements_generated_audio(st, ev, inp)
        assert rc == 0 and len(ramps) == 1
        played, prep = _flywheel_from_recent_on_cpu(port, st, sv[0:1], pieces, inp)
        assert np.array_equal(played, audio), at
        if with_silence:
            assert ((prep["flags"] & abi.F_SILENCE) != 0).any() and ((prep["flags"] & abi.F_SILENCE) == 0).any()
            with pytest.raises(capi.OhpError):
                capi.flywheel_plan(st, sv[0:1])
        else:
            assert np.array_equal(prep, capi.flywheel_plan(st, sv[0:1])[0])


@pytest.mark.parametrize("seed", [41, 42, 43, 44])
def test_recent_audio_plans_on_random_element_schedules(ref, port, seed):
    """Random element schedules again, planned from the recent audio: everything the PCM-only plan gives comes out the same,
    starvations with silence in their last millisecond are played as the real element plays them, and the plan says "the
    reference does not return" exactly where the harness finds the cut that never ends (-3)."""
    w = workloads.elements(seed, n_streams=30)
    inp = port.fill_pcm(w.in_bytes, w.seed)
    compared = with_silence = never = 0
    for s in range(len(w.streams)):
        st, ev = one_stream(w, s)
        if not (ev["op"] == abi.EV_STARVATION).any():
            continue
        sched = capi.schedule_build(st, ev)
        sv = sched.starvations
        playing_at = np.nonzero(sv["plays"] == 1)[0]
        verdicts = []
        for k in playing_at:
            try:
                verdicts.append(_flywheel_from_recent_on_cpu(port, st, sv[k:k + 1], sched.recent_of(int(k)), inp))
            except capi.OhpError as e:
                verdicts.append(e)
        says_never = any(isinstance(v, capi.OhpError) and v.status == abi.E_INVALID_DESC and "does not return" in str(v) for v in verdicts)
        rc, audio, ramps = ref.elements_generated_audio(st, ev, inp)
        assert (rc == -3) == says_never, (seed, s, rc)
        if rc == -3:
            never += 1
            continue
        assert rc == 0 and len(ramps) == len(playing_at)
        per = audio.size // max(1, len(playing_at))
        for j, (k, v) in enumerate(zip(playing_at, verdicts)):
            if isinstance(v, capi.OhpError):
                continue
            played, prep = v
            assert np.array_equal(played, audio[j * per:(j + 1) * per]), (seed, s, int(k))
            compared += 1
            if ((prep["flags"] & abi.F_SILENCE) != 0).any():
                with_silence += 1
            else:
                try:
                    assert np.array_equal(prep, capi.flywheel_plan(st, sv[k:k + 1])[0])
                except capi.OhpError:
                    pass  # (a change of attenuation inside the block: the PCM-only plan declines)
    assert compared >= 20


def test_a_batch_planned_from_the_recent_audio(ref, port):
    """ohp_flywheel_plan_batch_recent over a whole batch of random element schedules, silence under the starvations and all:
    three calls over the batch, and at out_off[k] what the k-th planned starving element of the reference played.  It plans
    everything the PCM-only batch plans, and more."""
    w = workloads.elements(51, n_streams=60)
    inp = port.fill_pcm(w.in_bytes, w.seed)
    sched = capi.schedule_build(w.streams, w.events)
    sv = sched.starvations
    b = capi.flywheel_plan_batch(w.streams, sv, recent=sched.recent, recent_begin=sched.recent_begin)
    plain = capi.flywheel_plan_batch(w.streams, sv)
    assert set(int(k) for k in plain.planned) < set(int(k) for k in b.planned)
    assert capi.validate(b.prep, w.in_bytes, b.training_bytes) == (abi.OK, 0)
    assert capi.flywheel_validate(b.jobs, b.training_bytes, b.generated_bytes) == (abi.OK, 0)
    rc, training = port.process_chunks(b.prep, inp, b.training_bytes)
    assert rc == 0
    rc, generated = port.flywheel(b.jobs, training, b.generated_bytes)
    assert rc == 0
    rc, out = port.process_chunks(b.blocks, generated, b.out_bytes)
    assert rc == 0
    at = {int(k): i for i, k in enumerate(b.planned)}
    playing = np.nonzero(sv["plays"] == 1)[0]
    compared = silent = 0
    for s in range(len(w.streams)):
        mine = [int(k) for k in playing if int(sv["stream"][k]) == s]
        if not mine:
            continue
        st, ev = one_stream(w, s)
        rc, audio, ramps = ref.elements_generated_audio(st, ev, inp)
        if rc == -3:
            continue
        assert rc == 0 and len(ramps) == len(mine)
        per = audio.size // len(mine)
        for j, k in enumerate(mine):
            if k in at:
                i = at[k]
                assert np.array_equal(out[int(b.out_off[i]):int(b.out_off[i]) + per], audio[j * per:(j + 1) * per]), (s, k)
                compared += 1
                silent += int((sched.recent_of(k)["silence"] != 0).any())
    assert compared >= 40 and silent >= 1


def test_a_reservoir_dry_exactly_where_the_stream_ends(ref, port):
    """The last message has passed and the reservoir is empty: the element starves there as anywhere (found by the campaign,
    seed 200924).  The stream's own playables are what they were; the starvation is recorded and its flywheel ramp planned."""
    w = workloads.config5(n_streams=1, seconds=0.1)
    st = w.streams.copy()
    inp = port.fill_pcm(w.in_bytes, 9)
    end = int(st[0]["total_frames"]) * abi.jiffies_per_sample(96000)
    ev = workloads._events([(end, 1, abi.EV_STARVATION, 50 * MS)])
    st[0]["first_event"], st[0]["num_events"] = 0, 1
    sched = capi.schedule_build(st, ev)
    sv = sched.starvations
    assert len(sv) == 1 and int(sv["plays"][0]) == 1 and int(sv["pcm_jiffies"][0]) == end
    rc, audio, ramps = ref.elements_generated_audio(st, ev, inp)
    assert rc == 0 and len(ramps) == 1 and int(ramps[0]) == int(sv["ramp"][0])
    assert np.array_equal(_flywheel_on_cpu(port, st, sv[0:1], inp), audio)
    assert np.array_equal(_flywheel_from_recent_on_cpu(port, st, sv[0:1], sched.recent_of(0), inp)[0], audio)
    # the walk the GPU compiles produces the same playables (it keeps no starvation records)
    assert np.array_equal(capi.schedule_build(st, ev, walk=True).chunks, sched.chunks)
