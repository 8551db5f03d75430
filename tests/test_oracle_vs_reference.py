"""Pins the oracle: the plain-C restatement (oracle/ohp_oracle.c) and the host message model against the REFERENCE'S OWN
CODE (oracle/_ref/libohref.so = /root/reference's Msg.cpp + ProcessorAudioUtils.cpp compiled unmodified).
Skipped where the linked reference is absent.  CPU only."""
import numpy as np
import pytest

from ohpipeline_b200 import abi, capi, workloads as W


def compare(port, ref, w):
    w.streams["out_fmt"] = abi.OUT_PACKED_BE  # the linked reference has the packed-BE sink (ProcessorPcmBufTest) only
    inp = port.fill_pcm(w.in_bytes, w.seed)
    rc_r, out_r, ch_r, inf_r = ref.run(w.streams, w.events, inp, w.out_bytes)
    rc_p, ch_p, inf_p, beg_p, ob_p = port.schedule_run(w.streams, w.events)
    assert (rc_p != 0) == (rc_r != 0), "assert behaviour differs: port %d reference %d" % (rc_p, rc_r)
    try:
        sched = capi.schedule_build(w.streams, w.events, threads=1)
        rc_m = 0
    except capi.OhpError:
        rc_m = -1
    assert (rc_m != 0) == (rc_r != 0)
    if rc_r != 0:
        return None
    assert np.array_equal(ch_p, ch_r) and np.array_equal(inf_p, inf_r), "oracle port descriptors differ"
    assert np.array_equal(sched.chunks, ch_r) and np.array_equal(sched.info, inf_r), "host mirror descriptors differ"
    rc, out_p = port.process_chunks(ch_p, inp, w.out_bytes)
    assert rc == 0
    assert np.array_equal(out_p, out_r), "oracle port audio differs from the reference"
    return ch_r


def test_ramp_table_is_the_references(port, ref):
    assert np.array_equal(port.ramp_array, ref.ramp_array)


def test_jiffies_per_sample(port, ref):
    for rate in abi.PCM_SAMPLE_RATES:
        assert port.jiffies_per_sample(rate) == ref.jiffies_per_sample(rate) == capi.jiffies_per_sample(rate) \
            == abi.jiffies_per_sample(rate)
    assert port.jiffies_per_sample(44000) == 0


def test_baseline_configs(port, ref):
    for w in (W.config1(6.5), W.config2(4, 0.3), W.config3(16, 1.0), W.config3(4, 0.2, rate=192000), W.config5(4, 0.3)):
        assert compare(port, ref, w) is not None


def test_every_rate_and_depth(port, ref):
    """All 18 rates x 8/16/24/32 bits through a starvation (SuiteStarvationRamper's sweep, TestStarvationRamper.cpp:861-915)."""
    assert compare(port, ref, W.all_rates()) is not None
    assert compare(port, ref, W.all_rates(channels=6, seconds=0.2)) is not None


@pytest.mark.parametrize("seed", [4, 31, 32])
def test_pipeline_shaped_mix(port, ref, seed):
    """BASELINE configs[3] as the bench runs it (workloads.config4): the reference, the port and the host mirror agree."""
    assert compare(port, ref, W.config4(n_streams=120, seconds=0.12, seed=seed)) is not None


@pytest.mark.parametrize("seed", [3, 51, 52])
def test_bulk_step_edge_cases(port, ref, seed):
    """Events on and next to message boundaries, whole-message ramps, Aiff-born streams, driver blocks (workloads.
    steady_edges) -- stream by stream, so that one the reference ASSERTs on does not hide its neighbours."""
    w = W.steady_edges(seed, n_streams=60)
    refused = 0
    for k in range(len(w.streams)):
        one = W.Workload(w.name, w.streams[k:k + 1].copy(), w.events, w.in_bytes, w.out_bytes, w.seed)
        if compare(port, ref, one) is None:
            refused += 1
    assert refused < 6


@pytest.mark.parametrize("seed", range(40))
def test_mixed_schedules(port, ref, seed):
    compare(port, ref, W.mixed(n_streams=24, seed=1000 + seed))


def test_mixed_schedules_cover_merge_and_split_paths(port, ref):
    """The fuzz must actually reach Ramp::Set's merge/intersect/split and the mute-split quirk."""
    kinds = {"split_up_down": 0, "flat_enabled": 0, "silence": 0, "zero_bytes": 0, "mute_split_flat0": 0}
    for seed in range(40):
        ch = compare(port, ref, W.mixed(n_streams=24, seed=1000 + seed))
        if ch is None:
            continue
        en = (ch["flags"] & abi.F_RAMP_ENABLED) != 0
        sil = (ch["flags"] & abi.F_SILENCE) != 0
        kinds["flat_enabled"] += int((en & ~sil & (ch["ramp_start"] == ch["ramp_end"])).sum())
        kinds["silence"] += int(sil.sum())
        kinds["zero_bytes"] += int((ch["bytes"] == 0).sum())
        kinds["mute_split_flat0"] += int((en & ~sil & (ch["ramp_start"] == 0) & (ch["ramp_end"] == 0)).sum())
        up_then_down = en[:-1] & en[1:] & (ch["ramp_start"][:-1] < ch["ramp_end"][:-1]) & (ch["ramp_start"][1:] > ch["ramp_end"][1:]) \
            & (ch["ramp_end"][:-1] == ch["ramp_start"][1:])
        kinds["split_up_down"] += int(up_then_down.sum())
    for k, v in kinds.items():
        assert v > 0, kinds


def test_ramp_set_and_split_fuzz(port, ref):
    rng = np.random.default_rng(11)
    K = abi.RAMP_MAX
    for i in range(4000):
        cs, ce, start = (int(x) for x in rng.integers(0, K + 1, 3))
        mode = rng.integers(0, 4)
        if mode == 0:
            cur = (K, K, abi.DIR_NONE, 0)
        elif mode == 1:
            cur = (0, 0, abi.DIR_MUTE, 1)
        else:
            cur = (cs, ce, abi.DIR_NONE if cs == ce else (abi.DIR_UP if cs < ce else abi.DIR_DOWN), 1)
        frag = int(rng.integers(1, 400000))
        dur = frag + int(rng.integers(0, 4000000))
        d = int(rng.choice((abi.DIR_UP, abi.DIR_DOWN)))
        a = ref.ramp_set(cur, start, frag, dur, d)
        for impl in (port.ramp_set, capi.ramp_set):
            b = impl(cur, start, frag, dur, d)
            assert (a[0] < 0) == (b[0] < 0)
            if a[0] >= 0:
                assert a == b, (cur, start, frag, dur, d)
        if cur[3]:
            size = int(rng.integers(2, 400000))
            new = int(rng.integers(1, size))
            a = ref.ramp_split(cur, new, size)
            for impl in (port.ramp_split, capi.ramp_split):
                b = impl(cur, new, size)
                assert (a[0] < 0) == (b[0] < 0)
                if a[0] >= 0:
                    assert a == b


def test_process_chunks_against_real_playables(port, ref):
    """MsgPlayable::Read on hand-built playables: every depth, channel count incl. 6-ch/32-bit tag, silence pattern."""
    from util import pack_chunks
    rng = np.random.default_rng(5)
    specs = []
    for bits in (8, 16, 24, 32):
        for ch in range(1, 9):
            fb = ch * bits // 8
            for _ in range(4):
                frames = int(rng.integers(1, 9216 // fb + 1))
                kind = rng.integers(0, 4)
                flags = [0, abi.F_RAMP_ENABLED, abi.F_RAMP_ENABLED | abi.F_IN_LITTLE_ENDIAN, abi.F_SILENCE][kind]
                specs.append(dict(bytes=frames * fb, bit_depth=bits, channels=ch, flags=flags,
                                  ramp_start=int(rng.integers(0, 16385)), ramp_end=int(rng.integers(0, 16385)),
                                  attenuation=int(rng.integers(0, 512)) if (bits == 16 and kind != 3 and rng.random() < 0.5) else 256))
    descs, in_bytes, out_bytes = pack_chunks(specs)
    inp = port.fill_pcm(in_bytes, 77)
    rc_r, out_r = ref.process_chunks(descs, inp, out_bytes)
    rc_p, out_p = port.process_chunks(descs, inp, out_bytes)
    assert rc_r == 0 and rc_p == 0
    assert np.array_equal(out_r, out_p)


def test_median_multiplier_against_real_message(port, ref):
    K = abi.RAMP_MAX
    rng = np.random.default_rng(3)
    for _ in range(500):
        s, e = (int(x) for x in rng.integers(0, K + 1, 2))
        d = abi.DIR_NONE if s == e else (abi.DIR_UP if s < e else abi.DIR_DOWN)
        med = s + (e - s) // 2 if d == abi.DIR_UP else (s - (s - e) // 2 if d == abi.DIR_DOWN else s)
        if (K - med + 16) >> 5 >= 512:
            continue  # the reference reads past kRampArray here
        assert port.median_multiplier((s, e, d, 1)) == ref.median_multiplier((s, e, d, 1))
