"""The class-free schedule walk (ohpipeline_b200/host/schedule_walk.h -- the source the GPU schedule kernels compile,
run here on host threads through ohp_schedule_build_walk) against the reference's playables (tests/golden) and
against the class-based host model.  CPU only; tests/test_gpu_schedule.py repeats this on the device."""
import glob
import os

import numpy as np
import pytest

from ohpipeline_b200 import abi, capi, workloads

GOLDEN = sorted(g for g in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not g.endswith(("ramp_algebra.npz", "flywheel.npz")))


def check(w_streams, w_events):
    try:
        host = capi.schedule_build(w_streams, w_events)
    except capi.OhpError as eh:
        with pytest.raises(capi.OhpError) as ew:
            capi.schedule_build(w_streams, w_events, walk=True)
        assert ew.value.status == eh.status
        return None
    walk = capi.schedule_build(w_streams, w_events, walk=True)
    assert np.array_equal(walk.stream_chunk_begin, host.stream_chunk_begin)
    assert np.array_equal(walk.stream_out_bytes, host.stream_out_bytes)
    assert np.array_equal(walk.chunks, host.chunks)
    assert np.array_equal(walk.info, host.info)
    return host


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_walk_reproduces_reference_playables(path):
    g = np.load(path)
    walk = capi.schedule_build(g["streams"], g["events"], threads=2, walk=True)
    assert np.array_equal(walk.chunks, g["chunks"])
    assert np.array_equal(walk.info, g["info"])


@pytest.mark.parametrize("seed", range(40, 52))
def test_walk_matches_class_model_on_mixed_workloads(seed):
    w = workloads.mixed(n_streams=80, seed=seed, max_frames=4000)
    check(w.streams, w.events)


def test_walk_matches_class_model_on_baseline_configs():
    for w in (workloads.config1(6.2), workloads.config2(n_streams=8, seconds=0.5),
              workloads.config3(n_streams=40, seconds=1.0), workloads.config5(n_streams=16, seconds=0.25)):
        assert check(w.streams, w.events) is not None


def test_walk_refuses_what_the_class_model_refuses():
    w = workloads.config5(n_streams=4, seconds=0.02)
    for field, index, value, status in (("sample_rate", 2, 12345, abi.E_INVALID_ARG), ("bit_depth", 1, 20, abi.E_INVALID_DESC),
                                        ("out_fmt", 0, abi.OUT_PLANAR32_BE, abi.E_INVALID_ARG),
                                        ("chunk_frames", 3, 0, abi.E_INVALID_ARG)):
        bad = w.streams.copy()
        bad[field][index] = value
        with pytest.raises(capi.OhpError) as e:
            capi.schedule_build(bad, w.events, walk=True)
        assert e.value.status == status, field
        assert check(bad, w.events) is None
    # seed 21 of the 3001-stream mix holds a MsgSilence split below one sample: the reference ASSERTs
    w = workloads.mixed(n_streams=3001, seed=21, max_frames=700)
    assert check(w.streams, w.events) is None


@pytest.mark.parametrize("seed", range(8))
def test_bulk_step_edges_match_class_model(seed):
    w = workloads.steady_edges(seed)
    # stream by stream, so that one the reference ASSERTs on does not hide its neighbours
    asserted = 0
    for k in range(len(w.streams)):
        s = w.streams[k:k + 1].copy()
        ev = w.events[int(s[0]["first_event"]):int(s[0]["first_event"]) + int(s[0]["num_events"])].copy()
        s[0]["first_event"] = 0
        if check(s, ev) is None:
            asserted += 1
    assert asserted < len(w.streams) // 4


@pytest.mark.parametrize("seed", [4, 21, 22, 23])
def test_walk_matches_class_model_on_pipeline_shaped_mix(seed):
    w = workloads.config4(n_streams=300, seconds=0.3, seed=seed)
    assert check(w.streams, w.events) is not None
