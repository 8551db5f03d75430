#!/usr/bin/env python
"""What a starved StarvationRamper plays: ohp_schedule_build's starvation records + ohp_flywheel_plan + the C port's three
steps (FlywheelInput sink, flywheel generator, ramped blocks) against the REFERENCE'S OWN element object
(oracle/_ref: StarvationRamper.cpp, FlywheelRamper.cpp, Msg.cpp linked unmodified; oracle/ref_elements.cpp stages the
starvations), on random element schedules (workloads.elements: Ramper / StarvationRamper / Muter stages, halts, inserted
silence, starvations wherever the PRNG put them).  CPU only; needs oracle/_ref (the build container).

    python profiles/starvation_fuzz.py FIRST_SEED LAST_SEED [SECONDS] > profiles/r02_starvation_fuzz.json

Where the reference's own cut loop would not terminate (a MsgSilence under the cut at a jiffy count that is not a whole sample:
include/ohp_schedule.h, ohp_starvation.recent_jiffies) the harness says so instead of hanging (-3) and the stream is counted,
not compared; starvations whose last millisecond holds silence but cuts cleanly are played by the reference and not planned
here (counted)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import __graft_entry__  # noqa: E402

__graft_entry__.build()
from ohpipeline_b200 import abi, capi, workloads  # noqa: E402
from oracle import pyoracle  # noqa: E402


def played_by_plan(port, st, starvation, recent, inp):
    prep, job, blocks = capi.flywheel_plan_recent(st, starvation, recent)
    rc, training = port.process_chunks(prep, inp, int(job["train_frames"][0]) * 4 * int(st[0]["channels"]))
    assert rc == 0
    fb = int(st[0]["channels"]) * int(st[0]["bit_depth"]) // 8
    rc, raw = port.flywheel(job, training, int(job["out_frames"][0]) * fb)
    assert rc == 0
    rc, out = port.process_chunks(blocks, raw, raw.size)
    assert rc == 0
    return out, len(prep) > 1


KEYS = ("seeds", "streams", "starvations_compared", "bytes_compared", "frame_too_many", "ramp_below_max", "played_nothing",
        "streams_the_reference_would_not_return_from", "streams_refused_by_model_and_reference", "not_planned",
        "with_silence_in_the_block")


def worker(first, last, only_stream=None):
    """Seeds [first, last) in THIS process (one stream of one seed when only_stream is given) -> totals."""
    budget = 1e9
    port, ref = pyoracle.Port(), pyoracle.Ref()
    t0 = time.time()
    tot = {"seeds": 0, "streams": 0, "starvations_compared": 0, "bytes_compared": 0, "frame_too_many": 0, "ramp_below_max": 0,
           "played_nothing": 0, "streams_the_reference_would_not_return_from": 0, "streams_refused_by_model_and_reference": 0,
           "not_planned": 0, "with_silence_in_the_block": 0, "differences": []}
    for seed in range(first, last):
        w = workloads.elements(seed, n_streams=24)
        inp = port.fill_pcm(w.in_bytes, w.seed)
        tot["seeds"] += 1
        for s in range(len(w.streams)):
            if only_stream is not None and s != only_stream:
                continue
            st = w.streams[s:s + 1].copy()
            ev = w.events[int(st[0]["first_event"]):int(st[0]["first_event"]) + int(st[0]["num_events"])].copy()
            st[0]["first_event"] = 0
            if not (ev["op"] == abi.EV_STARVATION).any():
                continue
            try:
                sched = capi.schedule_build(st, ev)
                sv = sched.starvations
            except capi.OhpError:
                sched = sv = None
            rc, audio, ramps = ref.elements_generated_audio(st, ev, inp)
            if rc == -3:
                # the reference's own cut loop would not terminate (oracle/ref_elements.cpp, CutNeverEnds): the model must have
                # seen silence under the cut too
                seen = False
                if sv is not None:
                    for k in np.nonzero(sv["plays"] == 1)[0]:
                        try:
                            capi.flywheel_plan_recent(st, sv[k:k + 1], sched.recent_of(int(k)))
                        except capi.OhpError as e:
                            seen = seen or (e.status == abi.E_INVALID_DESC and "does not return" in str(e))
                if not seen:
                    tot["differences"].append({"seed": seed, "stream": s, "what": "reference would not return, the plan does not say so"})
                tot["streams_the_reference_would_not_return_from"] += 1
                continue
            if (rc != 0) != (sv is None):
                tot["differences"].append({"seed": seed, "stream": s, "what": "status", "reference": rc})
                continue
            if sv is None:
                tot["streams_refused_by_model_and_reference"] += 1
                continue
            tot["streams"] += 1
            playing_at = np.nonzero(sv["plays"] == 1)[0]
            playing = sv[playing_at]
            tot["played_nothing"] += int((sv["plays"] == 0).sum())
            jps = abi.jiffies_per_sample(int(st[0]["sample_rate"]))
            per = abi.FLYWHEEL_RAMP_JIFFIES // jps * int(st[0]["channels"]) * int(st[0]["bit_depth"]) // 8
            if [int(r) for r in playing["ramp"]] != [int(r) for r in ramps] or audio.size != per * len(playing):
                tot["differences"].append({"seed": seed, "stream": s, "what": "starvations or ramp values"})
                continue
            for k in range(len(playing)):
                try:
                    pieces = sched.recent_of(int(playing_at[k]))
                    out, many = played_by_plan(port, st, playing[k:k + 1], pieces, inp)
                    tot["with_silence_in_the_block"] += int((pieces["silence"] != 0).any() and int(playing["recent_jiffies"][k]) < abi.FLYWHEEL_TRAINING_JIFFIES)
                except capi.OhpError as e:
                    if "does not return" in str(e):
                        tot["differences"].append({"seed": seed, "stream": s, "starvation": k, "what": "the plan says the reference does not return; it did"})
                    tot["not_planned"] += 1  # under 1 ms since the recent audio was emptied, too few frames for the planes, a shape the flywheel does not take
                    continue
                if not np.array_equal(out, audio[k * per:(k + 1) * per]):
                    tot["differences"].append({"seed": seed, "stream": s, "starvation": k, "what": "audio"})
                tot["starvations_compared"] += 1
                tot["bytes_compared"] += per
                tot["frame_too_many"] += int(many)
                tot["ramp_below_max"] += int(int(playing["ramp"][k]) != abi.RAMP_MAX)
    return tot


def run_isolated(args):
    """A worker in a process of its own: (totals or None, return code)."""
    import subprocess
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker"] + [str(a) for a in args],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0:
        return None, r.returncode
    return json.loads(r.stdout.strip().splitlines()[-1]), 0


def main():
    if sys.argv[1] == "--worker":
        only = int(sys.argv[4]) if len(sys.argv) > 4 else None
        print(json.dumps(worker(int(sys.argv[2]), int(sys.argv[3]), only)))
        return 0
    first, last = int(sys.argv[1]), int(sys.argv[2])
    budget = float(sys.argv[3]) if len(sys.argv) > 3 else 1e9
    t0 = time.time()
    tot = {k: 0 for k in KEYS}
    tot["differences"] = []
    # The reference itself can die: FlywheelRamper::BurgsMethod divides by a 32-bit sum of squares that is left to wrap
    # (FlywheelRamper.cpp:252-282); wrapped to exactly zero under a numerator that is not, it is an integer division by zero
    # (SIGFPE) -- with 8-bit audio, whose squares are multiples of 2^16, about once in 2^16 sums.  Seeds run in processes of
    # their own, 40 at a time; where one dies the seeds, then the streams, are run one by one and the stream is counted.
    tot["streams_the_reference_died_on"] = []

    def add(t):
        for k in KEYS:
            tot[k] += t[k]
        tot["differences"] += t["differences"]

    seed = first
    while seed < last and time.time() - t0 < budget:
        hi = min(seed + 40, last)
        t, rc = run_isolated([seed, hi])
        if t is not None:
            add(t)
        else:
            for one in range(seed, hi):
                t, rc = run_isolated([one, one + 1])
                if t is not None:
                    add(t)
                    continue
                for stream in range(24):
                    t, rc = run_isolated([one, one + 1, stream])
                    if t is not None:
                        t["seeds"] = 0
                        add(t)
                    else:
                        tot["streams_the_reference_died_on"].append({"seed": one, "stream": stream, "signal": -rc})
                tot["seeds"] += 1
        seed = hi
    tot["first_seed"], tot["seconds"] = first, round(time.time() - t0, 1)
    print(json.dumps(tot))
    return 1 if tot["differences"] else 0


if __name__ == "__main__":
    sys.exit(main())
