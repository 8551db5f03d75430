"""Flywheel ramp generator on the GPU (ohp_flywheel_device) against the C oracle, the reference's golden vectors
(tests/golden/flywheel.npz, recorded from the real RampGenerator) and, end to end, the whole starvation sequence
FlywheelInput (planar sink) -> flywheel kernel -> RampGenerator's ramped blocks -> ramp + convert kernel."""
import os

import numpy as np
import pytest

from ohpipeline_b200 import abi, capi
from flywheel_util import SHAPES, training_block, train_frames
from util import make_desc

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "flywheel.npz")


def run_jobs(ctx, jobs, inp, out_bytes, fill=0x5A):
    import torch
    d_jobs = torch.from_numpy(jobs.view(np.uint8).copy()).cuda()
    d_in = torch.from_numpy(np.ascontiguousarray(inp)).cuda()
    d_out = torch.full((max(out_bytes, 1),), fill, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.flywheel_device(d_jobs.data_ptr(), len(jobs), d_in.data_ptr(), int(inp.size), d_out.data_ptr(), out_bytes)
    ctx.sync()
    return d_out.cpu().numpy()[:out_bytes]


def batch(kinds, shapes, seed, dst_align=1):
    """One job per (kind, shape): training blocks and outputs laid out back to back (dst optionally misaligned)."""
    jobs, blocks = [], []
    src = dst = 0
    for n, (kind, (rate, ch, bits)) in enumerate((k, s) for k in kinds for s in shapes):
        j = capi.flywheel_job(rate, ch, bits, src_off=src, dst_off=dst)
        t = training_block(rate, ch, kind, seed + n)
        jobs.append(j)
        blocks.append(t)
        src += t.size
        dst += int(j["out_frames"][0]) * ch * bits // 8
        if dst_align > 1:
            dst = (dst + dst_align - 1) // dst_align * dst_align
        else:
            dst += n % 3  # ragged: exercises the byte-store path
    return np.concatenate(jobs), np.concatenate(blocks), dst + 16


@pytest.mark.parametrize("kind", ["tone", "noise", "dc", "zero", "step", "max"])
def test_flywheel_kernel_matches_oracle(ctx, port, kind):
    jobs, inp, out_bytes = batch([kind], SHAPES, seed=300)
    assert capi.flywheel_validate(jobs, inp.size, out_bytes)[0] == abi.OK
    rc, want = port.flywheel(jobs, inp, out_bytes)
    assert rc == 0
    got = run_jobs(ctx, jobs, inp, out_bytes, fill=0)
    if not np.array_equal(got, want):
        i = int(np.nonzero(got != want)[0][0])
        k = int(np.searchsorted(jobs["dst_off"], i, side="right") - 1)
        raise AssertionError("byte %d differs (got %#x want %#x) in job %d: %s" % (i, got[i], want[i], k, jobs[k]))


def test_flywheel_kernel_aligned_destinations_and_many_jobs(ctx, port):
    """Word-store path (4-byte aligned destinations), more jobs than one wave of warps."""
    jobs, inp, out_bytes = batch(["tone", "noise", "step"] * 40, SHAPES, seed=500, dst_align=16)
    assert len(jobs) == 1200
    rc, want = port.flywheel(jobs, inp, out_bytes)
    assert rc == 0
    got = run_jobs(ctx, jobs, inp, out_bytes, fill=0)
    assert np.array_equal(got, want)


def test_flywheel_kernel_reproduces_reference_golden_vectors(ctx):
    g = np.load(GOLDEN)
    for k in range(len(g["jobs"])):
        job = g["jobs"][k:k + 1].copy()
        raw = g["raw_%d" % k]
        got = run_jobs(ctx, job, g["training_%d" % k], raw.size)
        assert np.array_equal(got, raw), "case %d: generated audio differs from the reference's RampGenerator" % k


def test_whole_starvation_sequence_on_the_gpu(ctx, port):
    """recent audio (wire format) --PLANAR32 sink--> training block --flywheel--> generated audio --ramped blocks-->
    what the driver reads; every stage on the device, every stage checked against the oracle chain."""
    import torch
    rng = np.random.default_rng(77)
    cases = [(48000, 2, 24, True, abi.RAMP_MAX), (44100, 2, 16, False, 7000), (192000, 2, 24, False, abi.RAMP_MAX),
             (96000, 6, 32, True, abi.RAMP_MAX), (88200, 4, 8, False, 300)]
    for rate, ch, bits, le, start in cases:
        T = train_frames(rate)
        B = bits // 8
        # 1. the last 1 ms of audio as two messages (a split lands mid-block), unramped (FlywheelPlayableCreator clears ramps)
        t = np.arange(T)
        pcm = np.zeros((T, ch), dtype=np.int64)
        for c in range(ch):
            pcm[:, c] = (0.5 * np.sin(2 * np.pi * 300.0 * (c + 1) * t / rate) * 2 ** (bits - 1)).astype(np.int64)
        pcm += rng.integers(-3, 4, pcm.shape)
        pcm = np.clip(pcm, -2 ** (bits - 1), 2 ** (bits - 1) - 1)
        be = np.zeros((T, ch, B), dtype=np.uint8)
        for b in range(B):
            be[:, :, b] = (pcm >> (8 * (B - 1 - b))) & 0xff
        wire = (be[:, :, ::-1] if le else be).reshape(-1).copy()
        cut = (T // 3) * ch * B
        flags = abi.F_IN_LITTLE_ENDIAN if (le and B > 1) else 0
        d1 = make_desc(src_off=0, dst_off=0, bytes=cut, bit_depth=bits, channels=ch, flags=flags,
                       out_fmt=abi.OUT_PLANAR32_BE, aux=T)
        d2 = make_desc(src_off=cut, dst_off=(T // 3) * 4, bytes=wire.size - cut, bit_depth=bits, channels=ch, flags=flags,
                       out_fmt=abi.OUT_PLANAR32_BE, aux=T)
        prep = np.concatenate([d1, d2])
        train_bytes = T * 4 * ch
        rc, want_train = port.process_chunks(prep, wire, train_bytes)
        assert rc == 0
        # 2. the flywheel job and RampGenerator's ramped blocks
        job = capi.flywheel_job(rate, ch, bits)
        gen_bytes = int(job["out_frames"][0]) * ch * B
        rc, want_gen = port.flywheel(job, want_train, gen_bytes)
        assert rc == 0
        blocks, final = capi.flywheel_ramp_chunks(job, start, 0, 0)
        rc, want_out = port.process_chunks(blocks, want_gen, gen_bytes)
        assert rc == 0
        assert final == 0  # a flywheel ramp always ends in silence
        # the same three steps on the device, buffers chained in HBM
        d_wire = torch.from_numpy(wire).cuda()
        d_train = torch.zeros(train_bytes, dtype=torch.uint8, device="cuda")
        d_gen = torch.zeros(gen_bytes, dtype=torch.uint8, device="cuda")
        d_out = torch.zeros(gen_bytes, dtype=torch.uint8, device="cuda")
        d_prep = torch.from_numpy(prep.view(np.uint8).copy()).cuda()
        d_job = torch.from_numpy(job.view(np.uint8).copy()).cuda()
        d_blocks = torch.from_numpy(blocks.view(np.uint8).copy()).cuda()
        torch.cuda.synchronize()
        ctx.process_device(d_prep.data_ptr(), len(prep), d_wire.data_ptr(), wire.size, d_train.data_ptr(), train_bytes)
        ctx.flywheel_device(d_job.data_ptr(), 1, d_train.data_ptr(), train_bytes, d_gen.data_ptr(), gen_bytes)
        ctx.process_device(d_blocks.data_ptr(), len(blocks), d_gen.data_ptr(), gen_bytes, d_out.data_ptr(), gen_bytes)
        ctx.sync()
        assert np.array_equal(d_train.cpu().numpy(), want_train), (rate, ch, bits)
        assert np.array_equal(d_gen.cpu().numpy(), want_gen), (rate, ch, bits)
        assert np.array_equal(d_out.cpu().numpy(), want_out), (rate, ch, bits)


def test_device_rejects_bad_flywheel_jobs_loudly(ctx):
    ok = capi.flywheel_job(48000, 2, 24)
    bad = ok.copy()
    bad["bit_depth"] = 20
    jobs = np.concatenate([ok, bad])
    jobs["dst_off"][1] = 8192
    inp = training_block(48000, 2, "tone", 1)
    with pytest.raises(capi.OhpError) as e:
        run_jobs(ctx, jobs, np.concatenate([inp, inp]), 16384)
    assert e.value.status == abi.E_INVALID_DESC and "flywheel job 1" in str(e.value)
    far = ok.copy()
    far["dst_off"] = 1 << 40
    with pytest.raises(capi.OhpError) as e:
        run_jobs(ctx, far, inp, 16384)
    assert e.value.status == abi.E_OUT_OF_RANGE
    # and the context keeps working
    got = run_jobs(ctx, ok, inp, 5760)
    assert got.any()


def test_planned_starvations_on_the_gpu_play_what_the_reference_element_plays(ctx, port):
    """ohp_schedule_build's starvation records -> ohp_flywheel_plan -> the three launches on the device, against
    tests/golden/starvation_flywheel.npz: what the reference's own StarvationRamper object played when it was starved at the
    same positions (recorded by tests/golden/make_golden_starvation.py).  Some of these training blocks are a frame too long
    in the reference (44.1 kHz family): one-subsample planar descriptors at odd offsets."""
    import torch
    from flywheel_util import starved_streams
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "starvation_flywheel.npz"))
    many = 0
    for name, w, seed in starved_streams():
        inp = port.fill_pcm(w.in_bytes, seed)
        d_in = torch.from_numpy(inp).cuda()
        sv = capi.schedule_build(w.streams, w.events).starvations
        played = []
        for k in range(len(sv)):
            if not sv["plays"][k]:
                continue
            prep, job, blocks = capi.flywheel_plan(w.streams, sv[k:k + 1])
            many += len(prep) > 1
            ch, B = int(job["channels"][0]), int(job["bit_depth"][0]) // 8
            train_bytes = int(job["train_frames"][0]) * 4 * ch
            gen_bytes = int(job["out_frames"][0]) * ch * B
            d_train = torch.full((train_bytes,), 0xEE, dtype=torch.uint8, device="cuda")
            d_gen = torch.zeros(gen_bytes, dtype=torch.uint8, device="cuda")
            d_out = torch.zeros(gen_bytes, dtype=torch.uint8, device="cuda")
            d_prep = torch.from_numpy(prep.view(np.uint8).copy()).cuda()
            d_job = torch.from_numpy(job.view(np.uint8).copy()).cuda()
            d_blocks = torch.from_numpy(blocks.view(np.uint8).copy()).cuda()
            torch.cuda.synchronize()
            ctx.process_device(d_prep.data_ptr(), len(prep), d_in.data_ptr(), inp.size, d_train.data_ptr(), train_bytes)
            ctx.flywheel_device(d_job.data_ptr(), 1, d_train.data_ptr(), train_bytes, d_gen.data_ptr(), gen_bytes)
            ctx.process_device(d_blocks.data_ptr(), len(blocks), d_gen.data_ptr(), gen_bytes, d_out.data_ptr(), gen_bytes)
            ctx.sync()
            rc, want_train = port.process_chunks(prep, inp, train_bytes)
            assert rc == 0 and np.array_equal(d_train.cpu().numpy(), want_train), (name, k)
            played.append(d_out.cpu().numpy())
        assert np.array_equal(np.concatenate(played), g["audio_" + name]), name
    assert many >= 3
