// schedule_walk.h -- one stream's walk through the ramp-setting stages on plain data, compiled for BOTH the device
// (csrc/ohp_schedule_kernels.cuh: one thread per stream) and the host (schedule.cpp: ohp_schedule_walk_host, the
// CPU-testable twin the "-m 'not gpu'" suite compares with the class-based host model).
//
// It restates stage_chain.h + msg_model.cpp without classes, allocation, recursion or exceptions.  The ramp algebra
// (Ramp::Set / Ramp::Split, Msg.cpp:590-807) is ramp_core.h, the source the host mirror compiles too; the
// message-level steps follow
//     MsgAudio::Split             Msg.cpp:1949-1969     msg_split
//     MsgSilence::SplitCompleted  Msg.cpp:2522-2545     msg_split (silence branch)
//     MsgAudio::SetRamp           Msg.cpp:1989-2046     msg_set_ramp
//     MsgAudioPcm::CreatePlayable Msg.cpp:2234-2262     create_playable
//     MsgSilence::CreatePlayable  Msg.cpp:2472-2492     create_playable (silence branch)
//     MsgPlayable::Split          Msg.cpp:2591-2624     playable_split
// and the stage idiom of Ramper.cpp:114-134, Muter.cpp:210-262, StarvationRamper.cpp:579-603, 791-832.
#pragma once

#include <cstdint>

#include "../../include/ohp_schedule.h"
#include "ramp_core.h"
#include "codec_source.h"

namespace ohp {
namespace sched {

constexpr int kStages = (int)OHP_MAX_STAGES;
constexpr int kStackDepth = 16; // pending split remainders per stage (each Process pushes at most four)

// per-stream result codes
constexpr uint32_t kOk = 0u;
constexpr uint32_t kErrAssert = 1u;   // the reference would ASSERT                 -> OHP_E_INVALID_DESC
constexpr uint32_t kErrSpec = 2u;     // a spec the message model cannot represent  -> OHP_E_INVALID_ARG
constexpr uint32_t kErrDepth = 3u;    // more pending splits than kStackDepth       -> OHP_E_NO_MEMORY

struct Msg
{
    uint64_t cell;     // arena offset of the audio this message (and every split of it) refers to
    uint32_t size;     // jiffies (MsgAudio::iSize)
    uint32_t offset;   // jiffies into the cell (MsgAudio::iOffset)
    uint32_t total;    // MsgSilence::iSizeJiffiesTotal
    uint32_t atten;    // MsgAudioPcm::iAttenuation
    core::RampPod ramp;
    uint32_t silence;  // 1: MsgSilence
};

// 32-byte stack entry
struct Packed
{
    uint64_t cell;
    uint32_t size, offset, total;
    uint32_t ramp_se;  // start | end << 16
    uint32_t misc;     // atten | direction << 16 | enabled << 24 | silence << 25
    uint32_t pad;
};

OHP_HD Packed pack(const Msg& m)
{
    Packed p;
    p.cell = m.cell; p.size = m.size; p.offset = m.offset; p.total = m.total;
    p.ramp_se = m.ramp.start | (m.ramp.end << 16);
    p.misc = (m.atten & 0xffffu) | (m.ramp.direction << 16) | (m.ramp.enabled << 24) | (m.silence << 25);
    p.pad = 0;
    return p;
}
OHP_HD Msg unpack(const Packed& p)
{
    Msg m;
    m.cell = p.cell; m.size = p.size; m.offset = p.offset; m.total = p.total;
    m.ramp.start = p.ramp_se & 0xffffu; m.ramp.end = p.ramp_se >> 16;
    m.atten = p.misc & 0xffffu; m.ramp.direction = (p.misc >> 16) & 0xffu; m.ramp.enabled = (p.misc >> 24) & 1u;
    m.silence = (p.misc >> 25) & 1u;
    return m;
}

struct Playable
{
    uint64_t arena;
    uint32_t size;     // bytes
    uint32_t jiffies;
    uint32_t atten;
    core::RampPod ramp;
    uint32_t silence;
};

enum Mode : uint32_t { Running = 0, RampingDown = 1, RampingUp = 2, Muted = 3 };

struct Stage
{
    uint64_t pos;
    uint32_t mode, current, remaining, maxMsg, attenuation, nextEv, depth;
};

struct StreamCtx
{
    const ohp_ramp_event* ev;
    uint32_t nEv;
    uint32_t jps, frameBytes, channels, bits;
    uint32_t in_le, out_fmt;
    uint32_t blockBytes, blockFill;
    uint64_t dst_base;
    // sink
    uint64_t nChunks, outBytes;
    ohp_chunk_desc* descs; // EMIT: this stream's first descriptor
    ohp_chunk_info* info;  // EMIT, may be null
};

// MsgAudio::Split (+ SplitCompleted): m keeps the first aJiffies, rest gets what follows.
OHP_HD uint32_t msg_split(Msg& m, uint32_t aJiffies, Msg& rest, uint32_t jps)
{
    if (!(aJiffies > 0)) return kErrAssert;
    if (!(aJiffies < m.size)) return kErrAssert;
    rest.cell = m.cell;
    rest.atten = m.atten;
    rest.silence = m.silence;
    rest.total = 0;
    rest.offset = m.offset + aJiffies;
    rest.size = m.size - aJiffies;
    if (m.ramp.enabled) {
        if (core::ramp_split(m.ramp, aJiffies, m.size, rest.ramp) != 0) return kErrAssert;
    }
    else {
        core::ramp_reset(rest.ramp);
    }
    m.size = aJiffies;
    if (m.silence) {
        // silence only exists in whole samples: the first part gives its sub-sample remainder to the second
        const uint32_t spare = m.size % jps;
        m.size -= spare;
        m.total = m.size;
        rest.size += spare;
        rest.total = rest.size - rest.size % jps;
    }
    return kOk;
}

// MsgAudio::SetRamp.  Returns the new "current" ramp value in aCurrent; aHaveSplit says whether aSplit was produced.
OHP_HD uint32_t msg_set_ramp(Msg& m, uint32_t aStart, uint32_t& aRemaining, uint32_t aDirection,
                                                 Msg& aSplit, bool& aHaveSplit, uint32_t& aCurrent, uint32_t jps)
{
    const uint32_t duration = aRemaining;
    aHaveSplit = false;
    if (m.ramp.enabled && m.ramp.direction == core::kDirMute) {
        if (aDirection == core::kDirDown) aRemaining = 0;
        aCurrent = m.ramp.end;
        return kOk;
    }
    core::RampPod second;
    uint32_t splitPos;
    const int rc = core::ramp_set(m.ramp, aStart, m.size, duration, aDirection, second, splitPos);
    if (rc == core::kRampAssert) return kErrAssert;
    if (rc == 1) {
        if (splitPos == 0) {
            m.ramp = second;
        }
        else if (splitPos != m.size) {
            const core::RampPod first = m.ramp; // Split() rescales the ramp; the values Set() chose are the ones to keep
            const uint32_t e = msg_split(m, splitPos, aSplit, jps);
            if (e != kOk) return e;
            m.ramp = first;
            aSplit.ramp = second;
            aHaveSplit = true;
        }
    }
    aRemaining -= m.size;
    if (aHaveSplit && aSplit.ramp.direction != aDirection && aDirection == core::kDirUp) {
        aRemaining += aSplit.size; // the split part runs against the requested ramp (Msg.cpp:2031-2034)
    }
    if ((aDirection == core::kDirDown && m.ramp.end == core::kRampMin) || (aDirection == core::kDirUp && m.ramp.end == core::kRampMax)) {
        aRemaining = 0; // finished early (Msg.cpp:2037-2043)
    }
    aCurrent = m.ramp.end;
    return kOk;
}

OHP_HD Playable create_playable(const Msg& m, const StreamCtx& cx)
{
    Playable p;
    p.jiffies = m.size;
    if (!m.silence) {
        // offset and size are each rounded DOWN to a sample boundary, the size first being extended by whatever the
        // offset lost, so no audio is dropped between adjacent splits
        uint32_t offsetJiffies = m.offset;
        const uint32_t offsetBytes = core::jiffies_to_bytes(offsetJiffies, cx.jps, cx.channels, cx.bits);
        uint32_t sizeJiffies = m.size + (m.offset - offsetJiffies);
        p.size = core::jiffies_to_bytes(sizeJiffies, cx.jps, cx.channels, cx.bits);
        if (m.ramp.direction != core::kDirMute) {
            p.silence = 0;
            p.arena = m.cell + offsetBytes;
            p.ramp = m.ramp;
            p.atten = m.atten;
        }
        else {
            // muted audio is replaced by silence and its ramp dropped (Msg.cpp:2252-2257)
            p.silence = 1;
            p.arena = 0;
            core::ramp_reset(p.ramp);
            p.atten = OHP_UNITY_ATTENUATION;
        }
    }
    else {
        uint32_t total = m.total;
        p.size = core::jiffies_to_bytes(total, cx.jps, cx.channels, cx.bits);
        p.silence = 1;
        p.arena = 0;
        p.ramp = m.ramp; // travels with the playable but is never applied to silence
        p.atten = OHP_UNITY_ATTENUATION;
    }
    return p;
}

// MsgPlayable::Split: p keeps the first aBytes (< p.size), rest gets what follows.
OHP_HD uint32_t playable_split(Playable& p, uint32_t aBytes, Playable& rest, const StreamCtx& cx)
{
    if (!(aBytes <= p.size) || aBytes == 0) return kErrAssert;
    const uint32_t frames = aBytes / cx.frameBytes;
    const uint32_t splitJiffies = frames * cx.jps;
    rest.silence = p.silence;
    rest.arena = p.silence ? 0 : p.arena + aBytes;
    rest.size = p.size - aBytes;
    rest.jiffies = p.jiffies - splitJiffies;
    rest.atten = OHP_UNITY_ATTENUATION; // reference quirk: SplitCompleted does not pass iAttenuation on (Msg.cpp:2803-2807)
    if (p.ramp.enabled) {
        if (core::ramp_split(p.ramp, aBytes, p.size, rest.ramp) != 0) return kErrAssert;
    }
    else {
        core::ramp_reset(rest.ramp);
    }
    p.size = aBytes;
    p.jiffies = splitJiffies;
    return kOk;
}

template <bool EMIT>
OHP_HD void on_playable(StreamCtx& cx, const Playable& p)
{
    if (EMIT) {
        const uint32_t flags = (p.ramp.enabled ? OHP_F_RAMP_ENABLED : 0u) | (p.silence ? OHP_F_SILENCE : 0u)
                             | ((!p.silence && cx.in_le) ? OHP_F_IN_LITTLE_ENDIAN : 0u);
        const uint64_t src = p.silence ? 0 : p.arena;
        const uint64_t dst = cx.dst_base + cx.outBytes;
        // one 32-byte descriptor (layout: include/ohp_b200.h); two 128-bit stores on the device
        const uint32_t w4 = p.size;
        const uint32_t w5 = (p.ramp.start & 0xffffu) | (p.ramp.end << 16);
        const uint32_t w6 = (p.atten & 0xffffu) | (cx.bits << 16) | (cx.channels << 24);
        const uint32_t w7 = flags | (cx.out_fmt << 8);
#if defined(__CUDA_ARCH__)
        uint4* out = reinterpret_cast<uint4*>(cx.descs + cx.nChunks);
        out[0] = make_uint4((uint32_t)src, (uint32_t)(src >> 32), (uint32_t)dst, (uint32_t)(dst >> 32));
        out[1] = make_uint4(w4, w5, w6, w7);
        if (cx.info) {
            *reinterpret_cast<uint2*>(cx.info + cx.nChunks) = make_uint2(p.ramp.direction, p.jiffies);
        }
#else
        uint32_t* out = reinterpret_cast<uint32_t*>(cx.descs + cx.nChunks);
        out[0] = (uint32_t)src; out[1] = (uint32_t)(src >> 32); out[2] = (uint32_t)dst; out[3] = (uint32_t)(dst >> 32);
        out[4] = w4; out[5] = w5; out[6] = w6; out[7] = w7;
        if (cx.info) {
            cx.info[cx.nChunks].direction = p.ramp.direction;
            cx.info[cx.nChunks].jiffies = p.jiffies;
        }
#endif
    }
    cx.nChunks++;
    cx.outBytes += p.size;
}

// PreDriver -> CreatePlayable, then a driver that pulls fixed blocks (stage_chain.h Drive()).
template <bool EMIT>
OHP_HD uint32_t drive(StreamCtx& cx, const Msg& m)
{
    Playable p = create_playable(m, cx);
    const uint32_t block = cx.blockBytes;
    if (block == 0 || p.size == 0) {
        on_playable<EMIT>(cx, p);
        return kOk;
    }
    for (;;) {
        const uint32_t room = block - cx.blockFill;
        if (p.size > room) {
            Playable rest;
            const uint32_t e = playable_split(p, room, rest, cx);
            if (e != kOk) return e;
            on_playable<EMIT>(cx, p);
            cx.blockFill = 0;
            p = rest;
        }
        else {
            on_playable<EMIT>(cx, p);
            cx.blockFill += p.size;
            if (cx.blockFill == block) cx.blockFill = 0;
            return kOk;
        }
    }
}

OHP_HD bool next_stage_event(const StreamCtx& cx, Stage& s, uint32_t stage, uint32_t& idx)
{
    while (s.nextEv < cx.nEv) {
        const ohp_ramp_event& e = cx.ev[s.nextEv];
        if (e.stage == stage && e.op != OHP_EV_INSERT_SILENCE) {
            idx = s.nextEv;
            return true;
        }
        s.nextEv++;
    }
    return false;
}

OHP_HD void apply_event(Stage& s, uint32_t op, uint32_t arg)
{
    switch (op) {
    case OHP_EV_RAMP_DOWN:
        if (s.mode == Muted || s.current == core::kRampMin) { s.mode = Muted; s.current = core::kRampMin; s.remaining = 0; }
        else { s.mode = RampingDown; s.remaining = arg; }
        break;
    case OHP_EV_RAMP_UP:
        if (s.mode == Running && s.current == core::kRampMax) { /* already at full level */ }
        else { s.mode = RampingUp; s.remaining = arg; }
        break;
    case OHP_EV_MUTE: s.mode = Muted; s.current = core::kRampMin; s.remaining = 0; break;
    case OHP_EV_UNMUTE: s.mode = Running; s.current = core::kRampMax; s.remaining = 0; break;
    case OHP_EV_SET_ATTENUATION: s.attenuation = arg; break;
    case OHP_EV_MAX_MSG_JIFFIES: s.maxMsg = arg; break;
    default: break;
    }
}

// One stream's walk.  The per-stage queues of stage_chain.h only ever grow at the front while a stage is being served
// and are drained before the stage returns, so each is a stack; Feed()'s recursion becomes "serve the deepest
// non-empty stage".
template <bool EMIT>
OHP_HD uint32_t walk_stream(const ohp_stream_spec& sp, StreamCtx& cx)
{
    Packed stack[kStages][kStackDepth];
    Stage st[kStages];
    for (int i = 0; i < kStages; i++) {
        st[i].pos = 0; st[i].mode = Running; st[i].current = core::kRampMax; st[i].remaining = 0; st[i].maxMsg = 0;
        st[i].attenuation = OHP_UNITY_ATTENUATION; st[i].nextEv = 0; st[i].depth = 0;
    }
    if (cx.jps == 0 || cx.frameBytes == 0) return kErrSpec;
    if (sp.chunk_frames == 0 || sp.chunk_frames * cx.frameBytes > OHP_MAX_PCM_CHUNK_BYTES) return kErrSpec;

    auto push = [&](int stage, const Msg& m) -> bool {
        if (st[stage].depth == (uint32_t)kStackDepth) return false;
        stack[stage][st[stage].depth++] = pack(m);
        return true;
    };

    // stage_chain.h Process()
    auto process = [&](int stage, Msg& msg) -> uint32_t {
        Stage& s = st[stage];
        uint32_t ei;
        while (next_stage_event(cx, s, (uint32_t)stage, ei) && cx.ev[ei].at_jiffies <= s.pos) {
            apply_event(s, cx.ev[ei].op, cx.ev[ei].arg);
            s.nextEv = ei + 1;
        }
        Msg rest;
        if (next_stage_event(cx, s, (uint32_t)stage, ei) && cx.ev[ei].at_jiffies < s.pos + msg.size) {
            uint32_t at = (uint32_t)(cx.ev[ei].at_jiffies - s.pos);
            if (msg.silence) at -= at % cx.jps; // silence only splits on sample blocks
            if (at == 0) {
                apply_event(s, cx.ev[ei].op, cx.ev[ei].arg);
                s.nextEv = ei + 1;
            }
            else {
                const uint32_t e = msg_split(msg, at, rest, cx.jps);
                if (e != kOk) return e;
                if (!push(stage, rest)) return kErrDepth;
            }
        }
        if (s.maxMsg != 0 && msg.size > s.maxMsg) {
            if (s.maxMsg < cx.jps) return kErrSpec;
            const uint32_t e = msg_split(msg, s.maxMsg, rest, cx.jps);
            if (e != kOk) return e;
            if (!push(stage, rest)) return kErrDepth;
        }
        if (!msg.silence && s.attenuation != OHP_UNITY_ATTENUATION) {
            msg.atten = s.attenuation;
        }
        if (s.mode == RampingDown || s.mode == RampingUp) {
            if (s.remaining > 0) {
                if (msg.size > s.remaining) {
                    const uint32_t e = msg_split(msg, s.remaining, rest, cx.jps);
                    if (e != kOk) return e;
                    if (!push(stage, rest)) return kErrDepth;
                    if (msg.size == 0) return kErrAssert; // a MsgSilence split below one sample (see stage_chain.h)
                }
                bool haveSplit;
                const uint32_t e = msg_set_ramp(msg, s.current, s.remaining, s.mode == RampingDown ? core::kDirDown : core::kDirUp,
                                                rest, haveSplit, s.current, cx.jps);
                if (e != kOk) return e;
                if (haveSplit && !push(stage, rest)) return kErrDepth;
            }
            if (s.remaining == 0) {
                if (s.mode == RampingUp) { s.mode = Running; s.current = core::kRampMax; }
                else { s.mode = Muted; s.current = core::kRampMin; }
            }
        }
        else if (s.mode == Muted) {
            core::ramp_set_muted(msg.ramp);
        }
        s.pos += msg.size;
        return kOk;
    };

    // stage_chain.h Feed(0, item), iteratively
    auto feed = [&](const Msg& first) -> uint32_t {
        if (!push(0, first)) return kErrDepth;
        for (;;) {
            int stage = -1;
            for (int i = kStages - 1; i >= 0; i--) {
                if (stage < 0 && st[i].depth != 0) stage = i;
            }
            if (stage < 0) return kOk;
            Msg m = unpack(stack[stage][--st[stage].depth]);
            const uint32_t e = process(stage, m);
            if (e != kOk) return e;
            if (stage + 1 == kStages) {
                const uint32_t e2 = drive<EMIT>(cx, m);
                if (e2 != kOk) return e2;
            }
            else if (!push(stage + 1, m)) {
                return kErrDepth;
            }
        }
    };

    // stage_chain.h Run()
    uint64_t frame = 0;
    uint64_t srcJiffies = 0;
    uint32_t silEv = 0;
    core::CodecSource source;
    core::codec_source_init(source, sp.chunk_frames, sp.codec_read_frames, cx.frameBytes, cx.jps, sp.total_frames);
    for (;;) {
        const uint32_t frames = core::codec_source_next(source);
        if (frames == 0) break;
        for (; silEv < cx.nEv; silEv++) {
            const ohp_ramp_event& e = cx.ev[silEv];
            if (e.op != OHP_EV_INSERT_SILENCE) continue;
            if (e.at_jiffies > srcJiffies) break;
            // MsgFactory::CreateMsgSilence / MsgSilence::Initialise (Msg.cpp:2547-2560)
            Msg m;
            uint32_t jiffies = e.arg;
            core::round_down_non_zero_sample_block(jiffies, cx.jps);
            m.cell = 0; m.size = jiffies; m.total = jiffies; m.offset = 0; m.atten = OHP_UNITY_ATTENUATION; m.silence = 1;
            core::ramp_reset(m.ramp);
            const uint32_t rc = feed(m);
            if (rc != kOk) return rc;
        }
        // DecodedAudio::ConstructPcm ASSERTs on the bit depth (Msg.cpp:349-366)
        if (!(cx.bits == 8 || cx.bits == 16 || cx.bits == 24 || cx.bits == 32)) return kErrAssert;
        Msg m;
        m.cell = sp.src_base + frame * cx.frameBytes;
        m.size = frames * cx.jps;
        m.total = 0; m.offset = 0; m.atten = OHP_UNITY_ATTENUATION; m.silence = 0;
        core::ramp_reset(m.ramp);
        const uint32_t rc = feed(m);
        if (rc != kOk) return rc;
        frame += frames;
        srcJiffies += (uint64_t)frames * cx.jps;
    }
    return kOk;
}

// Fill the per-stream context from a spec and walk it.  aEvents is the WHOLE events array (aNumEvents entries).
template <bool EMIT>
OHP_HD uint32_t run_stream(const ohp_stream_spec& sp, const ohp_ramp_event* aEvents, uint64_t aNumEvents,
                           ohp_chunk_desc* aDescs, ohp_chunk_info* aInfo, uint64_t& aNumChunks, uint64_t& aOutBytes)
{
    aNumChunks = 0;
    aOutBytes = 0;
    if ((uint64_t)sp.first_event + sp.num_events > aNumEvents) return kErrSpec;
    if (!(sp.out_fmt == OHP_OUT_PACKED_BE || sp.out_fmt == OHP_OUT_PACKED_LE)) return kErrSpec;
    StreamCtx cx;
    cx.ev = aEvents + sp.first_event;
    cx.nEv = sp.num_events;
    cx.jps = core::jiffies_per_sample_or_zero(sp.sample_rate);
    cx.channels = sp.channels;
    cx.bits = sp.bit_depth;
    cx.frameBytes = sp.channels * (sp.bit_depth / 8u);
    cx.in_le = sp.in_little_endian ? 1u : 0u;
    cx.out_fmt = sp.out_fmt;
    cx.blockBytes = sp.driver_block_frames * cx.frameBytes;
    cx.blockFill = 0;
    cx.dst_base = sp.dst_base;
    cx.nChunks = 0;
    cx.outBytes = 0;
    cx.descs = EMIT ? aDescs : nullptr;
    cx.info = EMIT ? aInfo : nullptr;
    const uint32_t rc = walk_stream<EMIT>(sp, cx);
    if (rc != kOk) return rc;
    aNumChunks = cx.nChunks;
    aOutBytes = cx.outBytes;
    return kOk;
}

} // namespace sched
} // namespace ohp
