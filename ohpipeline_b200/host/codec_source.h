// codec_source.h -- the sizes of the MsgAudioPcm messages a PCM stream enters the pipeline with, on plain data
// (host and device): what a container codec reads, what CodecController cuts it into and what DecodedAudioAggregator
// packs together again.
//
//   Wav:   CodecWav::Process asks CodecController::GetAudioBuf for min(iMaxOutputSamples, 9216 B / frame) frames and
//          fills exactly that (Media/Codec/Wav.cpp:126-189, Media/Codec/CodecController.cpp:919-939);
//   Aiff:  CodecAiffBase::Process reads 9216 - 9216 % frame bytes (Media/Codec/AiffBase.cpp:59-86) and hands them to
//          CodecController::OutputAudioPcm, which cuts them into iMaxOutputBytes pieces (CodecController.cpp:800-827);
//   both:  iMaxOutputSamples = Jiffies::ToSamples(5 ms, rate) (CodecController.cpp:792-793, Pipeline.cpp:403-405,
//          Pipeline.h:176);
//   then   DecodedAudioAggregator::TryAggregate (Media/Pipeline/DecodedAudioAggregator.cpp:134-186): a message is
//          passed on when it holds 9216 bytes or at least 5 ms - 7680 jiffies, otherwise held and merged with what
//          follows while that fits in 9216 bytes.
// A 5 ms message is always "full", so Wav's uniform messages pass unchanged; Aiff's short tail piece of every read is
// merged with the first piece of the next one.  Pinned against the real DecodedAudioAggregator (oracle/_ref,
// tests/test_container.py).
#pragma once

#include <cstdint>

#include "ramp_core.h"

namespace ohp {
namespace core {

constexpr uint32_t kAudioCellBytes = 9216u;                         // AudioData::kMaxBytes (Msg.h:117)
constexpr uint32_t kAggregatorMaxJiffies = 5u * 56448u - 7680u;     // DecodedAudioAggregator::kMaxJiffies (.h:19)

struct CodecSource
{
    uint32_t chunk;        // frames per message out of CodecController (ohp_stream_spec::chunk_frames)
    uint32_t read;         // frames per codec read (ohp_stream_spec::codec_read_frames); 0: no aggregator in the way
    uint32_t frame_bytes, jps;
    uint64_t total_left;   // frames of the stream not yet read
    uint32_t read_left;    // frames of the current read not yet cut off
    uint32_t held;         // frames DecodedAudioAggregator is holding back
};

OHP_HD void codec_source_init(CodecSource& s, uint32_t chunk, uint32_t read, uint32_t frame_bytes, uint32_t jps, uint64_t total_frames)
{
    s.chunk = chunk; s.read = read; s.frame_bytes = frame_bytes; s.jps = jps;
    s.total_left = total_frames; s.read_left = 0; s.held = 0;
}

// DecodedAudioAggregator::AggregatorFull (DecodedAudioAggregator.cpp:129-132)
OHP_HD bool aggregator_full(const CodecSource& s, uint32_t frames)
{
    return frames * s.frame_bytes == kAudioCellBytes || frames * s.jps >= kAggregatorMaxJiffies;
}

// Frames of the next message to enter the stage chain; 0 when the stream is over.
OHP_HD uint32_t codec_source_next(CodecSource& s)
{
    if (s.read == 0) {
        const uint32_t f = (uint32_t)(s.total_left < s.chunk ? s.total_left : s.chunk);
        s.total_left -= f;
        return f;
    }
    for (;;) {
        if (s.total_left == 0) {
            const uint32_t f = s.held; // OutputAggregatedAudio when the stream ends (DecodedAudioAggregator.cpp:188-194)
            s.held = 0;
            return f;
        }
        if (s.read_left == 0) s.read_left = (uint32_t)(s.total_left < s.read ? s.total_left : s.read);
        const uint32_t p = s.read_left < s.chunk ? s.read_left : s.chunk;
        s.read_left -= p;
        s.total_left -= p;
        if (s.held == 0) {
            if (aggregator_full(s, p)) return p;
            s.held = p;
        }
        else if ((s.held + p) * s.frame_bytes <= kAudioCellBytes) {
            s.held += p;
            if (aggregator_full(s, s.held)) {
                const uint32_t f = s.held;
                s.held = 0;
                return f;
            }
        }
        else {
            const uint32_t f = s.held; // the lazy branch: pass on what is held, hold the newcomer
            s.held = p;
            return f;
        }
    }
}

} // namespace core
} // namespace ohp
