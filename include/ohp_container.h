/*
 * ohp_container.h -- C ABI of the PCM container front end: where a batch starts from WAV / AIFF / AIFC
 * bytes instead of raw PCM.  Exported by libohp_host.so (header parsing is control plane; the samples never move:
 * the container's data chunk IS the input arena, its byte order is handled inside the GPU kernel by
 * OHP_F_IN_LITTLE_ENDIAN, which is what CodecWav::WriteSamples / DecodedAudio::CopyToBigEndian* do per sample on the
 * reference's CPU path).  Citations are relative to the ohPipeline source root.
 *
 *   reference                                                        here
 *   CodecWav::Recognise / ProcessHeader (Riff, Fmt, Data, FindChunk)  ohp_container_parse      (Media/Codec/Wav.cpp:87-103, 225-353)
 *   CodecAiffBase::ProcessHeader (Form, Comm, Ssnd), DetermineRate    ohp_container_parse      (Media/Codec/AiffBase.cpp:112-281)
 *   CodecAiff / CodecAifc COMM handling ("NONE", "sowt")              ohp_container_parse      (Media/Codec/Aiff.cpp:44-52, Aifc.cpp:44-69)
 *   CodecController::GetAudioBuf / OutputAudioPcm message sizes       ohp_container_stream_spec (Media/Codec/CodecController.cpp:792-827, 919-939)
 *   DecodedAudioAggregator::TryAggregate                              ohp_stream_spec::codec_read_frames, ohp_codec_message_frames
 *                                                                     (Media/Pipeline/DecodedAudioAggregator.cpp:134-186)
 */
#ifndef OHP_CONTAINER_H
#define OHP_CONTAINER_H

#include "ohp_schedule.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ohp_container_kind {
    OHP_CONTAINER_WAV = 1,
    OHP_CONTAINER_AIFF = 2,
    OHP_CONTAINER_AIFC = 3
} ohp_container_kind;

/* what the reference's codecs THROW */
typedef enum ohp_container_status {
    OHP_CONTAINER_OK = 0,
    OHP_CONTAINER_E_UNRECOGNISED = 1, /* no codec's Recognise() accepts the first 12 bytes                      */
    OHP_CONTAINER_E_ENDED = 2,        /* CodecStreamEnded: the bytes stop inside the header                      */
    OHP_CONTAINER_E_CORRUPT = 3,      /* CodecStreamCorrupt                                                      */
    OHP_CONTAINER_E_UNSUPPORTED = 4,  /* CodecStreamFeatureUnsupported (compressed WAV, AIFC other than NONE/sowt, odd depths) */
    OHP_CONTAINER_E_ARG = 5
} ohp_container_status;

typedef struct ohp_container_info {
    uint32_t kind;              /* ohp_container_kind                                                              */
    uint32_t sample_rate;       /* AIFF: from the 80-bit extended COMM field (22255 -> 22050, 11127 -> 11025)       */
    uint32_t bit_depth_src;     /* as stored                                                                       */
    uint32_t bit_depth;         /* as output: min(src, max_bit_depth) for WAV; 20 -> 24 for AIFF                   */
    uint32_t channels;
    uint32_t little_endian;     /* byte order of the stored subsamples (WAV, AIFC "sowt": 1)                       */
    uint32_t bit_rate;
    uint32_t streaming;         /* WAV with a RIFF size of 0: continuous stream, audio_bytes unknown (0)            */
    uint64_t data_offset;       /* iTrackStart: offset of the first audio byte in the container                     */
    uint64_t audio_bytes;       /* playable audio bytes, a whole number of frames (iAudioBytesRemaining)            */
    uint64_t total_frames;
    uint64_t track_length_jiffies;
} ohp_container_info;

/* Parse the header at bytes[0..len).  max_bit_depth = the animator's limit (CodecController::MaxBitDepth; 32 = none). */
int ohp_container_parse(const uint8_t* bytes, uint64_t len, uint32_t max_bit_depth, ohp_container_info* out);

/*
 * The stream spec of a parsed container whose first byte sits at arena_offset of the input arena: src_base points at
 * the audio itself, chunk_frames = min(5 ms, 9216 B) as CodecController cuts it, codec_read_frames as the codec reads.
 * Audio present in [data_offset, len) beyond audio_bytes is ignored, audio missing (truncated file) shortens the stream.
 * Fails (OHP_CONTAINER_E_UNSUPPORTED) where output depth != stored depth: that is a re-quantising sink, not a stream.
 *
 * ASSUMPTION for WAV (and raw PCM): every message is a full chunk_frames (the last one excepted).  CodecWav::Process fills
 * the 5 ms buffer from what is left of the current MsgAudioEncoded plus ONE more (Wav.cpp:140-185), so over a transport
 * whose encoded messages are so small that two of them hold less than 5 ms of audio the reference's messages are shorter
 * and follow the transport's boundaries -- and Ramp::Set rounds per message, so a ramp over such a stretch differs in its
 * last bits.  Reads of a FIXED size can be expressed (codec_read_frames of the spec: the codec's read size in frames,
 * what AIFF's 9216-byte reads use; 0 = messages of chunk_frames); arbitrary transport boundaries cannot.
 */
int ohp_container_stream_spec(const ohp_container_info* info, uint64_t container_len, uint64_t arena_offset,
                              uint64_t dst_base, ohp_stream_spec* out);

/* Frames of every message the stream enters the pipeline with (codec reads, CodecController pieces,
 * DecodedAudioAggregator); returns their number (may exceed cap: only cap are written). */
size_t ohp_codec_message_frames(const ohp_stream_spec* spec, uint32_t* out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* OHP_CONTAINER_H */
