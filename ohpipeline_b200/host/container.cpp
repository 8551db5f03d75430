// container.cpp -- WAV / AIFF / AIFC header parsing (include/ohp_container.h), following the reference's codecs
// decision for decision (which chunk order they accept, which sizes, what they THROW where), on a byte range instead
// of a pull interface.  No sample is touched here.
#include "../../include/ohp_container.h"
#include "codec_source.h"

#include <cstring>

namespace {

struct Reader
{
    const uint8_t* p;
    uint64_t len;
    uint64_t pos;
    // iController->Read(buf, n): fewer than n bytes left means the stream ended
    const uint8_t* Read(uint64_t n)
    {
        if (n > len - pos) { pos = len; return nullptr; }
        const uint8_t* r = p + pos;
        pos += n;
        return r;
    }
};

uint32_t Le16(const uint8_t* b) { return (uint32_t)b[0] | ((uint32_t)b[1] << 8); }
uint32_t Le32(const uint8_t* b) { return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24); }
uint32_t Be16(const uint8_t* b) { return ((uint32_t)b[0] << 8) | b[1]; }
uint32_t Be32(const uint8_t* b) { return ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3]; }

// CodecWav::FindChunk (Wav.cpp:319-353) / CodecAiffBase::FindChunk (AiffBase.cpp:112-147): skip chunks until aId;
// sizes are padded to even.  Returns the (padded) chunk size through aBytes.
int FindChunk(Reader& r, const char* aId, bool aLittleEndian, uint32_t& aBytes)
{
    for (;;) {
        const uint8_t* h = r.Read(8);
        if (h == nullptr) return OHP_CONTAINER_E_ENDED;
        uint32_t bytes = aLittleEndian ? Le32(h + 4) : Be32(h + 4);
        bytes += bytes % 2;
        if (std::memcmp(h, aId, 4) == 0) {
            aBytes = bytes;
            return OHP_CONTAINER_OK;
        }
        if (r.Read(bytes) == nullptr) return OHP_CONTAINER_E_ENDED;
    }
}

int ParseWav(Reader& r, uint32_t aMaxBitDepth, ohp_container_info& o)
{
    // ProcessRiffChunk (Wav.cpp:247-269)
    r.pos = 0;
    const uint8_t* h = r.Read(12);
    if (h == nullptr) return OHP_CONTAINER_E_ENDED;
    const uint32_t fileSize = Le32(h + 4); // zero: a continuous stream
    // ProcessFmtChunk (Wav.cpp:271-313)
    uint32_t fmtBytes;
    int rc = FindChunk(r, "fmt ", true, fmtBytes);
    if (rc != OHP_CONTAINER_OK) return rc;
    if (fmtBytes != 16 && fmtBytes != 18 && fmtBytes != 40) return OHP_CONTAINER_E_CORRUPT;
    const uint8_t* f = r.Read(fmtBytes);
    if (f == nullptr) return OHP_CONTAINER_E_ENDED;
    const uint32_t audioFormat = Le16(f);
    if (audioFormat != 0x01 && audioFormat != 0xfffe) return OHP_CONTAINER_E_UNSUPPORTED; // PCM or WAVE_FORMAT_EXTENSIBLE
    o.channels = Le16(f + 2);
    o.sample_rate = Le32(f + 4);
    o.bit_rate = Le32(f + 8) * 8u;
    o.bit_depth_src = Le16(f + 14);
    o.bit_depth = o.bit_depth_src < aMaxBitDepth ? o.bit_depth_src : aMaxBitDepth;
    const uint32_t frameBytesSrc = o.channels * (o.bit_depth_src / 8u);
    if (o.channels == 0 || o.sample_rate == 0 || o.bit_rate == 0 || o.bit_depth == 0 || o.bit_depth % 8 != 0) {
        return OHP_CONTAINER_E_CORRUPT;
    }
    // ProcessDataChunk (Wav.cpp:315-337)
    uint32_t dataBytes;
    rc = FindChunk(r, "data", true, dataBytes);
    if (rc != OHP_CONTAINER_OK) return rc;
    o.streaming = fileSize == 0 ? 1u : 0u;
    uint64_t audio = o.streaming ? 0 : dataBytes;
    if (frameBytesSrc == 0) return OHP_CONTAINER_E_CORRUPT; // the reference divides by it (8 > depth > 0 cannot pass the % 8 test)
    audio -= audio % frameBytesSrc;
    o.audio_bytes = audio;
    o.data_offset = r.pos;
    o.total_frames = audio / frameBytesSrc;
    o.track_length_jiffies = (o.total_frames * OHP_JIFFIES_PER_SECOND) / o.sample_rate;
    o.little_endian = 1;
    o.kind = OHP_CONTAINER_WAV;
    return OHP_CONTAINER_OK;
}

// CodecAiffBase::DetermineRate (AiffBase.cpp:149-186): the 80-bit extended field's exponent and top 32 mantissa bits
uint32_t DetermineRate(uint32_t aExponent, uint32_t aMantissa)
{
    // The reference shifts a 32-bit value by whatever the exponent gives, 32 and more included -- undefined in C++, and on
    // the x86 and ARM64 targets it is built for the hardware takes the count modulo 32.  A damaged exponent therefore still
    // yields a rate there (0x020e reads like 0x400e); the same here, so that both accept and refuse the same headers.
    uint32_t rate;
    if (aExponent < 0x4013) { // kUnder65kHz (AiffBase.h:38 -- despite its name the switch-over is at 2^20 Hz)
        rate = aMantissa >> ((0x401eu - aExponent) & 31u);
    }
    else {
        rate = aMantissa >> ((aExponent - 0x4007u) & 31u);
    }
    if (rate == 22255) rate = 22050;      // old Macintosh rates
    else if (rate == 11127) rate = 11025;
    return rate;
}

int ParseAiff(Reader& r, bool aAifc, ohp_container_info& o)
{
    // ProcessFormChunk (AiffBase.cpp:200-221)
    r.pos = 0;
    if (r.Read(12) == nullptr) return OHP_CONTAINER_E_ENDED;
    // GetCommChunkHeader (Aiff.cpp:44-52: exactly 18 bytes; Aifc.cpp:44-53: at least 22)
    uint32_t commBytes;
    int rc = FindChunk(r, "COMM", false, commBytes);
    if (rc != OHP_CONTAINER_OK) return rc;
    if (aAifc ? commBytes < 22 : commBytes != 18) return OHP_CONTAINER_E_CORRUPT;
    // ParseCommChunk (AiffBase.cpp:223-259)
    const uint8_t* c = r.Read(commBytes);
    if (c == nullptr) return OHP_CONTAINER_E_ENDED;
    o.channels = Be16(c);
    const uint32_t samplesTotal = Be32(c + 2);
    o.bit_depth_src = Be16(c + 6);
    const uint32_t frameBytes = o.channels * (o.bit_depth_src / 8u);
    const uint32_t audioBytesTotal = samplesTotal * frameBytes; // TUint arithmetic as in the reference
    o.sample_rate = DetermineRate(Be16(c + 8), Be32(c + 10));
    if (o.sample_rate == 0) return OHP_CONTAINER_E_CORRUPT;     // the reference divides by it
    o.track_length_jiffies = ((uint64_t)samplesTotal * OHP_JIFFIES_PER_SECOND) / o.sample_rate;
    switch (o.bit_depth_src) {
    case 8: case 16: case 24: o.bit_depth = o.bit_depth_src; break;
    case 20: o.bit_depth = 24; break;
    default: return OHP_CONTAINER_E_UNSUPPORTED;
    }
    o.bit_rate = o.sample_rate * frameBytes * 8u;
    o.little_endian = 0;
    if (aAifc) {
        // CodecAifc::ProcessCommChunkExtra (Aifc.cpp:55-69)
        if (std::memcmp(c + 18, "sowt", 4) == 0 || std::memcmp(c + 18, "SOWT", 4) == 0) o.little_endian = 1;
        else if (std::memcmp(c + 18, "NONE", 4) != 0) return OHP_CONTAINER_E_UNSUPPORTED;
    }
    // ProcessSsndChunk (AiffBase.cpp:261-276)
    uint32_t ssndBytes;
    rc = FindChunk(r, "SSND", false, ssndBytes);
    if (rc != OHP_CONTAINER_OK) return rc;
    if (audioBytesTotal > ssndBytes) return OHP_CONTAINER_E_CORRUPT;
    if (r.Read(8) == nullptr) return OHP_CONTAINER_E_ENDED; // offset and block size
    o.data_offset = r.pos;
    o.audio_bytes = audioBytesTotal;
    o.total_frames = frameBytes ? audioBytesTotal / frameBytes : 0;
    o.streaming = 0;
    o.kind = aAifc ? OHP_CONTAINER_AIFC : OHP_CONTAINER_AIFF;
    return OHP_CONTAINER_OK;
}

} // namespace

extern "C" {

int ohp_container_parse(const uint8_t* bytes, uint64_t len, uint32_t max_bit_depth, ohp_container_info* out)
{
    if (!out || (!bytes && len)) return OHP_CONTAINER_E_ARG;
    std::memset(out, 0, sizeof *out);
    if (max_bit_depth == 0) max_bit_depth = 32;
    // Recognise() of each codec looks at the first 12 bytes (Wav.cpp:87-103, AiffBase.cpp:41-52)
    if (len < 12) return OHP_CONTAINER_E_UNRECOGNISED;
    Reader r{bytes, len, 0};
    if (std::memcmp(bytes, "RIFF", 4) == 0 && std::memcmp(bytes + 8, "WAVE", 4) == 0) return ParseWav(r, max_bit_depth, *out);
    if (std::memcmp(bytes, "FORM", 4) == 0 && std::memcmp(bytes + 8, "AIFF", 4) == 0) return ParseAiff(r, false, *out);
    if (std::memcmp(bytes, "FORM", 4) == 0 && std::memcmp(bytes + 8, "AIFC", 4) == 0) return ParseAiff(r, true, *out);
    return OHP_CONTAINER_E_UNRECOGNISED;
}

int ohp_container_stream_spec(const ohp_container_info* info, uint64_t container_len, uint64_t arena_offset, uint64_t dst_base,
                              ohp_stream_spec* out)
{
    if (!info || !out) return OHP_CONTAINER_E_ARG;
    std::memset(out, 0, sizeof *out);
    if (info->bit_depth != info->bit_depth_src) return OHP_CONTAINER_E_UNSUPPORTED;
    const uint32_t jps = ohp::core::jiffies_per_sample_or_zero(info->sample_rate);
    if (jps == 0) return OHP_CONTAINER_E_UNSUPPORTED; // Jiffies::PerSample throws SampleRateInvalid
    const uint32_t frameBytes = info->channels * (info->bit_depth / 8u);
    if (frameBytes == 0 || frameBytes > OHP_MAX_PCM_CHUNK_BYTES) return OHP_CONTAINER_E_UNSUPPORTED;
    uint64_t bytes = info->audio_bytes;
    if (info->streaming || info->data_offset + bytes > container_len) {
        bytes = container_len > info->data_offset ? container_len - info->data_offset : 0; // what is actually there
    }
    out->sample_rate = info->sample_rate;
    out->bit_depth = info->bit_depth;
    out->channels = info->channels;
    out->in_little_endian = info->little_endian && info->bit_depth > 8 ? 1u : 0u;
    // iMaxOutputSamples = Jiffies::ToSamples(5 ms, rate), capped by what fits a DecodedAudio cell
    const uint32_t maxSamples = (5u * OHP_JIFFIES_PER_MS) / jps;
    const uint32_t cellSamples = OHP_MAX_PCM_CHUNK_BYTES / frameBytes;
    out->chunk_frames = maxSamples < cellSamples ? maxSamples : cellSamples;
    out->codec_read_frames = info->kind == OHP_CONTAINER_WAV ? 0u : cellSamples; // AiffBase.cpp:66
    out->out_fmt = OHP_OUT_PACKED_BE;
    out->total_frames = bytes / frameBytes;
    out->src_base = arena_offset + info->data_offset;
    out->dst_base = dst_base;
    return OHP_CONTAINER_OK;
}

size_t ohp_codec_message_frames(const ohp_stream_spec* spec, uint32_t* out, size_t cap)
{
    if (!spec) return 0;
    const uint32_t jps = ohp::core::jiffies_per_sample_or_zero(spec->sample_rate);
    const uint32_t frameBytes = spec->channels * (spec->bit_depth / 8u);
    if (jps == 0 || frameBytes == 0 || spec->chunk_frames == 0) return 0;
    ohp::core::CodecSource s;
    ohp::core::codec_source_init(s, spec->chunk_frames, spec->codec_read_frames, frameBytes, jps, spec->total_frames);
    size_t n = 0;
    for (;;) {
        const uint32_t f = ohp::core::codec_source_next(s);
        if (f == 0) break;
        if (out && n < cap) out[n] = f;
        n++;
    }
    return n;
}

} // extern "C"
