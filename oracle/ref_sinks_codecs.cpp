// ref_sinks_codecs.cpp -- LINKED-REFERENCE ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE), second harness file.
//
// Drives reference code that oracle/ref_harness.cpp could not reach in round 1:
//   * the packed little-endian sink ProcessorPcmSwpEndianPacked (Media/Tests/TestCodecInteractiveMain.cpp:114-124,
//     540-593) and the Songcast sender's IPcmProcessor half (Av/Songcast/Sender.cpp:350-398): their text is cut out of
//     /root/reference at build time by oracle/extract_ref.py into oracle/_ref/gen/ (those files do not compile
//     stand-alone) and compiled next to this file;
//   * the container codecs CodecWav / CodecAiff / CodecAifc (Media/Codec/Wav.cpp, AiffBase.cpp, Aiff.cpp, Aifc.cpp),
//     compiled UNMODIFIED, behind a fake ICodecController that feeds them the container bytes and records what they
//     tell the pipeline (OutputDecodedStream, OutputAudioPcm).
// No reference source is copied into this file.

#include <OpenHome/Media/Pipeline/Msg.h>
#include <OpenHome/Media/Utils/ProcessorAudioUtils.h>
#include <OpenHome/Media/Codec/CodecController.h>
#include <OpenHome/Media/Codec/CodecFactory.h>
#include <OpenHome/Media/Codec/Container.h>
#include <OpenHome/Media/MimeTypeList.h>

#include "extracted_p2.h"
#include "extracted_sender.h"

#include <csetjmp>
#include <csignal>
#include <cstring>
#include <vector>

#include "ohp_oracle.h"
#include "../include/ohp_container.h"

using namespace OpenHome;
using namespace OpenHome::Media;

namespace {

class NullInfo : public IInfoAggregator
{
public:
    void Register(IInfoProvider&, std::vector<Brn>&) override {}
};

struct Factory
{
    NullInfo info;
    MsgFactory* factory;
    Factory()
    {
        MsgFactoryInitParams p;
        p.SetMsgAudioPcmCount(16, 16);
        p.SetMsgSilenceCount(16);
        p.SetMsgPlayableCount(16, 1, 16);
        factory = new MsgFactory(info, p);
        // see RefFactory in ref_harness.cpp: a never-used MsgPlayablePcm carries an uninitialised iAttenuation
        static const TByte frame[4] = {0, 0, 0, 0};
        MsgPlayable* warm[16];
        for (TUint i = 0; i < 16; i++) warm[i] = factory->CreateMsgAudioPcm(Brn(frame, 4), 2, 44100, 16, AudioDataEndian::Big, 0)->CreatePlayable();
        for (TUint i = 0; i < 16; i++) warm[i]->RemoveRef();
    }
    ~Factory() { delete factory; }
};

MsgPlayable* PlayableFor(Factory& f, const ohp_chunk_desc& d, const uint8_t* in)
{
    const TUint rate = 192000;
    const uint32_t frameBytes = d.channels * (d.bit_depth / 8u);
    if (d.flags & OHP_F_SILENCE) {
        TUint jiffies = (d.bytes / frameBytes) * Jiffies::PerSample(rate);
        MsgSilence* msg = f.factory->CreateMsgSilence(jiffies, rate, d.bit_depth, d.channels);
        msg->iRamp.iStart = d.ramp_start;
        msg->iRamp.iEnd = d.ramp_end;
        msg->iRamp.iEnabled = (d.flags & OHP_F_RAMP_ENABLED) != 0;
        return msg->CreatePlayable();
    }
    MsgAudioPcm* msg = f.factory->CreateMsgAudioPcm(Brn(in + d.src_off, d.bytes), d.channels, rate, d.bit_depth,
            (d.flags & OHP_F_IN_LITTLE_ENDIAN) ? AudioDataEndian::Little : AudioDataEndian::Big, 0);
    msg->iRamp.iStart = d.ramp_start;
    msg->iRamp.iEnd = d.ramp_end;
    msg->iRamp.iEnabled = (d.flags & OHP_F_RAMP_ENABLED) != 0;
    msg->iRamp.iDirection = !msg->iRamp.iEnabled ? Ramp::ENone
                          : (d.ramp_start < d.ramp_end ? Ramp::EUp : (d.ramp_start > d.ramp_end ? Ramp::EDown : Ramp::ENone));
    msg->SetAttenuation(d.attenuation);
    return msg->CreatePlayable();
}

} // namespace

extern "C" {

// MsgPlayable::Read for each descriptor through the sink its out_fmt names, on real messages:
//   OHP_OUT_PACKED_BE  ProcessorPcmBufTest            (linked, ProcessorAudioUtils.cpp)
//   OHP_OUT_PACKED_LE  ProcessorPcmSwpEndianPacked    (extracted)  -- what the sink HOLDS after the read: for a ramped 16/24-bit
//                      playable that is the last <= 256-byte fragment only (SwapEndianness* overwrite)
//   OHP_OUT_SONGCAST   Sender::ProcessFragment        (extracted)  -- iFirstChannelIndex = FirstChannelToSend(channels)
// out receives each chunk's bytes at dst_off, out_sizes[k] (may be NULL) how many.  Returns 0, or -(k+1) where the reference
// ASSERTs on chunk k (or the chunk does not fit / names another sink).
int64_t ref_process_chunks_sinks(const ohp_chunk_desc* descs, size_t n, const uint8_t* in, uint64_t in_bytes,
                                 uint8_t* out, uint64_t out_bytes, uint32_t* out_sizes)
{
    Factory* f = new Factory();
    int64_t rc = 0;
    for (size_t k = 0; k < n && rc == 0; k++) {
        const ohp_chunk_desc& d = descs[k];
        if (out_sizes) out_sizes[k] = 0;
        try {
            if (d.bytes == 0) continue;
            if (!(d.flags & OHP_F_SILENCE) && d.src_off + d.bytes > in_bytes) { rc = -(int64_t)(k + 1); break; }
            MsgPlayable* playable = PlayableFor(*f, d, in);
            const TByte* got = nullptr;
            TUint gotBytes = 0;
            ProcessorPcmBufTest be;
            ProcessorPcmSwpEndianPacked le;
            Av::Sender sender;
            Bwh packet(4 * 9216 + 64);
            try {
                if (d.out_fmt == OHP_OUT_PACKED_BE) {
                    playable->Read(be);
                    got = be.Ptr(); gotBytes = be.Buf().Bytes();
                }
                else if (d.out_fmt == OHP_OUT_PACKED_LE) {
                    playable->Read(le);
                    got = le.Ptr(); gotBytes = le.Buf().Bytes();
                }
                else if (d.out_fmt == OHP_OUT_SONGCAST) {
                    sender.iAudioBuf = &packet;                                     // Sender::SendPendingAudio, Sender.cpp:310-311
                    sender.iFirstChannelIndex = Av::Sender::FirstChannelToSend(d.channels); // Sender.cpp:230
                    playable->Read(sender);
                    got = packet.Ptr(); gotBytes = packet.Bytes();
                }
                else {
                    rc = -(int64_t)(k + 1);
                }
            }
            catch (AssertionFailed&) {
                playable->RemoveRef();
                throw;
            }
            if (rc == 0) {
                if (d.dst_off + gotBytes > out_bytes) rc = -(int64_t)(k + 1);
                else {
                    if (gotBytes) std::memcpy(out + d.dst_off, got, gotBytes);
                    if (out_sizes) out_sizes[k] = gotBytes;
                }
            }
            playable->RemoveRef();
        }
        catch (AssertionFailed&) {
            rc = -(int64_t)(k + 1);
            f = new Factory(); // messages in flight when the ASSERT unwound are leaked with their factory
        }
    }
    if (rc == 0) delete f;
    return rc;
}

} // extern "C"

// ---- container codecs ---------------------------------------------------------------------------------------------------

namespace {

using namespace OpenHome::Media::Codec;

class MimeSink : public IMimeTypeList
{
public:
    void Add(const TChar*) override {}
};

// What CodecController is to a codec, reduced to a byte buffer: Read() hands out the container's bytes in order, the Output
// calls record what the codec tells the pipeline.  Calls the header path never makes ASSERT.
class FakeController : public ICodecController
{
public:
    FakeController(const uint8_t* aBytes, uint64_t aLen, TUint aMaxBitDepth)
        : iBytes(aBytes), iLen(aLen), iPos(0), iMaxBitDepth(aMaxBitDepth), iStreams(0), iEndianLittle(0), iAudioBytes(0) {}
    void Rewind() { iPos = 0; }
public: // from ICodecController
    void Read(Bwx& aBuf, TUint aBytes) override
    {
        TUint n = aBytes;
        if (n > aBuf.BytesRemaining()) n = aBuf.BytesRemaining();
        if ((uint64_t)n > iLen - iPos) n = (TUint)(iLen - iPos);
        aBuf.Append(iBytes + iPos, n);
        iPos += n;
    }
    void ReadNextMsg(Bwx&) override { ASSERTS(); }
    MsgAudioEncoded* ReadNextMsg() override { ASSERTS(); return nullptr; }
    TBool Read(IWriter&, TUint64, TUint) override { ASSERTS(); return false; }
    TBool TrySeekTo(TUint, TUint64) override { return false; }
    TUint64 StreamLength() const override { return iLen; }
    TUint64 StreamPos() const override { return iPos; }
    void OutputDecodedStream(TUint aBitRate, TUint aBitDepth, TUint aSampleRate, TUint aNumChannels, const Brx&, TUint64 aLength,
                             TUint64, TBool, SpeakerProfile, TBool) override
    {
        iStreams++;
        iBitRate = aBitRate; iBitDepth = aBitDepth; iSampleRate = aSampleRate; iNumChannels = aNumChannels; iLength = aLength;
        iPosAtStream = iPos;
    }
    void OutputDecodedStreamDsd(TUint, TUint, const Brx&, TUint64, TUint64, SpeakerProfile) override { ASSERTS(); }
    TUint64 OutputAudioPcm(const Brx& aData, TUint aChannels, TUint aSampleRate, TUint aBitDepth, AudioDataEndian aEndian, TUint64) override
    {
        iEndianLittle = aEndian == AudioDataEndian::Little ? 1 : 0;
        iReads.push_back(aData.Bytes());
        iAudioBytes += aData.Bytes();
        const TUint frames = aData.Bytes() / (aChannels * (aBitDepth / 8));
        return (TUint64)frames * Jiffies::PerSample(aSampleRate);
    }
    TUint64 OutputAudioPcm(MsgAudioEncoded*, TUint, TUint, TUint, TUint64) override { ASSERTS(); return 0; }
    TUint64 OutputAudioDsd(const Brx&, TUint, TUint, TUint, TUint64, TUint) override { ASSERTS(); return 0; }
    TUint64 OutputAudioDsd(MsgAudioEncoded*, TUint, TUint, TUint, TUint64, TUint) override { ASSERTS(); return 0; }
    void OutputMetaText(const Brx&) override {}
    void OutputStreamInterrupted() override {}
    void GetAudioBuf(TByte*&, TUint&) override { ASSERTS(); }
    void OutputAudioBuf(TUint, TUint64&) override { ASSERTS(); }
    TUint MaxBitDepth() const override { return iMaxBitDepth; }
public:
    const uint8_t* iBytes; uint64_t iLen; uint64_t iPos; TUint iMaxBitDepth;
    TUint iStreams; TUint iBitRate = 0, iBitDepth = 0, iSampleRate = 0, iNumChannels = 0; TUint64 iLength = 0; uint64_t iPosAtStream = 0;
    TUint iEndianLittle; uint64_t iAudioBytes;
    std::vector<uint32_t> iReads;
};

} // namespace

namespace {
// Some headers make the reference's codecs divide by zero (a zero sample rate in an AIFF COMM chunk reaches
// AiffBase.cpp's track-length division, for one): the harness reports that instead of dying with the test process.
sigjmp_buf gFpeJump;
void OnFpe(int) { siglongjmp(gFpeJump, 1); }
} // namespace

extern "C" {

// The reference's own container codecs on bytes[0..len): Recognise() of CodecWav, CodecAifc, CodecAiff in turn (each from the
// start of the stream, as CodecController rewinds between codecs, CodecController.cpp:355-395), then StreamInitialise() and the
// first Process() of the one that accepted (WAV parses its header in Process, AIFF in StreamInitialise).  For AIFF / AIFC the
// Process() loop then runs to the end of the stream: reads[] receives the byte count of every OutputAudioPcm call (the codec's
// read sizes), *n_reads their number, and audio_bytes is their sum; WAV decodes through GetAudioBuf / MsgAudioEncoded, which
// this controller does not provide, so for WAV audio_bytes / total_frames are derived from the track length the codec reports.
// out->bit_depth_src is not observable from outside the codec and is left 0.  Returns an ohp_container_status, or 100 where
// the reference's code divides by zero (SIGFPE caught; whatever the codec held is leaked).
int ref_container_decode(const uint8_t* bytes, uint64_t len, uint32_t max_bit_depth, ohp_container_info* out,
                         uint32_t* reads, uint32_t cap, uint32_t* n_reads)
{
    std::memset(out, 0, sizeof *out);
    if (n_reads) *n_reads = 0;
    struct sigaction act, old;
    std::memset(&act, 0, sizeof act);
    act.sa_handler = OnFpe;
    sigaction(SIGFPE, &act, &old);
    struct Restore { struct sigaction* o; ~Restore() { sigaction(SIGFPE, o, nullptr); } } restore{&old};
    if (sigsetjmp(gFpeJump, 1) != 0) {
        return 100;
    }
    MimeSink mime;
    FakeController ctl(bytes, len, max_bit_depth);
    CodecBase* codecs[3] = {CodecFactory::NewWav(mime), CodecFactory::NewAifc(mime), CodecFactory::NewAiff(mime)};
    const uint32_t kinds[3] = {OHP_CONTAINER_WAV, OHP_CONTAINER_AIFC, OHP_CONTAINER_AIFF};
    int status = OHP_CONTAINER_E_UNRECOGNISED;
    EncodedStreamInfo info; // Format::Encoded
    for (int c = 0; c < 3 && status == OHP_CONTAINER_E_UNRECOGNISED; c++) {
        CodecBase* codec = codecs[c];
        codec->Construct(ctl);
        ctl.Rewind();
        TBool ok = false;
        try { ok = codec->Recognise(info); }
        catch (Exception&) { ok = false; }
        if (!ok) continue;
        ctl.Rewind();
        status = OHP_CONTAINER_OK;
        try {
            codec->StreamInitialise();
            if (ctl.iStreams == 0) codec->Process(); // CodecWav: header + MsgDecodedStream on the first Process()
        }
        catch (CodecStreamEnded&) { status = OHP_CONTAINER_E_ENDED; }
        catch (CodecStreamCorrupt&) { status = OHP_CONTAINER_E_CORRUPT; }
        catch (CodecStreamFeatureUnsupported&) { status = OHP_CONTAINER_E_UNSUPPORTED; }
        catch (AssertionFailed&) { status = OHP_CONTAINER_E_ARG; }
        if (status != OHP_CONTAINER_OK) break;
        out->kind = kinds[c];
        out->sample_rate = ctl.iSampleRate;
        out->bit_depth = ctl.iBitDepth;
        out->channels = ctl.iNumChannels;
        out->bit_rate = ctl.iBitRate;
        out->data_offset = ctl.iPosAtStream;
        out->track_length_jiffies = ctl.iLength;
        if (kinds[c] == OHP_CONTAINER_WAV) {
            out->little_endian = 1; // CodecWav::WriteSamples, Wav.cpp:365-427 swaps every subsample: the stored order is little-endian
            out->total_frames = ctl.iLength * ctl.iSampleRate / Jiffies::kPerSecond;
        }
        else {
            try {
                for (;;) codec->Process();
            }
            catch (CodecStreamEnded&) {}
            catch (Exception&) {
                // e.g. SampleRateInvalid from the first message of a rate the pipeline does not play: the header itself was
                // accepted (CodecController drops such a stream when the exception reaches its thread, CodecController.cpp:481)
                ctl.iReads.clear();
                ctl.iAudioBytes = 0;
            }
            out->little_endian = ctl.iEndianLittle;
            out->audio_bytes = ctl.iAudioBytes;
            const uint32_t frameBytes = ctl.iNumChannels * (ctl.iBitDepth / 8);
            out->total_frames = frameBytes ? ctl.iAudioBytes / frameBytes : 0;
            if (n_reads) *n_reads = (uint32_t)ctl.iReads.size();
            for (size_t i = 0; i < ctl.iReads.size() && i < cap; i++) reads[i] = ctl.iReads[i];
        }
    }
    for (CodecBase* c : codecs) delete c;
    return status;
}

} // extern "C"
