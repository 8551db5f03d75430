// msg_model.h -- host-side mirror of the slice of ohPipeline's message model that produces and carries
// ramp descriptors for the decoded-PCM hot path.  Names, argument meaning and error behaviour follow
// OpenHome/Media/Pipeline/Msg.h so that reference-style code (and restated reference tests) compile
// against it; the implementation is this repo's own and differs where a GPU batch backend needs it to:
//
//   * a MsgAudioPcm does not own a 9216-byte pooled cell: it records WHERE its bytes live in the batch's
//     input arena (wire format, not yet byte-swapped -- the LE->BE unpack of DecodedAudio::ConstructPcm is
//     fused into the GPU kernel);
//   * MsgPlayable::Read() does not walk samples: a playable is a 32-byte ohp_chunk_desc, and reading is
//     done for a whole batch by BatchPcmReader (batch_reader.h) through the C ABI in include/ohp_b200.h;
//   * attenuation is never applied in place to shared audio (the reference mutates the shared cell,
//     Msg.cpp:2736-2751), each chunk is attenuated once on its way through the kernel.
//
//   reference class / function                         here
//   Ramp (Msg.h:253-286, Msg.cpp:568-807)              ohp::media::Ramp
//   Jiffies (Msg.h:186-240, Msg.cpp:411-514)           ohp::media::Jiffies
//   MsgAudio (Msg.h:866-899, Msg.cpp:1938-2080)        ohp::media::MsgAudio
//   MsgAudioPcm (Msg.h:929-958, Msg.cpp:2216-2305)     ohp::media::MsgAudioPcm
//   MsgSilence (Msg.h:1004-1030, Msg.cpp:2466-2560)    ohp::media::MsgSilence
//   MsgPlayable{,Pcm,Silence} (Msg.h:1035-1160)        ohp::media::MsgPlayable
//   IPcmProcessor (Msg.h:1204-1240)                    ohp::media::IPcmProcessor
//   MsgFactory::CreateMsgAudioPcm/CreateMsgSilence     ohp::media::MsgFactory
//   ASSERT -> AssertionFailed                          ohp::AssertionFailed
#pragma once

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "../../include/ohp_b200.h"
#include "../../include/ohp_schedule.h"

namespace ohp {

// The reference's ASSERT() raises AssertionFailed through ohNet's assert handler; tests rely on it
// (TEST_THROWS(..., AssertionFailed), Media/Tests/TestMsg.cpp:833-835).
class AssertionFailed : public std::logic_error
{
public:
    AssertionFailed(const char* aFile, int aLine)
        : std::logic_error(std::string("assertion failed at ") + aFile + ":" + std::to_string(aLine)) {}
};
class SampleRateInvalid : public std::invalid_argument
{
public:
    SampleRateInvalid() : std::invalid_argument("sample rate not supported") {}
};

#define OHP_ASSERT(x) do { if (!(x)) throw ::ohp::AssertionFailed(__FILE__, __LINE__); } while (0)

// Non-owning byte range, standing in for ohNet's Brx/Brn in signatures.
struct Brn
{
    const uint8_t* iPtr;
    uint32_t iBytes;
    Brn() : iPtr(nullptr), iBytes(0) {}
    Brn(const uint8_t* aPtr, uint32_t aBytes) : iPtr(aPtr), iBytes(aBytes) {}
    const uint8_t* Ptr() const { return iPtr; }
    uint32_t Bytes() const { return iBytes; }
};
typedef Brn Brx;

namespace media {

enum class AudioDataEndian { Invalid, Little, Big };

class Jiffies
{
public:
    static const uint32_t kPerSecond = OHP_JIFFIES_PER_SECOND;
    static const uint32_t kPerMs = kPerSecond / 1000;
    // throws SampleRateInvalid like the reference (Msg.cpp:424-470)
    static uint32_t PerSample(uint32_t aSampleRate)
    {
        const uint32_t j = PerSampleOrZero(aSampleRate);
        if (j == 0) throw SampleRateInvalid();
        return j;
    }
    static uint32_t PerSampleOrZero(uint32_t aSampleRate)
    {
        static const uint32_t kRates[] = {7350, 8000, 11025, 12000, 14700, 16000, 22050, 24000, 29400, 32000,
                                          44100, 48000, 88200, 96000, 176400, 192000, 352800, 384000};
        for (uint32_t r : kRates) {
            if (r == aSampleRate) return kPerSecond / r;
        }
        return 0;
    }
    static bool IsValidSampleRate(uint32_t aSampleRate) { return PerSampleOrZero(aSampleRate) != 0; }
    // rounds aJiffies down to a whole sample, returns the byte count (Msg.cpp:476-489)
    static uint32_t ToBytes(uint32_t& aJiffies, uint32_t aJiffiesPerSample, uint32_t aNumChannels, uint32_t aBitsPerSubsample)
    {
        aJiffies -= aJiffies % aJiffiesPerSample;
        const uint32_t subsamples = (aJiffies / aJiffiesPerSample) * aNumChannels;
        return (subsamples * aBitsPerSubsample + 7) / 8;
    }
    static void RoundDown(uint32_t& aJiffies, uint32_t aSampleRate) { aJiffies -= aJiffies % PerSample(aSampleRate); }
    static void RoundUp(uint32_t& aJiffies, uint32_t aSampleRate)
    {
        const uint32_t jps = PerSample(aSampleRate);
        aJiffies += jps - 1;
        aJiffies -= aJiffies % jps;
    }
    // never returns 0: a request below one block rounds UP to one block (Msg.cpp:504-514)
    static void RoundDownNonZeroSampleBlock(uint32_t& aJiffies, uint32_t aSampleBlockJiffies)
    {
        uint32_t j = aJiffies - aJiffies % aSampleBlockJiffies;
        if (j == 0) {
            j = aJiffies + aSampleBlockJiffies - 1;
            j -= j % aSampleBlockJiffies;
        }
        aJiffies = j;
    }
    static uint32_t ToSamples(uint32_t aJiffies, uint32_t aSampleRate) { return aJiffies / PerSample(aSampleRate); }
    static uint32_t ToMs(uint32_t aJiffies) { return aJiffies / kPerMs; }
};

class Ramp
{
public:
    static const uint32_t kMax = OHP_RAMP_MAX;
    static const uint32_t kMin = OHP_RAMP_MIN;
    enum EDirection { ENone, EUp, EDown, EMute };

    Ramp() { Reset(); }
    void Reset()
    {
        iStart = iEnd = kMax;
        iDirection = ENone;
        iEnabled = false;
    }
    // Returns true iff aSplit was set: the existing and the requested ramp run in opposite directions and cross
    // inside this fragment, which then has to be split at aSplitPos (Msg.cpp:590-712).
    bool Set(uint32_t aStart, uint32_t aFragmentSize, uint32_t aRemainingDuration, EDirection aDirection, Ramp& aSplit, uint32_t& aSplitPos);
    void SetMuted()
    {
        iStart = iEnd = kMin;
        iDirection = EMute;
        iEnabled = true;
    }
    // *this keeps the first aNewSize of aCurrentSize; the remainder is returned (Msg.cpp:784-807)
    Ramp Split(uint32_t aNewSize, uint32_t aCurrentSize);
    uint32_t Start() const { return iStart; }
    uint32_t End() const { return iEnd; }
    EDirection Direction() const { return iDirection; }
    bool IsEnabled() const { return iEnabled; }
    // plain-data view used by the C ABI
    ohp_ramp ToAbi() const { return ohp_ramp{iStart, iEnd, (uint32_t)iDirection, iEnabled ? 1u : 0u}; }
    static Ramp FromAbi(const ohp_ramp& aRamp)
    {
        Ramp r;
        r.iStart = aRamp.start;
        r.iEnd = aRamp.end;
        r.iDirection = (EDirection)aRamp.direction;
        r.iEnabled = aRamp.enabled != 0;
        return r;
    }
private:
    uint32_t iStart;
    uint32_t iEnd;
    EDirection iDirection;
    bool iEnabled;
};

class MsgFactory;
class MsgPlayable;

// Intrusively ref-counted, recycled through its factory's free lists.
class Msg
{
public:
    void AddRef() { iRefCount++; }
    void RemoveRef();
protected:
    explicit Msg(MsgFactory& aFactory) : iFactory(aFactory), iRefCount(0), iNextFree(nullptr) {}
    virtual ~Msg() {}
    virtual void Recycle() = 0;
protected:
    friend class MsgFactory;
    MsgFactory& iFactory;
    uint32_t iRefCount;
    Msg* iNextFree; // free-list link while recycled
};

class MsgAudio : public Msg
{
    friend class MsgFactory;
public:
    MsgAudio* Split(uint32_t aJiffies); // returns block after aJiffies
    virtual MsgPlayable* CreatePlayable() = 0; // consumes this msg's reference
    uint32_t Jiffies() const { return iSize; }
    // returns iRamp.End(); may hand back a second message in aSplit (Msg.cpp:1989-2046)
    uint32_t SetRamp(uint32_t aStart, uint32_t& aRemainingDuration, Ramp::EDirection aDirection, MsgAudio*& aSplit);
    void ClearRamp() { iRamp.Reset(); }
    void SetMuted() { iRamp.SetMuted(); }
    const media::Ramp& Ramp() const { return iRamp; }
    // 0x8000 = unity.  Clears the ramp (Msg.cpp:2063-2074).
    uint32_t MedianRampMultiplier();
    uint32_t SampleRate() const { return iSampleRate; }
    uint32_t BitDepth() const { return iBitDepth; }
    uint32_t NumChannels() const { return iNumChannels; }
protected:
    explicit MsgAudio(MsgFactory& aFactory) : Msg(aFactory) {}
    void Initialise(uint32_t aSampleRate, uint32_t aBitDepth, uint32_t aChannels)
    {
        iRamp.Reset();
        iSampleRate = aSampleRate;
        iBitDepth = aBitDepth;
        iNumChannels = aChannels;
    }
private:
    virtual MsgAudio* Allocate() = 0;
    virtual void SplitCompleted(MsgAudio& aRemaining) = 0;
protected:
    uint32_t iSize;   // jiffies
    uint32_t iOffset; // jiffies
    media::Ramp iRamp;
    uint32_t iSampleRate;
    uint32_t iBitDepth;
    uint32_t iNumChannels;
};

class MsgAudioPcm : public MsgAudio
{
    friend class MsgFactory;
public:
    static const uint32_t kUnityAttenuation = OHP_UNITY_ATTENUATION;
    static const uint64_t kTrackOffsetInvalid = UINT64_MAX;
    void SetAttenuation(uint32_t aAttenuation) { iAttenuation = aAttenuation; }
    MsgPlayable* CreatePlayable() override;
    uint64_t TrackOffset() const { return iTrackOffset; }
private:
    explicit MsgAudioPcm(MsgFactory& aFactory) : MsgAudio(aFactory) {}
    MsgAudio* Allocate() override;
    void SplitCompleted(MsgAudio& aRemaining) override;
    void Recycle() override;
private:
    uint64_t iArenaOffset;  // where this message's audio data ("cell") starts in the batch input arena
    AudioDataEndian iEndian; // wire order of those bytes
    uint64_t iTrackOffset;
    uint32_t iAttenuation;
};

class MsgSilence : public MsgAudio
{
    friend class MsgFactory;
public:
    MsgPlayable* CreatePlayable() override;
private:
    explicit MsgSilence(MsgFactory& aFactory) : MsgAudio(aFactory) {}
    MsgAudio* Allocate() override;
    void SplitCompleted(MsgAudio& aRemaining) override;
    void Recycle() override;
private:
    uint32_t iSampleBlockJiffies; // one sample for PCM silence
    uint32_t iSizeJiffiesTotal;
};

// Used to retrieve PCM audio data from a MsgPlayable (Msg.h:1204-1240): data is packed big endian,
// always a complete number of samples.
class IPcmProcessor
{
public:
    virtual ~IPcmProcessor() {}
    virtual void BeginBlock() = 0;
    virtual void ProcessFragment(const Brx& aData, uint32_t aNumChannels, uint32_t aSubsampleBytes) = 0;
    virtual void ProcessSilence(const Brx& aData, uint32_t aNumChannels, uint32_t aSubsampleBytes) = 0;
    virtual void EndBlock() = 0;
    virtual void Flush() = 0;
};

class MsgPlayable : public Msg
{
    friend class MsgFactory;
    friend class MsgAudioPcm;
    friend class MsgSilence;
public:
    MsgPlayable* Split(uint32_t aBytes); // returns block after aBytes; nullptr when aBytes == Bytes()
    uint32_t Bytes() const { return iSize; }
    uint32_t Jiffies() const { return iJiffies; }
    const media::Ramp& Ramp() const { return iRamp; }
    bool IsSilence() const { return iSilence; }
    uint32_t Attenuation() const { return iAttenuation; }
    // The chunk descriptor the GPU path consumes.  aDstOffset / aOutFmt are the reader's choice.
    ohp_chunk_desc Descriptor(uint64_t aDstOffset, uint32_t aOutFmt) const;
    // Drop-in for MsgPlayable::Read(IPcmProcessor&) (Msg.cpp:2646-2653): synchronous, runs this one playable
    // through the factory's batch reader.  Batch many playables with BatchPcmReader instead.
    void Read(IPcmProcessor& aProcessor);
private:
    explicit MsgPlayable(MsgFactory& aFactory) : Msg(aFactory) {}
    void Recycle() override;
private:
    bool iSilence;
    uint32_t iSize;    // bytes
    uint32_t iJiffies;
    uint32_t iSampleRate;
    uint32_t iBitDepth;
    uint32_t iNumChannels;
    uint64_t iArenaOffset; // byte offset of the first payload byte in the input arena (cell start + iOffset)
    AudioDataEndian iEndian;
    media::Ramp iRamp;
    uint32_t iAttenuation;
    bool iPinned = false;  // holds a pin on the factory's input arena (PCM playables only)
};

// Hook through which MsgPlayable::Read reaches the GPU (implemented by BatchPcmReader).
class IPlayableReader
{
public:
    virtual ~IPlayableReader() {}
    virtual void ReadNow(MsgPlayable& aPlayable, IPcmProcessor& aProcessor) = 0;
};

// Where CreateMsgAudioPcm(const Brx&, ...) puts audio: the batch's host-side input arena.
class IInputArena
{
public:
    virtual ~IInputArena() {}
    // Returns the arena offset now holding aData (copied, or located in place if it already lies inside).
    virtual uint64_t Stage(const Brx& aData) = 0;
    // Every live MsgAudioPcm / PCM MsgPlayable holds one pin on the arena (taken when the message is created or split
    // off, dropped when it is recycled): the arena must not hand the staged bytes out again while a pin is held --
    // the counterpart of the reference's ref-counted DecodedAudio cell (Msg.cpp:2260, 2732).
    virtual void Pin() {}
    virtual void Unpin() {}
};

class MsgFactory
{
    friend class Msg;
    friend class MsgAudioPcm;
    friend class MsgSilence;
    friend class MsgPlayable;
public:
    explicit MsgFactory(IInputArena* aArena = nullptr, IPlayableReader* aReader = nullptr)
        : iArena(aArena), iReader(aReader), iFreePcm(nullptr), iFreeSilence(nullptr), iFreePlayable(nullptr) {}
    ~MsgFactory();
    void SetReader(IPlayableReader* aReader) { iReader = aReader; }
    // MsgFactory::CreateMsgAudioPcm (Msg.cpp:3961-3965, 4027-4039).  aData is staged into the input arena in wire
    // format; no byte swap happens on the host.
    MsgAudioPcm* CreateMsgAudioPcm(const Brx& aData, uint32_t aChannels, uint32_t aSampleRate, uint32_t aBitDepth,
                                   AudioDataEndian aEndian, uint64_t aTrackOffset);
    // Same, for audio already resident in the arena at aArenaOffset (zero-copy).
    MsgAudioPcm* CreateMsgAudioPcm(uint64_t aArenaOffset, uint32_t aBytes, uint32_t aChannels, uint32_t aSampleRate,
                                   uint32_t aBitDepth, AudioDataEndian aEndian, uint64_t aTrackOffset);
    // MsgFactory::CreateMsgSilence (Msg.cpp:3989-3994); aSizeJiffies is rounded to whole samples in place.
    MsgSilence* CreateMsgSilence(uint32_t& aSizeJiffies, uint32_t aSampleRate, uint32_t aBitDepth, uint32_t aChannels);
private:
    template <class T> T* Take(Msg*& aFreeList);
    MsgPlayable* TakePlayable() { return Take<MsgPlayable>(iFreePlayable); }
    void PinArena() { if (iArena != nullptr) iArena->Pin(); }
    void UnpinArena() { if (iArena != nullptr) iArena->Unpin(); }
private:
    IInputArena* iArena;
    IPlayableReader* iReader;
    Msg* iFreePcm;
    Msg* iFreeSilence;
    Msg* iFreePlayable;
};

} // namespace media
} // namespace ohp
