// msg_model.cpp -- see msg_model.h.  The arithmetic follows the reference exactly (descriptor parity is
// checked against the linked reference in tests/test_schedule_parity.py); the code is this repo's own.
#include "msg_model.h"
#include "ramp_core.h"

#include <algorithm>

namespace ohp {
namespace media {

static const uint16_t kRampTable[OHP_RAMP_TABLE_ENTRIES] = {
#include "../csrc/ramp_table.inc"
};

// MsgAudio::MedianRampMultiplier + RampApplicator::MedianMultiplier (Msg.cpp:2063-2074, 901-920).  The reference
// indexes kRampArray[512] (out of bounds) when the median ramp is 0; this returns 0 there.
static uint32_t ohp_median_multiplier_host(uint32_t aStart, uint32_t aEnd, uint32_t aDirection, int aEnabled)
{
    if (!aEnabled) return 0x8000u;
    if (aDirection == Ramp::EMute) return 0;
    uint32_t med = aStart;
    if (aDirection == Ramp::EUp) med = aStart + ((aEnd - aStart) / 2);
    else if (aDirection == Ramp::EDown) med = aStart - ((aStart - aEnd) / 2);
    const uint32_t idx = (Ramp::kMax - Ramp::kMin - med + (1u << 4)) >> 5;
    return idx < OHP_RAMP_TABLE_ENTRIES ? kRampTable[idx] : 0u;
}

// ------------------------------------------------------------------------------------------------
// Ramp

// The arithmetic lives in ramp_core.h (shared with the device-side schedule builder); these wrappers keep the
// reference's signatures and turn its ASSERTs into AssertionFailed.

static core::RampPod ToPod(const Ramp& aRamp)
{
    const ohp_ramp r = aRamp.ToAbi();
    return core::RampPod{r.start, r.end, r.direction, r.enabled};
}

static Ramp FromPod(const core::RampPod& aPod)
{
    return Ramp::FromAbi(ohp_ramp{aPod.start, aPod.end, aPod.direction, aPod.enabled});
}

bool Ramp::Set(uint32_t aStart, uint32_t aFragmentSize, uint32_t aRemainingDuration, EDirection aDirection,
               Ramp& aSplit, uint32_t& aSplitPos)
{
    core::RampPod r = ToPod(*this), split;
    const int rc = core::ramp_set(r, aStart, aFragmentSize, aRemainingDuration, (uint32_t)aDirection, split, aSplitPos);
    OHP_ASSERT(rc != core::kRampAssert); // Msg.cpp:598-599, 611, 620, 701-708
    *this = FromPod(r);
    aSplit = FromPod(split);
    return rc == 1;
}

Ramp Ramp::Split(uint32_t aNewSize, uint32_t aCurrentSize)
{
    core::RampPod r = ToPod(*this), rest;
    const int rc = core::ramp_split(r, aNewSize, aCurrentSize, rest);
    OHP_ASSERT(rc != core::kRampAssert);
    *this = FromPod(r);
    return FromPod(rest);
}

// ------------------------------------------------------------------------------------------------
// Msg / factory plumbing

void Msg::RemoveRef()
{
    OHP_ASSERT(iRefCount != 0);
    if (--iRefCount == 0) {
        Recycle();
    }
}

template <class T> T* MsgFactory::Take(Msg*& aFreeList)
{
    T* m;
    if (aFreeList != nullptr) {
        m = static_cast<T*>(aFreeList);
        aFreeList = m->iNextFree;
    }
    else {
        m = new T(*this);
    }
    m->iNextFree = nullptr;
    m->iRefCount = 1;
    return m;
}

static void FreeList(Msg* aHead, Msg* (*aNext)(Msg*), void (*aDelete)(Msg*))
{
    while (aHead != nullptr) {
        Msg* next = aNext(aHead);
        aDelete(aHead);
        aHead = next;
    }
}

MsgFactory::~MsgFactory()
{
    auto next = [](Msg* m) -> Msg* { return m->iNextFree; };
    FreeList(iFreePcm, next, [](Msg* m) { delete static_cast<MsgAudioPcm*>(m); });
    FreeList(iFreeSilence, next, [](Msg* m) { delete static_cast<MsgSilence*>(m); });
    FreeList(iFreePlayable, next, [](Msg* m) { delete static_cast<MsgPlayable*>(m); });
}

MsgAudioPcm* MsgFactory::CreateMsgAudioPcm(const Brx& aData, uint32_t aChannels, uint32_t aSampleRate, uint32_t aBitDepth,
                                           AudioDataEndian aEndian, uint64_t aTrackOffset)
{
    OHP_ASSERT(iArena != nullptr);
    OHP_ASSERT(aData.Bytes() <= OHP_MAX_PCM_CHUNK_BYTES); // a DecodedAudio cell holds at most AudioData::kMaxBytes
    const uint64_t off = iArena->Stage(aData);
    return CreateMsgAudioPcm(off, aData.Bytes(), aChannels, aSampleRate, aBitDepth, aEndian, aTrackOffset);
}

MsgAudioPcm* MsgFactory::CreateMsgAudioPcm(uint64_t aArenaOffset, uint32_t aBytes, uint32_t aChannels, uint32_t aSampleRate,
                                           uint32_t aBitDepth, AudioDataEndian aEndian, uint64_t aTrackOffset)
{
    // DecodedAudio::ConstructPcm asserts (Msg.cpp:349-350), MsgAudioPcm::Initialise (Msg.cpp:2264-2276),
    // MsgAudioDecoded::Initialise (Msg.cpp:2155-2168)
    OHP_ASSERT((aBitDepth & 7) == 0);
    OHP_ASSERT(aBitDepth == 8 || aBitDepth == 16 || aBitDepth == 24 || aBitDepth == 32);
    const uint32_t byteDepth = aBitDepth / 8;
    OHP_ASSERT(aBytes % byteDepth == 0);
    OHP_ASSERT(aBytes <= OHP_MAX_PCM_CHUNK_BYTES);
    OHP_ASSERT(aChannels != 0);
    const uint32_t subsamples = aBytes / byteDepth;
    OHP_ASSERT(subsamples % aChannels == 0);
    const uint32_t jps = Jiffies::PerSample(aSampleRate);
    MsgAudioPcm* msg = Take<MsgAudioPcm>(iFreePcm);
    PinArena(); // dropped in MsgAudioPcm::Recycle
    msg->Initialise(aSampleRate, aBitDepth, aChannels);
    msg->iSize = (subsamples / aChannels) * jps;
    if (msg->iSize == 0) {
        msg->RemoveRef();
        OHP_ASSERT(false); // ASSERT(iSize > 0), Msg.cpp:2166
    }
    msg->iOffset = 0;
    msg->iArenaOffset = aArenaOffset;
    msg->iEndian = aEndian;
    msg->iTrackOffset = aTrackOffset;
    msg->iAttenuation = MsgAudioPcm::kUnityAttenuation;
    return msg;
}

MsgSilence* MsgFactory::CreateMsgSilence(uint32_t& aSizeJiffies, uint32_t aSampleRate, uint32_t aBitDepth, uint32_t aChannels)
{
    // MsgSilence::Initialise, Msg.cpp:2547-2560
    const uint32_t jps = Jiffies::PerSample(aSampleRate);
    MsgSilence* msg = Take<MsgSilence>(iFreeSilence);
    msg->Initialise(aSampleRate, aBitDepth, aChannels);
    msg->iSampleBlockJiffies = jps;
    Jiffies::RoundDownNonZeroSampleBlock(aSizeJiffies, jps);
    msg->iSize = aSizeJiffies;
    msg->iSizeJiffiesTotal = aSizeJiffies;
    msg->iOffset = 0;
    return msg;
}

// ------------------------------------------------------------------------------------------------
// MsgAudio

MsgAudio* MsgAudio::Split(uint32_t aJiffies)
{
    // Msg.cpp:1949-1969
    OHP_ASSERT(aJiffies > 0);
    OHP_ASSERT(aJiffies < iSize);
    MsgAudio* rest = Allocate();
    rest->iOffset = iOffset + aJiffies;
    rest->iSize = iSize - aJiffies;
    rest->iSampleRate = iSampleRate;
    rest->iBitDepth = iBitDepth;
    rest->iNumChannels = iNumChannels;
    if (iRamp.IsEnabled()) {
        try {
            rest->iRamp = iRamp.Split(aJiffies, iSize);
        }
        catch (...) {
            rest->RemoveRef();
            throw;
        }
    }
    else {
        rest->iRamp.Reset();
    }
    iSize = aJiffies;
    SplitCompleted(*rest);
    return rest;
}

uint32_t MsgAudio::SetRamp(uint32_t aStart, uint32_t& aRemainingDuration, Ramp::EDirection aDirection, MsgAudio*& aSplit)
{
    // Msg.cpp:1989-2046
    const uint32_t duration = aRemainingDuration;
    aSplit = nullptr;
    OHP_ASSERT(aDirection == Ramp::EUp || aDirection == Ramp::EDown);
    if (iRamp.IsEnabled() && iRamp.Direction() == Ramp::EMute) {
        // already silent: nothing to ramp; a ramp down is complete by definition
        if (aDirection == Ramp::EDown) {
            aRemainingDuration = 0;
        }
        return iRamp.End();
    }
    media::Ramp second;
    uint32_t splitPos;
    if (iRamp.Set(aStart, iSize, duration, aDirection, second, splitPos)) {
        if (splitPos == 0) {
            iRamp = second;
        }
        else if (splitPos != iSize) {
            const media::Ramp first = iRamp; // Split() rescales iRamp; the values Set() chose are the ones to keep
            aSplit = Split(splitPos);
            iRamp = first;
            aSplit->iRamp = second;
        }
    }
    aRemainingDuration -= iSize;
    if (aSplit != nullptr && aSplit->iRamp.Direction() != aDirection && aDirection == Ramp::EUp) {
        aRemainingDuration += aSplit->iSize; // the split part runs against the requested ramp (Msg.cpp:2031-2034)
    }
    if ((aDirection == Ramp::EDown && iRamp.End() == Ramp::kMin) || (aDirection == Ramp::EUp && iRamp.End() == Ramp::kMax)) {
        aRemainingDuration = 0; // finished early (Msg.cpp:2037-2043)
    }
    return iRamp.End();
}

uint32_t MsgAudio::MedianRampMultiplier()
{
    const ohp_ramp r = iRamp.ToAbi();
    const uint32_t mult = ohp_median_multiplier_host(r.start, r.end, r.direction, (int)r.enabled);
    if (iRamp.IsEnabled() && iRamp.Direction() != Ramp::EMute) {
        iRamp.Reset(); // Msg.cpp:2071-2072
    }
    return mult;
}

// ------------------------------------------------------------------------------------------------
// MsgAudioPcm

MsgAudio* MsgAudioPcm::Allocate()
{
    iFactory.PinArena(); // the split remainder refers to the same staged bytes
    return iFactory.Take<MsgAudioPcm>(iFactory.iFreePcm);
}

void MsgAudioPcm::SplitCompleted(MsgAudio& aRemaining)
{
    // both halves keep referring to the same audio (Msg.cpp:2279-2285, 2184-2201)
    MsgAudioPcm& rest = static_cast<MsgAudioPcm&>(aRemaining);
    rest.iArenaOffset = iArenaOffset;
    rest.iEndian = iEndian;
    rest.iTrackOffset = (iTrackOffset == kTrackOffsetInvalid) ? iTrackOffset : iTrackOffset + iSize;
    rest.iAttenuation = iAttenuation;
}

void MsgAudioPcm::Recycle()
{
    iFactory.UnpinArena();
    iNextFree = iFactory.iFreePcm;
    iFactory.iFreePcm = this;
}

MsgPlayable* MsgAudioPcm::CreatePlayable()
{
    // Msg.cpp:2234-2262: offset and size are each rounded DOWN to a sample boundary, the size first being
    // extended by whatever the offset lost, so no audio is dropped between adjacent splits.
    const uint32_t jps = Jiffies::PerSample(iSampleRate);
    uint32_t offsetJiffies = iOffset;
    const uint32_t offsetBytes = Jiffies::ToBytes(offsetJiffies, jps, iNumChannels, iBitDepth);
    uint32_t sizeJiffies = iSize + (iOffset - offsetJiffies);
    const uint32_t sizeBytes = Jiffies::ToBytes(sizeJiffies, jps, iNumChannels, iBitDepth);
    MsgPlayable* p = iFactory.TakePlayable();
    p->iSize = sizeBytes;
    p->iJiffies = iSize;
    p->iSampleRate = iSampleRate;
    p->iBitDepth = iBitDepth;
    p->iNumChannels = iNumChannels;
    if (iRamp.Direction() != Ramp::EMute) {
        p->iSilence = false;
        p->iArenaOffset = iArenaOffset + offsetBytes;
        p->iEndian = iEndian;
        p->iRamp = iRamp;
        p->iAttenuation = iAttenuation;
        iFactory.PinArena();
        p->iPinned = true;
    }
    else {
        // muted audio is replaced by silence and its ramp dropped (Msg.cpp:2252-2257)
        p->iSilence = true;
        p->iArenaOffset = 0;
        p->iEndian = AudioDataEndian::Big;
        p->iRamp.Reset();
        p->iAttenuation = kUnityAttenuation;
    }
    RemoveRef();
    return p;
}

// ------------------------------------------------------------------------------------------------
// MsgSilence

MsgAudio* MsgSilence::Allocate() { return iFactory.Take<MsgSilence>(iFactory.iFreeSilence); }

void MsgSilence::SplitCompleted(MsgAudio& aRemaining)
{
    // silence only exists in whole samples: the first part gives its sub-sample remainder to the second
    // (Msg.cpp:2522-2545)
    MsgSilence& rest = static_cast<MsgSilence&>(aRemaining);
    rest.iSampleBlockJiffies = iSampleBlockJiffies;
    const uint32_t spare = iSize % iSampleBlockJiffies;
    iSize -= spare;
    iSizeJiffiesTotal = iSize;
    rest.iSize += spare;
    rest.iSizeJiffiesTotal = rest.iSize - rest.iSize % iSampleBlockJiffies;
}

void MsgSilence::Recycle()
{
    iNextFree = iFactory.iFreeSilence;
    iFactory.iFreeSilence = this;
}

MsgPlayable* MsgSilence::CreatePlayable()
{
    // Msg.cpp:2472-2492.  The ramp travels with the playable but is never applied to silence.
    const uint32_t jps = Jiffies::PerSample(iSampleRate);
    uint32_t total = iSizeJiffiesTotal;
    const uint32_t bytes = Jiffies::ToBytes(total, jps, iNumChannels, iBitDepth);
    if (bytes > 0) {
        OHP_ASSERT(total % iSampleBlockJiffies == 0);
    }
    MsgPlayable* p = iFactory.TakePlayable();
    p->iSilence = true;
    p->iSize = bytes;
    p->iJiffies = iSize;
    p->iSampleRate = iSampleRate;
    p->iBitDepth = iBitDepth;
    p->iNumChannels = iNumChannels;
    p->iArenaOffset = 0;
    p->iEndian = AudioDataEndian::Big;
    p->iRamp = iRamp;
    p->iAttenuation = MsgAudioPcm::kUnityAttenuation;
    RemoveRef();
    return p;
}

// ------------------------------------------------------------------------------------------------
// MsgPlayable

MsgPlayable* MsgPlayable::Split(uint32_t aBytes)
{
    // Msg.cpp:2591-2624
    OHP_ASSERT(aBytes <= iSize);
    OHP_ASSERT(aBytes != 0);
    if (aBytes == iSize) {
        return nullptr;
    }
    const uint32_t frames = aBytes / ((iBitDepth / 8) * iNumChannels);
    const uint32_t splitJiffies = frames * Jiffies::PerSample(iSampleRate);
    MsgPlayable* rest = iFactory.TakePlayable();
    rest->iSilence = iSilence;
    rest->iArenaOffset = iSilence ? 0 : iArenaOffset + aBytes;
    rest->iEndian = iEndian;
    rest->iSize = iSize - aBytes;
    rest->iJiffies = iJiffies - splitJiffies;
    rest->iSampleRate = iSampleRate;
    rest->iBitDepth = iBitDepth;
    rest->iNumChannels = iNumChannels;
    // Reference quirk kept for parity: MsgPlayablePcm::SplitCompleted (Msg.cpp:2803-2807) passes on the audio
    // but not iAttenuation, so the second part plays at unity (what Clear() left in the pooled object).
    rest->iAttenuation = MsgAudioPcm::kUnityAttenuation;
    if (!iSilence) {
        iFactory.PinArena();
        rest->iPinned = true;
    }
    if (iRamp.IsEnabled()) {
        try {
            rest->iRamp = iRamp.Split(aBytes, iSize);
        }
        catch (...) {
            rest->RemoveRef();
            throw;
        }
    }
    else {
        rest->iRamp.Reset();
    }
    iSize = aBytes;
    iJiffies = splitJiffies;
    return rest;
}

ohp_chunk_desc MsgPlayable::Descriptor(uint64_t aDstOffset, uint32_t aOutFmt) const
{
    ohp_chunk_desc d;
    std::memset(&d, 0, sizeof d);
    d.src_off = iSilence ? 0 : iArenaOffset;
    d.dst_off = aDstOffset;
    d.bytes = iSize;
    d.ramp_start = (uint16_t)iRamp.Start();
    d.ramp_end = (uint16_t)iRamp.End();
    d.attenuation = (uint16_t)iAttenuation;
    d.bit_depth = (uint8_t)iBitDepth;
    d.channels = (uint8_t)iNumChannels;
    d.flags = (uint8_t)((iRamp.IsEnabled() ? OHP_F_RAMP_ENABLED : 0u) | (iSilence ? OHP_F_SILENCE : 0u)
                        | ((!iSilence && iEndian == AudioDataEndian::Little) ? OHP_F_IN_LITTLE_ENDIAN : 0u));
    d.out_fmt = (uint8_t)aOutFmt;
    d.aux = aOutFmt == OHP_OUT_PACKED_LE ? OHP_LE_APPEND : 0; // a reader hands each playable's fragments on in order
    return d;
}

void MsgPlayable::Read(IPcmProcessor& aProcessor)
{
    OHP_ASSERT(iFactory.iReader != nullptr); // no CPU fallback: a GPU-backed reader must be attached
    iFactory.iReader->ReadNow(*this, aProcessor);
}

void MsgPlayable::Recycle()
{
    if (iPinned) {
        iPinned = false;
        iFactory.UnpinArena();
    }
    iNextFree = iFactory.iFreePlayable;
    iFactory.iFreePlayable = this;
}

} // namespace media
} // namespace ohp
