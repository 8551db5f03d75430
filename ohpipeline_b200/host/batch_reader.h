// batch_reader.h -- GPU-backed replacement for the loop { playable->Read(processor); playable->RemoveRef(); }
//
// In the reference a driver pulls one MsgPlayable at a time and reads it synchronously on the animator thread
// (e.g. Av/Utils/DriverSongcastSender.cpp:176-199: Pull -> Read(ProcessorPcmBufTest) -> send).  Here playables from
// many independent streams are queued, turned into chunk descriptors and processed by ONE call into the C ABI
// (ohp_process_host: H2D, the fused ramp + convert kernel, D2H); every processor then receives its audio through the
// reference's own IPcmProcessor calls: BeginBlock, ProcessFragment / ProcessSilence (packed big-endian, whole
// samples, Msg.h:1204-1240), EndBlock -- one fragment per playable instead of one per 256 bytes (Msg.cpp:2762-2779).
//
// Header-only; link against libohp_b200.so and compile msg_model.cpp alongside.  There is no CPU path: construction
// fails (throws) when no B200 is available.
#pragma once

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ohp_b200.h"
#include "msg_model.h"

namespace ohp {

class OhpError : public std::runtime_error
{
public:
    OhpError(int aStatus, const std::string& aWhat) : std::runtime_error(aWhat), iStatus(aStatus) {}
    int Status() const { return iStatus; }
private:
    int iStatus;
};

class BatchPcmReader : public media::IPlayableReader, public media::IInputArena
{
public:
    BatchPcmReader(int aDevice, uint64_t aInCapacityBytes, uint64_t aOutCapacityBytes)
        : iCtx(nullptr), iIn(nullptr), iOut(nullptr), iInCapacity(aInCapacityBytes), iOutCapacity(aOutCapacityBytes)
        , iInUsed(0), iOutUsed(0), iPins(0)
    {
        const int rc = ohp_create(aDevice, &iCtx);
        if (rc != OHP_OK) throw OhpError(rc, std::string("ohp_create: ") + ohp_last_error(nullptr));
        try {
            Check(ohp_host_alloc(iCtx, iInCapacity ? iInCapacity : 16, reinterpret_cast<void**>(&iIn)), "ohp_host_alloc");
            Check(ohp_host_alloc(iCtx, iOutCapacity ? iOutCapacity : 16, reinterpret_cast<void**>(&iOut)), "ohp_host_alloc");
        }
        catch (...) {
            Release(); // a constructor that throws runs no destructor
            throw;
        }
    }
    ~BatchPcmReader()
    {
        for (auto& q : iQueue) q.playable->RemoveRef();
        Release();
    }
    BatchPcmReader(const BatchPcmReader&) = delete;
    BatchPcmReader& operator=(const BatchPcmReader&) = delete;

    // IInputArena: MsgFactory::CreateMsgAudioPcm(const Brx&, ...) lands here.  Audio already inside the arena
    // (for instance decoded straight into Reserve()d space) is used in place.
    uint64_t Stage(const Brx& aData) override
    {
        // in place only inside what Reserve() has handed out: anything above iInUsed is not part of the batch
        if (aData.Ptr() >= iIn && aData.Ptr() + aData.Bytes() <= iIn + iInUsed) {
            return (uint64_t)(aData.Ptr() - iIn);
        }
        uint8_t* p = Reserve(aData.Bytes());
        std::memcpy(p, aData.Ptr(), aData.Bytes());
        return (uint64_t)(p - iIn);
    }
    // IInputArena: one pin per live message that refers to staged bytes (msg_model.h).  The arena is only recycled
    // while no pin is held, so a message created but not yet Add()ed, a split remainder or a playable the caller kept
    // across a Flush() still finds its audio where it was staged.
    void Pin() override { iPins++; }
    void Unpin() override { OHP_ASSERT(iPins != 0); iPins--; }
    uint64_t Pins() const { return iPins; }
    // Pinned, 16-byte aligned space inside the input arena for the caller to decode into.
    uint8_t* Reserve(uint32_t aBytes)
    {
        const uint64_t at = (iInUsed + 15u) & ~15ull;
        if (at + aBytes > iInCapacity) throw OhpError(OHP_E_NO_MEMORY, "BatchPcmReader: input arena full (Flush() first)");
        iInUsed = at + aBytes;
        return iIn + at;
    }

    // Queue a playable for the next Flush().  Takes over the caller's reference, like the driver that would have
    // called Read() then RemoveRef().  aOutFmt fuses the sink's format conversion (default: packed big-endian, what
    // every IPcmProcessor is handed by the reference).
    void Add(media::MsgPlayable* aPlayable, media::IPcmProcessor& aProcessor, uint32_t aOutFmt = OHP_OUT_PACKED_BE)
    {
        const uint64_t at = (iOutUsed + 15u) & ~15ull;
        ohp_chunk_desc d = aPlayable->Descriptor(at, aOutFmt);
        const uint32_t outBytes = ohp_chunk_out_bytes(&d);
        if (at + outBytes > iOutCapacity) throw OhpError(OHP_E_NO_MEMORY, "BatchPcmReader: output arena full (Flush() first)");
        iOutUsed = at + outBytes;
        iDescs.push_back(d);
        iQueue.push_back(Pending{aPlayable, &aProcessor, outBytes});
    }
    size_t Pending_() const { return iQueue.size(); }

    // Process everything queued on the GPU and deliver it, in queue order.
    void Flush()
    {
        if (iQueue.empty()) return;
        const int rc = ohp_process_host(iCtx, iDescs.data(), iDescs.size(), iIn, iInUsed, iOut, iOutUsed);
        if (rc != OHP_OK) {
            const std::string msg = std::string("ohp_process_host: ") + ohp_last_error(iCtx);
            Drop();
            if (rc == OHP_E_INVALID_DESC) throw AssertionFailed(__FILE__, __LINE__); // what the reference's ASSERT does
            throw OhpError(rc, msg);
        }
        for (size_t i = 0; i < iQueue.size(); i++) {
            Pending& q = iQueue[i];
            const ohp_chunk_desc& d = iDescs[i];
            // MsgPlayable::Read (Msg.cpp:2646-2653)
            q.processor->BeginBlock();
            if (d.bytes > 0) {
                const Brn data(iOut + d.dst_off, q.outBytes);
                const uint32_t subsampleBytes = d.bit_depth / 8u;
                if (d.flags & OHP_F_SILENCE) q.processor->ProcessSilence(data, d.channels, subsampleBytes);
                else q.processor->ProcessFragment(data, d.channels, subsampleBytes);
            }
            q.processor->EndBlock();
            q.playable->RemoveRef();
        }
        iQueue.clear();
        iDescs.clear();
        iOutUsed = 0;
        if (iPins == 0) iInUsed = 0; // nobody refers to the staged audio any more: the space can be handed out again
    }

    // IPlayableReader: MsgPlayable::Read() on a single playable (drop-in, synchronous, does not take the reference).
    void ReadNow(media::MsgPlayable& aPlayable, media::IPcmProcessor& aProcessor) override
    {
        OHP_ASSERT(iQueue.empty()); // mixing queued and synchronous reads would reorder deliveries
        aPlayable.AddRef();
        try {
            Add(&aPlayable, aProcessor);
        }
        catch (...) {
            aPlayable.RemoveRef();
            throw;
        }
        Flush(); // the caller's own reference keeps the arena pinned: it may Split() and read again
    }
    // Forget staged audio; only legal once every message that refers to it is gone.
    void ResetArena() { OHP_ASSERT(iQueue.empty()); OHP_ASSERT(iPins == 0); iInUsed = 0; }
    ohp_context* Context() { return iCtx; }

private:
    struct Pending
    {
        media::MsgPlayable* playable;
        media::IPcmProcessor* processor;
        uint32_t outBytes;
    };
    void Check(int aRc, const char* aWhat)
    {
        if (aRc != OHP_OK) throw OhpError(aRc, std::string(aWhat) + ": " + ohp_last_error(iCtx));
    }
    void Drop()
    {
        for (auto& q : iQueue) q.playable->RemoveRef();
        iQueue.clear();
        iDescs.clear();
        iOutUsed = 0;
        if (iPins == 0) iInUsed = 0;
    }
    void Release()
    {
        if (iCtx) {
            if (iIn) ohp_host_free(iCtx, iIn);
            if (iOut) ohp_host_free(iCtx, iOut);
            ohp_destroy(iCtx);
        }
        iCtx = nullptr; iIn = iOut = nullptr;
    }
private:
    ohp_context* iCtx;
    uint8_t* iIn;
    uint8_t* iOut;
    uint64_t iInCapacity, iOutCapacity, iInUsed, iOutUsed;
    uint64_t iPins; // live messages referring to staged bytes
    std::vector<ohp_chunk_desc> iDescs;
    std::vector<Pending> iQueue;
};

// The reference's ProcessorPcmBufTest (Media/Utils/ProcessorAudioUtils.cpp:18-60): BeginBlock resets, fragments append.
class ProcessorPcmBuf : public media::IPcmProcessor
{
public:
    const std::vector<uint8_t>& Buf() const { return iBuf; }
    void BeginBlock() override { iBuf.clear(); }
    void ProcessFragment(const Brx& aData, uint32_t aNumChannels, uint32_t aSubsampleBytes) override
    {
        OHP_ASSERT(aData.Bytes() % (aSubsampleBytes * aNumChannels) == 0);
        iBuf.insert(iBuf.end(), aData.Ptr(), aData.Ptr() + aData.Bytes());
    }
    void ProcessSilence(const Brx& aData, uint32_t aNumChannels, uint32_t aSubsampleBytes) override
    {
        OHP_ASSERT(aData.Bytes() % (aSubsampleBytes * aNumChannels) == 0);
        iBuf.insert(iBuf.end(), aData.Ptr(), aData.Ptr() + aData.Bytes());
    }
    void EndBlock() override {}
    void Flush() override {}
private:
    std::vector<uint8_t> iBuf;
};

} // namespace ohp
