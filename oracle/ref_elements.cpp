// ref_elements.cpp -- LINKED-REFERENCE ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE), third harness file.
//
// The reference's own ramp-setting ELEMENTS -- Ramper (Media/Pipeline/Ramper.cpp), Muter (Muter.cpp), StarvationRamper
// (StarvationRamper.cpp) and VolumeRamper (VolumeRamper.cpp), compiled unmodified -- pulled through the way the reference's
// own suites pull them (Media/Tests/TestRamper.cpp:164-170, TestMuter.cpp:295-338, TestStarvationRamper.cpp): a fake
// upstream hands out the stream's messages, a sink pulls from the last element and turns every audio message into its
// playable.  What the pipeline's other threads would do to an element -- IMute::Mute()/Unmute(), a MsgDecodedStream or a
// MsgHalt arriving, the reservoir running dry -- happens at the stream positions the stage's OHP_EV_* events name:
// a Gate in front of each element splits the message that straddles an event position (as the stage model does) and then
// makes the call or hands on the control message.  The playables are recorded exactly as ref_harness.cpp records them, so
// that tests/test_elements_vs_reference.py can hold the stage model (StageChain on the same reference message classes, the
// host mirror, the class-free walk, the C port, the device walk) against the element objects themselves.
// No reference source is copied here.

#include <OpenHome/Media/Pipeline/Msg.h>
#include <OpenHome/Media/Utils/ProcessorAudioUtils.h>
#include <OpenHome/Media/Pipeline/Ramper.h>
#include <OpenHome/Media/Pipeline/Muter.h>
#include <OpenHome/Media/Pipeline/VolumeRamper.h>
#include <OpenHome/Media/Pipeline/StarvationRamper.h>
#include <OpenHome/Media/Pipeline/ElementObserver.h>

#include <atomic>
#include <mutex>
#include <chrono>
#include <cstring>
#include <cstdlib>
#include <thread>
#include <unordered_map>
#include <vector>

#include "ohp_oracle.h"
#include "../ohpipeline_b200/host/codec_source.h"

using namespace OpenHome;
using namespace OpenHome::Media;

namespace {

class NullInfo : public IInfoAggregator
{
public:
    void Register(IInfoProvider&, std::vector<Brn>&) override {}
};

// Where each DecodedAudio cell's bytes came from in the stream's PCM.  Written by the Source (on a StarvationRamper's puller
// thread when there is one), read by the Sink.
class CellMap
{
public:
    void Set(const DecodedAudio* aCell, uint64_t aOff) { std::lock_guard<std::mutex> _(iLock); iMap[aCell] = aOff; }
    uint64_t Get(const DecodedAudio* aCell) { std::lock_guard<std::mutex> _(iLock); return iMap[aCell]; }
private:
    std::mutex iLock;
    std::unordered_map<const DecodedAudio*, uint64_t> iMap;
};

struct ElemFactory
{
    NullInfo info;
    MsgFactory* factory;
    ElemFactory()
    {
        MsgFactoryInitParams p;
        p.SetMsgAudioPcmCount(8192, 8192); // a StarvationRamper's puller thread reads the whole stream ahead
        p.SetMsgSilenceCount(1024);
        p.SetMsgPlayableCount(64, 1, 64);
        p.SetMsgDecodedStreamCount(16);
        p.SetMsgHaltCount(64);
        p.SetMsgQuitCount(2);
        p.SetMsgModeCount(2);
        factory = new MsgFactory(info, p);
        static const TByte frame[4] = {0, 0, 0, 0}; // see RefFactory in ref_harness.cpp
        MsgPlayable* warm[64];
        for (TUint i = 0; i < 64; i++) warm[i] = factory->CreateMsgAudioPcm(Brn(frame, 4), 2, 44100, 16, AudioDataEndian::Big, 0)->CreatePlayable();
        for (TUint i = 0; i < 64; i++) warm[i]->RemoveRef();
    }
};

class NullStreamHandler : public IStreamHandler
{
public:
    EStreamPlay OkToPlay(TUint) override { return ePlayYes; }
    TUint TrySeek(TUint, TUint64) override { return MsgFlush::kIdInvalid; }
    TUint TryDiscard(TUint) override { return MsgFlush::kIdInvalid; }
    TUint TryStop(TUint) override { return MsgFlush::kIdInvalid; }
    void NotifyStarving(const Brx&, TUint, TBool) override {}
};

class NullObserver : public IStarvationRamperObserver, public IPipelineElementObserverThread
{
public:
    void NotifyStarvationRamperBuffering(TBool) override {}
    TUint Register(Functor) override { return 0; }
    void Schedule(TUint) override {}
};

class FakeAnimator : public IPipelineAnimator
{
public:
    TUint PipelineAnimatorBufferJiffies() const override { return 0; }
    TUint PipelineAnimatorDelayJiffies(AudioFormat, TUint, TUint, TUint) const override { return 0; }
    void PipelineAnimatorDsdBlockConfiguration(TUint& aWords, TUint& aPad) const override { aWords = 1; aPad = 0; }
    TUint PipelineAnimatorMaxBitDepth() const override { return 32; }
    void PipelineAnimatorGetMaxSampleRates(TUint& aPcm, TUint& aDsd) const override { aPcm = 384000; aDsd = 5644800; }
};

// What feeds the first stage: a MsgDecodedStream, then the stream's PCM in the message sizes its codec delivers with
// MsgSilence where OHP_EV_INSERT_SILENCE says so (stage_chain.h Run()), then MsgQuit.
class Source : public IPipelineElementUpstream
{
public:
    Source(ElemFactory& aF, const ohp_stream_spec& aSpec, const ohp_ramp_event* aEvents, const uint8_t* aPcm,
           CellMap& aCellSrc, IStreamHandler& aHandler)
        : iF(aF), iSpec(aSpec), iEvents(aEvents), iPcm(aPcm), iCellSrc(aCellSrc), iHandler(aHandler)
        , iJps(Jiffies::PerSample(aSpec.sample_rate)), iFrameBytes(aSpec.channels * (aSpec.bit_depth / 8u))
        , iFrame(0), iSrcJiffies(0), iSilEv(0), iStarted(false), iPendingFrames(0), iDone(false)
    {
        ohp::core::codec_source_init(iCodec, aSpec.chunk_frames, aSpec.codec_read_frames, iFrameBytes, iJps, aSpec.total_frames);
    }
    Msg* Pull() override
    {
        if (!iStarted) {
            iStarted = true;
            SpeakerProfile profile(2);
            return iF.factory->CreateMsgDecodedStream(1, 0, iSpec.bit_depth, iSpec.sample_rate, iSpec.channels, Brn("PCM"), 0, 0, true, false,
                                                      false, false, AudioFormat::Pcm, Multiroom::Allowed, profile, &iHandler, RampType::Sample);
        }
        if (iPendingFrames == 0 && !iDone) {
            iPendingFrames = ohp::core::codec_source_next(iCodec);
            if (iPendingFrames == 0) iDone = true;
        }
        if (iDone) {
            return iF.factory->CreateMsgQuit();
        }
        for (; iSilEv < iSpec.num_events; iSilEv++) {
            const ohp_ramp_event& e = iEvents[iSilEv];
            if (e.op != OHP_EV_INSERT_SILENCE) continue;
            if (e.at_jiffies > iSrcJiffies) break;
            TUint jiffies = e.arg;
            iSilEv++;
            return iF.factory->CreateMsgSilence(jiffies, iSpec.sample_rate, iSpec.bit_depth, iSpec.channels);
        }
        const uint32_t frames = iPendingFrames;
        iPendingFrames = 0;
        const uint64_t off = iFrame * iFrameBytes;
        MsgAudioPcm* msg = iF.factory->CreateMsgAudioPcm(Brn(iPcm + off, frames * iFrameBytes), iSpec.channels, iSpec.sample_rate, iSpec.bit_depth,
                                                         iSpec.in_little_endian ? AudioDataEndian::Little : AudioDataEndian::Big,
                                                         iFrame * iJps);
        iCellSrc.Set(msg->iAudioData, off);
        iFrame += frames;
        iSrcJiffies += (uint64_t)frames * iJps;
        return msg;
    }
private:
    ElemFactory& iF;
    const ohp_stream_spec& iSpec;
    const ohp_ramp_event* iEvents;
    const uint8_t* iPcm;
    CellMap& iCellSrc;
    IStreamHandler& iHandler;
    uint32_t iJps, iFrameBytes;
    uint64_t iFrame, iSrcJiffies;
    uint32_t iSilEv;
    bool iStarted;
    uint32_t iPendingFrames;
    bool iDone;
    ohp::core::CodecSource iCodec;
};

class IsAudio : public IMsgProcessor
{
public:
    MsgAudio* audio = nullptr;
    bool silence = false;
    Msg* ProcessMsg(MsgMode* m) override { return m; }
    Msg* ProcessMsg(MsgTrack* m) override { return m; }
    Msg* ProcessMsg(MsgDrain* m) override { return m; }
    Msg* ProcessMsg(MsgDelay* m) override { return m; }
    Msg* ProcessMsg(MsgEncodedStream* m) override { return m; }
    Msg* ProcessMsg(MsgStreamSegment* m) override { return m; }
    Msg* ProcessMsg(MsgAudioEncoded* m) override { return m; }
    Msg* ProcessMsg(MsgMetaText* m) override { return m; }
    Msg* ProcessMsg(MsgStreamInterrupted* m) override { return m; }
    Msg* ProcessMsg(MsgHalt* m) override { return m; }
    Msg* ProcessMsg(MsgFlush* m) override { return m; }
    Msg* ProcessMsg(MsgWait* m) override { return m; }
    Msg* ProcessMsg(MsgDecodedStream* m) override { return m; }
    Msg* ProcessMsg(MsgAudioPcm* m) override { audio = m; silence = false; return m; }
    Msg* ProcessMsg(MsgAudioDsd* m) override { return m; }
    Msg* ProcessMsg(MsgSilence* m) override { audio = m; silence = true; return m; }
    Msg* ProcessMsg(MsgPlayable* m) override { return m; }
    Msg* ProcessMsg(MsgQuit* m) override { return m; }
};

struct Elements; // the chain's element objects, for the gates to call into

// In front of stage aStage's element: counts the audio that passes, cuts the message an event falls inside (the remainder
// waits for the next Pull, as the stage model re-queues it at the head), and at the event's position does what the event
// stands for.
class Gate : public IPipelineElementUpstream
{
public:
    Gate(ElemFactory& aF, IPipelineElementUpstream& aUp, const ohp_stream_spec& aSpec, const ohp_ramp_event* aEvents, uint32_t aStage,
         Elements& aElems, IStreamHandler& aHandler)
        : iF(aF), iUp(aUp), iSpec(aSpec), iEvents(aEvents), iStage(aStage), iElems(aElems), iHandler(aHandler)
        , iJps(Jiffies::PerSample(aSpec.sample_rate)), iPos(0), iNextEv(0), iHeld(nullptr), iNextStreamId(2)
        , iAtStarvation(false), iStarvationArg(0), iRelease("GATE", 0) {}
    Msg* Pull() override;
    // StarvationRamper: the puller thread is parked here until the sink has played the reservoir dry
    std::atomic<bool> iAtStarvationFlag{false};
    void ReleaseStarvation() { iAtStarvationFlag.store(false); iRelease.Signal(); }
private:
    bool NextEvent(uint32_t& aIndex)
    {
        while (iNextEv < iSpec.num_events) {
            const ohp_ramp_event& e = iEvents[iNextEv];
            if (e.stage == iStage && e.op != OHP_EV_INSERT_SILENCE) { aIndex = iNextEv; return true; }
            iNextEv++;
        }
        return false;
    }
    Msg* Fire(const ohp_ramp_event& e); // returns a control message to hand on, or nullptr
private:
    ElemFactory& iF;
    IPipelineElementUpstream& iUp;
    const ohp_stream_spec& iSpec;
    const ohp_ramp_event* iEvents;
    uint32_t iStage;
    Elements& iElems;
    IStreamHandler& iHandler;
    uint32_t iJps;
    uint64_t iPos;
    uint32_t iNextEv;
    Msg* iHeld;
    bool iStreamSeen = false;
    TUint iNextStreamId;
    bool iAtStarvation;
    uint32_t iStarvationArg;
    uint32_t iAttenuation = OHP_UNITY_ATTENUATION;
    Semaphore iRelease;
};

struct Elements
{
    Ramper* ramper[OHP_MAX_STAGES] = {};
    Muter* muter[OHP_MAX_STAGES] = {};
    StarvationRamper* starvation[OHP_MAX_STAGES] = {};
    Gate* gate[OHP_MAX_STAGES] = {};
    IPipelineElementUpstream* after[OHP_MAX_STAGES] = {};
    std::vector<std::thread> muteCallers; // IMute::Mute() blocks until the mute has taken effect: it gets a thread of its own
    std::atomic<bool> asserted{false};
};

Msg* Gate::Fire(const ohp_ramp_event& e)
{
    switch (e.op) {
    case OHP_EV_RAMPER_STREAM: {
        // a new stream reaches the Ramper: live (IsRampApplicable, Ramper.cpp:136-152) when a ramp is asked for
        SpeakerProfile profile(2);
        return iF.factory->CreateMsgDecodedStream(iNextStreamId++, 0, iSpec.bit_depth, iSpec.sample_rate, iSpec.channels, Brn("PCM"), 0, 0, true, false,
                                                  e.arg != 0, false, AudioFormat::Pcm, Multiroom::Allowed, profile, &iHandler, RampType::Sample);
    }
    case OHP_EV_MUTER_MUTE: {
        Muter* m = iElems.muter[iStage];
        m->iLock.Wait();
        const int before = (int)m->iState;
        const bool legal = m->iState == Muter::eRunning || m->iState == Muter::eRampingUp; // else Mute() ASSERTS (Muter.cpp:88-90)
        m->iLock.Signal();
        if (!legal) { iElems.asserted = true; return nullptr; }
        iElems.muteCallers.emplace_back([m] { m->Mute(); });
        for (;;) { // until the call has changed the Muter's state (it may then stay blocked until the mute completes)
            m->iLock.Wait();
            const bool changed = (int)m->iState != before;
            m->iLock.Signal();
            if (changed) break;
            std::this_thread::sleep_for(std::chrono::microseconds(50));
        }
        return nullptr;
    }
    case OHP_EV_MUTER_UNMUTE: {
        Muter* m = iElems.muter[iStage];
        m->iLock.Wait();
        const bool legal = !(m->iState == Muter::eRunning || m->iState == Muter::eRampingUp); // else Unmute() ASSERTS (Muter.cpp:107-111)
        m->iLock.Signal();
        if (!legal) { iElems.asserted = true; return nullptr; }
        m->Unmute();
        return nullptr;
    }
    case OHP_EV_HALT:
        return iF.factory->CreateMsgHalt();
    case OHP_EV_STARVATION:
        // nothing more comes from upstream until the sink has played the reservoir dry and the element has reacted
        iAtStarvationFlag.store(true);
        iRelease.Wait();
        return nullptr;
    case OHP_EV_SET_ATTENUATION:
        // what an Attenuator ahead of the element does to every MsgAudioPcm from here on (Attenuator.cpp:55-58)
        iAttenuation = e.arg;
        return nullptr;
    default:
        return nullptr; // OHP_EV_MAX_MSG_JIFFIES and the bare ramp ops have no element to go to
    }
}

Msg* Gate::Pull()
{
    if (!iStreamSeen) {
        iStreamSeen = true;
        return iUp.Pull(); // the Source's MsgDecodedStream reaches every element before anything happens to it
    }
    for (;;) {
        uint32_t ei;
        while (NextEvent(ei) && iEvents[ei].at_jiffies <= iPos) {
            iNextEv = ei + 1;
            Msg* control = Fire(iEvents[ei]);
            if (control != nullptr) return control;
        }
        Msg* msg = iHeld != nullptr ? iHeld : iUp.Pull();
        iHeld = nullptr;
        IsAudio probe;
        (void)msg->Process(probe);
        if (probe.audio == nullptr) return msg;
        MsgAudio* audio = probe.audio;
        if (NextEvent(ei) && iEvents[ei].at_jiffies < iPos + audio->Jiffies()) {
            uint32_t at = (uint32_t)(iEvents[ei].at_jiffies - iPos);
            if (probe.silence) at -= at % iJps; // silence only splits on sample blocks
            if (at == 0) {
                iNextEv = ei + 1;
                iHeld = msg;
                Msg* control = Fire(iEvents[ei]);
                if (control != nullptr) return control;
                continue;
            }
            iHeld = audio->Split(at);
        }
        iPos += audio->Jiffies();
        if (!probe.silence && iAttenuation != OHP_UNITY_ATTENUATION) static_cast<MsgAudioPcm*>(audio)->SetAttenuation(iAttenuation);
        return msg;
    }
}

// Behind stage aStage's element.  The stage model's events are per stage: the control messages a Gate makes for ITS element
// (MsgHalt, the MsgDecodedStream of a new stream) end here instead of travelling on to the elements downstream, as does the
// flywheel audio a StarvationRamper plays when it starves (not part of the stream the model describes; counted).  For a
// StarvationRamper this is also where its starvations are staged: its puller thread parks in the Gate at the event's
// position, the reservoir is played dry, and only then is the element pulled on an empty reservoir
// (StarvationRamper::Pull, StarvationRamper.cpp:622-673) -- never before, whatever the two threads' timing.
// StartFlywheelRamp cuts the element's recent audio to its last kTrainingJiffies by splitting the message under the cut
// (StarvationRamper.cpp:495-507).  A MsgSilence split at a jiffy count that is not a whole sample keeps the whole samples in
// front (MsgSilence::SplitCompleted, Msg.cpp:2530-2535): less is taken off than asked for, the loop comes round with under a
// sample of excess, splits off a message of ZERO jiffies, subtracts nothing, and never ends.  The harness looks before it
// lets the element in: such a starvation ends the stream with -3 instead of hanging the caller.
struct ReferenceWouldNotReturn {};

static bool CutNeverEnds(StarvationRamper& aSr)
{
    if (aSr.iRecentAudioJiffies <= StarvationRamper::kTrainingJiffies) return false;
    TUint excess = aSr.iRecentAudioJiffies - StarvationRamper::kTrainingJiffies;
    const TUint jps = Jiffies::PerSample(aSr.iSampleRate);
    for (Msg* m = aSr.iRecentAudio.iHead; m != nullptr && excess > 0; m = m->iNextMsg) {
        IsAudio probe;
        (void)m->Process(probe);
        const TUint jiffies = probe.audio->Jiffies();
        if (jiffies > excess) return probe.silence && excess % jps != 0; // the message that is split
        excess -= jiffies;
    }
    return false;
}

class After : public IPipelineElementUpstream
{
public:
    After(IPipelineElementUpstream& aElement, Gate& aGate, StarvationRamper* aStarvation, uint32_t& aGenerated, uint32_t& aHalts,
          std::vector<uint8_t>& aGeneratedAudio, std::vector<uint32_t>& aStarvedAtRamp)
        : iElement(aElement), iGate(aGate), iSr(aStarvation), iGenerated(aGenerated), iHalts(aHalts), iGeneratedAudio(aGeneratedAudio)
        , iStarvedAtRamp(aStarvedAtRamp), iProbe(*this) {}
    Msg* Pull() override
    {
        for (;;) {
            bool generating = false;
            if (iSr != nullptr) {
                for (;;) { // wait for something to pull: buffered messages, or the upstream parked at a starvation
                    iSr->iLock.Wait();
                    const bool empty = iSr->IsEmpty();
                    iSr->iLock.Signal();
                    if (!empty) break;
                    if (iGate.iAtStarvationFlag.load()) {
                        // the upstream thread is parked now and enqueues nothing more -- but it may have enqueued its last
                        // message between the look at the reservoir above and the look at the flag: look again
                        iSr->iLock.Wait();
                        const bool stillEmpty = iSr->IsEmpty();
                        iSr->iLock.Signal();
                        if (!stillEmpty) break;
                        const bool reacts = iSr->iState == StarvationRamper::State::Running
                                         || (iSr->iState == StarvationRamper::State::RampingUp && iSr->iCurrentRampValue != Ramp::kMin);
                        if (!reacts) { iGate.ReleaseStarvation(); continue; }
                        generating = true;
                        break;
                    }
                    std::this_thread::sleep_for(std::chrono::microseconds(20));
                }
            }
            if (generating) {
                // the flywheel ramp, then the MsgHalt that ends it (StarvationRamper.cpp:640-650); what a driver would read from
                // the generated messages is kept (PreDriver -> CreatePlayable -> Read, as for the stream's own audio)
                // (OHP_REF_LET_IT_RUN=1 lets the element in regardless: how tests show, under a timeout, that it does not come back)
                if (CutNeverEnds(*iSr) && std::getenv("OHP_REF_LET_IT_RUN") == nullptr) throw ReferenceWouldNotReturn();
                iStarvedAtRamp.push_back(iSr->iCurrentRampValue);
                for (;;) {
                    Msg* msg = iElement.Pull();
                    iProbe.Reset();
                    (void)msg->Process(iProbe);
                    const bool halt = iProbe.halt;
                    if (halt) { iHalts++; static_cast<MsgHalt*>(msg)->ReportHalted(); }
                    if (iProbe.audio && !iProbe.silence) {
                        iGenerated++;
                        MsgPlayable* playable = static_cast<MsgAudioPcm*>(iProbe.audio)->CreatePlayable(); // takes the message's reference
                        ProcessorPcmBufTest proc;
                        playable->Read(proc);
                        const Brn bytes = proc.Buf();
                        iGeneratedAudio.insert(iGeneratedAudio.end(), bytes.Ptr(), bytes.Ptr() + bytes.Bytes());
                        playable->RemoveRef();
                    }
                    else {
                        if (iProbe.audio) iGenerated++;
                        msg->RemoveRef();
                    }
                    if (halt) break;
                }
                iGate.ReleaseStarvation();
                continue;
            }
            Msg* msg = iElement.Pull();
            iProbe.Reset();
            (void)msg->Process(iProbe);
            if (iProbe.halt) { // this stage's own: the animator acknowledges at once (TestMuter.cpp's ProcessMsg(MsgHalt))
                iHalts++;
                static_cast<MsgHalt*>(msg)->ReportHalted();
                msg->RemoveRef();
                continue;
            }
            if (iProbe.stream && iProbe.streamId != 1) { msg->RemoveRef(); continue; }
            return msg;
        }
    }
private:
    class Probe : public IsAudio
    {
    public:
        explicit Probe(After&) {}
        bool halt = false, stream = false;
        TUint streamId = 0;
        void Reset() { audio = nullptr; silence = false; halt = false; stream = false; }
        Msg* ProcessMsg(MsgHalt* m) override { halt = true; return m; }
        Msg* ProcessMsg(MsgDecodedStream* m) override { stream = true; streamId = m->StreamInfo().StreamId(); return m; }
    };
    IPipelineElementUpstream& iElement;
    Gate& iGate;
    StarvationRamper* iSr;
    uint32_t& iGenerated;
    uint32_t& iHalts;
    std::vector<uint8_t>& iGeneratedAudio;
    std::vector<uint32_t>& iStarvedAtRamp;
    Probe iProbe;
};

struct StreamOut
{
    std::vector<ohp_chunk_desc> chunks;
    std::vector<ohp_chunk_info> info;
    uint64_t outBytes = 0;
    uint32_t generated = 0; // flywheel messages the StarvationRamper played (not part of the stream)
    uint32_t halts = 0;
    std::vector<uint8_t> generatedAudio;   // ... what a driver reads from them (every starvation of the stream, one after the other)
    std::vector<uint32_t> starvedAtRamp;   // ... the element's iCurrentRampValue when it starved, per starvation
};

// Pulls the last element until MsgQuit: PreDriver -> CreatePlayable, a driver that pulls fixed blocks (stage_chain.h Drive()).
class Sink : public IMsgProcessor
{
public:
    Sink(const ohp_stream_spec& aSpec, CellMap& aCellSrc, uint8_t* aOut, StreamOut& aRec)
        : iSpec(aSpec), iCellSrc(aCellSrc), iOut(aOut), iRec(aRec), iFrameBytes(aSpec.channels * (aSpec.bit_depth / 8u)), iBlockFill(0)
        , iQuit(false), iGenerating(false) {}
    bool Quit() const { return iQuit; }
    void SetGenerating(bool aOn) { iGenerating = aOn; }
    bool SawHalt() { const bool h = iSawHalt; iSawHalt = false; return h; }
private:
    void OnPlayable(MsgPlayable* p)
    {
        const uint32_t bytes = p->Bytes();
        ohp_chunk_desc d;
        std::memset(&d, 0, sizeof d);
        MsgPlayablePcm* pcm = dynamic_cast<MsgPlayablePcm*>(p);
        d.dst_off = iSpec.dst_base + iRec.outBytes;
        d.bytes = bytes;
        d.ramp_start = (uint16_t)p->Ramp().Start();
        d.ramp_end = (uint16_t)p->Ramp().End();
        d.bit_depth = (uint8_t)iSpec.bit_depth;
        d.channels = (uint8_t)iSpec.channels;
        d.out_fmt = (uint8_t)iSpec.out_fmt;
        d.aux = iSpec.out_fmt == OHP_OUT_PACKED_LE ? OHP_LE_APPEND : 0;
        d.flags = p->Ramp().IsEnabled() ? OHP_F_RAMP_ENABLED : 0;
        if (pcm != nullptr) {
            d.src_off = iSpec.src_base + iCellSrc.Get(pcm->iAudioData) + pcm->iOffset;
            d.attenuation = (uint16_t)pcm->iAttenuation;
            if (iSpec.in_little_endian) d.flags |= OHP_F_IN_LITTLE_ENDIAN;
        }
        else {
            d.attenuation = OHP_UNITY_ATTENUATION;
            d.flags |= OHP_F_SILENCE;
        }
        ohp_chunk_info ci;
        ci.direction = (uint32_t)p->Ramp().Direction();
        ci.jiffies = p->Jiffies();
        iRec.chunks.push_back(d);
        iRec.info.push_back(ci);
        if (iOut != nullptr) {
            p->Read(iProc);
            ASSERT(iProc.Buf().Bytes() == bytes);
            if (bytes > 0) std::memcpy(iOut + iSpec.dst_base + iRec.outBytes, iProc.Ptr(), bytes);
        }
        iRec.outBytes += bytes;
        p->RemoveRef();
    }
    Msg* Audio(MsgAudio* aMsg)
    {
        MsgPlayable* playable = aMsg->CreatePlayable();
        const uint32_t block = iSpec.driver_block_frames * iFrameBytes;
        if (block == 0 || playable->Bytes() == 0) {
            OnPlayable(playable);
            return nullptr;
        }
        for (;;) {
            const uint32_t room = block - iBlockFill;
            if (playable->Bytes() > room) {
                MsgPlayable* remaining = playable->Split(room);
                OnPlayable(playable);
                iBlockFill = 0;
                playable = remaining;
            }
            else {
                const uint32_t bytes = playable->Bytes();
                OnPlayable(playable);
                iBlockFill += bytes;
                if (iBlockFill == block) iBlockFill = 0;
                return nullptr;
            }
        }
    }
    Msg* Drop(Msg* m) { m->RemoveRef(); return nullptr; }
public: // from IMsgProcessor
    Msg* ProcessMsg(MsgMode* m) override { return Drop(m); }
    Msg* ProcessMsg(MsgTrack* m) override { return Drop(m); }
    Msg* ProcessMsg(MsgDrain* m) override { m->ReportDrained(); return Drop(m); }
    Msg* ProcessMsg(MsgDelay* m) override { return Drop(m); }
    Msg* ProcessMsg(MsgEncodedStream* m) override { return Drop(m); }
    Msg* ProcessMsg(MsgStreamSegment* m) override { return Drop(m); }
    Msg* ProcessMsg(MsgAudioEncoded* m) override { return Drop(m); }
    Msg* ProcessMsg(MsgMetaText* m) override { return Drop(m); }
    Msg* ProcessMsg(MsgStreamInterrupted* m) override { return Drop(m); }
    Msg* ProcessMsg(MsgHalt* m) override
    {
        iRec.halts++;
        iSawHalt = true;
        m->ReportHalted(); // the animator acknowledges at once (TestMuter.cpp's ProcessMsg(MsgHalt) without deferral)
        return Drop(m);
    }
    Msg* ProcessMsg(MsgFlush* m) override { return Drop(m); }
    Msg* ProcessMsg(MsgWait* m) override { return Drop(m); }
    Msg* ProcessMsg(MsgDecodedStream* m) override { return Drop(m); }
    Msg* ProcessMsg(MsgAudioPcm* m) override { return Audio(m); }
    Msg* ProcessMsg(MsgAudioDsd* m) override { return Drop(m); }
    Msg* ProcessMsg(MsgSilence* m) override { return Audio(m); }
    Msg* ProcessMsg(MsgPlayable* m) override { return Drop(m); }
    Msg* ProcessMsg(MsgQuit* m) override { iQuit = true; return Drop(m); }
private:
    const ohp_stream_spec& iSpec;
    CellMap& iCellSrc;
    uint8_t* iOut;
    StreamOut& iRec;
    uint32_t iFrameBytes, iBlockFill;
    bool iQuit, iGenerating;
    bool iSawHalt = false;
    ProcessorPcmBufTest iProc;
};

// Which element a stage is, by the ops of its events: 0 none, 1 Ramper, 2 Muter, 3 StarvationRamper, -1 not representable here
int StageKind(const ohp_stream_spec& sp, const ohp_ramp_event* ev, uint32_t stage, uint32_t& arg)
{
    int kind = 0;
    arg = 0;
    for (uint32_t i = 0; i < sp.num_events; i++) {
        if (ev[i].stage != stage) continue;
        int of = 0;
        switch (ev[i].op) {
        case OHP_EV_RAMPER_STREAM: of = 1; break;
        case OHP_EV_MUTER_MUTE: case OHP_EV_MUTER_UNMUTE: of = 2; break;
        case OHP_EV_STARVATION: of = 3; break;
        case OHP_EV_HALT: case OHP_EV_INSERT_SILENCE: break;
        case OHP_EV_SET_ATTENUATION: if (kind == 0) kind = 4; break; // the stage's Gate alone (an Attenuator ahead of the element, if any)
        default: return -1; // the bare ramp ops and message caps are the stage MODEL's; no element takes them
        }
        if (of != 0) {
            if (kind != 0 && kind != 4 && kind != of) return -1;
            kind = of;
            if (ev[i].arg != 0) {
                if (arg != 0 && arg != ev[i].arg) return -1; // an element has ONE ramp duration
                arg = ev[i].arg;
            }
        }
    }
    return kind;
}

// One stream through the real elements.  Returns 0, -1 where the reference ASSERTs (or this harness would have to make a
// call the reference ASSERTS on), -2 for a schedule this harness cannot stage, -3 where the reference would never return
// (CutNeverEnds above).
int RunStream(const ohp_stream_spec& sp, const ohp_ramp_event* events, const uint8_t* in, uint8_t* out, StreamOut& rec)
{
    ElemFactory* f = new ElemFactory(); // leaked when an ASSERT unwinds with messages in flight (see ref_harness.cpp)
    CellMap cellSrc;
    NullStreamHandler handler;
    NullObserver observer;
    FakeAnimator animator;
    const ohp_ramp_event* ev = events + sp.first_event;
    Elements elems;
    std::vector<IPipelineElementUpstream*> owned;
    Source source(*f, sp, ev, in + sp.src_base, cellSrc, handler);
    IPipelineElementUpstream* up = &source;
    int rc = 0;
    for (uint32_t s = 0; s < OHP_MAX_STAGES && rc == 0; s++) {
        uint32_t arg;
        const int kind = StageKind(sp, ev, s, arg);
        if (kind < 0) { rc = -2; break; }
        if (kind == 0) continue;
        Gate* gate = new Gate(*f, *up, sp, ev, s, elems, handler);
        elems.gate[s] = gate;
        if (kind == 1) {
            elems.ramper[s] = new Ramper(*gate, arg, arg);
            up = elems.ramper[s];
        }
        else if (kind == 2) {
            elems.muter[s] = new Muter(*f->factory, *gate, arg);
            elems.muter[s]->SetAnimator(animator);
            up = elems.muter[s];
        }
        else if (kind == 4) {
            up = gate; // no element: the attenuation alone
            continue;
        }
        else {
            elems.starvation[s] = new StarvationRamper(*f->factory, *gate, observer, observer, 0x7fffffffu /* never full */, kPriorityNormal,
                                                       arg, 1000);
            up = elems.starvation[s];
        }
        elems.after[s] = new After(*up, *gate, elems.starvation[s], rec.generated, rec.halts, rec.generatedAudio, rec.starvedAtRamp);
        up = elems.after[s];
    }
    if (rc == 0) {
        Sink sink(sp, cellSrc, out, rec);
        try {
            while (!sink.Quit() && !elems.asserted) {
                Msg* msg = up->Pull();
                msg = msg->Process(sink);
                ASSERT(msg == nullptr);
            }
            if (elems.asserted) rc = -1;
        }
        catch (Exception& e) {
            if (std::getenv("OHP_REF_TRACE") != nullptr) std::fprintf(stderr, "ref elements: %s at %s:%u\n", e.Message(), e.File(), e.Line());
            rc = -1;
        }
        catch (ReferenceWouldNotReturn&) {
            rc = -3;
        }
    }
    // callers of Mute() still waiting for their mute to complete (it never will: the stream is over)
    for (uint32_t s = 0; s < OHP_MAX_STAGES; s++) {
        if (elems.muter[s] != nullptr) {
            for (size_t k = 0; k < elems.muteCallers.size() + 1; k++) elems.muter[s]->iSemMuted.Signal();
        }
    }
    for (auto& t : elems.muteCallers) t.join();
    if (rc == 0) {
        // orderly teardown; after an ASSERT everything is abandoned with its factory
        for (int s = (int)OHP_MAX_STAGES - 1; s >= 0; s--) {
            if (elems.starvation[s] != nullptr) {
                // its puller thread is blocked in our Source (which has nothing more to give): let it go with a quit
                // (MsgQuit was already pulled through: iExit is set, the thread has left its loop)
                delete elems.starvation[s];
            }
            delete elems.ramper[s];
            delete elems.muter[s];
            delete elems.after[s];
            delete elems.gate[s];
        }
        delete f->factory;
        delete f;
    }
    return rc;
}

// ---- VolumeRamper -----------------------------------------------------------------------------------------------------

class ListUpstream : public IPipelineElementUpstream
{
public:
    std::vector<Msg*> msgs;
    size_t next = 0;
    Msg* Pull() override { return msgs[next++]; }
};

class VolumeRecorder : public IVolumeRamper
{
public:
    std::vector<TUint> multipliers;
    void ApplyVolumeMultiplier(TUint aValue) override { multipliers.push_back(aValue); }
};

} // namespace

extern "C" {

// The real VolumeRamper (VolumeRamper.cpp:77-122) fed a MsgDecodedStream -- RampType::Volume when `enabled`, the sample-ramped
// default otherwise -- and then one message per ramp: kind[i] 0 = MsgAudioPcm carrying ramps[i], 1 = MsgSilence.  multipliers
// receives every value handed to IVolumeRamper::ApplyVolumeMultiplier (returns their number), after[i] the ramp each message
// leaves with.  Callers keep the median ramp above 0 (index 512 is read out of bounds, see ref_median_multiplier).
int ref_volume_ramper(int enabled, const ohp_ramp* ramps, const uint32_t* kind, uint32_t n, uint32_t* multipliers, ohp_ramp* after)
{
    ElemFactory* f = new ElemFactory();
    int count = 0;
    try {
        ListUpstream up;
        VolumeRecorder rec;
        VolumeRamper vr(*f->factory, up);
        vr.SetVolumeRamper(rec);
        SpeakerProfile profile(2);
        up.msgs.push_back(f->factory->CreateMsgDecodedStream(1, 0, 16, 44100, 2, Brn("PCM"), 0, 0, true, false, false, false, AudioFormat::Pcm,
                                                             Multiroom::Allowed, profile, nullptr, enabled ? RampType::Volume : RampType::Sample));
        static const TByte data[64] = {0};
        for (uint32_t i = 0; i < n; i++) {
            MsgAudio* msg;
            if (kind[i] == 0) msg = f->factory->CreateMsgAudioPcm(Brn(data, 64), 2, 44100, 16, AudioDataEndian::Big, 0);
            else { TUint j = 16 * Jiffies::PerSample(44100); msg = f->factory->CreateMsgSilence(j, 44100, 16, 2); }
            msg->iRamp.iStart = ramps[i].start;
            msg->iRamp.iEnd = ramps[i].end;
            msg->iRamp.iDirection = (Ramp::EDirection)ramps[i].direction;
            msg->iRamp.iEnabled = ramps[i].enabled != 0;
            up.msgs.push_back(msg);
        }
        vr.Pull()->RemoveRef();
        for (uint32_t i = 0; i < n; i++) {
            Msg* m = vr.Pull();
            MsgAudio* audio = static_cast<MsgAudio*>(m);
            after[i].start = audio->Ramp().Start();
            after[i].end = audio->Ramp().End();
            after[i].direction = (uint32_t)audio->Ramp().Direction();
            after[i].enabled = audio->Ramp().IsEnabled() ? 1u : 0u;
            m->RemoveRef();
        }
        count = (int)rec.multipliers.size();
        for (int i = 0; i < count; i++) multipliers[i] = rec.multipliers[i];
    }
    catch (AssertionFailed&) {
        return -1;
    }
    delete f->factory;
    delete f;
    return count;
}

// Every stream of the batch through the reference's element objects (one after the other).  out == NULL: descriptors only.
// generated (may be NULL, n_streams entries): flywheel messages each stream's StarvationRamper played.
// Returns 0, -1 (the reference ASSERTs somewhere), -2 (a schedule the harness cannot stage: bare ramp ops, two ramp durations
// for one element, ...).
int ref_elements_run(const ohp_stream_spec* streams, size_t n_streams, const ohp_ramp_event* events, size_t n_events,
                     const uint8_t* in, uint8_t* out, ohpo_schedule_result* res, uint32_t* generated)
{
    (void)n_events;
    std::vector<StreamOut> recs(n_streams);
    for (size_t s = 0; s < n_streams; s++) {
        const int rc = RunStream(streams[s], events, in, out, recs[s]);
        if (rc != 0) return rc;
        if (generated) generated[s] = recs[s].generated;
    }
    std::memset(res, 0, sizeof *res);
    size_t total = 0;
    for (auto& r : recs) total += r.chunks.size();
    res->chunks = (ohp_chunk_desc*)std::malloc((total ? total : 1) * sizeof(ohp_chunk_desc));
    res->info = (ohp_chunk_info*)std::malloc((total ? total : 1) * sizeof(ohp_chunk_info));
    res->stream_chunk_begin = (uint64_t*)std::calloc(n_streams + 1, sizeof(uint64_t));
    res->stream_out_bytes = (uint64_t*)std::calloc(n_streams ? n_streams : 1, sizeof(uint64_t));
    size_t k = 0;
    for (size_t s = 0; s < n_streams; s++) {
        res->stream_chunk_begin[s] = k;
        res->stream_out_bytes[s] = recs[s].outBytes;
        if (!recs[s].chunks.empty()) {
            std::memcpy(res->chunks + k, recs[s].chunks.data(), recs[s].chunks.size() * sizeof(ohp_chunk_desc));
            std::memcpy(res->info + k, recs[s].info.data(), recs[s].info.size() * sizeof(ohp_chunk_info));
        }
        k += recs[s].chunks.size();
    }
    res->stream_chunk_begin[n_streams] = k;
    res->num_chunks = k;
    return 0;
}

// ONE stream through the element objects, for what its StarvationRamper plays when it starves: the bytes a driver reads from
// the flywheel messages (every starvation, one after the other) and the ramp value the element had when each began.
// Returns the number of bytes (copied up to cap), -1 / -2 as ref_elements_run.
long ref_elements_generated_audio(const ohp_stream_spec* stream, const ohp_ramp_event* events, const uint8_t* in,
                                  uint8_t* audio, size_t cap, uint32_t* ramps, size_t ramps_cap, uint32_t* n_starvations)
{
    StreamOut rec;
    const int rc = RunStream(*stream, events, in, nullptr, rec);
    if (rc != 0) return rc;
    const size_t n = rec.generatedAudio.size() < cap ? rec.generatedAudio.size() : cap;
    if (n) std::memcpy(audio, rec.generatedAudio.data(), n);
    for (size_t i = 0; i < rec.starvedAtRamp.size() && i < ramps_cap; i++) ramps[i] = rec.starvedAtRamp[i];
    if (n_starvations) *n_starvations = (uint32_t)rec.starvedAtRamp.size();
    return (long)rec.generatedAudio.size();
}

} // extern "C"
