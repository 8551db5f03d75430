// stage_chain.h -- drives a stream's audio messages through a chain of ramp-setting stages, then
// through CreatePlayable and a block-pulling driver, exactly as ohPipeline's elements do:
//
//     if (msg->Jiffies() > remaining) split = msg->Split(remaining);        // requeued at the head
//     current = msg->SetRamp(current, remaining, direction, split);         // split requeued too
//
// (Ramper.cpp:114-134, Muter.cpp:210-262, StarvationRamper.cpp:579-603, 791-832), then
// PreDriver -> CreatePlayable (PreDriver.cpp:115-133) and a driver that Split()s playables into
// fixed blocks (Av/Utils/DriverSongcastSender.cpp:176-199).
//
// It is a template over the message API so that ONE statement of the stage logic runs against
//   * this repo's host mirror (ohpipeline_b200/host/msg_model.h)  -> chunk descriptors for the GPU, and
//   * the reference's own classes (OpenHome::Media::MsgAudio ...) -> the linked-reference oracle
//     (oracle/ref_harness.cpp), which is how descriptor parity is checked.
// The message API must offer the reference's names: Jiffies(), Split(), SetRamp(), SetMuted(),
// SetAttenuation(), CreatePlayable(), RemoveRef(), MsgPlayable::Split()/Bytes().
#pragma once

#include <cstdint>
#include <deque>
#include <vector>

#include "../../include/ohp_schedule.h"
#include "codec_source.h"

namespace ohp {

// Api must provide:
//   types   MsgAudio, MsgAudioPcm, MsgSilence, MsgPlayable, Factory
//   consts  kRampMax, kRampMin, kDirUp, kDirDown (of the type SetRamp takes)
//   static  MsgAudioPcm* CreatePcm(Factory&, const ohp_stream_spec&, uint64_t first_frame, uint32_t frames);
//   static  MsgSilence*  CreateSilence(Factory&, const ohp_stream_spec&, uint32_t& jiffies);
//   static  uint32_t     JiffiesPerSample(uint32_t rate);
//   static  void         Assert(bool) -- raises the API's AssertionFailed
template <class Api, class Sink>
class StageChain
{
    using MsgAudio = typename Api::MsgAudio;
    using MsgPlayable = typename Api::MsgPlayable;

    struct Item
    {
        MsgAudio* msg;
        bool silence;
    };
public:
    // A piece of a StarvationRamper stage's recent audio: a MsgSilence as it passed, or PCM (adjacent messages of one
    // attenuation run together: MsgAudioPcm can be cut anywhere, a MsgSilence only between samples -- Msg.cpp:2522-2545 --
    // which is why silence stays message by message)
    struct Recent
    {
        uint64_t pcmJiffies; // PCM: where in the stream's PCM it begins
        uint32_t jiffies;
        uint32_t silence;
        uint32_t attenuation;
    };
private:
    enum Mode { Running = 0, RampingDown = 1, RampingUp = 2, Muted = 3 };
    // which of the reference's elements a stage is (decided by the ops of its events, see StageElement below)
    enum Element { Generic = 0, ElemRamper = 1, ElemMuter = 2, ElemStarvation = 3, ElemBad = 4 };
    struct Stage
    {
        std::deque<Item> queue;
        Mode mode = Running;
        uint32_t current = Api::kRampMax;
        uint32_t remaining = 0;
        uint32_t maxMsg = 0;
        uint32_t attenuation = OHP_UNITY_ATTENUATION;
        uint64_t pos = 0;
        uint32_t nextEv = 0;
        Element elem = Generic;
        bool halted = true; // Muter::iHalted / StarvationRamper's Halted-or-Starting: no PCM has passed since the start or the last halt
        uint64_t pcmJiffies = 0; // PCM that has passed the stage (pos counts MsgSilence too)
        uint64_t pcmRun = 0;     // ... the latest unbroken run of it: nothing but PCM with one attenuation (pcmRunAtt) since the
        uint32_t pcmRunAtt = OHP_UNITY_ATTENUATION; // stream began, a MsgSilence passed or a flywheel ramp used the recent audio up
        std::deque<Recent> recent;         // StarvationRamper::iRecentAudio: what the element has handed on, oldest first, cut back
        uint64_t recentJiffies = 0;        // (whole pieces only) to what still covers the last millisecond
        uint32_t elemRamp = Api::kRampMax; // StarvationRamper::iCurrentRampValue as the element itself keeps it: what SetRamp last
                                           // returned -- NOT reset to kMax when its ramp up completes, so where the messages carried
                                           // a lower ramp from upstream it stays at where THAT ramp stood (StarvationRamper.cpp:812-817)
    };
public:
    // What a StarvationRamper stage was doing when its reservoir ran dry (OHP_EV_STARVATION): what the flywheel ramp it then
    // plays is made from (ohp_schedule.h, ohp_starvation).  Stream-local: `event` indexes the stream's own events.
    struct Starvation
    {
        uint32_t event;
        uint32_t ramp;          // the element's ramp value: what RampGenerator starts from
        uint32_t plays;         // running, or ramping up and audible: the flywheel ramp plays
        uint32_t recentJiffies; // the unbroken run of PCM the element's recent audio ends with, saturated
        uint32_t attenuation;   // ... and the attenuation its messages carry
        uint64_t pcmJiffies;    // PCM that had passed the element
        std::vector<Recent> recent; // the element's recent audio, oldest first (covers the last millisecond where that much has passed)
    };

    StageChain(typename Api::Factory& aFactory, const ohp_stream_spec& aSpec,
               const ohp_ramp_event* aEvents, Sink& aSink)
        : iFactory(aFactory), iSpec(aSpec), iEvents(aEvents), iNumEvents(aSpec.num_events), iSink(aSink)
        , iJps(Api::JiffiesPerSample(aSpec.sample_rate))
        , iFrameBytes(aSpec.channels * (aSpec.bit_depth / 8u))
        , iBlockFill(0)
    {
    }
    ~StageChain()
    {
        // only non-empty after an assertion unwound Run()
        for (auto& st : iStages) {
            for (auto& it : st.queue) {
                it.msg->RemoveRef();
            }
        }
    }
    // Where Run() leaves a record per OHP_EV_STARVATION it applies (null: nowhere).
    void SetStarvationLog(std::vector<Starvation>* aLog) { iStarvationLog = aLog; }
    // Returns 0, or -2 for a spec the message model cannot represent.  The API's AssertionFailed propagates.
    int Run()
    {
        if (iJps == 0 || iFrameBytes == 0) return -2;
        if (iSpec.channels > 32u) return -2; // ohp_chunk_desc::channels is 8 bits wide and checked against 1..32
        if (iSpec.chunk_frames == 0 || (uint64_t)iSpec.chunk_frames * iFrameBytes > OHP_MAX_PCM_CHUNK_BYTES) return -2;
        if (iSpec.total_frames > (~0ull) / iFrameBytes) return -2;
        for (unsigned i = 0; i < OHP_MAX_STAGES; i++) {
            iStages[i].elem = StageElement(i);
            if (iStages[i].elem == ElemBad) return -2;
            if (iStages[i].elem == ElemStarvation) iStages[i].maxMsg = 5 * OHP_JIFFIES_PER_MS; // kMaxAudioOutJiffies, StarvationRamper.cpp:376
        }
        uint64_t frame = 0;
        uint64_t srcJiffies = 0;
        uint32_t silEv = 0;
        core::CodecSource source; // message sizes as the codec, CodecController and DecodedAudioAggregator leave them
        core::codec_source_init(source, iSpec.chunk_frames, iSpec.codec_read_frames, iFrameBytes, iJps, iSpec.total_frames);
        for (;;) {
            const uint32_t frames = core::codec_source_next(source);
            if (frames == 0) break;
            for (; silEv < iNumEvents; silEv++) {
                const ohp_ramp_event& e = iEvents[silEv];
                if (e.op != OHP_EV_INSERT_SILENCE) continue;
                if (e.at_jiffies > srcJiffies) break;
                uint32_t jiffies = e.arg;
                Item it{Api::CreateSilence(iFactory, iSpec, jiffies), true};
                Feed(0, it);
            }
            Item it{Api::CreatePcm(iFactory, iSpec, frame, frames), false};
            Feed(0, it);
            frame += frames;
            srcJiffies += (uint64_t)frames * iJps;
        }
        // A reservoir that runs dry exactly where the stream ends still starves its StarvationRamper (there is no message left
        // for the event to be applied ahead of, and nothing of the stream's own audio for it to change): recorded
        for (unsigned i = 0; i < OHP_MAX_STAGES && iErr == 0 && iStarvationLog != nullptr; i++) {
            Stage& s = iStages[i];
            uint32_t ei;
            while (s.elem == ElemStarvation && NextStageEvent(i, ei) && iEvents[ei].at_jiffies <= s.pos) {
                NoteStarvation(s, ei);
                (void)ApplyElementEvent(s, iEvents[ei]); // OHP_EV_HALT / OHP_EV_STARVATION: state only
                s.nextEv = ei + 1;
            }
        }
        return iErr;
    }

private:
    bool NextStageEvent(unsigned aStage, uint32_t& aIndex)
    {
        Stage& s = iStages[aStage];
        while (s.nextEv < iNumEvents) {
            const ohp_ramp_event& e = iEvents[s.nextEv];
            if (e.stage == aStage && e.op != OHP_EV_INSERT_SILENCE) {
                aIndex = s.nextEv;
                return true;
            }
            s.nextEv++;
        }
        return false;
    }
    // A stage whose events use an element's own ops IS that element from the first message on; the bare ramp ops
    // (OHP_EV_RAMP_DOWN ... OHP_EV_UNMUTE) do not mix with them, two different elements' ops neither.
    Element StageElement(unsigned aStage) const
    {
        Element elem = Generic;
        bool bare = false;
        for (uint32_t i = 0; i < iNumEvents; i++) {
            const ohp_ramp_event& e = iEvents[i];
            if (e.stage != aStage) continue;
            Element of = Generic;
            switch (e.op) {
            case OHP_EV_RAMPER_STREAM: of = ElemRamper; break;
            case OHP_EV_MUTER_MUTE: case OHP_EV_MUTER_UNMUTE: of = ElemMuter; break;
            case OHP_EV_STARVATION: of = ElemStarvation; break;
            case OHP_EV_RAMP_DOWN: case OHP_EV_RAMP_UP: case OHP_EV_MUTE: case OHP_EV_UNMUTE: bare = true; break;
            default: break;
            }
            if (of != Generic) {
                if (elem != Generic && elem != of) return ElemBad;
                elem = of;
            }
        }
        if (elem != Generic && bare) return ElemBad;
        return elem;
    }
    // The element's own calls and control messages.  Returns false where the reference ASSERTS.
    static bool ApplyElementEvent(Stage& s, const ohp_ramp_event& e)
    {
        switch (e.op) {
        case OHP_EV_RAMPER_STREAM: // Ramper::ProcessMsg(MsgDecodedStream), Ramper.cpp:72-93
            if (e.arg != 0) { s.mode = RampingUp; s.current = Api::kRampMin; s.remaining = e.arg; }
            else { s.mode = Running; s.current = Api::kRampMax; s.remaining = 0; }
            break;
        case OHP_EV_MUTER_MUTE: // Muter::Mute, Muter.cpp:57-99 (eMuting and eMuted differ in who is told when, not in the audio)
            if (s.mode == Running) {
                if (s.halted) { s.mode = Muted; }
                else { s.mode = RampingDown; s.remaining = e.arg; s.current = Api::kRampMax; }
            }
            else if (s.mode == RampingUp) {
                if (s.remaining == e.arg) { s.mode = Muted; }
                else { s.mode = RampingDown; s.remaining = e.arg - s.remaining; }
            }
            else return false;
            break;
        case OHP_EV_MUTER_UNMUTE: // Muter::Unmute, Muter.cpp:101-137
            if (s.mode == RampingDown) {
                if (s.remaining == e.arg) { s.mode = Running; }
                else { s.mode = RampingUp; s.remaining = e.arg - s.remaining; }
            }
            else if (s.mode == Muted) {
                if (s.halted) { s.mode = Running; }
                else { s.mode = RampingUp; s.remaining = e.arg; s.current = Api::kRampMin; }
            }
            else return false;
            break;
        case OHP_EV_HALT:
            if (s.elem == ElemRamper) { // Ramper::ProcessMsg(MsgHalt), Ramper.cpp:65-70
                if (s.mode == RampingUp) s.mode = Running;
            }
            else if (s.elem == ElemMuter) { // Muter::ProcessMsg(MsgHalt) -> BeginHalting, then Halted(), Muter.cpp:159-167, 264-280
                if (s.mode == RampingDown) { s.mode = Muted; s.remaining = 0; s.current = Api::kRampMin; }
                s.halted = true;
            }
            else if (s.elem == ElemStarvation) { // StarvationRamper::ProcessMsgOut(MsgHalt), StarvationRamper.cpp:728-736
                s.mode = Running;
                s.halted = true;
            }
            break;
        case OHP_EV_STARVATION: // StarvationRamper::Pull on an empty reservoir, StarvationRamper.cpp:622-673
            if ((s.mode == Running && !s.halted) || (s.mode == RampingUp && s.current != Api::kRampMin)) {
                s.mode = RampingUp; s.current = Api::kRampMin; s.remaining = e.arg;
            }
            break;
        default: break;
        }
        return true;
    }
    // What a MsgSilence passing through does to the element (the message itself is handed on untouched).
    static void ElementSeesSilence(Stage& s)
    {
        if (s.elem == ElemRamper) { // Ramper.cpp:106-112
            s.mode = Running; s.current = Api::kRampMax; s.remaining = 0;
        }
        else if (s.elem == ElemMuter) { // Muter.cpp:188-208
            if (s.mode == RampingDown) { s.mode = Muted; s.remaining = 0; s.current = Api::kRampMin; }
            else if (s.mode == RampingUp) { s.mode = Running; s.remaining = 0; s.current = Api::kRampMax; }
        }
        // StarvationRamper::ProcessMsgOut(MsgSilence), StarvationRamper.cpp:893-911: a ramp up in progress waits for the next audio
    }
    static void ApplyEvent(Stage& s, const ohp_ramp_event& e)
    {
        switch (e.op) {
        case OHP_EV_RAMP_DOWN:
            if (s.mode == Muted || s.current == Api::kRampMin) { s.mode = Muted; s.current = Api::kRampMin; s.remaining = 0; }
            else { s.mode = RampingDown; s.remaining = e.arg; }
            break;
        case OHP_EV_RAMP_UP:
            if (s.mode == Running && s.current == Api::kRampMax) { /* already at full level */ }
            else { s.mode = RampingUp; s.remaining = e.arg; }
            break;
        case OHP_EV_MUTE: s.mode = Muted; s.current = Api::kRampMin; s.remaining = 0; break;
        case OHP_EV_UNMUTE: s.mode = Running; s.current = Api::kRampMax; s.remaining = 0; break;
        case OHP_EV_SET_ATTENUATION: s.attenuation = e.arg; break;
        case OHP_EV_MAX_MSG_JIFFIES: s.maxMsg = e.arg; break;
        default: Api::Assert(ApplyElementEvent(s, e)); break;
        }
    }
    void Process(unsigned aStage, Item& aItem)
    {
        Stage& s = iStages[aStage];
        MsgAudio* msg = aItem.msg;
        uint32_t ei;
        while (NextStageEvent(aStage, ei) && iEvents[ei].at_jiffies <= s.pos) {
            NoteStarvation(s, ei);
            ApplyEvent(s, iEvents[ei]);
            s.nextEv = ei + 1;
        }
        if (NextStageEvent(aStage, ei) && iEvents[ei].at_jiffies < s.pos + msg->Jiffies()) {
            uint32_t at = (uint32_t)(iEvents[ei].at_jiffies - s.pos);
            if (aItem.silence) at -= at % iJps; // silence only splits on sample blocks
            if (at == 0) {
                NoteStarvation(s, ei);
                ApplyEvent(s, iEvents[ei]);
                s.nextEv = ei + 1;
            }
            else {
                s.queue.push_front(Item{msg->Split(at), aItem.silence});
            }
        }
        if (s.maxMsg != 0 && msg->Jiffies() > s.maxMsg) {
            if (s.maxMsg < iJps) { iErr = -2; return; }
            s.queue.push_front(Item{msg->Split(s.maxMsg), aItem.silence});
        }
        if (!aItem.silence && s.attenuation != OHP_UNITY_ATTENUATION) {
            static_cast<typename Api::MsgAudioPcm*>(msg)->SetAttenuation(s.attenuation);
        }
        if (s.elem != Generic) {
            if (aItem.silence) {
                ElementSeesSilence(s);
                s.pos += msg->Jiffies();
                s.pcmRun = 0;
                if (s.elem == ElemStarvation) NoteRecent(s, Recent{0, msg->Jiffies(), 1u, OHP_UNITY_ATTENUATION});
                return;
            }
            s.halted = false; // Muter::ProcessAudio, Muter.cpp:212; StarvationRamper::ProcessMsgOut(MsgAudioPcm), StarvationRamper.cpp:797-799
        }
        if (s.mode == RampingDown || s.mode == RampingUp) {
            if (s.remaining > 0) {
                if (msg->Jiffies() > s.remaining) {
                    s.queue.push_front(Item{msg->Split(s.remaining), aItem.silence});
                    Api::Assert(msg->Jiffies() != 0); // see oracle/ohp_oracle.c stage_process
                }
                MsgAudio* split = nullptr;
                s.current = msg->SetRamp(s.current, s.remaining, s.mode == RampingDown ? Api::kDirDown : Api::kDirUp, split);
                s.elemRamp = s.current;
                if (split != nullptr) {
                    s.queue.push_front(Item{split, aItem.silence});
                }
            }
            if (s.remaining == 0) {
                if (s.mode == RampingUp) { s.mode = Running; s.current = Api::kRampMax; }
                else { s.mode = Muted; s.current = Api::kRampMin; }
            }
        }
        else if (s.mode == Muted) {
            msg->SetMuted();
        }
        s.pos += msg->Jiffies();
        if (!aItem.silence) {
            if (s.elem == ElemStarvation) {
                // the attenuation the message carries: set by the last stage up to this one that has one (every stage whose
                // attenuation is not unity calls SetAttenuation, above; a message is handed down the chain as soon as a stage
                // lets go of it, so the stages before this one are still as they were when it passed them)
                uint32_t att = OHP_UNITY_ATTENUATION;
                for (unsigned i = aStage + 1; i-- > 0;) {
                    if (iStages[i].attenuation != OHP_UNITY_ATTENUATION) { att = iStages[i].attenuation; break; }
                }
                if (att != s.pcmRunAtt) { s.pcmRun = 0; s.pcmRunAtt = att; }
                NoteRecent(s, Recent{s.pcmJiffies, msg->Jiffies(), 0u, att});
            }
            s.pcmJiffies += msg->Jiffies();
            s.pcmRun += msg->Jiffies();
        }
    }
    // StarvationRamper keeps a clone of every MsgAudioPcm and MsgSilence it hands on, trimmed to about 1 ms
    // (ProcessAudioOut, StarvationRamper.cpp:548-577); StartFlywheelRamp (:491-536) cuts that to the last kTrainingJiffies,
    // reads it through FlywheelInput and leaves it empty.  Only that and a new stream (NewStream, :539-546) empty it: a
    // MsgHalt does not.  pcmRun is how much of its tail is PCM and nothing else.
    // ProcessAudioOut (StarvationRamper.cpp:548-577): the clone joins the queue; pieces the last millisecond no longer needs go
    void NoteRecent(Stage& s, const Recent& aPiece)
    {
        if (aPiece.jiffies == 0) return;
        if (!aPiece.silence && !s.recent.empty()) {
            Recent& last = s.recent.back();
            if (!last.silence && last.attenuation == aPiece.attenuation && last.pcmJiffies + last.jiffies == aPiece.pcmJiffies
                && (uint64_t)last.jiffies + aPiece.jiffies <= 0x7fffffffu) {
                last.jiffies += aPiece.jiffies;
                s.recentJiffies += aPiece.jiffies;
                TrimRecent(s);
                return;
            }
        }
        s.recent.push_back(aPiece);
        s.recentJiffies += aPiece.jiffies;
        TrimRecent(s);
    }
    static void TrimRecent(Stage& s)
    {
        while (s.recent.size() > 1 && s.recentJiffies - s.recent.front().jiffies >= OHP_FLYWHEEL_TRAINING_JIFFIES) {
            s.recentJiffies -= s.recent.front().jiffies;
            s.recent.pop_front();
        }
        // a long run of PCM: only its tail matters (kept generously: a MsgAudioPcm can be cut anywhere)
        Recent& front = s.recent.front();
        if (!front.silence && s.recent.size() == 1 && front.jiffies > 4u * OHP_FLYWHEEL_TRAINING_JIFFIES) {
            const uint32_t drop = front.jiffies - 2u * OHP_FLYWHEEL_TRAINING_JIFFIES;
            front.pcmJiffies += drop;
            front.jiffies -= drop;
            s.recentJiffies -= drop;
        }
    }
    void NoteStarvation(Stage& s, uint32_t aEvent)
    {
        const ohp_ramp_event& e = iEvents[aEvent];
        if (s.elem != ElemStarvation) return;
        if (e.op != OHP_EV_STARVATION) return;
        const bool plays = (s.mode == Running && !s.halted) || (s.mode == RampingUp && s.current != Api::kRampMin); // as ApplyElementEvent
        if (iStarvationLog != nullptr) {
            Starvation st;
            st.event = aEvent;
            st.plays = plays ? 1u : 0u;
            st.ramp = s.elemRamp;
            st.recentJiffies = s.pcmRun > 0xffffffffull ? 0xffffffffu : (uint32_t)s.pcmRun;
            st.attenuation = s.pcmRunAtt;
            st.pcmJiffies = s.pcmJiffies;
            st.recent.assign(s.recent.begin(), s.recent.end());
            iStarvationLog->push_back(st);
        }
        if (plays) { s.pcmRun = 0; s.elemRamp = Api::kRampMin; s.recent.clear(); s.recentJiffies = 0; } // Pull(): FlywheelRamping -> RampingUp from kMin (:651-656)
    }
    void Feed(unsigned aStage, Item aItem)
    {
        if (iErr) { aItem.msg->RemoveRef(); return; }
        if (aStage == OHP_MAX_STAGES) { Drive(aItem); return; }
        Stage& s = iStages[aStage];
        s.queue.push_back(aItem);
        while (!s.queue.empty() && !iErr) {
            Item it = s.queue.front();
            s.queue.pop_front();
            try {
                Process(aStage, it);
            }
            catch (...) {
                it.msg->RemoveRef();
                throw;
            }
            if (iErr) { it.msg->RemoveRef(); return; }
            Feed(aStage + 1, it);
        }
    }
    void Drive(Item& aItem)
    {
        MsgPlayable* playable = aItem.msg->CreatePlayable(); // consumes the msg's reference
        const uint32_t block = iSpec.driver_block_frames * iFrameBytes;
        if (block == 0 || playable->Bytes() == 0) {
            iSink.OnPlayable(playable);
            return;
        }
        for (;;) {
            const uint32_t room = block - iBlockFill;
            if (playable->Bytes() > room) {
                MsgPlayable* remaining = playable->Split(room);
                iSink.OnPlayable(playable);
                iBlockFill = 0;
                playable = remaining;
            }
            else {
                const uint32_t bytes = playable->Bytes();
                iSink.OnPlayable(playable);
                iBlockFill += bytes;
                if (iBlockFill == block) iBlockFill = 0;
                return;
            }
        }
    }

private:
    typename Api::Factory& iFactory;
    const ohp_stream_spec& iSpec;
    const ohp_ramp_event* iEvents;
    uint32_t iNumEvents;
    Sink& iSink;
    uint32_t iJps;
    uint32_t iFrameBytes;
    uint32_t iBlockFill;
    int iErr = 0;
    std::vector<Starvation>* iStarvationLog = nullptr;
    Stage iStages[OHP_MAX_STAGES];
};

} // namespace ohp
