// ref_harness.cpp -- LINKED-REFERENCE ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// Compiled together with the reference's own, unmodified sources
//     /root/reference/OpenHome/Media/Pipeline/Msg.cpp
//     /root/reference/OpenHome/Media/Utils/ProcessorAudioUtils.cpp
//     /root/reference/OpenHome/Media/FlywheelRamper.cpp
//     /root/reference/OpenHome/Media/Pipeline/StarvationRamper.cpp   (for FlywheelInput and RampGenerator)
//     /root/reference/OpenHome/Media/Pipeline/DecodedAudioAggregator.cpp
// (against the ohNet header shim in oracle/shim/) into oracle/_ref/libohref.so by oracle/Makefile.
// It calls the reference's public API the way the reference's own unit tests do
// (Media/Tests/TestMsg.cpp SuiteRamp / SuiteMsgPlayable): MsgFactory::CreateMsgAudioPcm ->
// MsgAudio::Split/SetRamp/SetMuted -> CreatePlayable -> MsgPlayable::Split -> Read(ProcessorPcmBufTest).
// No reference source is copied here.  Built with -fno-access-control so the harness can read the
// private members (iOffset, iAudioData, iAttenuation) that make up a chunk descriptor.
//
// Used to (1) validate the plain-C restatement in ohp_oracle.c, (2) generate tests/golden/,
// (3) serve as the CPU baseline ("kind": "reference") that bench.py times on the GPU box's host cores.

#include <OpenHome/Media/Pipeline/Msg.h>
#include <OpenHome/Media/Utils/ProcessorAudioUtils.h>
#include <OpenHome/Media/Pipeline/RampArray.h>
#include <OpenHome/Media/Pipeline/StarvationRamper.h>
#include <OpenHome/Media/FlywheelRamper.h>
#include <OpenHome/Media/Pipeline/DecodedAudioAggregator.h>

#include <cstring>
#include <cstdlib>
#include <thread>
#include <unordered_map>
#include <vector>

#include "ohp_oracle.h"
#include "../ohpipeline_b200/host/stage_chain.h"

using namespace OpenHome;
using namespace OpenHome::Media;

namespace {

class NullInfoAggregator : public IInfoAggregator
{
public:
    void Register(IInfoProvider& /*aProvider*/, std::vector<Brn>& /*aSupportedQueries*/) override {}
};

struct RefFactory
{
    NullInfoAggregator info;
    MsgFactory* factory;
    const uint8_t* pcm; // stream's PCM base (wire format)
    std::unordered_map<const DecodedAudio*, uint64_t> cellSrc;
    RefFactory()
    {
        MsgFactoryInitParams p;
        p.SetMsgAudioPcmCount(64, 64);
        p.SetMsgSilenceCount(64);
        p.SetMsgPlayableCount(16, 1, 16);
        factory = new MsgFactory(info, p);
        pcm = nullptr;
        // MsgPlayablePcm's constructor leaves iAttenuation uninitialised (Msg.cpp:2719-2722); only Clear()
        // sets it to unity (Msg.cpp:2809-2814).  MsgPlayable::Split hands out such a cell as the remainder
        // without initialising it (Msg.cpp:2605, 2803-2807), so a never-used pool cell carries garbage.
        // A running pipeline has cycled its 10 playables long ago; cycle ours once for the same steady state.
        static const TByte frame[4] = {0, 0, 0, 0};
        MsgPlayable* warm[16];
        for (TUint i = 0; i < 16; i++) {
            warm[i] = factory->CreateMsgAudioPcm(Brn(frame, 4), 2, 44100, 16, AudioDataEndian::Big, 0)->CreatePlayable();
        }
        for (TUint i = 0; i < 16; i++) {
            warm[i]->RemoveRef();
        }
    }
    ~RefFactory() { delete factory; }
};

struct RefApi
{
    using MsgAudio = OpenHome::Media::MsgAudio;
    using MsgAudioPcm = OpenHome::Media::MsgAudioPcm;
    using MsgSilence = OpenHome::Media::MsgSilence;
    using MsgPlayable = OpenHome::Media::MsgPlayable;
    using Factory = RefFactory;
    static const uint32_t kRampMax = Ramp::kMax;
    static const uint32_t kRampMin = Ramp::kMin;
    static const Ramp::EDirection kDirUp = Ramp::EUp;
    static const Ramp::EDirection kDirDown = Ramp::EDown;
    static MsgAudioPcm* CreatePcm(Factory& f, const ohp_stream_spec& sp, uint64_t firstFrame, uint32_t frames)
    {
        const uint32_t frameBytes = sp.channels * (sp.bit_depth / 8u);
        const uint64_t off = firstFrame * frameBytes;
        Brn data(f.pcm + off, frames * frameBytes);
        MsgAudioPcm* msg = f.factory->CreateMsgAudioPcm(data, sp.channels, sp.sample_rate, sp.bit_depth,
                                                        sp.in_little_endian ? AudioDataEndian::Little : AudioDataEndian::Big,
                                                        firstFrame * Jiffies::PerSample(sp.sample_rate));
        f.cellSrc[msg->iAudioData] = off;
        return msg;
    }
    static MsgSilence* CreateSilence(Factory& f, const ohp_stream_spec& sp, uint32_t& jiffies)
    {
        return f.factory->CreateMsgSilence(jiffies, sp.sample_rate, sp.bit_depth, sp.channels);
    }
    static uint32_t JiffiesPerSample(uint32_t rate)
    {
        try { return Jiffies::PerSample(rate); }
        catch (SampleRateInvalid&) { return 0; }
    }
    static void Assert(bool ok) { ASSERT(ok); }
};

struct StreamOut
{
    std::vector<ohp_chunk_desc> chunks;
    std::vector<ohp_chunk_info> info;
    uint64_t outBytes = 0;
};

class RefSink
{
public:
    RefSink(RefFactory& f, const ohp_stream_spec& sp, uint8_t* out, StreamOut* rec)
        : iFactory(f), iSpec(sp), iOut(out), iRec(rec), iOutBytes(0) {}
    void OnPlayable(MsgPlayable* p)
    {
        const uint32_t bytes = p->Bytes();
        if (iRec != nullptr) {
            ohp_chunk_desc d;
            std::memset(&d, 0, sizeof d);
            MsgPlayablePcm* pcm = dynamic_cast<MsgPlayablePcm*>(p);
            d.dst_off = iSpec.dst_base + iOutBytes;
            d.bytes = bytes;
            d.ramp_start = (uint16_t)p->Ramp().Start();
            d.ramp_end = (uint16_t)p->Ramp().End();
            d.bit_depth = (uint8_t)iSpec.bit_depth;
            d.channels = (uint8_t)iSpec.channels;
            d.out_fmt = (uint8_t)iSpec.out_fmt;
            d.aux = iSpec.out_fmt == OHP_OUT_PACKED_LE ? OHP_LE_APPEND : 0;
            d.flags = p->Ramp().IsEnabled() ? OHP_F_RAMP_ENABLED : 0;
            if (pcm != nullptr) {
                d.src_off = iSpec.src_base + iFactory.cellSrc[pcm->iAudioData] + pcm->iOffset;
                d.attenuation = (uint16_t)pcm->iAttenuation;
                if (iSpec.in_little_endian) d.flags |= OHP_F_IN_LITTLE_ENDIAN;
            }
            else {
                d.src_off = 0;
                d.attenuation = OHP_UNITY_ATTENUATION;
                d.flags |= OHP_F_SILENCE;
            }
            ohp_chunk_info ci;
            ci.direction = (uint32_t)p->Ramp().Direction();
            ci.jiffies = p->Jiffies();
            iRec->chunks.push_back(d);
            iRec->info.push_back(ci);
        }
        if (iOut != nullptr) {
            p->Read(iProc);
            ASSERT(iProc.Buf().Bytes() == bytes);
            if (bytes > 0) std::memcpy(iOut + iSpec.dst_base + iOutBytes, iProc.Ptr(), bytes);
        }
        iOutBytes += bytes;
        p->RemoveRef();
    }
    uint64_t OutBytes() const { return iOutBytes; }
private:
    RefFactory& iFactory;
    const ohp_stream_spec& iSpec;
    uint8_t* iOut;
    StreamOut* iRec;
    uint64_t iOutBytes;
    ProcessorPcmBufTest iProc;
};

// One stream through the real message model.  Returns 0, -1 on AssertionFailed, -2 bad spec.
int RunStream(RefFactory*& f, const ohp_stream_spec& sp, const ohp_ramp_event* events,
              const uint8_t* in, uint8_t* out, StreamOut* rec, uint64_t* outBytes)
{
    f->pcm = in + sp.src_base;
    f->cellSrc.clear();
    RefSink sink(*f, sp, out, rec);
    int rc;
    try {
        ohp::StageChain<RefApi, RefSink> chain(*f, sp, events + sp.first_event, sink);
        rc = chain.Run();
    }
    catch (Exception& e) { // AssertionFailed and the typed THROW()s
        if (std::getenv("OHP_REF_TRACE") != nullptr) {
            std::fprintf(stderr, "ref: %s at %s:%u\n", e.Message(), e.File(), e.Line());
        }
        rc = -1;
    }
    if (rc != 0) {
        // messages in flight when an ASSERT unwound may have leaked; the reference's allocators assert on
        // destruction with cells outstanding, so abandon this factory rather than destroy it
        f = new RefFactory();
    }
    *outBytes = sink.OutBytes();
    return rc;
}

ohp_ramp FromRamp(const Ramp& r)
{
    ohp_ramp o;
    o.start = r.Start();
    o.end = r.End();
    o.direction = (uint32_t)r.Direction();
    o.enabled = r.IsEnabled() ? 1u : 0u;
    return o;
}
void ToRamp(const ohp_ramp& o, Ramp& r)
{
    r.iStart = o.start;
    r.iEnd = o.end;
    r.iDirection = (Ramp::EDirection)o.direction;
    r.iEnabled = o.enabled != 0;
}

} // namespace

extern "C" {

uint32_t ref_jiffies_per_sample(uint32_t rate) { return RefApi::JiffiesPerSample(rate); }

const uint32_t* ref_ramp_array(void) { return (const uint32_t*)kRampArray; }
uint32_t ref_ramp_array_count(void) { return kRampArrayCount; }

int ref_ramp_set(ohp_ramp* ramp, uint32_t start, uint32_t frag, uint32_t dur, uint32_t dir, ohp_ramp* split, uint32_t* splitPos)
{
    Ramp r, s;
    ToRamp(*ramp, r);
    TUint pos = 0;
    try {
        const TBool ret = r.Set(start, frag, dur, (Ramp::EDirection)dir, s, pos);
        *ramp = FromRamp(r);
        *split = FromRamp(s);
        *splitPos = pos;
        return ret ? 1 : 0;
    }
    catch (AssertionFailed&) {
        return -1;
    }
}

int ref_ramp_split(ohp_ramp* ramp, uint32_t newSize, uint32_t curSize, ohp_ramp* remaining)
{
    if (curSize == 0) return -1; // integer divide by zero in the reference
    Ramp r;
    ToRamp(*ramp, r);
    try {
        Ramp rem = r.Split(newSize, curSize);
        *ramp = FromRamp(r);
        *remaining = FromRamp(rem);
        return 0;
    }
    catch (AssertionFailed&) {
        return -1;
    }
}

// MsgAudio::MedianRampMultiplier on a real message.  median ramp 0 (index 512) is an out-of-bounds read in the
// reference, so callers avoid that case.
uint32_t ref_median_multiplier(const ohp_ramp* ramp)
{
    static RefFactory* f = new RefFactory();
    static const TByte data[4] = {0, 0, 0, 0};
    MsgAudioPcm* msg = f->factory->CreateMsgAudioPcm(Brn(data, 4), 2, 44100, 16, AudioDataEndian::Big, 0);
    ToRamp(*ramp, msg->iRamp);
    const uint32_t m = msg->MedianRampMultiplier();
    msg->RemoveRef();
    return m;
}

// MsgPlayable::Read(ProcessorPcmBufTest) for each descriptor, through real messages.  Packed-BE sink only
// (the other IPcmProcessors live in files that do not compile stand-alone).
int64_t ref_process_chunks(const ohp_chunk_desc* descs, size_t n, const uint8_t* in, uint64_t in_bytes,
                           uint8_t* out, uint64_t out_bytes)
{
    RefFactory* f = new RefFactory();
    ProcessorPcmBufTest proc;
    int64_t rc = 0;
    for (size_t k = 0; k < n && rc == 0; k++) {
        const ohp_chunk_desc& d = descs[k];
        try {
            if (d.out_fmt != OHP_OUT_PACKED_BE) { rc = -(int64_t)(k + 1); break; }
            if (d.bytes == 0) continue;
            const uint32_t frameBytes = d.channels * (d.bit_depth / 8u);
            MsgPlayable* playable;
            const TUint rate = 192000;
            if (d.flags & OHP_F_SILENCE) {
                TUint jiffies = (d.bytes / frameBytes) * Jiffies::PerSample(rate);
                MsgSilence* msg = f->factory->CreateMsgSilence(jiffies, rate, d.bit_depth, d.channels);
                msg->iRamp.iStart = d.ramp_start;
                msg->iRamp.iEnd = d.ramp_end;
                msg->iRamp.iEnabled = (d.flags & OHP_F_RAMP_ENABLED) != 0;
                playable = msg->CreatePlayable();
            }
            else {
                if (d.src_off + d.bytes > in_bytes) { rc = -(int64_t)(k + 1); break; }
                MsgAudioPcm* msg = f->factory->CreateMsgAudioPcm(Brn(in + d.src_off, d.bytes), d.channels, rate, d.bit_depth,
                        (d.flags & OHP_F_IN_LITTLE_ENDIAN) ? AudioDataEndian::Little : AudioDataEndian::Big, 0);
                msg->iRamp.iStart = d.ramp_start;
                msg->iRamp.iEnd = d.ramp_end;
                msg->iRamp.iEnabled = (d.flags & OHP_F_RAMP_ENABLED) != 0;
                msg->iRamp.iDirection = !msg->iRamp.iEnabled ? Ramp::ENone
                                      : (d.ramp_start < d.ramp_end ? Ramp::EUp : (d.ramp_start > d.ramp_end ? Ramp::EDown : Ramp::ENone));
                msg->SetAttenuation(d.attenuation);
                playable = msg->CreatePlayable();
            }
            try {
                playable->Read(proc);
            }
            catch (AssertionFailed&) {
                playable->RemoveRef();
                throw;
            }
            if (proc.Buf().Bytes() != d.bytes || d.dst_off + d.bytes > out_bytes) {
                rc = -(int64_t)(k + 1);
            }
            else {
                std::memcpy(out + d.dst_off, proc.Ptr(), d.bytes);
            }
            playable->RemoveRef();
        }
        catch (AssertionFailed&) {
            rc = -(int64_t)(k + 1);
            f = new RefFactory(); // see RunStream
        }
    }
    if (rc == 0) delete f;
    return rc;
}

// The whole reference path for a batch of streams, threaded over streams with one MsgFactory per thread.
//   out  == NULL : descriptors only;   res == NULL : audio only (the timed CPU-baseline configuration)
// Returns 0, -1 (AssertionFailed in some stream), -2 (bad spec).
int ref_schedule_run(const ohp_stream_spec* streams, size_t n_streams,
                     const ohp_ramp_event* events, size_t n_events,
                     const uint8_t* in, uint8_t* out, ohpo_schedule_result* res, int threads)
{
    (void)n_events;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    if ((size_t)threads > n_streams) threads = (int)(n_streams ? n_streams : 1);
    std::vector<StreamOut> recs(res ? n_streams : 0);
    std::vector<uint64_t> outBytes(n_streams, 0);
    std::vector<int> rcs(threads, 0);
    auto worker = [&](int t) {
        RefFactory* f = new RefFactory();
        const size_t lo = n_streams * (size_t)t / (size_t)threads;
        const size_t hi = n_streams * (size_t)(t + 1) / (size_t)threads;
        for (size_t s = lo; s < hi; s++) {
            const int rc = RunStream(f, streams[s], events, in, out, res ? &recs[s] : nullptr, &outBytes[s]);
            if (rc != 0) { rcs[t] = rc; break; }
        }
        delete f;
    };
    if (threads == 1) {
        worker(0);
    }
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++) pool.emplace_back(worker, t);
        for (auto& th : pool) th.join();
    }
    for (int rc : rcs) if (rc != 0) return rc;
    if (res != nullptr) {
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
            res->stream_out_bytes[s] = outBytes[s];
            if (!recs[s].chunks.empty()) {
                std::memcpy(res->chunks + k, recs[s].chunks.data(), recs[s].chunks.size() * sizeof(ohp_chunk_desc));
                std::memcpy(res->info + k, recs[s].info.data(), recs[s].info.size() * sizeof(ohp_chunk_info));
            }
            k += recs[s].chunks.size();
        }
        res->stream_chunk_begin[n_streams] = k;
        res->num_chunks = k;
    }
    return 0;
}

// ---- flywheel ramp generator ---------------------------------------------------------------------------------------

// FlywheelRamper::BurgsMethod as the reference's own Test6 calls it (TestFlywheelRamper.cpp:568-608).
void ref_burgs_method(const int16_t* samples, uint32_t n, uint32_t degree, int16_t* output, int16_t* h, int16_t* per, int16_t* pef)
{
    FlywheelRamper::BurgsMethod(const_cast<TInt16*>(samples), n, degree, output, h, per, pef);
}

// FeedbackModel as the reference's Test1-Test5 use it: `count` outputs from the given coefficients and initial states.
void ref_feedback_model(uint32_t states, uint32_t descale, uint32_t coeffFormat, uint32_t dataFormat, uint32_t outFormat,
                        int32_t* coeffs, int32_t* samples, int32_t* out, uint32_t count)
{
    FeedbackModel fb(states, descale, coeffFormat, dataFormat, outFormat);
    fb.Initialise(coeffs, samples);
    for (uint32_t i = 0; i < count; i++) out[i] = fb.NextSample();
}

// FlywheelInput::Prepare (StarvationRamper.cpp:90-111) over a queue of real messages made from `descs`
// (unramped PCM or silence chunks of one stream).  planar receives channels * (jiffies / jps) * 4 bytes.
int ref_flywheel_input(const ohp_chunk_desc* descs, size_t n, const uint8_t* in, uint32_t rate, uint32_t jiffies,
                       uint8_t* planar, uint32_t cap, uint32_t* planarBytes)
{
    RefFactory* f = new RefFactory();
    try {
        MsgQueueLite queue;
        uint32_t bits = 16, ch = 2;
        for (size_t k = 0; k < n; k++) {
            const ohp_chunk_desc& d = descs[k];
            bits = d.bit_depth; ch = d.channels;
            if (d.flags & OHP_F_SILENCE) {
                TUint j = (d.bytes / (d.channels * (d.bit_depth / 8u))) * Jiffies::PerSample(rate);
                queue.Enqueue(f->factory->CreateMsgSilence(j, rate, d.bit_depth, d.channels));
            }
            else {
                queue.Enqueue(f->factory->CreateMsgAudioPcm(Brn(in + d.src_off, d.bytes), d.channels, rate, d.bit_depth,
                              (d.flags & OHP_F_IN_LITTLE_ENDIAN) ? AudioDataEndian::Little : AudioDataEndian::Big, 0));
            }
        }
        FlywheelInput input(OHP_FLYWHEEL_TRAINING_JIFFIES);
        const Brx& buf = input.Prepare(queue, jiffies, rate, bits, ch);
        if (buf.Bytes() > cap) { delete f; return -2; }
        std::memcpy(planar, buf.Ptr(), buf.Bytes());
        *planarBytes = buf.Bytes();
        delete f;
        return 0;
    }
    catch (AssertionFailed&) {
        return -1;
    }
}

// The real RampGenerator (StarvationRamper.cpp:196-371): Start() on a training block in FlywheelInput's layout, then
// every MsgAudioPcm it produces.  raw = the messages' payloads (what FlywheelRamperManager generated, repacked by
// RampGenerator::ProcessFragment); ramped = the same messages through CreatePlayable -> Read(ProcessorPcmBufTest) with
// the ramp RampGenerator::EndBlock set; descs = one per message (src_off / dst_off = offset in raw / ramped).
// Callers keep to jobs ohp_flywheel_validate accepts: an ASSERT on the generator's own thread cannot be caught here.
int ref_flywheel(uint32_t rate, uint32_t channels, uint32_t bits, uint32_t currentRamp,
                 const uint8_t* training, uint32_t trainingBytes,
                 uint8_t* raw, uint8_t* ramped, uint32_t cap, uint32_t* outBytes,
                 ohp_chunk_desc* descs, ohp_chunk_info* info, uint32_t descCap, uint32_t* numDescs, uint32_t* finalRamp)
{
    RefFactory* f = new RefFactory();
    int rc = 0;
    uint32_t bytes = 0, nd = 0;
    try {
        RampGenerator gen(*f->factory, OHP_FLYWHEEL_TRAINING_JIFFIES, OHP_FLYWHEEL_RAMP_JIFFIES, kPriorityNormal);
        Brn recent(training, trainingBytes);
        gen.Start(recent, rate, channels, bits, currentRamp);
        ProcessorPcmBufTest proc;
        Msg* m = nullptr;
        while (gen.TryGetAudio(m)) {
            MsgAudioPcm* pcm = static_cast<MsgAudioPcm*>(m);
            MsgAudioPcm* clone = static_cast<MsgAudioPcm*>(pcm->Clone());
            clone->ClearRamp();
            MsgPlayable* rawPlayable = clone->CreatePlayable();
            rawPlayable->Read(proc);
            const uint32_t n = proc.Buf().Bytes();
            if (bytes + n > cap || nd == descCap) { rc = -2; rawPlayable->RemoveRef(); pcm->RemoveRef(); continue; }
            std::memcpy(raw + bytes, proc.Ptr(), n);
            rawPlayable->RemoveRef();
            const Ramp ramp = pcm->Ramp();
            const uint32_t jiffies = pcm->Jiffies();
            MsgPlayable* playable = pcm->CreatePlayable();
            playable->Read(proc);
            if (proc.Buf().Bytes() != n) rc = -3;
            else std::memcpy(ramped + bytes, proc.Ptr(), n);
            ohp_chunk_desc& d = descs[nd];
            std::memset(&d, 0, sizeof d);
            const bool silence = dynamic_cast<MsgPlayablePcm*>(playable) == nullptr;
            d.src_off = silence ? 0 : bytes;
            d.dst_off = bytes;
            d.bytes = n;
            d.ramp_start = (uint16_t)playable->Ramp().Start();
            d.ramp_end = (uint16_t)playable->Ramp().End();
            d.attenuation = OHP_UNITY_ATTENUATION;
            d.bit_depth = (uint8_t)bits;
            d.channels = (uint8_t)channels;
            d.flags = (playable->Ramp().IsEnabled() ? OHP_F_RAMP_ENABLED : 0) | (silence ? OHP_F_SILENCE : 0);
            d.out_fmt = OHP_OUT_PACKED_BE;
            info[nd].direction = (uint32_t)ramp.Direction();
            info[nd].jiffies = jiffies;
            nd++;
            bytes += n;
            playable->RemoveRef();
        }
        *finalRamp = gen.iCurrentRampValue;
    }
    catch (AssertionFailed&) {
        return -1; // the factory is leaked on purpose (see RunStream)
    }
    *outBytes = bytes;
    *numDescs = nd;
    delete f;
    return rc;
}

// ---- DecodedAudioAggregator -------------------------------------------------------------------------------------------

namespace {
class CollectFrames : public IPipelineElementDownstream
{
public:
    CollectFrames(uint32_t* out, uint32_t cap, uint32_t jps) : iOut(out), iCap(cap), iJps(jps), iCount(0) {}
    void Push(Msg* aMsg) override
    {
        MsgAudioPcm* pcm = dynamic_cast<MsgAudioPcm*>(aMsg);
        if (pcm != nullptr && iCount < iCap) iOut[iCount] = pcm->Jiffies() / iJps;
        if (pcm != nullptr) iCount++;
        aMsg->RemoveRef();
    }
    uint32_t Count() const { return iCount; }
private:
    uint32_t* iOut; uint32_t iCap; uint32_t iJps; uint32_t iCount;
};
} // namespace

// The real DecodedAudioAggregator fed with a MsgDecodedStream and then MsgAudioPcm messages of the given frame counts
// (what CodecController queues); out receives the frame counts of the messages it passes on, a final MsgQuit flushes.
int ref_aggregate(uint32_t rate, uint32_t channels, uint32_t bits, const uint32_t* frames, uint32_t n, uint32_t* out, uint32_t cap)
{
    NullInfoAggregator info;
    MsgFactoryInitParams p;
    p.SetMsgAudioPcmCount(64, 64);
    p.SetMsgDecodedStreamCount(2);
    MsgFactory factory(info, p);
    try {
        const uint32_t jps = Jiffies::PerSample(rate);
        CollectFrames sink(out, cap, jps);
        DecodedAudioAggregator agg(sink);
        SpeakerProfile profile(2); // the aggregator does not look at it (its constructor ASSERTs above 3 fronts)
        agg.Push(factory.CreateMsgDecodedStream(0, 0, bits, rate, channels, Brn("PCM"), 0, 0, true, true, false, false,
                                                AudioFormat::Pcm, Multiroom::Allowed, profile, nullptr, RampType::Sample));
        static TByte cell[AudioData::kMaxBytes];
        TUint64 offset = 0;
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t bytes = frames[i] * channels * (bits / 8);
            MsgAudioPcm* msg = factory.CreateMsgAudioPcm(Brn(cell, bytes), channels, rate, bits, AudioDataEndian::Big, offset);
            offset += msg->Jiffies();
            agg.Push(msg);
        }
        agg.Push(factory.CreateMsgQuit());
        return (int)sink.Count();
    }
    catch (AssertionFailed& e) {
        if (std::getenv("OHP_REF_VERBOSE")) std::fprintf(stderr, "ref_aggregate: ASSERT at %s:%u\n", e.File(), e.Line());
        return -1;
    }
}

int ref_hardware_threads(void)
{
    const unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

} // extern "C"
