// test_host_api.cpp -- reference-style tests of the C++ host mirror (ohpipeline_b200/host): the assertions of the
// reference's SuiteRamp / SuiteMsgPlayable / SuiteMsgAudio (OpenHome/Media/Tests/TestMsg.cpp), restated against
// ohp::media::{Ramp, MsgFactory, MsgAudioPcm, MsgSilence, MsgPlayable, IPcmProcessor}.
//
//   test_host_api          CPU part: ramp algebra, message model, assertion behaviour (no GPU needed)
//   test_host_api --gpu    additionally reads playables through BatchPcmReader on cuda:0 and compares every byte with
//                          the C oracle (oracle/libohp_oracle.so, loaded with dlopen -- test infrastructure)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>

#include "../../ohpipeline_b200/host/batch_reader.h"

using namespace ohp;
using namespace ohp::media;

static int gChecks = 0, gFailures = 0;
#define TEST(x) do { gChecks++; if (!(x)) { gFailures++; std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #x); } } while (0)
#define TEST_THROWS(expr, Ex) do { gChecks++; bool t_ = false; try { expr; } catch (Ex&) { t_ = true; } \
    if (!t_) { gFailures++; std::printf("FAILED %s:%d  %s did not throw " #Ex "\n", __FILE__, __LINE__, #expr); } } while (0)

// an arena that just hands out offsets (CPU tests never touch audio bytes)
class NullArena : public IInputArena
{
public:
    uint64_t Stage(const Brx& aData) override { const uint64_t at = iUsed; iUsed += (aData.Bytes() + 15u) & ~15u; return at; }
    void Pin() override { pins++; }
    void Unpin() override { pins--; }
    long pins = 0; // live messages that refer to staged audio
private:
    uint64_t iUsed = 0;
};

static void SuiteRampAlgebra()
{
    // TestMsg.cpp:1393-1443
    const uint32_t jiffies = Jiffies::kPerMs;
    Ramp ramp, split;
    uint32_t splitPos;
    TEST(!ramp.Set(Ramp::kMax, jiffies, jiffies, Ramp::EDown, split, splitPos));
    TEST(ramp.Start() == Ramp::kMax); TEST(ramp.End() == Ramp::kMin); TEST(ramp.Direction() == Ramp::EDown);
    ramp.Reset();
    TEST_THROWS(ramp.Set(Ramp::kMax, jiffies, jiffies, Ramp::EUp, split, splitPos), AssertionFailed);
    ramp.Reset();
    TEST(!ramp.Set(Ramp::kMin, jiffies, jiffies, Ramp::EUp, split, splitPos));
    TEST(ramp.Start() == Ramp::kMin); TEST(ramp.End() == Ramp::kMax); TEST(ramp.Direction() == Ramp::EUp);
    ramp.Reset();
    TEST(!ramp.Set(Ramp::kMax, jiffies, 2 * jiffies, Ramp::EDown, split, splitPos));
    TEST(ramp.End() == (Ramp::kMax - Ramp::kMin) / 2);
    ramp.Reset();
    TEST(!ramp.Set(Ramp::kMin, jiffies, 2 * jiffies, Ramp::EUp, split, splitPos));
    TEST(ramp.End() == (Ramp::kMax - Ramp::kMin) / 2);
    ramp.Reset();
    uint32_t start = (Ramp::kMax - Ramp::kMin) / 2;
    TEST(!ramp.Set(start, jiffies, 2 * jiffies, Ramp::EDown, split, splitPos));
    TEST(ramp.End() == (Ramp::kMax - Ramp::kMin) / 4);
    ramp.Reset();
    TEST(!ramp.Set(start, jiffies, 2 * jiffies, Ramp::EUp, split, splitPos));
    TEST(ramp.End() == Ramp::kMax - ((Ramp::kMax - Ramp::kMin) / 4));
    // TestMsg.cpp:1593-1625: opposing ramps split at the crossing
    ramp.Reset();
    TEST(!ramp.Set(Ramp::kMax / 2, jiffies, jiffies, Ramp::EDown, split, splitPos));
    TEST(ramp.Set(Ramp::kMin, jiffies, 2 * jiffies, Ramp::EUp, split, splitPos));
    TEST(ramp.Start() == 0); TEST(ramp.End() == Ramp::kMax / 4); TEST(ramp.Direction() == Ramp::EUp);
    TEST(split.Start() == ramp.End()); TEST(split.End() == 0); TEST(split.Direction() == Ramp::EDown);
    TEST(ramp.IsEnabled()); TEST(split.IsEnabled());
    ramp.Reset();
    TEST(!ramp.Set(Ramp::kMax / 2, jiffies, 4 * jiffies, Ramp::EDown, split, splitPos));
    const uint32_t s0 = ramp.Start(), e0 = ramp.End();
    TEST(!ramp.Set((uint32_t)(((uint64_t)10 * Ramp::kMax) / 7), jiffies, (5 * jiffies) / 2, Ramp::EDown, split, splitPos));
    TEST(ramp.Start() == s0); TEST(ramp.End() == e0);
    ramp.Reset();
    TEST(!ramp.Set(Ramp::kMax / 2, jiffies, 2 * jiffies, Ramp::EDown, split, splitPos));
    start = (uint32_t)(((uint64_t)2 * Ramp::kMax) / 5);
    TEST(!ramp.Set(start, jiffies, jiffies, Ramp::EDown, split, splitPos));
    TEST(ramp.Start() == start); TEST(ramp.End() == 0); TEST(ramp.Direction() == Ramp::EDown);
}

static void SuiteMsgModel()
{
    NullArena arena;
    MsgFactory factory(&arena);
    uint8_t data[1024];
    std::memset(data, 0x7f, sizeof data);
    // MsgAudioPcm sizes: 2 ch / 16 bit / 44.1 kHz, 64 frames
    MsgAudioPcm* msg = factory.CreateMsgAudioPcm(Brn(data, 256), 2, 44100, 16, AudioDataEndian::Big, 0);
    const uint32_t jps = Jiffies::PerSample(44100);
    TEST(jps == 1280);
    TEST(msg->Jiffies() == 64 * jps);
    // Split at a non-sample boundary: playables round down (TestMsg.cpp:1133-1236)
    MsgAudio* rest = msg->Split(10 * jps + 100);
    TEST(msg->Jiffies() == 10 * jps + 100);
    TEST(rest->Jiffies() == 54 * jps - 100);
    TEST(arena.pins == 2);                    // both halves refer to the staged cell (Msg.cpp:2279-2285)
    MsgPlayable* p1 = msg->CreatePlayable();
    MsgPlayable* p2 = rest->CreatePlayable();
    TEST(arena.pins == 2);                    // the playables took the messages' place (Msg.cpp:2260)
    TEST(p1->Bytes() == 10 * 4);
    TEST(p2->Bytes() == 54 * 4);              // offset rounded down to frame 10, size extended by what the offset lost
    TEST(p2->Descriptor(0, 0).src_off == p1->Descriptor(0, 0).src_off + 40);
    p1->RemoveRef();
    TEST(arena.pins == 1);
    p2->RemoveRef();
    TEST(arena.pins == 0);
    {
        // muted audio becomes a silence playable that refers to no audio; silence never pins; MsgPlayable::Split pins
        MsgAudioPcm* m2 = factory.CreateMsgAudioPcm(Brn(data, 256), 2, 44100, 16, AudioDataEndian::Big, 0);
        m2->SetMuted();
        MsgPlayable* q = m2->CreatePlayable();
        TEST(q->IsSilence() && arena.pins == 0);
        q->RemoveRef();
        uint32_t sj = 100 * jps;
        MsgPlayable* sp = factory.CreateMsgSilence(sj, 44100, 16, 2)->CreatePlayable();
        MsgPlayable* sp2 = sp->Split(40);
        TEST(arena.pins == 0);
        sp->RemoveRef(); sp2->RemoveRef();
        MsgPlayable* a = factory.CreateMsgAudioPcm(Brn(data, 256), 2, 44100, 16, AudioDataEndian::Big, 0)->CreatePlayable();
        MsgPlayable* b = a->Split(40);
        TEST(arena.pins == 2);
        a->RemoveRef(); b->RemoveRef();
        TEST(arena.pins == 0);
    }
    // 1-jiffy split -> 0 bytes
    msg = factory.CreateMsgAudioPcm(Brn(data, 256), 2, 44100, 16, AudioDataEndian::Big, 0);
    rest = msg->Split(1);
    p1 = msg->CreatePlayable();
    TEST(p1->Bytes() == 0);
    p1->RemoveRef();
    rest->RemoveRef();
    // invalid arguments assert
    TEST_THROWS(factory.CreateMsgAudioPcm(Brn(data, 255), 2, 44100, 16, AudioDataEndian::Big, 0), AssertionFailed);
    TEST_THROWS(factory.CreateMsgAudioPcm(Brn(data, 256), 2, 44101, 16, AudioDataEndian::Big, 0), SampleRateInvalid);
    msg = factory.CreateMsgAudioPcm(Brn(data, 256), 2, 44100, 16, AudioDataEndian::Big, 0);
    TEST_THROWS(msg->Split(0), AssertionFailed);
    TEST_THROWS(msg->Split(msg->Jiffies()), AssertionFailed);
    // SetRamp over two messages closes exactly (TestMsg.cpp:1690-1703, silence 17 ms + 23 ms)
    msg->RemoveRef();
    uint32_t j17 = 17 * Jiffies::kPerMs, j23 = 23 * Jiffies::kPerMs;
    MsgSilence* sil1 = factory.CreateMsgSilence(j17, 44100, 8, 2);
    MsgSilence* sil2 = factory.CreateMsgSilence(j23, 44100, 8, 2);
    uint32_t remaining = sil1->Jiffies() + sil2->Jiffies();
    MsgAudio* split = nullptr;
    uint32_t cur = sil1->SetRamp(Ramp::kMax, remaining, Ramp::EDown, split);
    TEST(split == nullptr);
    cur = sil2->SetRamp(cur, remaining, Ramp::EDown, split);
    TEST(cur == Ramp::kMin); TEST(remaining == 0);
    MsgPlayable* ps = sil1->CreatePlayable();
    TEST(ps->IsSilence());
    ps->RemoveRef();
    sil2->RemoveRef();
    // muted audio becomes a silence playable without ramp (Msg.cpp:2252-2257)
    msg = factory.CreateMsgAudioPcm(Brn(data, 256), 2, 44100, 16, AudioDataEndian::Big, 0);
    msg->SetMuted();
    uint32_t rem2 = 1000000;
    TEST(msg->SetRamp(Ramp::kMax, rem2, Ramp::EDown, split) == Ramp::kMin); // already muted: ramp down is complete
    TEST(rem2 == 0);
    p1 = msg->CreatePlayable();
    TEST(p1->IsSilence()); TEST(!p1->Ramp().IsEnabled()); TEST(p1->Bytes() == 256);
    p1->RemoveRef();
    // MsgPlayable::Split: byte ratio ramp split; attenuation is dropped on the remainder (reference quirk)
    msg = factory.CreateMsgAudioPcm(Brn(data, 256), 2, 44100, 16, AudioDataEndian::Big, 0);
    msg->SetAttenuation(128);
    uint32_t rem3 = msg->Jiffies();
    msg->SetRamp(Ramp::kMax, rem3, Ramp::EDown, split);
    p1 = msg->CreatePlayable();
    p2 = p1->Split(64);
    TEST(p1->Bytes() == 64); TEST(p2->Bytes() == 192);
    TEST(p1->Ramp().End() == Ramp::kMax - Ramp::kMax / 4); TEST(p2->Ramp().Start() == p1->Ramp().End()); TEST(p2->Ramp().End() == 0);
    TEST(p1->Attenuation() == 128); TEST(p2->Attenuation() == MsgAudioPcm::kUnityAttenuation);
    TEST(p1->Split(64) == nullptr);
    TEST_THROWS(p1->Split(0), AssertionFailed);
    TEST_THROWS(p1->Split(65), AssertionFailed);
    // Read() without a GPU-backed reader attached must fail loudly, never fall back
    ProcessorPcmBuf proc;
    TEST_THROWS(p1->Read(proc), AssertionFailed);
    p1->RemoveRef();
    p2->RemoveRef();
    // MedianRampMultiplier (VolumeRamper.cpp:111-122)
    msg = factory.CreateMsgAudioPcm(Brn(data, 256), 2, 44100, 16, AudioDataEndian::Big, 0);
    TEST(msg->MedianRampMultiplier() == 0x8000);
    uint32_t rem4 = msg->Jiffies();
    msg->SetRamp(Ramp::kMax, rem4, Ramp::EDown, split);
    const uint32_t med = msg->MedianRampMultiplier();
    TEST(med > 0 && med < 0x7fff); TEST(!msg->Ramp().IsEnabled());
    msg->SetMuted();
    TEST(msg->MedianRampMultiplier() == 0);
    msg->RemoveRef();
}

typedef int64_t (*OracleProcessFn)(const ohp_chunk_desc*, size_t, const uint8_t*, uint64_t, uint8_t*, uint64_t);

// sink that remembers every delivery (per-stream concatenation, as a driver would send it)
class Collect : public IPcmProcessor
{
public:
    std::vector<uint8_t> all;
    unsigned begins = 0, ends = 0, silences = 0;
    void BeginBlock() override { begins++; }
    void ProcessFragment(const Brx& d, uint32_t ch, uint32_t b) override { OHP_ASSERT(d.Bytes() % (ch * b) == 0); all.insert(all.end(), d.Ptr(), d.Ptr() + d.Bytes()); }
    void ProcessSilence(const Brx& d, uint32_t ch, uint32_t b) override { OHP_ASSERT(d.Bytes() % (ch * b) == 0); silences++; all.insert(all.end(), d.Ptr(), d.Ptr() + d.Bytes()); }
    void EndBlock() override { ends++; }
    void Flush() override {}
};

static void SuiteGpuRead(const char* aOraclePath)
{
    void* so = dlopen(aOraclePath, RTLD_NOW);
    if (!so) { std::printf("cannot load oracle %s: %s\n", aOraclePath, dlerror()); gFailures++; return; }
    OracleProcessFn oracle = (OracleProcessFn)dlsym(so, "ohpo_process_chunks");
    BatchPcmReader reader(0, 1u << 20, 1u << 20);
    MsgFactory factory(&reader, &reader);
    // three streams of different formats, the way three pipelines would feed one batch
    struct Fmt { uint32_t ch, rate, bits; AudioDataEndian endian; } fmts[3] = {
        {2, 44100, 16, AudioDataEndian::Big}, {2, 192000, 24, AudioDataEndian::Big}, {6, 48000, 32, AudioDataEndian::Little}};
    Collect sinks[3];
    std::vector<ohp_chunk_desc> descs;
    std::vector<uint8_t> pcmCopy(1u << 20, 0);
    std::vector<std::vector<uint8_t>> expect(3);
    uint32_t seed = 12345;
    for (int round = 0; round < 2; round++) {
        uint64_t outAt = 0;
        descs.clear();
        std::vector<int> owner;
        for (int s = 0; s < 3; s++) {
            const Fmt& f = fmts[s];
            const uint32_t frameBytes = f.ch * f.bits / 8;
            uint32_t current = Ramp::kMax;
            uint32_t remaining = 3 * 200 * Jiffies::PerSample(f.rate);
            for (int m = 0; m < 4; m++) {
                const uint32_t frames = 200;
                uint8_t* dst = reader.Reserve(frames * frameBytes);   // "decode" straight into pinned memory
                for (uint32_t i = 0; i < frames * frameBytes; i++) { seed = seed * 1664525u + 1013904223u; dst[i] = (uint8_t)(seed >> 24); }
                MsgAudioPcm* msg = factory.CreateMsgAudioPcm(Brn(dst, frames * frameBytes), f.ch, f.rate, f.bits, f.endian, 0);
                MsgAudio* split = nullptr;
                if (m < 3) current = msg->SetRamp(current, remaining, Ramp::EDown, split);   // Ramper-style: 3 messages ramp down
                else msg->SetMuted();                                                         // Muter-style: then muted
                TEST(split == nullptr);
                MsgPlayable* playable = msg->CreatePlayable();
                MsgPlayable* tail = (m == 1) ? playable->Split(64 * frameBytes) : nullptr;   // driver-style block split
                for (MsgPlayable* p : {playable, tail}) {
                    if (p == nullptr) continue;
                    ohp_chunk_desc d = p->Descriptor((outAt + 15u) & ~15ull, OHP_OUT_PACKED_BE);
                    outAt = d.dst_off + d.bytes;
                    descs.push_back(d);
                    owner.push_back(s);
                    reader.Add(p, sinks[s]);
                }
            }
            TEST(current == Ramp::kMin);
        }
        // what the oracle says those playables read as (same arena bytes)
        const uint8_t* arena = reader.Reserve(0) - 0; // current end; arena base is end - used: recover base via Stage()
        (void)arena;
        std::vector<uint8_t> want((size_t)outAt + 64, 0);
        // the input arena is private to the reader: re-stage() of a pointer inside it returns its offset, so base = p - off
        uint8_t* probe = reader.Reserve(16);
        const uint64_t off = reader.Stage(Brn(probe, 16));
        const uint8_t* base = probe - off;
        const int64_t orc = oracle(descs.data(), descs.size(), base, off + 16, want.data(), want.size());
        TEST(orc == 0);
        for (size_t i = 0; i < descs.size(); i++) {
            expect[owner[i]].insert(expect[owner[i]].end(), want.begin() + descs[i].dst_off, want.begin() + descs[i].dst_off + descs[i].bytes);
        }
        reader.Flush();
    }
    for (int s = 0; s < 3; s++) {
        TEST(sinks[s].begins == sinks[s].ends);
        TEST(sinks[s].begins == 2 * 5);
        TEST(sinks[s].silences == 2);
        TEST(sinks[s].all.size() == expect[s].size());
        TEST(sinks[s].all == expect[s]);
    }
    // arena lifetime: audio staged before a Flush() and read after it is still where it was staged; the arena is only
    // recycled once nothing refers to it
    {
        TEST(reader.Pins() == 0);
        uint8_t a[96], b[96], c[96];
        std::memset(a, 0x11, sizeof a); std::memset(b, 0x22, sizeof b); std::memset(c, 0x33, sizeof c);
        MsgAudioPcm* keep = factory.CreateMsgAudioPcm(Brn(a, sizeof a), 2, 48000, 24, AudioDataEndian::Big, 0); // staged, not added yet
        ProcessorPcmBuf pa, pb, pc;
        reader.Add(factory.CreateMsgAudioPcm(Brn(b, sizeof b), 2, 48000, 24, AudioDataEndian::Big, 0)->CreatePlayable(), pb);
        reader.Flush();
        TEST(reader.Pins() == 1);
        TEST(pb.Buf() == std::vector<uint8_t>(b, b + sizeof b));
        reader.Add(factory.CreateMsgAudioPcm(Brn(c, sizeof c), 2, 48000, 24, AudioDataEndian::Big, 0)->CreatePlayable(), pc); // must not land on keep's bytes
        MsgPlayable* kp = keep->CreatePlayable();
        MsgPlayable* kp2 = kp->Split(48);
        reader.Add(kp, pa);
        reader.Flush();
        TEST(pc.Buf() == std::vector<uint8_t>(c, c + sizeof c));
        TEST(pa.Buf() == std::vector<uint8_t>(a, a + 48));
        TEST(reader.Pins() == 1);             // the split remainder is still out
        TEST_THROWS(reader.ResetArena(), AssertionFailed);
        reader.Add(kp2, pa);
        reader.Flush();
        TEST(pa.Buf() == std::vector<uint8_t>(a + 48, a + 96));
        TEST(reader.Pins() == 0);
        reader.ResetArena();
    }
    // drop-in synchronous MsgPlayable::Read
    {
        uint8_t pcm[240];
        for (unsigned i = 0; i < sizeof pcm; i++) pcm[i] = 0x7f;
        MsgAudioPcm* msg = factory.CreateMsgAudioPcm(Brn(pcm, sizeof pcm), 2, 48000, 24, AudioDataEndian::Big, 0);
        uint32_t rem = msg->Jiffies();
        MsgAudio* split = nullptr;
        msg->SetRamp(Ramp::kMax, rem, Ramp::EDown, split);
        MsgPlayable* p = msg->CreatePlayable();
        ProcessorPcmBuf proc;
        p->Read(proc);
        TEST(proc.Buf().size() == sizeof pcm);
        TEST(proc.Buf()[0] == 0x7f && proc.Buf()[1] == 0x7e && proc.Buf()[2] == 0x00);  // TestMuter.cpp:332
        TEST(proc.Buf()[sizeof pcm - 3] == 0 && proc.Buf()[sizeof pcm - 2] == 0);        // ramp reached kMin
        p->RemoveRef();
    }
    dlclose(so);
}

int main(int argc, char** argv)
{
    bool gpu = false;
    std::string oracle = "oracle/libohp_oracle.so";
    for (int i = 1; i < argc; i++) {
        if (std::strcmp(argv[i], "--gpu") == 0) gpu = true;
        else if (std::strcmp(argv[i], "--oracle") == 0 && i + 1 < argc) oracle = argv[++i];
    }
    try {
        SuiteRampAlgebra();
        SuiteMsgModel();
        if (gpu) SuiteGpuRead(oracle.c_str());
    }
    catch (const std::exception& e) {
        std::printf("FAILED: unexpected exception: %s\n", e.what());
        gFailures++;
    }
    std::printf("%s: %d checks, %d failures%s\n", gFailures ? "FAIL" : "PASS", gChecks, gFailures, gpu ? " (with GPU suite)" : "");
    return gFailures ? 1 : 0;
}
