// Sanitizer harness (AddressSanitizer + UndefinedBehaviorSanitizer): built and fed by tests/test_sanitizers.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "ohp_schedule.h"
// Every starvation the model recorded through ohp_flywheel_plan: what it plans must lie inside the stream's PCM, the training
// block and the generated audio, and no two FlywheelInput descriptors may write the same slot.
static int plan_starvations(const ohp_stream_spec& one, const ohp_schedule* a, unsigned long& planned, unsigned long& unplanned)
{
    const ohp_starvation* sv = ohp_schedule_starvations(a);
    const uint64_t fb = (uint64_t)one.channels * (one.bit_depth / 8u);
    for (size_t k = 0; k < ohp_schedule_num_starvations(a); k++) {
      // twice: the PCM-only plan, then the plan from the recent audio piece by piece (exact-size copy of the pieces)
      for (int from_recent = 0; from_recent < 2; from_recent++) {
        const size_t prep_cap = from_recent ? 48 : OHP_FLYWHEEL_MAX_PREP;
        ohp_chunk_desc* prep = (ohp_chunk_desc*)malloc(sizeof(ohp_chunk_desc) * prep_cap);
        ohp_chunk_desc* blocks = (ohp_chunk_desc*)malloc(sizeof(ohp_chunk_desc) * 24);
        ohp_flywheel_job job;
        size_t np = 0, nb = 0;
        int rc;
        if (from_recent) {
            const uint64_t* rb = ohp_schedule_recent_begin(a);
            const size_t nr = (size_t)(rb[k + 1] - rb[k]);
            ohp_recent_audio* pieces = (ohp_recent_audio*)malloc(sizeof(ohp_recent_audio) * (nr ? nr : 1));
            if (nr) memcpy(pieces, ohp_schedule_recent_audio(a) + rb[k], nr * sizeof(ohp_recent_audio));
            rc = ohp_flywheel_plan_recent(&one, &sv[k], pieces, nr, 0, 0, 0, prep, prep_cap, &np, &job, blocks, 24, &nb);
            free(pieces);
        }
        else rc = ohp_flywheel_plan(&one, &sv[k], 0, 0, 0, prep, &np, &job, blocks, 24, &nb);
        if (rc == OHP_OK) {
            planned++;
            const uint64_t slots = (uint64_t)job.train_frames * one.channels;
            std::vector<char> written(slots, 0);
            for (size_t i = 0; i < np; i++) {
                const ohp_chunk_desc& d = prep[i];
                const uint64_t dfb = (uint64_t)d.channels * (d.bit_depth / 8u);
                const bool silent = (d.flags & OHP_F_SILENCE) != 0;
                if ((!silent && (d.src_off < one.src_base || d.src_off + d.bytes > one.src_base + one.total_frames * fb)) || d.bytes % dfb || d.dst_off % 4) { printf("plan reads outside the stream\n"); return 1; }
                const uint64_t frames = d.bytes / dfb;
                for (uint32_t c = 0; c < d.channels; c++) {
                    for (uint64_t i2 = 0; i2 < frames; i2++) {
                        const uint64_t slot = d.dst_off / 4 + (uint64_t)c * d.aux + i2;
                        if (slot >= slots || written[slot]) { printf("plan writes a training slot twice or outside the block\n"); return 1; }
                        written[slot] = 1;
                    }
                }
            }
            for (uint64_t i = 0; i < slots; i++) if (!written[i]) { printf("plan leaves a training slot unwritten\n"); return 1; }
            uint64_t out = 0;
            for (size_t i = 0; i < nb; i++) out += blocks[i].bytes;
            if (out != (uint64_t)job.out_frames * fb) { printf("ramped blocks do not cover the generated audio\n"); return 1; }
        }
        else if (rc == OHP_E_INVALID_ARG || rc == OHP_E_INVALID_DESC) unplanned++;
        else { printf("plan status %d\n", rc); return 1; }
        free(prep); free(blocks);
      }
    }
    return 0;
}
int main(int argc, char** argv) {
    FILE* f = argc > 1 ? fopen(argv[1], "rb") : nullptr; if (!f) return 2;
    uint32_t hdr[2]; unsigned n = 0, refused = 0; unsigned long chunks = 0, planned = 0, unplanned = 0;
    while (fread(hdr, 4, 2, f) == 2) {
        // exact-size heap buffers so that ASAN sees any over-read of the inputs
        ohp_stream_spec* s = (ohp_stream_spec*)malloc(sizeof(ohp_stream_spec) * (hdr[0] ? hdr[0] : 1));
        ohp_ramp_event* e = (ohp_ramp_event*)malloc(sizeof(ohp_ramp_event) * (hdr[1] ? hdr[1] : 1));
        if (fread(s, sizeof *s, hdr[0], f) != hdr[0] || fread(e, sizeof *e, hdr[1], f) != hdr[1]) return 3;
        // stream by stream: one the reference ASSERTs on must not hide its neighbours
        for (uint32_t k = 0; k < hdr[0]; k++) {
            ohp_stream_spec one = s[k];
            const ohp_ramp_event* ev = e + one.first_event;
            one.first_event = 0;
            ohp_schedule *a = nullptr, *b = nullptr;
            int ra = ohp_schedule_build(&one, 1, ev, one.num_events, 1, &a);
            int rb = ohp_schedule_build_walk(&one, 1, ev, one.num_events, 1, &b);
            if ((ra != 0) != (rb != 0)) { printf("status differs %d %d\n", ra, rb); return 1; }
            if (ra == 0) {
                size_t na = ohp_schedule_num_chunks(a), nb = ohp_schedule_num_chunks(b);
                if (na != nb || memcmp(ohp_schedule_chunks(a), ohp_schedule_chunks(b), na * sizeof(ohp_chunk_desc)) != 0) { printf("schedules differ\n"); return 1; }
                chunks += na;
                if (plan_starvations(one, a, planned, unplanned) != 0) return 1;
            } else refused++;
            if (a) ohp_schedule_free(a);
            if (b) ohp_schedule_free(b);
        }
        free(s); free(e); n++;
    }
    printf("%u workloads, %lu chunks, %u streams refused by both, %lu starvations planned, %lu not (nothing to play, silence in the block, shapes the flywheel does not take), no sanitizer report\n", n, chunks, refused, planned, unplanned);
    return 0;
}
