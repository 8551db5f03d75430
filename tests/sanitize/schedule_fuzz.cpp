// Sanitizer harness (AddressSanitizer + UndefinedBehaviorSanitizer): built and fed by tests/test_sanitizers.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "ohp_schedule.h"
int main(int argc, char** argv) {
    FILE* f = argc > 1 ? fopen(argv[1], "rb") : nullptr; if (!f) return 2;
    uint32_t hdr[2]; unsigned n = 0, refused = 0; unsigned long chunks = 0;
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
            } else refused++;
            if (a) ohp_schedule_free(a);
            if (b) ohp_schedule_free(b);
        }
        free(s); free(e); n++;
    }
    printf("%u workloads, %lu chunks, %u streams refused by both, no sanitizer report\n", n, chunks, refused);
    return 0;
}
