// How much of a workload the schedule walk's bulk step covers: messages taken 32 at a time vs one at a time.
// Built and fed by tests/test_sanitizers.py (a CPU-side guard for a GPU-side performance property: a stream that
// falls off the bulk step costs ~1 us per message on the device instead of ~0.3).
#include <cstdio>
#include <cstdlib>
#include <vector>
static unsigned long g_bulk_calls = 0, g_bulk_msgs = 0, g_general = 0;
#define OHP_WALK_STATS 1
#include "../../ohpipeline_b200/host/schedule_walk.h"

int main(int argc, char** argv)
{
    FILE* f = argc > 1 ? fopen(argv[1], "rb") : nullptr;
    if (!f) return 2;
    uint32_t hdr[2];
    while (fread(hdr, 4, 2, f) == 2) {
        std::vector<ohp_stream_spec> s(hdr[0]);
        std::vector<ohp_ramp_event> e(hdr[1] ? hdr[1] : 1);
        if (fread(s.data(), sizeof s[0], hdr[0], f) != hdr[0] || fread(e.data(), sizeof e[0], hdr[1], f) != hdr[1]) return 3;
        g_bulk_calls = g_bulk_msgs = g_general = 0;
        unsigned long chunks = 0;
        for (uint32_t k = 0; k < hdr[0]; k++) {
            uint64_t n = 0, bytes = 0;
            if (ohp::sched::run_stream<false>(s[k], e.data(), hdr[1], nullptr, nullptr, n, bytes) == ohp::sched::kOk) chunks += n;
        }
        std::printf("%lu %lu %lu %lu\n", chunks, g_bulk_calls, g_bulk_msgs, g_general);
    }
    return 0;
}
