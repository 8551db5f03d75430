// Sanitizer harness (AddressSanitizer + UndefinedBehaviorSanitizer): built and fed by tests/test_sanitizers.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "ohp_container.h"
int main(int argc, char** argv) {
    FILE* f = argc > 1 ? fopen(argv[1], "rb") : nullptr; if (!f) return 2;
    unsigned n = 0, ok = 0; uint32_t len;
    while (fread(&len, 4, 1, f) == 1) {
        uint8_t* buf = (uint8_t*)malloc(len ? len : 1);   // exact size: ASAN sees any over-read
        if (len && fread(buf, 1, len, f) != len) return 3;
        ohp_container_info info;
        int rc = ohp_container_parse(len ? buf : nullptr, len, 32, &info);
        if (rc == OHP_CONTAINER_OK) {
            ohp_stream_spec sp;
            if (ohp_container_stream_spec(&info, len, 0, 0, &sp) == OHP_CONTAINER_OK) {
                ok++;
                uint64_t fb = (uint64_t)sp.channels * (sp.bit_depth / 8);
                if (sp.src_base + sp.total_frames * fb > len) { printf("spec reaches past the file\n"); return 1; }
                std::vector<uint32_t> frames(64);
                ohp_codec_message_frames(&sp, frames.data(), frames.size());
            }
        }
        free(buf); n++;
    }
    printf("%u buffers parsed, %u usable, no sanitizer report\n", n, ok);
    return 0;
}
