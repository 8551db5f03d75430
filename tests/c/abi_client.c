/*
 * abi_client.c -- a plain C99 program on the drop-in boundary: it includes nothing but the headers under include/, links libohp_b200.so and
 * libohp_host.so, and does what a cgo / JNI / N-API stub would do through the same symbols.  Built and run
 * by tests/test_abi.py: the control plane of one starved stream (schedule, starvation record, the three flywheel launches
 * as data, validated the way the device calls validate them), and -- there being no GPU in that run -- every way into the
 * compute path refusing loudly.  Prints "key value" lines; exit code 0 when everything behaved as the headers say.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ohp_b200.h"
#include "ohp_flywheel.h"
#include "ohp_multi.h"
#include "ohp_schedule.h"
#include "ohp_schedule_device.h"

#define CHECK(cond) do { if (!(cond)) { printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)

int main(void)
{
    /* one stream: 48 kHz stereo 24-bit, 100 ms, its StarvationRamper (stage 1) starved 30.4 ms in, a Ramper (stage 0) ahead */
    enum { kRate = 48000, kFrames = 4800, kFrameBytes = 6 };
    ohp_stream_spec spec;
    ohp_ramp_event events[2];
    ohp_schedule* sch = NULL;
    const ohp_starvation* sv;
    ohp_chunk_desc prep[OHP_FLYWHEEL_MAX_PREP], blocks[24];
    ohp_flywheel_job job;
    size_t n_prep = 0, n_blocks = 0, first = 0, count = 0, bad = 0;
    uint64_t total_out;
    int devices, rc;
    const uint32_t jps = ohp_jiffies_per_sample(kRate);

    CHECK(ohp_abi_version() == OHP_ABI_VERSION);
    CHECK(jps == 1176u && ohp_jiffies_per_sample(12345) == 0u);
    memset(&spec, 0, sizeof spec);
    spec.sample_rate = kRate; spec.bit_depth = 24; spec.channels = 2; spec.chunk_frames = 240; spec.out_fmt = OHP_OUT_PACKED_BE;
    spec.total_frames = kFrames; spec.num_events = 2;
    memset(events, 0, sizeof events);
    events[0].at_jiffies = 0; events[0].stage = 0; events[0].op = OHP_EV_RAMPER_STREAM; events[0].arg = 50u * OHP_JIFFIES_PER_MS;
    events[1].at_jiffies = 30u * OHP_JIFFIES_PER_MS + 23456u; events[1].stage = 1; events[1].op = OHP_EV_STARVATION;
    events[1].arg = 50u * OHP_JIFFIES_PER_MS;

    /* the control plane on the host: playables, and what the starved element was doing */
    CHECK(ohp_schedule_build(&spec, 1, events, 2, 1, &sch) == OHP_OK);
    CHECK(ohp_schedule_num_chunks(sch) > 20);
    total_out = ohp_schedule_stream_out_bytes(sch)[0];
    CHECK(total_out == (uint64_t)kFrames * kFrameBytes);
    CHECK(ohp_schedule_num_starvations(sch) == 1);
    sv = ohp_schedule_starvations(sch);
    CHECK(sv[0].plays == 1 && sv[0].pcm_jiffies == events[1].at_jiffies && sv[0].event == 1);
    CHECK(sv[0].ramp == 16384u); /* the element has not ramped anything itself yet: Ramp::kMax, whatever the Ramper ahead of it put on the messages */
    CHECK(ohp_flywheel_plan(&spec, &sv[0], 0, 0, 0, prep, &n_prep, &job, blocks, 24, &n_blocks) == OHP_OK);
    CHECK(n_prep == 1 && n_blocks == 20 && job.train_frames == 48 && job.out_frames == 960);
    CHECK(ohp_validate(ohp_schedule_chunks(sch), ohp_schedule_num_chunks(sch), total_out, total_out, &bad) == OHP_OK);
    CHECK(ohp_validate(prep, n_prep, total_out, 48u * 4u * 2u, &bad) == OHP_OK);
    CHECK(ohp_flywheel_validate(&job, 1, 48u * 4u * 2u, 960u * kFrameBytes, &bad) == OHP_OK);
    CHECK(ohp_validate(blocks, n_blocks, 960u * kFrameBytes, 960u * kFrameBytes, &bad) == OHP_OK);
    printf("chunks %lu\n", (unsigned long)ohp_schedule_num_chunks(sch));
    printf("starved_at_ramp %u\n", (unsigned)sv[0].ramp);

    ohp_multi_shard(65536, 8, 3, &first, &count);
    CHECK(first == 24576 && count == 8192);

    devices = ohp_device_count();
    printf("devices %d\n", devices);
    if (devices == 0) {
        /* no CPU fallback anywhere: every way in says so */
        ohp_context* ctx = NULL;
        ohp_multi* m = NULL;
        int dev0 = 0;
        rc = ohp_create(0, &ctx);
        CHECK(rc == OHP_E_NO_DEVICE && ctx == NULL && strlen(ohp_last_error(NULL)) > 0);
        rc = ohp_multi_create(&dev0, 1, &m);
        CHECK((rc == OHP_E_NO_DEVICE || rc == OHP_E_CUDA) && m == NULL);
        printf("refused %s\n", ohp_last_error(NULL));
    }
    else {
        /* with a device: a context comes and goes (the compute calls are exercised through the same symbols by tests -m gpu) */
        ohp_context* ctx = NULL;
        CHECK(ohp_create(0, &ctx) == OHP_OK && ctx != NULL);
        CHECK(ohp_destroy(ctx) == OHP_OK);
    }
    ohp_schedule_free(sch);
    printf("ok\n");
    return 0;
}
