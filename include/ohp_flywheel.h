/*
 * ohp_flywheel.h -- C ABI of the B200-native batch implementation of ohPipeline's flywheel ramp generator:
 * what StarvationRamper plays when its reservoir runs dry.  Citations are relative to the ohPipeline
 * source root.
 *
 *   reference                                                              this ABI
 *   -----------------------------------------------------------------     ------------------------------------
 *   FlywheelInput::Prepare             (Media/Pipeline/StarvationRamper.cpp:90-111)   ohp_chunk_desc with OHP_OUT_PLANAR32_BE
 *                                                                           (ohp_b200.h; fills the training block)
 *   FlywheelRamperManager::Ramp        (Media/FlywheelRamper.cpp:46-68)     ohp_flywheel_device, one ohp_flywheel_job
 *     InitChannels / FlywheelRamper::Initialise    (:70-83, 178-231)          per starving stream
 *     FlywheelRamper::BurgsMethod      (:253-328)   integer Burg LPC, degree 3
 *     CorrectBurgCoeffs / CoeffOverflow (:347-388)
 *     PrepareFeedbackCoeffs            (:233-240)
 *     FeedbackModel::NextSample        (:447-487)   the all-pole extrapolation
 *     RenderChannels                   (:85-135)    sample-hold decimation, 1 ms blocks
 *   RampGenerator::ProcessFragment     (StarvationRamper.cpp:281-326)       fused: 32-bit -> packed bit_depth BE
 *   RampGenerator::Start / EndBlock    (StarvationRamper.cpp:235-247, 351-364)  ohp_flywheel_ramp_chunks (libohp_host.so):
 *                                                                           one ramped ohp_chunk_desc per 1 ms block,
 *                                                                           consumed by ohp_process_device
 *
 * Everything is integer (the reference's double helpers, FlywheelRamper.cpp:390-443, are unused by the path), so
 * results are bit-exact: tests/test_flywheel*.py compare with the reference's own FlywheelRamper.cpp /
 * StarvationRamper.cpp linked into oracle/_ref and with the known answers of Media/Tests/TestFlywheelRamper.cpp.
 */
#ifndef OHP_FLYWHEEL_H
#define OHP_FLYWHEEL_H

#include "ohp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define OHP_FLYWHEEL_DEGREE 3u                    /* kDegree, FlywheelRamper.cpp:14                         */
#define OHP_FLYWHEEL_TRAINING_JIFFIES 56448u      /* StarvationRamper::kTrainingJiffies = 1 ms (:374)       */
#define OHP_FLYWHEEL_RAMP_JIFFIES (20u * 56448u)  /* StarvationRamper::kRampDownJiffies = 20 ms (:375)      */
#define OHP_FLYWHEEL_BLOCK_JIFFIES 56448u         /* FlywheelRamperManager::kMaxOutputJiffiesBlockSize (:22) */
#define OHP_FLYWHEEL_MAX_CHANNELS 8u              /* RampGenerator::kMaxChannels (StarvationRamper.h:65)    */
#define OHP_FLYWHEEL_MAX_TRAIN_FRAMES 384u        /* 1 ms at 384 kHz                                        */
#define OHP_FLYWHEEL_MAX_BLOCK_BYTES 6144u        /* RampGenerator's iFlywheelAudio: 192 samples x 4 B x 8 ch (StarvationRamper.cpp:215-219) */
#define OHP_FLYWHEEL_MAX_INPUT_BYTES 7680u        /* FlywheelInput's buffer: 192 samples x 4 B x 10 ch (StarvationRamper.cpp:76-83)          */

/*
 * One starving stream.  32 bytes, 16-byte aligned.
 * Training block (input arena): planar, 4 bytes per subsample, big-endian, left-justified -- FlywheelInput's layout,
 * i.e. what an OHP_OUT_PLANAR32_BE chunk with aux = train_frames writes; channel c starts at src_off + c*train_frames*4.
 * Generated audio (output arena): out_frames interleaved frames, packed big-endian at bit_depth (32-bit: three bytes
 * and a zero, StarvationRamper.cpp:313-321) = the concatenation of the MsgAudioPcm payloads RampGenerator enqueues.
 */
typedef struct ohp_flywheel_job {
    uint64_t src_off;
    uint64_t dst_off;
    uint32_t sample_rate;   /* any rate Jiffies::PerSample accepts                                              */
    uint32_t out_frames;    /* Jiffies::ToSamples(ramp jiffies, rate): 20 ms for StarvationRamper                */
    uint16_t train_frames;  /* must equal Jiffies::ToSamples(OHP_FLYWHEEL_TRAINING_JIFFIES, rate)                */
    uint8_t  channels;      /* 1..OHP_FLYWHEEL_MAX_CHANNELS                                                      */
    uint8_t  bit_depth;     /* 8, 16, 24, 32: depth of the generated audio (RampGenerator::iBitDepth)            */
    uint32_t reserved;
} ohp_flywheel_job;

/* Bytes job j writes at dst_off. */
uint32_t ohp_flywheel_out_bytes(const ohp_flywheel_job* job);
/* Check `n` jobs against the reference's ASSERTs / buffer sizes and the arena sizes. */
int ohp_flywheel_validate(const ohp_flywheel_job* jobs, size_t n, uint64_t in_bytes, uint64_t out_bytes, size_t* bad_index);
/*
 * Replaces FlywheelRamperManager::Ramp + RampGenerator::ProcessFragment for `n` streams at once.  DEVICE pointers;
 * asynchronous on `stream` (NULL = the context's own).  Jobs are checked on the device; a violation is reported by the
 * next ohp_sync as OHP_E_INVALID_DESC / OHP_E_OUT_OF_RANGE and that job is skipped.
 */
int ohp_flywheel_device(ohp_context* ctx, const ohp_flywheel_job* d_jobs, size_t n,
                        const uint8_t* d_in, uint64_t in_bytes, uint8_t* d_out, uint64_t out_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OHP_FLYWHEEL_H */
