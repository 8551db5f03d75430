/*
 * ohp_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See ohp_oracle.h.
 *
 * Plain-C restatement of the reference's ramp + format-convert path.  Scalar, one thread, written
 * for clarity: it is the checker, never the thing shipped or measured as the product.
 * Types: the reference's TUint/TInt are 32-bit (ohNet Types.h); uint32_t/int32_t here.
 */
#include "ohp_oracle.h"

#include <stdlib.h>
#include <string.h>

#define KMAX 16384u /* Ramp::kMax, Msg.h:257 */
#define KMIN 0u     /* Ramp::kMin, Msg.h:258 */
#define CELL_MAX 9216u /* AudioData::kMaxBytes, Msg.h:117 */

const uint32_t ohpo_ramp_array[512] = {
#include "ramp_table.inc"
};

/* ------------------------------------------------------------------------------------------- */
/* Jiffies (Msg.cpp:424-489, Msg.h:193)                                                         */

uint32_t ohpo_jiffies_per_sample(uint32_t rate)
{
    /* Jiffies::PerSample, Msg.cpp:424-470: a switch over the supported rates, each constant
     * being kPerSecond / rate (Msg.h:193-230).  DSD rates are not PCM and are excluded here. */
    switch (rate) {
    case 7350: case 8000: case 11025: case 12000: case 14700: case 16000:
    case 22050: case 24000: case 29400: case 32000: case 44100: case 48000:
    case 88200: case 96000: case 176400: case 192000: case 352800: case 384000:
        return OHP_JIFFIES_PER_SECOND / rate;
    default:
        return 0;
    }
}

uint32_t ohpo_jiffies_to_bytes(uint32_t* jiffies, uint32_t jps, uint32_t channels, uint32_t bits)
{
    /* Jiffies::ToBytesSampleBlock with aSamplesPerBlock == 1, Msg.cpp:481-489 */
    *jiffies -= *jiffies % jps;
    const uint32_t num_samples = *jiffies / jps;
    const uint32_t num_subsamples = num_samples * channels;
    return ((num_subsamples * bits) + 7) / 8;
}

static void round_down_non_zero_sample_block(uint32_t* jiffies, uint32_t block)
{
    /* Jiffies::RoundDownNonZeroSampleBlock, Msg.cpp:504-514 */
    uint32_t j = *jiffies;
    j -= j % block;
    if (j == 0) {
        j = *jiffies;
        j += block - 1;
        j -= j % block;
    }
    *jiffies = j;
}

/* ------------------------------------------------------------------------------------------- */
/* Ramp (Msg.cpp:568-807)                                                                       */

static void ramp_reset(ohp_ramp* r)
{
    /* Ramp::Reset, Msg.cpp:582-588 */
    r->start = KMAX;
    r->end = KMAX;
    r->direction = OHP_DIR_NONE;
    r->enabled = 0;
}

static int ramp_do_validate(const ohp_ramp* r)
{
    /* Ramp::DoValidate, Msg.cpp:745-782 */
    if (r->start > KMAX) return 0;
    if (r->end > KMAX) return 0;
    switch (r->direction) {
    case OHP_DIR_NONE: return r->start == r->end;
    case OHP_DIR_UP:   return r->start < r->end;
    case OHP_DIR_DOWN: return r->start > r->end;
    case OHP_DIR_MUTE: return r->start == r->end && r->start == KMIN;
    default: return 0;
    }
}

static void ramp_select_lower(ohp_ramp* r, uint32_t req_start, uint32_t req_end)
{
    /* Ramp::SelectLowerRampPoints, Msg.cpp:721-734 */
    if (req_start < r->start) r->start = req_start;
    if (req_end < r->end) r->end = req_end;
    if (r->start == r->end) r->direction = OHP_DIR_NONE;
    else if (r->start > r->end) r->direction = OHP_DIR_DOWN;
    else r->direction = OHP_DIR_UP;
}

int ohpo_ramp_set(ohp_ramp* r, uint32_t a_start, uint32_t frag, uint32_t dur, uint32_t dir,
                  ohp_ramp* split, uint32_t* split_pos)
{
    /* Ramp::Set, Msg.cpp:590-712 */
    if (!(dur >= frag)) return -1;          /* :598 */
    if (dir == OHP_DIR_NONE) return -1;     /* :599 */
    if (dur == 0) return -1;                /* the reference would divide by zero */
    r->enabled = 1;
    ramp_reset(split);
    *split_pos = 0xffffffffu;
    const uint32_t ramp_remaining = (dir == OHP_DIR_DOWN ? a_start : KMAX - a_start);
    /* ceil so that rounding can never stop a ramp completing inside its duration (:604-605) */
    const uint32_t ramp_delta = (uint32_t)(((ramp_remaining * (uint64_t)frag) + dur - 1) / dur);
    uint32_t ramp_end;
    if (dir == OHP_DIR_DOWN) {
        if (ramp_delta > a_start) {
            if (!(ramp_delta - a_start <= frag - 1)) return -1; /* :611 */
            ramp_end = 0;
        } else {
            ramp_end = a_start - ramp_delta;
        }
    } else {
        if (a_start + ramp_delta > KMAX) {
            if (!(a_start + ramp_delta - KMAX <= frag - 1)) return -1; /* :620 */
            ramp_end = KMAX;
        } else {
            ramp_end = a_start + ramp_delta;
        }
    }
    if (r->direction == OHP_DIR_NONE) {
        /* no previous ramp (:627-632) */
        r->direction = dir;
        r->start = a_start;
        r->end = ramp_end;
    } else if (r->direction == dir) {
        ramp_select_lower(r, a_start, ramp_end); /* :633-636 */
    } else {
        /* opposite directions: intersect the two lines (:637-699) */
        int64_t y1, y2, y3, y4;
        if (r->start < a_start) {
            y1 = r->start; y2 = r->end; y3 = a_start; y4 = ramp_end;
        } else {
            y1 = a_start; y2 = ramp_end; y3 = r->start; y4 = r->end;
        }
        if ((y2 - y1) == (y4 - y3)) {
            ramp_select_lower(r, a_start, ramp_end); /* parallel */
        } else {
            const int64_t ix = ((int64_t)frag * (y3 - y1)) / ((y2 - y1) - (y4 - y3));
            const int64_t iy = (((y2 - y1) * (y3 - y1)) / ((y2 - y1) - (y4 - y3))) + y1;
            if (ix <= 0 || (uint32_t)ix >= frag) {
                ramp_select_lower(r, a_start, ramp_end);
            } else {
                *split_pos = (uint32_t)ix;
                split->start = (uint32_t)iy;
                split->end = (r->end < ramp_end ? r->end : ramp_end);
                split->direction = (split->start == split->end ? OHP_DIR_NONE : OHP_DIR_DOWN);
                split->enabled = 1;
                const uint32_t start = (r->start < a_start ? r->start : a_start);
                const uint32_t end = (uint32_t)iy;
                r->direction = (start == end ? OHP_DIR_NONE : OHP_DIR_UP);
                r->start = start;
                r->end = end;
            }
        }
    }
    if (!ramp_do_validate(r)) return -1; /* :701-708 ASSERTS */
    return split->enabled ? 1 : 0;
}

int ohpo_ramp_split(ohp_ramp* r, uint32_t new_size, uint32_t cur_size, ohp_ramp* remaining)
{
    /* Ramp::Split, Msg.cpp:784-807 */
    if (cur_size == 0) return -1;
    ramp_reset(remaining);
    remaining->end = r->end;
    remaining->direction = r->direction;
    remaining->enabled = 1;
    if (r->direction == OHP_DIR_UP) {
        const uint32_t ramp = (uint32_t)(((uint32_t)(r->end - r->start) * (uint64_t)new_size) / cur_size);
        r->end = r->start + ramp;
    } else {
        const uint32_t ramp = (uint32_t)(((uint32_t)(r->start - r->end) * (uint64_t)new_size) / cur_size);
        r->end = r->start - ramp;
    }
    if (r->start == r->end) r->direction = OHP_DIR_NONE;
    remaining->start = r->end; /* :802, the reference's FIXME: no one-step advance */
    if (!ramp_do_validate(r)) return -1;
    if (!ramp_do_validate(remaining)) return -1;
    return 0;
}

uint32_t ohpo_median_multiplier(const ohp_ramp* ramp)
{
    /* MsgAudio::MedianRampMultiplier, Msg.cpp:2063-2074 */
    if (!ramp->enabled) return 0x8000u;
    if (ramp->direction == OHP_DIR_MUTE) return 0;
    /* RampApplicator::MedianMultiplier, Msg.cpp:901-920 */
    uint32_t med;
    switch (ramp->direction) {
    case OHP_DIR_UP:   med = ramp->start + ((ramp->end - ramp->start) / 2); break;
    case OHP_DIR_DOWN: med = ramp->start - ((ramp->start - ramp->end) / 2); break;
    default:           med = ramp->start; break;
    }
    const uint32_t idx = (KMAX - KMIN - med + (1u << 4)) >> 5;
    if (idx >= 512) return 0; /* reference reads past the table here (undefined); documented deviation */
    return ohpo_ramp_array[idx];
}

/* ------------------------------------------------------------------------------------------- */
/* Input unpack (Msg.cpp:347-408)                                                               */

void ohpo_unpack_to_be(const uint8_t* src, uint8_t* dst, uint32_t bytes, uint32_t bits, int le)
{
    if (!le || bits == 8) {          /* ConstructPcm :352-354 */
        memcpy(dst, src, bytes);
        return;
    }
    if (bits == 16) {                /* CopyToBigEndian16 :380-387 */
        for (uint32_t i = 0; i < bytes; i += 2) { *dst++ = src[i + 1]; *dst++ = src[i]; }
    } else if (bits == 24) {         /* CopyToBigEndian24 :389-397 */
        for (uint32_t i = 0; i < bytes; i += 3) { *dst++ = src[i + 2]; *dst++ = src[i + 1]; *dst++ = src[i]; }
    } else {                         /* CopyToBigEndian32 :399-408 */
        for (uint32_t i = 0; i < bytes; i += 4) { *dst++ = src[i + 3]; *dst++ = src[i + 2]; *dst++ = src[i + 1]; *dst++ = src[i]; }
    }
}

/* ------------------------------------------------------------------------------------------- */
/* MsgPlayable::Read (Msg.cpp:2646-2653, 2736-2786, 820-899, 2874-2893)                          */

static void apply_attenuation(uint8_t* p, uint32_t bytes, uint32_t attenuation)
{
    /* MsgPlayablePcm::ApplyAttenuation, Msg.cpp:2736-2751.  iAttenuation is a TUint, so
     * ((TInt)sample) * iAttenuation / 256 is evaluated in unsigned 32-bit arithmetic. */
    const uint32_t samples = bytes / 2;
    for (uint32_t i = 0; i < samples; i++) {
        int16_t sample = (int16_t)(p[0] << 8);
        sample = (int16_t)(sample + p[1]);
        const int16_t att = (int16_t)(((uint32_t)(int32_t)sample) * attenuation / OHP_UNITY_ATTENUATION);
        *p++ = (uint8_t)(att >> 8);
        *p++ = (uint8_t)att;
    }
}

static void ramp_apply(const uint8_t* src, uint8_t* dst, uint32_t bytes, uint32_t bits, uint32_t ch,
                       uint32_t start, uint32_t end)
{
    /* RampApplicator::Start (Msg.cpp:820-830) + GetNextSample (:832-899), called once per frame by
     * MsgPlayablePcm::ReadBlock (:2761-2780).  Fragmenting into <=256-byte ProcessFragment calls only
     * changes call granularity, not the concatenated bytes. */
    const uint32_t byte_depth = bits / 8;
    const int32_t num_samples = (int32_t)(bytes / (byte_depth * ch));
    const int32_t total_ramp = (int32_t)(start - end);
    for (int32_t loop = 0; loop < num_samples; loop++) {
        const uint16_t ramp = (num_samples == 1)
            ? (uint16_t)start
            : (uint16_t)(start - (uint32_t)((loop * total_ramp) / (num_samples - 1)));
        uint32_t ramp_index = (KMAX - KMIN - ramp + (1u << 4)) >> 5;
        if (ramp_index > 511) ramp_index = 511;
        for (uint32_t c = 0; c < ch; c++) {
            int16_t s16 = 0;
            switch (bits) {
            case 8:  s16 = (int16_t)(*src++ << 8); break;
            case 16: s16 = (int16_t)(*src++ << 8); s16 = (int16_t)(s16 + *src++); break;
            case 24: s16 = (int16_t)(*src++ << 8); s16 = (int16_t)(s16 + *src++); src++; break;
            default: s16 = (int16_t)(*src++ << 8); s16 = (int16_t)(s16 + *src++); src++; src++; break;
            }
            const uint16_t mult = (uint16_t)ohpo_ramp_array[ramp_index];
            const int32_t ramped = ((int32_t)s16 * (int32_t)mult) >> 15; /* arithmetic shift, as gcc does for TInt */
            switch (bits) {
            case 8:  *dst++ = (uint8_t)(ramped >> 8); break;
            case 16: *dst++ = (uint8_t)(ramped >> 8); *dst++ = (uint8_t)ramped; break;
            case 24: *dst++ = (uint8_t)(ramped >> 8); *dst++ = (uint8_t)ramped; *dst++ = 0; break;
            default:
                *dst++ = (uint8_t)(ramped >> 8); *dst++ = (uint8_t)ramped; *dst++ = 0;
                *dst++ = (ch == 6) ? (uint8_t)(c << 4) : 0; /* :886-891 channel tag on 6-channel 32-bit */
                break;
            }
        }
    }
}

static void silence_fill(uint8_t* dst, uint32_t bytes, uint32_t bits, uint32_t ch)
{
    /* MsgPlayableSilence::ReadBlock, Msg.cpp:2874-2893: zeros in blocks of maxBytes; for 6 channels the
     * block source is silence6ch whose first 32 bytes carry 00 00 00 c0 with c0 = 0x00,0x10..0x70. */
    const uint32_t subsample_bytes = bits / 8;
    const uint32_t max_bytes = CELL_MAX - (CELL_MAX % (ch * subsample_bytes));
    uint32_t remaining = bytes;
    memset(dst, 0, bytes);
    while (remaining > 0) {
        const uint32_t n = remaining > max_bytes ? max_bytes : remaining;
        if (ch == 6) {
            for (uint32_t i = 0; i < 8 && (4 * i + 3) < n; i++) dst[4 * i + 3] = (uint8_t)(i << 4);
        }
        dst += n;
        remaining -= n;
    }
}

uint32_t ohpo_chunk_out_bytes(const ohp_chunk_desc* d)
{
    const uint32_t b = d->bit_depth / 8u;
    if (b == 0 || d->channels == 0) return 0;
    const uint32_t frames = d->bytes / (b * d->channels);
    switch (d->out_fmt) {
    case OHP_OUT_PACKED_BE: return d->bytes;
    case OHP_OUT_PACKED_LE:
        if (d->aux != OHP_LE_APPEND && (d->flags & OHP_F_RAMP_ENABLED) && b >= 2 && frames != 0) {
            /* SwapEndianness16/24 set the sink's size to the fragment's (TestCodecInteractiveMain.cpp:570-590): after the
             * read it holds the LAST of the fragments ReadBlock made, 256 / frame bytes frames each (Msg.cpp:2765-2779) */
            const uint32_t per_fragment = 256u / (b * d->channels);
            return (frames - ((frames - 1u) / per_fragment) * per_fragment) * b * d->channels;
        }
        return d->bytes;
    case OHP_OUT_PLANAR32_BE: return frames * d->channels * 4u;
    case OHP_OUT_FROM32_BE: return (d->bytes / 4u) * (d->aux / 8u);
    case OHP_OUT_SONGCAST: return frames * (d->channels < 2 ? d->channels : 2u) * (b < 3 ? b : 3u);
    default: return 0;
    }
}

/* Deliver `n` bytes of packed-BE audio (what ReadBlock hands to ProcessFragment/ProcessSilence) to the
 * IPcmProcessor selected by the descriptor.  Returns 0, or -1 where that processor ASSERTs. */
static int sink_write(const ohp_chunk_desc* d, const uint8_t* be, uint32_t n, int silence, uint8_t* out)
{
    const uint32_t b = d->bit_depth / 8u;
    const uint32_t ch = d->channels;
    switch (d->out_fmt) {
    case OHP_OUT_PACKED_BE:
        /* ProcessorPcmBufTest::ProcessFragment/ProcessSilence, ProcessorAudioUtils.cpp:43-53 */
        if (n % (b * ch) != 0) return -1;
        memcpy(out, be, n);
        return 0;
    case OHP_OUT_PACKED_LE:
        /* ProcessorPcmSwpEndianPacked, TestCodecInteractiveMain.cpp:546-590 (32-bit and silence ASSERT) */
        if (silence) return -1;
        if (d->aux > OHP_LE_APPEND) return -1;
        {
            /* aux 0: only what the sink is left holding (see ohpo_chunk_out_bytes) */
            const uint32_t keep = ohpo_chunk_out_bytes(d);
            be += n - keep;
            n = keep;
        }
        if (b == 1) { memcpy(out, be, n); return 0; }
        if (b == 2) { for (uint32_t i = 0; i < n; i += 2) { out[i] = be[i + 1]; out[i + 1] = be[i]; } return 0; }
        if (b == 3) { for (uint32_t i = 0; i < n; i += 3) { out[i] = be[i + 2]; out[i + 1] = be[i + 1]; out[i + 2] = be[i]; } return 0; }
        return -1;
    case OHP_OUT_PLANAR32_BE: {
        /* FlywheelInput::DoProcessFragment + AppendSubsample8/16/24/32, StarvationRamper.cpp:117-186:
         * channel j's plane starts at j * (frames per plane * 4) (Prepare, :92-100) */
        const uint32_t frames = n / (b * ch);
        const uint32_t plane = (uint32_t)d->aux * 4u;
        for (uint32_t i = 0; i < frames; i++) {
            for (uint32_t j = 0; j < ch; j++) {
                uint8_t* o = out + (uint64_t)j * plane + (uint64_t)i * 4u;
                for (uint32_t k = 0; k < 4; k++) o[k] = (k < b) ? *be++ : 0;
            }
        }
        return 0;
    }
    case OHP_OUT_FROM32_BE: {
        /* RampGenerator::ProcessFragment, StarvationRamper.cpp:281-326 (ProcessSilence ASSERTs, :339-343) */
        if (silence) return -1;
        if (b != 4) return -1;
        const uint32_t subsamples = n / 4u;
        const uint32_t ob = d->aux;
        if (ob != 8 && ob != 16 && ob != 24 && ob != 32) return -1;
        for (uint32_t i = 0; i < subsamples; i++, be += 4) {
            switch (ob) {
            case 8:  *out++ = be[0]; break;
            case 16: *out++ = be[0]; *out++ = be[1]; break;
            case 24: *out++ = be[0]; *out++ = be[1]; *out++ = be[2]; break;
            default: *out++ = be[0]; *out++ = be[1]; *out++ = be[2]; *out++ = 0; break;
            }
        }
        return 0;
    }
    case OHP_OUT_SONGCAST: {
        /* Sender::DoProcessFragment, Av/Songcast/Sender.cpp:356-377.  The second memcpy is unconditional;
         * for mono its bytes land beyond the counted output and are overwritten by the next frame, so the
         * visible result is one subsample per frame. */
        const uint32_t first = d->aux;
        const uint32_t stride = b * ch;
        const uint32_t frames = n / stride;
        const uint32_t db = b < 3 ? b : 3u;
        const uint32_t och = ch < 2 ? ch : 2u;
        if (first + och > ch) return -1;
        const uint8_t* s = be + b * first;
        for (uint32_t i = 0; i < frames; i++) {
            memcpy(out, s, db);
            if (och == 2) memcpy(out + db, s + b, db);
            s += stride;
            out += och * db;
        }
        return 0;
    }
    default:
        return -1;
    }
}

static int desc_basic_ok(const ohp_chunk_desc* d)
{
    if (d->bit_depth != 8 && d->bit_depth != 16 && d->bit_depth != 24 && d->bit_depth != 32) return 0;
    if (d->channels == 0 || d->channels > 32) return 0;
    if (d->bytes % ((d->bit_depth / 8u) * d->channels) != 0) return 0;
    if (d->ramp_start > KMAX || d->ramp_end > KMAX) return 0;
    if (!(d->flags & OHP_F_SILENCE) && d->bytes > CELL_MAX) return 0;
    return 1;
}

int64_t ohpo_process_chunks(const ohp_chunk_desc* descs, size_t n, const uint8_t* in, uint64_t in_bytes,
                            uint8_t* out, uint64_t out_bytes)
{
    uint8_t cell[CELL_MAX];
    uint8_t ramped[CELL_MAX];
    for (size_t k = 0; k < n; k++) {
        const ohp_chunk_desc* d = &descs[k];
        if (!desc_basic_ok(d)) return -(int64_t)(k + 1);
        uint8_t* o = out + d->dst_off;
        (void)out_bytes;
        if (d->bytes == 0) continue; /* MsgPlayable::Read: ReadBlock only when iSize > 0 (Msg.cpp:2649) */
        if (d->flags & OHP_F_SILENCE) {
            /* silence may exceed one cell; emit block by block through the sink */
            const uint32_t b = d->bit_depth / 8u;
            uint8_t* tmp = (uint8_t*)malloc(d->bytes);
            if (!tmp) return -(int64_t)(k + 1);
            silence_fill(tmp, d->bytes, d->bit_depth, d->channels);
            (void)b;
            const int rc = sink_write(d, tmp, d->bytes, 1, o);
            free(tmp);
            if (rc) return -(int64_t)(k + 1);
            continue;
        }
        if (d->src_off + d->bytes > in_bytes) return -(int64_t)(k + 1);
        /* a7: the cell holds big-endian data whatever the wire format was */
        ohpo_unpack_to_be(in + d->src_off, cell, d->bytes, d->bit_depth, (d->flags & OHP_F_IN_LITTLE_ENDIAN) != 0);
        if (d->attenuation != OHP_UNITY_ATTENUATION) {
            if (d->bit_depth != 16) return -(int64_t)(k + 1); /* ASSERT(iBitDepth == 16), Msg.cpp:2741 */
            apply_attenuation(cell, d->bytes, d->attenuation);
        }
        const uint8_t* be = cell;
        if (d->flags & OHP_F_RAMP_ENABLED) {
            ramp_apply(cell, ramped, d->bytes, d->bit_depth, d->channels, d->ramp_start, d->ramp_end);
            be = ramped;
        }
        if (sink_write(d, be, d->bytes, 0, o)) return -(int64_t)(k + 1);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* Message model + stage chain                                                                  */

enum { MSG_PCM = 0, MSG_SILENCE = 1 };

typedef struct msg {
    int kind;
    uint32_t size;       /* MsgAudio::iSize (jiffies)   */
    uint32_t offset;     /* MsgAudio::iOffset (jiffies) */
    ohp_ramp ramp;
    uint32_t attenuation;
    uint64_t cell_src;   /* arena offset of the DecodedAudio cell's first byte (PCM) */
    /* MsgSilence */
    uint32_t size_total; /* iSizeJiffiesTotal */
} msg;

typedef struct playable {
    int kind;
    uint32_t size;    /* bytes */
    uint32_t jiffies;
    uint32_t offset;  /* bytes into the cell */
    ohp_ramp ramp;
    uint32_t attenuation;
    uint64_t cell_src;
} playable;

#define QCAP 64
typedef struct stage {
    msg q[QCAP];
    int head, count;
    int mode; /* 0 running, 1 ramping down, 2 ramping up, 3 muted */
    uint32_t current;
    uint32_t remaining;
    uint32_t max_msg;
    uint32_t attenuation;
    uint64_t pos;
    uint32_t next_ev; /* index into the stream's event slice */
    int elem;         /* 0: the bare ramp idiom; 1 Ramper, 2 Muter, 3 StarvationRamper (by the ops of the stage's events); 4 mixed up */
    int halted;       /* Muter::iHalted / StarvationRamper Halted-or-Starting: no PCM since the start or the last halt */
} stage;

typedef struct run {
    const ohp_stream_spec* spec;
    const ohp_ramp_event* ev;
    uint32_t nev;
    uint32_t jps;
    uint32_t frame_bytes;
    stage st[OHP_MAX_STAGES];
    uint32_t block_fill; /* driver block bytes filled so far */
    uint64_t out_bytes;
    ohpo_schedule_result* res;
    size_t cap;
    int err;
} run;

static void q_push_front(stage* s, const msg* m, int* err)
{
    if (s->count == QCAP) { *err = -3; return; }
    s->head = (s->head + QCAP - 1) % QCAP;
    s->q[s->head] = *m;
    s->count++;
}
static void q_push_back(stage* s, const msg* m, int* err)
{
    if (s->count == QCAP) { *err = -3; return; }
    s->q[(s->head + s->count) % QCAP] = *m;
    s->count++;
}
static msg q_pop_front(stage* s)
{
    msg m = s->q[s->head];
    s->head = (s->head + 1) % QCAP;
    s->count--;
    return m;
}

static int msg_split(run* r, msg* m, uint32_t jiffies, msg* rem)
{
    /* MsgAudio::Split, Msg.cpp:1949-1969 */
    if (!(jiffies > 0)) return -1;
    if (!(jiffies < m->size)) return -1;
    *rem = *m;
    rem->offset = m->offset + jiffies;
    rem->size = m->size - jiffies;
    if (m->ramp.enabled) {
        if (ohpo_ramp_split(&m->ramp, jiffies, m->size, &rem->ramp)) return -1;
    } else {
        ramp_reset(&rem->ramp);
    }
    m->size = jiffies;
    if (m->kind == MSG_SILENCE) {
        /* MsgSilence::SplitCompleted, Msg.cpp:2522-2545 (sample block = one sample for PCM silence) */
        const uint32_t block = r->jps;
        const uint32_t extra = m->size % block;
        m->size -= extra;
        m->size_total = (m->size / block) * block;
        rem->size += extra;
        rem->size_total = (rem->size / block) * block;
    }
    /* MsgAudioPcm::SplitCompleted (Msg.cpp:2279-2285): shares the cell, copies iAttenuation -- done by *rem = *m */
    return 0;
}

static int msg_set_ramp(run* r, msg* m, uint32_t start, uint32_t* remaining_duration, uint32_t dir,
                        msg* split, int* have_split, uint32_t* out_end)
{
    /* MsgAudio::SetRamp, Msg.cpp:1989-2046 */
    const uint32_t remaining = *remaining_duration;
    ohp_ramp sp;
    uint32_t split_pos;
    *have_split = 0;
    if (!(dir == OHP_DIR_UP || dir == OHP_DIR_DOWN)) return -1;
    if (m->ramp.enabled && m->ramp.direction == OHP_DIR_MUTE) {
        if (dir == OHP_DIR_DOWN) *remaining_duration = 0;
        *out_end = m->ramp.end;
        return 0;
    }
    const int rc = ohpo_ramp_set(&m->ramp, start, m->size, remaining, dir, &sp, &split_pos);
    if (rc < 0) return -1;
    if (rc == 1) {
        if (split_pos == 0) {
            m->ramp = sp;
        } else if (split_pos != m->size) {
            const ohp_ramp keep = m->ramp;
            if (msg_split(r, m, split_pos, split)) return -1;
            m->ramp = keep;
            split->ramp = sp;
            *have_split = 1;
        }
    }
    *remaining_duration -= m->size;
    if (*have_split && split->ramp.direction != dir && dir == OHP_DIR_UP) {
        *remaining_duration += split->size; /* :2032-2034 */
    }
    if (dir == OHP_DIR_DOWN && m->ramp.end == KMIN) *remaining_duration = 0;
    else if (dir == OHP_DIR_UP && m->ramp.end == KMAX) *remaining_duration = 0;
    *out_end = m->ramp.end;
    return 0;
}

static void msg_set_muted(msg* m)
{
    /* MsgAudio::SetMuted -> Ramp::SetMuted, Msg.cpp:2053-2056, 714-719 */
    m->ramp.start = m->ramp.end = KMIN;
    m->ramp.direction = OHP_DIR_MUTE;
    m->ramp.enabled = 1;
}

static int msg_create_playable(run* r, const msg* m, playable* p)
{
    const ohp_stream_spec* sp = r->spec;
    if (m->kind == MSG_PCM) {
        /* MsgAudioPcm::CreatePlayable, Msg.cpp:2234-2262 */
        uint32_t offset_j = m->offset;
        const uint32_t offset_bytes = ohpo_jiffies_to_bytes(&offset_j, r->jps, sp->channels, sp->bit_depth);
        uint32_t size_j = m->size + (m->offset - offset_j);
        const uint32_t size_bytes = ohpo_jiffies_to_bytes(&size_j, r->jps, sp->channels, sp->bit_depth);
        p->size = size_bytes;
        p->jiffies = m->size;
        p->cell_src = m->cell_src;
        if (m->ramp.direction != OHP_DIR_MUTE) {
            p->kind = MSG_PCM;
            p->offset = offset_bytes;
            p->attenuation = m->attenuation;
            p->ramp = m->ramp;
        } else {
            p->kind = MSG_SILENCE; /* muted audio becomes MsgPlayableSilence with no ramp (:2252-2257) */
            p->offset = 0;
            p->attenuation = OHP_UNITY_ATTENUATION;
            ramp_reset(&p->ramp);
        }
    } else {
        /* MsgSilence::CreatePlayable, Msg.cpp:2472-2492 */
        uint32_t total = m->size_total;
        const uint32_t bytes = ohpo_jiffies_to_bytes(&total, r->jps, sp->channels, sp->bit_depth);
        if (bytes > 0 && (total % r->jps) != 0) return -1;
        p->kind = MSG_SILENCE;
        p->size = bytes;
        p->jiffies = m->size;
        p->offset = 0;
        p->attenuation = OHP_UNITY_ATTENUATION;
        p->ramp = m->ramp;
        p->cell_src = 0;
    }
    return 0;
}

static int playable_split(run* r, playable* p, uint32_t bytes, playable* rem, int* have_rem)
{
    /* MsgPlayable::Split, Msg.cpp:2591-2624 */
    *have_rem = 0;
    if (!(bytes <= p->size)) return -1;
    if (!(bytes != 0)) return -1;
    if (bytes == p->size) return 0;
    const uint32_t num_samples = bytes / r->frame_bytes;
    const uint32_t split_jiffies = num_samples * r->jps;
    *rem = *p;
    rem->offset = p->offset + bytes;
    rem->size = p->size - bytes;
    rem->jiffies = p->jiffies - split_jiffies;
    /* MsgPlayablePcm::SplitCompleted (Msg.cpp:2803-2807) hands over only the cell: the remainder comes fresh
     * from the allocator, where Clear() left iAttenuation at unity (Msg.cpp:2809-2814).  A driver split
     * therefore DROPS the attenuation on the second part -- a reference quirk kept for parity. */
    rem->attenuation = OHP_UNITY_ATTENUATION;
    if (p->ramp.enabled) {
        if (ohpo_ramp_split(&p->ramp, bytes, p->size, &rem->ramp)) return -1;
    } else {
        ramp_reset(&rem->ramp);
    }
    p->size = bytes;
    p->jiffies = split_jiffies;
    *have_rem = 1;
    return 0;
}

static void emit(run* r, const playable* p)
{
    ohpo_schedule_result* res = r->res;
    if (res->num_chunks == r->cap) {
        r->cap = r->cap ? r->cap * 2 : 1024;
        res->chunks = (ohp_chunk_desc*)realloc(res->chunks, r->cap * sizeof(ohp_chunk_desc));
        res->info = (ohp_chunk_info*)realloc(res->info, r->cap * sizeof(ohp_chunk_info));
    }
    ohp_chunk_desc* d = &res->chunks[res->num_chunks];
    ohp_chunk_info* ci = &res->info[res->num_chunks];
    memset(d, 0, sizeof *d);
    const ohp_stream_spec* sp = r->spec;
    d->src_off = (p->kind == MSG_PCM) ? sp->src_base + p->cell_src + p->offset : 0;
    d->dst_off = sp->dst_base + r->out_bytes;
    d->bytes = p->size;
    d->ramp_start = (uint16_t)p->ramp.start;
    d->ramp_end = (uint16_t)p->ramp.end;
    d->attenuation = (uint16_t)p->attenuation;
    d->bit_depth = (uint8_t)sp->bit_depth;
    d->channels = (uint8_t)sp->channels;
    d->flags = (uint8_t)((p->ramp.enabled ? OHP_F_RAMP_ENABLED : 0) | (p->kind == MSG_SILENCE ? OHP_F_SILENCE : 0) |
                         ((sp->in_little_endian && p->kind == MSG_PCM) ? OHP_F_IN_LITTLE_ENDIAN : 0));
    d->out_fmt = (uint8_t)sp->out_fmt;
    d->aux = sp->out_fmt == OHP_OUT_PACKED_LE ? OHP_LE_APPEND : 0;
    ci->direction = p->ramp.direction;
    ci->jiffies = p->jiffies;
    r->out_bytes += p->size;
    res->num_chunks++;
}

static void drive(run* r, const msg* m)
{
    /* PreDriver::ProcessMsg -> CreatePlayable (PreDriver.cpp:115-133), then a driver that pulls fixed blocks
     * and Split()s playables to fit (e.g. DriverSongcastSender.cpp:176-199). */
    playable p, rem;
    if (msg_create_playable(r, m, &p)) { r->err = -1; return; }
    const uint32_t block = r->spec->driver_block_frames * r->frame_bytes;
    if (block == 0 || p.size == 0) { emit(r, &p); return; }
    for (;;) {
        const uint32_t room = block - r->block_fill;
        if (p.size > room) {
            int have = 0;
            if (playable_split(r, &p, room, &rem, &have)) { r->err = -1; return; }
            emit(r, &p);
            r->block_fill = 0;
            p = rem;
        } else {
            emit(r, &p);
            r->block_fill += p.size;
            if (r->block_fill == block) r->block_fill = 0;
            return;
        }
    }
}

/* Which of the reference's elements a stage is: decided by the ops of its events (include/ohp_schedule.h). */
static int stage_element(const ohp_ramp_event* ev, uint32_t nev, uint32_t si)
{
    int elem = 0, bare = 0;
    for (uint32_t i = 0; i < nev; i++) {
        if (ev[i].stage != si) continue;
        int of = 0;
        switch (ev[i].op) {
        case OHP_EV_RAMPER_STREAM: of = 1; break;
        case OHP_EV_MUTER_MUTE: case OHP_EV_MUTER_UNMUTE: of = 2; break;
        case OHP_EV_STARVATION: of = 3; break;
        case OHP_EV_RAMP_DOWN: case OHP_EV_RAMP_UP: case OHP_EV_MUTE: case OHP_EV_UNMUTE: bare = 1; break;
        default: break;
        }
        if (of) {
            if (elem && elem != of) return 4;
            elem = of;
        }
    }
    return (elem && bare) ? 4 : elem;
}

/* The elements' own calls and control messages; returns -1 where the reference ASSERTS. */
static int apply_element_event(stage* s, const ohp_ramp_event* e)
{
    switch (e->op) {
    case OHP_EV_RAMPER_STREAM: /* Ramper::ProcessMsg(MsgDecodedStream), Ramper.cpp:72-93 */
        if (e->arg != 0) { s->mode = 2; s->current = KMIN; s->remaining = e->arg; }
        else { s->mode = 0; s->current = KMAX; s->remaining = 0; }
        break;
    case OHP_EV_MUTER_MUTE: /* Muter::Mute, Muter.cpp:57-99 */
        if (s->mode == 0) {
            if (s->halted) s->mode = 3;
            else { s->mode = 1; s->remaining = e->arg; s->current = KMAX; }
        } else if (s->mode == 2) {
            if (s->remaining == e->arg) s->mode = 3;
            else { s->mode = 1; s->remaining = e->arg - s->remaining; }
        } else return -1; /* ASSERTS(), Muter.cpp:88-90 */
        break;
    case OHP_EV_MUTER_UNMUTE: /* Muter::Unmute, Muter.cpp:101-137 */
        if (s->mode == 1) {
            if (s->remaining == e->arg) s->mode = 0;
            else { s->mode = 2; s->remaining = e->arg - s->remaining; }
        } else if (s->mode == 3) {
            if (s->halted) s->mode = 0;
            else { s->mode = 2; s->remaining = e->arg; s->current = KMIN; }
        } else return -1; /* ASSERTS(), Muter.cpp:107-111 */
        break;
    case OHP_EV_HALT:
        if (s->elem == 1) { /* Ramper::ProcessMsg(MsgHalt), Ramper.cpp:65-70 */
            if (s->mode == 2) s->mode = 0;
        } else if (s->elem == 2) { /* Muter::ProcessMsg(MsgHalt) -> BeginHalting; the animator reports halted: Muter.cpp:159-167, 264-280 */
            if (s->mode == 1) { s->mode = 3; s->remaining = 0; s->current = KMIN; }
            s->halted = 1;
        } else if (s->elem == 3) { /* StarvationRamper::ProcessMsgOut(MsgHalt), StarvationRamper.cpp:728-736 */
            s->mode = 0;
            s->halted = 1;
        }
        break;
    case OHP_EV_STARVATION: /* StarvationRamper::Pull on an empty reservoir, StarvationRamper.cpp:622-673 */
        if ((s->mode == 0 && !s->halted) || (s->mode == 2 && s->current != KMIN)) {
            s->mode = 2; s->current = KMIN; s->remaining = e->arg;
        }
        break;
    default: break;
    }
    return 0;
}

/* What a MsgSilence passing through does to the element; the message itself is handed on untouched. */
static void element_sees_silence(stage* s)
{
    if (s->elem == 1) { /* Ramper::ProcessMsg(MsgSilence), Ramper.cpp:106-112 */
        s->mode = 0; s->current = KMAX; s->remaining = 0;
    } else if (s->elem == 2) { /* Muter::ProcessMsg(MsgSilence), Muter.cpp:188-208 */
        if (s->mode == 1) { s->mode = 3; s->remaining = 0; s->current = KMIN; }
        else if (s->mode == 2) { s->mode = 0; s->remaining = 0; s->current = KMAX; }
    }
    /* StarvationRamper::ProcessMsgOut(MsgSilence), StarvationRamper.cpp:893-911: a ramp up in progress waits for the next audio */
}

static int apply_event(stage* s, const ohp_ramp_event* e)
{
    switch (e->op) {
    case OHP_EV_RAMP_DOWN:
        if (s->mode == 3 || s->current == KMIN) { s->mode = 3; s->current = KMIN; s->remaining = 0; }
        else { s->mode = 1; s->remaining = e->arg; }
        break;
    case OHP_EV_RAMP_UP:
        if (s->mode == 0 && s->current == KMAX) { /* already at full level */ }
        else { s->mode = 2; s->remaining = e->arg; }
        break;
    case OHP_EV_MUTE: s->mode = 3; s->current = KMIN; s->remaining = 0; break;
    case OHP_EV_UNMUTE: s->mode = 0; s->current = KMAX; s->remaining = 0; break;
    case OHP_EV_SET_ATTENUATION: s->attenuation = e->arg; break;
    case OHP_EV_MAX_MSG_JIFFIES: s->max_msg = e->arg; break;
    default: return apply_element_event(s, e);
    }
    return 0;
}

static int next_stage_event(run* r, int si, uint32_t* idx)
{
    stage* s = &r->st[si];
    while (s->next_ev < r->nev) {
        const ohp_ramp_event* e = &r->ev[s->next_ev];
        if (e->stage == (uint32_t)si && e->op != OHP_EV_INSERT_SILENCE) { *idx = s->next_ev; return 1; }
        s->next_ev++;
    }
    return 0;
}

static void feed(run* r, int si, const msg* in);

static void stage_process(run* r, int si, msg* m)
{
    stage* s = &r->st[si];
    uint32_t ei;
    msg rem;
    /* events due at or before the current position fire first */
    while (next_stage_event(r, si, &ei) && r->ev[ei].at_jiffies <= s->pos) {
        if (apply_event(s, &r->ev[ei])) { r->err = -1; return; }
        s->next_ev = ei + 1;
    }
    /* an event inside this message: Split() there; the remainder is re-queued at the head */
    if (next_stage_event(r, si, &ei) && r->ev[ei].at_jiffies < s->pos + m->size) {
        uint32_t at = (uint32_t)(r->ev[ei].at_jiffies - s->pos);
        if (m->kind == MSG_SILENCE) at -= at % r->jps; /* silence only splits on sample blocks */
        if (at == 0) {
            if (apply_event(s, &r->ev[ei])) { r->err = -1; return; }
            s->next_ev = ei + 1;
        } else {
            if (msg_split(r, m, at, &rem)) { r->err = -1; return; }
            q_push_front(s, &rem, &r->err);
        }
    }
    if (s->max_msg != 0 && m->size > s->max_msg) {
        /* StarvationRamper::ProcessMsgOut, StarvationRamper.cpp:802-805 */
        if (s->max_msg < r->jps) { r->err = -2; return; } /* a silence split would make no progress */
        if (msg_split(r, m, s->max_msg, &rem)) { r->err = -1; return; }
        q_push_front(s, &rem, &r->err);
    }
    if (m->kind == MSG_PCM && s->attenuation != OHP_UNITY_ATTENUATION) {
        m->attenuation = s->attenuation; /* Attenuator::ProcessMsg, Attenuator.cpp:55-58 */
    }
    if (s->elem != 0) {
        if (m->kind == MSG_SILENCE) {
            element_sees_silence(s);
            s->pos += m->size;
            return;
        }
        s->halted = 0; /* Muter::ProcessAudio, Muter.cpp:212; StarvationRamper::ProcessMsgOut(MsgAudioPcm), StarvationRamper.cpp:797-799 */
    }
    if (s->mode == 1 || s->mode == 2) {
        if (s->remaining > 0) {
            /* the shared idiom: Ramper.cpp:114-134, Muter.cpp:221-247, StarvationRamper.cpp:807-826 */
            if (m->size > s->remaining) {
                if (msg_split(r, m, s->remaining, &rem)) { r->err = -1; return; }
                q_push_front(s, &rem, &r->err);
                /* a MsgSilence split below one sample leaves a zero-length first part: the reference
                 * then ASSERTs in Ramp::Set's validation or stops making progress; treat as ASSERT */
                if (m->size == 0) { r->err = -1; return; }
            }
            msg split;
            int have_split = 0;
            const uint32_t dir = (s->mode == 1) ? OHP_DIR_DOWN : OHP_DIR_UP;
            if (msg_set_ramp(r, m, s->current, &s->remaining, dir, &split, &have_split, &s->current)) { r->err = -1; return; }
            if (have_split) q_push_front(s, &split, &r->err);
        }
        if (s->remaining == 0) {
            if (s->mode == 2) { s->mode = 0; s->current = KMAX; }
            else { s->mode = 3; s->current = KMIN; }
        }
    } else if (s->mode == 3) {
        msg_set_muted(m); /* Muter eMuted, Muter.cpp:257-259 */
    }
    s->pos += m->size;
}

static void feed(run* r, int si, const msg* in)
{
    if (r->err) return;
    if (si == (int)OHP_MAX_STAGES) { drive(r, in); return; }
    stage* s = &r->st[si];
    q_push_back(s, in, &r->err);
    while (s->count > 0 && !r->err) {
        msg m = q_pop_front(s);
        stage_process(r, si, &m);
        if (r->err) return;
        feed(r, si + 1, &m);
    }
}

static int run_stream(run* r)
{
    const ohp_stream_spec* sp = r->spec;
    r->jps = ohpo_jiffies_per_sample(sp->sample_rate);
    if (r->jps == 0) return -2;
    if (sp->bit_depth != 8 && sp->bit_depth != 16 && sp->bit_depth != 24 && sp->bit_depth != 32) return -2;
    if (sp->channels == 0 || sp->channels > 32) return -2;
    r->frame_bytes = sp->channels * (sp->bit_depth / 8u);
    if (sp->chunk_frames == 0 || (uint64_t)sp->chunk_frames * r->frame_bytes > CELL_MAX) return -2;
    if (sp->total_frames > (~0ull) / r->frame_bytes) return -2;
    for (unsigned i = 0; i < OHP_MAX_STAGES; i++) {
        stage* s = &r->st[i];
        memset(s, 0, sizeof *s);
        s->current = KMAX;
        s->attenuation = OHP_UNITY_ATTENUATION;
        s->elem = stage_element(r->ev, r->nev, i);
        s->halted = 1;
        if (s->elem == 4) return -2;
        if (s->elem == 3) s->max_msg = 5u * OHP_JIFFIES_PER_MS; /* kMaxAudioOutJiffies, StarvationRamper.cpp:376 */
    }
    r->block_fill = 0;
    r->out_bytes = 0;
    uint64_t frame = 0;
    uint64_t src_jiffies = 0; /* PCM jiffies fed so far */
    uint32_t sil_ev = 0;
    /* message sizes: codec reads -> CodecController::OutputAudioPcm pieces (CodecController.cpp:800-827) ->
     * DecodedAudioAggregator::TryAggregate (DecodedAudioAggregator.cpp:134-186) */
    uint64_t total_left = sp->total_frames;
    uint32_t read_left = 0, held = 0;
    while (!r->err) {
        uint32_t frames = 0;
        if (sp->codec_read_frames == 0) {
            frames = (uint32_t)(total_left < sp->chunk_frames ? total_left : sp->chunk_frames);
            total_left -= frames;
        } else {
            while (frames == 0) {
                if (total_left == 0) { frames = held; held = 0; break; } /* OutputAggregatedAudio, :188-194 */
                if (read_left == 0) read_left = (uint32_t)(total_left < sp->codec_read_frames ? total_left : sp->codec_read_frames);
                const uint32_t piece = read_left < sp->chunk_frames ? read_left : sp->chunk_frames;
                read_left -= piece;
                total_left -= piece;
#define AGG_FULL(f) ((f) * r->frame_bytes == CELL_MAX || (f) * r->jps >= 5u * OHP_JIFFIES_PER_MS - 7680u) /* :129-132, .h:19 */
                if (held == 0) {
                    if (AGG_FULL(piece)) frames = piece; else held = piece;
                } else if ((held + piece) * r->frame_bytes <= CELL_MAX) {
                    held += piece;
                    if (AGG_FULL(held)) { frames = held; held = 0; }
                } else {
                    frames = held;
                    held = piece;
                }
#undef AGG_FULL
            }
        }
        if (frames == 0) break;
        /* MsgSilence entering ahead of the next PCM message */
        for (; sil_ev < r->nev && !r->err; sil_ev++) {
            const ohp_ramp_event* e = &r->ev[sil_ev];
            if (e->op != OHP_EV_INSERT_SILENCE) continue;
            if (e->at_jiffies > src_jiffies) break;
            msg m;
            memset(&m, 0, sizeof m);
            m.kind = MSG_SILENCE;
            uint32_t j = e->arg;
            round_down_non_zero_sample_block(&j, r->jps); /* MsgSilence::Initialise, Msg.cpp:2547-2560 */
            m.size = j;
            m.size_total = j;
            m.offset = 0;
            ramp_reset(&m.ramp);
            m.attenuation = OHP_UNITY_ATTENUATION;
            feed(r, 0, &m);
        }
        if (r->err) break;
        msg m;
        memset(&m, 0, sizeof m);
        m.kind = MSG_PCM;
        /* MsgAudioDecoded::Initialise, Msg.cpp:2155-2168 */
        m.size = frames * r->jps;
        m.offset = 0;
        ramp_reset(&m.ramp);
        m.attenuation = OHP_UNITY_ATTENUATION; /* MsgAudioPcm::Initialise, Msg.cpp:2275 */
        m.cell_src = frame * r->frame_bytes;
        feed(r, 0, &m);
        frame += frames;
        src_jiffies += (uint64_t)frames * r->jps;
    }
    return r->err;
}

int ohpo_schedule_run(const ohp_stream_spec* streams, size_t n_streams,
                      const ohp_ramp_event* events, size_t n_events, ohpo_schedule_result* out)
{
    memset(out, 0, sizeof *out);
    out->stream_chunk_begin = (uint64_t*)calloc(n_streams + 1, sizeof(uint64_t));
    out->stream_out_bytes = (uint64_t*)calloc(n_streams ? n_streams : 1, sizeof(uint64_t));
    run* r = (run*)calloc(1, sizeof(run));
    r->res = out;
    int rc = 0;
    for (size_t s = 0; s < n_streams; s++) {
        const ohp_stream_spec* sp = &streams[s];
        if ((uint64_t)sp->first_event + sp->num_events > n_events) { rc = -2; break; }
        r->spec = sp;
        r->ev = events + sp->first_event;
        r->nev = sp->num_events;
        r->err = 0;
        out->stream_chunk_begin[s] = out->num_chunks;
        rc = run_stream(r);
        if (rc) break;
        out->stream_out_bytes[s] = r->out_bytes;
    }
    out->stream_chunk_begin[n_streams] = out->num_chunks;
    free(r);
    if (rc) ohpo_schedule_result_free(out);
    return rc;
}

void ohpo_schedule_result_free(ohpo_schedule_result* r)
{
    free(r->chunks);
    free(r->info);
    free(r->stream_chunk_begin);
    free(r->stream_out_bytes);
    memset(r, 0, sizeof *r);
}

/* ------------------------------------------------------------------------------------------- */

uint64_t ohpo_checksum(const uint8_t* data, uint64_t n)
{
    uint64_t sum = 0;
    for (uint64_t i = 0; i < n; i++) sum += ((uint64_t)data[i] + 1u) * (i + 1u);
    return sum;
}

void ohpo_fill_pcm(uint8_t* dst, uint64_t bytes, uint64_t seed)
{
    /* splitmix64 (public-domain constants), 8 bytes per step, little-endian */
    uint64_t x = seed;
    uint64_t i = 0;
    while (i < bytes) {
        x += 0x9E3779B97F4A7C15ull;
        uint64_t z = x;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z = z ^ (z >> 31);
        for (int k = 0; k < 8 && i < bytes; k++, i++) dst[i] = (uint8_t)(z >> (8 * k));
    }
}

/* ==================================================================================================================
 * Flywheel ramp generator: Media/FlywheelRamper.cpp + RampGenerator (Media/Pipeline/StarvationRamper.cpp).
 * The reference computes in TInt16 / TInt32 with wrap-around on overflow (two's complement on every target it
 * builds for); the sums below are done in unsigned arithmetic and cast back so that C leaves nothing undefined. */

static int16_t wrap16(int32_t v) { return (int16_t)(uint16_t)(uint32_t)v; }
static int32_t mul32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }

void ohpo_burgs_method(const int16_t* samples, uint32_t n, uint32_t degree, int16_t* output, int16_t* h,
                       int16_t* per, int16_t* pef)
{
    /* FlywheelRamper::BurgsMethod, FlywheelRamper.cpp:253-328; kBurgScaleShift = 16 - kBurgOutputFormat = 13 (:11-12) */
    const uint32_t shift = 13;
    uint32_t limit1 = n - 1;
    uint32_t limit2 = limit1;
    for (uint32_t k = 0; k < degree; k++) {
        uint32_t sn = 0, sd = 0; /* TInt32 accumulators, :260-261 */
        for (uint32_t j = 0; j < limit1; j++) {
            const int16_t t1 = wrap16((int32_t)samples[j + k + 1] + pef[j]); /* :265 */
            const int16_t t2 = wrap16((int32_t)samples[j] + per[j]);         /* :266 */
            const int32_t t1t1 = (int32_t)t1 * t1;
            const int32_t t2t2 = (int32_t)t2 * t2;
            const int32_t t1t2 = (int32_t)t1 * t2;
            sn -= 2u * (uint32_t)t1t2;                     /* :270 */
            sd += (uint32_t)t1t1 + (uint32_t)t2t2;         /* :271 */
        }
        limit1--;
        int16_t t3 = 0;
        if ((int32_t)sn != 0) {
            /* :278-282.  sd == 0 with sn != 0 needs the accumulator to wrap to exactly zero; the reference would
             * divide by zero there -- t3 stays 0 in this restatement (documented deviation). */
            if ((int32_t)sd != 0) {
                const int64_t ratio = (int64_t)((uint64_t)(int64_t)(int32_t)sn << shift) / (int64_t)(int32_t)sd;
                t3 = (int16_t)(uint16_t)(uint64_t)ratio;
            }
        }
        output[k] = t3;
        if (k > 0) {
            for (uint32_t j = 0; j < k; j++) {
                const int32_t prod = (int32_t)t3 * output[k - j - 1];              /* :290 */
                h[j] = wrap16(prod >> shift);                                       /* :291 */
                h[j] = wrap16((int32_t)h[j] + output[j]);                           /* :292 */
            }
            for (uint32_t j = 0; j < k; j++) output[j] = h[j];
            limit2--;
        }
        if (k == degree - 1) break;
        for (uint32_t j = 0; j < limit2; j++) {
            const uint32_t i = j + 1;
            int32_t p = (int32_t)pef[j] + samples[i + k];                           /* :312 */
            p = mul32(p, t3);
            per[j] = wrap16((int32_t)per[j] + wrap16(p >> shift));                  /* :314 */
            int32_t f = (int32_t)per[i] + samples[i];                               /* :316 */
            f = mul32(f, t3);
            pef[j] = wrap16(f >> shift);                                            /* :318 */
            pef[j] = wrap16((int32_t)pef[j] + pef[i]);                              /* :319 */
        }
    }
}

int16_t ohpo_coeff_overflow(const int16_t* coeffs, uint32_t count, uint32_t format)
{
    /* FlywheelRamper::CoeffOverflow, FlywheelRamper.cpp:355-388 */
    const int16_t one = (int16_t)(1 << (16 - format));
    int16_t total = 0;
    for (uint32_t j = 0; j < count; j++) total = wrap16((int32_t)total + coeffs[j]);
    if (total <= one && total >= -one) return 0;
    if (total & 0x8000) return wrap16((int32_t)total + one);
    return wrap16((int32_t)total - one);
}

void ohpo_feedback_init(ohpo_feedback* f, uint32_t state_count, uint32_t descale_bits, uint32_t coeff_format,
                        uint32_t data_format, uint32_t output_format, int32_t* coeffs, int32_t* samples)
{
    /* FeedbackModel::FeedbackModel + Initialise, FlywheelRamper.cpp:420-445 */
    f->coeffs = coeffs;
    f->samples = samples;
    f->state_count = state_count;
    f->descale_bits = descale_bits;
    f->coeff_format = coeff_format;
    f->scale_shift_for_output = (int32_t)(data_format + descale_bits) - (int32_t)output_format;
    for (uint32_t j = 0; j < state_count; j++) samples[j] >>= descale_bits;
}

int32_t ohpo_feedback_next(ohpo_feedback* f)
{
    /* FeedbackModel::NextSample, FlywheelRamper.cpp:447-487 */
    uint32_t sum = 0;
    for (uint32_t j = 0; j < f->state_count; j++) {
        const int64_t product = (int64_t)f->samples[j] * (int64_t)f->coeffs[j];
        sum += (uint32_t)(int32_t)(product >> 32);
    }
    for (uint32_t j = f->state_count - 1; j > 0; j--) f->samples[j] = f->samples[j - 1];
    sum <<= f->coeff_format;
    f->samples[0] = (int32_t)sum;
    int32_t out = (int32_t)sum;
    if (f->scale_shift_for_output < 0) out >>= -f->scale_shift_for_output;
    else out = (int32_t)((uint32_t)out << f->scale_shift_for_output);
    return out;
}

uint32_t ohpo_decimation_factor(uint32_t rate)
{
    /* FlywheelRamper::DecimationFactor, FlywheelRamper.cpp:330-345 */
    if (rate == 192000 || rate == 176400) return 4;
    if (rate == 88200 || rate == 96000) return 2;
    return 1;
}

uint32_t ohpo_flywheel_out_bytes(const ohp_flywheel_job* job)
{
    return job->out_frames * job->channels * (job->bit_depth / 8u);
}

static int flywheel_job_ok(const ohp_flywheel_job* job, uint64_t in_bytes, uint64_t out_bytes)
{
    const uint32_t jps = ohpo_jiffies_per_sample(job->sample_rate);
    if (jps == 0) return 0;
    if (!(job->bit_depth == 8 || job->bit_depth == 16 || job->bit_depth == 24 || job->bit_depth == 32)) return 0; /* StarvationRamper.cpp:322-324 */
    if (job->channels < 1 || job->channels > OHP_FLYWHEEL_MAX_CHANNELS) return 0;
    /* FlywheelRamper::Initialise wants exactly the training length (FlywheelRamper.cpp:180-195; with more it reads
     * past its channel's block) */
    if (job->train_frames != OHP_FLYWHEEL_TRAINING_JIFFIES / jps) return 0;
    if (job->train_frames / ohpo_decimation_factor(job->sample_rate) < OHP_FLYWHEEL_DEGREE + 1u) return 0;
    if ((uint32_t)job->train_frames * 4u * job->channels > OHP_FLYWHEEL_MAX_INPUT_BYTES) return 0;
    const uint32_t block_frames = OHP_FLYWHEEL_BLOCK_JIFFIES / jps;
    if (block_frames * job->channels * (job->bit_depth / 8u) > OHP_FLYWHEEL_MAX_BLOCK_BYTES) return 0; /* Bwh capacity ASSERT */
    if (block_frames * job->channels * 4u > 384u * 10u * 4u) return 0; /* FlywheelRamperManager::iOutBuf, FlywheelRamper.cpp:28 */
    const uint64_t in_need = (uint64_t)job->train_frames * 4u * job->channels;
    if (job->src_off > in_bytes || in_need > in_bytes - job->src_off) return 0;
    const uint64_t out_need = ohpo_flywheel_out_bytes(job);
    if (job->dst_off > out_bytes || out_need > out_bytes - job->dst_off) return 0;
    return 1;
}

int64_t ohpo_flywheel(const ohp_flywheel_job* jobs, size_t n, const uint8_t* in, uint64_t in_bytes,
                      uint8_t* out, uint64_t out_bytes)
{
    for (size_t q = 0; q < n; q++) {
        const ohp_flywheel_job* job = &jobs[q];
        if (!flywheel_job_ok(job, in_bytes, out_bytes)) return -(int64_t)(q + 1);
        const uint32_t ch = job->channels;
        const uint32_t dec = ohpo_decimation_factor(job->sample_rate);
        const uint32_t jps = ohpo_jiffies_per_sample(job->sample_rate);
        const uint32_t bytes_per_chan = (uint32_t)job->train_frames * 4u;
        const uint32_t ob = job->bit_depth / 8u;
        int16_t input[OHP_FLYWHEEL_MAX_TRAIN_FRAMES], per[OHP_FLYWHEEL_MAX_TRAIN_FRAMES], pef[OHP_FLYWHEEL_MAX_TRAIN_FRAMES];
        int16_t burg[OHP_FLYWHEEL_DEGREE], h[OHP_FLYWHEEL_DEGREE];
        int32_t fb_samples[OHP_FLYWHEEL_MAX_CHANNELS][OHP_FLYWHEEL_DEGREE], fb_coeffs[OHP_FLYWHEEL_MAX_CHANNELS][OHP_FLYWHEEL_DEGREE];
        ohpo_feedback fb[OHP_FLYWHEEL_MAX_CHANNELS];
        /* FlywheelRamperManager::InitChannels (FlywheelRamper.cpp:70-83) -> FlywheelRamper::Initialise (:178-231) */
        for (uint32_t c = 0; c < ch; c++) {
            const uint8_t* p = in + job->src_off + (uint64_t)c * bytes_per_chan;
            const uint32_t count = bytes_per_chan / (4u * dec);
            memset(per, 0, sizeof per);
            memset(pef, 0, sizeof pef);
            for (uint32_t i = 0; i < count; i++) {
                const uint32_t hi = ((uint32_t)p[0] << 8) + p[1];
                const int16_t s16 = (int16_t)(uint16_t)hi;
                const uint32_t full = (hi << 16) + ((uint32_t)p[2] << 8) + p[3];
                p += 4u * dec;
                if (i >= count - OHP_FLYWHEEL_DEGREE) fb_samples[c][count - i - 1] = (int32_t)full; /* reverse order, :214-218 */
                input[i] = (int16_t)(s16 >> 1); /* kBurgDataDescaleBitCount = 1, :220 */
            }
            ohpo_burgs_method(input, count, OHP_FLYWHEEL_DEGREE, burg, h, per, pef);
            const int16_t excess = ohpo_coeff_overflow(burg, OHP_FLYWHEEL_DEGREE, 3); /* CorrectBurgCoeffs, :347-353 */
            if (excess != 0) burg[0] = wrap16((int32_t)burg[0] - (int32_t)excess * 2);
            for (uint32_t i = 0; i < OHP_FLYWHEEL_DEGREE; i++) {
                fb_coeffs[c][i] = (int32_t)(0u - ((uint32_t)(int32_t)burg[i] << 16)); /* PrepareFeedbackCoeffs, :233-240 */
            }
            /* FeedbackModel(iDegree, 0, kBurgOutputFormat = 3, kFeedbackDataFormat = 1, 1), :150 */
            ohpo_feedback_init(&fb[c], OHP_FLYWHEEL_DEGREE, 0, 3, 1, 1, fb_coeffs[c], fb_samples[c]);
        }
        /* FlywheelRamperManager::Ramp (:46-68): 1 ms blocks through RenderChannels (:85-135), each delivered to
         * RampGenerator::ProcessFragment (StarvationRamper.cpp:281-326) */
        const uint32_t block_frames = OHP_FLYWHEEL_BLOCK_JIFFIES / jps;
        uint32_t remaining = job->out_frames;
        uint8_t* dst = out + job->dst_off;
        int32_t prev[OHP_FLYWHEEL_MAX_CHANNELS];
        while (remaining > 0) {
            const uint32_t frames = remaining > block_frames ? block_frames : remaining;
            remaining -= frames;
            uint32_t hold = 0; /* restarts with every block, :89 */
            for (uint32_t j = 0; j < frames; j++) {
                for (uint32_t c = 0; c < ch; c++) {
                    int32_t sample;
                    if (hold == 0) {
                        sample = ohpo_feedback_next(&fb[c]);
                        prev[c] = sample;
                    } else {
                        sample = prev[c];
                    }
                    const uint32_t u = (uint32_t)sample;
                    const uint8_t b[4] = {(uint8_t)(u >> 24), (uint8_t)(u >> 16), (uint8_t)(u >> 8), (uint8_t)u};
                    switch (ob) {
                    case 1: *dst++ = b[0]; break;
                    case 2: *dst++ = b[0]; *dst++ = b[1]; break;
                    case 3: *dst++ = b[0]; *dst++ = b[1]; *dst++ = b[2]; break;
                    default: *dst++ = b[0]; *dst++ = b[1]; *dst++ = b[2]; *dst++ = 0; break; /* :313-321 */
                    }
                }
                if (++hold == dec) hold = 0;
            }
        }
    }
    return 0;
}

int ohpo_flywheel_ramp_chunks(const ohp_flywheel_job* job, uint32_t current_ramp, uint64_t src_off, uint64_t dst_off,
                              ohp_chunk_desc* out, uint32_t cap, uint32_t* final_ramp)
{
    /* RampGenerator::Start (StarvationRamper.cpp:235-247) */
    const uint32_t jps = ohpo_jiffies_per_sample(job->sample_rate);
    if (jps == 0) return -1;
    const uint32_t frame_bytes = job->channels * (job->bit_depth / 8u);
    const uint32_t block_frames = OHP_FLYWHEEL_BLOCK_JIFFIES / jps;
    uint32_t remaining_ramp = jps * job->out_frames;
    uint32_t remaining = job->out_frames;
    uint32_t n = 0;
    while (remaining > 0) {
        const uint32_t frames = remaining > block_frames ? block_frames : remaining;
        remaining -= frames;
        if (n == cap) return -1;
        /* RampGenerator::EndBlock (StarvationRamper.cpp:351-364) on a fresh MsgAudioPcm, then CreatePlayable
         * (Msg.cpp:2234-2262) */
        ohp_chunk_desc* d = &out[n++];
        memset(d, 0, sizeof *d);
        d->dst_off = dst_off;
        d->bytes = frames * frame_bytes;
        d->attenuation = OHP_UNITY_ATTENUATION;
        d->bit_depth = job->bit_depth;
        d->channels = job->channels;
        d->out_fmt = OHP_OUT_PACKED_BE;
        if (current_ramp == KMIN) {
            /* SetMuted -> CreatePlayable hands out silence with no ramp */
            d->flags = OHP_F_SILENCE;
            d->ramp_start = d->ramp_end = KMAX;
        } else {
            ohp_ramp r, split;
            uint32_t split_pos;
            ramp_reset(&r);
            const int rc = ohpo_ramp_set(&r, current_ramp, frames * jps, remaining_ramp, OHP_DIR_DOWN, &split, &split_pos);
            if (rc != 0) return -1; /* ASSERT(split == nullptr), :359 */
            remaining_ramp -= frames * jps;
            if (r.end == KMIN) remaining_ramp = 0;
            current_ramp = r.end;
            d->src_off = src_off;
            d->flags = OHP_F_RAMP_ENABLED;
            d->ramp_start = (uint16_t)r.start;
            d->ramp_end = (uint16_t)r.end;
        }
        src_off += d->bytes;
        dst_off += d->bytes;
    }
    if (final_ramp) *final_ramp = current_ramp;
    return (int)n;
}
