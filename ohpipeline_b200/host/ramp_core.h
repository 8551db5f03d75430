// ramp_core.h -- the ramp algebra of ohPipeline's message model on plain data, compiled for BOTH the host
// (msg_model.cpp: Ramp::Set / Ramp::Split delegate here) and the device (ohp_schedule_kernels.cuh: one thread walks
// one stream's messages).  One statement of the arithmetic, so the host mirror and the GPU schedule builder cannot
// drift apart; it follows the reference exactly (Msg.cpp:568-807) and is pinned by tests/golden/ramp_algebra.npz.
//
// Where the reference ASSERTs these functions return kRampAssert instead of throwing: the host wrapper turns that
// into AssertionFailed, the device kernel into a per-stream error code.
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define OHP_HD __host__ __device__ __forceinline__
#else
#define OHP_HD inline
#endif

namespace ohp {
namespace core {

constexpr uint32_t kRampMax = 16384u; // Ramp::kMax (Msg.h:257)
constexpr uint32_t kRampMin = 0u;     // Ramp::kMin (Msg.h:258)
// Ramp::EDirection (Msg.h:259-265)
constexpr uint32_t kDirNone = 0u, kDirUp = 1u, kDirDown = 2u, kDirMute = 3u;

constexpr int kRampAssert = -1;

struct RampPod
{
    uint32_t start, end, direction, enabled;
};

OHP_HD void ramp_reset(RampPod& r)
{
    r.start = r.end = kRampMax;
    r.direction = kDirNone;
    r.enabled = 0;
}

OHP_HD void ramp_set_muted(RampPod& r)
{
    r.start = r.end = kRampMin;
    r.direction = kDirMute;
    r.enabled = 1;
}

// Ramp::DoValidate, Msg.cpp:745-782
OHP_HD bool ramp_is_valid(const RampPod& r)
{
    if (r.start > kRampMax || r.end > kRampMax) return false;
    switch (r.direction) {
    case kDirNone: return r.start == r.end;
    case kDirUp:   return r.start < r.end;
    case kDirDown: return r.start > r.end;
    case kDirMute: return r.start == kRampMin && r.end == kRampMin;
    default:       return false;
    }
}

// (a * b + add) / d in 64 bits, as the reference computes it; 32-bit divide when the numerator fits (the GPU has no
// 64-bit divider: __udivdi3-style code is ~10x the cost of a 32-bit divide, and most ramp steps fit).
OHP_HD uint32_t mul_add_div(uint32_t a, uint32_t b, uint32_t add, uint32_t d)
{
    const uint64_t num = a * (uint64_t)b + add;
    if ((num >> 32) == 0) return (uint32_t)num / d;
    return (uint32_t)(num / d);
}

// two ramps over the same audio: the quieter one wins at both ends (Msg.cpp:721-734)
OHP_HD void ramp_take_lower(RampPod& r, uint32_t aStart, uint32_t aEnd)
{
    r.start = r.start < aStart ? r.start : aStart;
    r.end = r.end < aEnd ? r.end : aEnd;
    r.direction = (r.start == r.end) ? kDirNone : (r.start > r.end ? kDirDown : kDirUp);
}

// Ramp::Set (Msg.cpp:590-712).  Returns 1 iff aSplit was set (the existing and the requested ramp run in opposite
// directions and cross inside this fragment, which then has to be split at aSplitPos), 0 if not, kRampAssert where
// the reference ASSERTs.
OHP_HD int ramp_set(RampPod& r, uint32_t aStart, uint32_t aFragmentSize, uint32_t aRemainingDuration, uint32_t aDirection,
                    RampPod& aSplit, uint32_t& aSplitPos)
{
    if (!(aRemainingDuration >= aFragmentSize)) return kRampAssert; // Msg.cpp:598
    if (aDirection == kDirNone) return kRampAssert;                 // Msg.cpp:599
    if (aRemainingDuration == 0) return kRampAssert;                // the reference divides by it
    r.enabled = 1;
    ramp_reset(aSplit);
    aSplitPos = 0xffffffffu;

    // How far this fragment moves the ramp: its share of what is left, rounded UP so that a ramp always completes
    // within its duration (Msg.cpp:603-605); an overshoot of less than the fragment size is rounding, anything more
    // is a caller bug (Msg.cpp:611, 620).
    const uint32_t distance = (aDirection == kDirDown) ? aStart : kRampMax - aStart;
    const uint32_t delta = mul_add_div(distance, aFragmentSize, aRemainingDuration - 1, aRemainingDuration);
    uint32_t end;
    if (aDirection == kDirDown) {
        if (delta > aStart) {
            if (!(delta - aStart <= aFragmentSize - 1)) return kRampAssert;
            end = kRampMin;
        }
        else {
            end = aStart - delta;
        }
    }
    else {
        if (aStart + delta > kRampMax) {
            if (!(aStart + delta - kRampMax <= aFragmentSize - 1)) return kRampAssert;
            end = kRampMax;
        }
        else {
            end = aStart + delta;
        }
    }

    if (r.direction == kDirNone) {
        r.direction = aDirection;
        r.start = aStart;
        r.end = end;
    }
    else if (r.direction == aDirection) {
        ramp_take_lower(r, aStart, end);
    }
    else {
        // Opposite directions.  Treat both as lines over x in [0, aFragmentSize]; (a0,a1) is the one starting lower.
        // If they cross strictly inside the fragment, the fragment becomes "rise to the crossing" and aSplit becomes
        // "fall from the crossing" (Msg.cpp:637-699).  All in 64-bit signed, truncating division.
        int64_t a0, a1, b0, b1;
        if (r.start < aStart) { a0 = r.start; a1 = r.end; b0 = aStart;  b1 = end; }
        else                  { a0 = aStart;  a1 = end;   b0 = r.start; b1 = r.end; }
        const int64_t slopeDiff = (a1 - a0) - (b1 - b0);
        bool crossed = false;
        if (slopeDiff != 0) {
            const int64_t x = ((int64_t)aFragmentSize * (b0 - a0)) / slopeDiff;
            const int64_t y = ((a1 - a0) * (b0 - a0)) / slopeDiff + a0;
            if (x > 0 && (uint32_t)x < aFragmentSize) {
                crossed = true;
                aSplitPos = (uint32_t)x;
                aSplit.start = (uint32_t)y;
                aSplit.end = r.end < end ? r.end : end;
                aSplit.direction = (aSplit.start == aSplit.end) ? kDirNone : kDirDown;
                aSplit.enabled = 1;
                const uint32_t first = r.start < aStart ? r.start : aStart;
                r.start = first;
                r.end = (uint32_t)y;
                r.direction = (r.start == r.end) ? kDirNone : kDirUp;
            }
        }
        if (!crossed) {
            ramp_take_lower(r, aStart, end);
        }
    }
    if (!ramp_is_valid(r)) return kRampAssert; // Msg.cpp:701-708
    return aSplit.enabled ? 1 : 0;
}

// Ramp::Split (Msg.cpp:784-807): r keeps the first aNewSize of aCurrentSize; the remainder goes to aRest.
// Returns 0, or kRampAssert.
OHP_HD int ramp_split(RampPod& r, uint32_t aNewSize, uint32_t aCurrentSize, RampPod& aRest)
{
    if (aCurrentSize == 0) return kRampAssert;
    aRest.end = r.end;
    aRest.direction = r.direction;
    aRest.enabled = 1;
    // proportional share, truncated; unsigned 32-bit span as in the reference (Msg.cpp:791-798)
    if (r.direction == kDirUp) {
        r.end = r.start + mul_add_div((uint32_t)(r.end - r.start), aNewSize, 0, aCurrentSize);
    }
    else {
        r.end = r.start - mul_add_div((uint32_t)(r.start - r.end), aNewSize, 0, aCurrentSize);
    }
    if (r.start == r.end) {
        r.direction = kDirNone; // also turns the first part of a muted message into an enabled flat ramp at 0
    }
    aRest.start = r.end; // no one-step advance (the reference's FIXME, Msg.cpp:802)
    if (!ramp_is_valid(r)) return kRampAssert;
    if (!ramp_is_valid(aRest)) return kRampAssert;
    return 0;
}

// Jiffies::PerSample (Msg.cpp:424-470); 0 for a rate the reference throws SampleRateInvalid on.
OHP_HD uint32_t jiffies_per_sample_or_zero(uint32_t aSampleRate)
{
    switch (aSampleRate) {
    case 7350: case 8000: case 11025: case 12000: case 14700: case 16000: case 22050: case 24000: case 29400:
    case 32000: case 44100: case 48000: case 88200: case 96000: case 176400: case 192000: case 352800: case 384000:
        return 56448000u / aSampleRate;
    default:
        return 0;
    }
}

// Jiffies::ToBytes (Msg.cpp:476-489): rounds aJiffies down to a whole sample, returns the byte count.
OHP_HD uint32_t jiffies_to_bytes(uint32_t& aJiffies, uint32_t aJiffiesPerSample, uint32_t aNumChannels, uint32_t aBitsPerSubsample)
{
    aJiffies -= aJiffies % aJiffiesPerSample;
    const uint32_t subsamples = (aJiffies / aJiffiesPerSample) * aNumChannels;
    return (subsamples * aBitsPerSubsample + 7) / 8;
}

// Jiffies::RoundDownNonZeroSampleBlock (Msg.cpp:504-514): never returns 0 for a non-zero request.
OHP_HD void round_down_non_zero_sample_block(uint32_t& aJiffies, uint32_t aSampleBlockJiffies)
{
    uint32_t j = aJiffies - aJiffies % aSampleBlockJiffies;
    if (j == 0) {
        j = aJiffies + aSampleBlockJiffies - 1;
        j -= j % aSampleBlockJiffies;
    }
    aJiffies = j;
}

} // namespace core
} // namespace ohp
