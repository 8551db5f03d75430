// ohp_flywheel_kernels.cuh -- device code of the flywheel ramp generator (include/ohp_flywheel.h).
//
// One WARP per starving stream (job):
//   1. per channel, all 32 lanes together: decimate the 1 ms training block (FlywheelRamper::Initialise,
//      Media/FlywheelRamper.cpp:178-231) into shared memory and run the integer Burg recursion of degree 3
//      (BurgsMethod, :253-328).  The two accumulators are sums in wrap-around 32-bit arithmetic, so a warp-shuffle
//      reduction gives exactly the reference's sequential result; the in-place update of the forward / backward
//      prediction errors only ever reads values the reference has not overwritten yet (index j and j+1 while writing
//      j), so "every lane loads, __syncwarp, every lane stores" reproduces it.  Lane c keeps channel c's coefficients
//      (CorrectBurgCoeffs :347-353, PrepareFeedbackCoeffs :233-240) and the last three samples as its filter state;
//   2. lanes 0..channels-1 run their channel's all-pole recurrence (FeedbackModel::NextSample, :447-487: three
//      multiply-highs and a shift per output) 32 frames at a time into a shared-memory tile, holding each value for
//      `decimation` frames and restarting the hold at every 1 ms block like RenderChannels (:85-135);
//   3. all 32 lanes repack the tile to the stream's bit depth (RampGenerator::ProcessFragment,
//      Media/Pipeline/StarvationRamper.cpp:281-326) and store it, 4 bytes per lane when the destination allows.
// Integer only; bit-exact against the reference (oracle/_ref links FlywheelRamper.cpp and StarvationRamper.cpp).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/ohp_flywheel.h"
#include "../host/ramp_core.h"

namespace ohp {
namespace fly {

constexpr uint32_t kWarpsPerCta = 4;
constexpr uint32_t kTileFrames = 32;

// FlywheelRamper::DecimationFactor (FlywheelRamper.cpp:330-345)
OHP_HD uint32_t decimation_factor(uint32_t rate)
{
    if (rate == 192000u || rate == 176400u) return 4u;
    if (rate == 88200u || rate == 96000u) return 2u;
    return 1u;
}

// 0 = ok, 1 = invalid (the reference would ASSERT or overrun one of its fixed buffers), 2 = outside the arenas.
// Shared by ohp_flywheel_validate (host) and the kernel so both refuse exactly the same jobs.
OHP_HD uint32_t check_job(const ohp_flywheel_job& j, uint64_t in_bytes, uint64_t out_bytes)
{
    const uint32_t jps = core::jiffies_per_sample_or_zero(j.sample_rate);
    if (jps == 0) return 1u;
    if (!(j.bit_depth == 8 || j.bit_depth == 16 || j.bit_depth == 24 || j.bit_depth == 32)) return 1u; // StarvationRamper.cpp:322-324
    if (j.channels < 1 || j.channels > OHP_FLYWHEEL_MAX_CHANNELS) return 1u;
    // FlywheelRamper::Initialise wants exactly the training length (FlywheelRamper.cpp:180-195; given more it reads
    // past its channel's block)
    if (j.train_frames != OHP_FLYWHEEL_TRAINING_JIFFIES / jps) return 1u;
    if (j.train_frames / decimation_factor(j.sample_rate) < OHP_FLYWHEEL_DEGREE + 1u) return 1u;
    if ((uint32_t)j.train_frames * 4u * j.channels > OHP_FLYWHEEL_MAX_INPUT_BYTES) return 1u;
    const uint32_t block_frames = OHP_FLYWHEEL_BLOCK_JIFFIES / jps;
    if (block_frames * j.channels * (j.bit_depth / 8u) > OHP_FLYWHEEL_MAX_BLOCK_BYTES) return 1u; // Bwh capacity ASSERT
    const uint64_t in_need = (uint64_t)j.train_frames * 4u * j.channels;
    if (j.src_off > in_bytes || in_need > in_bytes - j.src_off) return 2u;
    const uint64_t out_need = (uint64_t)j.out_frames * j.channels * (j.bit_depth / 8u);
    if (j.dst_off > out_bytes || out_need > out_bytes - j.dst_off) return 2u;
    return 0u;
}

struct FlywheelParams
{
    const ohp_flywheel_job* jobs;
    uint64_t n;
    const uint8_t* in;
    uint64_t in_bytes;
    uint8_t* out;
    uint64_t out_bytes;
    uint32_t* status; // [0] error bits (1 invalid, 2 out of range), [1] first offending job + 1
};

struct WarpScratch
{
    int16_t x[OHP_FLYWHEEL_MAX_TRAIN_FRAMES];    // decimated, descaled training samples of the current channel
    int16_t per[OHP_FLYWHEEL_MAX_TRAIN_FRAMES];  // backward prediction error
    int16_t pef[OHP_FLYWHEEL_MAX_TRAIN_FRAMES];  // forward prediction error
    int32_t tile[kTileFrames * OHP_FLYWHEEL_MAX_CHANNELS];
};

__device__ __forceinline__ int16_t wrap16(int32_t v) { return (int16_t)(uint16_t)(uint32_t)v; }

__device__ __forceinline__ uint32_t warp_sum(uint32_t v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Burg's method, degree 3, on ws.x[0..count) with all 32 lanes; every lane returns the same coefficients.
__device__ __forceinline__ void burg3(WarpScratch& ws, uint32_t count, uint32_t lane, int16_t (&coef)[3])
{
    const int shift = 13; // kBurgScaleShift = 16 - kBurgOutputFormat (FlywheelRamper.cpp:11-12)
    uint32_t limit1 = count - 1;
    uint32_t limit2 = limit1;
    int16_t h[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        uint32_t sn = 0, sd = 0;
        for (uint32_t j = lane; j < limit1; j += 32) {
            const int32_t t1 = wrap16((int32_t)ws.x[j + k + 1] + ws.pef[j]);
            const int32_t t2 = wrap16((int32_t)ws.x[j] + ws.per[j]);
            sn -= 2u * (uint32_t)(t1 * t2);
            sd += (uint32_t)(t1 * t1) + (uint32_t)(t2 * t2);
        }
        sn = warp_sum(sn);
        sd = warp_sum(sd);
        limit1--;
        int16_t t3 = 0;
        // sd == 0 with sn != 0 (the accumulator wrapped to exactly zero) is a division by zero in the reference;
        // t3 stays 0 here, as in the oracle
        if ((int32_t)sn != 0 && (int32_t)sd != 0) {
            const long long ratio = (long long)((unsigned long long)(long long)(int32_t)sn << shift) / (long long)(int32_t)sd;
            t3 = (int16_t)(uint16_t)(unsigned long long)ratio;
        }
        coef[k] = t3;
        if (k > 0) {
#pragma unroll
            for (int j = 0; j < k; j++) {
                const int32_t prod = (int32_t)t3 * coef[k - j - 1];
                h[j] = wrap16((int32_t)wrap16(prod >> shift) + coef[j]);
            }
#pragma unroll
            for (int j = 0; j < k; j++) coef[j] = h[j];
            limit2--;
        }
        if (k == 2) break;
        for (uint32_t base = 0; base < limit2; base += 32) {
            const uint32_t j = base + lane;
            const bool on = j < limit2;
            int32_t pef_j = 0, pef_i = 0, per_j = 0, per_i = 0, x_ik = 0, x_i = 0;
            if (on) {
                pef_j = ws.pef[j]; pef_i = ws.pef[j + 1]; per_j = ws.per[j]; per_i = ws.per[j + 1];
                x_ik = ws.x[j + 1 + k]; x_i = ws.x[j + 1];
            }
            __syncwarp();
            if (on) {
                const int32_t p = (int32_t)((uint32_t)(pef_j + x_ik) * (uint32_t)(int32_t)t3);
                ws.per[j] = wrap16(per_j + wrap16(p >> shift));
                const int32_t f = (int32_t)((uint32_t)(per_i + x_i) * (uint32_t)(int32_t)t3);
                ws.pef[j] = wrap16((int32_t)wrap16(f >> shift) + pef_i);
            }
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(kWarpsPerCta * 32) flywheel_kernel(const FlywheelParams p)
{
    __shared__ WarpScratch scratch[kWarpsPerCta];
    const uint32_t lane = threadIdx.x & 31;
    WarpScratch& ws = scratch[threadIdx.x >> 5];
    const uint64_t q = (uint64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (q >= p.n) return;
    const uint4* jp = reinterpret_cast<const uint4*>(p.jobs + q);
    const uint4 j0 = __ldg(jp), j1 = __ldg(jp + 1);
    ohp_flywheel_job job;
    job.src_off = (uint64_t)j0.x | ((uint64_t)j0.y << 32);
    job.dst_off = (uint64_t)j0.z | ((uint64_t)j0.w << 32);
    job.sample_rate = j1.x;
    job.out_frames = j1.y;
    job.train_frames = (uint16_t)(j1.z & 0xffffu);
    job.channels = (uint8_t)((j1.z >> 16) & 0xffu);
    job.bit_depth = (uint8_t)(j1.z >> 24);
    job.reserved = j1.w;
    const uint32_t err = check_job(job, p.in_bytes, p.out_bytes);
    if (err) {
        if (lane == 0) {
            atomicOr(&p.status[0], err);
            atomicCAS(&p.status[1], 0u, (uint32_t)(q + 1 > 0xffffffffull ? 0xffffffffull : q + 1));
        }
        return;
    }
    const uint32_t ch = job.channels;
    const uint32_t dec = decimation_factor(job.sample_rate);
    const uint32_t jps = core::jiffies_per_sample_or_zero(job.sample_rate);
    const uint32_t count = job.train_frames / dec; // aSamples.Bytes() / (kBytesPerSample * decFactor), FlywheelRamper.cpp:197
    const uint32_t ob = job.bit_depth >> 3;

    // ---- 1. per channel: decimate, Burg, coefficients and initial states into lane c's registers
    int32_t s0 = 0, s1 = 0, s2 = 0, c0 = 0, c1 = 0, c2 = 0;
    for (uint32_t c = 0; c < ch; c++) {
        const uint8_t* plane = p.in + job.src_off + (uint64_t)c * job.train_frames * 4u;
        for (uint32_t i = lane; i < count; i += 32) {
            const uint8_t* s = plane + (uint64_t)i * 4u * dec;
            const int16_t s16 = (int16_t)(uint16_t)(((uint32_t)s[0] << 8) | s[1]);
            ws.x[i] = (int16_t)(s16 >> 1); // kBurgDataDescaleBitCount, FlywheelRamper.cpp:220
            ws.per[i] = 0;
            ws.pef[i] = 0;
        }
        __syncwarp();
        int16_t coef[3];
        burg3(ws, count, lane, coef);
        if (lane == c) {
            // the last kDegree decimated samples at full width, newest first (FlywheelRamper.cpp:214-218)
            auto full = [&](uint32_t i) -> int32_t {
                const uint8_t* s = plane + (uint64_t)i * 4u * dec;
                return (int32_t)(((uint32_t)s[0] << 24) | ((uint32_t)s[1] << 16) | ((uint32_t)s[2] << 8) | s[3]);
            };
            s0 = full(count - 1); s1 = full(count - 2); s2 = full(count - 3);
            // CorrectBurgCoeffs / CoeffOverflow (FlywheelRamper.cpp:347-388) with kOne = 1 << 13
            const int16_t one = (int16_t)(1 << 13);
            const int16_t total = wrap16((int32_t)wrap16((int32_t)coef[0] + coef[1]) + coef[2]);
            if (!(total <= one && total >= -one)) {
                const int16_t excess = (total & 0x8000) ? wrap16((int32_t)total + one) : wrap16((int32_t)total - one);
                coef[0] = wrap16((int32_t)coef[0] - (int32_t)excess * 2);
            }
            // PrepareFeedbackCoeffs (FlywheelRamper.cpp:233-240): invert and widen
            c0 = (int32_t)(0u - ((uint32_t)(int32_t)coef[0] << 16));
            c1 = (int32_t)(0u - ((uint32_t)(int32_t)coef[1] << 16));
            c2 = (int32_t)(0u - ((uint32_t)(int32_t)coef[2] << 16));
        }
        __syncwarp();
    }

    // ---- 2 + 3. generate in 1 ms blocks, tile by tile
    const uint32_t block_frames = OHP_FLYWHEEL_BLOCK_JIFFIES / jps;
    uint8_t* dst = p.out + job.dst_off;
    uint32_t remaining = job.out_frames;
    int32_t held = 0;
    while (remaining > 0) {
        const uint32_t frames = remaining > block_frames ? block_frames : remaining;
        remaining -= frames;
        uint32_t hold = 0; // RenderChannels restarts the sample hold with every block (FlywheelRamper.cpp:89)
        for (uint32_t f0 = 0; f0 < frames; f0 += kTileFrames) {
            const uint32_t tf = frames - f0 < kTileFrames ? frames - f0 : kTileFrames;
            if (lane < ch) {
                for (uint32_t f = 0; f < tf; f++) {
                    if (hold == 0) {
                        // FeedbackModel::NextSample (FlywheelRamper.cpp:447-487): coefficient format 3, no output shift
                        uint32_t sum = (uint32_t)__mulhi(s0, c0) + (uint32_t)__mulhi(s1, c1) + (uint32_t)__mulhi(s2, c2);
                        s2 = s1;
                        s1 = s0;
                        sum <<= 3;
                        s0 = (int32_t)sum;
                        held = (int32_t)sum;
                    }
                    ws.tile[f * ch + lane] = held;
                    if (++hold == dec) hold = 0;
                }
            }
            else {
                // keep the idle lanes' hold counter in step (it is block-uniform state)
                hold = (hold + tf) % dec;
            }
            __syncwarp();
            const uint32_t nsub = tf * ch;
            const uint32_t nbytes = nsub * ob;
            auto byte_at = [&](uint32_t i) -> uint32_t {
                const uint32_t sub = i / ob, b = i - sub * ob;
                if (b == 3u) return 0u; // 32-bit: the least significant byte is dropped (StarvationRamper.cpp:313-321)
                return ((uint32_t)ws.tile[sub] >> (24u - 8u * b)) & 0xffu;
            };
            if ((reinterpret_cast<uint64_t>(dst) & 3u) == 0) {
                const uint32_t words = nbytes >> 2;
                uint32_t* d4 = reinterpret_cast<uint32_t*>(dst);
                for (uint32_t w = lane; w < words; w += 32) {
                    const uint32_t i = 4u * w;
                    d4[w] = byte_at(i) | (byte_at(i + 1) << 8) | (byte_at(i + 2) << 16) | (byte_at(i + 3) << 24);
                }
                for (uint32_t i = 4u * words + lane; i < nbytes; i += 32) dst[i] = (uint8_t)byte_at(i);
            }
            else {
                for (uint32_t i = lane; i < nbytes; i += 32) dst[i] = (uint8_t)byte_at(i);
            }
            dst += nbytes;
            __syncwarp();
        }
    }
}

} // namespace fly
} // namespace ohp
