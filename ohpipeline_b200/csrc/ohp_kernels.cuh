// ohp_kernels.cuh -- device code of the fused ramp + format-convert path (sm_100a).
//
// One MsgPlayable ("chunk") is the unit of work.  Per chunk a CTA
//   1. fetches the 32-byte descriptor (two 128-bit loads, broadcast to the CTA),
//   2. stages the chunk's source bytes into shared memory with coalesced 128-bit loads of the
//      16-byte-aligned span that covers it (chunks start at arbitrary byte offsets: a 24-bit stereo
//      frame is 6 bytes, a split playable starts wherever the ramp ended),
//   3. transforms "units" of four subsamples held in registers: unpack (BE or LE wire order,
//      DecodedAudio::CopyToBigEndian*, Msg.cpp:380-408), attenuate (MsgPlayablePcm::ApplyAttenuation,
//      Msg.cpp:2736-2751), ramp (RampApplicator::GetNextSample, Msg.cpp:832-899) and repack for the
//      IPcmProcessor sink (packed BE / packed LE) -- byte shuffles are PRMT, the ramp is one IMAD,
//   4. writes the result with 128-bit stores to the 16-byte-aligned span of the destination and
//      byte stores for the ragged head/tail.
// The 512-entry ramp curve sits in shared memory (as 2*multiplier so that the Q15 product's high half
// is the answer), and the per-frame ramp position trunc(i*total/(N-1)) is an exact multiply-high by a
// per-chunk magic reciprocal instead of the reference's per-frame integer divide.
//
// Everything is integer; results are bit-exact against the reference (see oracle/).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/ohp_b200.h"

namespace ohp {

constexpr int kThreads = 128;                 // threads per CTA
constexpr uint32_t kMaxChunk = OHP_MAX_PCM_CHUNK_BYTES;
constexpr uint32_t kSinBytes = kMaxChunk + 48;  // aligned span (<= 9216+30 rounded) + one pad word for funnel reads
constexpr uint32_t kSoutBytes = kMaxChunk + 32; // transformed image at offset 0 + pad for the last partial unit / funnel

// device status word bits (OR-ed by the kernel, read back by ohp_sync)
constexpr uint32_t kErrInvalidDesc = 1u;
constexpr uint32_t kErrOutOfRange = 2u;

struct KernelParams
{
    const ohp_chunk_desc* descs;
    uint64_t n;
    const uint8_t* in;
    uint64_t in_bytes;
    uint8_t* out;
    uint64_t out_bytes;
    const uint16_t* table2;  // 512 x (2 * kRampArray[i])
    uint32_t* status;        // [0] error bits, [1] index of first offending chunk + 1
};

// ---------------------------------------------------------------------------------------------
// small helpers

__device__ __forceinline__ uint4 ldg128_stream(const void* p)
{
    // streaming data, read once: do not allocate in L1
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ void stg128_stream(void* p, const uint4& v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// floor(k / ch) for k < 2^27, ch in 1..32: multiply-high by ceil(2^32/ch) (exact: error < ch * k / 2^32 < 1)
__device__ __forceinline__ uint32_t div_channels(uint32_t k, uint32_t ch_magic)
{
    return ch_magic == 0 ? k : __umulhi(k, ch_magic);
}

// Per-chunk ramp constants, computed once by one thread (Msg.cpp:820-837 restated).
struct RampConst
{
    uint32_t start;     // Ramp::Start()
    uint32_t total;     // |Start - End|
    uint32_t up;        // End > Start: ramp value rises with the frame index
    uint32_t magic;     // ceil(2^(32+shift) / (N-1)); q = umulhi(i*total, magic) >> shift == (i*total)/(N-1) exactly
    uint32_t shift;
    uint32_t div_by_one; // N-1 == 1: q = i*total
};

__device__ __forceinline__ RampConst make_ramp_const(uint32_t start, uint32_t end, uint32_t frames)
{
    RampConst rc;
    rc.start = start;
    rc.up = end > start;
    rc.total = rc.up ? end - start : start - end;
    const uint32_t d = frames > 1 ? frames - 1 : 0;
    rc.div_by_one = (d == 1);
    rc.magic = 0;
    rc.shift = 0;
    if (d > 1) {
        // x = i*total < 9216*16384 < 2^28.  With L = ceil(log2 d) and p = 31 + L:
        //   magic = ceil(2^p / d) < 2^32 and magic*d - 2^p < d <= 2^L, so the error term x*2^L/2^p < 2^(28-31) < 1.
        const uint32_t L = 32 - __clz(d - 1);
        const uint32_t p = 31 + L;
        rc.magic = (uint32_t)((((uint64_t)1 << p) + d - 1) / d);
        rc.shift = p - 32;
    }
    return rc;
}

// 2 * kRampArray[rampIndex] for frame i (RampApplicator::GetNextSample, Msg.cpp:835-837, 864)
__device__ __forceinline__ uint32_t ramp_mult2(const RampConst& rc, const uint16_t* s_table2, uint32_t frame)
{
    const uint32_t x = frame * rc.total;
    const uint32_t q = rc.div_by_one ? x : (__umulhi(x, rc.magic) >> rc.shift);
    const uint32_t ramp = rc.up ? rc.start + q : rc.start - q;
    // (kFullRampSpan - ramp + 16) >> 5, clamped to the last table entry.  Frames past the end of a chunk
    // (tail of the last unit) can take ramp out of range; the unsigned wrap lands on the clamp.
    const uint32_t idx = min(511u, (OHP_RAMP_MAX + 16u - ramp) >> 5);
    return s_table2[idx];
}

// ---------------------------------------------------------------------------------------------
// unit transform: four subsamples of B bytes each (4*B bytes = B words) in registers

// Left-justified big-endian value of subsample j (low 4-B bytes are don't-care).
template <int B>
__device__ __forceinline__ uint32_t unit_extract(const uint32_t (&r)[B + 1], int j, bool le)
{
    const int o = j * B;
    const int wi = o >> 2;
    const int oo = o & 3;
    uint32_t sel_be = 0, sel_le = 0;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const int ibe = (t < B) ? oo + t : oo;
        const int ile = (t < B) ? oo + (B - 1 - t) : oo;
        sel_be |= (uint32_t)ibe << (4 * (3 - t));
        sel_le |= (uint32_t)ile << (4 * (3 - t));
    }
    return __byte_perm(r[wi], r[wi + 1], le ? sel_le : sel_be);
}

// Pack four left-justified results into B output words in BE or LE subsample byte order.
template <int B>
__device__ __forceinline__ void unit_pack(const uint32_t (&o)[4], uint32_t (&w)[B], bool le)
{
    if constexpr (B == 1) {
        // top byte of each result
        const uint32_t lo = __byte_perm(o[0], o[1], 0x0073);
        const uint32_t hi = __byte_perm(o[2], o[3], 0x0073);
        w[0] = __byte_perm(lo, hi, 0x5410);
    } else {
#pragma unroll
    for (int m = 0; m < B; m++) {
        const int j_lo = (4 * m) / B;
        uint32_t sel_be = 0, sel_le = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int q = 4 * m + i;
            const int j = q / B;
            const int t = q % B;
            const int src = (j == j_lo) ? 0 : 4;          // first or second PRMT operand
            const int bbe = 3 - t;                        // MSB first
            const int ble = 3 - (B - 1 - t);              // LSB first
            sel_be |= (uint32_t)(src + bbe) << (4 * i);
            sel_le |= (uint32_t)(src + ble) << (4 * i);
        }
        const int j_hi = (4 * m + 3) / B;
        w[m] = __byte_perm(o[j_lo], o[j_hi], le ? sel_le : sel_be);
    }
    }
}

struct ChunkCtx
{
    uint32_t bytes;        // payload bytes (== output bytes for packed sinks)
    uint32_t channels;
    uint32_t ch_magic;     // ceil(2^32/ch), 0 for mono
    uint32_t attenuation;  // 256 = unity
    bool ramped;
    bool in_le;
    bool out_le;
    bool tag6;             // 6-channel 32-bit: channel id in the low byte of ramped subsamples
};

// Transform the whole chunk from the staged source image (s_in + head) into s_out (offset 0).
template <int B>
__device__ __forceinline__ void transform_chunk(const ChunkCtx& cx, const RampConst& rc, const uint16_t* s_table2,
                                                const uint8_t* s_in, uint32_t head, uint8_t* s_out)
{
    const uint32_t subsamples = cx.bytes / B;
    const uint32_t units = (subsamples + 3) >> 2;
    const uint32_t* in_w = reinterpret_cast<const uint32_t*>(s_in) + (head >> 2);
    const uint32_t fshift = (head & 3) * 8;
    uint32_t* out_w = reinterpret_cast<uint32_t*>(s_out);
    for (uint32_t u = threadIdx.x; u < units; u += kThreads) {
        // B source words, realigned when the chunk starts off a word boundary
        uint32_t raw[B + 1];
#pragma unroll
        for (int i = 0; i <= B; i++) raw[i] = in_w[u * B + i];
        uint32_t r[B + 1];
#pragma unroll
        for (int i = 0; i < B; i++) r[i] = __funnelshift_r(raw[i], raw[i + 1], fshift);
        r[B] = 0;

        uint32_t o[4];
        uint32_t frame = 0, chan = 0, mult2 = 0;
        if (cx.ramped) {
            const uint32_t k0 = 4 * u;
            frame = div_channels(k0, cx.ch_magic);
            chan = k0 - frame * cx.channels;
            mult2 = ramp_mult2(rc, s_table2, frame);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t v = unit_extract<B>(r, j, cx.in_le);
            if (B == 2 && cx.attenuation != OHP_UNITY_ATTENUATION) {
                // ((TInt)sample) * iAttenuation / 256 in UNSIGNED 32-bit arithmetic, truncated to 16 bits (Msg.cpp:2746)
                const uint32_t s = (uint32_t)((int32_t)v >> 16);
                v = ((s * cx.attenuation) >> 8) << 16;
            }
            if (cx.ramped) {
                if (j > 0) {
                    chan++;
                    if (chan == cx.channels) {
                        chan = 0;
                        frame++;
                        mult2 = ramp_mult2(rc, s_table2, frame);
                    }
                }
                // subsample16 from the two most significant bytes (8-bit: byte << 8), Msg.cpp:840-862
                const int32_t s16 = (B == 1) ? (((int32_t)v >> 24) << 8) : ((int32_t)v >> 16);
                // (s16 * mult) >> 15 kept to 16 bits == high half of s16 * (2*mult); low bytes are zero (Msg.cpp:865-895)
                const uint32_t prod2 = (uint32_t)(s16 * (int32_t)mult2);
                v = prod2 & (B == 1 ? 0xFF000000u : 0xFFFF0000u);
                if (B == 4 && cx.tag6) v |= chan << 4;
            }
            o[j] = v;
        }
        uint32_t w[B];
        unit_pack<B>(o, w, cx.out_le);
#pragma unroll
        for (int i = 0; i < B; i++) out_w[u * B + i] = w[i];
    }
}

// ---------------------------------------------------------------------------------------------
// staging in / out

// Load the 16-byte-aligned span covering [src, src+bytes) into s_in; returns src & 15.
__device__ __forceinline__ uint32_t stage_in(const uint8_t* in_base, uint64_t in_bytes, uint64_t src_off, uint32_t bytes,
                                             uint8_t* s_in)
{
    const uint64_t addr = reinterpret_cast<uint64_t>(in_base) + src_off;
    const uint32_t head = (uint32_t)(addr & 15u);
    const uint8_t* al = reinterpret_cast<const uint8_t*>(addr - head);
    const uint32_t words = (head + bytes + 15u) >> 4;
    const uint8_t* in_end = in_base + in_bytes;
    uint4* s4 = reinterpret_cast<uint4*>(s_in);
    for (uint32_t w = threadIdx.x; w < words; w += kThreads) {
        const uint8_t* p = al + 16u * w;
        if (p + 16 <= in_end) {
            s4[w] = ldg128_stream(p);
        } else {
            // last word of the arena: touch only bytes that exist
            for (uint32_t i = 0; i < 16; i++) s_in[16u * w + i] = (p + i < in_end) ? p[i] : 0;
        }
    }
    return head;
}

// Store `bytes` bytes found at s_src + s_off to the (arbitrarily aligned) global address dst.
__device__ __forceinline__ void stage_out(const uint8_t* s_src, uint32_t s_off, uint8_t* dst, uint32_t bytes)
{
    const uint32_t lead = (uint32_t)((16u - (reinterpret_cast<uint64_t>(dst) & 15u)) & 15u);
    const uint32_t head_n = min(lead, bytes);
    const uint32_t words = (bytes - head_n) >> 4;
    const uint32_t tail_at = head_n + (words << 4);
    // ragged head and tail: at most 15 bytes each
    if (threadIdx.x < head_n) dst[threadIdx.x] = s_src[s_off + threadIdx.x];
    if (threadIdx.x >= 32 && threadIdx.x - 32 < bytes - tail_at) {
        const uint32_t i = tail_at + threadIdx.x - 32;
        dst[i] = s_src[s_off + i];
    }
    const uint32_t base = s_off + head_n;          // byte offset in s_src of the first aligned store
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(s_src) + (base >> 2);
    const uint32_t fshift = (base & 3) * 8;
    uint4* d4 = reinterpret_cast<uint4*>(dst + head_n);
    if (fshift == 0) {
        for (uint32_t w = threadIdx.x; w < words; w += kThreads) {
            uint4 v;
            v.x = sw[4 * w + 0]; v.y = sw[4 * w + 1]; v.z = sw[4 * w + 2]; v.w = sw[4 * w + 3];
            stg128_stream(d4 + w, v);
        }
    } else {
        for (uint32_t w = threadIdx.x; w < words; w += kThreads) {
            const uint32_t a0 = sw[4 * w + 0], a1 = sw[4 * w + 1], a2 = sw[4 * w + 2], a3 = sw[4 * w + 3], a4 = sw[4 * w + 4];
            uint4 v;
            v.x = __funnelshift_r(a0, a1, fshift);
            v.y = __funnelshift_r(a1, a2, fshift);
            v.z = __funnelshift_r(a2, a3, fshift);
            v.w = __funnelshift_r(a3, a4, fshift);
            stg128_stream(d4 + w, v);
        }
    }
}

// MsgPlayableSilence::ReadBlock (Msg.cpp:2874-2893): zeros; with 6 channels every emitted block of
// maxBytes starts with 00 00 00 c0 for c0 = 0x00,0x10..0x70 (32 bytes, whatever the bit depth).
__device__ __forceinline__ void write_silence(uint8_t* dst, uint32_t bytes, uint32_t channels, uint32_t B)
{
    const uint32_t block = kMaxChunk - (kMaxChunk % (channels * B));
    const uint32_t lead = (uint32_t)((16u - (reinterpret_cast<uint64_t>(dst) & 15u)) & 15u);
    const uint32_t head_n = min(lead, bytes);
    const uint32_t words = (bytes - head_n) >> 4;
    const uint32_t tail_at = head_n + (words << 4);
    auto value_at = [&](uint32_t i) -> uint32_t {
        if (channels != 6) return 0u;
        const uint32_t r = i % block;
        return (r < 32u && (r & 3u) == 3u) ? ((r >> 2) << 4) : 0u;
    };
    if (threadIdx.x < head_n) dst[threadIdx.x] = (uint8_t)value_at(threadIdx.x);
    if (threadIdx.x >= 32 && threadIdx.x - 32 < bytes - tail_at) {
        const uint32_t i = tail_at + threadIdx.x - 32;
        dst[i] = (uint8_t)value_at(i);
    }
    uint4* d4 = reinterpret_cast<uint4*>(dst + head_n);
    for (uint32_t w = threadIdx.x; w < words; w += kThreads) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (channels == 6) {
            const uint32_t i0 = head_n + 16u * w;
            if ((i0 % block) < 32u || (i0 % block) + 16u > block) {
                uint32_t t[4] = {0, 0, 0, 0};
                for (uint32_t i = 0; i < 16; i++) t[i >> 2] |= value_at(i0 + i) << (8 * (i & 3));
                v = make_uint4(t[0], t[1], t[2], t[3]);
            }
        }
        stg128_stream(d4 + w, v);
    }
}

} // namespace ohp
