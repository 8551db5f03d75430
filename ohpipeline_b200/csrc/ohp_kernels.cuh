// ohp_kernels.cuh -- device code of the fused ramp + format-convert path (sm_100a).
//
// One MsgPlayable ("chunk") is the unit of work.  The kernel is persistent and warp-specialised; two CTAs per SM,
// each with a 103 KB shared-memory BYTE RING, one loader warp and eight consumer warps:
//
//   loader  (warp 0)    walks this CTA's chunks: its 32 lanes fetch and decode 32 descriptors at a time (restating the
//                       reference's ASSERTs and precomputing the per-chunk ramp constants; the next batch's descriptors
//                       are already in flight), then the whole warp runs the issue loop in lockstep, everything it
//                       needs of chunk j arriving by shuffle from lane j:  (1) lanes test the "empty" barriers of all
//                       slots in flight at once and reclaim, in order, what has been released;  (2) as many of the next
//                       chunks as fit are given contiguous slots in the ring -- all at once, by a prefix sum of their
//                       sizes over the lanes (SERIAL_PLACE = false), or one after the other (SERIAL_PLACE = true: fewer
//                       dependent shuffles when the ring is full and only one to three fit);  (3) one lane per placed
//                       chunk starts ONE TMA bulk copy (cp.async.bulk.shared::cluster.global, completion on the slot's
//                       "full" mbarrier) of the 16-byte-aligned span covering the chunk's source bytes.  With large
//                       chunks the ring is full and the loader places one chunk per released slot; with small chunks
//                       the loader is what limits, and it places up to OHP_ISSUE_WIDTH per round;
//   consumers           take chunks by TICKET (a shared-memory counter), first come first served, so a warp that met
//                       a run of expensive chunks does not hold up the in-order ring while its neighbours idle.  The
//                       warp waits on the chunk's "full" mbarrier, transforms the chunk IN PLACE in "units" of four
//                       subsamples held in registers -- unpack (BE or LE wire order: DecodedAudio::CopyToBigEndian*,
//                       Msg.cpp:380-408), attenuate (MsgPlayablePcm::ApplyAttenuation, Msg.cpp:2736-2751), ramp
//                       (RampApplicator::GetNextSample, Msg.cpp:832-899), repack for the IPcmProcessor sink; byte
//                       shuffles are PRMT, the ramp one IMAD per subsample -- then writes it out itself: the
//                       16-byte-aligned interior of the destination with ONE TMA bulk store
//                       (cp.async.bulk.global.shared::cta), the ragged head/tail (chunks start at arbitrary byte
//                       offsets: a 24-bit stereo frame is 6 bytes, a split playable starts wherever the ramp ended)
//                       with byte stores, and hands the slot back ("empty" mbarrier) once the store has read it.
//
// In place means a chunk in flight costs one buffer, not two, so the 227 KB of an SM hold twice as many chunks between
// "load issued" and "store drained" -- that, not arithmetic, is what bounds an HBM-bound kernel.  A warp per chunk means
// no cross-warp synchronisation on the data path and one setup per chunk.  (profiles/README.md has the measurements
// behind the shape: CTAs per SM, ring size, warps, issue width, polling discipline.)
//
// The 512-entry ramp curve sits in shared memory as 2*multiplier, so that the high half of the 16x16 product is
// the reference's (s16 * mult) >> 15; the per-frame ramp position trunc(i*total/(N-1)) is an exact multiply-high by
// a per-chunk magic reciprocal instead of the reference's per-frame integer divide.  Silence chunks
// (MsgPlayableSilence::ReadBlock, Msg.cpp:2874-2893) are written straight to global memory by their consumer warp.
//
// Everything is integer; results are bit-exact against the reference (see oracle/).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/ohp_b200.h"

namespace ohp {

// Tunables (overridable with -D for experiments; the defaults are what ships)
#ifndef OHP_CONSUMER_WARPS
#define OHP_CONSUMER_WARPS 8
#endif
#ifndef OHP_RING_BYTES
#define OHP_RING_BYTES 105472
#endif
#ifndef OHP_RING_SLOTS
#define OHP_RING_SLOTS 32
#endif
#ifndef OHP_CHUNK_BLOCK
#define OHP_CHUNK_BLOCK 16
#endif
#ifndef OHP_LOADER
#define OHP_LOADER 1        /* 1: place and start up to OHP_ISSUE_WIDTH chunks per round; 0: one chunk at a time */
#endif
#ifndef OHP_CONSUMER_POLL
#define OHP_CONSUMER_POLL 0 /* 0: every lane waits on the barrier; 1: lane 0 waits, __syncwarp; 2: one look by all, then 1 */
#endif
#ifndef OHP_ISSUE_WIDTH
#define OHP_ISSUE_WIDTH 16
#endif
#ifndef OHP_GROUPS_PER_STEP
#define OHP_GROUPS_PER_STEP 1
#endif
#ifndef OHP_DEFER_RELEASE
#define OHP_DEFER_RELEASE 2 /* 0: never, 1: always, 2: serial-placement instantiation only (see the consumer loop) */
#endif
#ifndef OHP_ANY_UNITS
#define OHP_ANY_UNITS 1     /* 4-subsample units a lane of the compact general transform works on at once; 2 measured
                               WORSE (configs[3] 0.905 -> 0.87, stress mix 0.40 -> 0.38: +7 KB of code in four functions
                               that are hot together costs more in instruction fetch than the second chain gains) */
#endif
#ifndef OHP_DYNAMIC
#define OHP_DYNAMIC 1       /* 1: consumer warps take chunks by ticket (first come, first served); 0: chunk k -> warp k % warps */
#endif
constexpr uint32_t kIssueWidth = OHP_ISSUE_WIDTH;     // chunks the loader warp can place and start in one round (one per lane)
constexpr uint32_t kSerialWidth = 8;                  // ... and at most this many per round in serial-placement mode
constexpr int kAnyUnits = OHP_ANY_UNITS;
constexpr int kGroupsPerStep = OHP_GROUPS_PER_STEP; // independent 16-subsample groups a lane works on at once
constexpr uint32_t kChunkBlock = OHP_CHUNK_BLOCK;     // chunks are dealt to CTAs in runs of (up to) this many consecutive chunks
constexpr uint32_t kMinChunkBlock = 4;                // ... shortened for small batches so that every CTA gets work
constexpr int kConsumerWarps = OHP_CONSUMER_WARPS;
constexpr int kThreads = 32 + kConsumerWarps * 32; // loader warp + consumer warps
constexpr int kRecSlots = 64;                      // two batches of 32 decoded chunk records
constexpr uint32_t kRingSlots = OHP_RING_SLOTS;    // chunks in flight per CTA (barrier pairs); <= 32
constexpr uint32_t kRingBytes = OHP_RING_BYTES;    // shared-memory byte ring the chunk slots are carved from
constexpr uint32_t kMaxChunk = OHP_MAX_PCM_CHUNK_BYTES;
constexpr uint32_t kSlotFront = 0;                 // (the output image never starts below the input image any more)
constexpr uint32_t kSlotBack = 80;                 // over-read / over-write of the last 16-subsample group + funnel word
static_assert(kRingBytes % 16 == 0 && kRingBytes >= 2 * (kSlotFront + kMaxChunk + 16 + kSlotBack), "ring too small");
static_assert(kRingSlots <= 32, "ring slots: at most one decode batch");
// A barrier pair is waited on by parity, which only tells "the previous phase is over" from "it is not": whoever waits
// for phase q of a pair must know that phase q - 1 has completed.
//   * static assignment (chunk k -> warp k % warps, pair k % kRingSlots): every phase of one pair must be consumed by
//     the SAME warp (it finishes phase q - 1 before it waits for q), so the slot count is a multiple of the warp count;
//   * ticket assignment (OHP_DYNAMIC): any warp may hold any ticket, and a warp can be handed a ticket whose chunk has
//     not even been started.  With S = kRingSlots chunks in flight at most and W <= S warps, a ticket y is never more
//     than S + W ahead of the oldest chunk o still loading (finished and running tickets are loaded, hence < o + S;
//     at most W tickets are blocked), so with 2 S pairs the previous user of y's pair, chunk y - 2 S < o, has landed.
#if OHP_DYNAMIC
constexpr uint32_t kBarPairs = 2u * kRingSlots;
static_assert((uint32_t)kConsumerWarps <= kRingSlots, "ticket assignment needs at most one warp per ring slot");
#else
constexpr uint32_t kBarPairs = kRingSlots;
static_assert(kRingSlots % kConsumerWarps == 0, "ring slots must be a multiple of the consumer warp count");
#endif

// device status word bits (OR-ed by the kernel, read back by ohp_sync)
constexpr uint32_t kErrInvalidDesc = 1u;
constexpr uint32_t kErrOutOfRange = 2u;
constexpr uint32_t kErrWatchdog = 4u;

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
    uint32_t cap_bytes;      // ring bytes a CTA may have in flight (<= kRingBytes)
    uint32_t cap_chunks;     // chunks a CTA may have in flight (<= kRingSlots)
    uint32_t chunk_block;    // chunks are dealt to the CTAs in runs of this many consecutive chunks
    uint32_t serial_place;   // host side only: which instantiation runs (1 = loader places chunks one after the other, up to
                             // kSerialWidth a round; 0 = all at once by prefix sum, up to kIssueWidth)
};

// One descriptor, unpacked; the checks are the reference's ASSERTs restated, shared by ohp_validate (host) and the
// loader warp (device) so that both reject exactly the same descriptors.
struct DescFields
{
    uint64_t src_off, dst_off;
    uint32_t bytes, ramp_start, ramp_end, attenuation, bit_depth, channels, flags, out_fmt, aux;
};
struct DescDerived
{
    uint32_t frames;
    uint32_t out_bytes;   // bytes the sink receives
    uint64_t out_extent;  // distance from dst_off to one past the last byte written (planar output is strided)
};
// 0 = ok, else kErrInvalidDesc / kErrOutOfRange
__host__ __device__ inline uint32_t check_desc_fields(const DescFields& d, uint64_t in_bytes, uint64_t out_total, DescDerived& o)
{
    o.frames = 0; o.out_bytes = 0; o.out_extent = 0;
    // an all-zero descriptor is an EMPTY SLOT (ohp_run_streams_device pads each stream's region of the descriptor array with
    // them): nothing to read, nothing to write, nothing to reject
    if ((d.bytes | d.ramp_start | d.ramp_end | d.attenuation | d.bit_depth | d.channels | d.flags | d.out_fmt | d.aux) == 0) return 0;
    const uint32_t bd = d.bit_depth;
    if (!(bd == 8 || bd == 16 || bd == 24 || bd == 32)) return 1u;           // ConstructPcm ASSERTs, Msg.cpp:349-366
    if (d.channels < 1 || d.channels > 32) return 1u;
    if (d.ramp_start > OHP_RAMP_MAX || d.ramp_end > OHP_RAMP_MAX) return 1u;  // Ramp::DoValidate, Msg.cpp:747-752
    if (d.out_fmt > OHP_OUT_SONGCAST) return 1u;
    const uint32_t B = bd >> 3;
    const bool silence = (d.flags & OHP_F_SILENCE) != 0;
    const uint32_t frame_bytes = B * d.channels;
    const uint32_t frames = d.bytes / frame_bytes;
    if (frames * frame_bytes != d.bytes) return 1u;                           // whole frames (ProcessorAudioUtils.cpp:45)
    if (!silence && d.bytes > OHP_MAX_PCM_CHUNK_BYTES) return 1u;             // a DecodedAudio cell
    if (!silence && d.attenuation != OHP_UNITY_ATTENUATION && bd != 16) return 1u; // Msg.cpp:2741
    o.frames = frames;
    o.out_bytes = d.bytes;
    o.out_extent = d.bytes;
    switch (d.out_fmt) {
    case OHP_OUT_PACKED_BE:
        break;
    case OHP_OUT_PACKED_LE:                                                   // TestCodecInteractiveMain.cpp:546-567
        if (silence || B > 3 || d.aux > OHP_LE_APPEND) return 1u;
        if (d.aux != OHP_LE_APPEND && (d.flags & OHP_F_RAMP_ENABLED) && B >= 2 && frames != 0) {
            // SwapEndianness16/24 overwrite (TestCodecInteractiveMain.cpp:570-590): the sink is left holding the last of the
            // fragments MsgPlayablePcm::ReadBlock made, 256 / frame bytes frames each (Msg.cpp:2765-2779)
            const uint32_t per_fragment = 256u / frame_bytes;
            const uint32_t last = frames - ((frames - 1u) / per_fragment) * per_fragment;
            o.out_bytes = last * frame_bytes;
            o.out_extent = o.out_bytes;
        }
        break;
    case OHP_OUT_PLANAR32_BE:                                                 // StarvationRamper.cpp:90-111
        if (d.aux < frames) return 1u;
        if (silence && d.bytes > OHP_MAX_PCM_CHUNK_BYTES) return 1u;
        o.out_bytes = frames * d.channels * 4u;
        o.out_extent = frames ? (uint64_t)(d.channels - 1u) * d.aux * 4u + (uint64_t)frames * 4u : 0;
        break;
    case OHP_OUT_FROM32_BE:                                                   // StarvationRamper.cpp:281-343
        if (silence || B != 4) return 1u;
        if (!(d.aux == 8 || d.aux == 16 || d.aux == 24 || d.aux == 32)) return 1u;
        o.out_bytes = (d.bytes / 4u) * (d.aux / 8u);
        o.out_extent = o.out_bytes;
        break;
    default: {                                                                // OHP_OUT_SONGCAST, Sender.cpp:356-377
        const uint32_t och = d.channels < 2 ? d.channels : 2u;
        if (d.aux + och > d.channels) return 1u;
        if (silence && d.bytes > OHP_MAX_PCM_CHUNK_BYTES) return 1u;
        o.out_bytes = frames * och * (B < 3 ? B : 3u);
        o.out_extent = o.out_bytes;
        break;
    }
    }
    if (d.dst_off > out_total || o.out_extent > out_total - d.dst_off) return 2u;
    if (!silence && (d.src_off > in_bytes || d.bytes > in_bytes - d.src_off)) return 2u;
    return 0;
}

// What the loader hands to the consumers (and the storer) for one chunk: one 64-byte shared-memory record.
enum ChunkKind : uint32_t { kSkip = 0, kPcm = 1, kSilence = 2, kSilenceConv = 3 /* silence into a converting sink */ };

struct ChunkRec
{
    uint32_t kind;
    uint32_t bytes;       // payload bytes (== output bytes for the packed sinks)
    uint32_t head;        // src & 15: where the chunk starts inside the staged span
    uint32_t mode;        // kMode* bits
    uint32_t channels;
    uint32_t ch_magic;    // ceil(2^32 / channels); 0 for mono
    uint32_t attenuation;
    uint32_t variant;     // consumer dispatch: (B-1) | chm << 2 | aligned << 4   (see kVar*)
    // ramp: table index for frame i is min(511, (ramp_c + ramp_sign * q(i)) >> 5), q(i) = (i * total) / (N - 1)
    uint32_t ramp_c;      // 16384 + 16 - Start
    int32_t  ramp_sign;   // +1 ramp down, -1 ramp up
    uint32_t ramp_total;  // |Start - End|   (doubled when N - 1 == 1, see make_ramp_const)
    uint32_t ramp_magic;  // ceil(2^(32+shift) / (N-1)) ...
    uint32_t ramp_shift;  // ... so that q = umulhi(i*total, magic) >> shift exactly
    uint32_t units;       // ceil(subsamples / 4)
    uint32_t dst_lo, dst_hi;
    uint32_t frames;
    uint32_t out_fmt;     // ohp_out_fmt
    uint32_t aux;         // per-format parameter (ohp_chunk_desc::aux)
    uint32_t out_bytes;   // bytes the sink receives (== bytes for the packed sinks)
};
static_assert(sizeof(ChunkRec) == 80, "ChunkRec is 80 bytes (16-byte multiple)");

constexpr uint32_t kModeRamped = 1u, kModeInLe = 2u, kModeOutLe = 4u, kModeTag6 = 8u, kModeTransform = 16u;
// how the four subsamples of a unit map to frames
constexpr uint32_t kChmOther = 0u, kChmMono = 1u, kChmStereo = 2u, kChmMul4 = 3u;
constexpr uint32_t kChmAny = 4u; // decided at run time (mono or the general two-frame case): ONE instantiation for every channel count

struct __align__(128) SharedStorage
{
    uint8_t ring[kRingBytes];
    ChunkRec rec[kRecSlots];          // indexed by (chunk ordinal & (kRecSlots-1)); decoded 32 at a time
    uint32_t ring_off[kRecSlots];     // where the chunk's slot starts in the ring (written when the load is issued)
    uint16_t table2[OHP_RAMP_TABLE_ENTRIES];
    uint64_t full[kBarPairs];
    uint64_t empty[kBarPairs];
    unsigned long long next_ticket;   // OHP_DYNAMIC: ordinal of the next chunk a consumer warp will take
};

// Work distribution: chunks are dealt to the CTAs block-cyclically, kChunkBlock consecutive chunks at a time, so that
// the chunks one CTA has in flight are neighbours in HBM (its loads and stores walk DRAM pages instead of hopping
// grid * chunk bytes apart) while the grid as a whole still sweeps one narrow window of the arenas.
__device__ __forceinline__ uint64_t cta_chunk_count(uint64_t n, uint32_t cta, uint32_t grid, uint32_t block)
{
    const uint64_t round = (uint64_t)grid * block;
    const uint64_t full = n / round;
    const uint64_t rem = n - full * round;
    const uint64_t mine = (uint64_t)cta * block;
    const uint64_t extra = rem > mine ? (rem - mine < block ? rem - mine : block) : 0;
    return full * block + extra;
}
// chunk index of this CTA's k-th chunk
__device__ __forceinline__ uint64_t cta_chunk_index(uint64_t k, uint32_t cta, uint32_t grid, uint32_t block)
{
    return ((k / block) * grid + cta) * block + (k % block);
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers (mbarrier, TMA bulk copies, shared-memory accesses by 32-bit address)

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Wait with a watchdog: a protocol bug must surface as an error, never as a hung GPU.
__device__ __forceinline__ long long mbar_wait(uint32_t bar, uint32_t parity, uint32_t* status)
{
#ifdef OHP_PROFILE_WAITS
    const long long t0 = clock64(); // try_wait itself suspends the thread, so time the first call too
#else
    if (mbar_try_wait(bar, parity)) return 0;
    const long long t0 = clock64();
#endif
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) { // ~2 s at 1.9 GHz: far beyond any legitimate wait
            atomicOr(&status[0], kErrWatchdog);
            __trap();
        }
    }
    return clock64() - t0; // cycles spent blocked (used by the OHP_PROFILE_WAITS instrumentation only)
}
// L2 eviction policy for data that is touched exactly once (both arenas stream through)
__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// global -> shared bulk copy (TMA), completion counted in bytes on an mbarrier.  16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void tma_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar)
{
#ifdef OHP_L2_HINT_LOAD
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "l"(l2_evict_first_policy()) : "memory");
#else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
#endif
}
// shared -> global bulk copy (TMA), tracked by bulk async-groups.
__device__ __forceinline__ void tma_store(void* dst, uint32_t src_smem, uint32_t bytes)
{
#ifndef OHP_NO_L2_HINT_STORE /* output is written once and never re-read: evict-first keeps it from crowding L2 (+1.5 % on configs[1]) */
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 :: "l"(dst), "r"(src_smem), "r"(bytes), "l"(l2_evict_first_policy()) : "memory");
#else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(dst), "r"(src_smem), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void tma_commit()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_wait_all()
{
    asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (the TMA store that follows)
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// PRMT with the full selector semantics (bit 3 of a selector nibble replicates the sign of the chosen byte);
// __byte_perm only honours the low three bits of each nibble.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr)
{
    uint32_t v;
    asm volatile("{\n .reg .u16 t;\n ld.shared.u16 t, [%1];\n cvt.u32.u16 %0, t;\n}" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t lds8(uint32_t addr)
{
    uint32_t v;
    asm volatile("{\n .reg .u16 t;\n ld.shared.u8 t, [%1];\n cvt.u32.u16 %0, t;\n}" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ void stg128_stream(void* p, const uint4& v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---------------------------------------------------------------------------------------------
// ramp position -> multiplier

// Per-chunk constants (Msg.cpp:820-837 restated).  Called by the loader.
__device__ __forceinline__ void make_ramp_const(ChunkRec& r, uint32_t start, uint32_t end, uint32_t frames)
{
    const bool up = end > start;
    r.ramp_c = OHP_RAMP_MAX + 16u - start;
    r.ramp_sign = up ? -1 : 1;
    uint32_t total = up ? end - start : start - end;
    uint32_t d = frames > 1 ? frames - 1 : 0;
    if (d == 1) {
        // (i * total) / 1 == (i * 2 total) / 2: keeps the divisor >= 2 so that one formula serves every chunk
        d = 2;
        total *= 2;
    }
    r.ramp_total = total;
    r.ramp_magic = 0;   // N == 1: q = 0, the ramp value is Start() (Msg.cpp:835)
    r.ramp_shift = 0;
    if (d > 1) {
        // x = i*total < 9216*16384 < 2^28.  With L = ceil(log2 d) and p = 31 + L:
        //   magic = ceil(2^p / d) < 2^32 and magic*d - 2^p < d <= 2^L, so the error term x*2^L/2^p < 2^(28-31) < 1.
        const uint32_t L = 32 - __clz(d - 1);
        const uint32_t p = 31 + L;
        r.ramp_magic = (uint32_t)((((uint64_t)1 << p) + d - 1) / d);
        r.ramp_shift = p - 32;
    }
}

struct RampRegs
{
    uint32_t c;
    int32_t sign;
    uint32_t total, magic, shift;
    uint32_t table; // shared-memory address of table2
};

// 2 * kRampArray[rampIndex] for frame i (RampApplicator::GetNextSample, Msg.cpp:835-837, 864).
// Frames past the end of a chunk (tail of the last group) can push the index out of range; the unsigned clamp
// catches both directions and the value is never stored.
__device__ __forceinline__ uint32_t ramp_mult2(const RampRegs& rr, uint32_t frame)
{
    const uint32_t x = frame * rr.total;
    const uint32_t q = __umulhi(x, rr.magic) >> rr.shift;
    const uint32_t idx = min(511u, (rr.c + (uint32_t)(rr.sign * (int32_t)q)) >> 5);
    return lds16(rr.table + 2u * idx);
}

// ---------------------------------------------------------------------------------------------
// unit transform: four subsamples of B bytes each (4*B bytes = B words) in registers

// Subsample j's two most significant bytes as a sign-extended 16-bit value (8-bit: byte << 8): one PRMT with
// sign replication (selector nibble | 8).  r holds the unit's B words (+1 zero word).
template <int B>
__device__ __forceinline__ int32_t unit_s16(const uint32_t (&r)[B + 1], int j, bool le)
{
    const int o = j * B;
    const int wi = o >> 2;
    const int oo = o & 3;
    // byte indices (within the register pair) of the most significant and next byte
    const int msb_be = oo, nxt_be = oo + 1;
    const int msb_le = oo + B - 1, nxt_le = oo + B - 2;
    if (B == 1) {
        // (int8)b << 8: low byte zero -- take it from the always-zero top word r[B]... use a shift instead
        const uint32_t sel = (uint32_t)msb_be | ((uint32_t)msb_be << 4) | ((uint32_t)(msb_be | 8) << 8) | ((uint32_t)(msb_be | 8) << 12);
        return (int32_t)(prmt(r[wi], r[wi + 1], sel) & 0xffffff00u);
    }
    const uint32_t sel_be = (uint32_t)nxt_be | ((uint32_t)msb_be << 4) | ((uint32_t)(msb_be | 8) << 8) | ((uint32_t)(msb_be | 8) << 12);
    const uint32_t sel_le = (uint32_t)nxt_le | ((uint32_t)msb_le << 4) | ((uint32_t)(msb_le | 8) << 8) | ((uint32_t)(msb_le | 8) << 12);
    return (int32_t)prmt(r[wi], r[wi + 1], le ? sel_le : sel_be);
}

// Left-justified big-endian value of subsample j (low 4-B bytes are don't-care) -- the unramped paths.
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

// Pack four left-justified results into B output words in BE or LE subsample byte order.  With `ramped` the bytes
// below the 16-bit result must read as zero (Msg.cpp:868-895): masked per output word after packing.
template <int B>
__device__ __forceinline__ void unit_pack(const uint32_t (&o)[4], uint32_t (&w)[B], bool le, bool ramped)
{
    if constexpr (B == 1) {
        const uint32_t lo = __byte_perm(o[0], o[1], 0x0073);
        const uint32_t hi = __byte_perm(o[2], o[3], 0x0073);
        w[0] = __byte_perm(lo, hi, 0x5410);
    } else {
#pragma unroll
        for (int m = 0; m < B; m++) {
            const int j_lo = (4 * m) / B;
            const int j_hi = (4 * m + 3) / B;
            uint32_t sel_be = 0, sel_le = 0, keep_be = 0, keep_le = 0;
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
                // a ramped result only has its two top bytes (index 3 and 2) populated
                if (bbe >= 2) keep_be |= 0xffu << (8 * i);
                if (ble >= 2) keep_le |= 0xffu << (8 * i);
            }
            uint32_t v = __byte_perm(o[j_lo], o[j_hi], le ? sel_le : sel_be);
            if (B == 3 && ramped) v &= (le ? keep_le : keep_be); // 32-bit results carry their own zero/tag bytes
            w[m] = v;
        }
    }
}

struct UnitCtx
{
    uint32_t channels, ch_magic, attenuation;
    bool ramped, in_le, out_le, tag6;
};

// One unit: four subsamples (B words in r[0..B-1]; r[B] may be anything) -> B output words.
// CHM says how the unit's subsamples map to frames (kChm*); u is the unit's index inside the chunk.
template <int B, uint32_t CHM>
__device__ __forceinline__ void process_unit(const UnitCtx& cx, const RampRegs& rr, uint32_t u, const uint32_t (&r)[B + 1], uint32_t (&w)[B])
{
    uint32_t o[4];
    if (cx.ramped) {
        uint32_t m[4];
        uint32_t chan0 = 0;
        if (CHM == kChmMono || (CHM == kChmAny && cx.channels == 1u)) {
#pragma unroll
            for (int j = 0; j < 4; j++) m[j] = ramp_mult2(rr, 4u * u + j);
        } else if (CHM == kChmStereo) {
            m[0] = m[1] = ramp_mult2(rr, 2u * u);
            m[2] = m[3] = ramp_mult2(rr, 2u * u + 1u);
        } else if (CHM == kChmMul4) {
            const uint32_t f0 = __umulhi(4u * u, cx.ch_magic);
            chan0 = 4u * u - f0 * cx.channels;
            m[0] = m[1] = m[2] = m[3] = ramp_mult2(rr, f0);
        } else {
            // two channels or more: the unit touches frame f0 and possibly f0+1
            const uint32_t f0 = __umulhi(4u * u, cx.ch_magic);
            chan0 = 4u * u - f0 * cx.channels;
            const uint32_t ma = ramp_mult2(rr, f0);
            const uint32_t mb = ramp_mult2(rr, f0 + 1u);
#pragma unroll
            for (int j = 0; j < 4; j++) m[j] = (chan0 + j >= cx.channels) ? mb : ma;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int32_t s16 = unit_s16<B>(r, j, cx.in_le);
            if (B == 2 && cx.attenuation != OHP_UNITY_ATTENUATION) {
                // ((TInt)sample) * iAttenuation / 256 in UNSIGNED 32-bit arithmetic, truncated to 16 bits (Msg.cpp:2746)
                s16 = (int32_t)(int16_t)(((uint32_t)s16 * cx.attenuation) >> 8);
            }
            // (s16 * mult) >> 15 kept to 16 bits == high half of s16 * (2*mult) (Msg.cpp:865)
            uint32_t v = (uint32_t)(s16 * (int32_t)m[j]);
            if (B == 4) {
                // bytes: hi, lo, 00, channel tag on 6-channel audio (Msg.cpp:880-891)
                uint32_t c = chan0 + j;
                if (CHM == kChmOther || CHM == kChmAny) c = (c >= cx.channels) ? c - cx.channels : c;
                v = (v & 0xffff0000u) | (cx.tag6 ? (c << 4) : 0u);
            }
            o[j] = v;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t v = unit_extract<B>(r, j, cx.in_le);
            if (B == 2 && cx.attenuation != OHP_UNITY_ATTENUATION) {
                const uint32_t s = (uint32_t)((int32_t)v >> 16);
                v = ((s * cx.attenuation) >> 8) << 16;
            }
            o[j] = v;
        }
    }
    unit_pack<B>(o, w, cx.out_le, cx.ramped);
}

__device__ __forceinline__ void load_ctx(const ChunkRec& cr, UnitCtx& cx, RampRegs& rr)
{
    cx.channels = cr.channels;
    cx.ch_magic = cr.ch_magic;
    cx.attenuation = cr.attenuation;
    cx.ramped = (cr.mode & kModeRamped) != 0;
    cx.in_le = (cr.mode & kModeInLe) != 0;
    cx.out_le = (cr.mode & kModeOutLe) != 0;
    cx.tag6 = (cr.mode & kModeTag6) != 0;
    rr.c = cr.ramp_c; rr.sign = cr.ramp_sign; rr.total = cr.ramp_total; rr.magic = cr.ramp_magic; rr.shift = cr.ramp_shift;
}

// The image of a chunk lands in its slot at in_addr + head, head = source address & 15 (the bulk load moves whole
// 16-byte words).  Every lane takes groups of four units (16 subsamples = B x 16 bytes), whatever the head:
//   ALIGNED (head == 0)  B 128-bit shared-memory loads per group, nothing else;
//   otherwise            B + 1 of them from the 16-byte boundary below the group, then the group's 4 B words are cut
//                        out of the 4 B + 4 loaded: a word offset (head / 4: the same for the whole chunk, so a
//                        warp-uniform switch with compile-time register indices in every arm) and a funnel shift by the
//                        byte offset (head % 4).
// The transformed group goes back with B 128-bit stores to in_addr + g * 16 B -- the output image starts ON the 16-byte
// boundary, up to 15 bytes below the input image.  In place: the stores of group g cover [g * 16 B, (g + 1) * 16 B) of the
// slot, the loads of a LATER step start at or above that, and inside a step every lane loads before any lane stores
// (__syncwarp).  Which way the image then leaves the slot is the store's business (store_image_warp).
template <int B>
__device__ __forceinline__ void cut_group_words(const uint32_t (&v)[4 * B + 4], uint32_t (&r)[4 * B + 1], uint32_t word_off, uint32_t bit_off)
{
    switch (word_off) {
    case 0:
#pragma unroll
        for (int k = 0; k < 4 * B; k++) r[k] = __funnelshift_r(v[k], v[k + 1], bit_off);
        break;
    case 1:
#pragma unroll
        for (int k = 0; k < 4 * B; k++) r[k] = __funnelshift_r(v[k + 1], v[k + 2], bit_off);
        break;
    case 2:
#pragma unroll
        for (int k = 0; k < 4 * B; k++) r[k] = __funnelshift_r(v[k + 2], v[k + 3], bit_off);
        break;
    default:
#pragma unroll
        for (int k = 0; k < 4 * B; k++) r[k] = __funnelshift_r(v[k + 3], v[k + 4], bit_off);
        break;
    }
    r[4 * B] = 0;
}

template <int B, uint32_t CHM, bool ALIGNED>
__device__ __noinline__ void transform_wide(const ChunkRec& cr, uint32_t table, uint32_t in_addr, uint32_t t)
{
    UnitCtx cx;
    RampRegs rr;
    rr.table = table;
    load_ctx(cr, cx, rr);
    const uint32_t groups = (cr.units + 3u) >> 2;
    const uint32_t word_off = cr.head >> 2;
    const uint32_t bit_off = (cr.head & 3u) * 8u;
    // kGroupsPerStep independent groups per lane per step: one basic block, so their dependency chains interleave
    // (a lone warp per chunk has no other warp to hide LDS / IMAD latency behind)
    for (uint32_t gb = 0; gb < groups; gb += 32u * kGroupsPerStep) { // warp-uniform trip count (the step synchronises the warp)
        const uint32_t g0 = gb + t;
        uint32_t r[kGroupsPerStep][4 * B + 1];
        uint32_t w[kGroupsPerStep][4 * B];
#pragma unroll
        for (int s = 0; s < kGroupsPerStep; s++) {
            const uint32_t g = min(g0 + 32u * s, groups - 1u); // clamp: the surplus copy recomputes the last group, unstored
            const uint32_t a = in_addr + g * (16u * B);
            if constexpr (ALIGNED) {
#pragma unroll
                for (int i = 0; i < B; i++) {
                    const uint4 v = lds128(a + 16u * i);
                    r[s][4 * i + 0] = v.x; r[s][4 * i + 1] = v.y; r[s][4 * i + 2] = v.z; r[s][4 * i + 3] = v.w;
                }
                r[s][4 * B] = 0;
            } else {
                uint32_t v[4 * B + 4];
#pragma unroll
                for (int i = 0; i <= B; i++) {
                    const uint4 q = lds128(a + 16u * i);
                    v[4 * i + 0] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
                }
                cut_group_words<B>(v, r[s], word_off, bit_off);
            }
        }
        __syncwarp(); // (clamped lanes re-read a group another lane is about to overwrite in place)
#pragma unroll
        for (int s = 0; s < kGroupsPerStep; s++) {
            const uint32_t g = min(g0 + 32u * s, groups - 1u);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint32_t ru[B + 1], wu[B];
#pragma unroll
                for (int k = 0; k <= B; k++) ru[k] = r[s][i * B + k];
                process_unit<B, CHM>(cx, rr, 4u * g + i, ru, wu);
#pragma unroll
                for (int k = 0; k < B; k++) w[s][i * B + k] = wu[k];
            }
        }
#pragma unroll
        for (int s = 0; s < kGroupsPerStep; s++) {
            const uint32_t g = g0 + 32u * s;
            if (g < groups) {
                const uint32_t d = in_addr + g * (16u * B);
#pragma unroll
                for (int i = 0; i < B; i++) {
                    sts128(d + 16u * i, make_uint4(w[s][4 * i + 0], w[s][4 * i + 1], w[s][4 * i + 2], w[s][4 * i + 3]));
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The other IPcmProcessor sinks of the reference tree (output size differs from the input size).  One subsample at a
// time, any alignment: these carry 1 ms flywheel blocks and Songcast frames, not the bulk of the traffic.

// Left-justified big-endian value of the B-byte subsample whose first byte is at shared address a (low bytes zero).
__device__ __forceinline__ uint32_t load_subsample(uint32_t a, uint32_t B, bool le)
{
    const uint32_t aw = a & ~3u;
    const uint32_t x = __funnelshift_r(lds32(aw), lds32(aw + 4u), (a & 3u) * 8u); // memory order, first byte lowest
    const uint32_t v = le ? (x << (8u * (4u - B))) : __byte_perm(x, 0u, 0x0123);
    return v & (0xffffffffu << (8u * (4u - B)));
}

// MsgPlayablePcm::ReadBlock's arithmetic on one subsample (attenuation, then ramp), on the left-justified value.
__device__ __forceinline__ uint32_t ramp_subsample(uint32_t v, uint32_t B, const UnitCtx& cx, const RampRegs& rr, uint32_t frame, uint32_t chan)
{
    if (B == 2 && cx.attenuation != OHP_UNITY_ATTENUATION) {
        const uint32_t s = (uint32_t)((int32_t)v >> 16);
        v = ((s * cx.attenuation) >> 8) << 16;
    }
    if (cx.ramped) {
        const int32_t s16 = (int32_t)v >> 16; // 8-bit: the value's low byte is already zero
        const uint32_t prod2 = (uint32_t)(s16 * (int32_t)ramp_mult2(rr, frame));
        v = prod2 & (B == 1 ? 0xff000000u : 0xffff0000u);
        if (B == 4 && cx.tag6) v |= chan << 4;
    }
    return v;
}

// FlywheelInput::DoProcessFragment (StarvationRamper.cpp:117-186): interleaved -> planar, 4 bytes per subsample,
// big-endian left-justified, zero padded; channel c's plane starts at dst + c * plane_bytes.  Written straight to global.
__device__ __noinline__ void convert_planar32(const ChunkRec& cr, uint32_t table, uint32_t in_addr, uint8_t* dst, uint32_t plane_bytes, uint32_t lane)
{
    UnitCtx cx;
    RampRegs rr;
    rr.table = table;
    load_ctx(cr, cx, rr);
    const uint32_t B = (cr.variant & 3u) + 1u;
    const uint32_t ch = cr.channels;
    const uint32_t frames = cr.frames;
    const bool words = ((reinterpret_cast<uint64_t>(dst) | plane_bytes) & 3u) == 0;
    for (uint32_t c = 0; c < ch; c++) {
        uint8_t* plane = dst + (uint64_t)c * plane_bytes;
        for (uint32_t f = lane; f < frames; f += 32) {
            uint32_t v = load_subsample(in_addr + cr.head + (f * ch + c) * B, B, cx.in_le);
            v = ramp_subsample(v, B, cx, rr, f, c);
            const uint32_t be = __byte_perm(v, 0u, 0x0123); // most significant byte first in memory
            if (words) {
                reinterpret_cast<uint32_t*>(plane)[f] = be;
            } else {
                for (uint32_t i = 0; i < 4; i++) plane[4u * f + i] = (uint8_t)(be >> (8u * i));
            }
        }
    }
}

// RampGenerator::ProcessFragment (StarvationRamper.cpp:281-326): 32-bit big-endian in, the top OB bytes out
// (32 -> three bytes and a zero).  Contracts in place: unit u (4 subsamples, 16 bytes) -> OB words.
template <int OB>
__device__ __noinline__ void convert_from32(const ChunkRec& cr, uint32_t table, uint32_t in_addr, uint32_t out_addr, uint32_t lane)
{
    UnitCtx cx;
    RampRegs rr;
    rr.table = table;
    load_ctx(cr, cx, rr);
    const uint32_t ch = cr.channels;
    const uint32_t units = cr.units;
    for (uint32_t u0 = 0; u0 < units; u0 += 32) {
        const uint32_t u = u0 + lane;
        uint32_t o[4];
        if (u < units) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t k = 4u * u + j;
                const uint32_t f = cr.ch_magic ? __umulhi(k, cr.ch_magic) : k; // k / channels
                uint32_t v = load_subsample(in_addr + cr.head + 4u * k, 4u, cx.in_le);
                v = ramp_subsample(v, 4u, cx, rr, f, k - f * ch);
                o[j] = (OB == 4) ? (v & 0xffffff00u) : v;
            }
        } else {
            o[0] = o[1] = o[2] = o[3] = 0;
        }
        __syncwarp();
        if (u < units) {
            uint32_t w[OB];
            unit_pack<OB>(o, w, false, false);
#pragma unroll
            for (int i = 0; i < OB; i++) sts32(out_addr + u * (4u * OB) + 4u * i, w[i]);
        }
    }
}

// Sender::DoProcessFragment (Av/Songcast/Sender.cpp:356-377): OCH (1 or 2) channels starting at `first`, the top
// DB = min(B, 3) bytes of each subsample.  Contracts in place: unit u = 4 frames -> OCH * DB words.
template <int DB, int OCH>
__device__ __noinline__ void convert_songcast(const ChunkRec& cr, uint32_t table, uint32_t in_addr, uint32_t out_addr, uint32_t first, uint32_t lane)
{
    UnitCtx cx;
    RampRegs rr;
    rr.table = table;
    load_ctx(cr, cx, rr);
    const uint32_t B = (cr.variant & 3u) + 1u;
    const uint32_t ch = cr.channels;
    const uint32_t units = (cr.frames + 3u) >> 2;
    for (uint32_t u0 = 0; u0 < units; u0 += 32) {
        const uint32_t u = u0 + lane;
        uint32_t o[4 * OCH];
        if (u < units) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t f = min(4u * u + j, cr.frames - 1u); // surplus frames of the last unit: recomputed, not stored beyond
#pragma unroll
                for (int c = 0; c < OCH; c++) {
                    uint32_t v = load_subsample(in_addr + cr.head + (f * ch + first + c) * B, B, cx.in_le);
                    o[j * OCH + c] = ramp_subsample(v, B, cx, rr, f, first + c);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4 * OCH; j++) o[j] = 0;
        }
        __syncwarp();
        if (u < units) {
#pragma unroll
            for (int h = 0; h < OCH; h++) {
                const uint32_t (&oh)[4] = reinterpret_cast<const uint32_t (&)[4]>(o[4 * h]);
                uint32_t w[DB];
                unit_pack<DB>(oh, w, false, false);
#pragma unroll
                for (int i = 0; i < DB; i++) sts32(out_addr + u * (4u * OCH * DB) + (h * DB + i) * 4u, w[i]);
            }
        }
    }
}

// Fill `bytes` bytes at shared address a (16-byte aligned) with what MsgPlayableSilence::ReadBlock would emit:
// used when silence goes to a sink that converts it (planar / Songcast).
__device__ __noinline__ void silence_to_smem(uint32_t a, uint32_t bytes, uint32_t channels, uint32_t lane)
{
    const uint32_t vecs = (bytes + 15u) >> 4;
    for (uint32_t v = lane; v < vecs; v += 32) {
        uint4 z = make_uint4(0, 0, 0, 0);
        if (channels == 6 && v < 2) {
            // 00 00 00 c0 with c0 = 0x00,0x10..0x70 over the first 32 bytes (Msg.cpp:2877)
            const uint32_t c0 = 4u * v;
            z = make_uint4((c0 << 4) << 24, ((c0 + 1u) << 4) << 24, ((c0 + 2u) << 4) << 24, ((c0 + 3u) << 4) << 24);
        }
        sts128(a + 16u * v, z);
    }
}

// The general path: any channel count, the image anywhere in its slot.  One unit (four subsamples) per lane per step, word
// loads realigned with a funnel shift -- a quarter of the work per step of the group path above, and a quarter of its CODE:
// that is the point.  A batch of mixed formats has every instantiation hot at the same time (16 consumer warps per SM, each
// in whatever its chunk needs) and the SM's instruction caches are small (L0 ~6 KB, L1.5 32 KB): with one group-wide
// instantiation per (depth, channel layout, alignment) -- 32 functions, 143 KB of SASS -- the mixed-format BASELINE config
// ran at 0.53 of the copy peak with "no instruction" as its first stall reason (4.5 stalled warps per issue,
// profiles/r02_config4_wide_ncu_summary.txt).  So the two layouts the uniform configs consist of (stereo, channel counts
// that are multiples of four, image on a 16-byte boundary) keep their lean group-wide instantiation, and everything else
// shares ONE small instantiation per depth.
// In place: the output image starts at in_addr, at or below the input image (in_addr + head), so a step's stores never
// reach bytes a LATER step still has to read; inside a step every lane loads before any lane stores (__syncwarp).
template <int B>
__device__ __noinline__ void transform_any(const ChunkRec& cr, uint32_t table, uint32_t in_addr, uint32_t lane)
{
    UnitCtx cx;
    RampRegs rr;
    rr.table = table;
    load_ctx(cr, cx, rr);
    const uint32_t units = cr.units;
    const uint32_t src = in_addr + (cr.head & ~3u);
    const uint32_t fshift = (cr.head & 3u) * 8u;
    // kAnyUnits units per lane and step, independent of each other (1: see OHP_ANY_UNITS)
    for (uint32_t u0 = 0; u0 < units; u0 += 32 * kAnyUnits) {
        uint32_t raw[kAnyUnits][B + 1], r[kAnyUnits][B + 1], w[kAnyUnits][B];
#pragma unroll
        for (int k = 0; k < kAnyUnits; k++) {
            const uint32_t u = u0 + 32 * k + lane;
            const uint32_t a = src + u * (4u * B);
            if (u < units) {
#pragma unroll
                for (int i = 0; i <= B; i++) raw[k][i] = lds32(a + 4u * i);
            } else {
#pragma unroll
                for (int i = 0; i <= B; i++) raw[k][i] = 0;
            }
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < kAnyUnits; k++) {
            const uint32_t u = u0 + 32 * k + lane;
#pragma unroll
            for (int i = 0; i < B; i++) r[k][i] = __funnelshift_r(raw[k][i], raw[k][i + 1], fshift);
            r[k][B] = 0;
            process_unit<B, kChmAny>(cx, rr, u, r[k], w[k]);
        }
#pragma unroll
        for (int k = 0; k < kAnyUnits; k++) {
            const uint32_t u = u0 + 32 * k + lane;
            if (u < units) {
                const uint32_t d = in_addr + u * (4u * B);
#pragma unroll
                for (int i = 0; i < B; i++) sts32(d + 4u * i, w[k][i]);
            }
        }
    }
}

template <int B>
__device__ __forceinline__ void transform_dispatch(const ChunkRec& cr, uint32_t table, uint32_t in_addr, uint32_t t)
{
    const uint32_t chm = (cr.variant >> 2) & 3u;
    const bool aligned = (cr.variant & 16u) != 0; // head == 0
    if (aligned && chm == kChmStereo) transform_wide<B, kChmStereo, true>(cr, table, in_addr, t);
    else if (aligned && chm == kChmMul4) transform_wide<B, kChmMul4, true>(cr, table, in_addr, t);
    else transform_any<B>(cr, table, in_addr, t);
}

// ---------------------------------------------------------------------------------------------
// storing a finished image: `bytes` bytes at shared address s_addr (ANY byte address) -> global dst, by ONE warp.
// The <= 15 bytes before the destination's first 16-byte boundary and after its last go out as byte stores.  The
// interior leaves with ONE TMA bulk store when image and destination agree mod 16; otherwise every lane assembles
// 16-byte destination words from two 128-bit shared-memory loads (word offset by a warp-uniform switch, byte offset by
// a funnel shift) and writes them with coalesced 128-bit streaming stores.

__device__ __forceinline__ void store_ragged(uint32_t s_addr, uint8_t* dst, uint32_t from, uint32_t to, uint32_t lane)
{
    // at most 15 bytes
    const uint32_t i = from + lane;
    if (i < to) dst[i] = (uint8_t)lds8(s_addr + i);
}

__device__ __forceinline__ uint4 cut_vector(const uint4& a, const uint4& b, uint32_t word_off, uint32_t bit_off)
{
    uint4 v;
    switch (word_off) {
    case 0:
        v.x = __funnelshift_r(a.x, a.y, bit_off); v.y = __funnelshift_r(a.y, a.z, bit_off);
        v.z = __funnelshift_r(a.z, a.w, bit_off); v.w = __funnelshift_r(a.w, b.x, bit_off);
        break;
    case 1:
        v.x = __funnelshift_r(a.y, a.z, bit_off); v.y = __funnelshift_r(a.z, a.w, bit_off);
        v.z = __funnelshift_r(a.w, b.x, bit_off); v.w = __funnelshift_r(b.x, b.y, bit_off);
        break;
    case 2:
        v.x = __funnelshift_r(a.z, a.w, bit_off); v.y = __funnelshift_r(a.w, b.x, bit_off);
        v.z = __funnelshift_r(b.x, b.y, bit_off); v.w = __funnelshift_r(b.y, b.z, bit_off);
        break;
    default:
        v.x = __funnelshift_r(a.w, b.x, bit_off); v.y = __funnelshift_r(b.x, b.y, bit_off);
        v.z = __funnelshift_r(b.y, b.z, bit_off); v.w = __funnelshift_r(b.z, b.w, bit_off);
        break;
    }
    return v;
}

// (out of line: keeps the consumer loop's own code short, see transform_any)
__device__ __noinline__ void store_words_cut(uint32_t base, uint4* d4, uint32_t words, uint32_t lane)
{
    const uint32_t abase = base & ~15u;
    const uint32_t word_off = (base & 15u) >> 2;
    const uint32_t bit_off = (base & 3u) * 8u;
    for (uint32_t w = lane; w < words; w += 32) {
        const uint4 a = lds128(abase + 16u * w);
        const uint4 b = lds128(abase + 16u * w + 16u); // at most 16 bytes past the image: inside the slot (kSlotBack)
        stg128_stream(d4 + w, cut_vector(a, b, word_off, bit_off));
    }
}

__device__ __forceinline__ void store_image_warp(uint32_t s_addr, uint8_t* dst, uint32_t bytes, uint32_t lane)
{
    const uint32_t lead = (uint32_t)((16u - (reinterpret_cast<uint64_t>(dst) & 15u)) & 15u);
    const uint32_t head_n = min(lead, bytes);
    const uint32_t words = (bytes - head_n) >> 4;
    const uint32_t tail_at = head_n + (words << 4);
    store_ragged(s_addr, dst, 0, head_n, lane);
    store_ragged(s_addr, dst, tail_at, bytes, lane);
    const uint32_t base = s_addr + head_n;
    if ((base & 15u) == 0) {
        if (lane == 0 && words != 0) {
            tma_store(dst + head_n, base, words << 4);
        }
    } else {
        store_words_cut(base, reinterpret_cast<uint4*>(dst + head_n), words, lane);
    }
}

// MsgPlayableSilence::ReadBlock (Msg.cpp:2874-2893): zeros; with 6 channels every emitted block of
// maxBytes starts with 00 00 00 c0 for c0 = 0x00,0x10..0x70 (32 bytes, whatever the bit depth).
// Written straight to global memory by the chunk's consumer warp (t = lane).
__device__ __noinline__ void write_silence(uint8_t* dst, uint32_t bytes, uint32_t channels, uint32_t B, uint32_t t)
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
    if (t < head_n) dst[t] = (uint8_t)value_at(t);
    if (t < bytes - tail_at) {
        const uint32_t i = tail_at + t;
        dst[i] = (uint8_t)value_at(i);
    }
    uint4* d4 = reinterpret_cast<uint4*>(dst + head_n);
    for (uint32_t w = t; w < words; w += 32) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (channels == 6) {
            const uint32_t i0 = head_n + 16u * w;
            if ((i0 % block) < 32u || (i0 % block) + 16u > block) {
                uint32_t tmp[4] = {0, 0, 0, 0};
                for (uint32_t i = 0; i < 16; i++) tmp[i >> 2] |= value_at(i0 + i) << (8 * (i & 3));
                v = make_uint4(tmp[0], tmp[1], tmp[2], tmp[3]);
            }
        }
        stg128_stream(d4 + w, v);
    }
}

} // namespace ohp
