// ohp_ramp_convert_kernel.cuh -- the kernels of the hot path: ramp_convert_kernel (persistent, warp-specialised: one
// loader warp + eight consumer warps per CTA over a shared-memory byte ring; the device helpers it uses and the full
// description are in ohp_kernels.cuh) and checksum_kernel (per-stream 64-bit checksums for the multi-GPU gather).
// Launched from ohp_capi.cu.
#pragma once

#include "ohp_kernels.cuh"
#include "../../include/ohp_schedule.h"

namespace ohp {

__constant__ uint32_t c_ch_magic[33]; // ceil(2^32 / ch); [0],[1] = 0 (mono divides by one)

// Decode and check one descriptor (the reference's ASSERTs, restated); fill the consumer record.
// Returns 0, or the error bits to report (the chunk is then skipped).
__device__ __forceinline__ uint32_t decode_chunk(const KernelParams& p, const uint4& d0, const uint4& d1, ChunkRec& r,
                                                 uint64_t& src_off)
{
    DescFields d;
    d.src_off = src_off = (uint64_t)d0.x | ((uint64_t)d0.y << 32);
    d.dst_off = (uint64_t)d0.z | ((uint64_t)d0.w << 32);
    d.bytes = d1.x;
    d.ramp_start = d1.y & 0xffffu;
    d.ramp_end = d1.y >> 16;
    d.attenuation = d1.z & 0xffffu;
    d.bit_depth = (d1.z >> 16) & 0xffu;
    d.channels = d1.z >> 24;
    d.flags = d1.w & 0xffu;
    d.out_fmt = (d1.w >> 8) & 0xffu;
    d.aux = d1.w >> 16;
    r.kind = kSkip;
    DescDerived dv;
    const uint32_t err = check_desc_fields(d, p.in_bytes, p.out_bytes, dv);
    if (err) return err == 1u ? kErrInvalidDesc : kErrOutOfRange;
    if (d.bytes == 0) return 0; // MsgPlayable::Read only calls ReadBlock when iSize > 0 (Msg.cpp:2649)

    const bool silence = (d.flags & OHP_F_SILENCE) != 0;
    const bool packed = d.out_fmt == OHP_OUT_PACKED_BE || d.out_fmt == OHP_OUT_PACKED_LE;
    const uint32_t B = d.bit_depth >> 3;
    const uint32_t channels = d.channels;
    const uint64_t dst = reinterpret_cast<uint64_t>(p.out) + d.dst_off;
    r.kind = silence ? (packed ? kSilence : kSilenceConv) : kPcm;
    r.bytes = d.bytes;
    r.head = silence ? 0u : (uint32_t)((reinterpret_cast<uint64_t>(p.in) + src_off) & 15u);
    r.channels = channels;
    r.ch_magic = c_ch_magic[channels];
    r.attenuation = silence ? OHP_UNITY_ATTENUATION : d.attenuation;
    r.units = (d.bytes / B + 3u) >> 2;
    r.frames = dv.frames;
    r.out_fmt = d.out_fmt;
    r.aux = d.aux;
    r.out_bytes = dv.out_bytes;
    r.dst_lo = (uint32_t)dst;
    r.dst_hi = (uint32_t)(dst >> 32);
    const bool ramped = (d.flags & OHP_F_RAMP_ENABLED) != 0 && !silence; // silence is never ramped (Msg.cpp:2874-2893)
    const bool in_le = (d.flags & OHP_F_IN_LITTLE_ENDIAN) != 0 && B > 1 && !silence;
    const bool out_le = (d.out_fmt == OHP_OUT_PACKED_LE) && B > 1;
    const bool transform = ramped || (in_le != out_le) || r.attenuation != OHP_UNITY_ATTENUATION;
    r.mode = (ramped ? kModeRamped : 0u) | (in_le ? kModeInLe : 0u) | (out_le ? kModeOutLe : 0u)
           | ((channels == 6) ? kModeTag6 : 0u) | (transform ? kModeTransform : 0u);
    const uint32_t chm = channels == 2 ? kChmStereo : ((channels & 3u) == 0 ? kChmMul4 : (channels == 1 ? kChmMono : kChmOther));
    const bool aligned = r.head == 0; // the image starts on a 16-byte boundary of its slot
    r.variant = (B - 1u) | (chm << 2) | (aligned ? 16u : 0u);
    make_ramp_const(r, d.ramp_start, d.ramp_end, dv.frames);
    return 0;
}

__device__ __forceinline__ void report(const KernelParams& p, uint32_t bits, uint64_t chunk)
{
    atomicOr(&p.status[0], bits);
    atomicCAS(&p.status[1], 0u, (uint32_t)(chunk + 1 > 0xffffffffull ? 0xffffffffull : chunk + 1));
}

// Optional instrumentation (-DOHP_PROFILE_WAITS): cycles each role spends blocked on each barrier, summed over CTAs
// into status[4..] in units of 4096 cycles.  [4] loader/empty_in [5] consumer/full_in [6] consumer/empty_out
// [7] storer/full_out [8] CTA lifetime [9] loader busy (decode) [10] consumer busy (transform)
#ifdef OHP_PROFILE_WAITS
#define OHP_ACC(var, expr) (var) += (expr)
#define OHP_FLUSH(slot, var) atomicAdd(&p.status[slot], (uint32_t)((var) >> 12))
#else
#define OHP_ACC(var, expr) (void)(expr)
#define OHP_FLUSH(slot, var) (void)0
#endif

// Persistent, warp-specialised CTAs (see ohp_kernels.cuh): chunks are dealt block-cyclically to the CTAs; inside the
// CTA the chunk with ordinal k uses barrier pair k % kBarPairs and is transformed by whichever consumer warp draws
// ticket k (OHP_DYNAMIC) -- so a warp that met a run of expensive chunks does not hold the in-order ring up while
// its neighbours idle -- or by warp k % kConsumerWarps (static).
// SERIAL_PLACE: how the loader warp places chunks in the ring (compiled twice; the context picks per batch shape).
template <bool SERIAL_PLACE>
__global__ void __launch_bounds__(kThreads) ramp_convert_kernel(const KernelParams p)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    SharedStorage& sm = *reinterpret_cast<SharedStorage*>(smem_raw);
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < kBarPairs; s++) {
            mbar_init(smem_u32(&sm.full[s]), 1);
            mbar_init(smem_u32(&sm.empty[s]), 1);
        }
        sm.next_ticket = 0;
        fence_mbar_init();
    }
    for (uint32_t i = threadIdx.x; i < OHP_RAMP_TABLE_ENTRIES; i += kThreads) sm.table2[i] = p.table2[i];
    __syncthreads();

    [[maybe_unused]] uint32_t n_rounds = 0, n_inflight = 0;
    [[maybe_unused]] long long w_loader = 0, w_full = 0, w_store = 0, w_xform = 0, w_fence = 0, w_issue = 0;
#ifdef OHP_PROFILE_WAITS
    const long long t_begin = clock64();
#endif
    // chunks of this CTA: ordinal k <-> chunk cta_chunk_index(k)
    const uint64_t my_n = cta_chunk_count(p.n, blockIdx.x, gridDim.x, p.chunk_block);
    const uint32_t ring = smem_u32(&sm.ring[0]);

    if (warp == 0) {
        // ------------------------------------------------------------------ loader
        const uint4* dp = reinterpret_cast<const uint4*>(p.descs);
        uint32_t wr = 0;              // next free byte of the ring
        uint32_t free_bytes = p.cap_bytes;
        uint64_t rd = 0;              // oldest chunk whose slot has not been reclaimed yet
        uint32_t my_slot_bytes = 0;   // lane s: bytes to give back when ring slot s (barrier pair s) is released
        // The whole warp walks the issue loop in lockstep (warp-uniform control flow) and lane 0 performs the side
        // effects: what the loop needs of chunk j lives in lane j's registers and arrives by shuffle, so issuing a chunk
        // costs a few dozen ALU cycles instead of a chain of dependent shared-memory reads.  The descriptors of the NEXT
        // batch are fetched before the loop, so their global-memory latency hides behind it.
        uint4 d0 = make_uint4(0, 0, 0, 0), d1 = d0;
        if (lane < my_n) {
            const uint64_t c = cta_chunk_index(lane, blockIdx.x, gridDim.x, p.chunk_block);
            d0 = __ldg(dp + 2 * c);
            d1 = __ldg(dp + 2 * c + 1);
        }
        for (uint64_t base = 0; base < my_n; base += 32) {
            // all lanes: decode 32 descriptors into the record table (the half the consumers are done with:
            // at most kRingSlots <= 32 chunks are ever in flight)
            const uint64_t k = base + lane;
            ChunkRec r;
            r.kind = kSkip;
            uint64_t src_al = 0;
            if (k < my_n) {
                uint64_t src_off;
                const uint32_t err = decode_chunk(p, d0, d1, r, src_off);
                if (err) report(p, err, cta_chunk_index(k, blockIdx.x, gridDim.x, p.chunk_block));
                src_al = reinterpret_cast<uint64_t>(p.in) + src_off - r.head;
            }
            if (k + 32 < my_n) {
                const uint64_t c = cta_chunk_index(k + 32, blockIdx.x, gridDim.x, p.chunk_block);
                d0 = __ldg(dp + 2 * c);
                d1 = __ldg(dp + 2 * c + 1);
            }
            sm.rec[(uint32_t)(k & (kRecSlots - 1))] = r;
            const uint32_t my_kind = r.kind;
            uint32_t my_span = 0;
            if (my_kind == kPcm || my_kind == kSilenceConv) my_span = (r.head + r.bytes + 15u) & ~15u;
            __syncwarp();
            const uint32_t count = (uint32_t)(my_n - base < 32 ? my_n - base : 32);
            [[maybe_unused]] const uint32_t my_need = my_span ? kSlotFront + my_span + kSlotBack : 0u;
#if OHP_LOADER == 1
            uint32_t j = 0;
            while (j < count) {
                const uint64_t it0 = base + j;
                // (1) sweep the slots in flight, oldest first, without blocking: lane l tests slot rd + l
                {
                    const uint32_t inflight = (uint32_t)(it0 - rd); // <= kRingSlots <= 32
                    const uint64_t r = rd + lane;
                    const uint32_t rs = (uint32_t)(r % kRingSlots); // lane rs remembers what the slot holds
                    const bool released = lane < inflight
                        && mbar_test(smem_u32(&sm.empty[(uint32_t)(r % kBarPairs)]), (uint32_t)(r / kBarPairs) & 1u);
                    const uint32_t mask = __ballot_sync(0xffffffffu, released);
                    const uint32_t nrel = mask == 0xffffffffu ? 32u : (uint32_t)__ffs((int)~mask) - 1u; // reclaim is in order
                    uint32_t bytes = __shfl_sync(0xffffffffu, my_slot_bytes, rs);
                    bytes = lane < nrel ? bytes : 0u;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
                    free_bytes += bytes;
                    rd += nrel;
                }
                // (2) place as many of the next chunks as fit right now, all at once: lane g stands for chunk j + g.
                //     Slots are contiguous, so the first chunk that would cross the end of the ring moves to offset 0
                //     and the bytes it skipped are charged to it; a prefix sum of the needs gives every candidate its
                //     offset and the running cost, and "fits" is monotone in g, so one ballot counts the chunks placed.
                uint32_t fit = 0, my_wr = 0;
                // Two ways to place (the kernel is compiled with each; the context measures which one a batch shape
                // prefers, see TuneEntry).  When the ring is nearly full (large uniform chunks: the consumers set the
                // pace) a round places one to three chunks and what counts is how soon after a release the next load
                // starts: a short serial loop.  When there is room (small or mixed chunks: the loader sets the pace) what
                // counts is chunks per round: the scan.
                if (SERIAL_PLACE) {
                    uint32_t wr_s = wr, free_s = free_bytes;
                    for (uint32_t g = 0; g < kSerialWidth && j + g < count; g++) {
                        const uint32_t need = __shfl_sync(0xffffffffu, my_need, j + g);
                        const bool wrap = wr_s + need > kRingBytes;       // the slot must be contiguous: skip the end of the ring
                        const uint32_t waste = wrap ? kRingBytes - wr_s : 0u;
                        if (free_s < need + waste || it0 + g - rd >= p.cap_chunks) break;
                        if (wrap) wr_s = 0;
                        if (lane == g) my_wr = wr_s;
                        if (lane == (uint32_t)((it0 + g) % kRingSlots)) my_slot_bytes = need + waste;
                        free_s -= need + waste;
                        wr_s += need;
                        fit++;
                    }
                    if (fit == 0) {
                        // ring full: block on the oldest slot (one lane polls), then sweep again
                        if (lane == 0) {
                            OHP_ACC(w_loader, mbar_wait(smem_u32(&sm.empty[(uint32_t)(rd % kBarPairs)]), (uint32_t)(rd / kBarPairs) & 1u, p.status));
                        }
                        __syncwarp();
                        continue;
                    }
                    wr = wr_s;
                    free_bytes = free_s;
                }
                else {
                    const uint32_t width = count - j < kIssueWidth ? count - j : kIssueWidth;
                    uint32_t need = __shfl_sync(0xffffffffu, my_need, (j + lane) & 31u);
                    need = lane < width ? need : 0u;
                    uint32_t incl = need;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= (uint32_t)o) incl += y;
                    }
                    const uint32_t excl = incl - need;
                    const uint32_t wmask = __ballot_sync(0xffffffffu, wr + incl > kRingBytes);
                    const uint32_t wl = wmask ? (uint32_t)__ffs((int)wmask) - 1u : 32u;   // the lane that wraps (32: none)
                    const uint32_t excl_w = __shfl_sync(0xffffffffu, excl, wl & 31u);
                    const uint32_t waste = wl < 32u ? kRingBytes - (wr + excl_w) : 0u;
                    const uint32_t cost_incl = incl + (lane >= wl ? waste : 0u);
                    const bool ok = lane < width && cost_incl <= free_bytes && it0 + lane - rd < p.cap_chunks;
                    const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
                    fit = okmask == 0xffffffffu ? 32u : (uint32_t)__ffs((int)~okmask) - 1u;
                    if (fit == 0) {
                        // ring full: block on the oldest slot (one lane polls), then sweep again
                        if (lane == 0) {
                            OHP_ACC(w_loader, mbar_wait(smem_u32(&sm.empty[(uint32_t)(rd % kBarPairs)]), (uint32_t)(rd / kBarPairs) & 1u, p.status));
                        }
                        __syncwarp();
                        continue;
                    }
                    my_wr = lane < wl ? wr + excl : excl - excl_w;
                    // ring slot (it0 + g) % kRingSlots remembers what chunk g took, for the sweep that reclaims it
                    const uint32_t cost = need + (lane == wl ? waste : 0u);
                    const uint32_t g = (lane + kRingSlots - (uint32_t)(it0 % kRingSlots)) % kRingSlots;
                    const uint32_t got = __shfl_sync(0xffffffffu, cost, g & 31u);
                    if (lane < kRingSlots && g < fit) my_slot_bytes = got;
                    const uint32_t last_incl = __shfl_sync(0xffffffffu, incl, fit - 1u);
                    const uint32_t last_cost = __shfl_sync(0xffffffffu, cost_incl, fit - 1u);
                    wr = fit > wl ? last_incl - excl_w : wr + last_incl;
                    free_bytes -= last_cost;
                }
                // (3) lanes 0..fit-1 start chunk j + lane
                {
                    const uint32_t from = (j + lane) & 31u;
                    const uint32_t kind = __shfl_sync(0xffffffffu, my_kind, from);
                    uint32_t span = __shfl_sync(0xffffffffu, my_span, from);
                    const uint64_t src = ((uint64_t)__shfl_sync(0xffffffffu, (uint32_t)(src_al >> 32), from) << 32)
                                       | __shfl_sync(0xffffffffu, (uint32_t)src_al, from);
                    if (lane < fit) {
                        const uint64_t it = it0 + lane;
                        sm.ring_off[(uint32_t)(it & (kRecSlots - 1))] = my_wr;
                        const uint32_t full = smem_u32(&sm.full[(uint32_t)(it % kBarPairs)]);
                        if (kind == kPcm) {
                            const uint8_t* al = reinterpret_cast<const uint8_t*>(src);
                            const uint32_t dst_smem = ring + my_wr + kSlotFront;
                            const uint64_t room = (uint64_t)(p.in + p.in_bytes - al);
                            if (span > room) {
                                // last 16-byte word of the arena is partial: fetch its bytes one by one
                                const uint32_t whole = (uint32_t)(room & ~15ull);
                                for (uint32_t i = whole; i < (uint32_t)room; i++) sm.ring[my_wr + kSlotFront + i] = al[i];
                                span = whole;
                            }
                            if (span != 0) {
                                mbar_arrive_expect_tx(full, span);
                                tma_load(dst_smem, al, span, full);
                            } else {
                                mbar_arrive(full);
                            }
                        } else {
                            mbar_arrive(full); // nothing to load (silence; a converting sink still gets its slot)
                        }
                    }
                }
#ifdef OHP_PROFILE_WAITS
                n_rounds++; n_inflight += (uint32_t)(it0 - rd);
#endif
                j += fit;
            }
#else
            for (uint32_t j = 0; j < count; j++) {
                const uint64_t it = base + j;
                const uint32_t sl = (uint32_t)(it & (kRecSlots - 1));
                const uint32_t bs = (uint32_t)(it % kRingSlots);
                const uint32_t kind = __shfl_sync(0xffffffffu, my_kind, j);
                uint32_t span = __shfl_sync(0xffffffffu, my_span, j);
                const uint64_t src = ((uint64_t)__shfl_sync(0xffffffffu, (uint32_t)(src_al >> 32), j) << 32)
                                   | __shfl_sync(0xffffffffu, (uint32_t)src_al, j);
                const uint32_t need = span ? kSlotFront + span + kSlotBack : 0u;
                const bool wrap = wr + need > kRingBytes;           // the slot must be contiguous: skip the end of the ring
                const uint32_t waste = wrap ? kRingBytes - wr : 0u;
                // reclaim, oldest first, until the slot fits and its barrier pair is free
                while (free_bytes < need + waste || it - rd >= p.cap_chunks) {
                    // one lane polls (32 lanes hammering the same mbarrier slow the SM's barrier unit down measurably)
                    if (lane == 0) {
                        OHP_ACC(w_loader, mbar_wait(smem_u32(&sm.empty[(uint32_t)(rd % kBarPairs)]), (uint32_t)(rd / kBarPairs) & 1u, p.status));
                    }
                    __syncwarp();
                    free_bytes += __shfl_sync(0xffffffffu, my_slot_bytes, (uint32_t)(rd % kRingSlots));
                    rd++;
                }
                if (wrap) wr = 0;
                if (lane == bs) my_slot_bytes = need + waste;
                free_bytes -= need + waste;
                if (lane == 0) {
                    sm.ring_off[sl] = wr;
                    const uint32_t full = smem_u32(&sm.full[(uint32_t)(it % kBarPairs)]);
                    if (kind == kPcm) {
                        const uint8_t* al = reinterpret_cast<const uint8_t*>(src);
                        const uint32_t dst_smem = ring + wr + kSlotFront;
                        const uint64_t room = (uint64_t)(p.in + p.in_bytes - al);
                        if (span > room) {
                            const uint32_t whole = (uint32_t)(room & ~15ull);
                            for (uint32_t i = whole; i < (uint32_t)room; i++) sm.ring[wr + kSlotFront + i] = al[i];
                            span = whole;
                        }
                        if (span != 0) {
                            mbar_arrive_expect_tx(full, span);
                            tma_load(dst_smem, al, span, full);
                        } else {
                            mbar_arrive(full);
                        }
                    } else {
                        mbar_arrive(full);
                    }
                }
                wr += need;
            }
#endif
            __syncwarp();
        }
        if (lane == 0) {
            OHP_FLUSH(4, w_loader);
#ifdef OHP_PROFILE_WAITS
            atomicAdd(&p.status[7], n_rounds);      // issue rounds that placed something
            atomicAdd(&p.status[14], n_inflight >> 4); // chunks in flight when they started, summed (/16)
#endif
#ifdef OHP_PROFILE_WAITS
            atomicAdd(&p.status[8], (uint32_t)((clock64() - t_begin) >> 12));
#endif
        }
    } else {
        // ------------------------------------------------------------------ consumers: one warp per chunk
        const uint32_t cw = warp - 1;
        const uint32_t table = smem_u32(&sm.table2[0]);
        // Deferred release: a consumer hands its slot back when it takes its NEXT chunk, so that the bulk store drains
        // while the ticket is fetched.  OHP_DEFER_RELEASE 0: never, 1: always, 2: in the serial-placement instantiation
        // only -- measured (profiles/README.md, round 2): with the ring kept nearly full by large uniform chunks
        // (configs[1], 12 in flight) it is worth +2 %, where the loader places by prefix sum into a ring with room it
        // costs 0.5-1.7 % (slots come back later and the loader is what sets the pace there).
        constexpr bool kDefer = OHP_DEFER_RELEASE == 1 || (OHP_DEFER_RELEASE == 2 && SERIAL_PLACE);
        constexpr uint32_t kNoPending = 0xffffffffu;
        [[maybe_unused]] uint32_t pending = kNoPending; // barrier pair of the chunk whose slot this warp has yet to hand back
#if OHP_DYNAMIC
        for (;;) {
            unsigned long long ticket = 0;
            if (lane == 0) ticket = atomicAdd(&sm.next_ticket, 1ull);
            const uint64_t it = __shfl_sync(0xffffffffu, ticket, 0);
            if (it >= my_n) break;
#else
        for (uint64_t it = cw; it < my_n; it += kConsumerWarps) {
#endif
            const uint32_t bs = (uint32_t)(it % kBarPairs);
            const uint32_t ph = (uint32_t)(it / kBarPairs) & 1u;
            if constexpr (kDefer) {
                // The previous chunk's slot MUST be handed back before this warp blocks: the loader reclaims in order, and
                // the chunk waited for may be the one that needs that very slot.
                if (pending != kNoPending) {
                    const bool ready = __all_sync(0xffffffffu, mbar_test(smem_u32(&sm.full[bs]), ph));
                    if (!ready) {
                        if (lane == 0) { tma_wait_read<0>(); mbar_arrive(smem_u32(&sm.empty[pending])); }
                        pending = kNoPending;
                    }
                }
            }
#if OHP_CONSUMER_POLL == 0
            OHP_ACC(w_full, mbar_wait(smem_u32(&sm.full[bs]), ph, p.status));
#elif OHP_CONSUMER_POLL == 1
            if (lane == 0) OHP_ACC(w_full, mbar_wait(smem_u32(&sm.full[bs]), ph, p.status));
            __syncwarp();
#else
            if (!__all_sync(0xffffffffu, mbar_test(smem_u32(&sm.full[bs]), ph))) {
                if (lane == 0) OHP_ACC(w_full, mbar_wait(smem_u32(&sm.full[bs]), ph, p.status));
                __syncwarp();
            }
#endif
            if constexpr (kDefer) {
                if (pending != kNoPending) { // the data was there already: the store had the ticket fetch to drain in
                    if (lane == 0) { tma_wait_read<0>(); mbar_arrive(smem_u32(&sm.empty[pending])); }
                    pending = kNoPending;
                }
            }
            const uint32_t sl = (uint32_t)(it & (kRecSlots - 1));
            const ChunkRec& cr = sm.rec[sl];
            const uint32_t kind = cr.kind;
            if (kind == kPcm || kind == kSilenceConv) {
                uint8_t* dst = reinterpret_cast<uint8_t*>((uint64_t)cr.dst_lo | ((uint64_t)cr.dst_hi << 32));
                const uint32_t head = cr.head;
                const uint32_t in_addr = ring + sm.ring_off[sl] + kSlotFront;          // 16-byte aligned; image at +head
                // where the finished image will start: a transform writes it back from the slot's 16-byte boundary, a
                // verbatim chunk stays where it landed (the store takes an image at any byte address)
                uint32_t out_addr = in_addr;
                const uint32_t fmt = cr.out_fmt;
                const uint32_t B = (cr.variant & 3u) + 1u;
                if (kind == kSilenceConv) {
                    silence_to_smem(in_addr, cr.bytes, cr.channels, lane);
                    __syncwarp();
                }
#ifdef OHP_PROFILE_WAITS
                const long long tx0 = clock64();
#endif
                if (fmt <= OHP_OUT_PACKED_LE) {
                    if (cr.mode & kModeTransform) {
                        switch (cr.variant & 3u) {
                        case 0: transform_dispatch<1>(cr, table, in_addr, lane); break;
                        case 1: transform_dispatch<2>(cr, table, in_addr, lane); break;
                        case 2: transform_dispatch<3>(cr, table, in_addr, lane); break;
                        default: transform_dispatch<4>(cr, table, in_addr, lane); break;
                        }
                    } else {
                        out_addr = in_addr + head; // Msg.cpp:2782-2784: the bytes as they are
                    }
                } else if (fmt == OHP_OUT_PLANAR32_BE) {
                    convert_planar32(cr, table, in_addr, dst, cr.aux * 4u, lane);
                } else if (fmt == OHP_OUT_FROM32_BE) {
                    switch (cr.aux) {
                    case 8: convert_from32<1>(cr, table, in_addr, out_addr, lane); break;
                    case 16: convert_from32<2>(cr, table, in_addr, out_addr, lane); break;
                    case 24: convert_from32<3>(cr, table, in_addr, out_addr, lane); break;
                    default: convert_from32<4>(cr, table, in_addr, out_addr, lane); break;
                    }
                } else {
                    const uint32_t db = B < 3 ? B : 3u;
                    if (cr.channels >= 2) {
                        if (db == 1) convert_songcast<1, 2>(cr, table, in_addr, out_addr, cr.aux, lane);
                        else if (db == 2) convert_songcast<2, 2>(cr, table, in_addr, out_addr, cr.aux, lane);
                        else convert_songcast<3, 2>(cr, table, in_addr, out_addr, cr.aux, lane);
                    } else {
                        if (db == 1) convert_songcast<1, 1>(cr, table, in_addr, out_addr, cr.aux, lane);
                        else if (db == 2) convert_songcast<2, 1>(cr, table, in_addr, out_addr, cr.aux, lane);
                        else convert_songcast<3, 1>(cr, table, in_addr, out_addr, cr.aux, lane);
                    }
                }
#ifdef OHP_PROFILE_WAITS
                const long long tx1 = clock64();
                w_xform += tx1 - tx0;
#endif
                fence_proxy_async(); // this lane's shared-memory writes -> visible to the TMA store
                __syncwarp();
#ifdef OHP_PROFILE_WAITS
                const long long tx2 = clock64();
                w_fence += tx2 - tx1;
#endif
                // the finished image sits at out_addr: one TMA bulk store for its 16-byte aligned interior when out_addr is
                // congruent to dst mod 16, 128-bit loads cut to the destination's alignment + streaming stores otherwise.
                // The planar sink has already written global memory itself.
                if (fmt != OHP_OUT_PLANAR32_BE && cr.out_bytes != 0) {
                    // (the packed little-endian sink to the letter keeps the tail of a ramped image only)
                    const uint32_t skip = fmt == OHP_OUT_PACKED_LE ? cr.bytes - cr.out_bytes : 0u;
                    store_image_warp(out_addr + skip, dst, cr.out_bytes, lane);
                }
                __syncwarp();
                if (lane == 0) {
                    tma_commit();
#ifdef OHP_PROFILE_WAITS
                    const long long ts = clock64();
                    w_issue += ts - tx2;
#endif
                    if constexpr (!kDefer) {
                        tma_wait_read<0>(); // the slot can be reused once the bulk store has READ it
#ifdef OHP_PROFILE_WAITS
                        w_store += clock64() - ts;
#endif
                        mbar_arrive(smem_u32(&sm.empty[bs]));
                    }
                }
                if constexpr (kDefer) pending = bs;
                __syncwarp();
            } else {
                if (kind == kSilence) {
                    uint8_t* dst = reinterpret_cast<uint8_t*>((uint64_t)cr.dst_lo | ((uint64_t)cr.dst_hi << 32));
                    write_silence(dst, cr.bytes, cr.channels, (cr.variant & 3u) + 1u, lane);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&sm.empty[bs]));
            }
        }
        if constexpr (kDefer) {
            if (pending != kNoPending && lane == 0) { tma_wait_read<0>(); mbar_arrive(smem_u32(&sm.empty[pending])); }
        }
        if (lane == 0) {
            tma_wait_all<0>(); // every bulk store complete before the CTA (and its shared memory) goes away
            if (cw == 0) { OHP_FLUSH(5, w_full); OHP_FLUSH(6, w_store); OHP_FLUSH(9, w_xform); OHP_FLUSH(10, w_fence); OHP_FLUSH(11, w_issue); }
        }
    }
}

// Per-stream checksum: sum_i (byte_i + 1) * (i + 1) mod 2^64 over [off[s * stride], off[s * stride + 1]): stride 1 for
// streams that tile the arena (n_streams + 1 offsets), 2 for a (begin, end) pair per stream.  One CTA per stream.
__global__ void __launch_bounds__(256) checksum_kernel(const uint8_t* __restrict__ out, const uint64_t* __restrict__ off,
                                                       uint64_t n_streams, uint64_t* __restrict__ sums, uint32_t stride)
{
    __shared__ uint64_t s_part[8];
    for (uint64_t s = blockIdx.x; s < n_streams; s += gridDim.x) {
        const uint64_t lo = off[s * stride], hi = off[s * stride + 1];
        const uint8_t* base = out + lo;
        const uint64_t n = hi - lo;
        uint64_t acc = 0;
        // 16-byte aligned middle with 128-bit loads, ragged edges bytewise
        const uint64_t lead0 = (16u - (reinterpret_cast<uint64_t>(base) & 15u)) & 15u;
        const uint64_t lead = lead0 < n ? lead0 : n;
        const uint64_t words = (n - lead) >> 4;
        const uint64_t tail_at = lead + (words << 4);
        if (threadIdx.x < lead) acc += ((uint64_t)base[threadIdx.x] + 1u) * (threadIdx.x + 1u);
        if (threadIdx.x < n - tail_at) acc += ((uint64_t)base[tail_at + threadIdx.x] + 1u) * (tail_at + threadIdx.x + 1u);
        const uint4* b4 = reinterpret_cast<const uint4*>(base + lead);
        for (uint64_t w = threadIdx.x; w < words; w += blockDim.x) {
            const uint4 v = b4[w];
            const uint32_t t[4] = {v.x, v.y, v.z, v.w};
            const uint64_t i0 = lead + (w << 4) + 1u; // 1-based index of the word's first byte
            uint32_t sum = 0, wsum = 0;               // sum of (byte+1), sum of k*(byte+1) for k=0..15
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const uint32_t b = ((t[k >> 2] >> (8 * (k & 3))) & 0xffu) + 1u;
                sum += b;
                wsum += b * (uint32_t)k;
            }
            acc += i0 * sum + wsum;
        }
        // block reduce
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint64_t total = 0;
            for (unsigned i = 0; i < (blockDim.x >> 5); i++) total += s_part[i];
            sums[s] = total;
        }
        __syncthreads();
    }
}

// Synthetic PCM for benchmarks and tests: stream s's bytes are the splitmix64 sequence (public-domain constants) seeded
// with seed_base | (first_stream_id + s), eight bytes per step, little-endian -- a function of the stream's global id
// only, so a stream reads the same wherever it is sharded to, and the CPU reference arm can generate the same bytes.
// One CTA per stream (grid-stride); a thread writes 16 bytes at a time.
__global__ void __launch_bounds__(256) fill_streams_kernel(uint8_t* __restrict__ in, const ohp_stream_spec* __restrict__ streams,
                                                           uint64_t n_streams, uint64_t seed_base, uint64_t first_stream_id)
{
    for (uint64_t s = blockIdx.x; s < n_streams; s += gridDim.x) {
        const ohp_stream_spec sp = streams[s];
        const uint64_t bytes = sp.total_frames * (uint64_t)(sp.channels * (sp.bit_depth >> 3));
        uint8_t* dst = in + sp.src_base;
        const uint64_t seed = seed_base | (first_stream_id + s);
        auto word = [seed](uint64_t j) -> uint64_t { // the j-th 8-byte step of the sequence
            uint64_t z = seed + (j + 1u) * 0x9E3779B97F4A7C15ull;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            return z ^ (z >> 31);
        };
        const bool vec = (reinterpret_cast<uint64_t>(dst) & 15u) == 0;
        const uint64_t pairs = vec ? bytes >> 4 : 0;
        for (uint64_t v = threadIdx.x; v < pairs; v += blockDim.x) {
            const uint64_t a = word(2 * v), b = word(2 * v + 1);
            reinterpret_cast<uint4*>(dst)[v] = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
        }
        for (uint64_t i = (pairs << 4) + threadIdx.x; i < bytes; i += blockDim.x) {
            dst[i] = (uint8_t)(word(i >> 3) >> (8u * (i & 7u)));
        }
    }
}

} // namespace ohp
