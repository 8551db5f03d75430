/*
 * ohp_multi.h -- the whole stage for a batch of streams over several B200s from ONE process.
 *
 * Streams are fully independent -- no cross-stream state anywhere on the path (Msg.cpp: every MsgAudioPcm carries its own
 * ramp, every MsgPlayablePcm::Read its own RampApplicator) -- so a batch shards by stream and the data path has no
 * collective.  The layout is 8e's: a contiguous block of streams per device (device g of G takes streams
 * [S*g/G, S*(g+1)/G), the rule of ohpipeline_b200/sharding.py), each device with its own context (input and output arenas,
 * descriptors built on that device, ramp table), ONE HOST THREAD AND ITS OWN CUDA STREAMS PER DEVICE; afterwards a
 * host-side gather of per-stream 64-bit checksums (8 bytes per stream D2H per device), the only thing the devices'
 * results are ever put together for.
 *
 * The torchrun-style alternative -- one process per GPU, each calling ohp_run_streams_host on its shard (bench.py
 * --gpus N) -- computes the same bytes; this is the form for a host program that owns all the GPUs itself, as an
 * ohPipeline host process would.
 *
 * Strict C99.  Threading: one call at a time per ohp_multi.
 */
#ifndef OHP_MULTI_H
#define OHP_MULTI_H

#include "ohp_b200.h"
#include "ohp_schedule.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ohp_multi ohp_multi;

/*
 * One context and one host thread per entry of `devices` (CUDA ordinals; the same ordinal may appear more than once: two
 * contexts sharing a GPU, which is how the sharding is tested on a one-GPU box).  Each thread binds itself to the cores of
 * its GPU's NUMA node where the system tells which they are.  OHP_E_NO_DEVICE / OHP_E_CUDA as ohp_create, with the text
 * in ohp_last_error(NULL).
 */
int    ohp_multi_create(const int* devices, size_t n_devices, ohp_multi** out);
int    ohp_multi_destroy(ohp_multi* m);
size_t ohp_multi_num_devices(const ohp_multi* m);
const char* ohp_multi_last_error(const ohp_multi* m);

/* The block of streams device `index` of n_devices takes out of n_streams: [*first, *first + *count). */
void   ohp_multi_shard(size_t n_streams, size_t n_devices, size_t index, size_t* first, size_t* count);

/*
 * ohp_run_streams_host (ohp_schedule_device.h) over all the devices at once: every device thread takes its block of
 * streams, re-based so that its device arenas hold that block only, moves the block's PCM host -> device, builds the
 * descriptors there, runs the ramp + convert kernel and moves the block's output back, all devices concurrently.
 * h_in / h_out as there (pinned -- ohp_multi_host_alloc -- for full PCIe speed); streams may lie anywhere in the arenas.
 *   h_stream_out_bytes (n_streams, may be NULL): bytes each stream produced at h_out + dst_base.
 *   h_checksums (n_streams, may be NULL): per stream, SUM_i (byte_i + 1) * (i + 1) mod 2^64 over the bytes it produced
 *     (the sum ohp_checksums_device defines), computed on the device that produced them from what sits in ITS memory
 *     and gathered on the host.
 *   total_chunks (may be NULL): playables read, all devices.
 * Synchronous; on any error nothing of the call is still in flight on any device when it returns.  The status is that of
 * the first device (in `devices` order) that failed, its message prefixed with the device's index.
 */
int    ohp_multi_run_streams_host(ohp_multi* m, const ohp_stream_spec* h_streams, size_t n_streams,
                                  const ohp_ramp_event* h_events, size_t n_events,
                                  const uint8_t* h_in, uint64_t in_bytes, uint8_t* h_out, uint64_t out_bytes,
                                  uint64_t* h_stream_out_bytes, uint64_t* h_checksums, uint64_t* total_chunks);

/* Pinned host memory every device of `m` can DMA from and to (unified addressing makes one allocation do for all). */
int    ohp_multi_host_alloc(ohp_multi* m, uint64_t bytes, void** out_hptr);
int    ohp_multi_host_free(ohp_multi* m, void* hptr);

/* The context of device `index` (owned by `m`; for ohp_inflight_cap, ohp_last_error and the like -- not for launches
 * while an ohp_multi call is running). */
ohp_context* ohp_multi_context(ohp_multi* m, size_t index);

#ifdef __cplusplus
}
#endif

#endif /* OHP_MULTI_H */
