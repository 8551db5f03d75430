"""Multi-GPU layout of the path: streams are independent (no cross-stream state anywhere on the path), so they shard
across ranks with NO data-path collective.  The only exchange is a host-side gather of per-stream 64-bit checksums.

One process per GPU; `torch.distributed` (nccl on the GPU box, gloo in CPU tests) is plumbing for that gather only.
"""
import numpy as np


def shard_range(n_streams, world_size, rank):
    """Contiguous block of streams for `rank` (stream id -> rank id * world // n), sizes differing by at most one."""
    lo = n_streams * rank // world_size
    hi = n_streams * (rank + 1) // world_size
    return lo, hi


def shard_workload(streams, events, world_size, rank):
    """The rank's slice of a STREAM_SPEC array, re-based so that its arenas start at 0.  Events are shared
    (first_event indexes stay valid)."""
    lo, hi = shard_range(len(streams), world_size, rank)
    sub = streams[lo:hi].copy()
    if len(sub):
        sub["src_base"] -= sub["src_base"][0]
        sub["dst_base"] -= sub["dst_base"][0]
    return sub, events, lo, hi


def gather_checksums(local_sums, n_streams, world_size, rank, dist=None):
    """Host-side gather of per-stream checksums to every rank (uint64 array of n_streams)."""
    local_sums = np.ascontiguousarray(local_sums, dtype=np.uint64)
    if dist is None or world_size == 1:
        assert len(local_sums) == n_streams
        return local_sums
    gathered = [None] * world_size
    dist.all_gather_object(gathered, local_sums.tobytes())
    out = np.concatenate([np.frombuffer(b, dtype=np.uint64) for b in gathered])
    assert len(out) == n_streams
    return out
