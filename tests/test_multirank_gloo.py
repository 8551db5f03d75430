"""The N > 1 path on CPU: two gloo ranks shard a batch of streams, each runs its shard (through the C oracle here, the
CUDA path on the GPU box), and the host-side checksum gather reproduces the single-process result."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from ohpipeline_b200 import capi, sharding, workloads as W


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _stream_sums(port_lib, streams, events, pcm_of_stream):
    """Per-stream checksum of the shard's output (descriptors from the product's host code, audio from the oracle)."""
    sched = capi.schedule_build(streams, events, threads=1)
    in_bytes = int(streams["src_base"][-1]) + len(pcm_of_stream[-1]) + 64
    inp = np.zeros(in_bytes, dtype=np.uint8)
    for s, pcm in zip(streams, pcm_of_stream):
        inp[int(s["src_base"]):int(s["src_base"]) + len(pcm)] = pcm
    out_bytes = int(streams["dst_base"][-1] + sched.stream_out_bytes[-1]) + 64
    rc, out = port_lib.process_chunks(sched.chunks, inp, out_bytes)
    assert rc == 0
    sums = np.zeros(len(streams), dtype=np.uint64)
    for i, s in enumerate(streams):
        lo = int(s["dst_base"])
        sums[i] = port_lib.checksum(out[lo:lo + int(sched.stream_out_bytes[i])])
    return sums


def _pcm(port_lib, w, idx):
    s = w.streams[idx]
    n = int(s["total_frames"]) * int(s["channels"]) * int(s["bit_depth"]) // 8
    return port_lib.fill_pcm(n, (w.seed << 8) + idx)   # seeded by GLOBAL stream id


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import pyoracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    port_lib = pyoracle.Port()
    w = W.mixed(n_streams=20, seed=9, max_frames=800)
    sub, events, lo, hi = sharding.shard_workload(w.streams, w.events, world, rank)
    sums = _stream_sums(port_lib, sub, events, [_pcm(port_lib, w, i) for i in range(lo, hi)])
    allsums = sharding.gather_checksums(sums, len(w.streams), world, rank, dist)
    if rank == 0:
        q.put(allsums.tobytes())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_checksum_gather():
    from oracle import pyoracle
    port_lib = pyoracle.Port()
    w = W.mixed(n_streams=20, seed=9, max_frames=800)
    single = _stream_sums(port_lib, w.streams.copy(), w.events, [_pcm(port_lib, w, i) for i in range(len(w.streams))])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = np.frombuffer(q.get(timeout=120), dtype=np.uint64)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(got, single)


def test_shard_ranges_tile_the_batch():
    for n in (1, 7, 1024, 65536):
        for world in (1, 2, 4, 8):
            edges = [sharding.shard_range(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
