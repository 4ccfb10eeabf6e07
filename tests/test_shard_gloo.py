"""Host-side multi-GPU logic on CPU: world_size-2 gloo run of the track partition + length exchange
(the only cross-rank traffic of the path).  The oracle stands in for the per-rank encoder
here -- this test checks sharding and concatenation offsets, not the CUDA path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import oracle, pcm16_to_f32, synth_pcm16
from flo_b200 import shard

TRACKS = [(9000, 1, 8000), (20000, 2, 8000), (500, 2, 8000), (0, 1, 8000), (16001, 1, 8000), (8000, 2, 4000), (12345, 1, 8000)]


def _encode(i):
    n, ch, sr = TRACKS[i]
    x = pcm16_to_f32(synth_pcm16(n, ch, sr, seed=100 + i))
    return oracle.encode(x, sr, ch, 16, 5, b"t%d" % i)


def test_partition_is_contiguous_and_balanced():
    fc = [shard.frames_of_track(n * ch, sr, ch) for n, ch, sr in TRACKS]
    assert fc == [2, 3, 1, 0, 3, 2, 2]
    for world in (1, 2, 3, 4, 8, 16):
        r = shard.partition_tracks(fc, world)
        assert len(r) == world and r[0][0] == 0 and r[-1][1] == len(fc)
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        loads = [sum(fc[s:e]) for s, e in r]
        assert sum(loads) == sum(fc)
        if world <= 4:
            assert max(loads) <= -(-sum(fc) // world) + max(fc)
    assert shard.partition_tracks([180] * 10000, 8) == [(1250 * i, 1250 * (i + 1)) for i in range(8)]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fc = [shard.frames_of_track(n * ch, sr, ch) for n, ch, sr in TRACKS]
        ranges = shard.partition_tracks(fc, world)
        s, e = ranges[rank]
        mine = [_encode(i) for i in range(s, e)]
        all_lens, offsets = shard.exchange_lengths([len(b) for b in mine], ranges, rank)
        q.put((rank, all_lens, offsets, [bytes(b) for b in mine], (s, e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_length_exchange_and_concatenation():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [_encode(i) for i in range(len(TRACKS))]
    archive = bytearray(sum(len(w) for w in want))
    for rank, all_lens, offsets, mine, (s, e) in got:
        assert all_lens == [len(w) for w in want]            # every rank knows every length
        assert offsets == list(np.cumsum([0] + all_lens[:-1]))
        for i, b in zip(range(s, e), mine):
            archive[offsets[i]:offsets[i] + len(b)] = b
    assert bytes(archive) == b"".join(want)


def test_exchange_without_process_group_is_identity():
    lens, offs = shard.exchange_lengths([5, 7, 9], [(0, 3)], 0)
    assert lens == [5, 7, 9] and offs == [0, 5, 12]


def _worker_async(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ranges = [(0, 3), (3, 4)]
        s, e = ranges[rank]
        mine = [10 * (i + 1) + rank for i in range(s, e)]
        h = shard.exchange_lengths_async(mine, ranges)
        q.put((rank,) + h.result())
    finally:
        dist.destroy_process_group()


def test_async_length_exchange_gloo():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_async, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, lens, offs in got:
        assert lens == [10, 20, 30, 41] and offs == [0, 10, 30, 60]
