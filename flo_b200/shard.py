"""Multi-GPU sharding of a track batch.

The encode path has no exchange step: tracks are independent (encoder.rs:32-45 is a pure function of
one track), so every rank encodes a contiguous range of tracks with the full single-GPU pipeline.  The
only data that crosses ranks is one byte length per track, for placing the file images in a
concatenated corpus (archive) -- an all_gather of int64, never sample or bitstream data."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def frames_of_track(n_interleaved: int, sample_rate: int, channels: int) -> int:
    """encode_frames, encoder.rs:47-51: ceil((len / channels) / sample_rate)."""
    total = n_interleaved // channels
    return -(-total // sample_rate)


def partition_tracks(frame_counts: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Contiguous track ranges [start, end) per rank, balanced by frame count (the unit of device work).

    Greedy on the running prefix: rank r ends at the first track whose cumulative frame count reaches
    (r + 1) / world of the total.  Deterministic and identical on every rank."""
    n = len(frame_counts)
    total = sum(frame_counts)
    out, start, acc = [], 0, 0
    for r in range(world):
        if r == world - 1:
            end = n
        else:
            target = total * (r + 1) / world
            end = start
            while end < n and (acc + frame_counts[end] <= target or end == start and acc < target and n - end > world - 1 - r):
                acc += frame_counts[end]
                end += 1
            end = min(end, n)
        out.append((start, end))
        start = end
    return out


class LengthExchange:
    """Handle of an asynchronous exchange_lengths_async(); result() blocks and returns (all_lens, offsets)."""

    def __init__(self, work, gathered, ranges):
        self._work, self._gathered, self._ranges = work, gathered, ranges

    def result(self):
        if self._work is not None:
            self._work.wait()
        world = len(self._ranges)
        g = self._gathered.cpu().view(world, -1)
        all_lens = []
        for r, (s, e) in enumerate(self._ranges):
            all_lens.extend(int(v) for v in g[r, :e - s])
        offsets, pos = [], 0
        for v in all_lens:
            offsets.append(pos)
            pos += v
        return all_lens, offsets


def exchange_lengths_async(local_lens: Sequence[int], ranges: Sequence[Tuple[int, int]], group=None) -> LengthExchange:
    """Non-blocking form of exchange_lengths (needs an initialised process group): the all_gather is only
    waited for when the concatenation offsets are needed, so it never stalls the encode stream."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    width = max(1, max((e - s) for s, e in ranges))
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.zeros(width, dtype=torch.int64)
    if len(local_lens):
        mine[:len(local_lens)] = torch.as_tensor([int(v) for v in local_lens], dtype=torch.int64)
    mine = mine.to(dev, non_blocking=True)
    gathered = torch.empty(world * width, dtype=torch.int64, device=dev)
    work = dist.all_gather_into_tensor(gathered, mine, group=group, async_op=True)
    return LengthExchange(work, gathered, list(ranges))


def exchange_lengths(local_lens: Sequence[int], ranges: Sequence[Tuple[int, int]], rank: int, group=None):
    """All ranks learn every track's byte length; returns (all_lens, offsets) where offsets[i] is the
    position of track i's file image in the rank-ordered concatenation.  Uses torch.distributed
    (nccl on GPUs, gloo in the CPU tests); with no process group it is the identity."""
    import torch
    import torch.distributed as dist

    n_total = ranges[-1][1] if ranges else 0
    if not (dist.is_available() and dist.is_initialized()):
        all_lens = list(int(v) for v in local_lens)
    else:
        world = dist.get_world_size(group)
        width = max((e - s) for s, e in ranges) if ranges else 0
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        mine = torch.zeros(max(width, 1), dtype=torch.int64, device=dev)
        if len(local_lens):
            mine[:len(local_lens)] = torch.as_tensor([int(v) for v in local_lens], dtype=torch.int64, device=dev)
        gathered = torch.empty(world * max(width, 1), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gathered, mine, group=group)
        g = gathered.cpu().view(world, -1)
        all_lens = []
        for r, (s, e) in enumerate(ranges):
            all_lens.extend(int(v) for v in g[r, :e - s])
    assert len(all_lens) == n_total or not (dist.is_available() and dist.is_initialized())
    offsets, pos = [], 0
    for v in all_lens:
        offsets.append(pos)
        pos += v
    return all_lens, offsets
