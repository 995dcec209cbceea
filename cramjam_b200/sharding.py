"""Multi-GPU sharding of a batch of independent units.

Every unit (snappy raw block / LZ4 block / frame) is independent, so the batch is cut into contiguous
ranges of the unit index space, one per rank, balanced by uncompressed bytes.  No collective is needed
when each rank generates or already holds its range (bench.py's mode).  For the "batch lives on rank 0"
mode, `scatter_units` / `gather_units` move the ranges with point-to-point sends over torch.distributed
(NCCL over NVLink on GPUs; gloo in the CPU tests) — a trivial exchange, no reduction.
"""
from typing import Callable, List, Sequence, Tuple

import numpy as np


def partition_units(weights: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous [start, end) unit ranges per rank with ~equal total weight (e.g. uncompressed bytes).
    Always returns world_size ranges; ranges may be empty when there are fewer units than ranks."""
    w = np.asarray(weights, dtype=np.float64)
    n = len(w)
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    if n == 0:
        return [(0, 0)] * world_size
    csum = np.concatenate([[0.0], np.cumsum(w)])
    total = csum[-1]
    cuts = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        k = int(np.searchsorted(csum, target, side="left"))
        # pick the boundary closer to the ideal cut, never moving backwards
        if k > 0 and abs(csum[k - 1] - target) <= abs(csum[min(k, n)] - target):
            k -= 1
        cuts.append(min(max(k, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def _dist():
    import torch.distributed as dist
    return dist


def scatter_units(payload, offsets: np.ndarray, lengths: np.ndarray, ranges: List[Tuple[int, int]], src_rank: int = 0, device=None):
    """Rank `src_rank` holds `payload` (1-D uint8 tensor) with unit i at [offsets[i], offsets[i]+lengths[i]).
    Returns (local_payload, local_offsets, local_lengths) for this rank's range.  Descriptors travel as one
    broadcast; payload ranges as point-to-point sends (ranges are contiguous spans of the arena)."""
    import torch
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    device = device if device is not None else (payload.device if payload is not None else torch.device("cpu"))
    n_t = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == src_rank:
        n_t[0] = len(offsets)
    dist.broadcast(n_t, src_rank)
    n = int(n_t.item())
    desc = torch.zeros(2 * n, dtype=torch.int64, device=device)
    if rank == src_rank:
        desc[:n] = torch.from_numpy(np.ascontiguousarray(offsets, dtype=np.int64)).to(device)
        desc[n:] = torch.from_numpy(np.ascontiguousarray(lengths, dtype=np.int64)).to(device)
    dist.broadcast(desc, src_rank)
    off = desc[:n].cpu().numpy().astype(np.uint64)
    ln = desc[n:].cpu().numpy().astype(np.uint64)

    def span(r):
        s, e = ranges[r]
        if s == e:
            return 0, 0
        return int(off[s]), int(off[e - 1] + ln[e - 1])

    s, e = ranges[rank]
    lo, hi = span(rank)
    local = torch.empty(hi - lo, dtype=torch.uint8, device=device)
    if rank == src_rank:
        reqs = []
        for r in range(world):
            a, b = span(r)
            if r == rank:
                local.copy_(payload[a:b])
            elif b > a:
                reqs.append(dist.isend(payload[a:b].contiguous(), r))
        for q in reqs:
            q.wait()
    elif hi > lo:
        dist.recv(local, src_rank)
    return local, (off[s:e] - np.uint64(lo)), ln[s:e]


def gather_units(local_out, local_lengths: np.ndarray, ranges: List[Tuple[int, int]], dst_rank: int = 0):
    """Inverse of scatter_units for the outputs: every rank sends its densely packed output bytes to `dst_rank`,
    which returns (payload, offsets, lengths) in global unit order; other ranks return None."""
    import torch
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    device = local_out.device
    n_total = ranges[-1][1]
    lens = torch.zeros(n_total, dtype=torch.int64, device=device)
    s, e = ranges[rank]
    if e > s:
        lens[s:e] = torch.from_numpy(np.ascontiguousarray(local_lengths, dtype=np.int64)).to(device)
    dist.all_reduce(lens)  # disjoint ranges: the sum is the concatenation
    ln = lens.cpu().numpy().astype(np.uint64)
    off = np.zeros(n_total, dtype=np.uint64)
    if n_total:
        off[1:] = np.cumsum(ln[:-1])
    total = int(ln.sum())
    sizes = [int(ln[a:b].sum()) for a, b in ranges]
    if rank == dst_rank:
        out = torch.empty(total, dtype=torch.uint8, device=device)
        pos = [int(off[a]) if b > a else 0 for a, b in ranges]
        reqs = []
        for r in range(world):
            if sizes[r] == 0:
                continue
            if r == rank:
                out[pos[r]:pos[r] + sizes[r]].copy_(local_out[:sizes[r]])
            else:
                reqs.append(dist.irecv(out[pos[r]:pos[r] + sizes[r]], r))
        for q in reqs:
            q.wait()
        return out, off, ln
    if sizes[rank]:
        dist.send(local_out[:sizes[rank]].contiguous(), dst_rank)
    return None


def run_sharded(payload, offsets, lengths, weights, codec_fn: Callable, src_rank: int = 0, device=None):
    """scatter -> per-rank codec_fn(local_payload, local_offsets, local_lengths) -> gather.
    codec_fn returns (dense_output_tensor, output_lengths)."""
    dist = _dist()
    ranges = partition_units(weights, dist.get_world_size())
    local, loff, llen = scatter_units(payload, offsets, lengths, ranges, src_rank, device)
    out, out_len = codec_fn(local, loff, llen)
    return gather_units(out, out_len, ranges, src_rank)
