"""Multi-GPU sharding of a batch of independent units.

Every unit (snappy raw block / LZ4 block / frame) is independent, so the batch is cut into contiguous
ranges of the unit index space, one per rank, balanced by uncompressed bytes.  No collective is needed
when each rank generates or already holds its range (bench.py's mode).  For the "batch lives on rank 0"
mode, a `ShardPlan` (one descriptor broadcast, made before the timed path) says which contiguous span of rank 0's arena
every rank decodes; `scatter_payload` / `gather_payload` then move those spans as one batched point-to-point group over
torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests) — a trivial exchange, no reduction, no host round trip.
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


class ShardPlan:
    """Everything the payload moves need, known on every rank before the timed path: the unit ranges per rank and every
    unit's offset / length inside the source rank's input arena and inside the gathered output."""

    def __init__(self, ranges, in_off, in_len, out_len, src_rank):
        self.ranges, self.src_rank = ranges, src_rank
        self.in_off, self.in_len, self.out_len = in_off, in_len, out_len
        n = len(out_len)
        self.out_off = np.zeros(n, dtype=np.uint64)
        if n:
            self.out_off[1:] = np.cumsum(out_len[:-1])

    def count(self, r):
        return self.ranges[r][1] - self.ranges[r][0]

    def in_span(self, r):
        s, e = self.ranges[r]
        return (0, 0) if s == e else (int(self.in_off[s]), int(self.in_off[e - 1] + self.in_len[e - 1]))

    def in_bytes(self, r):
        a, b = self.in_span(r)
        return b - a

    def out_span(self, r):
        s, e = self.ranges[r]
        return (0, 0) if s == e else (int(self.out_off[s]), int(self.out_off[e - 1] + self.out_len[e - 1]))

    def local_offsets(self, r):
        s, e = self.ranges[r]
        return self.in_off[s:e] - np.uint64(self.in_span(r)[0])

    def local_lengths(self, r):
        s, e = self.ranges[r]
        return self.in_len[s:e]


def make_plan(offsets, lengths, out_lengths, weights, src_rank: int = 0, device=None) -> ShardPlan:
    """One broadcast of a packed int64 tensor [n | offsets | lengths | out_lengths] from `src_rank` (the other ranks pass
    None for the three arrays); `weights` (same on every rank, e.g. uncompressed bytes per unit) decides the ranges."""
    import torch
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    device = device if device is not None else torch.device("cpu")
    n = len(weights)
    desc = torch.zeros(3 * n, dtype=torch.int64, device=device)
    if rank == src_rank:
        host = np.concatenate([np.ascontiguousarray(a, dtype=np.uint64).view(np.int64) for a in (offsets, lengths, out_lengths)])
        desc.copy_(torch.from_numpy(host))
    dist.broadcast(desc, src_rank)
    h = desc.cpu().numpy().view(np.uint64)
    return ShardPlan(partition_units(weights, world), h[:n].copy(), h[n:2 * n].copy(), h[2 * n:].copy(), src_rank)


def scatter_payload(plan: ShardPlan, payload, local):
    """Moves every rank's contiguous span of the source arena into its `local` tensor: all sends and receives of the step
    are posted as one batched point-to-point group (one NCCL group on GPUs), nothing touches the host."""
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    ops = []
    if rank == plan.src_rank:
        for r in range(world):
            a, b = plan.in_span(r)
            if b <= a:
                continue
            if r == rank:
                local[: b - a].copy_(payload[a:b])
            else:
                ops.append(dist.P2POp(dist.isend, payload[a:b], r))
    else:
        a, b = plan.in_span(rank)
        if b > a:
            ops.append(dist.P2POp(dist.irecv, local[: b - a], plan.src_rank))
    if ops:
        for q in dist.batch_isend_irecv(ops):
            q.wait()


def gather_payload(plan: ShardPlan, local_out, gathered):
    """Inverse move for the outputs (their lengths are part of the plan): every rank's dense output lands at its place in
    `gathered` on the source rank, again as one batched group."""
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    ops = []
    if rank == plan.src_rank:
        for r in range(world):
            a, b = plan.out_span(r)
            if b <= a:
                continue
            if r == rank:
                gathered[a:b].copy_(local_out[: b - a])
            else:
                ops.append(dist.P2POp(dist.irecv, gathered[a:b], r))
    else:
        a, b = plan.out_span(rank)
        if b > a:
            ops.append(dist.P2POp(dist.isend, local_out[: b - a], plan.src_rank))
    if ops:
        for q in dist.batch_isend_irecv(ops):
            q.wait()


def run_sharded(payload, offsets, lengths, out_lengths, codec_fn: Callable, src_rank: int = 0, device=None):
    """plan -> scatter -> per-rank codec_fn(local_payload, local_offsets, local_lengths, local_out_lengths) -> gather.
    `out_lengths` (known on every rank: for decompression the units' uncompressed sizes) also balances the ranges;
    codec_fn returns this rank's densely packed output tensor.  Returns (gathered, out_offsets, out_lengths) on
    `src_rank`, None elsewhere."""
    import torch
    dist = _dist()
    rank = dist.get_rank()
    device = device if device is not None else torch.device("cpu")
    out_lengths = np.ascontiguousarray(out_lengths, dtype=np.uint64)
    plan = make_plan(offsets, lengths, out_lengths, out_lengths, src_rank, device)
    local = torch.empty(max(plan.in_bytes(rank), 1), dtype=torch.uint8, device=device)
    scatter_payload(plan, payload, local)
    s, e = plan.ranges[rank]
    out = codec_fn(local[: plan.in_bytes(rank)], plan.local_offsets(rank), plan.local_lengths(rank), out_lengths[s:e])
    gathered = torch.empty(int(out_lengths.sum()), dtype=torch.uint8, device=device) if rank == src_rank else None
    gather_payload(plan, out, gathered)
    return (gathered, plan.out_off, plan.out_len) if rank == src_rank else None
