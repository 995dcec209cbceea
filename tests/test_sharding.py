"""Multi-GPU plumbing on CPU: partition arithmetic, and the scatter -> codec -> gather path over
torch.distributed with the gloo backend at world_size 2 (the N>1 path of SURVEY.md §8e).  The codec
step is the CPU oracle here (test infrastructure); on GPUs it is cj_decompress_batch."""
import os
import socket
import sys

import numpy as np
import pytest

from cramjam_b200.sharding import partition_units

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_covers_everything_contiguously():
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 7, 100, 65536):
        for world in (1, 2, 3, 8):
            w = rng.integers(1, 100000, size=n)
            r = partition_units(w, world)
            assert len(r) == world and r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert all(a <= b for a, b in r)


def test_partition_balances_bytes():
    w = np.full(65536, 65536)
    r = partition_units(w, 8)
    assert [b - a for a, b in r] == [8192] * 8
    rng = np.random.default_rng(1)
    w = rng.integers(1000, 70000, size=10000)
    r = partition_units(w, 4)
    loads = [w[a:b].sum() for a, b in r]
    assert max(loads) - min(loads) <= 2 * w.max()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import torch
    import torch.distributed as dist
    import oracle as O
    from cramjam_b200 import _capi as capi
    from cramjam_b200.sharding import run_sharded
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, U = 37, 65536
        payload = offsets = lengths = weights = None
        data = capi.synth_host(n, U, seed=9)
        if rank == 0:
            blocks = [O.snappy_raw_compress(data[i * U:(i + 1) * U].tobytes()) for i in range(n)]
            lengths = np.array([len(b) for b in blocks], dtype=np.uint64)
            offsets = np.zeros(n, dtype=np.uint64); offsets[1:] = np.cumsum(lengths[:-1])
            payload = torch.from_numpy(np.frombuffer(b"".join(blocks), dtype=np.uint8).copy())
        weights = np.full(n, U)

        def codec(local, loff, llen, lout):
            src = local.numpy()
            k = len(loff)
            assert (lout == U).all()
            dst = np.zeros(k * U, dtype=np.uint8)
            do = np.arange(k, dtype=np.uint64) * U
            out_len, _ = O.batch(O.SNAPPY_RAW, 0, src if src.size else np.zeros(1, np.uint8), loff, llen, dst if k else np.zeros(1, np.uint8), do,
                                 np.full(k, U, np.uint64), nthreads=2)
            assert (out_len == U).all()
            return torch.from_numpy(dst)

        res = run_sharded(payload, offsets, lengths, weights, codec, src_rank=0, device=torch.device("cpu"))
        if rank == 0:
            out, off, ln = res
            ok = out.numpy().tobytes() == data.tobytes() and int(ln.sum()) == n * U and len(off) == n
            q.put(("ok" if ok else "mismatch"))
        else:
            assert res is None
            q.put("ok")
    finally:
        dist.destroy_process_group()


def test_scatter_codec_gather_gloo_world2():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(30)
    assert results == ["ok", "ok"]
