#!/usr/bin/env python
"""bench.py — headline benchmark of the cramjam_b200 engine (contract: see the task statement).

Workload (BASELINE.json configs[1]): snappy raw block decompress of 65 536 x 64 KiB synthetic "Silesia-like" blocks per
GPU (4 GiB uncompressed out, ~2.2 GiB compressed in).  One step = one pass of the hot path (one batched decode) over the
whole batch.  The compressed streams are made on the HOST by Google snappy (the library `snap`, the reference's encoder,
is a port of) so that both arms decode byte-identical input.

  value         uncompressed GB/s, whole job, inputs and outputs resident in HBM, CUDA-event timed
  e2e           same metric through the C-ABI call with PINNED HOST buffers (H2D + kernel + D2H inside the timed region)
  roofline      algorithmic bytes (compressed read + uncompressed written) / kernel time vs measured HBM copy bandwidth
  cpu_baseline  the fastest CPU implementation in the image (Arrow's bundled Google snappy, "stand-in") and the oracle
                port, all host cores and one core, bounded sample
  extras        per-codec sub-records (decode / encode of snappy, lz4, zstd at the BASELINE.json sizes, each with its
                own roofline numbers), the mixed configs[4] batch, a real-corpus row, the single-buffer Python API

`--impl reference` times only the CPU stand-in on the same blocks (the product library is never loaded on that arm).
N>1 (torchrun): one process per GPU, blocks sharded by global index (rank r owns blocks [r*B, (r+1)*B)), no data-path
collective; weak scaling; time = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

U = 65536
SEED = 0xC0FFEE
METRIC = "snappy raw block decompress, uncompressed GB/s (64 KiB blocks)"
O_SNAPPY, O_LZ4, O_ZSTD = 0, 2, 4   # oracle / stand-in codec ids (oracle/cj_oracle.h)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--blocks", type=int, default=65536, help="64 KiB blocks per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="length of the bounded CPU baseline sample")
    ap.add_argument("--no-extras", action="store_true", help="skip the per-codec extra rows")
    ap.add_argument("--zstd-frames", type=int, default=16384, help="256 KiB frames of the zstd row (BASELINE.json configs[3])")
    ap.add_argument("--mixed-blocks", type=int, default=131072, help="blocks per GPU of the mixed snappy+lz4 row (configs[4])")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(key, blocks):
    """Per-launch DRAM bytes of a kernel from the committed ncu captures (profiles/traffic.json), if taken at this size."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = t.get(key)
        if e and int(e["units"]) == int(blocks):
            return float(e["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def host_threads(world=1):
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, n // max(1, world))


def bind_to_gpu_numa(local_rank):
    """Pins this process to the CPUs of the NUMA node its GPU hangs off, so that pinned staging buffers (first touch) and the
    copy threads are local to the GPU's PCIe root.  Best effort; returns what it did."""
    try:
        import torch
        bdf = torch.cuda.get_device_properties(local_rank)
        bus = f"{bdf.pci_domain_id:04x}:{bdf.pci_bus_id:02x}:{bdf.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return {"numa_node": None}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception as e:
        return {"numa_node": None, "error": repr(e)[:80]}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                mx = float(f[1])
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[0]))
                    for nm, v in zip(names, f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(nm)
            except ValueError:
                continue
        if not sm:  # region shorter than the sampling period: use every sample taken
            for ts, line in self.rows:
                try:
                    sm.append(float(line.split(",")[0]))
                except ValueError:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# host-side workload (oracle/ only: shared by both arms, so that they see byte-identical input)
# ------------------------------------------------------------------------------------------------
def pack16(lens):
    off = np.zeros(len(lens), dtype=np.uint64)
    if len(lens) > 1:
        off[1:] = np.cumsum((lens[:-1] + np.uint64(15)) & ~np.uint64(15))
    return off


def compress_host(ocodec, data, n, unit, nthreads, level=3):
    """Compresses n units of `unit` bytes with the stand-in CPU encoder (oracle encoder if the stand-in is absent) and packs
    the streams into a dense 16-byte aligned arena.  Returns (arena u8, off u64[n], len u64[n], encoder name)."""
    import oracle as O
    bound = {O_SNAPPY: 32 + unit + unit // 6, O_LZ4: unit + unit // 255 + 16, O_ZSTD: unit + unit // 128 + 1024}[ocodec]
    slot = (bound + 15) // 16 * 16
    CH = max(1, min(n, (1 << 30) // slot))   # bounded scratch: one chunk of worst-case slots at a time
    lens = np.zeros(n, dtype=np.uint64)
    parts, enc = [], None
    scratch = np.empty(CH * slot, dtype=np.uint8)
    for a in range(0, n, CH):
        k = min(CH, n - a)
        so = (np.arange(k, dtype=np.uint64) + np.uint64(a)) * np.uint64(unit)
        do = np.arange(k, dtype=np.uint64) * np.uint64(slot)
        ul, cap = np.full(k, unit, np.uint64), np.full(k, slot, np.uint64)
        r = O.standin_batch(ocodec, 1, data, so, ul, scratch, do, cap, nthreads=nthreads, level=level)
        if r is not None and (r[0] > 0).all():
            cl, enc = r[0].astype(np.uint64), "stand-in"
        else:
            if ocodec == O_ZSTD:
                raise RuntimeError("no CPU zstd encoder available")
            cl = O.batch({O_SNAPPY: O.SNAPPY_RAW, O_LZ4: O.LZ4_BLOCK}[ocodec], 1, data, so, ul, scratch, do, cap, nthreads=nthreads)[0]
            assert (cl > 0).all()
            cl, enc = cl.astype(np.uint64), "oracle"
        lens[a:a + k] = cl
        lo = pack16(cl)
        part = np.zeros((int(lo[-1] + cl[-1]) + 15) // 16 * 16, dtype=np.uint8)
        O.pack_units(scratch, do, cl, part, lo, nthreads)
        parts.append(part)
    arena = np.concatenate(parts + [np.zeros(64, dtype=np.uint8)])
    return arena, pack16(lens), lens, enc


def headline_workload(blocks, first_index, nthreads):
    """The configs[1] input: synthetic blocks (oracle/synth.c) and their Google-snappy streams."""
    import oracle as O
    data = O.synth(blocks, U, SEED, first_index, nthreads)
    comp, coff, clen, enc = compress_host(O_SNAPPY, data, blocks, U, nthreads)
    return data, comp, coff, clen, enc


def workload_config(blocks, ratio, comp_bytes, enc):
    streams = ("Google snappy (Arrow's bundled copy; the library the reference's snap 1.1.1 is a port of), made on the host; both arms decode the same bytes"
               if enc == "stand-in" else "oracle/snappy.c encoder, made on the host; both arms decode the same bytes")
    return {"workload": f"snappy raw block decompress, {blocks} x 64 KiB synthetic Silesia-like blocks per GPU (BASELINE.json configs[1])",
            "blocks_per_gpu": blocks, "block_bytes": U, "ratio": round(ratio, 4), "compressed_bytes_per_gpu": int(comp_bytes), "streams": streams,
            "l2": "inputs+outputs per step (~6 GiB) far exceed the 126 MB L2; no flush needed",
            "sharding": "contiguous global block ranges per rank, no data-path collective"}


def cpu_decode_rate(ocodec, comp, coff, clen, unit, n, nthreads, seconds, use_standin, check=None):
    """Throughput (uncompressed GB/s) of the CPU decoder over units [0, n): repeated passes for about `seconds`."""
    import oracle as O
    out = np.empty(n * unit, dtype=np.uint8)
    do = np.arange(n, dtype=np.uint64) * np.uint64(unit)
    cap = np.full(n, unit, np.uint64)
    ocodec_port = {O_SNAPPY: O.SNAPPY_RAW, O_LZ4: O.LZ4_BLOCK, O_ZSTD: O.ZSTD}[ocodec]

    def one():
        if use_standin:
            r = O.standin_batch(ocodec, 0, comp, coff[:n], clen[:n], out, do, cap, nthreads=nthreads)
        else:
            r = O.batch(ocodec_port, 0, comp, coff[:n], clen[:n], out, do, cap, nthreads=nthreads)
        assert r is not None and (r[0] == unit).all()
        return r[1]
    one()   # warm (page faults of `out`)
    if check is not None:
        assert np.array_equal(out, check[: n * unit]), "CPU decode differs from the original blocks"
    t, passes = 0.0, 0
    while t < seconds and passes < 1000:
        t += one()
        passes += 1
    return n * unit * passes / t / 1e9, passes


def run_reference(args):
    """The reference arm: the CPU stand-in decoding the same blocks on all host cores.  Never loads the product library."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as O
    nt = host_threads()
    B = args.blocks
    data, comp, coff, clen, enc = headline_workload(B, 0, nt)
    use_standin = O.standin() is not None
    out = np.empty(B * U, dtype=np.uint8)
    do = np.arange(B, dtype=np.uint64) * np.uint64(U)
    cap = np.full(B, U, np.uint64)

    def one():
        r = (O.standin_batch(O_SNAPPY, 0, comp, coff, clen, out, do, cap, nthreads=nt) if use_standin
             else O.batch(O.SNAPPY_RAW, 0, comp, coff, clen, out, do, cap, nthreads=nt))
        assert (r[0] == U).all()
        return r[1]
    for _ in range(max(args.warmup, 1)):
        one()
    assert np.array_equal(out, data)
    t = sum(one() for _ in range(args.steps))
    ms = 1e3 * t / args.steps
    gbs = B * U / (ms * 1e6)
    what = "Arrow's bundled Google snappy via arrow::util::Codec (oracle/standin.cpp)" if use_standin else "oracle/snappy.c (port)"
    sample = f"all {B} x 64 KiB blocks of one GPU's share per step (the full configs[1] batch, same bytes as the GPU arm), {what}, {nt} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(B, B * U / float(clen.sum()), int(coff[-1] + clen[-1]), enc),
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": nt, "kind": "stand-in" if use_standin else "port", "sample": sample},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from cramjam_b200 import _capi as capi
    import oracle as O   # workload generation + the CPU baseline leg only

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; cramjam_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local) if world > 1 else {"numa_node": None}
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = args.blocks
    nt = host_threads(world)
    ctx = capi.Context(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    i64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).to(dev)
    peak, peak_src = peaks()

    def timed(fn, k=5, warm=2):
        for _ in range(warm):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(k):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / k

    def record(ms, units, unit_bytes, comp_bytes, kernel, **kw):
        """A per-codec sub-record: uncompressed GB/s and the roofline numbers of SURVEY.md 8(d) (algorithmic bytes = compressed +
        uncompressed bytes of the batch, whichever way they flow)."""
        alg = float(comp_bytes) + float(units) * unit_bytes
        r = {"GBps": units * unit_bytes / (ms * 1e6), "ms": ms, "units": units, "unit_bytes": unit_bytes, "ratio": units * unit_bytes / float(comp_bytes),
             "algorithmic_bytes": alg, "achieved_GBps": alg / (ms * 1e6), "frac": alg / (ms * 1e6) / peak, "kernel": kernel}
        r.update(kw)
        return r

    # ---- set-up (untimed): this rank's blocks and their host-made snappy streams; the device generator must agree ----
    data, h_comp_np, coff, clen, enc = headline_workload(B, rank * B, nt)
    comp_bytes = int(coff[-1] + clen[-1])
    comp_span = len(h_comp_np) - 64
    ratio = B * U / float(clen.sum())
    raw = torch.empty(B * U, dtype=torch.uint8, device=dev)
    ctx.synth_device(raw, B, U, SEED, first_index=rank * B)
    h_comp = torch.from_numpy(h_comp_np).pin_memory()
    comp = h_comp.to(dev)
    raw_off = np.arange(B, dtype=np.uint64) * U
    t_raw_off, t_raw_len = i64(raw_off), i64(np.full(B, U, np.uint64))
    t_coff, t_clen = i64(coff), i64(clen)
    t_st = torch.zeros(B, dtype=torch.int32, device=dev)
    out = torch.zeros(B * U, dtype=torch.uint8, device=dev)
    t_dl = torch.zeros(B, dtype=torch.int64, device=dev)

    def step_device():
        ctx.decompress_batch(capi.SNAPPY_RAW, capi.DEVICE, B, comp, t_coff, t_clen, out, t_raw_off, t_raw_len, t_dl, t_st)

    # ---- correctness of the timed path on this exact input (untimed): output == the device generator's blocks, and the
    #      host generator (what the CPU arm decodes to) made the same bytes ----
    step_device()
    ctx.synchronize()
    assert int((t_st != 0).sum()) == 0 and bool(torch.equal(out, raw)), "decode output differs from the original blocks"
    assert bool(torch.equal(raw[: 256 * U].cpu(), torch.from_numpy(data[: 256 * U]))), "host and device generators disagree"
    redo = ctx.last_redo_count()

    # ---- value: device-resident, CUDA events on the launching stream ----
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    wall1 = time.time()
    launches = ctx.launch_count - launches0
    ms_dev = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    clocks = sampler.stop(wall0, wall1) if sampler else None
    total_blocks = B * world
    gen, min_units = ctx.decode_path()
    tpb = f"g{gen}_kernel" if (gen in (4, 7) and B >= min_units + min_units // 2) else None   # the thread-per-block kernel that decodes batches of this size
    kernel_name = f"{tpb}<snappy> (one thread per block) + lz_decode_list_kernel (redo list)" if tpb else "lz_decode_kernel<snappy, lane-parallel> (one warp per block)"
    value = total_blocks * U / (ms_dev * 1e6)
    alg_bytes = float(clen.sum()) + float(B) * U   # per launch on this rank: compressed read + uncompressed written
    achieved = alg_bytes / (ms_dev * 1e6)

    # ---- e2e: pinned host buffers through the same C-ABI call (H2D + kernel + D2H inside) ----
    h_out = torch.empty(B * U, dtype=torch.uint8).pin_memory()
    h_dl = np.zeros(B, dtype=np.uint64)
    h_st = np.zeros(B, dtype=np.int32)
    h_cap = np.full(B, U, np.uint64)

    def step_e2e():
        ctx.decompress_batch(capi.SNAPPY_RAW, capi.PINNED, B, h_comp, coff, clen, h_out, raw_off, h_cap, h_dl, h_st)

    e2e_steps = max(2, min(args.steps, 5))
    step_e2e()
    assert int((h_st != 0).sum()) == 0
    chk = torch.equal(h_out[: 64 * U], raw[: 64 * U].cpu()) and torch.equal(h_out[-64 * U:], raw[-64 * U:].cpu())
    assert chk, "e2e output differs"
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    ms_e2e = max_over_ranks(1e3 * (time.perf_counter() - t0) / e2e_steps)
    e2e_val = total_blocks * U / (ms_e2e * 1e6)
    h2d = comp_span + 32 * B   # payload span + descriptor arrays
    d2h = B * U + 12 * B       # output + dst_len/status arrays
    # what the link can do with exactly these bytes, all ranks at once (the ceiling of e2e): H2D and D2H streams side by side
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def link():
        with torch.cuda.stream(s1):
            comp.copy_(h_comp, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(out, non_blocking=True)
    link()
    barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        link()
    barrier()
    ms_link = max_over_ranks(1e3 * (time.perf_counter() - t0) / 2)
    torch.cuda.set_stream(stream)
    e2e = {"value": e2e_val, "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e, "steps": e2e_steps,
           "path": "cj_decompress_batch(CJ_SNAPPY_RAW, CJ_PINNED) from pinned host arenas",
           "bound": "pcie", "link_only_ms_per_step": ms_link,
           "link_ceiling_GBps": total_blocks * U / (ms_link * 1e6), "frac_of_link_ceiling": ms_link / ms_e2e,
           "note": "link_only = the same H2D + D2H bytes moved by plain cudaMemcpyAsync on two streams, all ranks at once, no kernel", "numa": numa}

    extras = {}
    # ---- extras, every N: the mixed snappy+lz4 batch of BASELINE.json configs[4] at its per-GPU size, generated in place per
    #      rank from the global block index (even global index = snappy, odd = lz4), GPU-encoded, decoded by two batched calls ----
    if not args.no_extras:
        try:
            MB = args.mixed_blocks
            del out
            torch.cuda.empty_cache()
            mraw = torch.empty(MB * U, dtype=torch.uint8, device=dev)
            ctx.synth_device(mraw, MB, U, SEED, first_index=rank * MB)
            half = MB // 2
            bound_s = (32 + U + U // 6 + 15) // 16 * 16
            mslots = torch.empty(half * bound_s, dtype=torch.uint8, device=dev)
            t_slot_off, t_slot_cap = i64(np.arange(half, dtype=np.uint64) * bound_s), i64(np.full(half, bound_s, np.uint64))
            t_hl = i64(np.full(half, U, np.uint64))
            parts = {}
            for name, codec, par in (("snappy", capi.SNAPPY_RAW, 0), ("lz4", capi.LZ4_BLOCK, 1)):
                idx = np.arange(par, MB, 2, dtype=np.uint64)
                t_ro = i64(idx * np.uint64(U))
                t_cl = torch.zeros(half, dtype=torch.int64, device=dev)
                t_s = torch.zeros(half, dtype=torch.int32, device=dev)
                ctx.compress_batch(codec, capi.DEVICE, half, mraw, t_ro, t_hl, mslots, t_slot_off, t_slot_cap, t_cl, t_s)
                ctx.synchronize()
                assert int((t_s != 0).sum()) == 0
                cl = t_cl.cpu().numpy().astype(np.uint64)
                co = pack16(cl)
                arena = torch.zeros(int(co[-1] + cl[-1]) + 64, dtype=torch.uint8, device=dev)
                ctx.copy_units(half, mslots, t_slot_off, t_cl, arena, i64(co))
                parts[name] = (codec, arena, i64(co), t_cl, t_ro, t_s, float(cl.sum()))
            del mslots
            mout = torch.zeros(MB * U, dtype=torch.uint8, device=dev)
            t_mdl = torch.zeros(half, dtype=torch.int64, device=dev)

            def step_mixed():
                for codec, arena, t_co, t_cl, t_ro, t_s, _ in parts.values():
                    ctx.decompress_batch(codec, capi.DEVICE, half, arena, t_co, t_cl, mout, t_ro, t_hl, t_mdl, t_s)
            step_mixed()
            ctx.synchronize()
            assert bool(torch.equal(mout, mraw)), "mixed batch output differs"
            barrier()
            ms_m = max_over_ranks(timed(step_mixed, k=3, warm=1))
            cbytes = sum(p[6] for p in parts.values())
            extras["mixed_snappy_lz4_configs4"] = record(ms_m, MB, U, cbytes, f"{tpb or 'lz_decode_kernel'}<snappy> then {tpb or 'lz_decode_kernel'}<lz4>, one batched call per codec",
                                                         whole_job_GBps=MB * world * U / (ms_m * 1e6), blocks_per_gpu=MB, n_gpus=world,
                                                         streams="GPU encoders; even global block index = snappy, odd = lz4; generated per rank from the global index")
            del mraw, mout, parts
            torch.cuda.empty_cache()
            out = torch.zeros(B * U, dtype=torch.uint8, device=dev)
        except Exception as e:
            extras["mixed_error"] = repr(e)[:300]
            if "out" not in dir():
                out = torch.zeros(B * U, dtype=torch.uint8, device=dev)

    # ---- extras (N == 1): the other legs of "GB/s per codec", device resident, at the BASELINE.json sizes ----
    if world == 1 and not args.no_extras:
        codecs = {}
        slot = (32 + U + U // 6 + 15) // 16 * 16
        eslots = torch.empty(B * slot, dtype=torch.uint8, device=dev)
        t_slot_off, t_slot_cap = i64(np.arange(B, dtype=np.uint64) * slot), i64(np.full(B, slot, np.uint64))
        t_ecl = torch.zeros(B, dtype=torch.int64, device=dev)
        for name, codec, ocodec in (("snappy", capi.SNAPPY_RAW, O_SNAPPY), ("lz4", capi.LZ4_BLOCK, O_LZ4)):
            try:
                # encode (GPU) + decode of the GPU encoder's streams
                ms_c = timed(lambda: ctx.compress_batch(codec, capi.DEVICE, B, raw, t_raw_off, t_raw_len, eslots, t_slot_off, t_slot_cap, t_ecl, t_st))
                assert int((t_st != 0).sum()) == 0
                cb = float(t_ecl.sum().item())
                codecs[f"{name}_block_compress"] = record(ms_c, B, U, cb, f"lz_encode_kernel<{name}>", traffic=ncu_traffic(f"{name}_block_compress", B))
                ms_d = timed(lambda: ctx.decompress_batch(codec, capi.DEVICE, B, eslots, t_slot_off, t_ecl, out, t_raw_off, t_raw_len, t_dl, t_st))
                assert int((t_st != 0).sum()) == 0 and bool(torch.equal(out, raw))
                codecs[f"{name}_block_decompress_gpu_encoded"] = record(ms_d, B, U, cb, f"{tpb or 'lz_decode_kernel'}<{name}>", streams="this engine's GPU encoder",
                                                                         traffic=ncu_traffic(f"{name}_block_decompress_gpu_encoded", B))
                if name == "lz4":   # the headline already is snappy on CPU-encoder streams
                    lcomp, lcoff, lclen, lenc = compress_host(O_LZ4, data, B, U, nt)
                    t_lcomp, t_lco, t_lcl = torch.from_numpy(lcomp).to(dev), i64(lcoff), i64(lclen)
                    ms_d = timed(lambda: ctx.decompress_batch(codec, capi.DEVICE, B, t_lcomp, t_lco, t_lcl, out, t_raw_off, t_raw_len, t_dl, t_st))
                    assert int((t_st != 0).sum()) == 0 and bool(torch.equal(out, raw))
                    codecs["lz4_block_decompress"] = record(ms_d, B, U, float(lclen.sum()), f"{tpb or 'lz_decode_kernel'}<lz4>", traffic=ncu_traffic("lz4_block_decompress", B),
                                                             streams="lz4 (Arrow's bundled liblz4, LZ4_compress_default) made on the host" if lenc == "stand-in" else "oracle/lz4.c encoder")
                    del t_lcomp
            except Exception as e:
                codecs[f"{name}_error"] = repr(e)[:300]
        del eslots
        torch.cuda.empty_cache()
        # the warp-per-block kernel on the headline batch (what small batches and the pinned pipeline's chunks run on)
        default_path = ctx.decode_path()
        try:
            ctx.set_decode_path(2, 1)
            ms_g = timed(step_device, k=3)
            assert int((t_st != 0).sum()) == 0
            codecs["snappy_block_decompress_warp_per_block"] = record(ms_g, B, U, float(clen.sum()), "lz_decode_kernel<snappy, lane-parallel> (rings staged by cp.async.bulk)",
                                                                      streams="as the headline", traffic=ncu_traffic("snappy_block_decompress_warp_per_block", B))
        except Exception as e:
            codecs["gen2_error"] = repr(e)[:200]
        finally:
            ctx.set_decode_path(*default_path)
        # zstd level-3 frame decompress (BASELINE.json configs[3]): frames made on the host by zstd (Arrow's bundled libzstd), level 3
        try:
            ZF, ZU = args.zstd_frames, 262144
            zn = ZF * ZU // U
            zdata = data if zn <= B else O.synth(zn, U, SEED, rank * B, nt)
            zcomp, zo, zl, _ = compress_host(O_ZSTD, zdata, ZF, ZU, nt, level=3)
            t_zsrc, t_zo, t_zl = torch.from_numpy(zcomp).to(dev), i64(zo), i64(zl)
            t_zdo, t_zdc = i64(np.arange(ZF, dtype=np.uint64) * ZU), i64(np.full(ZF, ZU, np.uint64))
            zout = out if ZF * ZU <= out.numel() else torch.zeros(ZF * ZU, dtype=torch.uint8, device=dev)
            t_zdl, t_zst = torch.zeros(ZF, dtype=torch.int64, device=dev), torch.zeros(ZF, dtype=torch.int32, device=dev)
            ms_z = timed(lambda: ctx.decompress_batch(capi.ZSTD, capi.DEVICE, ZF, t_zsrc, t_zo, t_zl, zout, t_zdo, t_zdc, t_zdl, t_zst), k=3, warm=1)
            want = raw if zn <= B else torch.from_numpy(zdata).to(dev)
            assert int((t_zst != 0).sum()) == 0 and bool(torch.equal(zout[: ZF * ZU], want[: ZF * ZU]))
            codecs["zstd_l3_frame_decompress"] = record(ms_z, ZF, ZU, float(zl.sum()), "zstd_decode_kernel", traffic=ncu_traffic("zstd_l3_frame_decompress", ZF),
                                                         streams="zstd level 3 (Arrow's bundled libzstd) made on the host, one frame per 256 KiB")
            # the zstd encoder of this engine on the same frames
            zslot = (capi.lib().cj_compress_bound(capi.ZSTD, ZU) + 15) // 16 * 16
            zs = torch.empty(ZF * zslot, dtype=torch.uint8, device=dev)
            t_zso, t_zsc = i64(np.arange(ZF, dtype=np.uint64) * zslot), i64(np.full(ZF, zslot, np.uint64))
            ms_zc = timed(lambda: ctx.compress_batch(capi.ZSTD, capi.DEVICE, ZF, want, t_zdo, t_zdc, zs, t_zso, t_zsc, t_zdl, t_zst, level=3), k=3, warm=1)
            assert int((t_zst != 0).sum()) == 0
            codecs["zstd_frame_compress"] = record(ms_zc, ZF, ZU, float(t_zdl.sum().item()), "zstd_encode_kernel")
            del zs, t_zsrc
        except Exception as e:
            codecs["zstd_error"] = repr(e)[:300]
        extras["codecs"] = codecs
        # real-corpus row: six Silesia files as the reference ships them (tests/golden/corpus, from benchmarks/data), cut into
        # 64 KiB blocks, compressed on the host by Google snappy / lz4; the batch tiles the block descriptors up to the
        # headline's block count (distinct output slots; the compressed input of the tiled batch is L2-resident, which is said here)
        try:
            import bz2
            cdir = os.path.join(ROOT, "tests", "golden", "corpus")
            files = sorted(f for f in os.listdir(cdir) if f.endswith(".bz2"))
            blobs = [np.frombuffer(bz2.decompress(open(os.path.join(cdir, f), "rb").read()), dtype=np.uint8) for f in files]
            blobs = [b[: len(b) // U * U] for b in blobs]
            cdata = np.concatenate(blobs)
            nb = len(cdata) // U
            rep = max(1, B // nb)
            rc = {"files": [f[:-4] for f in files], "distinct_blocks": nb, "tiled_to": nb * rep,
                  "note": "descriptors tiled (distinct output slots); the ~%d MB of distinct compressed input stays L2-resident" % (nb * U // 2 >> 20)}
            t_want = torch.from_numpy(cdata).to(dev)
            for name, codec, ocodec in (("snappy", capi.SNAPPY_RAW, O_SNAPPY), ("lz4", capi.LZ4_BLOCK, O_LZ4)):
                ccomp, cco, ccl, cenc = compress_host(ocodec, cdata, nb, U, nt)
                t_c = torch.from_numpy(ccomp).to(dev)
                t_co, t_cl = i64(np.tile(cco, rep)), i64(np.tile(ccl, rep))
                n2 = nb * rep
                t_do2, t_dc2 = i64(np.arange(n2, dtype=np.uint64) * U), i64(np.full(n2, U, np.uint64))
                t_dl2, t_st2 = torch.zeros(n2, dtype=torch.int64, device=dev), torch.zeros(n2, dtype=torch.int32, device=dev)
                ms_r = timed(lambda: ctx.decompress_batch(codec, capi.DEVICE, n2, t_c, t_co, t_cl, out, t_do2, t_dc2, t_dl2, t_st2), k=3, warm=1)
                got = out[: n2 * U].view(rep, nb * U)
                assert int((t_st2 != 0).sum()) == 0 and bool(torch.equal(got[0], t_want)) and bool(torch.equal(got[rep - 1], t_want))
                rc[f"{name}_block_decompress"] = record(ms_r, n2, U, float(ccl.sum()) * rep, f"{tpb or 'lz_decode_kernel'}<{name}>", encoder=cenc)
                del t_c
            extras["real_corpus"] = rc
        except Exception as e:
            extras["real_corpus_error"] = repr(e)[:300]
        # single-buffer calls through the cramjam-compatible Python module (BASELINE configs[0] shape and a large buffer)
        try:
            from cramjam_b200 import cramjam as cj_mod
            api = {}
            for mib in (1, 256):
                buf = raw[: mib << 20].cpu().numpy().tobytes()
                for name, mod in (("snappy", cj_mod.snappy), ("lz4", cj_mod.lz4), ("zstd", cj_mod.zstd)):
                    def wall(f, reps):
                        f()
                        t0 = time.perf_counter()
                        for _ in range(reps):
                            r = f()
                        return (time.perf_counter() - t0) / reps, r
                    reps = 10 if mib == 1 else 2
                    tc, cbuf = wall(lambda: mod.compress(buf), reps)
                    cb = bytes(cbuf)
                    td, dbuf = wall(lambda: mod.decompress(cb), reps)
                    assert bytes(dbuf) == buf
                    api[f"{name}_{mib}MiB"] = {"compress_ms": round(tc * 1e3, 2), "decompress_ms": round(td * 1e3, 2), "ratio": round(len(buf) / len(cb), 3)}
            extras["api_single_buffer"] = api
        except Exception as e:
            extras["api_error"] = repr(e)[:200]

    # ---- N > 1 extra: the "batch lives on rank 0" mode of configs[4] — NCCL point-to-point scatter of compressed ranges,
    #      per-rank decode, gather of the outputs (north_star: NCCL only for the trivial scatter/gather) ----
    if world > 1 and not args.no_extras:
        try:
            from cramjam_b200.sharding import make_plan, scatter_payload, gather_payload
            SB = min(B, args.mixed_blocks)               # blocks held by rank 0 for this leg
            plan = make_plan(coff[:SB] if rank == 0 else None, clen[:SB] if rank == 0 else None, np.full(SB, U, np.uint64) if rank == 0 else None,
                             np.full(SB, U), 0, dev)     # descriptors: one broadcast, outside the timed region
            k = plan.count(rank)
            t_so, t_sc = i64(np.arange(max(k, 1), dtype=np.uint64) * U), i64(np.full(max(k, 1), U, np.uint64))
            t_sdl, t_sst = torch.zeros(max(k, 1), dtype=torch.int64, device=dev), torch.zeros(max(k, 1), dtype=torch.int32, device=dev)
            sout = torch.empty(max(k, 1) * U, dtype=torch.uint8, device=dev)
            local_in = torch.empty(max(plan.in_bytes(rank), 16), dtype=torch.uint8, device=dev)
            gathered = torch.empty(SB * U, dtype=torch.uint8, device=dev) if rank == 0 else None
            t_lo, t_ll = i64(plan.local_offsets(rank)), i64(plan.local_lengths(rank))

            def scatter_decode_gather():
                scatter_payload(plan, comp if rank == 0 else None, local_in)
                if k:
                    ctx.decompress_batch(capi.SNAPPY_RAW, capi.DEVICE, k, local_in, t_lo, t_ll, sout, t_so, t_sc, t_sdl, t_sst)
                torch.cuda.current_stream().synchronize()
                gather_payload(plan, sout, gathered)

            scatter_decode_gather()
            torch.cuda.synchronize()
            if rank == 0:
                assert bool(torch.equal(gathered, raw[: SB * U])), "scatter/decode/gather output differs"
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                scatter_decode_gather()
            barrier()
            ms_sg = max_over_ranks(1e3 * (time.perf_counter() - t0) / 3)
            moved = (float(clen[:SB].sum()) + SB * U) * (world - 1) / world if rank == 0 else 0.0
            extras["scatter_decode_gather"] = {"GBps": SB * U / (ms_sg * 1e6), "ms": ms_sg, "blocks": SB, "n_gpus": world,
                                               "rank0_nvlink_bytes": moved, "rank0_nvlink_GBps": moved / (ms_sg * 1e6),
                                               "frac_of_nvlink_770GBps": moved / (ms_sg * 1e6) / 770.0,
                                               "note": "rank 0 holds the batch; NCCL p2p scatter of compressed ranges, per-rank decode, gather of outputs; "
                                                       "bounded by rank 0's NVLink egress+ingress"}
        except Exception as e:
            extras["scatter_error"] = repr(e)[:300]

    # ---- CPU baseline beside it (rank 0 at N == 1 only; bounded sample): the fastest CPU path in the image ----
    cpu = None
    if rank == 0 and world == 1:
        ntc = host_threads()
        cb = min(8192, B)
        have = O.standin() is not None
        rows = {}
        if have:
            rows["standin_all_cores"], p_all = cpu_decode_rate(O_SNAPPY, h_comp_np, coff, clen, U, cb, ntc, args.cpu_seconds, True, check=data)
            rows["standin_one_core"], _ = cpu_decode_rate(O_SNAPPY, h_comp_np, coff, clen, U, min(cb, 1024), 1, 2.0, True)
        rows["port_all_cores"], p_port = cpu_decode_rate(O_SNAPPY, h_comp_np, coff, clen, U, cb, ntc, args.cpu_seconds if not have else 3.0, False, check=data)
        rows["port_one_core"], _ = cpu_decode_rate(O_SNAPPY, h_comp_np, coff, clen, U, min(cb, 1024), 1, 2.0, False)
        best_standin = have and rows["standin_all_cores"] >= rows["port_all_cores"]
        cpu = {"value": rows["standin_all_cores"] if best_standin else rows["port_all_cores"], "unit": "GB/s", "cores": ntc,
               "kind": "stand-in" if best_standin else "port",
               "sample": f"repeated passes (~{args.cpu_seconds:.0f} s) over the first {cb} x 64 KiB blocks of the GPU run's own input; "
                         + ("Arrow's bundled Google snappy via arrow::util::Codec (oracle/standin.cpp)" if best_standin else "oracle/snappy.c"),
               "all": {k: round(v, 3) for k, v in rows.items()}}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(B, ratio, comp_bytes, enc),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic("snappy_block_decompress", B), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel": kernel_name, "redo_units": int(redo)},
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "cpu_baseline": cpu,
            "extras": extras,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
