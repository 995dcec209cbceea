#!/usr/bin/env python
"""bench.py — headline benchmark of the cramjam_b200 engine (contract: see the task statement).

Workload (BASELINE.json configs[1]): snappy raw block decompress of 65 536 x 64 KiB synthetic
"Silesia-like" blocks per GPU (4 GiB uncompressed out, ~2.1 GiB compressed in).  One step = one
pass of the hot path (one batched decode launch) over the whole batch.

  value      uncompressed GB/s, whole job, inputs and outputs resident in HBM, CUDA-event timed
  e2e        same metric through the C-ABI call with PINNED HOST buffers (H2D + kernel + D2H inside
             the timed region)
  roofline   algorithmic bytes (compressed read + uncompressed written) / kernel time vs measured
             HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the CPU oracle (a stated stand-in port of the reference's Rust path, which cannot be
             built here) on all host cores, bounded sample

`--impl reference` times only that CPU path (no GPU code on it).  N>1 (torchrun): one process per
GPU, blocks are sharded by global index (rank r owns blocks [r*B, (r+1)*B)), no data-path
collective is needed; weak scaling; time = max over ranks.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--blocks", type=int, default=65536, help="64 KiB blocks per GPU")
    ap.add_argument("--cpu-blocks", type=int, default=8192, help="blocks in the bounded CPU sample")
    ap.add_argument("--no-extras", action="store_true", help="skip the per-codec extra rows")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(blocks):
    """Per-launch DRAM bytes of the decode kernel from the committed ncu capture, if it was taken at this size."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p))
        if int(t["blocks"]) == int(blocks):
            return float(t["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


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
# CPU path (the oracle port of the reference's Rust codecs), shared by cpu_baseline and --impl reference
# ------------------------------------------------------------------------------------------------
def cpu_sample(blocks, first_index=0):
    """Host-side synthetic blocks + their snappy-raw compressed form (oracle encoder), untimed set-up."""
    import oracle as O
    from cramjam_b200 import _capi as capi
    nt = os.cpu_count() or 1
    data = capi.synth_host(blocks, U, SEED, first_index)
    slot = (capi.lib().cj_compress_bound(capi.SNAPPY_RAW, U) + 15) // 16 * 16
    comp = np.zeros(blocks * slot, dtype=np.uint8)
    so = np.arange(blocks, dtype=np.uint64) * U
    do = np.arange(blocks, dtype=np.uint64) * slot
    clen, _ = O.batch(O.SNAPPY_RAW, 1, data, so, np.full(blocks, U, np.uint64), comp, do, np.full(blocks, slot, np.uint64), nthreads=nt)
    assert (clen > 0).all()
    return data, comp, do, clen.astype(np.uint64), so


def cpu_decode_pass(comp, do, clen, out, so, nthreads):
    import oracle as O
    n = len(do)
    dl, sec = O.batch(O.SNAPPY_RAW, 0, comp, do, clen, out, so, np.full(n, U, np.uint64), nthreads=nthreads)
    assert (dl == U).all()
    return sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nt = os.cpu_count() or 1
    blocks = args.cpu_blocks
    data, comp, do, clen, so = cpu_sample(blocks)
    out = np.zeros(blocks * U, dtype=np.uint8)
    for _ in range(max(args.warmup, 1)):
        cpu_decode_pass(comp, do, clen, out, so, nt)
    assert np.array_equal(out, data)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_decode_pass(comp, do, clen, out, so, nt)
    ms = 1e3 * t / args.steps
    gbs = blocks * U / (ms * 1e6)
    sample = f"{blocks} x 64 KiB synthetic blocks per step (bounded sample of the {args.blocks}-block workload), oracle/snappy.c, {nt} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"snappy raw block decompress, {args.blocks} x 64 KiB synthetic Silesia-like blocks per GPU (BASELINE.json configs[1])",
                   "sample_blocks": blocks, "ratio": float(blocks * U / clen.sum())},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": nt, "kind": "port", "sample": sample},
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

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; cramjam_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
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

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    B = args.blocks
    ctx = capi.Context(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    i64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).to(dev)

    # ---- set-up (untimed): synthesize this rank's blocks on the device, compress them with the GPU
    #      encoder, pack the compressed blocks into a dense 16-byte aligned arena ----
    raw = torch.empty(B * U, dtype=torch.uint8, device=dev)
    ctx.synth_device(raw, B, U, SEED, first_index=rank * B)
    slot = (capi.lib().cj_compress_bound(capi.SNAPPY_RAW, U) + 15) // 16 * 16
    slots = torch.empty(B * slot, dtype=torch.uint8, device=dev)
    raw_off = np.arange(B, dtype=np.uint64) * U
    t_raw_off, t_raw_len = i64(raw_off), i64(np.full(B, U, np.uint64))
    t_slot_off, t_slot_cap = i64(np.arange(B, dtype=np.uint64) * slot), i64(np.full(B, slot, np.uint64))
    t_clen = torch.zeros(B, dtype=torch.int64, device=dev)
    t_st = torch.zeros(B, dtype=torch.int32, device=dev)
    ctx.compress_batch(capi.SNAPPY_RAW, capi.DEVICE, B, raw, t_raw_off, t_raw_len, slots, t_slot_off, t_slot_cap, t_clen, t_st)
    ctx.synchronize()
    assert int((t_st != 0).sum()) == 0, "GPU snappy encoder failed on the synthetic corpus"
    clen = t_clen.cpu().numpy().astype(np.uint64)
    coff = np.zeros(B, dtype=np.uint64)
    coff[1:] = np.cumsum((clen[:-1] + np.uint64(15)) & ~np.uint64(15))
    comp_bytes = int(coff[-1] + clen[-1])
    comp_span = (comp_bytes + 15) // 16 * 16
    comp = torch.zeros(comp_span + 64, dtype=torch.uint8, device=dev)
    t_coff = i64(coff)
    ctx.copy_units(B, slots, t_slot_off, t_clen, comp, t_coff)
    ctx.synchronize()
    del slots
    torch.cuda.empty_cache()
    ratio = B * U / float(clen.sum())

    out = torch.zeros(B * U, dtype=torch.uint8, device=dev)
    t_dl = torch.zeros(B, dtype=torch.int64, device=dev)

    def step_device():
        ctx.decompress_batch(capi.SNAPPY_RAW, capi.DEVICE, B, comp, t_coff, t_clen, out, t_raw_off, t_raw_len, t_dl, t_st)

    # ---- correctness of the timed path on this exact input (untimed) ----
    step_device()
    ctx.synchronize()
    assert int((t_st != 0).sum()) == 0 and bool(torch.equal(out, raw)), "decode output differs from the original blocks"

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
    kernel_name = ("g4_kernel<snappy> (thread per block) || lz_decode_kernel<snappy, lane-parallel> (warp per block), co-scheduled halves of one batch"
                   if gen == 5 and B >= min_units else
                   {3: "g3_index_kernel + g3_exec_kernel<snappy>", 4: "g4_kernel<snappy>"}.get(gen if B >= min_units else 2, "lz_decode_kernel<snappy, lane-parallel>"))
    value = total_blocks * U / (ms_dev * 1e6)
    alg_bytes = float(clen.sum()) + float(B) * U   # per launch on this rank: compressed read + uncompressed written
    peak, peak_src = peaks()
    achieved = alg_bytes / (ms_dev * 1e6)

    # ---- e2e: pinned host buffers through the same C-ABI call (H2D + kernel + D2H inside) ----
    h_comp = torch.empty(comp_span + 64, dtype=torch.uint8).pin_memory()
    h_comp.copy_(comp.cpu())
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

    # ---- extras (rank 0, N==1): the other legs of "GB/s per codec", device resident ----
    extras = {}
    if world == 1 and not args.no_extras:
        EB = B   # the configs[2] size (65 536 x 64 KiB by default): large enough for the thread-per-block decode path
        def timed(fn, k=5):
            for _ in range(2):
                fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(k):
                fn()
            b.record(stream)
            torch.cuda.synchronize()
            return a.elapsed_time(b) / k
        eslots = torch.empty(EB * slot, dtype=torch.uint8, device=dev)
        for name, codec in (("snappy", capi.SNAPPY_RAW), ("lz4", capi.LZ4_BLOCK)):
            ms_c = timed(lambda: ctx.compress_batch(codec, capi.DEVICE, EB, raw, t_raw_off, t_raw_len, eslots, t_slot_off, t_slot_cap, t_clen, t_st))
            r = EB * U / float(t_clen[:EB].sum().item())
            extras[f"{name}_block_compress_GBps"] = EB * U / (ms_c * 1e6)
            extras[f"{name}_gpu_ratio"] = r
            if codec == capi.LZ4_BLOCK:
                ms_d = timed(lambda: ctx.decompress_batch(codec, capi.DEVICE, EB, eslots, t_slot_off, t_clen, out, t_raw_off, t_raw_len, t_dl, t_st))
                assert int((t_st[:EB] != 0).sum()) == 0 and bool(torch.equal(out[: EB * U], raw[: EB * U]))
                extras["lz4_block_decompress_GBps"] = EB * U / (ms_d * 1e6)
        extras["blocks"] = EB
        del eslots
        # zstd level-3 frame decompress (BASELINE configs[3] shape, reduced count): frames made by the system libzstd on the host
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import syslibs as S
            from concurrent.futures import ThreadPoolExecutor
            if S.have_zstd:
                ZF, ZU = 4096, 262144
                zraw = raw[: ZF * ZU].cpu().numpy()
                with ThreadPoolExecutor(os.cpu_count()) as ex:
                    frames = list(ex.map(lambda i: S.zstd_compress(zraw[i * ZU:(i + 1) * ZU].tobytes(), 3), range(ZF)))
                zl = np.array([len(f) for f in frames], dtype=np.uint64)
                zo = np.zeros(ZF, dtype=np.uint64); zo[1:] = np.cumsum((zl[:-1] + np.uint64(15)) & ~np.uint64(15))
                zsrc = np.zeros(int(zo[-1] + zl[-1]) + 64, dtype=np.uint8)
                for i, f in enumerate(frames):
                    zsrc[int(zo[i]):int(zo[i]) + len(f)] = np.frombuffer(f, dtype=np.uint8)
                t_zsrc = torch.from_numpy(zsrc).to(dev)
                t_zo, t_zl = i64(zo), i64(zl)
                t_zdo, t_zdc = i64(np.arange(ZF, dtype=np.uint64) * ZU), i64(np.full(ZF, ZU, np.uint64))
                ms_z = timed(lambda: ctx.decompress_batch(capi.ZSTD, capi.DEVICE, ZF, t_zsrc, t_zo, t_zl, out, t_zdo, t_zdc, t_dl, t_st), k=3)
                assert int((t_st[:ZF] != 0).sum()) == 0 and bool(torch.equal(out[: ZF * ZU], raw[: ZF * ZU]))
                extras["zstd_l3_frame_decompress_GBps"] = ZF * ZU / (ms_z * 1e6)
                extras["zstd_frames"] = ZF
                extras["zstd_ratio_libzstd_l3"] = ZF * ZU / float(zl.sum())
        except Exception as e:  # extras never fail the headline line
            extras["zstd_error"] = repr(e)[:200]
        # the other block-decode paths on the headline batch: warp per block, index walk + lane
        # state machines, and thread-per-block and warp-per-block side by side on a split batch; the headline `value` is the
        # default path (thread per block)
        default_path = ctx.decode_path()
        try:
            t_clen_in = i64(clen)   # the extras above reused t_clen for other codecs
            for gen, key in ((2, "snappy_block_decompress_gen2_GBps"), (3, "snappy_block_decompress_gen3_GBps"), (5, "snappy_block_decompress_gen5_GBps")):
                ctx.set_decode_path(gen, 4096)
                ms_g = timed(lambda: ctx.decompress_batch(capi.SNAPPY_RAW, capi.DEVICE, B, comp, t_coff, t_clen_in, out, t_raw_off, t_raw_len, t_dl, t_st), k=3)
                assert int((t_st[:B] != 0).sum()) == 0
                extras[key] = B * U / (ms_g * 1e6)
        except Exception as e:
            extras["decode_paths_error"] = repr(e)[:200]
        finally:
            ctx.set_decode_path(*default_path)
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

    # ---- N > 1 extra: the "batch lives on rank 0" mode — NCCL point-to-point scatter of compressed ranges,
    #      per-rank decode, gather of the outputs (north_star: NCCL only for the trivial scatter/gather) ----
    if world > 1 and not args.no_extras:
        from cramjam_b200.sharding import partition_units, scatter_units, gather_units
        SB = 16384                                   # blocks held by rank 0 for this leg (1 GiB uncompressed)
        ranges = partition_units(np.full(SB, U), world)
        s_lo, s_hi = ranges[rank]
        k = s_hi - s_lo
        t_so = i64(np.arange(max(k, 1), dtype=np.uint64) * U)
        t_sc = i64(np.full(max(k, 1), U, np.uint64))
        t_sdl = torch.zeros(max(k, 1), dtype=torch.int64, device=dev)
        t_sst = torch.zeros(max(k, 1), dtype=torch.int32, device=dev)
        sout = torch.empty(max(k, 1) * U, dtype=torch.uint8, device=dev)

        def scatter_decode_gather():
            pay = comp if rank == 0 else None
            local, loff, llen = scatter_units(pay, coff[:SB] if rank == 0 else None, clen[:SB] if rank == 0 else None, ranges, 0, dev)
            if k:
                ctx.decompress_batch(capi.SNAPPY_RAW, capi.DEVICE, k, local, i64(loff), i64(llen), sout, t_so, t_sc, t_sdl, t_sst)
            torch.cuda.current_stream().synchronize()
            return gather_units(sout, np.full(k, U, np.uint64), ranges, 0)

        res = scatter_decode_gather()
        if rank == 0:
            assert bool(torch.equal(res[0], raw[: SB * U])), "scatter/decode/gather output differs"
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            scatter_decode_gather()
        barrier()
        ms_sg = max_over_ranks(1e3 * (time.perf_counter() - t0) / 3)
        extras = {"scatter_decode_gather_GBps": SB * U / (ms_sg * 1e6), "scatter_blocks": SB,
                  "note": "rank 0 holds the batch; NCCL p2p scatter of compressed ranges + gather of outputs; bounded by rank 0's NVLink ingress"}

    # ---- CPU baseline beside it (rank 0 at N == 1 only; bounded sample) ----
    cpu = None
    if rank == 0 and world == 1:
        nt = os.cpu_count() or 1
        cb = min(args.cpu_blocks, B)
        data, ccomp, cdo, cclen, cso = cpu_sample(cb)
        cout = np.zeros(cb * U, dtype=np.uint8)
        cpu_decode_pass(ccomp, cdo, cclen, cout, cso, nt)
        assert np.array_equal(cout, data)
        passes, t = 0, 0.0
        while t < 10.0 and passes < 200:
            t += cpu_decode_pass(ccomp, cdo, cclen, cout, cso, nt)
            passes += 1
        t1 = cpu_decode_pass(ccomp, cdo, cclen, cout, cso, 1) if cb <= 8192 else None
        cpu = {"value": cb * U * passes / t / 1e9, "unit": "GB/s", "cores": nt, "kind": "port",
               "sample": f"{passes} passes over {cb} x 64 KiB synthetic blocks (same generator/seed as the GPU run), oracle/snappy.c, {nt} threads",
               "single_thread_GBps": (cb * U / t1 / 1e9) if t1 else None}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"snappy raw block decompress, {B} x 64 KiB synthetic Silesia-like blocks per GPU (BASELINE.json configs[1])",
                       "blocks_per_gpu": B, "block_bytes": U, "ratio": ratio, "compressed_bytes_per_gpu": comp_bytes,
                       "l2": "inputs+outputs per step (~6 GiB) far exceed the 126 MB L2; no flush needed",
                       "sharding": "contiguous global block ranges per rank, no data-path collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(B), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel": kernel_name},
            "e2e": {"value": e2e_val, "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e,
                    "steps": e2e_steps, "path": "cj_decompress_batch(CJ_SNAPPY_RAW, CJ_PINNED) from pinned host arenas"},
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
