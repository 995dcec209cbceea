"""Device-resident decode timing of the thread-per-block paths (generations 4 and 7) over knob settings, one process
(development tool; bench.py is the contract).  usage: python tools/g7_sweep.py [n_blocks] [snappy|lz4 ...]
Streams are made by this engine's GPU encoder; every variant's output is compared with the generator's bytes."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
if os.environ.get("EARLY_L2_FETCH"):   # experiment: cudaLimitMaxL2FetchGranularity set before anything else touches the device
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    v = ctypes.c_size_t()
    print("cudaDeviceSetLimit ->", rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["EARLY_L2_FETCH"]))), flush=True)
    print("cudaDeviceGetLimit ->", rt.cudaDeviceGetLimit(ctypes.byref(v), 5), v.value, flush=True)
import torch
from cramjam_b200 import _capi as capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
REAL = "--real" in sys.argv   # the six Silesia files of tests/golden/corpus instead of the synthetic generator (oracle-encoded, descriptors tiled)
codecs = [a for a in sys.argv[2:] if not a.startswith("--")] or ["snappy", "lz4"]
U = 65536
dev = torch.device("cuda:0")
c = capi.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
c.set_stream(stream.cuda_stream)
i64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).to(dev)
ORACLE = "--oracle" in sys.argv   # synthetic blocks encoded on the host by the oracle encoder (streams like the bench headline's Google-snappy ones)
NEAR = next((int(a.split("=")[1]) for a in sys.argv if a.startswith("--near=")), 0)   # synthetic blocks whose offsets are folded below this distance
def near_blocks(nb, cap, seed=1):
    rng = np.random.default_rng(seed)
    out = np.zeros((nb, U), dtype=np.uint8)
    for b in range(nb):
        o = out[b]
        pos = 0
        r_ll = rng.integers(0, 16, size=20000); r_l2 = rng.integers(0, 7, size=20000); r_ml = rng.integers(0, 13, size=20000)
        r_oc = rng.integers(0, 100, size=20000); r_of = rng.integers(0, 1 << 30, size=20000)
        lits = (32 + (rng.integers(0, 96, size=U) * rng.integers(0, 96, size=U)) // 96).astype(np.uint8)
        i = 0
        while pos < U:
            ll = 0 if r_ll[i] < 10 else 1 + int(r_l2[i])
            ll = min(ll, U - pos)
            o[pos:pos + ll] = lits[pos:pos + ll]
            pos += ll
            if pos >= U:
                break
            if pos >= 8:
                ml = min(5 + int(r_ml[i]), U - pos)
                oc = r_oc[i]
                off = 16 + r_of[i] % 240 if oc < 24 else (256 + r_of[i] % 3840 if oc < 65 else 4096 + r_of[i] % 61440)
                off = int(off)
                if off > cap: off = 33 + (off - 33) % (cap - 32)
                if off > pos: off = 1 + (off - 1) % pos
                for j in range(ml):
                    o[pos + j] = o[pos + j - off]
                pos += ml
            i += 1
    return out.reshape(-1)
if REAL or NEAR or ORACLE:
    import bz2
    import oracle as O
    cdir = os.path.join(ROOT, "tests", "golden", "corpus")
    blobs = [] if not REAL else [np.frombuffer(bz2.decompress(open(os.path.join(cdir, f), "rb").read()), dtype=np.uint8) for f in sorted(os.listdir(cdir)) if f.endswith(".bz2")]
    CACHE = os.environ.get("SWEEP_CACHE")   # directory: host-made inputs are kept there between the processes of one A/B run
    def cached(tag, make):
        if not CACHE:
            return make()
        f = os.path.join(CACHE, f"g7sweep_{tag}_{n}.npy")
        if os.path.exists(f):
            return np.load(f)
        a = make()
        os.makedirs(CACHE, exist_ok=True)
        np.save(f, a)
        return a
    cdata = np.concatenate([b[: len(b) // U * U] for b in blobs]) if REAL else (near_blocks(128, NEAR) if NEAR else cached("data", lambda: capi.synth_host(n, U)))
    nb = len(cdata) // U
    rep = max(1, n // nb)
    n = nb * rep
    data = torch.from_numpy(cdata).to(dev)
else:
    data = torch.from_numpy(capi.synth_host(n, U)).to(dev)
for name in codecs:
    codec = capi.LZ4_BLOCK if name == "lz4" else capi.SNAPPY_RAW
    slot = (capi.lib().cj_compress_bound(codec, U) + 15) // 16 * 16
    t_cmp = torch.zeros(n * slot + 64, dtype=torch.uint8, device=dev)
    t_uo, t_ul = i64(np.arange(n, dtype=np.uint64) * U), i64(np.full(n, U, np.uint64))
    t_co, t_cc = i64(np.arange(n, dtype=np.uint64) * slot), i64(np.full(n, slot, np.uint64))
    t_cl = torch.zeros(n, dtype=torch.int64, device=dev)
    t_st = torch.zeros(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    if REAL or NEAR or ORACLE:
        comp = np.zeros(nb * slot + 64, dtype=np.uint8)
        so1, do1 = np.arange(nb, dtype=np.uint64) * np.uint64(U), np.arange(nb, dtype=np.uint64) * np.uint64(slot)
        def encode():
            lens = O.batch(O.LZ4_BLOCK if name == "lz4" else O.SNAPPY_RAW, 1, cdata, so1, np.full(nb, U, np.uint64), comp, do1, np.full(nb, slot, np.uint64), nthreads=os.cpu_count())[0]
            return np.concatenate([lens.astype(np.uint64).view(np.uint8), comp])
        if ORACLE and not REAL and not NEAR:
            both = cached(f"comp_{name}", encode)
        else:
            both = encode()
        lens, comp = both[:nb * 8].view(np.uint64).astype(np.int64), both[nb * 8:]
        assert (lens > 0).all()
        t_cmp = torch.from_numpy(comp).to(dev)
        t_co, t_cl = i64(np.tile(do1, rep)), i64(np.tile(lens.astype(np.uint64), rep))
    else:
        c.compress_batch(codec, capi.DEVICE, n, data, t_uo, t_ul, t_cmp, t_co, t_cc, t_cl, t_st)
        c.synchronize()
        assert int((t_st != 0).sum()) == 0
    ratio = n * U / float(t_cl.sum().item())
    t_dst = torch.zeros(n * U + 64, dtype=torch.uint8, device=dev)
    t_dl = torch.zeros(n, dtype=torch.int64, device=dev)
    variants = [(4, {}), (7, {"CJ_G7_D": "2"}), (7, {"CJ_G7_D": "3"}), (7, {"CJ_G7_D": "4"})]
    if os.environ.get("SWEEP_ONLY"):   # e.g. SWEEP_ONLY=7:3 -> generation 7 with CJ_G7_D=3 only
        g_, d_ = os.environ["SWEEP_ONLY"].split(":")
        variants = [(int(g_), {"CJ_G7_D": d_})]
    for w in os.environ.get("SWEEP_WARPS", "").split(","):
        if w:
            variants.append((7, {"CJ_G7_D": "3", "CJ_G7_WARPS": w}))
    for gen, env in variants:
        for k in ("CJ_G7_D", "CJ_G7_WARPS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        c.set_decode_path(gen, 1)
        t_dst.zero_()
        torch.cuda.synchronize()
        for _ in range(2):
            c.decompress_batch(codec, capi.DEVICE, n, t_cmp, t_co, t_cl, t_dst, t_uo, t_ul, t_dl, t_st)
        c.synchronize()
        if REAL or NEAR or ORACLE:
            got = t_dst[:n * U].view(rep, nb * U)
            ok = int((t_st != 0).sum()) == 0 and torch.equal(got[0], data) and torch.equal(got[rep - 1], data)
        else:
            ok = int((t_st != 0).sum()) == 0 and torch.equal(t_dst[:n * U], data)
        redo = c.last_redo_count()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        K = 5
        ev[0].record(stream)
        for _ in range(K):
            c.decompress_batch(codec, capi.DEVICE, n, t_cmp, t_co, t_cl, t_dst, t_uo, t_ul, t_dl, t_st)
        ev[1].record(stream)
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / K
        gbs = n * U / ms / 1e6
        print(f"[{name}{' real' if REAL else ''}] gen {gen} {env}: {ms:.3f} ms  {gbs:.1f} GB/s uncompressed  frac {gbs * (1 + 1 / ratio) / 6537:.4f}  "
              f"bit-exact {ok}  redo {redo}  ratio {ratio:.3f}", flush=True)
