"""Helpers shared by the -m gpu parity tests (all calls go through the C ABI)."""
import numpy as np
import pytest

import oracle as O
from cramjam_b200 import _capi as capi

_ctx = None


def ctx():
    global _ctx
    if _ctx is None:
        _ctx = capi.Context(0)
    return _ctx


ORACLE_DEC = {
    capi.LZ4_BLOCK: O.LZ4_BLOCK, capi.SNAPPY_RAW: O.SNAPPY_RAW, capi.SNAPPY_FRAMED: O.SNAPPY_FRAMED,
    capi.LZ4_FRAME: O.LZ4_FRAME, capi.ZSTD: O.ZSTD,
}


def arena(units, align=16, lead=0):
    """Packs byte strings into one uint8 arena; returns (arena, off u64[n], len u64[n])."""
    n = len(units)
    lens = np.array([len(u) for u in units], dtype=np.uint64)
    off = np.zeros(n, dtype=np.uint64)
    pos = lead
    for i in range(n):
        pos = (pos + align - 1) // align * align if align > 1 else pos
        off[i] = pos
        pos += int(lens[i])
    a = np.zeros(pos + 64, dtype=np.uint8)
    for i, u in enumerate(units):
        a[int(off[i]):int(off[i]) + len(u)] = np.frombuffer(u, dtype=np.uint8)
    return a, off, lens


def oracle_batch(codec, direction, src, so, sl, caps, dst_align=16):
    n = len(so)
    do = np.zeros(n, dtype=np.uint64)
    pos = 0
    for i in range(n):
        pos = (pos + dst_align - 1) // dst_align * dst_align
        do[i] = pos
        pos += int(caps[i])
    dst = np.zeros(pos + 64, dtype=np.uint8)
    out, _ = O.batch(ORACLE_DEC[codec], direction, src, so, sl, dst, do, np.asarray(caps, dtype=np.uint64), nthreads=8)
    return dst, do, out


def gpu_decode_host(codec, units, caps, where=capi.HOST):
    return ctx().run_host_units(codec, False, units, caps, where=where)


def gpu_decode_device(codec, src, so, sl, caps, dst_align=16, dst_lead=0):
    """Device-resident batch through torch tensors.  Returns (dst np, dst_off, dst_len, status)."""
    import torch
    n = len(so)
    do = np.zeros(n, dtype=np.uint64)
    pos = dst_lead
    for i in range(n):
        pos = (pos + dst_align - 1) // dst_align * dst_align if dst_align > 1 else pos
        do[i] = pos
        pos += int(caps[i])
    dev = torch.device("cuda:0")
    t_src = torch.from_numpy(src).to(dev)
    t_dst = torch.zeros(pos + 64, dtype=torch.uint8, device=dev)
    as_i64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).to(dev)
    t_so, t_sl, t_do, t_dc = as_i64(so), as_i64(sl), as_i64(do), as_i64(np.asarray(caps, dtype=np.uint64))
    t_dl = torch.zeros(n, dtype=torch.int64, device=dev)
    t_st = torch.full((n,), -99, dtype=torch.int32, device=dev)
    c = ctx()
    torch.cuda.synchronize()
    c.decompress_batch(codec, capi.DEVICE, n, t_src, t_so, t_sl, t_dst, t_do, t_dc, t_dl, t_st)
    c.synchronize()
    return t_dst.cpu().numpy(), do, t_dl.cpu().numpy().astype(np.uint64), t_st.cpu().numpy()


def assert_same_as_oracle(codec, units, caps, via="host"):
    """GPU result == oracle result (status per unit; bytes where status is OK)."""
    src, so, sl = arena(units)
    odst, odo, olen = oracle_batch(codec, 0, src, so, sl, caps)
    if via == "host":
        outs, st = gpu_decode_host(codec, units, caps)
        for i in range(len(units)):
            want_st = 0 if olen[i] >= 0 else int(-olen[i])
            assert int(st[i]) == want_st, (i, int(st[i]), want_st, len(units[i]), int(caps[i]))
            if want_st == 0:
                assert outs[i] == odst[int(odo[i]):int(odo[i]) + int(olen[i])].tobytes(), i
    else:
        gdst, gdo, glen, st = gpu_decode_device(codec, src, so, sl, caps)
        for i in range(len(units)):
            want_st = 0 if olen[i] >= 0 else int(-olen[i])
            assert int(st[i]) == want_st, (i, int(st[i]), want_st)
            if want_st == 0:
                assert int(glen[i]) == int(olen[i])
                assert np.array_equal(gdst[int(gdo[i]):int(gdo[i]) + int(glen[i])], odst[int(odo[i]):int(odo[i]) + int(olen[i])]), i
