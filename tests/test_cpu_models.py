"""CPU-only checks of host-compilable pieces of the engine and of the benchmark's own inputs:
  * oracle/synth.c (what the CPU arms decode to) == the product's generator, byte for byte;
  * the stand-in harness (Arrow's bundled Google snappy / lz4 / zstd) and the oracle port agree on every stream;
  * the zstd encoder's Huffman literal stage (cramjam_b200/csrc/zstd_huf.cuh, host + device code) builds literal-only
    frames that libzstd and the oracle decode;
  * the generation-6 model (tools/g6_model: checkpoint walk + out-of-order range lanes) decodes bit-exact."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import corpus
import oracle as O
import syslibs as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_synth_equals_product_generator():
    from cramjam_b200 import _capi as capi
    for n, U, first in ((40, 65536, 0), (7, 1000, 123456), (3, 262144, 9)):
        assert np.array_equal(O.synth(n, U, first_index=first, nthreads=3), capi.synth_host(n, U, first_index=first))


def test_standin_and_port_agree():
    if O.standin() is None:
        pytest.skip("pyarrow (libarrow + headers) not present: no stand-in harness")
    n, U = 96, 65536
    data = O.synth(n, U, first_index=500)
    so, ul = np.arange(n, dtype=np.uint64) * U, np.full(n, U, np.uint64)
    for codec, port in ((0, O.SNAPPY_RAW), (2, O.LZ4_BLOCK)):
        slot = 80000
        comp = np.zeros(n * slot, dtype=np.uint8)
        do, cap = np.arange(n, dtype=np.uint64) * slot, np.full(n, slot, np.uint64)
        cl, _ = O.standin_batch(codec, 1, data, so, ul, comp, do, cap, nthreads=4)
        assert (cl > 0).all()
        a, b = np.zeros(n * U, dtype=np.uint8), np.zeros(n * U, dtype=np.uint8)
        la, _ = O.standin_batch(codec, 0, comp, do, cl.astype(np.uint64), a, so, ul, nthreads=4)
        lb, _ = O.batch(port, 0, comp, do, cl.astype(np.uint64), b, so, ul, nthreads=4)
        assert (la == U).all() and (lb == U).all() and np.array_equal(a, data) and np.array_equal(b, data)
        # and the other way round: the stand-in decodes the oracle encoder's streams
        cl2, _ = O.batch(port, 1, data, so, ul, comp, do, cap, nthreads=4)
        la, _ = O.standin_batch(codec, 0, comp, do, cl2.astype(np.uint64), a, so, ul, nthreads=4)
        assert (la == U).all() and np.array_equal(a, data)
    ZU = 262144
    zn = n * U // ZU
    zs, zu = np.arange(zn, dtype=np.uint64) * ZU, np.full(zn, ZU, np.uint64)
    zcomp = np.zeros(zn * 270000, dtype=np.uint8)
    zdo, zcap = np.arange(zn, dtype=np.uint64) * 270000, np.full(zn, 270000, np.uint64)
    zl, _ = O.standin_batch(4, 1, data, zs, zu, zcomp, zdo, zcap, nthreads=4, level=3)
    out = np.zeros(n * U, dtype=np.uint8)
    lo, _ = O.batch(O.ZSTD, 0, zcomp, zdo, zl.astype(np.uint64), out, zs, zu, nthreads=4)
    assert (lo == ZU).all() and np.array_equal(out, data)


def _build(src, out, *extra):
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", src, "-o", out, *extra])
    return C.CDLL(out)


ZH_HARNESS = r'''
#include <stdint.h>
#include <string.h>
#include <vector>
#include "%s/cramjam_b200/csrc/zstd_huf.cuh"
using namespace cj::zh;
// a single-segment zstd frame whose only block holds n literals and no sequences
extern "C" long zh_frame(const uint8_t* lit, uint32_t n, uint8_t* out) {
    uint32_t count[256] = {0};
    for (uint32_t i = 0; i < n; i++) count[lit[i]]++;
    uint8_t nbits[256]; uint16_t code[256]; uint8_t tree[160];
    const int mb = build_lengths(count, nbits);
    if (!mb) return 0;
    assign_codes(nbits, mb, code);
    const int tl = write_tree_direct(nbits, mb, tree);
    if (!tl) return 0;
    const bool four = n >= 1024;
    std::vector<uint8_t> body(tree, tree + tl), tmp(2 * n + 64);
    if (!four) { const uint32_t s = encode_stream(lit, n, code, tmp.data()); body.insert(body.end(), tmp.begin(), tmp.begin() + s); }
    else {
        const uint32_t seg = (n + 3) / 4;
        std::vector<uint8_t> st[4];
        for (int k = 0; k < 4; k++) {
            const uint32_t a = k * seg, b = k == 3 ? n : (k + 1) * seg;
            const uint32_t s = encode_stream(lit + a, b - a, code, tmp.data());
            st[k].assign(tmp.begin(), tmp.begin() + s);
        }
        for (int k = 0; k < 3; k++) { body.push_back(st[k].size() & 0xff); body.push_back(st[k].size() >> 8); }
        for (int k = 0; k < 4; k++) body.insert(body.end(), st[k].begin(), st[k].end());
    }
    const uint32_t comp = (uint32_t)body.size();
    if ((n < 1024 && comp >= 1024) || (n < 16384 && comp >= 16384)) return 0;
    const int hl = header_len(n);
    uint8_t hdr[5];
    write_header(hdr, n, comp, four);
    uint32_t op;
    out[0] = 0x28; out[1] = 0xB5; out[2] = 0x2F; out[3] = 0xFD;
    if (n < 256) { out[4] = 0x20; out[5] = (uint8_t)n; op = 6; }
    else if (n < 65536 + 256) { out[4] = 0x60; const uint32_t v = n - 256; out[5] = (uint8_t)v; out[6] = (uint8_t)(v >> 8); op = 7; }
    else { out[4] = 0xA0; memcpy(out + 5, &n, 4); op = 9; }
    const uint32_t bh = 1u | (2u << 1) | ((hl + comp + 1) << 3);
    out[op] = (uint8_t)bh; out[op + 1] = (uint8_t)(bh >> 8); out[op + 2] = (uint8_t)(bh >> 16); op += 3;
    memcpy(out + op, hdr, hl); op += hl;
    memcpy(out + op, body.data(), comp); op += comp;
    out[op++] = 0;
    return op;
}
'''


def test_zstd_huffman_literal_stage_on_the_host(tmp_path):
    src = tmp_path / "zh.cpp"
    src.write_text(ZH_HARNESS % ROOT)
    L = _build(str(src), str(tmp_path / "zh.so"))
    L.zh_frame.restype = C.c_long
    L.zh_frame.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    rng = np.random.default_rng(0)
    p = 0.5 ** np.arange(1, 60)
    cases = [corpus.text(n, n) for n in (70, 300, 1023, 1024, 5000, 16383, 16384, 60000, 131072)]
    cases += [bytes(rng.choice(59, size=100000, p=p / p.sum()).astype(np.uint8)),       # a deep tree: the 11-bit limit is hit
              bytes(rng.integers(0, 128, size=50000).astype(np.uint8)), bytes(rng.choice([65, 66], size=3000).astype(np.uint8)),
              O.synth(1, 40000).tobytes()]
    made = 0
    for t in cases:
        a = np.frombuffer(t, dtype=np.uint8)
        out = np.zeros(len(t) + 1024, dtype=np.uint8)
        r = L.zh_frame(a.ctypes.data, len(a), out.ctypes.data)
        if r <= 0:
            continue
        f = out[:r].tobytes()
        assert O.zstd_decompress(f) == t
        if S.have_zstd:
            assert S.zstd_decompress(f, len(t)) == t
        made += 1
    assert made >= 11
    big = bytes(rng.integers(0, 256, size=5000).astype(np.uint8))                        # 256 symbols: no direct tree description
    a = np.frombuffer(big, dtype=np.uint8)
    assert L.zh_frame(a.ctypes.data, len(a), np.zeros(8000, dtype=np.uint8).ctypes.data) == 0


def test_generation6_model_decodes_bit_exact(tmp_path):
    L = _build(os.path.join(ROOT, "tools", "g6_model", "g6_emu.cpp"), str(tmp_path / "g6.so"))
    L.g6_emu_decode.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_long), C.POINTER(C.c_long)]
    L.g6_emu_decode.restype = C.c_long

    def emu(codec, comp, cap, lanes):
        a = np.frombuffer(comp, dtype=np.uint8)
        out = np.zeros(70000, dtype=np.uint8)
        it, mv = C.c_long(), C.c_long()
        r = L.g6_emu_decode(codec, a.ctypes.data, len(a), out.ctypes.data, cap, lanes, C.byref(it), C.byref(mv))
        return r, out[:max(r, 0)].tobytes()
    data = O.synth(6, 65536, first_index=11)
    cases = [data[i * 65536:(i + 1) * 65536].tobytes() for i in range(6)] + [d for d in corpus.edge_cases() if 0 < len(d) <= 65536]
    for codec, comp in ((0, O.snappy_raw_compress), (2, O.lz4_block_compress)):
        for d in cases:
            c = comp(d)
            for lanes in (1, 64, 512):
                r, out = emu(codec, c, len(d), lanes)
                assert r == len(d) and out == d, (codec, len(d), lanes, r)
