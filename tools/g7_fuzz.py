"""One-off fuzz of the generation-7 decode path against the oracle on the GPU: many more mutated / truncated streams than the
test-suite carries (status code and bytes of every unit must agree).  usage: python tools/g7_fuzz.py [streams_per_codec] [seed]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import corpus
import oracle as O
from cramjam_b200 import _capi as capi
from gpu_util import assert_same_as_oracle, ctx
from test_gpu_lz_decode import _mutations

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1234
rng = np.random.default_rng(seed)
ctx().set_decode_path(7, 1)
synth = capi.synth_host(8, 65536, seed=seed)
bases = [corpus.text(3000, 1), corpus.lz_model(5000, 2), corpus.random_bytes(300, 3), b"a" * 700, corpus.lz_model(70000, 5), corpus.text(65536, 7),
         b"ab" * 3000, bytes(range(256)) * 8] + [synth[i * 65536:(i + 1) * 65536].tobytes() for i in range(8)]
for name, codec, comp in (("snappy", capi.SNAPPY_RAW, O.snappy_raw_compress), ("lz4", capi.LZ4_BLOCK, O.lz4_block_compress)):
    units, caps = [], []
    per = N // len(bases) + 1
    for d in bases:
        c = comp(d)
        for m in _mutations(c, rng, per):
            units.append(m)
            caps.append(max(0, len(d) + int(rng.integers(-8, 64))))
    for lo in range(0, len(units), 4000):
        assert_same_as_oracle(codec, units[lo:lo + 4000], caps[lo:lo + 4000], "device")
    print(f"{name}: {len(units)} mutated streams agree with the oracle (status and bytes), redo count of the last batch {ctx().last_redo_count()}", flush=True)
