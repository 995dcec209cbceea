"""Copies the reference's own fixtures for the snappy / lz4 / zstd path into tests/golden/ (run in the build
container, where /root/reference exists) or, with --check, verifies the committed copies against it."""
import hashlib
import os
import shutil
import sys

REF = "/root/reference/tests/data/integration"
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ["plaintext.txt", "plaintext.txt.snappy", "plaintext.txt.lz4", "plaintext.txt.zst"]
# Six files of the Silesia corpus as the reference ships them for its own benchmarks (benchmarks/data/*.bz2, used by
# benchmarks/test_bench.py:63-64): the real-corpus row of bench.py cuts them into 64 KiB blocks.  Data, not source.
CORPUS_REF = "/root/reference/benchmarks/data"
CORPUS = ["dickens.bz2", "xml.bz2", "mr.bz2", "nci.bz2", "ooffice.bz2", "reymont.bz2"]


def md5(p):
    return hashlib.md5(open(p, "rb").read()).hexdigest()


if __name__ == "__main__":
    check = "--check" in sys.argv
    os.makedirs(os.path.join(HERE, "corpus"), exist_ok=True)
    pairs = [(os.path.join(REF, f), os.path.join(HERE, f), f) for f in FILES]
    pairs += [(os.path.join(CORPUS_REF, f), os.path.join(HERE, "corpus", f), "corpus/" + f) for f in CORPUS]
    for src, dst, f in pairs:
        if check:
            print(f, md5(dst), "OK" if md5(src) == md5(dst) else "DIFFERS")
        else:
            shutil.copyfile(src, dst)
            print("copied", f, md5(dst))
