"""Copies the reference's own fixtures for the snappy / lz4 / zstd path into tests/golden/ (run in the build
container, where /root/reference exists) or, with --check, verifies the committed copies against it."""
import hashlib
import os
import shutil
import sys

REF = "/root/reference/tests/data/integration"
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ["plaintext.txt", "plaintext.txt.snappy", "plaintext.txt.lz4", "plaintext.txt.zst"]


def md5(p):
    return hashlib.md5(open(p, "rb").read()).hexdigest()


if __name__ == "__main__":
    check = "--check" in sys.argv
    for f in FILES:
        src, dst = os.path.join(REF, f), os.path.join(HERE, f)
        if check:
            print(f, md5(dst), "OK" if md5(src) == md5(dst) else "DIFFERS")
        else:
            shutil.copyfile(src, dst)
            print("copied", f, md5(dst))
