"""Builds cramjam_b200/libcramjam_cuda.so in-tree with nvcc for sm_100a (no JIT cache, so the
built library travels with the repo snapshot to the GPU box)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libcramjam_cuda.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
         "-Xcompiler", "-fPIC,-O3,-Wall", "-shared", "-cudart", "shared"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "cramjam_cuda.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + sources()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed building libcramjam_cuda.so")
    return OUT


HOST_SRC = os.path.join(HERE, "host", "cramjam_module.cpp")


def host_module_path():
    import sysconfig
    return os.path.join(HERE, "cramjam" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_host(force=False):
    """Builds the C++ (pybind11) host binding cramjam_b200/cramjam.*.so against libcramjam_cuda.so."""
    import sysconfig
    import pybind11
    out = host_module_path()
    deps = [HOST_SRC, os.path.join(HERE, "..", "include", "cramjam_cuda.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    build()
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", "-Wall",
           "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"], HOST_SRC, "-o", out,
           "-L", HERE, "-lcramjam_cuda", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building the cramjam host module")
    return out


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    build_host(force=True)
    print(OUT)
    print(host_module_path())
