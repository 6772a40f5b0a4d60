"""Build rlsim_b200/librlsim_gpu.so in-tree with nvcc for sm_100a (B200).

    python -m rlsim_b200.build [--force]

-fmad=false is part of the numerical contract: the deterministic math kernels (csrc/detmath.cuh) and the
variate algorithms must evaluate exactly as written so the CUDA path reproduces the "--rng=philox" spec bit for bit.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librlsim_gpu.so")
SOURCES = ["api.cu", "k_transcripts.cu", "k_fragment.cu", "k_pcr.cu", "k_select.cu", "k_fasta.cu", "k_fasta_in.cu", "comm.cu", "hostcopy.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared", "-ldl"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not os.path.isdir(os.path.join(CSRC, f))] + \
           [os.path.join(HERE, "..", "include", "rlsim_gpu.h")]
    return any(os.path.getmtime(d) > t for d in deps)


HOST_DIR = os.path.join(CSRC, "host")
HOST_BIN = os.path.join(HERE, "rlsim_gpu")
HOST_SOURCES = ["args.cpp", "io.cpp", "main.cpp"]


def _host_stale():
    if not os.path.exists(HOST_BIN):
        return True
    t = os.path.getmtime(HOST_BIN)
    deps = [os.path.join(HOST_DIR, f) for f in os.listdir(HOST_DIR)] + [os.path.join(HERE, "..", "include", "rlsim_gpu.h"), LIB]
    return any(os.path.getmtime(d) > t for d in deps)


def build_host(force=False):
    """The C++ twin of the Go host: rlsim_b200/rlsim_gpu (same flags / FASTA / report), linked against librlsim_gpu.so."""
    if not force and not _host_stale():
        return HOST_BIN
    cmd = ["g++", "-O2", "-std=c++17", "-Wall"] + [os.path.join(HOST_DIR, s) for s in HOST_SOURCES] + \
          ["-L" + HERE, "-lrlsim_gpu", "-Wl,-rpath,$ORIGIN", "-o", HOST_BIN]
    subprocess.check_call(cmd)
    return HOST_BIN


def build(force=False, verbose=False):
    if force or _stale():
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
        subprocess.check_call(cmd)
    build_host(force)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
