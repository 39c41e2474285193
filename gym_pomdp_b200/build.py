"""Build libpomdp_b200.so in-tree with nvcc for sm_100a (B200).

    python -m gym_pomdp_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with
the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = [os.path.join(CSRC, "pomdp_kernels.cu")]
HEADERS = [os.path.join(CSRC, "pomdp_core.h"), os.path.join(CSRC, "pomdp_envs.h"), os.path.join(CSRC, "pomdp_host.h"),
           os.path.join(os.path.dirname(HERE), "include", "pomdp_b200.h")]
OUT = os.path.join(CSRC, "libpomdp_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(p) <= t for p in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", OUT] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    log = os.path.join(CSRC, "ptxas.log")
    with open(log, "w") as f:
        f.write(res.stderr)
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
