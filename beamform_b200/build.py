"""In-tree build of the CUDA extension (sm_100a only) and nothing else.

`python -m beamform_b200.build` or `beamform_b200.build.build()`.  The resulting
`libbeamform_b200.so` sits next to this file so it travels with the tree (git-ignored).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbeamform_b200.so")
SOURCES = ["capi.cu", "frames_kernel.cu", "das_kernel.cu", "sel_kernel.cu", "generic_kernel.cu", "srp_kernel.cu", "srp_tc_kernel.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "beamform_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("BF_NVCC_EXTRA", "").split()   # profiling builds only, e.g. -DBF_PHASE_TIMERS
    out = os.environ.get("BF_LIB_OUT", LIB)                # profiling builds: write a variant library elsewhere
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libbeamform_b200.so")
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
