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
SOURCES = ["capi.cu", "frames_kernel.cu", "das_kernel.cu", "sel_kernel.cu", "sel_stream_kernel.cu", "generic_kernel.cu", "phase_n_kernel.cu", "mcra_kernel.cu", "srp_kernel.cu", "srp_tc_kernel.cu"]
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
    """Every .cu file is compiled to an object in parallel (one nvcc per file), then linked into the shared library."""
    if not force and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    import tempfile
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("BF_NVCC_EXTRA", "").split()   # profiling builds only, e.g. -DBF_PHASE_TIMERS
    out = os.environ.get("BF_LIB_OUT", LIB)                # profiling builds: write a variant library elsewhere
    cflags = [f for f in NVCC_FLAGS if f != "-shared"] + extra + (["-Xptxas", "-v"] if verbose else [])
    with tempfile.TemporaryDirectory(prefix="bf_build_") as td:
        def one(src):
            obj = os.path.join(td, src.replace(".cu", ".o"))
            r = subprocess.run([nvcc] + cflags + ["-c", "-o", obj, os.path.join(CSRC, src)], capture_output=True, text=True)
            return src, obj, r
        with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
            results = list(ex.map(one, SOURCES))
        failed = [src for src, _, r in results if r.returncode != 0]
        if verbose or failed:
            for _, _, r in results:
                sys.stderr.write(r.stdout + r.stderr)
        if failed:
            raise RuntimeError("nvcc failed building " + ", ".join(failed))
        r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + [obj for _, obj, _ in results],
                           capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed linking libbeamform_b200.so")
    return out


def build_tools():
    """tools/bf_offline: the C++ offline file driver (WAV / yaml / launch file in, WAV out) on top of the C ABI."""
    root = os.path.join(HERE, "..")
    out = os.path.join(root, "tools", "bf_offline")
    src = os.path.join(root, "tools", "bf_offline.cpp")
    if os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(src), os.path.getmtime(LIB)):
        return out
    r = subprocess.run([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-o", out, src, "-L" + HERE, "-lbeamform_b200",
                        "-Wl,-rpath,$ORIGIN/../beamform_b200"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building tools/bf_offline")
    return out


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    build_tools()
    print(LIB)
