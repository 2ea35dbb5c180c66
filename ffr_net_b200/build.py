"""Builds libffr_sm100.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot. Rebuilds only when a source
is newer than the library.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libffr_sm100.so")
PROBE_LIB_PATH = os.path.join(LIB_DIR, "libffr_sm100_probe.so")     # debug probes / micro-benchmarks, not the product

SOURCES = ["api.cu", "conv_gemm.cu", "backbone_kernels.cu", "recnet_kernels.cu", "train_kernels.cu", "bn_train_kernels.cu", "recnet_train_kernels.cu", "loss_kernels.cu", "head_kernels.cu", "scoring_kernels.cu", "host.cpp"]
PROBE_SOURCES = ["probe.cu", "host.cpp"]
HEADERS = ["ptx.cuh", "conv_gemm.cuh", "host.h", "kernels.h", os.path.join("..", "..", "include", "ffr_sm100.h"),
           os.path.join("..", "..", "include", "ffr_sm100_probe.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale():
    if not (os.path.exists(LIB_PATH) and os.path.exists(PROBE_LIB_PATH)):
        return True
    t = min(os.path.getmtime(LIB_PATH), os.path.getmtime(PROBE_LIB_PATH))
    deps = [os.path.join(CSRC, s) for s in SOURCES + PROBE_SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu into objects (parallel) and link the shared library. Returns the library path."""
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    objs = []
    for src in SOURCES + [p for p in PROBE_SOURCES if p not in SOURCES]:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(obj_dir, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ["-x", "cu", "-c", path, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write("== %s ==\n%s\n" % (src, out))
        if p.returncode != 0:
            failed = True
    if failed:
        raise RuntimeError("nvcc failed")
    # cudart is linked statically (nvcc default); no -lcuda: the driver entry point is looked up at run time
    def obj(src):
        return os.path.join(obj_dir, os.path.splitext(src)[0] + ".o")
    for lib_path, srcs in ((LIB_PATH, SOURCES), (PROBE_LIB_PATH, PROBE_SOURCES)):
        cmd = [nvcc, "-shared", "-o", lib_path + ".tmp"] + [obj(s_) for s_ in srcs] + ["-gencode", "arch=compute_100a,code=sm_100a"]
        subprocess.check_call(cmd)
        os.replace(lib_path + ".tmp", lib_path)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
