"""Build the native library (hand-written sm_100a CUDA behind the C ABI of include/ifk.h).

    python -m inverse_flow_b200.build [--force] [--verbose]

Produces inverse_flow_b200/lib/libifk_b200.so IN-TREE (git-ignored, shipped to the GPU box
with the repository snapshot).  nvcc cross-compiles for sm_100a without a GPU.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libifk_b200.so")
SOURCES = ["ifk_api.cu", "ifk_env.cu", "ifk_comm.cu", "ifk_microbench.cu", "ifk_solve_wave.cu", "ifk_solve_split.cu", "ifk_prepare.cu", "ifk_solve.cu", "ifk_solve_v1.cu", "ifk_solve_v2.cu",
           "ifk_solve_v4.cu", "ifk_solve_stream.cu", "ifk_solve_window.cu", "ifk_solve_shfl.cu", "ifk_conv.cu", "ifk_bwd_weight.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--use_fast_math=false",
]


def _host_compiler():
    """nvcc's host compiler: $CXX, else the first g++ on PATH"""
    import shutil
    return os.environ.get("CXX") or shutil.which("g++") or "g++"


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "ifk.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force=False, verbose=False):
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    objdir = os.path.join(LIB_DIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [_nvcc(), "-ccbin", _host_compiler(), *flags, "-I", os.path.join(ROOT, "include"), "-I", CSRC,
               "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                                 text=True)))
    objs = []
    log = []
    for src, obj, pr in procs:
        out, _ = pr.communicate()
        log.append("== %s\n%s" % (src, out))
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, out))
        objs.append(obj)
    with open(os.path.join(LIB_DIR, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    cmd = [_nvcc(), "-ccbin", _host_compiler(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
           "-o", LIB_PATH, *objs]
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    path = build_native(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
