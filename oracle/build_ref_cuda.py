"""Compile the reference's own CUDA extension `inv_conv_with_bp` into oracle/_ref/ for sm_100a.

This is the GPU implementation the hot path replaces
(inf/utils/inv_conv_cuda/inv_conv_with_bp_general.cpp:115-120 binds inverse / forward / dy / dw;
the kernels are in inv_conv_with_bp_kernel_general.cu).  It is test infrastructure like the rest
of oracle/: the GPU parity tests compare against it where its literal semantics are well defined
(groups = C/4 ... see tests/test_reference_cuda_gpu.py) and tools/ref_gpu_bench.py times it
beside our kernels on the same B200.  Nothing under inverse_flow_b200/ may import it.

Nothing is copied into the repository: the two source files are read where they lie under the
reference checkout, the one change current PyTorch needs -- `AT_DISPATCH_FLOATING_TYPES(
input.type(), ...)` takes a ScalarType now, so `.type()` becomes `.scalar_type()` at
.cu:112,246,428,465,679 -- is applied to a scratch copy inside the (git-ignored) build
directory, which is deleted after the build; only the shared object stays in oracle/_ref/.
The reference's setup.py (inv_conv_cuda/setup.py:1-16) is not run: it would write next to the
read-only sources and uses torch's arch list instead of sm_100a.
"""
import argparse
import os
import re
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC_DIR = "inf/utils/inv_conv_cuda"
CPP = "inv_conv_with_bp_general.cpp"
CU = "inv_conv_with_bp_kernel_general.cu"
NAME = "inv_conv_with_bp_ref"


def so_path():
    ext = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    return os.path.join(OUT, NAME + ext)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    args = ap.parse_args()
    cpp = os.path.join(args.ref, SRC_DIR, CPP)
    cu = os.path.join(args.ref, SRC_DIR, CU)
    if not (os.path.exists(cpp) and os.path.exists(cu)):
        print("reference checkout not present (%s): nothing to build" % cu)
        return 0
    so = so_path()
    if os.path.exists(so) and not args.force and os.path.getmtime(so) >= max(os.path.getmtime(cpp), os.path.getmtime(cu)):
        print("up to date:", so)
        return 0

    import torch
    from torch.utils import cpp_extension as ce

    build = os.path.join(OUT, "_cuda_build")
    shutil.rmtree(build, ignore_errors=True)
    os.makedirs(build)
    with open(cu) as f:
        text = f.read()
    patched, n = re.subn(r"AT_DISPATCH_FLOATING_TYPES\(\s*input\.type\(\)", "AT_DISPATCH_FLOATING_TYPES(input.scalar_type()", text)
    print("patched %d AT_DISPATCH sites (.type() -> .scalar_type())" % n)
    cu_tmp = os.path.join(build, CU)
    with open(cu_tmp, "w") as f:
        f.write(patched)

    inc = ["-I" + p for p in ce.include_paths("cuda")] + ["-I" + sysconfig.get_paths()["include"]]
    defs = ["-DTORCH_EXTENSION_NAME=" + NAME, "-DTORCH_API_INCLUDE_EXTENSION_H",
            "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    o_cpp, o_cu = os.path.join(build, "binding.o"), os.path.join(build, "kernels.o")
    subprocess.check_call([gxx, "-O2", "-fPIC", "-std=c++17", "-w", "-c", cpp, "-o", o_cpp] + inc + defs)
    subprocess.check_call(["nvcc", "-ccbin", gxx, "-O2", "-std=c++17", "-w", "-Xcompiler", "-fPIC",
                           "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                           "-c", cu_tmp, "-o", o_cu] + inc + defs)
    libs = ["-L" + p for p in ce.library_paths("cuda")]
    rpath = ["-Wl,-rpath," + p for p in ce.library_paths("cuda")]
    subprocess.check_call([gxx, "-shared", o_cpp, o_cu, "-o", so] + libs + rpath +
                          ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"])
    shutil.rmtree(build, ignore_errors=True)
    print("built", so)
    return 0


if __name__ == "__main__":
    sys.exit(main())
