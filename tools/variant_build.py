"""Compile-time variants of one translation unit as separate libraries (development aid).

    python tools/variant_build.py NAME=-DMACRO[,-DMACRO2] ...   [--tu ifk_solve_wave.cu]

Each variant is the in-tree build (inverse_flow_b200/lib/obj/*.o) with ONE translation unit recompiled with
the given macros, linked into inverse_flow_b200/lib/exp/libifk_NAME.so; run a tool against it with
IFK_LIBRARY=<that path>.  ptxas' scheduling of the wavefront loop moves by a few percent with any change
to the surrounding code, so alternatives are measured side by side in one GPU session rather than argued.
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inverse_flow_b200 import build as B  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    tu = "ifk_solve_wave.cu"
    if "--tu" in sys.argv:
        tu = sys.argv[sys.argv.index("--tu") + 1]
        args.remove(tu)
    objdir = os.path.join(B.LIB_DIR, "obj")
    expdir = os.path.join(B.LIB_DIR, "exp")
    os.makedirs(expdir, exist_ok=True)
    flags = [f for f in B.NVCC_FLAGS if f != "--use_fast_math=false"]
    procs = []
    for a in args:
        name, _, macros = a.partition("=")
        obj = os.path.join(expdir, "%s_%s.o" % (tu.replace(".cu", ""), name))
        cmd = [B._nvcc(), "-ccbin", B._host_compiler(), *flags, *[m for m in macros.split(",") if m],
               "-I", os.path.join(B.ROOT, "include"), "-I", B.CSRC, "-c", os.path.join(B.CSRC, tu), "-o", obj]
        procs.append((name, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, obj, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise SystemExit("nvcc failed for %s:\n%s" % (name, out))
        with open(os.path.join(expdir, "ptxas_%s.log" % name), "w") as f:
            f.write(out)
        others = [os.path.join(objdir, s.replace(".cu", ".o")) for s in B.SOURCES if s != tu]
        lib = os.path.join(expdir, "libifk_%s.so" % name)
        subprocess.check_call([B._nvcc(), "-ccbin", B._host_compiler(), "-shared", "-gencode",
                               "arch=compute_100a,code=sm_100a", "-o", lib, obj, *others])
        os.remove(obj)
        print(lib)


if __name__ == "__main__":
    main()
