"""Compile the reference's own CPU inverse solver into oracle/_ref/.

The only CPU implementation of this path that the reference ships is the Cython solver
inf/utils/fastflow_inverse/solve_parallel_mc.pyx:77-126 (float64, inverse only).  This
script compiles it FROM WHERE IT LIES under the reference checkout -- nothing is copied
into the repository; the generated C file and the two shared objects go to the
git-ignored oracle/_ref/ and travel to the GPU box with the snapshot.

  solve_parallel_mc.so      as the reference builds it (fastflow_inverse/setup.py:1-5:
                            no -fopenmp, so its prange is serial -> 1 core)
  solve_parallel_mc_omp.so  same source with -fopenmp (num_threads=30 hard-coded at
                            .pyx:104, capped by the host)

The reference's setup.py is not run (it omits numpy's include dir and would write next to
the read-only source); this is the short recipe the task allows.
"""
import argparse
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
PYX = "inf/utils/fastflow_inverse/solve_parallel_mc.pyx"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    pyx = os.path.join(args.ref, PYX)
    if not os.path.exists(pyx):
        print("reference checkout not present (%s): nothing to build" % pyx)
        return 0
    import numpy

    os.makedirs(OUT, exist_ok=True)
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    inc = [sysconfig.get_paths()["include"], numpy.get_include()]
    ext = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    for name, extra in (("solve_parallel_mc", []), ("solve_parallel_mc_omp", ["-fopenmp"])):
        cfile = os.path.join(OUT, name + ".c")
        # --module-name makes PyInit_<name> match the file name of each variant
        subprocess.check_call([sys.executable, "-m", "cython", "-3", "--module-name", name,
                               "-o", cfile, pyx])
        so = os.path.join(OUT, name + ext)
        cmd = [gcc, "-O2", "-fPIC", "-shared", "-w", "-DNPY_NO_DEPRECATED_API=0"]
        cmd += ["-I" + i for i in inc] + extra + [cfile, "-o", so]
        subprocess.check_call(cmd)
        os.remove(cfile)
        print("built", so)
    shutil.rmtree(os.path.join(OUT, "__pycache__"), ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
