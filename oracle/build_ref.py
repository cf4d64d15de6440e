"""Build the UNMODIFIED reference MAS kernel into oracle/_ref/ (test infrastructure only).

The reference's only native component is the Cython file
/root/reference/src/model/monotonic_align/core.pyx (core.pyx:9-45).  Its own build
(src/model/monotonic_align/setup.py:7-11) is `cythonize` + the default C compiler with no
OpenMP flag, so `prange` is a serial loop there.  This script does the same two steps by
hand, reading the .pyx where it lies and writing ONLY under oracle/_ref/ (git-ignored, but
shipped to the GPU box with the repo snapshot):

    oracle/_ref/core.c                 Cython-generated C (build product, never committed)
    oracle/_ref/serial/core*.so        reference-faithful build (no -fopenmp, -O2)
    oracle/_ref/omp/core*.so           same source, -fopenmp -O3 (the "Cython/OpenMP" baseline)

Nothing is copied from the reference into tracked files.  When /root/reference is absent
(e.g. on the GPU box) the script is a no-op and the prebuilt files are used as they are.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PYX = "/root/reference/src/model/monotonic_align/core.pyx"
OUT = os.path.join(HERE, "_ref")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"


def ext_suffix() -> str:
    return sysconfig.get_config_var("EXT_SUFFIX") or ".so"


def ref_so(kind: str) -> str:
    return os.path.join(OUT, kind, "core" + ext_suffix())


def have_ref(kind: str = "serial") -> bool:
    return os.path.exists(ref_so(kind))


def build(force: bool = False, verbose: bool = False) -> bool:
    """Returns True when oracle/_ref holds both builds afterwards."""
    if have_ref("serial") and have_ref("omp") and not force:
        return True
    if not os.path.exists(REF_PYX):
        return have_ref("serial") and have_ref("omp")
    try:
        import numpy  # noqa: F401
        import Cython  # noqa: F401
    except Exception:
        return False
    import numpy as np

    os.makedirs(OUT, exist_ok=True)
    c_file = os.path.join(OUT, "core.c")
    run = lambda cmd: subprocess.run(  # noqa: E731
        cmd, check=True, stdout=None if verbose else subprocess.DEVNULL,
        stderr=None if verbose else subprocess.DEVNULL)
    run([sys.executable, "-m", "cython", "-3", REF_PYX, "-o", c_file])
    inc = ["-I" + sysconfig.get_paths()["include"], "-I" + np.get_include()]
    common = ["-shared", "-fPIC", "-fwrapv", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"]
    for kind, flags in (("serial", ["-O2"]), ("omp", ["-O3", "-fopenmp"])):
        os.makedirs(os.path.join(OUT, kind), exist_ok=True)
        run([GCC, *common, *flags, *inc, c_file, "-o", ref_so(kind)])
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv, verbose=True)
    print("oracle/_ref:", "built" if ok else "unavailable")
