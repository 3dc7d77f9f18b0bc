#!/usr/bin/env python
"""Build the UNMODIFIED reference (Cython -> C) into oracle/_ref/.   TEST INFRASTRUCTURE ONLY.

The reference's hot path lives in five Cython modules (reference setup.py:39-83).
Its own `setup.py build_ext --inplace` writes into the source tree, and
/root/reference is read-only, so the recipe is:

  1. copy /root/reference/optical_networking_gym to a scratch dir under /tmp,
  2. cythonize + compile the five .pyx files there with the reference's release
     directives (setup.py:31-34) and flags `-O3 -ffast-math`
     (`-march=native` of setup.py:29 is dropped on purpose: the built .so has to
     run on the GPU box, whose host CPU is not this container's),
  3. install the result (compiled .so + the package's pure-Python modules, no
     .pyx/.c) into oracle/_ref/optical_networking_gym/, and the topology input
     files (examples/topologies: data) into oracle/_ref/topologies/.

oracle/_ref/ is git-ignored (never committed) but NOT gpurun-ignored, so the
compiled reference travels to the GPU box, where /root/reference does not exist.
Nothing under oracle/ is imported by the product package.

Usage:  python oracle/build_ref.py [--force]
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("QRMSA_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

SETUP = r'''
import numpy as np
from setuptools import setup, Extension
from Cython.Build import cythonize

mods = ["utils", "core.osnr", "topology", "envs.rmsa", "envs.qrmsa"]
exts = [
    Extension(
        "optical_networking_gym." + m,
        ["optical_networking_gym/" + m.replace(".", "/") + ".pyx"],
        include_dirs=[np.get_include()],
        define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")],
        extra_compile_args=["-O3", "-ffast-math"],
        extra_link_args=["-O3"],
    )
    for m in mods
]
setup(
    name="optical_networking_gym",
    ext_modules=cythonize(
        exts,
        compiler_directives=dict(language_level="3", boundscheck=False, wraparound=False,
                                 nonecheck=False, cdivision=True),
        nthreads=0,
    ),
)
'''


def ref_available() -> bool:
    return bool(glob.glob(os.path.join(OUT, "optical_networking_gym", "envs", "qrmsa*.so"))) and os.path.isdir(
        os.path.join(OUT, "topologies"))


def build(force: bool = False) -> bool:
    """Returns True when oracle/_ref holds a usable compiled reference."""
    if ref_available() and not force:
        return True
    if not os.path.isdir(os.path.join(REF, "optical_networking_gym")):
        return False  # GPU box: only the prebuilt files are used
    tmp = tempfile.mkdtemp(prefix="qrmsa_refbuild_")
    try:
        shutil.copytree(os.path.join(REF, "optical_networking_gym"), os.path.join(tmp, "optical_networking_gym"))
        with open(os.path.join(tmp, "setup_oracle.py"), "w") as f:
            f.write(SETUP)
        env = dict(os.environ)
        subprocess.run(
            [sys.executable, "setup_oracle.py", "build_ext", "--inplace", "-j", str(os.cpu_count() or 1)],
            cwd=tmp, check=True, env=env, stdout=subprocess.DEVNULL,
        )
        dst = os.path.join(OUT, "optical_networking_gym")
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(
            os.path.join(tmp, "optical_networking_gym"), dst,
            ignore=shutil.ignore_patterns("*.pyx", "*.c", "*.pxd", "__pycache__", "*.html"),
        )
        # topology input files (data, not source) so that the compiled reference can be timed on the GPU
        # box, where /root/reference does not exist; git-ignored like the rest of oracle/_ref
        tdst = os.path.join(OUT, "topologies")
        if os.path.isdir(tdst):
            shutil.rmtree(tdst)
        shutil.copytree(os.path.join(REF, "examples", "topologies"), tdst)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return ref_available()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "NOT built (reference sources not present)")
    sys.exit(0 if ok else 1)
