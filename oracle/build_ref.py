"""Builds the reference's own native module -- the Cython Viterbi tracker
pyfasst/SeparateLeadStereo/tracking/_tracking.pyx -- from the source where it lies under
/root/reference into oracle/_ref/ (git-ignored; it travels to the GPU box with the snapshot).
TEST INFRASTRUCTURE ONLY: tests/ and oracle/make_golden.py use it to pin oracle/viterbi_oracle.py.

The source is compiled from a temporary copy with two type-NAME substitutions that current
Cython / NumPy need (`np.int_t` -> `np.int64_t`, `dtype=np.int` -> `dtype=np.int64`: the same
64-bit integers on this platform); no arithmetic is touched and nothing of the reference is
copied into the repository.  The rest of the reference is Python 2 and cannot be built.

    python oracle/build_ref.py
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PYFASST_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
PYX = os.path.join(REF, "pyfasst", "SeparateLeadStereo", "tracking", "_tracking.pyx")

SETUP = """
from setuptools import setup, Extension
from Cython.Build import cythonize
import numpy
setup(ext_modules=cythonize([Extension("_tracking", ["_tracking.pyx"],
                                       include_dirs=[numpy.get_include()])], language_level=2))
"""


def built():
    return sorted(glob.glob(os.path.join(OUT, "_tracking*.so")))


def build(force=False):
    """Returns the path of oracle/_ref/_tracking*.so, building it if the reference is present;
    None when neither the reference nor a previous build is available."""
    have = built()
    if have and not force:
        return have[0]
    if not os.path.exists(PYX):
        return None
    tmp = tempfile.mkdtemp(prefix="pyfasst_ref_tracking_")
    try:
        with open(PYX) as fh:
            src = fh.read()
        src = src.replace("dtype=np.int)", "dtype=np.int64)").replace("np.int_t", "np.int64_t")
        with open(os.path.join(tmp, "_tracking.pyx"), "w") as fh:
            fh.write(src)
        with open(os.path.join(tmp, "setup.py"), "w") as fh:
            fh.write(SETUP)
        p = subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=tmp,
                           capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("cython build of the reference tracker failed:\n" + p.stdout[-2000:]
                               + p.stderr[-2000:])
        os.makedirs(OUT, exist_ok=True)
        for so in glob.glob(os.path.join(tmp, "_tracking*.so")):
            shutil.copy(so, OUT)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return built()[0]


def load():
    """The compiled reference module, or None if it is not available."""
    path = build()
    if path is None:
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("_tracking", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
