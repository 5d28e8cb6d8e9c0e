"""Build recipe for the oracle's native pieces (TEST INFRASTRUCTURE).

1. ``oracle/libmcov.so``  <- ``oracle/mcov.c`` (our C restatement), gcc.
2. ``oracle/_ref/multinomial_cov*.so`` <- the REFERENCE's own
   ``/root/reference/nucleoatac/multinomial_cov.pyx`` compiled with Cython, when
   the reference tree is present (build container only; the GPU box uses the
   prebuilt file that travels in ``oracle/_ref/``).  The only edit is the
   one-token py3/numpy-2 fix on line 14 (``np.float`` -> ``np.float64``), applied
   to a scratch copy under /tmp -- no reference source enters the repository and
   ``oracle/_ref/`` is git-ignored.

Run: ``python -m oracle.build``.
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PYX = "/root/reference/nucleoatac/multinomial_cov.pyx"


def build_mcov(force=False):
    src = os.path.join(HERE, "mcov.c")
    out = os.path.join(HERE, "libmcov.so")
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", out, src])
    return out


def build_ref(force=False):
    """Compile the reference's multinomial_cov.pyx into oracle/_ref (if the tree exists)."""
    outdir = os.path.join(HERE, "_ref")
    have = glob.glob(os.path.join(outdir, "multinomial_cov*.so"))
    if have and not force:
        return have[0]
    if not os.path.exists(REF_PYX):
        return None
    os.makedirs(outdir, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="nb200_ref_")
    try:
        with open(REF_PYX) as fh:
            text = fh.read()
        text = text.replace("DTYPE = np.float\n", "DTYPE = np.float64\n")
        with open(os.path.join(tmp, "multinomial_cov.pyx"), "w") as fh:
            fh.write(text)
        setup = (
            "from setuptools import setup, Extension\n"
            "from Cython.Build import cythonize\n"
            "import numpy as np\n"
            "setup(script_args=['build_ext','--inplace'], ext_modules=cythonize("
            "[Extension('multinomial_cov',['multinomial_cov.pyx'],include_dirs=[np.get_include()])],"
            "language_level=2, quiet=True))\n")
        with open(os.path.join(tmp, "setup_ref.py"), "w") as fh:
            fh.write(setup)
        subprocess.check_call([sys.executable, "setup_ref.py"], cwd=tmp,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        built = glob.glob(os.path.join(tmp, "multinomial_cov*.so"))
        if not built:
            return None
        dst = os.path.join(outdir, os.path.basename(built[0]))
        shutil.copy(built[0], dst)
        return dst
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    print("libmcov:", build_mcov(force=True))
    print("_ref   :", build_ref(force=True))


if __name__ == "__main__":
    main()
