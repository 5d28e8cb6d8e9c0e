"""Build libnucleo_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache).

    python -m nucleoatac_b200.build [--force]

The shared library sits next to this file so that it travels with the repository snapshot to
the GPU box; it is git-ignored.  Each .cu is compiled to an object (in parallel) and linked
with the static CUDA runtime; NCCL is dlopen'ed at run time, nothing else is linked.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
_TAG = os.environ.get("NB200_BUILD_TAG", "")   # developer kernel variants: own object directory and library name
OBJ = os.path.join(HERE, "build" + ("_" + _TAG if _TAG else ""))
LIB = os.path.join(HERE, "libnucleo_b200%s.so" % ("_" + _TAG if _TAG else ""))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["nb200_ctx.cu", "nb200_batch.cu", "nb200_occ.cu", "nb200_nuc.cu", "nb200_prims.cu", "nb200_xcor_tc.cu", "nb200_hostfmt.cu", "nb200_hostio.cu", "nb200_pyatac.cu", "nb200_bamio.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--fmad=true", "-Xptxas", "-v"]


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "nucleo_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _extra_defs():
    """Developer kernel variants: NB200_NVCC_DEFS="-DTC_EPI_MODE=2 -DTC_N=224" (with NB200_LIB for a separate output file)."""
    return os.environ.get("NB200_NVCC_DEFS", "").split()


def _compile(src, force, hdr_mtime, log):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    path = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(path), hdr_mtime):
        return obj
    res = subprocess.run([NVCC] + FLAGS + _extra_defs() + ["-c", path, "-o", obj], capture_output=True, text=True)
    log.append((src, res.stderr))
    if res.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s" % (src, res.stderr))
    return obj


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdr_mtime = _deps_mtime()
    log = []
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, hdr_mtime, log), SOURCES))
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        res = subprocess.run([NVCC, "-shared", "-o", LIB] + objs + ["-cudart", "static", "-ldl", "-lpthread", "-lz"],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s" % res.stderr)
    if verbose:
        for src, err in log:
            sys.stderr.write("== %s\n%s\n" % (src, err))
    with open(os.path.join(OBJ, "ptxas.log"), "a" if not force else "w") as fh:
        for src, err in log:
            fh.write("== %s\n%s\n" % (src, err))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
