"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/nucleo_b200.h declares, the ctypes structs match the header, and creating a context without a
GPU fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from nucleoatac_b200 import _lib
    return _lib.load()


def header_functions():
    h = open(os.path.join(ROOT, "include", "nucleo_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(nb200_[a-z0-9_]+)\s*\(", h)))


def test_every_declared_symbol_is_exported(lib):
    from nucleoatac_b200 import _lib
    decl = header_functions()
    assert len(decl) >= 50
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert not [d for d in decl if d not in exported]
    assert sorted(_lib.EXPORTS) == decl  # the ctypes stub binds exactly the header's surface
    for name in decl:
        assert getattr(lib, name).argtypes is not None, name


def test_struct_layouts_match_header():
    from nucleoatac_b200 import _lib
    h = open(os.path.join(ROOT, "include", "nucleo_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    for cname, cls in (("nb200_occ_params", _lib.OccParams), ("nb200_nuc_params", _lib.NucParams), ("nb200_batch", _lib.Batch),
                       ("nb200_occ_out", _lib.OccOut), ("nb200_nuc_out", _lib.NucOut), ("nb200_occ_out32", _lib.OccOut32),
                       ("nb200_nuc_out32", _lib.NucOut32)):
        body = re.search(r"typedef struct \{([^{}]*)\} " + cname + ";", h).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*$", part.strip())[0])
        assert names == [f[0] for f in cls._fields_], cname
    # the float32 variants differ from the float64 ones in the type of the per-position tracks only
    for c64, c32, ntr in ((_lib.OccOut, _lib.OccOut32, 7), (_lib.NucOut, _lib.NucOut32, 6)):
        for i, ((n64, t64), (n32, t32)) in enumerate(zip(c64._fields_, c32._fields_)):
            assert n64 == n32 and (t32 is _lib.c_float_p if i < ntr else t32 is t64), n64
        body = re.search(r"typedef struct \{([^{}]*)\} nb200_%s_out32;" % ("occ" if ntr == 7 else "nuc"), h).group(1)
        assert len(re.findall(r"\bfloat \*", body)) >= 1 and "double *" in body


def test_no_cpu_fallback(lib):
    """Without a CUDA device nb200_ctx_create must fail with a message (this container has no GPU)."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    h = C.c_void_p()
    st = lib.nb200_ctx_create(0, C.byref(h))
    assert st != 0 and not h.value
    assert b"no CPU fallback" in lib.nb200_last_error(None)
    from nucleoatac_b200.engine import Engine
    from nucleoatac_b200._lib import NB200Error
    with pytest.raises(NB200Error):
        Engine(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "nucleoatac_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
