"""Golden vectors made by RUNNING THE REFERENCE'S OWN PYTHON in the build container.

    python tests/golden/make_golden_pyref.py          # writes tests/golden/pyref_synth.npz

The reference is Python 2 and imports pysam / matplotlib, neither of which exists here, so its modules are loaded from
where they lie under /root/reference (nothing is copied into the repository) through a small compatibility loader:

  * source fixes that do not change meaning: `print x` -> `print(x)`, `except E, e` -> `except E as e`, `raise E(..), x`;
  * every `/` is rewritten (on the AST) into a helper with Python-2 semantics: floor division when both operands are
    integers (Python or numpy), true division otherwise; `map` / `zip` / `filter` return lists; `xrange`, `string.maketrans`;
  * `scipy.signal.gaussian` = `scipy.signal.windows.gaussian` (moved by scipy);
  * stubs for what the path under test does not compute: matplotlib, pyximport, `pysam.FastaFile` / `Samfile` (serve the
    synthetic genome), and `fragments.makeFragmentMat` -- the one Cython function that cannot be built without pysam; the
    fragment matrix of the synthetic reads comes from oracle.refalgo.make_fragment_mat, which is pinned bit-exactly on the
    reference's own tests/test_chunkmat2d.py;
  * `nucleoatac.multinomial_cov` = the reference's own .pyx compiled into oracle/_ref (oracle/build.py).

What then runs is the reference's code, unmodified in meaning: `OccChunk.process` (nucleoatac/Occupancy.py:241-248),
`NucChunk.process` (nucleoatac/NucleosomeCalling.py:328-340), `ChunkMat2D.get(flip=True)` (pyatac/chunkmat2d.py:21-54) and
`_vplotHelper` (pyatac/make_vplot.py:22-43)
on chunks of the synthetic workload (nucleoatac_b200/synth.py) -- inputs that are not the shipped example.  The vectors
pin oracle/ (tests/test_oracle_pyref.py) and, through the fixture, the device path (tests/test_gpu_pyref.py).
"""
import ast
import builtins
import glob
import importlib.util
import os
import re
import string
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
PWM_PATH = REF + "/pyatac/pwm/Human.PWM.txt"

# chunk index, length, density of the synthetic chunks that are scored (nucleoatac_b200/synth.make_chunk)
CASES = [(2, 10000, 0.25), (7, 3000, 0.25), (11, 2500, 0.02)]
# further cases for the oracle only (CPU test): chunk, length, density, Tn5 bias model on/off, VMat rows x columns, first size
CASES2 = [(3, 4000, 0.25, 0, 251, 251, 0), (4, 5000, 0.3, 1, 151, 101, 60), (6, 4000, 0.3, 0, 201, 151, 30)]
SEQ_MARGIN = 700


def py2div(a, b):
    def is_int(x):
        return isinstance(x, (bool, int, np.integer)) or (isinstance(x, np.ndarray) and np.issubdtype(x.dtype, np.integer))
    return a // b if is_int(a) and is_int(b) else a / b


class _Div(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.copy_location(ast.Call(func=ast.Name(id="__py2div__", ctx=ast.Load()), args=[node.left, node.right], keywords=[]), node)
        return node

    def visit_AugAssign(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            load = ast.parse(ast.unparse(node.target), mode="eval").body
            return ast.copy_location(ast.Assign(targets=[node.target], value=ast.Call(
                func=ast.Name(id="__py2div__", ctx=ast.Load()), args=[load, node.value], keywords=[])), node)
        return node


def _py2_text(text):
    out = []
    for line in text.split("\n"):
        m = re.match(r"^(\s*)print\s+(?!\()(.*)$", line)
        if m:
            line = "%sprint(%s)" % (m.group(1), m.group(2))
        line = re.sub(r"except\s+([\w\.]+)\s*,\s*(\w+)\s*:", r"except \1 as \2:", line)
        m = re.match(r"^(\s*raise\s+\w+\(.*\))\s*,\s*\w+\s*$", line)
        if m:
            line = m.group(1)
        out.append(line)
    return "\n".join(out)


def load_py2(modname, path):
    tree = _Div().visit(ast.parse(_py2_text(open(path).read()), filename=path))
    ast.fix_missing_locations(tree)
    m = types.ModuleType(modname)
    m.__file__ = path
    m.__dict__.update(__py2div__=py2div, xrange=range, map=lambda *a: list(builtins.map(*a)),
                      zip=lambda *a: list(builtins.zip(*a)), filter=lambda *a: list(builtins.filter(*a)))
    sys.modules[modname] = m
    exec(compile(tree, path, "exec"), m.__dict__)
    return m


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


GENOME, READS = {}, {}


def load_reference():
    import scipy.signal
    import scipy.signal.windows
    from oracle import refalgo as ra
    string.maketrans = str.maketrans
    scipy.signal.gaussian = scipy.signal.windows.gaussian
    mpl = _stub("matplotlib", use=lambda *a, **k: None)
    mpl.pyplot, mpl.cm = _stub("matplotlib.pyplot"), _stub("matplotlib.cm")
    _stub("pyximport", install=lambda **k: None)

    class FakeFasta:
        def __init__(self, path, mode="rb"):
            pass

        def fetch(self, chrom, start, end):
            return GENOME[chrom][start:end]

        references = property(lambda self: list(GENOME))
        lengths = property(lambda self: [len(v) for v in GENOME.values()])

        def close(self):
            pass
    _stub("pysam", FastaFile=FakeFasta, Samfile=FakeFasta, AlignmentFile=FakeFasta, tabix_compress=None, tabix_index=None, TabixFile=None)

    def makeFragmentMat(bamfile, chrom, start, end, lower, upper, atac=1):
        pos, tlen = READS[chrom]
        return ra.make_fragment_mat(pos, tlen, start, end, lower, upper, atac)
    fr = _stub("fragments", makeFragmentMat=makeFragmentMat, getInsertions=None, getStrandedInsertions=None,
               getAllFragmentSizes=None, getFragmentSizesFromChunkList=None)
    sys.modules["pyatac.fragments"] = fr
    _stub("pyatac").__path__ = [REF + "/pyatac"]
    _stub("nucleoatac").__path__ = [REF + "/nucleoatac"]
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "multinomial_cov*.so"))
    if not so:
        raise SystemExit("oracle/_ref is not built (python -m oracle.build)")
    spec = importlib.util.spec_from_file_location("multinomial_cov", so[0])
    mc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mc)
    sys.modules["multinomial_cov"] = sys.modules["nucleoatac.multinomial_cov"] = mc
    M = {}
    for pkg, names in (("pyatac", ["utils", "chunk", "seq", "bedgraph", "fragmentsizes", "tracks", "bias", "chunkmat2d", "VMat"]),
                       ("nucleoatac", ["Occupancy", "NucleosomeCalling"])):
        for n in names:
            M[n] = load_py2(pkg + "." + n, "%s/%s/%s.py" % (REF, pkg, n))
    M["bias"].pwm_parse = lambda name: name     # the PWM is given by its path (pkg_resources has no installed distribution)
    return M


def main():
    from nucleoatac_b200 import synth
    M = load_reference()
    wl = synth.Workload(251, 251)
    FS, Chunk = M["fragmentsizes"].FragmentSizes, M["chunk"].Chunk

    class Dist:
        pass
    d = Dist()
    d.nuc_fit, d.nfr_fit = FS(0, 251, vals=wl.nuc_probs.copy()), FS(0, 251, vals=wl.nfr_probs.copy())
    out = dict(cases=np.array(CASES, dtype=np.float64), seq_margin=np.array(SEQ_MARGIN))
    for ci, (k, length, density) in enumerate(CASES):
        s, e, pos, tlen, seq, s0 = synth.make_chunk(int(k), length=int(length), density=density, seq_margin=SEQ_MARGIN)
        GENOME.clear()
        READS.clear()
        GENOME["chrS"] = "N" * s0 + bytes(seq).decode() + "N" * 1000
        READS["chrS"] = (pos, tlen)
        op = M["Occupancy"].OccupancyParameters(d, 251, "synthetic.fa", PWM_PATH, bam="synthetic.bam")
        oc = M["Occupancy"].OccChunk(Chunk("chrS", s, e))
        oc.process(op)
        pk = sorted(oc.peaks.keys())
        p = "c%d_" % ci
        out.update({p + "occ_vals": oc.occ.vals, p + "occ_lower": oc.occ.lower_bound, p + "occ_upper": oc.occ.upper_bound,
                    p + "occ_smoothed_vals": oc.occ.smoothed_vals, p + "occ_smoothed_lower": oc.occ.smoothed_lower,
                    p + "occ_smoothed_upper": oc.occ.smoothed_upper, p + "occ_cov": oc.cov.vals,
                    p + "occ_peak_pos": np.array([oc.peaks[x].start for x in pk], dtype=np.int64),
                    p + "occ_peak_stats": np.array([[oc.peaks[x].occ, oc.peaks[x].occ_lower, oc.peaks[x].occ_upper, oc.peaks[x].reads] for x in pk],
                                                   dtype=np.float64).reshape(len(pk), 4),
                    p + "occ_nuc_dist": oc.getNucDist()})
        vm = M["VMat"].VMat(wl.vmat.copy(), wl.v_lower, wl.v_upper)
        npar = M["NucleosomeCalling"].NucParameters(vm, FS(0, wl.upper, vals=wl.fragmentsizes.copy()), "synthetic.bam", "synthetic.fa",
                                                    PWM_PATH, sd=10)
        nc = M["NucleosomeCalling"].NucChunk(Chunk("chrS", s, e))
        nc.process(npar)
        keys = sorted(nc.nuc_collection.keys())
        cols = ("z", "lr", "norm_signal", "nuc_signal", "nuc_cov", "nfr_cov", "fuzz", "weight", "fit_pos")
        out.update({p + "nuc_signal": nc.nuc_signal.vals, p + "nuc_background": nc.bias.vals, p + "nuc_norm_signal": nc.norm_signal.vals,
                    p + "nuc_smoothed": nc.smoothed.vals, p + "nuc_nuc_cov": nc.nuc_cov.vals, p + "nuc_nfr_cov": nc.nfr_cov.vals,
                    p + "nuc_call_pos": np.array([nc.nuc_collection[x].start for x in keys], dtype=np.int64),
                    p + "nuc_call_stats": np.array([[getattr(nc.nuc_collection[x], c) for c in cols] for x in keys], dtype=np.float64).reshape(len(keys), len(cols)),
                    p + "nuc_nonredundant": np.array(sorted(int(x) + s for x in nc.nonredundant), dtype=np.int64),
                    p + "nuc_redundant": np.array(sorted(int(x) + s for x in nc.redundant), dtype=np.int64)})
        print("case %d (chunk %d, %d bp, density %g): %d occupancy peaks, %d nucleosome calls" % (ci, k, length, density, len(pk), len(keys)))
    out["cases2"] = np.array(CASES2, dtype=np.float64)
    for ci, (k, length, density, use_bias, R, W, lower) in enumerate(CASES2):
        wl2 = synth.Workload(R, W, lower=lower)
        s, e, pos, tlen, seq, s0 = synth.make_chunk(int(k), length=int(length), density=density, seq_margin=SEQ_MARGIN)
        GENOME.clear()
        READS.clear()
        GENOME["chrS"] = "N" * s0 + bytes(seq).decode() + "N" * 1000
        READS["chrS"] = (pos, tlen)
        fasta = "synthetic.fa" if use_bias else None
        p = "d%d_" % ci
        if R == 251:   # the occupancy model of the synthetic workload is defined on sizes [0, 251)
            import pyatac.utils as pu
            M["Occupancy"].read_chrom_sizes_from_fasta = lambda f: {"chrS": len(GENOME["chrS"])}   # fasta=None: the reference crashes here
            op = M["Occupancy"].OccupancyParameters(d, 251, fasta, PWM_PATH, bam="synthetic.bam")
            oc = M["Occupancy"].OccChunk(Chunk("chrS", s, e))
            oc.process(op)
            pk = sorted(oc.peaks.keys())
            out.update({p + "occ_vals": oc.occ.vals, p + "occ_lower": oc.occ.lower_bound, p + "occ_upper": oc.occ.upper_bound,
                        p + "occ_smoothed_vals": oc.occ.smoothed_vals, p + "occ_cov": oc.cov.vals,
                        p + "occ_peak_pos": np.array([oc.peaks[x].start for x in pk], dtype=np.int64)})
        vm = M["VMat"].VMat(wl2.vmat.copy(), wl2.v_lower, wl2.v_upper)
        npar = M["NucleosomeCalling"].NucParameters(vm, FS(0, wl2.upper, vals=wl2.fragmentsizes.copy()), "synthetic.bam", fasta, PWM_PATH, sd=10)
        nc = M["NucleosomeCalling"].NucChunk(Chunk("chrS", s, e))
        nc.process(npar)
        keys = sorted(nc.nuc_collection.keys())
        out.update({p + "nuc_signal": nc.nuc_signal.vals, p + "nuc_background": nc.bias.vals, p + "nuc_norm_signal": nc.norm_signal.vals,
                    p + "nuc_smoothed": nc.smoothed.vals, p + "nuc_nuc_cov": nc.nuc_cov.vals,
                    p + "nuc_call_pos": np.array([nc.nuc_collection[x].start for x in keys], dtype=np.int64),
                    p + "nuc_call_zlr": np.array([[nc.nuc_collection[x].z, nc.nuc_collection[x].lr] for x in keys], dtype=np.float64).reshape(len(keys), 2)})
        print("case2 %d (chunk %d, %d bp, bias %d, VMat %dx%d from size %d): %d nucleosome calls" % (ci, k, length, use_bias, R, W, lower, len(keys)))
    # FragmentMixDistribution.modelNFR (nucleoatac/Occupancy.py:29-66): scipy brute + fmin, on two size distributions
    np.float = float                    # `np.float('inf')` of the reference (an alias numpy has dropped)
    for name, vals in (("nfr_synth", wl.fragmentsizes[:251] / wl.fragmentsizes[:251].sum()),
                       ("nfr_bumpy", (lambda v: v / v.sum())(np.abs(wl.fragmentsizes[:251] * (1 + 0.3 * np.random.RandomState(2).standard_normal(251)))))):
        fm = M["Occupancy"].FragmentMixDistribution(0, upper=251)
        fm.fragmentsizes = FS(0, 251, vals=vals.copy())
        fm.modelNFR()
        out[name + "_sizes"], out[name + "_nfr_fit"], out[name + "_nuc_fit"] = vals, fm.nfr_fit.get(), fm.nuc_fit.get()
    print("modelNFR: two fits")
    # ChunkList.read -> slop -> merge -> split (pyatac/chunk.py:101-207): what run_occ / run_nuc do with the BED (run_occ.py:83-90)
    import tempfile
    rng = np.random.RandomState(13)
    chroms = {"chrA": 60000, "chrB": 25000, "chrC": 9000}
    rows = []
    for c in ("chrA", "chrB", "chrC", "chrUnknown"):
        at = 0
        for _ in range(25):
            at += int(rng.randint(1, 1500))
            w = int(rng.randint(50, 2500))
            rows.append((c, at, at + w))                # sorted by start inside a chromosome, many overlaps
    with tempfile.NamedTemporaryFile("w", suffix=".bed", delete=False) as fh:
        for r in rows:
            fh.write("%s\t%d\t%d\n" % r)
        bed = fh.name
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cl = M["chunk"].ChunkList.read(bed, chromDict=chroms, min_offset=300, min_length=240)
    cl.slop(chroms, up=60, down=60)
    cl.merge()
    os.remove(bed)
    out["chunklist_bed"] = np.array([(list(chroms) + ["chrUnknown"]).index(c) for c, _, _ in rows] and
                                    [((list(chroms) + ["chrUnknown"]).index(c), a, b) for c, a, b in rows], dtype=np.int64)
    out["chunklist_merged"] = np.array([(list(chroms).index(c.chrom), c.start, c.end) for c in cl], dtype=np.int64)
    out["chunklist_split3"] = np.array([len(g) for g in cl.split(items=3)], dtype=np.int64)
    print("chunk list: %d BED rows -> %d merged chunks" % (len(rows), len(cl)))
    # ChunkMat2D.get with the strand flip (pyatac/chunkmat2d.py:41-54): integer matrices, odd and even first sizes
    rng = np.random.RandomState(7)
    CM = M["chunkmat2d"].ChunkMat2D
    flips = []
    for fi, (lower, upper, start, ncol, g0, g1, r0, r1) in enumerate(((0, 12, 100, 41, 101, 140, 0, 12), (3, 20, 50, 61, 55, 106, 3, 20),
                                                                     (4, 9, 10, 25, 11, 34, 4, 9), (1, 30, 1000, 51, 1002, 1049, 1, 30))):
        m = CM("chrS", start, start + ncol, lower, upper)
        m.mat = rng.randint(0, 9, size=(upper - lower, ncol)).astype(np.float64)
        got = m.get(lower=r0, upper=r1, start=g0, end=g1, flip=True)
        out["flip%d_mat" % fi], out["flip%d_out" % fi] = m.mat, got
        out["flip%d_args" % fi] = np.array([lower, upper, start, g0, g1, r0, r1], dtype=np.int64)
        flips.append(got.shape)
    out["n_flip"] = np.array(len(flips))
    # _vplotHelper (pyatac/make_vplot.py:22-43): the aggregate V-plot of a set of stranded sites, plain and --scale
    sys.modules["VMat"] = M["VMat"]
    mv = load_py2("pyatac.make_vplot", REF + "/pyatac/make_vplot.py")
    s, e, pos, tlen, seq, s0 = synth.make_chunk(5, length=6000, density=0.4, seq_margin=SEQ_MARGIN)
    READS.clear()
    READS["chrS"] = (pos, tlen)
    rng = np.random.RandomState(3)
    sites = [(int(a), int(a + w), st) for a, w, st in zip(rng.randint(s + 400, e - 400, 40), rng.randint(1, 9, 40), rng.choice(["+", "-", "*"], 40))]
    out["vplot_chunk"] = np.array([5, 6000, 0.4])
    out["vplot_sites"] = np.array([(a, b, {"+": 1, "-": -1, "*": 0}[st]) for a, b, st in sites], dtype=np.int64)
    for name, scale in (("vplot_plain", False), ("vplot_scaled", True)):
        chunks = [Chunk("chrS", a, b, strand=st) for a, b, st in sites]
        out[name] = mv._vplotHelper((chunks, mv._VplotParams(60, 30, 250, "synthetic.bam", 1, scale)))
    print("vplot: %d sites, %g fragments in the plain plot" % (len(sites), out["vplot_plain"].sum()))
    np.savez_compressed(os.path.join(HERE, "pyref_synth.npz"), **out)
    print("wrote", os.path.join(HERE, "pyref_synth.npz"), os.path.getsize(os.path.join(HERE, "pyref_synth.npz")), "bytes")


if __name__ == "__main__":
    main()
