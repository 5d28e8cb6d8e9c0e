"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the pyatac tools either side of the scoring path:
the per-site V-plot matrix with the strand flip, the coverage helper and the insertion helper, literally as the
reference computes them (dense float64 matrices).

Parity status: PINNED where the reference holds data -- `ins_chunk` reproduces the shipped example_results/example.ins.bedgraph.gz
value for value (the track `nucleoatac nfr` wrote with the same InsertionTrack.calculateInsertions), `cov_chunk` the 160 read
counts of the shipped example.occpeaks.bed.gz (OccChunk.cov at the peaks = the same CoverageTrack.calculateCoverage), `mat_get`
without flip the reference's own known answer (tests/test_chunkmat2d.py:12-17); see tests/test_oracle_pyatac.py.  The reference
holds NO vector for the strand flip of ChunkMat2D.get (chunkmat2d.py:41-54) nor for `pyatac vplot` as a whole (its only pyatac
test is a smoke run of `pyatac sizes`, tests/test_cli.py:38-46; example.VMat comes from the bundled S. cer V-plot, not from
example.bam): the flip and vplot_site below are pinned on vectors made by running the reference's own ChunkMat2D.get(flip=True) and _vplotHelper in the build
container (tests/golden/make_golden_pyref.py -> tests/golden/pyref_synth.npz, tests/test_oracle_pyref.py) and checked by its
mirror-image property.
"""
import numpy as np

from . import refalgo as ra


def mat_get(mat, mat_start, mat_lower, lower, upper, start, end, flip=False):
    """ChunkMat2D.get, pyatac/chunkmat2d.py:21-54."""
    y1, y2 = lower - mat_lower, upper - mat_lower
    x1, x2 = start - mat_start, end - mat_start
    if not flip:
        return mat[y1:y2, x1:x2]
    if x1 < 1 or x2 > mat.shape[1] or y1 < 0 or y2 > mat.shape[0]:
        raise Exception("Looks like dimensions from get probaby don't match Mat")
    ncol = x2 - x1
    if ncol % 2 == 0:
        raise Exception("Can only flip mat if the width is odd!")
    new = np.zeros((y2 - y1, ncol))
    for j in range(y1, y2):
        if (j + mat_lower) % 2 == 1:
            new[j, :] = mat[j, x1:x2][::-1]
        else:
            new[j, :] = mat[j, (x1 - 1):x2][::-1][1:]
    return new


def center(start, end, strand):
    """Chunk.center, pyatac/chunk.py:41-54 (Python-2 integer division)."""
    half = (end - start) // 2
    if strand == "-":
        e = end - half
        return e - 1, e
    s = start + half
    return s, s + 1


def vplot_site(pos, tlen, start, end, strand, flank, lower, upper, atac=True, scale=False):
    """_vplotHelper for one region, pyatac/make_vplot.py:29-36; pos/tlen = the reads the BAM fetch returns."""
    s, e = center(start, end, strand)
    m_start, m_end = s - flank - 1, e + flank
    mat = ra.make_fragment_mat(pos, tlen, m_start, m_end, lower, upper, atac)
    add = mat_get(mat, m_start, lower, lower, upper, s - flank, e + flank, flip=(strand == "-"))
    if scale:
        with np.errstate(invalid="ignore", divide="ignore"):
            add = add / np.sum(add)
    return add


def cov_chunk(pos, tlen, start, end, lower, upper, window, scale, atac=True):
    """_covHelper, pyatac/get_cov.py:22-31."""
    offset = window // 2
    mat = ra.make_fragment_mat(pos, tlen, start - offset, end + offset, lower, upper, atac)
    vals = ra.calculate_coverage(mat, start - offset, lower, start, lower, upper, window)
    return vals * (scale / float(window))


def ins_chunk(pos, tlen, start, end, lower, upper, smooth=None, atac=True):
    """_insHelper / _insHelperSmooth, pyatac/get_ins.py:20-46 -> (track start, values)."""
    if not smooth:
        return start, ra.get_insertions(pos, tlen, start, end, lower, upper, atac)
    offset = smooth // 2
    vals = ra.get_insertions(pos, tlen, start - offset, end + offset, lower, upper, atac)
    return start, ra.smooth(vals, smooth, window="gaussian", mode="valid")
