"""CPU checks of oracle/refpyatac.py (the checker of the pyatac tools): the strand flip of ChunkMat2D.get is pinned by a
size-independent property -- mirroring every fragment about the site centre and flipping the plot gives back the
unflipped plot -- and the coverage helper by its definition as a windowed count of fragment centres."""
import numpy as np

from oracle import refalgo as ra, refpyatac as rp


def _reads(rng, n, lo, hi, max_size=400):
    size = rng.integers(1, max_size, n)
    left = rng.integers(lo, hi, n)
    return (left - 4).astype(np.int32), (size + 8).astype(np.int32)


def test_flip_is_mirror_image():
    rng = np.random.default_rng(3)
    c, flank, lower, upper = 4000, 150, 0, 300
    pos, tlen = _reads(rng, 3000, c - 500, c + 300)
    l, i = ra.shift_fragments(pos, tlen, True)
    r = l + i - 1
    mpos, mtlen = (2 * c - r - 4).astype(np.int32), tlen  # mirrored fragment: left end 2c - r, same size
    plain = rp.vplot_site(pos, tlen, c, c + 1, "+", flank, lower, upper)
    flipped = rp.vplot_site(mpos, mtlen, c, c + 1, "-", flank, lower, upper)
    assert plain.sum() > 500
    np.testing.assert_array_equal(plain, flipped)
    # odd rows of a flipped plot are the plain plot reversed; even rows are reversed and shifted by one
    f2 = rp.vplot_site(pos, tlen, c, c + 1, "-", flank, lower, upper)
    np.testing.assert_array_equal(f2[1::2], plain[1::2, ::-1])
    np.testing.assert_array_equal(f2[0::2, :-1], plain[0::2, ::-1][:, 1:])


def test_center_and_even_width_flip():
    assert rp.center(100, 107, "+") == (103, 104) and rp.center(100, 107, "-") == (103, 104)
    assert rp.center(100, 108, "+") == (104, 105) and rp.center(100, 108, "-") == (103, 104)  # chunk.py:42-47
    try:
        rp.mat_get(np.zeros((4, 10)), 0, 0, 0, 4, 1, 5, flip=True)
        assert False
    except Exception as e:
        assert "odd" in str(e)


def test_cov_chunk_is_windowed_centre_count():
    rng = np.random.default_rng(5)
    start, end, lower, upper = 1000, 1400, 20, 250
    pos, tlen = _reads(rng, 2000, 600, 1500)
    l, i = ra.shift_fragments(pos, tlen, True)
    centre = l + (i - 1) // 2
    ok = (i >= lower) & (i < upper)
    for window in (121, 10, 1):
        half, weff = window // 2, window + (window % 2 == 0)
        exp = np.array([np.sum(ok & (centre >= x - half) & (centre < x - half + weff)) for x in range(start, end)], dtype=float)
        np.testing.assert_array_equal(rp.cov_chunk(pos, tlen, start, end, lower, upper, window, float(window)), exp)


# ----------------------------------------------------------------------------- pins on reference-held data
def test_mat_get_reference_kat():
    """ChunkMat2D.get without flip: the reference's own known answer (tests/test_chunkmat2d.py:12-17)."""
    mat = np.zeros((200, 500))
    mat[100, 5] = 1
    np.testing.assert_array_equal(rp.mat_get(mat, 500, 0, 100, 102, 505, 507), np.array([[1, 0], [0, 0]]))


def test_ins_chunk_reproduces_shipped_ins_track():
    """`pyatac ins` without smoothing writes InsertionTrack.calculateInsertions (pyatac/get_ins.py:20-28 -> tracks.py:159-163);
    the reference shipped exactly that track for the example regions as example_results/example.ins.bedgraph.gz
    (written by `nucleoatac nfr`, NFRCalling.py:73-76).  The checker of the tool must reproduce it value for value."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nfr_golden.npz"))
    n = len(z["chunk_start"])
    total = 0
    for i in range(n):
        s, e = int(z["chunk_start"][i]), int(z["chunk_end"][i])
        a, b = int(z["frag_off"][i]), int(z["frag_off"][i + 1])
        t0, t1 = int(z["track_off"][i]), int(z["track_off"][i + 1])
        start, vals = rp.ins_chunk(z["frag_pos"][a:b], z["frag_tlen"][a:b], s, e, 0, 2000)
        assert start == s and len(vals) == t1 - t0
        np.testing.assert_array_equal(vals, z["gold_ins"][t0:t1])
        total += int(vals.sum())
    assert total > 10000


def test_cov_chunk_reproduces_shipped_occpeak_read_counts(example, golden):
    """`pyatac cov` = CoverageTrack.calculateCoverage times scale / window (pyatac/get_cov.py:22-31).  The reference shipped 160
    values of that very function (window 121, sizes [0, 251)): the `reads` column of example_results/example.occpeaks.bed.gz
    is OccChunk.cov at the peak (Occupancy.py:155-168, 221-227).  With scale = window the tool's helper must give them."""
    starts = np.array([example.chunk(i)[1] for i in range(example.n_chunks)])
    ends = np.array([example.chunk(i)[2] for i in range(example.n_chunks)])
    chroms = [example.chunk(i)[0] for i in range(example.n_chunks)]
    hits = 0
    for c, p, row in zip(golden["occpeaks_chrom"], golden["occpeaks_pos"], golden["occpeaks_vals"]):
        cname = example.chrom_names[int(c)]
        i = [k for k in range(example.n_chunks) if chroms[k] == cname and starts[k] <= p < ends[k]][0]
        pos, tlen = example.reads(i)
        got = rp.cov_chunk(pos, tlen, int(p), int(p) + 1, 0, 251, 121, 121.0)
        assert got.shape == (1,) and got[0] == row[3], (cname, int(p), got, row[3])
        hits += 1
    assert hits == 160
