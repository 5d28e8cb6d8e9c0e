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
