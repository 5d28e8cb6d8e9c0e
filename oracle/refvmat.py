"""Oracle restatement of the V-plot processing of pyatac/VMat.py + nucleoatac/run_vprocess.py.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Once-per-run host work in the reference; used here
to derive the VMat templates the synthetic workload needs and to check the product's VMat class.
"""
import numpy as np
from scipy import ndimage


def trim(mat, lower0, w0, lower, upper, w):
    """pyatac/VMat.py:38-54 -> new mat (rows [lower,upper), columns centre +-w)."""
    up, dn = upper - lower0, lower - lower0
    left, right = w0 - w, w0 + w + 1
    if up > mat.shape[0] or dn < 0 or left < 0 or right > mat.shape[1]:
        raise Exception("Mat is smaller than desired trim")
    return np.array(mat[dn:up, left:right], dtype=np.float64, copy=True)


def symmetrize(mat, lower, upper):
    """pyatac/VMat.py:55-64 (odd sizes mirror about the centre column; even about centre-0.5)."""
    mat = np.array(mat, dtype=np.float64, copy=True)
    w = mat.shape[1] // 2
    for j in range(lower, upper):
        i = j - lower
        if j % 2 == 1:
            lefthalf = (mat[i, :(w + 1)] + mat[i, w:][::-1]) * 0.5
            mat[i, :] = np.hstack((lefthalf, lefthalf[:-1][::-1]))
        else:
            righthalf = (mat[i, w:-1] + mat[i, :w][::-1]) * 0.5
            mat[i, :] = np.hstack((righthalf[::-1], righthalf, mat[i, -1]))
    return mat


def norm_y(mat, lower, dist_vals, dist_lower):
    """pyatac/VMat.py:104-107: scale each row to the supplied insert-size distribution.
    NOTE FragmentSizes.get(size=0) falls through to the slice branch (`if size:`); lower>0 here."""
    mat = np.array(mat, dtype=np.float64, copy=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(mat.shape[0]):
            mat[i] = mat[i] * (dist_vals[i + lower - dist_lower] / np.sum(mat[i]))
    return mat


def smooth(mat, sd):
    """pyatac/VMat.py:90-93."""
    return ndimage.gaussian_filter(mat, sd, mode="constant")


def norm(mat):
    """pyatac/VMat.py:98-103."""
    tmp1 = mat / np.sum(mat)
    tmp2 = np.ones(mat.shape) * (1.0 / mat.size)
    mat = mat / (np.sum(mat * tmp1) - np.sum(mat * tmp2))
    return (mat / mat.shape[1]) * 10.0


def vprocess(raw, raw_lower, lower=105, upper=251, flank=60, sizes=None, sizes_lower=0, smooth_sd=0.75):
    """nucleoatac/run_vprocess.py:15-32 -> processed mat."""
    m = trim(raw, raw_lower, raw.shape[1] // 2, lower, upper, flank)
    m = symmetrize(m, lower, upper)
    if sizes is not None:
        m = norm_y(m, lower, sizes, sizes_lower)
    if smooth_sd > 0:
        m = smooth(m, smooth_sd)
    return norm(m)
