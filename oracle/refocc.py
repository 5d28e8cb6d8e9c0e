"""Oracle restatement of nucleoatac/Occupancy.py (per-chunk occupancy path).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows the reference function by
function; float64 throughout; Python-2 integer division written as ``//``.
"""
import numpy as np
from scipy import stats

from . import refalgo as ra


class OccCalcParams:
    """nucleoatac/Occupancy.py:89-102 (``nuc_fit`` / ``nfr_fit`` given as arrays over [lower,upper))."""

    def __init__(self, lower, upper, nuc_fit, nfr_fit, ci=0.9):
        self.lower = lower
        self.upper = upper
        nuc_probs = np.asarray(nuc_fit, dtype=np.float64)
        self.nuc_probs = nuc_probs / np.sum(nuc_probs)
        nfr_probs = np.asarray(nfr_fit, dtype=np.float64)
        self.nfr_probs = nfr_probs / np.sum(nfr_probs)
        self.alphas = np.linspace(0, 1, 101)
        self.l = len(self.alphas)
        self.cutoff = stats.chi2.ppf(ci, 1)


def calculate_occupancy(inserts, bias, params):
    """nucleoatac/Occupancy.py:104-120 -> (occ, lower, upper)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        nuc_probs = params.nuc_probs * bias
        nuc_probs = nuc_probs / np.sum(nuc_probs)
        nfr_probs = params.nfr_probs * bias
        nfr_probs = nfr_probs / np.sum(nfr_probs)
        x = [np.log(alpha * nuc_probs + (1 - alpha) * nfr_probs) for alpha in params.alphas]
        logliks = np.array([np.sum(x[j] * inserts) for j in range(params.l)])
    logliks[np.isnan(logliks)] = -float("inf")
    occ = params.alphas[np.argmax(logliks)]
    with np.errstate(invalid="ignore"):
        ratios = 2 * (max(logliks) - logliks)
    ok = np.where(ratios < params.cutoff)[0]
    lower = params.alphas[min(ok)]
    upper = params.alphas[max(ok)]
    return occ, lower, upper


class OccParams:
    """nucleoatac/Occupancy.py:175-193 without the file handles."""

    def __init__(self, nuc_fit, nfr_fit, upper=251, sep=120, min_occ=0.1, flank=60, ci=0.9, step=5):
        self.sep = sep
        self.window = flank * 2 + 1
        self.min_occ = min_occ
        self.flank = flank
        self.upper = upper
        self.occ_calc_params = OccCalcParams(0, upper, nuc_fit, nfr_fit, ci=ci)
        if step % 2 == 0:
            step = step - 1
        self.step = step
        self.halfstep = (self.step - 1) // 2


def occupancy_mle(mat, bias_mat, mat_start, start, end, params):
    """nucleoatac/Occupancy.py:128-146 -> (vals, lower_bound, upper_bound), NaN where no reads."""
    offset = start - mat_start
    if offset < params.flank:
        raise Exception("For calculateOccupancyMLE, mat does not have sufficient flanking regions")
    n = end - start
    vals = np.ones(n) * float("nan")
    lower_bound = np.ones(n) * float("nan")
    upper_bound = np.ones(n) * float("nan")
    for i in range(params.halfstep, n, params.step):
        x1 = start + i - params.flank - mat_start
        x2 = start + i + params.flank + 1 - mat_start
        new_inserts = np.sum(mat[0:params.upper, x1:x2], axis=1)
        new_bias = np.sum(bias_mat[0:params.upper, x1:x2], axis=1)
        if sum(new_inserts) > 0:
            left = i - params.halfstep
            right = min(i + params.halfstep + 1, n)
            vals[left:right], lower_bound[left:right], upper_bound[left:right] = calculate_occupancy(
                new_inserts, new_bias, params.occ_calc_params)
    return vals, lower_bound, upper_bound


def occ_bias_track_span(start, end, params):
    """Span of the InsertionBiasTrack of Occupancy.py:212-213 (before the PWM slop)."""
    return (start - params.window - params.upper // 2, end + params.window + params.upper // 2 + 1)


def process_occ_chunk(pos, tlen, start, end, params, bias_track=None, bias_track_start=None, atac=True):
    """OccChunk.process + getNucDist, nucleoatac/Occupancy.py:195-253 and run_occ.py:23-39.

    ``pos``/``tlen``: kept reads of the chunk's chromosome (any superset of the chunk).
    ``bias_track``: log-bias values over occ_bias_track_span (or None = no --fasta,
    bias matrix stays all ones, Occupancy.py:209-211).
    Returns a dict of every intermediate the reference exposes.
    """
    fl, up = params.flank, params.upper
    mat_start, mat_end = start - fl, end + fl
    mat = ra.make_fragment_mat(pos, tlen, mat_start, mat_end, 0, up, atac)  # :204-207
    if bias_track is not None:  # :208-215
        x1 = (mat_start - up // 2) - bias_track_start
        x2 = (mat_end + up // 2) - bias_track_start
        bias_mat = ra.make_bias_mat(np.asarray(bias_track)[x1:x2], 0, up)
    else:
        bias_mat = np.ones(mat.shape)
    vals, lb, ub = occupancy_mle(mat, bias_mat, mat_start, start, end, params)  # :216-219
    sd = params.flank / 3.0
    sm_vals = ra.smooth(vals, params.window, window="gaussian", sd=sd, mode="same", norm=True)  # :147-153,220
    sm_lower = ra.smooth(lb, params.window, window="gaussian", sd=sd, mode="same", norm=True)
    sm_upper = ra.smooth(ub, params.window, window="gaussian", sd=sd, mode="same", norm=True)
    cov = ra.calculate_coverage(mat, mat_start, 0, start, 0, up, params.window)  # :221-224
    peaks_idx = ra.call_peaks(sm_vals, sep=params.sep, min_signal=params.min_occ)  # :227 (mutates sm_vals)
    peaks = []
    for p in peaks_idx:
        p = int(p)
        occ, occ_lower, occ_upper, reads = sm_vals[p], sm_lower[p], sm_upper[p], cov[p]
        if occ_lower > params.min_occ and reads > 0:  # :230
            peaks.append((p + start, occ, occ_lower, occ_upper, reads))
    nuc_dist = np.zeros(up)  # :232-240
    for (ppos, _o, _l, _u, _r) in peaks:
        sub = mat[:, (ppos - fl - mat_start):(ppos + 1 + fl - mat_start)]
        sub_sum = np.sum(sub, axis=1)
        nuc_dist += sub_sum / float(sum(sub_sum))
    peaks.sort(key=lambda t: t[0])
    return dict(mat=mat, bias_mat=bias_mat, vals=vals, lower_bound=lb, upper_bound=ub,
                smoothed_vals=sm_vals, smoothed_lower=sm_lower, smoothed_upper=sm_upper,
                cov=cov, peaks=peaks, nuc_dist=nuc_dist)


def occ_peak_bed(chrom, peak):
    """OccPeak.asBed, nucleoatac/Occupancy.py:166-168."""
    pos, occ, lo, up, reads = peak
    return "\t".join([chrom, str(pos), str(pos + 1), ra.fmt12(occ), ra.fmt12(lo), ra.fmt12(up), ra.fmt12(reads)])
