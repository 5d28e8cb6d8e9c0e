"""Oracle restatement of nucleoatac/NucleosomeCalling.py (per-chunk nucleosome path).

TEST INFRASTRUCTURE (see oracle/__init__.py).  float64; Python-2 integer division
written as ``//``; follows the reference function by function.
"""
from bisect import bisect_left

import numpy as np
from scipy import optimize, signal

from . import mcov
from . import refalgo as ra


class NucParams:
    """nucleoatac/NucleosomeCalling.py:204-226 without the file handles.

    ``vmat`` = (mat f64[R,W], lower, upper); ``fragmentsizes`` = f64[upper] frequencies over
    sizes [0, upper) (FragmentSizes with lower 0, as run_nuc.py:155-159 builds it).
    """

    def __init__(self, vmat, fragmentsizes, atac=True, sd=25, nonredundant_sep=120,
                 redundant_sep=25, min_z=3, min_lr=0, min_reads=1):
        self.vmat, self.lower, self.upper = vmat
        self.vmat = np.asarray(self.vmat, dtype=np.float64)
        if self.vmat.shape[0] != self.upper - self.lower:  # VMat.py:33-34
            raise Exception("mat shape is not consistent with insert limits")
        self.w = self.vmat.shape[1] // 2  # VMat.py:37
        self.window = self.vmat.shape[1]  # :214
        self.fragmentsizes = np.asarray(fragmentsizes, dtype=np.float64)
        self.atac = atac
        self.min_reads = min_reads
        self.min_z = min_z
        self.min_lr = min_lr
        self.smooth_sd = sd
        self.redundant_sep = redundant_sep
        self.nonredundant_sep = nonredundant_sep


def nuc_geometry(start, end, params):
    """Spans of NucleosomeCalling.py:239-247: (mat span, bias_mat span, bias-track span)."""
    pad = max(params.window, params.upper // 2 + 1)
    mat_span = (start - pad, end + pad)
    bias_span = (start - params.window, end + params.window)
    track_span = (start - params.window - params.upper // 2, end + params.window + params.upper // 2 + 1)
    return mat_span, bias_span, track_span


def calculate_signal(mat, mat_start, mat_end, mat_lower, start, vmat, v_lower, v_upper, method="auto"):
    """SignalTrack.calculateSignal, NucleosomeCalling.py:29-36 (also the xcor of :60-63)."""
    w = vmat.shape[1] // 2
    offset = start - mat_start - w
    if offset < 0:
        raise Exception("Insufficient flanking region on mat to calculate signal")
    ncol = mat_end - mat_start
    sub = mat[(v_lower - mat_lower):(v_upper - mat_lower), offset:(ncol - offset)]
    return signal.correlate(sub, vmat, mode="valid", method=method)[0]


def smooth_signal(norm_vals, sd):
    """NucChunk.smoothSignal, NucleosomeCalling.py:274-283."""
    window_len = 6 * sd + 1
    tmp = np.array(norm_vals, dtype=np.float64, copy=True)
    tmp[tmp < 0] = 0
    return ra.smooth(tmp, window_len, window="gaussian", sd=sd, mode="same", norm=True)


def _norm_pdf(x, v, w, mean):
    """NucleosomeCalling.py:92-97."""
    n = (1.0 / (np.sqrt(2 * np.pi * v)) * np.exp(-(x - mean) ** 2 / (2 * v)))
    return n * (w / max(n))


def get_fuzz(index, sorted_keys, smoothed_vals, params):
    """Nucleosome.getFuzz, NucleosomeCalling.py:137-194 -> (fuzz, weight, fit_pos).  Host scipy
    in the product as well."""
    def add_norms(x, p):
        fit = np.zeros(len(x))
        for j in range(len(p) // 3):
            fit += _norm_pdf(x, p[j * 3], p[3 * j + 1], p[3 * j + 2])
        return fit

    def err_func(pars, y):
        x = np.linspace(0, len(y) - 1, len(y))
        return sum((add_norms(x, pars) - y) ** 2)

    sep = params.nonredundant_sep
    allnucs = list(sorted_keys)
    x = bisect_left(allnucs, index)
    if x == 0:
        left = index - sep // 3
        means = (sep // 3,)
    elif index - allnucs[x - 1] < sep:
        left = allnucs[x - 1]
        means = (index - allnucs[x - 1], 0)
    else:
        left = index - sep // 3
        means = (sep // 3,)
    if x == len(allnucs) - 1:
        right = index + sep // 3 + 1
    elif allnucs[x + 1] - index < sep:
        right = allnucs[x + 1]
        means += (allnucs[x + 1] - left,)
    else:
        right = index + sep // 3 + 1
    sig = smoothed_vals[left:right]
    sig[sig < 0] = 0
    bounds, guesses = (), ()
    for m in means:
        bounds += ((2 ** 2, 50 ** 2), (0.001, max(sig) * 1.1), (m - 10, m + 10))
        guesses += (params.smooth_sd ** 2, max(sig) * 0.9, m)
    res = optimize.minimize(err_func, guesses, args=(sig,), bounds=bounds, method="L-BFGS-B")
    return np.sqrt(res["x"][0]), res["x"][1], res["x"][2] + left


def process_nuc_chunk(pos, tlen, start, end, params, bias_track=None, bias_track_start=None,
                      occ_tracks=None, fit=True, closed_cov=True, xcor_method="auto", want_ins=False):
    """NucChunk.process, nucleoatac/NucleosomeCalling.py:328-340, + run_nuc.py:22-39.

    ``bias_track``: log-bias over nuc_geometry()[2] (None = no --fasta, :248).
    ``occ_tracks``: optional (occ, lower, upper) arrays over [start,end) (:284-293).
    ``closed_cov``: use the O(n) identity instead of the reference's O(n^2) loop for the variance.
    """
    V, lv, uv, w = params.vmat, params.lower, params.upper, params.w
    (m0, m1), (b0, b1), _ = nuc_geometry(start, end, params)
    n = end - start
    mat = ra.make_fragment_mat(pos, tlen, m0, m1, 0, uv, params.atac)  # :239-242
    if bias_track is not None:  # :243-250
        x1 = (b0 - uv // 2) - bias_track_start
        x2 = (b1 + uv // 2) - bias_track_start
        bias_prenorm = ra.make_bias_mat(np.asarray(bias_track)[x1:x2], 0, uv)
    else:
        bias_prenorm = np.ones((uv, b1 - b0))
    bias_mat = ra.norm_by_insert_dist(bias_prenorm, params.fragmentsizes[0:uv])  # :251-254
    # getNucSignal :255-268
    nuc_cov = ra.calculate_coverage(mat, m0, 0, start, lv, uv, params.window)
    bias_cov = ra.calculate_coverage(bias_mat, b0, 0, start, lv, uv, w * 2 + 1)  # :56-58
    bx = calculate_signal(bias_mat, b0, b1, 0, start, V, lv, uv, xcor_method)  # :60-63
    with np.errstate(divide="ignore", invalid="ignore"):
        background = bx * nuc_cov / bias_cov  # :64
    nuc_signal = calculate_signal(mat, m0, m1, 0, start, V, lv, uv, xcor_method)  # :264-266
    norm_signal = nuc_signal - background  # :42-43
    nfr_cov = ra.calculate_coverage(mat, m0, 0, start, 0, lv, params.window)  # :269-273
    smoothed = smooth_signal(norm_signal, params.smooth_sd)  # :274-283
    # findAllNucs :294-315
    combined = norm_signal + smoothed
    cands = ra.call_peaks(combined, min_signal=0, sep=params.redundant_sep,
                          boundary=params.nonredundant_sep // 2, order=params.redundant_sep // 2)
    flatv = np.ravel(V)
    nucs = {}
    cand_stats = []
    for i in cands:
        i = int(i)
        rec = dict(pos=i + start, nfr_cov=nfr_cov[i], nuc_cov=nuc_cov[i], nuc_signal=nuc_signal[i],
                   norm_signal=norm_signal[i], smoothed=smoothed[i], lr=np.nan, z=np.nan)
        cand_stats.append(rec)
        if rec["nuc_cov"] > params.min_reads:
            p = i + start
            mwin = mat[lv:uv, (p - w - m0):(p + w + 1 - m0)]  # getLR :110-122
            null_mat = bias_mat[lv:uv, (p - w - b0):(p + w + 1 - b0)]
            bwin = bias_prenorm[lv:uv, (p - w - b0):(p + w + 1 - b0)]
            with np.errstate(divide="ignore", invalid="ignore"):
                nuc_model = V * bwin
                nuc_model = nuc_model / np.sum(nuc_model)
                null_model = null_mat / np.sum(null_mat)
                nuc_lik = np.sum(np.log(nuc_model) * mwin)
                null_lik = np.sum(np.log(null_model) * mwin)
            rec["lr"] = nuc_lik - null_lik
            if rec["lr"] > params.min_lr:
                probs = (null_mat / np.sum(null_mat)).flatten()  # SignalDistribution :70-76
                var = mcov.calculate_cov(probs, flatv, rec["nuc_cov"], closed=closed_cov)  # :83-86
                with np.errstate(divide="ignore", invalid="ignore"):
                    rec["z"] = rec["norm_signal"] / np.sqrt(var)  # :123-127
                if rec["z"] >= params.min_z:
                    if occ_tracks is not None:  # getOcc :128-136
                        rec["occ"], rec["occ_lower"], rec["occ_upper"] = (
                            occ_tracks[0][i], occ_tracks[1][i], occ_tracks[2][i])
                    else:
                        rec["occ"] = rec["occ_lower"] = rec["occ_upper"] = np.nan
                    nucs[i] = rec
    sorted_keys = np.array(sorted(nucs.keys()), dtype=np.int64)
    nonredundant = ra.reduce_peaks(sorted_keys, [nucs[k]["z"] for k in sorted_keys], params.nonredundant_sep)
    redundant = np.setdiff1d(sorted_keys, nonredundant)
    if fit:  # :316-324
        for k in sorted_keys:
            nucs[int(k)]["fuzz"], nucs[int(k)]["weight"], nucs[int(k)]["fit_pos"] = get_fuzz(
                int(k), sorted_keys, smoothed, params)
    out = dict(mat=mat, bias_mat=bias_mat, bias_mat_prenorm=bias_prenorm, nuc_cov=nuc_cov, bias_cov=bias_cov,
               bias=background, nuc_signal=nuc_signal, norm_signal=norm_signal, nfr_cov=nfr_cov,
               smoothed=smoothed, cands=np.asarray(cands, dtype=np.int64), cand_stats=cand_stats,
               nuc_collection=nucs, sorted_nuc_keys=sorted_keys,
               nonredundant=np.asarray(nonredundant, dtype=np.int64), redundant=redundant)
    if want_ins:  # :325-327
        out["ins"] = ra.get_ins(mat, m0, m1, 0, uv)
    return out


def nuc_bed(chrom, rec):
    """Nucleosome.asBed, nucleoatac/NucleosomeCalling.py:195-199."""
    f = ra.fmt12
    return "\t".join([chrom, str(rec["pos"]), str(rec["pos"] + 1), f(rec["z"]), f(rec["occ"]), f(rec["occ_lower"]),
                      f(rec["occ_upper"]), f(rec["lr"]), f(rec["norm_signal"]), f(rec["nuc_signal"]),
                      f(rec["nuc_cov"]), f(rec["nfr_cov"]), f(rec.get("fuzz", np.nan))])
