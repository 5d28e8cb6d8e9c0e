"""Oracle restatement of the array primitives under the occ/nuc scoring path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  float64 numpy/scipy, Python-3
restatement of the reference's L0/L1 layer; every function cites the reference
file:line it follows.  The reference is Python 2: every ``/`` on ints there is a
floor division and is written ``//`` here.

Geometry convention used throughout: a chunk matrix spans
genomic columns [mat_start, mat_end) and insert-size rows [lower, upper); a track
spans [start, end).  Fragments arrive as raw BAM fields ``pos`` (0-based leftmost
coordinate of the forward mate) and ``tlen`` for reads that are proper-pair and
forward (filter applied by the reader, fragments.pyx:25).
"""
import numpy as np
from scipy import signal

# ----------------------------------------------------------------------------
# fragments.pyx
# ----------------------------------------------------------------------------


def shift_fragments(pos, tlen, atac=True):
    """ATAC insertion-to-insertion correction, pyatac/fragments.pyx:26-35."""
    pos = np.asarray(pos, dtype=np.int64)
    tlen = np.asarray(tlen, dtype=np.int64)
    if atac:
        return pos + 4, np.abs(tlen) - 8
    return pos, np.abs(tlen)


def make_fragment_mat(pos, tlen, start, end, lower, upper, atac=True):
    """pyatac/fragments.pyx:17-40 -> f64[(upper-lower), (end-start)] of counts.

    row = ilen-lower ; col = (ilen-1)//2 + l_pos - start (floor division);
    a cell is incremented iff 0<=col<ncol and 0<=row<nrow.
    """
    nrow, ncol = upper - lower, end - start
    mat = np.zeros((nrow, ncol), dtype=np.float64)
    l_pos, ilen = shift_fragments(pos, tlen, atac)
    row = ilen - lower
    col = (ilen - 1) // 2 + l_pos - start
    ok = (col >= 0) & (col < ncol) & (row < nrow) & (row >= 0)
    np.add.at(mat, (row[ok], col[ok]), 1.0)
    return mat


def get_insertions(pos, tlen, start, end, lower, upper, atac=True):
    """pyatac/fragments.pyx:43-67 (used by the reference's test_tracks KAT)."""
    out = np.zeros(end - start, dtype=np.float64)
    l_pos, ilen = shift_fragments(pos, tlen, atac)
    r_pos = l_pos + ilen - 1
    size_ok = (ilen >= lower) & (ilen < upper)
    lo = size_ok & (l_pos >= start) & (l_pos < end)
    ro = size_ok & (r_pos >= start) & (r_pos < end)
    np.add.at(out, l_pos[lo] - start, 1.0)
    np.add.at(out, r_pos[ro] - start, 1.0)
    return out


def fragment_size_counts(pos, tlen, chunks, lower, upper, atac=True):
    """pyatac/fragments.pyx:122-145 for the fragments of ONE chromosome.

    ``chunks`` is an iterable of (start, end).  A fragment is counted once per
    chunk whose [start,end) contains its centre l_pos+(ilen-1)//2.
    """
    sizes = np.zeros(upper - lower, dtype=np.float64)
    l_pos, ilen = shift_fragments(pos, tlen, atac)
    center = l_pos + (ilen - 1) // 2
    size_ok = (ilen < upper) & (ilen >= lower)
    for (cs, ce) in chunks:
        sel = size_ok & (center >= cs) & (center < ce)
        np.add.at(sizes, ilen[sel] - lower, 1.0)
    return sizes


def normalize_sizes(counts):
    """pyatac/fragmentsizes.py:27."""
    tot = np.sum(counts)
    return counts / (tot + (tot == 0))


# ----------------------------------------------------------------------------
# seq.py / bias.py
# ----------------------------------------------------------------------------


def seq_to_mat(sequence, nucleotides):
    """pyatac/seq.py:37-45 for single-letter nucleotides: one-hot len(nuc) x len(seq)."""
    arr = np.frombuffer(sequence.encode() if isinstance(sequence, str) else bytes(sequence), dtype=np.uint8)
    mat = np.zeros((len(nucleotides), arr.size), dtype=np.float64)
    for i, nuc in enumerate(nucleotides):
        mat[i] = (arr == ord(nuc))
    return mat


def log_bias_track(sequence, pwm_mat, nucleotides=("A", "C", "G", "T")):
    """pyatac/bias.py:87-92: ``correlate(onehot, log(pwm), 'valid')[0]``.

    ``sequence`` must already span [track_start-pwm.up, track_end+pwm.down) and be
    upper-cased (seq.py:22).  Output length = len(sequence) - pwm_width + 1.
    """
    seqmat = seq_to_mat(sequence, nucleotides)
    return signal.correlate(seqmat, np.log(np.asarray(pwm_mat, dtype=np.float64)), mode="valid")[0]


def read_pwm(path):
    """pyatac/bias.py:47-76: returns (mat f64[4,21], up, down, nucleotides)."""
    state = ""
    mat = []
    up = down = nucleotides = None
    with open(path) as fh:
        for line in fh:
            if "#up" in line:
                state = "up"
            elif "#down" in line:
                state = "down"
            elif "#mat" in line:
                state = "mat"
            elif "#nucleotides" in line:
                state = "nucleotides"
            elif state == "up":
                up = int(line.strip("\n"))
            elif state == "down":
                down = int(line.strip("\n"))
            elif state == "nucleotides":
                nucleotides = line.strip("\n").split()
            elif state == "mat":
                mat.append([float(x) for x in line.strip("\n").split("\t")])
    if up is None or down is None or nucleotides is None:
        raise Exception("PWM decriptor file appeas to be missing some needed components")
    return np.array(mat), up, down, nucleotides


# ----------------------------------------------------------------------------
# chunkmat2d.py
# ----------------------------------------------------------------------------


def _two_tap_pattern(lower, upper):
    """The pattern matrix of pyatac/chunkmat2d.py:76-80 and :147-151."""
    pattern = np.zeros((upper - lower, upper + (upper - 1) % 2))
    mid = upper // 2
    for i in range(lower, upper):
        pattern[i - lower, mid + (i - 1) // 2] = 1
        pattern[i - lower, mid - (i // 2)] = 1
    return pattern


def make_bias_mat_literal(bias_vals, lower, upper):
    """pyatac/chunkmat2d.py:140-153, literally (251 ``np.convolve`` calls).

    ``bias_vals`` = log-bias over genomic [mat_start-upper//2, mat_end+upper//2).
    """
    pattern = _two_tap_pattern(lower, upper)
    ncol = len(bias_vals) - pattern.shape[1] + 1
    mat = np.ones((upper - lower, ncol))
    for i in range(upper - lower):
        mat[i] = np.exp(np.convolve(bias_vals, pattern[i, :], mode="valid"))
    return mat


def make_bias_mat(bias_vals, lower, upper):
    """Same result as :func:`make_bias_mat_literal` by the derived two-tap gather
    (SURVEY App. A): cell(i, c) = exp(b[c-(i-1)//2] + b[c+i//2]); when both taps
    coincide (i == 1) the single tap is used once."""
    off = upper // 2
    plen = upper + (upper - 1) % 2
    ncol = len(bias_vals) - plen + 1
    # np.convolve 'valid' index algebra: out[n] = sum_m a[n + plen-1-m] v[m]
    base = plen - 1 - upper // 2  # == off for every upper
    assert base == off
    cols = np.arange(ncol)
    mat = np.empty((upper - lower, ncol))
    for i in range(lower, upper):
        a = base + cols - (i - 1) // 2
        b = base + cols + i // 2
        if (i - 1) // 2 == -(i // 2):
            mat[i - lower] = np.exp(bias_vals[a])
        else:
            mat[i - lower] = np.exp(bias_vals[b] + bias_vals[a])
    return mat


def norm_by_insert_dist(mat, inserts):
    """pyatac/chunkmat2d.py:154-156: row i scaled by inserts[i]."""
    return mat * np.reshape(np.tile(inserts, mat.shape[1]), mat.shape, order="F")


def get_ins(mat, mat_start, mat_end, lower, upper, literal=False):
    """pyatac/chunkmat2d.py:74-84 -> (ins_vals, ins_start, ins_end).

    ins[p] = number of fragments (size in [lower,upper)) with left or right end at p.
    """
    pattern = _two_tap_pattern(lower, upper)
    plen = pattern.shape[1]
    if literal:
        ins = signal.correlate2d(mat, pattern, mode="valid")[0]
    else:
        n = mat.shape[1] - plen + 1
        ins = np.zeros(n)
        mid = upper // 2
        for i in range(lower, upper):
            t1, t2 = mid + (i - 1) // 2, mid - (i // 2)
            ins += mat[i - lower, t1:t1 + n]
            if t2 != t1:
                ins += mat[i - lower, t2:t2 + n]
    return ins, mat_start + plen // 2, mat_end - plen // 2


# ----------------------------------------------------------------------------
# utils.py
# ----------------------------------------------------------------------------


def smooth(sig, window_len, window="flat", sd=None, mode="valid", norm=True):
    """pyatac/utils.py:23-52."""
    if window not in ["flat", "gaussian"]:
        raise Exception("Incorrect window input for smooth. Options are flat, gaussian")
    if window_len % 2 != 1:
        window_len += 1
    if window == "gaussian" and sd is None:
        sd = (window_len - 1) / 6.0
    if window == "gaussian":
        w = signal.windows.gaussian(window_len, sd)
    else:
        w = np.ones(window_len)
    sig = np.asarray(sig, dtype=np.float64)
    sig_nonan = sig.copy()
    sig_nonan[np.isnan(sig)] = 0
    smoothed = np.convolve(w, sig_nonan, mode=mode)
    if norm:
        norm_sig = np.ones(len(sig))
        norm_sig[np.isnan(sig)] = 0
        smoothed_norm = np.convolve(w, norm_sig, mode=mode)
        smoothed_norm[smoothed_norm == 0] = np.nan
        smoothed = smoothed / smoothed_norm
    return smoothed


def reduce_peaks(peaks, sig, sep):
    """pyatac/utils.py:56-78: greedy NMS walking np.argsort(sig) from the top."""
    peaks = np.asarray(peaks)
    exclude = np.zeros(peaks.size)
    keep = np.zeros(peaks.size)
    st = np.argsort(sig)
    j = peaks.size - 1
    while j >= 0:
        ind = st[j]
        j += -1
        if exclude[ind] == 0:
            keep[ind] = 1
            exclude[ind] = 1
            k = ind - 1
            while k >= 0 and (peaks[ind] - peaks[k]) < sep:
                exclude[k] = 1
                k += -1
            k = ind + 1
            while k < peaks.size and (peaks[k] - peaks[ind]) < sep:
                exclude[k] = 1
                k += 1
    return peaks[keep == 1]


_JITTER_SEED = 25


def peak_jitter(n):
    """The multiplicative jitter of pyatac/utils.py:94-97: RandomState(25).uniform(0,1e-12,n).
    The first n draws of the legacy MT19937 stream, so a prefix of any longer draw."""
    return np.random.RandomState(seed=_JITTER_SEED).uniform(0, 10 ** -12, n)


def call_peaks(sigvals, min_signal=0, sep=120, boundary=None, order=1):
    """pyatac/utils.py:82-102.  NOTE: mutates ``sigvals`` (NaN -> min) like the reference."""
    nan = np.isnan(sigvals)
    if nan.sum() > 0:
        if nan.sum() == len(sigvals):
            return np.array([])
        sigvals[nan] = np.min(sigvals[~nan])
    if boundary is None:
        boundary = sep // 2
    l = len(sigvals)
    peaks = signal.argrelmax(sigvals * (1 + peak_jitter(l)), order=order)[0]
    peaks = peaks[sigvals[peaks] >= min_signal]
    peaks = peaks[peaks >= boundary]
    peaks = peaks[peaks < (l - boundary)]
    return reduce_peaks(peaks, sigvals[peaks], sep)


# ----------------------------------------------------------------------------
# tracks.py
# ----------------------------------------------------------------------------


def calculate_coverage(mat, mat_start, mat_lower, start, lower, upper, window_len):
    """pyatac/tracks.py:209-222: flat-window coverage of fragment centres.

    Returns f64[(end-start)] given that mat spans far enough on both sides.
    """
    offset = start - mat_start - (window_len // 2)
    if offset < 0:
        raise Exception("Insufficient flanking region on mat to calculate coverage with desired window")
    lo, up = lower - mat_lower, upper - mat_lower
    if offset != 0:
        collapsed = np.sum(mat[lo:up, offset:-offset], axis=0)
    else:
        collapsed = np.sum(mat[lo:up, ], axis=0)
    return smooth(collapsed, window_len, window="flat", mode="valid", norm=False)


def fmt12(x):
    """Python-2 ``str(float)`` (= ``'%.12g'`` with a forced '.0' on integral values);
    every text output of the reference goes through it (tracks.py:63, Occupancy.py:167,
    NucleosomeCalling.py:196, VMat.py:189, fragmentsizes.py:53)."""
    if isinstance(x, (int, np.integer)):
        return str(int(x))
    x = float(x)
    if x != x:
        return "nan"
    if x in (float("inf"), float("-inf")):
        return "inf" if x > 0 else "-inf"
    s = "%.12g" % x
    if "." not in s and "e" not in s and "n" not in s:
        s += ".0"
    return s


def write_track(chrom, start, end, vals, write_zero=True):
    """pyatac/tracks.py:37-74: run-length bedgraph text of a track (NaN runs skipped)."""
    if len(vals) != end - start:
        raise Exception("Error! Inconsistency between length of values and start/end values")
    out = []
    prev_value = None
    start_range = 0
    for i in range(len(vals)):
        v = vals[i]
        if prev_value is not None and v == prev_value:
            pass
        elif np.isnan(v):
            # literal: the run that precedes a NaN is NOT flushed (tracks.py:59-60)
            prev_value = v
        elif prev_value is not None and not np.isnan(prev_value):
            if write_zero or prev_value != 0:
                out.append("\t".join([chrom, str(start_range), str(start + i), fmt12(prev_value)]))
            start_range = start + i
            prev_value = v
        else:
            start_range = start + i
            prev_value = v
    if prev_value is not None:
        if prev_value == 0:
            if write_zero:
                out.append("\t".join([chrom, str(start_range), str(end), fmt12(prev_value)]))
        elif not np.isnan(prev_value):
            out.append("\t".join([chrom, str(start_range), str(end), fmt12(prev_value)]))
    return "".join(s + "\n" for s in out)


# ----------------------------------------------------------------------------
# chunk.py
# ----------------------------------------------------------------------------


def read_bed_chunks(path, chrom_sizes=None, min_offset=None, min_length=1):
    """pyatac/chunk.py:132-175 -> list of [chrom, start, end]."""
    out = []
    with open(path) as fh:
        for line in fh:
            f = line.rstrip("\n").split("\t")
            if len(f) < 3:
                continue
            chrom, start, end = f[0], int(f[1]), int(f[2])
            if chrom_sizes is not None and chrom not in chrom_sizes:
                continue
            if min_offset:
                if start < min_offset:
                    start = min_offset
                if end > chrom_sizes[chrom] - min_offset:
                    end = chrom_sizes[chrom] - min_offset
            if end - start >= min_length:
                out.append([chrom, start, end])
    return out


def slop_chunks(chunks, chrom_sizes, up, down):
    """pyatac/chunk.py:26-40,101-108 for strand '+' / '*'."""
    return [[c, max(0, s - up), min(chrom_sizes[c], e + down)] for (c, s, e) in chunks]


def merge_chunks(chunks, sep=-1):
    """pyatac/chunk.py:109-125 on an already sorted list."""
    out = []
    previous = list(chunks[0])
    for i in range(1, len(chunks)):
        c, s, e = chunks[i]
        if c == previous[0] and s <= previous[2] + sep:
            previous[2] = max(e, previous[2])
        else:
            out.append(previous)
            previous = [c, s, e]
    out.append(previous)
    return out


# ----------------------------------------------------------------------------
# VMat.py / fragmentsizes.py text IO
# ----------------------------------------------------------------------------


def read_vmat(path):
    """pyatac/VMat.py:191-218 -> (mat, lower, upper)."""
    state = ""
    mat = []
    lower = upper = None
    with open(path) as fh:
        for line in fh:
            if "#lower" in line:
                state = "lower"
            elif "#upper" in line:
                state = "upper"
            elif "#mat" in line:
                state = "mat"
            elif "#" in line:
                state = "other"
            elif state == "lower":
                lower = int(line.strip("\n"))
            elif state == "upper":
                upper = int(line.strip("\n"))
            elif state == "mat":
                mat.append([float(x) for x in line.strip("\n").split("\t")])
    mat = np.array(mat)
    if mat.shape[0] != upper - lower:  # VMat.py:33-34
        raise Exception("mat shape is not consistent with insert limits")
    return mat, lower, upper


def read_sizes(path):
    """pyatac/fragmentsizes.py:56-81 -> (vals, lower, upper)."""
    state = ""
    lower = upper = vals = None
    with open(path) as fh:
        for line in fh:
            if "#lower" in line:
                state = "lower"
            elif "#upper" in line:
                state = "upper"
            elif "#sizes" in line:
                state = "sizes"
            elif "#" in line:
                state = "other"
            elif state == "lower":
                lower = int(line.strip("\n"))
            elif state == "upper":
                upper = int(line.strip("\n"))
            elif state == "sizes":
                vals = np.array([float(x) for x in line.rstrip("\n").split("\t")])
    return vals, lower, upper
