"""smooth / call_peaks / reduce_peaks on the device + chromosome-size readers (pyatac/utils.py:23-134)."""
import numpy as np

from . import hostio
from .engine import default_engine, gaussian_window


def fmt12(x):
    """Python-2 ``str(float)`` = '%.12g' (+ '.0' on integral values): the reference's text outputs
    (tracks.py:63, Occupancy.py:167, NucleosomeCalling.py:196, VMat.py:189, fragmentsizes.py:53) are printed
    with it, so the drop-in writers use it too."""
    if isinstance(x, (int, np.integer)):
        return str(int(x))
    x = float(x)
    if x != x:
        return "nan"
    if x in (float("inf"), float("-inf")):
        return "inf" if x > 0 else "-inf"
    s = "%.12g" % x
    if "." not in s and "e" not in s:
        s += ".0"
    return s


def smooth(sig, window_len, window="flat", sd=None, mode="valid", norm=True):
    """NaN-aware smoothing (pyatac/utils.py:23-52), evaluated on the device (nb200_smooth)."""
    if window not in ("flat", "gaussian"):
        raise Exception("Incorrect window input for smooth. Options are flat, gaussian")
    if window_len % 2 != 1:
        window_len += 1  # the reference only warns; the window must be odd
    if window == "gaussian":
        if sd is None:
            sd = (window_len - 1) / 6.0
        w = gaussian_window(window_len, sd)
    else:
        w = np.ones(window_len)
    return default_engine().smooth(np.asarray(sig, dtype=np.float64), w, mode=mode, norm=norm)


def reduce_peaks(peaks, sig, sep):
    """Greedy non-maximum suppression by descending `sig` (pyatac/utils.py:56-78)."""
    peaks = np.asarray(peaks)
    if peaks.size == 0:
        return peaks
    return default_engine().reduce_peaks(peaks, np.asarray(sig, dtype=np.float64), sep)


def call_peaks(sigvals, min_signal=0, sep=120, boundary=None, order=1):
    """Local maxima of the jittered signal + NMS (pyatac/utils.py:82-102).  NaNs in `sigvals` are replaced by
    the minimum IN PLACE when it is a float64 array, like the reference."""
    inplace = isinstance(sigvals, np.ndarray) and sigvals.dtype == np.float64 and sigvals.flags["C_CONTIGUOUS"]
    arr = sigvals if inplace else np.ascontiguousarray(sigvals, dtype=np.float64)
    if np.isnan(arr).all():
        return np.array([])
    return default_engine().call_peaks(arr, min_signal=min_signal, sep=sep, boundary=boundary, order=order)


def read_chrom_sizes_from_fasta(fastafile):
    fa = hostio.FastaFile(fastafile)
    out = dict(zip(fa.references, fa.lengths))
    fa.close()
    return out


def read_chrom_sizes_from_bam(bamfile):
    bam = hostio.BamFile(bamfile)
    out = dict(zip(bam.references, bam.lengths))
    bam.close()
    return out


def read_chrom_sizes(sizesFile):
    out = {}
    with open(sizesFile) as fh:
        for line in fh:
            f = line.split()
            if len(f) >= 2:
                out[f[0]] = int(f[1])
    return out
