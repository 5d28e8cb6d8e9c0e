"""Track / InsertionTrack / CoverageTrack (pyatac/tracks.py:16-242): 1-D values on [start, end)."""
import numpy as np

from .bedgraph import BedGraphFile
from .chunk import Chunk
from .engine import default_engine
from .fragments import getInsertions
from .utils import smooth


class Track(Chunk):
    def __init__(self, chrom, start, end, name="track", vals=None, log=False):
        Chunk.__init__(self, chrom, start, end, name=name)
        self.log = log
        if vals is not None and len(vals) != self.length():
            raise Exception("Input vals must be of length as set by start and end!")
        self.vals = vals

    def assign_track(self, vals, start=None, end=None):
        if start:
            self.start = start
        if end:
            self.end = end
        if len(vals) != self.end - self.start:
            raise Exception("The values being assigned to track do not span the start to end of the track")
        self.vals = vals

    def format_track(self, start=None, end=None, vals=None, write_zero=True):
        """The rows write_track writes, as bytes: run-length bedgraph text, NaN runs skipped, numbers as the reference prints
        them (tracks.py:37-74).  Formatted by the library's host-side writer (nb200_format_track: same rows, ~100x faster than
        Python); pure host code that releases the GIL, so the drivers format the tracks of a batch on a pool of threads."""
        start = self.start if start is None else start
        end = self.end if end is None else end
        vals = self.vals if vals is None else vals
        if len(vals) != self.end - self.start:
            raise Exception("Error! Inconsistency between length of values and start/end values")
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        n = len(vals)
        if n == 0:
            return b""
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        chrom = self.chrom.encode()
        cap = 48 * n + 64
        buf = np.empty(cap, dtype=np.uint8)
        used = lib.nb200_format_track(chrom, int(start), _lib.ptr(vals, C.c_double), n, int(bool(write_zero)), buf.ctypes.data_as(C.c_char_p), cap)
        if used > cap:
            buf = np.empty(used, dtype=np.uint8)
            used = lib.nb200_format_track(chrom, int(start), _lib.ptr(vals, C.c_double), n, int(bool(write_zero)), buf.ctypes.data_as(C.c_char_p), used)
        return buf[:used].tobytes()

    def write_track(self, handle, start=None, end=None, vals=None, write_zero=True):
        """Track.write_track (tracks.py:37-74) to a text handle, or to a writer that takes bytes (dist.ShardWriter)."""
        text = self.format_track(start=start, end=end, vals=vals, write_zero=write_zero)
        if not text:
            return
        if hasattr(handle, "write_bytes"):
            handle.write_bytes(text)
        else:
            handle.write(text.decode())

    def read_track(self, bedgraph, start=None, end=None, empty=np.nan, flank=None):
        if start:
            self.start = start
        if end:
            self.end = end
        if flank:
            self.start -= flank
            self.end += flank
        self.vals = BedGraphFile(bedgraph).read(self.chrom, self.start, self.end, empty=empty)

    def exp(self):
        self.vals = np.exp(self.vals)
        self.log = False

    def smooth_track(self, window_len, window="flat", sd=None, mode="valid", norm=True):
        self.smoothed = True
        self.vals = smooth(self.vals, window_len, window=window, sd=sd, mode=mode, norm=norm)
        if mode == "valid":
            self.start += window_len // 2
            self.end -= window_len // 2

    def get(self, start=None, end=None, pos=None):
        if pos:  # pos == 0 falls through to the slice branch, as in the reference (tracks.py:112)
            try:
                return self.vals[pos - self.start]
            except Exception:
                raise Exception("Looks like position given doesn't match track")
        start = self.start if start is None else start
        end = self.end if end is None else end
        try:
            return self.vals[start - self.start:end - self.start]
        except Exception:
            raise Exception("Looks like dimensions from get probaby don't match track, or there are no vals in track")

    def slop(self, chromDict, up=0, down=0, new=False):
        lo, hi = (down, up) if self.strand == "-" else (up, down)
        s, e = max(0, self.start - lo), min(chromDict[self.chrom], self.end + hi)
        if new:
            return Track(self.chrom, s, e, name=self.name)
        self.start, self.end = s, e


class InsertionTrack(Track):
    def __init__(self, chrom, start, end):
        Track.__init__(self, chrom, start, end, "insertions")

    def calculateInsertions(self, bamfile, flank=0, lower=0, upper=2000, atac=True):
        self.start -= flank
        self.end += flank
        self.vals = getInsertions(bamfile, self.chrom, self.start, self.end, lower, upper, atac)


class CoverageTrack(Track):
    def __init__(self, chrom, start, end):
        Track.__init__(self, chrom, start, end, "coverage")

    def calculateCoverage(self, mat, lower, upper, window_len):
        """Flat-window coverage of fragment centres (tracks.py:209-222), column sums + window on the device."""
        offset = self.start - mat.start - (window_len // 2)
        if offset < 0:
            raise Exception("Insufficient flanking region on mat to calculate coverage with desired window")
        sub = mat.mat[:, offset:mat.mat.shape[1] - offset] if offset != 0 else mat.mat
        self.vals = default_engine().coverage_dense(sub, lower - mat.lower, upper - mat.lower, window_len)
