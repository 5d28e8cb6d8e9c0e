"""BedGraphFile: random access to a (bgzip'd) bedgraph (pyatac/bedgraph.py:6-16)."""
import numpy as np

from . import hostio

_cache = {}


class BedGraphFile:
    """Uses the tabix index when `<file>.tbi` exists (pysam.Tabixfile.fetch), else one scan of the file."""

    def __init__(self, bedgraph):
        import os
        if bedgraph not in _cache:
            _cache[bedgraph] = hostio.TabixFile(bedgraph) if os.path.exists(bedgraph + ".tbi") else hostio.BedGraphReader(bedgraph)
        self.reader = _cache[bedgraph]

    def read(self, chrom, start, end, empty=np.nan):
        if isinstance(self.reader, hostio.TabixFile):
            # pyatac/bedgraph.py:10-14 in one native call (nb200_bedgraph_fetch: .tbi lookup, inflate, row scan): ~50x the
            # Python reader, which matters because `nuc --occ_track` and `nfr` read three tracks per chunk
            import ctypes as C
            from . import _lib
            out = np.empty(max(0, end - start), dtype=np.float64)
            err = C.create_string_buffer(256)
            st = _lib.load().nb200_bedgraph_fetch(self.reader.path.encode(), chrom.encode(), int(start), int(end), float(empty),
                                                  _lib.ptr(out, C.c_double), err, 256)
            if st != 0:
                raise IOError("reading %s failed: %s" % (self.reader.path, err.value.decode()))
            return out
        return self.reader.read(chrom, start, end, empty=empty)

    def read_python(self, chrom, start, end, empty=np.nan):
        """The same through the Python tabix reader (cross-check of the native one in the tests)."""
        out = np.ones(end - start) * empty
        for row in self.reader.fetch(chrom, start, end):
            out[max(int(row[1]) - start, 0):min(int(row[2]) - start, end - start)] = float(row[3])
        return out

    def close(self):
        pass
