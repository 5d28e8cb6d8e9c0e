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
            out = np.ones(end - start) * empty  # pyatac/bedgraph.py:10-14
            for row in self.reader.fetch(chrom, start, end):
                out[max(int(row[1]) - start, 0):min(int(row[2]) - start, end - start)] = float(row[3])
            return out
        return self.reader.read(chrom, start, end, empty=empty)

    def close(self):
        pass
