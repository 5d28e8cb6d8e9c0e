"""BedGraphFile: random access to a (bgzip'd) bedgraph (pyatac/bedgraph.py:6-16)."""
import numpy as np

from . import hostio

_cache = {}


class BedGraphFile:
    def __init__(self, bedgraph):
        if bedgraph not in _cache:
            _cache[bedgraph] = hostio.BedGraphReader(bedgraph)
        self.reader = _cache[bedgraph]

    def read(self, chrom, start, end, empty=np.nan):
        return self.reader.read(chrom, start, end, empty=empty)

    def close(self):
        pass
