"""FragmentSizes: insert-size distribution container + text IO (pyatac/fragmentsizes.py:15-81)."""
import numpy as np

from .fragments import getAllFragmentSizes, getFragmentSizesFromChunkList
from .utils import fmt12


class FragmentSizes:
    def __init__(self, lower, upper, atac=True, vals=None):
        self.lower, self.upper, self.vals, self.atac = lower, upper, vals, atac

    def calculateSizes(self, bamfile, chunks=None):
        if chunks is None:
            sizes = getAllFragmentSizes(bamfile, self.lower, self.upper, atac=self.atac)
        else:
            sizes = getFragmentSizesFromChunkList(chunks, bamfile, self.lower, self.upper, atac=self.atac)
        total = np.sum(sizes)
        self.vals = sizes / (total + (total == 0))

    def get(self, lower=None, upper=None, size=None):
        if size:  # size 0 falls through to the slice branch, like the reference (fragmentsizes.py:29)
            try:
                return self.vals[size - self.lower]
            except Exception:
                raise Exception("Looks like size doesn't match FragmentSizes")
        lower = self.lower if lower is None else lower
        upper = self.upper if upper is None else upper
        try:
            return self.vals[lower - self.lower:upper - self.lower]
        except Exception:
            raise Exception("Looks like dimensions from get probaby don't match FragmentSizes")

    def save(self, filename):
        with open(filename, "w") as fh:
            fh.write("#lower\n%s\n#upper\n%s\n#sizes\n" % (self.lower, self.upper))
            fh.write("\t".join(fmt12(v) for v in self.get()) + "\n")

    @staticmethod
    def open(filename):
        state, lower, upper, vals = "", None, None, None
        with open(filename) as fh:
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
        if lower is None or upper is None or vals is None:
            raise Exception("FragmentSizes file appears to be missing some needed components")
        return FragmentSizes(lower, upper, vals=vals)
