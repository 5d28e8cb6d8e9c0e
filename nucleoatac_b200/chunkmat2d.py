"""ChunkMat2D / FragmentMat2D / BiasMat2D (pyatac/chunkmat2d.py:9-156): dense insert-size x position float64
matrices with coordinate slicing.  The object API keeps them dense like the reference; the batched occ/nuc paths
never build them."""
import numpy as np

from .engine import default_engine
from .fragments import makeFragmentMat


class ChunkMat2D:
    def __init__(self, chrom, start, end, lower, upper):
        self.chrom, self.lower, self.upper, self.start, self.end = chrom, lower, upper, start, end
        self.ncol, self.nrow = end - start, upper - lower
        self.mat = np.zeros((self.nrow, self.ncol))

    def get(self, lower=None, upper=None, start=None, end=None, flip=False):
        lower = self.lower if lower is None else lower
        upper = self.upper if upper is None else upper
        start = self.start if start is None else start
        end = self.end if end is None else end
        if flip:
            raise Exception("flip is only used by the pyatac vplot tool and is not part of the scoring path")
        try:
            return self.mat[lower - self.lower:upper - self.lower, start - self.start:end - self.start]
        except Exception:
            raise Exception("Looks like dimensions from get probaby don't match Mat")

    def assign(self, mat):
        if mat.shape != self.mat.shape:
            raise Exception("Dimensions of input mat are wrong.  Uh oh!")
        self.mat = mat

    def save(self, filename):
        np.savetxt(filename, self.mat, delimiter="\t",
                   header=",".join(str(x) for x in (self.chrom, self.start, self.end, self.lower, self.upper)))

    @staticmethod
    def open(filename):
        with open(filename) as fh:
            el = fh.readline().rstrip("\n").lstrip("# ").split(",")
        new = ChunkMat2D(el[0], int(el[1]), int(el[2]), int(el[3]), int(el[4]))
        new.assign(np.loadtxt(filename, skiprows=1))
        return new

    def getIns(self):
        """Collapse the matrix into fragment-end counts; the span shrinks by the pattern width (chunkmat2d.py:74-84)."""
        from .tracks import InsertionTrack
        half = (self.upper + (self.upper - 1) % 2) // 2
        track = InsertionTrack(self.chrom, self.start + half, self.end - half)
        track.assign_track(default_engine().get_ins(self.mat, self.lower, self.upper))
        return track


class FragmentMat2D(ChunkMat2D):
    def __init__(self, chrom, start, end, lower, upper, atac=True):
        ChunkMat2D.__init__(self, chrom, start, end, lower, upper)
        self.atac = atac

    def makeFragmentMat(self, bamfile):
        self.mat = makeFragmentMat(bamfile, self.chrom, self.start, self.end, self.lower, self.upper, self.atac)


class BiasMat2D(ChunkMat2D):
    def __init__(self, chrom, start, end, lower, upper):
        ChunkMat2D.__init__(self, chrom, start, end, lower, upper)
        self.mat = np.ones(self.mat.shape)

    def makeBiasMat(self, bias_track):
        """cell(i, c) = exp(b[c-(i-1)//2] + b[c+i//2]) (chunkmat2d.py:140-153)."""
        offset = self.upper // 2
        bias = bias_track.get(self.start - offset, self.end + offset)
        if not bias_track.log:
            bias = np.log(bias + np.min(bias[bias != 0]))
        if len(bias) != self.ncol + 2 * offset:
            raise Exception("Insufficient flanking region on bias track for the bias matrix")
        self.mat = default_engine().biasmat(bias, self.lower, self.upper)

    def normByInsertDist(self, insertsizes):
        self.mat = self.mat * np.asarray(insertsizes.get(self.lower, self.upper))[:, None]
