"""`nucleoatac merge` (nucleoatac/merge.py:15-104): combine the occupancy peaks of `occ` with the nucleosome calls of
`nuc` into one position map -- nuc calls win; an occupancy peak is kept only when no nuc call lies within `sep`."""
import gzip
import os

from . import hostio
from .chunk import Chunk, ChunkList
from .utils import fmt12


class MergedNuc(Chunk):
    def __init__(self, chrom, start, end, occ, occ_lower, occ_upper, reads, source):
        self.chrom, self.start, self.end = chrom, start, end
        self.occ, self.occ_lower, self.occ_upper, self.reads, self.source = occ, occ_lower, occ_upper, reads, source

    def asBed(self):
        return "\t".join([self.chrom, str(self.start), str(self.end), fmt12(self.occ), fmt12(self.occ_lower),
                          fmt12(self.occ_upper), fmt12(self.reads), self.source])

    def write(self, handle):
        handle.write(self.asBed() + "\n")


class NucList(ChunkList):
    @staticmethod
    def read(bedfile, source, min_occ=0):
        """occpeaks.bed (source 'occ') or nucpos.bed (source 'nuc') rows with occ_lower >= min_occ."""
        if source not in ("occ", "nuc"):
            raise Exception("source must be 'occ' or 'nuc'")
        out = NucList()
        opener = gzip.open if bedfile.endswith(".gz") else open
        with opener(bedfile, "rt") as fh:
            for line in fh:
                f = line.rstrip("\n").split("\t")
                if len(f) < 7:
                    continue
                if source == "occ":
                    occ, lo, up, reads = float(f[3]), float(f[4]), float(f[5]), float(f[6])
                else:
                    occ, lo, up, reads = float(f[4]), float(f[5]), float(f[6]), float(f[10]) + float(f[11])
                if lo >= min_occ:  # NaN (nuc run without --occ_track) compares False, like the reference
                    out.append(MergedNuc(f[0], int(f[1]), int(f[2]), occ, lo, up, reads, source))
        return out


def merge(occ_peaks, nuc_calls, sep=120):
    """Two-pointer walk over both sorted lists (merge.py:66-91)."""
    keep = NucList()
    i = j = 0
    while i < len(occ_peaks) and j < len(nuc_calls):
        o, n = occ_peaks[i], nuc_calls[j]
        if o.chrom < n.chrom or (o.chrom == n.chrom and o.start < n.start - sep):
            keep.append(o)
            i += 1
        elif o.chrom > n.chrom or o.start > n.start + sep:
            keep.append(n)
            j += 1
        else:
            i += 1  # an occupancy peak next to a nuc call is dropped
    keep.extend(nuc_calls[j:])
    keep.extend(occ_peaks[i:])
    return keep


def run_merge(args):
    if not args.out:
        args.out = ".".join(os.path.basename(args.nucpos).split(".")[0:-3])
    occ = NucList.read(args.occpeaks, "occ", float(args.min_occ))
    nuc = NucList.read(args.nucpos, "nuc", float(args.min_occ))
    new = merge(occ, nuc, int(args.sep))
    plain = args.out + ".nucmap_combined.bed"
    with open(plain, "w") as fh:
        fh.write(new.asBed())
    hostio.bgzip_tabix(plain, plain + ".gz")
    os.remove(plain)
    return new
