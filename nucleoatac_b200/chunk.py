"""Chunk / ChunkList: BED regions = the unit of work and of multi-GPU sharding (pyatac/chunk.py:11-217)."""
import gzip
import warnings


class Chunk:
    """A genomic interval [start, end) (pyatac/chunk.py:11-54)."""

    def __init__(self, chrom, start, end, weight=1, name="region", strand="*"):
        self.chrom, self.start, self.end = chrom, start, end
        self.weight, self.strand, self.name = weight, strand, name

    def length(self):
        return self.end - self.start

    def asBed(self):
        return "\t".join(str(x) for x in (self.chrom, self.start, self.end, self.weight, self.name, self.strand))

    def slop(self, chromDict, up=0, down=0, new=False):
        """Extend within the chromosome; `up`/`down` swap on the minus strand (chunk.py:26-40)."""
        lo, hi = (down, up) if self.strand == "-" else (up, down)
        s, e = max(0, self.start - lo), min(chromDict[self.chrom], self.end + hi)
        if new:
            return Chunk(self.chrom, s, e, weight=self.weight, name=self.name, strand=self.strand)
        self.start, self.end = s, e

    def center(self, new=False):
        half = self.length() // 2  # Python-2 integer division in the reference (chunk.py:43-47)
        if self.strand == "-":
            e = self.end - half
            s = e - 1
        else:
            s = self.start + half
            e = s + 1
        if new:
            return Chunk(self.chrom, s, e, weight=self.weight, name=self.name, strand=self.strand)
        self.start, self.end = s, e


class ChunkList(list):
    """List of Chunks with BED reading, slop / merge / split (pyatac/chunk.py:71-217)."""

    def __init__(self, *args):
        list.__init__(self, args)

    def sort(self):
        # pyatac/chunk.py:57-69,97 orders by chromosome name then start (its comparator never returns 1 on
        # start, :66 -- only already-sorted BEDs are safe in the reference; a total order is used here)
        list.sort(self, key=lambda c: (c.chrom, c.start, c.end))

    def isSorted(self):
        return all((self[i].chrom, self[i].start) <= (self[i + 1].chrom, self[i + 1].start) for i in range(len(self) - 1))

    def slop(self, chromDict, up=0, down=0, new=False):
        if new:
            return ChunkList(*[c.slop(chromDict, up=up, down=down, new=True) for c in self])
        for c in self:
            c.slop(chromDict, up=up, down=down)

    def merge(self, new=False, sep=-1):
        """Merge neighbours closer than `sep` (chunk.py:109-125); an unsorted list is sorted first, as the reference does."""
        if not self.isSorted():
            self.sort()
        out = ChunkList()
        if len(self):
            prev = Chunk(self[0].chrom, self[0].start, self[0].end, weight=self[0].weight, name=self[0].name, strand=self[0].strand)
            for c in self[1:]:
                if c.chrom == prev.chrom and c.start <= prev.end + sep:
                    prev.end = max(c.end, prev.end)
                else:
                    out.append(prev)
                    prev = Chunk(c.chrom, c.start, c.end, weight=c.weight, name=c.name, strand=c.strand)
            out.append(prev)
        if new:
            return out
        self[:] = out

    def asBed(self):
        return "".join(c.asBed() + "\n" for c in self)

    @staticmethod
    def read(bedfile, weight_col=None, strand_col=None, name_col=None, chromDict=None, min_offset=None, min_length=1,
             chrom_source="FASTA file"):
        """Tab-delimited BED -> ChunkList, clipping to `min_offset` from the chromosome ends (chunk.py:132-175)."""
        opener = gzip.open if bedfile.endswith(".gz") else open
        out, bad = ChunkList(), set()
        weight, strand, name = None, "+", None
        with opener(bedfile, "rt") as fh:
            for line in fh:
                f = line.rstrip("\n").split("\t")
                if len(f) < 3:
                    continue
                if weight_col:
                    weight = f[weight_col - 1]
                if strand_col:
                    strand = f[strand_col - 1]
                if name_col:
                    name = f[name_col - 1]
                chrom, start, end = f[0], int(f[1]), int(f[2])
                if chromDict is not None and chrom not in chromDict:
                    bad.add(chrom)
                    continue
                if min_offset:
                    start = max(start, min_offset)
                    end = min(end, chromDict[chrom] - min_offset)
                if end - start >= min_length:
                    out.append(Chunk(chrom, start, end, weight=weight, strand=strand, name=name))
        if bad:
            warnings.warn("%d chromosome names in bed file not included in %s:\n%s\n These regions will be ignored in "
                          "subsequent analysis" % (len(bad), chrom_source, "\n".join(sorted(bad))))
        return out

    @staticmethod
    def convertChromSizes(chromDict, splitsize=None, offset=0):
        out = ChunkList()
        for chrom in sorted(chromDict):
            if splitsize is None:
                out.append(Chunk(chrom, offset, chromDict[chrom] - offset))
            else:
                for i in range(offset, chromDict[chrom] - offset, splitsize):
                    out.append(Chunk(chrom, i, min(i + splitsize, chromDict[chrom] - offset)))
        return out

    def split(self, bases=None, items=None):
        """Sub-lists of at most `items` chunks, or of roughly `bases` bp (chunk.py:188-208)."""
        if bases is not None:
            out, i, acc, k = [], 0, 0, 0
            for k in range(len(self)):
                acc += self[k].length()
                if acc > bases:
                    out.append(self[i:k + 1])
                    acc, i = 0, k + 1
            if len(self) and k >= i:
                out.append(self[i:k + 1])
            return out
        if items is not None:
            return [self[i:i + items] for i in range(0, len(self), items)]
        raise Exception("Need to provide items or bases argument!")

    def checkChroms(self, chroms, chunklist_source="bed file", chrom_source="fasta file",
                    warn="Regions on these chromosomes will be ignored in analysis"):
        bad = {c.chrom for c in self if c.chrom not in chroms}
        if bad:
            self[:] = [c for c in self if c.chrom in chroms]
            warnings.warn("%d chromosome names in %s not included in %s:\n%s\n %s" % (
                len(bad), chunklist_source, chrom_source, "\n".join(sorted(bad)), warn))
