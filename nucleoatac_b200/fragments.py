"""The pyatac/fragments.pyx seam: BAM decode on the host, binning arithmetic on the device."""
import numpy as np

from . import hostio
from .engine import default_engine

_open_bams = {}


def _bam(bamfile):
    if isinstance(bamfile, hostio.BamFile):
        return bamfile
    if bamfile not in _open_bams:
        _open_bams[bamfile] = hostio.BamFile(bamfile)
    return _open_bams[bamfile]


def fetch_reads(bamfile, chrom, start, end):
    """(pos, tlen) int32 arrays of the proper-pair forward reads overlapping [start, end) (fragments.pyx:21-25)."""
    return _bam(bamfile).fetch_fragments(chrom, max(0, start), end)


def fetch_reads_many(bamfile, regions):
    """fetch_reads for a list of (chrom, start, end) in one native, multi-threaded BAM decode (nb200_bam_fetch_many):
    -> list of (pos, tlen) array views, one per region."""
    off, pos, tlen = _bam(bamfile).fetch_fragments_many([(c, max(0, s), e) for c, s, e in regions])
    return [(pos[off[i]:off[i + 1]], tlen[off[i]:off[i + 1]]) for i in range(len(regions))]


def makeFragmentMat(bamfile, chrom, start, end, lower, upper, atac=1):
    """pyatac/fragments.pyx:17-40 -> float64 [(upper-lower), (end-start)] count matrix."""
    pos, tlen = fetch_reads(bamfile, chrom, start - upper, end + upper)
    return default_engine().fragmat(pos, tlen, start, end, lower, upper, atac)


def getInsertions(bamfile, chrom, start, end, lower, upper, atac=1):
    """pyatac/fragments.pyx:43-67 -> float64 [end-start] insertion counts."""
    pos, tlen = fetch_reads(bamfile, chrom, start - upper, end + upper)
    return default_engine().insertions(pos, tlen, start, end, lower, upper, atac)


def getFragmentSizesFromChunkList(chunks, bamfile, lower, upper, atac=1):
    """pyatac/fragments.pyx:122-145 -> float64 [upper-lower] counts of fragments centred inside the chunks."""
    chunks = list(chunks)
    if not chunks:
        return np.zeros(upper - lower)
    off, pos, tlen = _bam(bamfile).fetch_fragments_many([(c.chrom, max(0, c.start - upper), c.end + upper) for c in chunks])
    starts, ends = [c.start for c in chunks], [c.end for c in chunks]
    return default_engine().fragment_sizes(starts, ends, off, pos, tlen, lower, upper, atac).astype(np.float64)


def getAllFragmentSizes(bamfile, lower, upper, atac=1):
    """pyatac/fragments.pyx:100-119: histogram over every proper-pair forward read of the file."""
    bam = _bam(bamfile)
    sizes = np.zeros(upper - lower)
    for chrom, length in zip(bam.references, bam.lengths):
        pos, tlen = bam.fetch_fragments(chrom, 0, length)
        ilen = np.abs(tlen.astype(np.int64)) - (8 if atac else 0)
        ok = (ilen >= lower) & (ilen < upper)
        sizes += np.bincount(ilen[ok] - lower, minlength=upper - lower)
    return sizes
