"""Stdlib readers for the reference's example inputs (oracle side, test infrastructure).

The reference reads BAM / FASTA / tabix bedgraph through pysam (htslib), which is
not in this image.  These readers decode the same bytes with zlib/struct only and
are used to (1) validate the oracle against ``example/example_results`` and
(2) extract the compact fixtures in ``tests/golden``.  They replace
``AlignmentFile.fetch`` (pyatac/fragments.pyx:21-24), ``FastaFile.fetch``
(pyatac/seq.py:17-18) and ``Tabixfile.fetch`` (pyatac/bedgraph.py:9-14) for whole
small files; they are not region-indexed.
"""
import gzip
import struct

import numpy as np


def read_bam_fragments(path):
    """Decode a BAM into per-reference arrays of the reads the reference keeps.

    Keeps reads with ``is_proper_pair and not is_reverse`` (flag&0x2 and not
    flag&0x10), exactly the filter of pyatac/fragments.pyx:25,50,131.  Returns
    ``(chrom_sizes: dict name->len, frags: dict name->(pos int32[], tlen int32[]))``
    in file (coordinate-sorted) order.
    """
    with gzip.open(path, "rb") as fh:  # BGZF = concatenated gzip members
        raw = fh.read()
    if raw[:4] != b"BAM\x01":
        raise ValueError("not a BAM file: %s" % path)
    (l_text,) = struct.unpack_from("<i", raw, 4)
    off = 8 + l_text
    (n_ref,) = struct.unpack_from("<i", raw, off)
    off += 4
    names, sizes = [], {}
    for _ in range(n_ref):
        (l_name,) = struct.unpack_from("<i", raw, off)
        off += 4
        name = raw[off:off + l_name - 1].decode()
        off += l_name
        (l_ref,) = struct.unpack_from("<i", raw, off)
        off += 4
        names.append(name)
        sizes[name] = l_ref
    pos_l = {n: [] for n in names}
    tlen_l = {n: [] for n in names}
    n = len(raw)
    while off < n:
        (block_size,) = struct.unpack_from("<i", raw, off)
        ref_id, pos, _lrn, _mapq, _bin, _ncig, flag, _lseq, _nref, _npos, tlen = struct.unpack_from(
            "<iiBBHHHiiii", raw, off + 4)
        off += 4 + block_size
        if ref_id < 0:
            continue
        if (flag & 0x2) and not (flag & 0x10):
            name = names[ref_id]
            pos_l[name].append(pos)
            tlen_l[name].append(tlen)
    frags = {k: (np.asarray(pos_l[k], dtype=np.int32), np.asarray(tlen_l[k], dtype=np.int32))
             for k in names}
    return sizes, frags


class Fasta:
    """``.fai``-indexed FASTA fetch (replaces pysam.FastaFile, pyatac/seq.py:17-22)."""

    def __init__(self, path):
        self.path = path
        self.index = {}
        with open(path + ".fai") as fh:
            for line in fh:
                name, length, offset, linebases, linewidth = line.rstrip("\n").split("\t")[:5]
                self.index[name] = (int(length), int(offset), int(linebases), int(linewidth))
        self.references = list(self.index.keys())
        self.lengths = [self.index[k][0] for k in self.references]

    def chrom_sizes(self):
        """pyatac/utils.py:104-113."""
        return {k: self.index[k][0] for k in self.references}

    def fetch(self, chrom, start, end):
        length, offset, linebases, linewidth = self.index[chrom]
        start = max(0, start)
        end = min(length, end)
        if end <= start:
            return ""
        b0 = offset + (start // linebases) * linewidth + start % linebases
        b1 = offset + ((end - 1) // linebases) * linewidth + (end - 1) % linebases + 1
        with open(self.path, "rb") as fh:
            fh.seek(b0)
            data = fh.read(b1 - b0)
        return data.replace(b"\n", b"").replace(b"\r", b"").decode().upper()  # seq.py:22 upper()


def read_bedgraph_gz(path):
    """Whole bgzip'd bedgraph/bed as list of split rows (replaces Tabixfile.fetch)."""
    rows = []
    with gzip.open(path, "rt") as fh:
        for line in fh:
            if line.strip():
                rows.append(line.rstrip("\n").split("\t"))
    return rows


def bedgraph_region(rows, chrom, start, end, empty=np.nan):
    """pyatac/bedgraph.py:9-14 on pre-read rows: fill [start,end) from overlapping rows."""
    out = np.ones(end - start) * empty
    for r in rows:
        if r[0] != chrom:
            continue
        s, e = int(r[1]), int(r[2])
        if e <= start or s >= end:
            continue
        out[max(s - start, 0):min(e - start, end - start)] = float(r[3])
    return out
