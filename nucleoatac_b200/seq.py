"""Sequence fetch + one-hot encoding (pyatac/seq.py:11-45)."""
import numpy as np

from . import hostio

_open = {}


def _fasta(fastafile):
    if isinstance(fastafile, hostio.FastaFile):
        return fastafile
    if fastafile not in _open:
        _open[fastafile] = hostio.FastaFile(fastafile)
    return _open[fastafile]


def get_sequence(chunk, fastafile):
    """Upper-cased sequence of the chunk; reverse-complemented on the minus strand (seq.py:11-22)."""
    seq = _fasta(fastafile).fetch(chunk.chrom, chunk.start, chunk.end).upper()
    return reverse_complement(seq) if chunk.strand == "-" else seq


_COMP = str.maketrans("ACGTNacgtn", "TGCANtgcan")


def complement(sequence):
    return sequence.translate(_COMP)


def reverse_complement(sequence):
    return complement(sequence)[::-1]


def seq_to_mat(sequence, nucleotides):
    """One-hot len(nucleotides) x len(sequence); other letters give an all-zero column (seq.py:37-45)."""
    arr = np.frombuffer(sequence.encode(), dtype=np.uint8)
    mat = np.zeros((len(nucleotides), len(arr)))
    for i, nuc in enumerate(nucleotides):
        mat[i] = arr == ord(nuc)
    return mat
