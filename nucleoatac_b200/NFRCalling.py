"""NFR / NFRParameters / NFRChunk (nucleoatac/NFRCalling.py:16-111): nucleosome-free regions between adjacent calls.

The per-chunk work is the insertion track (pyatac/fragments.pyx:43-67 -> nb200_insertions) and the Tn5 bias track
(pyatac/bias.py:85-92 -> nb200_bias_track), both on the device; the per-region statistics are a handful of numpy means."""
import numpy as np

from . import hostio
from .bias import PWM, InsertionBiasTrack
from .chunk import Chunk
from .tracks import InsertionTrack, Track
from .utils import fmt12, read_chrom_sizes_from_fasta


class NFR(Chunk):
    """One NFR region: [left, right) with its mean occupancy, the minimum of the occupancy upper bound, the mean insertion
    density and the mean Tn5 bias (NFRCalling.py:16-33)."""

    def __init__(self, left, right, nfrtrack):
        self.chrom = nfrtrack.chrom
        self.start = left
        self.end = right
        self.strand = "*"
        self.occ = np.mean(nfrtrack.occ.get(left, right))
        self.min_upper = np.min(nfrtrack.occ_upper.get(left, right))
        self.ins_density = np.mean(nfrtrack.ins.get(left, right))
        self.bias_density = np.mean(nfrtrack.bias.get(left, right, log=False))

    def asBed(self):
        return "\t".join([self.chrom, str(self.start), str(self.end), fmt12(self.occ), fmt12(self.min_upper),
                          fmt12(self.ins_density), fmt12(self.bias_density)])

    def write(self, handle):
        handle.write(self.asBed() + "\n")


class NFRParameters:
    def __init__(self, occ_track, calls, ins_track=None, bam=None, max_occ=0.25, max_occ_upper=0.25, fasta=None, pwm=None):
        self.bam = bam
        self.ins_track = ins_track
        self.occ_track = occ_track
        self.calls = calls
        self.max_occ = max_occ
        self.max_occ_upper = max_occ_upper
        self.fasta = fasta
        if fasta is not None:
            self.pwm = PWM.open(pwm)
            self.chrs = read_chrom_sizes_from_fasta(fasta)


_calls_cache = {}


def _calls(path):
    if path not in _calls_cache:
        _calls_cache[path] = hostio.TabixFile(path)
    return _calls_cache[path]


class NFRChunk(Chunk):
    def __init__(self, chunk):
        self.start = chunk.start
        self.end = chunk.end
        self.chrom = chunk.chrom
        self.nfrs = []

    def initialize(self, parameters):
        self.params = parameters

    def getOcc(self):
        """Occupancy and its upper bound from the bedgraphs `occ` wrote (NFRCalling.py:60-70)."""
        self.occ = Track(self.chrom, self.start, self.end, "Occupancy")
        self.occ.read_track(self.params.occ_track)
        upper_file = self.params.occ_track[:-11] + "upper_bound.bedgraph.gz"
        self.occ_upper = Track(self.chrom, self.start, self.end, "Occupancy")
        self.occ_upper.read_track(upper_file)

    def getIns(self):
        if self.params.ins_track is None:
            self.ins = InsertionTrack(self.chrom, self.start, self.end)
            self.ins.calculateInsertions(self.params.bam)
        else:
            self.ins = Track(self.chrom, self.start, self.end, "Insertion")
            self.ins.read_track(self.params.ins_track)

    def getBias(self):
        self.bias = InsertionBiasTrack(self.chrom, self.start, self.end, log=True)
        if self.params.fasta is not None:
            self.bias.computeBias(self.params.fasta, self.params.chrs, self.params.pwm)

    def findNFRs(self):
        """Gaps [dyad + 73, next dyad - 72) between adjacent calls that pass the occupancy thresholds (NFRCalling.py:86-103)."""
        tbx = _calls(self.params.calls)
        nucs = []
        if self.chrom in tbx.contigs:
            for row in tbx.fetch(self.chrom, self.start, self.end):
                nucs.append(int(row[1]))
        for j in range(1, len(nucs)):
            left = nucs[j - 1] + 73
            right = nucs[j] - 72
            if right <= left:
                continue
            candidate = NFR(left, right, self)
            if candidate.min_upper < self.params.max_occ_upper and candidate.occ < self.params.max_occ:
                self.nfrs.append(candidate)

    def process(self, params):
        self.initialize(params)
        self.getOcc()
        self.getIns()
        self.getBias()
        self.findNFRs()

    def removeData(self):
        for name in list(self.__dict__.keys()):
            delattr(self, name)
