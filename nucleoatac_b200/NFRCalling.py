"""Nucleosome-free regions between adjacent nucleosome calls: the host-side mirror of nucleoatac/NFRCalling.py:16-111
(`NFR`, `NFRParameters`, `NFRChunk`; API names kept so that run_nfr and user code read like the reference's).

Per chunk the device supplies the insertion track (pyatac/fragments.pyx:43-67 -> nb200_insertions) and the Tn5 bias track
(pyatac/bias.py:85-92 -> nb200_bias_track); an NFR record itself is four slice statistics, taken exactly as the reference
takes them (np.mean / np.min over Track.get slices) so that the written rows are text-identical to the shipped
example.nfrpos.bed.gz."""
import numpy as np

from . import hostio
from .bias import PWM, InsertionBiasTrack
from .chunk import Chunk
from .tracks import InsertionTrack, Track
from .utils import fmt12, read_chrom_sizes_from_fasta

NUC_HALF_LEFT, NUC_HALF_RIGHT = 73, 72   # a gap runs from dyad + 73 to the next dyad - 72 (NFRCalling.py:96-97)
_BED_FIELDS = ("occ", "min_upper", "ins_density", "bias_density")


def _slice_stats(tracks, left, right):
    """(mean occupancy, min of the occupancy upper bound, mean insertion density, mean Tn5 bias) over [left, right)."""
    return (np.mean(tracks.occ.get(left, right)), np.min(tracks.occ_upper.get(left, right)),
            np.mean(tracks.ins.get(left, right)), np.mean(tracks.bias.get(left, right, log=False)))


class NFR(Chunk):
    """One region [left, right) with the statistics the nfrpos row carries (NFRCalling.py:16-33)."""

    def __init__(self, left, right, nfrtrack):
        self.chrom, self.start, self.end, self.strand = nfrtrack.chrom, left, right, "*"
        for name, value in zip(_BED_FIELDS, _slice_stats(nfrtrack, left, right)):
            setattr(self, name, value)

    def asBed(self):
        cols = [self.chrom, str(self.start), str(self.end)] + [fmt12(getattr(self, name)) for name in _BED_FIELDS]
        return "\t".join(cols)

    def write(self, handle):
        handle.write(self.asBed() + "\n")


class NFRParameters:
    """What `nucleoatac nfr` is run with (NFRCalling.py:36-48); the PWM and chromosome sizes are only read with a FASTA."""

    def __init__(self, occ_track, calls, ins_track=None, bam=None, max_occ=0.25, max_occ_upper=0.25, fasta=None, pwm=None):
        self.occ_track, self.calls = occ_track, calls
        self.ins_track, self.bam = ins_track, bam
        self.max_occ, self.max_occ_upper = max_occ, max_occ_upper
        self.fasta = fasta
        if fasta is not None:
            self.pwm = PWM.open(pwm)
            self.chrs = read_chrom_sizes_from_fasta(fasta)


_open_calls = {}


def _dyads(path, chrom, start, end):
    """Dyad positions of the calls file rows inside [start, end), in file order (one TabixFile per path and process)."""
    tbx = _open_calls.get(path)
    if tbx is None:
        tbx = _open_calls[path] = hostio.TabixFile(path)
    if chrom not in tbx.contigs:
        return []
    return [int(row[1]) for row in tbx.fetch(chrom, start, end)]


def _upper_bound_path(occ_path):
    """'<out>.occ.bedgraph.gz' -> '<out>.occ.upper_bound.bedgraph.gz' (NFRCalling.py:66)."""
    return occ_path[:-len("bedgraph.gz")] + "upper_bound.bedgraph.gz"


class NFRChunk(Chunk):
    """The tracks of one chunk and the NFRs found in it (NFRCalling.py:51-111)."""

    def __init__(self, chunk):
        self.chrom, self.start, self.end = chunk.chrom, chunk.start, chunk.end
        self.nfrs = []

    def initialize(self, parameters):
        self.params = parameters

    def _bedgraph_track(self, path, name):
        track = Track(self.chrom, self.start, self.end, name)
        track.read_track(path)
        return track

    def getOcc(self):
        """Occupancy and its upper bound, read back from the bedgraphs `occ` wrote (NFRCalling.py:60-70)."""
        self.occ = self._bedgraph_track(self.params.occ_track, "Occupancy")
        self.occ_upper = self._bedgraph_track(_upper_bound_path(self.params.occ_track), "Occupancy")

    def getIns(self):
        """Insertion track: from the BAM on the device, or a bedgraph given with --ins_track (NFRCalling.py:71-78)."""
        if self.params.ins_track is not None:
            self.ins = self._bedgraph_track(self.params.ins_track, "Insertion")
            return
        self.ins = InsertionTrack(self.chrom, self.start, self.end)
        self.ins.calculateInsertions(self.params.bam)

    def getBias(self):
        """log Tn5 bias (all zeros without --fasta, NFRCalling.py:79-84)."""
        self.bias = InsertionBiasTrack(self.chrom, self.start, self.end, log=True)
        if self.params.fasta is not None:
            self.bias.computeBias(self.params.fasta, self.params.chrs, self.params.pwm)

    def findNFRs(self):
        """Gaps between adjacent calls whose mean occupancy and minimal upper bound stay under the thresholds (:86-103)."""
        dyads = _dyads(self.params.calls, self.chrom, self.start, self.end)
        gaps = ((a + NUC_HALF_LEFT, b - NUC_HALF_RIGHT) for a, b in zip(dyads, dyads[1:]))
        for left, right in gaps:
            if right <= left:
                continue
            nfr = NFR(left, right, self)
            if nfr.min_upper < self.params.max_occ_upper and nfr.occ < self.params.max_occ:
                self.nfrs.append(nfr)

    def process(self, params):
        self.initialize(params)
        for step in (self.getOcc, self.getIns, self.getBias, self.findNFRs):
            step()

    def removeData(self):
        self.__dict__.clear()
