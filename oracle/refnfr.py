"""CPU restatement of NFRChunk.process (nucleoatac/NFRCalling.py:51-111) -- TEST INFRASTRUCTURE, not shipped code paths.

An NFR candidate is the gap between two adjacent nucleosome calls of the combined map: [dyad_j-1 + 73, dyad_j - 72)
(NFRCalling.py:95-100).  It is kept when the mean occupancy and the minimum of the occupancy upper bound over the gap are
below the thresholds (:101-103); its record carries the mean insertion density and the mean Tn5 bias exp(b) (:18-28).

Pinned on the outputs the reference shipped (example_results/example.nfrpos.bed.gz, example.ins.bedgraph.gz) by
tests/test_oracle_golden.py through tests/golden/nfr_golden.npz.
"""
import numpy as np

from . import refalgo as ra


def find_nfrs(start, end, nuc_positions, occ, occ_upper, ins, log_bias, max_occ=0.25, max_occ_upper=0.25):
    """nuc_positions: dyads of the calls fetched for [start, end) in file order (NFRCalling.py:90-94);
    occ / occ_upper / ins / log_bias: float64 tracks over [start, end).  -> list of
    (left, right, occ, min_upper, ins_density, bias_density)."""
    out = []
    nucs = [int(p) for p in nuc_positions]
    for j in range(1, len(nucs)):
        left = nucs[j - 1] + 73
        right = nucs[j] - 72
        if right <= left:
            continue
        a, b = left - start, right - start  # Track.get: plain slice of vals (pyatac/tracks.py:108-127)
        o = np.mean(occ[a:b])
        mu = np.min(occ_upper[a:b])
        if mu < max_occ_upper and o < max_occ:
            out.append((left, right, o, mu, np.mean(ins[a:b]), np.mean(np.exp(log_bias[a:b]))))
    return out


def process_nfr_chunk(pos, tlen, start, end, nuc_positions, occ, occ_upper, log_bias, max_occ=0.25, max_occ_upper=0.25,
                      atac=True):
    """-> (nfr records, insertion track).  Insertions: InsertionTrack.calculateInsertions with its defaults
    lower=0, upper=2000 (pyatac/tracks.py:159-163 -> fragments.pyx:43-67)."""
    ins = ra.get_insertions(pos, tlen, start, end, 0, 2000, atac)
    return find_nfrs(start, end, nuc_positions, occ, occ_upper, ins, log_bias, max_occ, max_occ_upper), ins


def nfr_bed(chrom, rec):
    """NFR.asBed, NFRCalling.py:29-31 (py2 str() of floats = 12 significant digits)."""
    left, right, o, mu, idens, bdens = rec
    return "\t".join([chrom, str(left), str(right), ra.fmt12(o), ra.fmt12(mu), ra.fmt12(idens), ra.fmt12(bdens)])
