"""Nucleosome calling (nucleoatac/NucleosomeCalling.py:25-345): V-plot cross-correlation signal, bias background,
normalised + smoothed signal, candidate calls with likelihood ratio and z-score.  `NucChunk.process` runs the chunk on
the device (nb200_nuc_run); only the gaussian-mixture fuzziness fit (getFuzz, L-BFGS-B) stays on the host."""
from bisect import bisect_left

import numpy as np

from . import seq as _seq
from .bias import PWM, InsertionBiasTrack
from .chunk import Chunk
from .chunkmat2d import BiasMat2D, FragmentMat2D
from .engine import FLAG_NONREDUNDANT, FLAG_Z, PackedBatch, default_engine
from .fragments import fetch_reads, fetch_reads_many
from .multinomial_cov import calculateCov
from .tracks import CoverageTrack, InsertionTrack, Track
from .utils import call_peaks, fmt12, read_chrom_sizes_from_bam, reduce_peaks


class SignalTrack(Track):
    def __init__(self, chrom, start, end):
        Track.__init__(self, chrom, start, end, "signal")

    def calculateSignal(self, mat, vmat):
        """Valid 2-D cross-correlation of the fragment matrix with the VMat (NucleosomeCalling.py:29-36)."""
        offset = self.start - mat.start - vmat.w
        if offset < 0:
            raise Exception("Insufficient flanking region on mat to calculate signal")
        eng = default_engine()
        eng.set_vmat(vmat.mat, vmat.lower, vmat.upper)
        self.vals = eng.xcor_dense(mat.get(vmat.lower, vmat.upper, mat.start + offset, mat.end - offset))


class NormSignalTrack(Track):
    def __init__(self, chrom, start, end):
        Track.__init__(self, chrom, start, end, "normalized signal")

    def calculateNormSignal(self, raw, bias):
        self.vals = raw.get(self.start, self.end) - bias.get(self.start, self.end)


class BiasTrack(Track):
    def __init__(self, chrom, start, end):
        Track.__init__(self, chrom, start, end, "bias")

    def calculateBackgroundSignal(self, mat, vmat, nuc_cov):
        """Expected signal under the bias model, scaled to the observed coverage (NucleosomeCalling.py:49-64)."""
        offset = self.start - mat.start - vmat.w
        if offset < 0:
            raise Exception("Insufficient flanking region on mat to calculate signal")
        self.vmat, self.bias_mat = vmat, mat
        self.cov = CoverageTrack(self.chrom, self.start, self.end)
        self.cov.calculateCoverage(self.bias_mat, vmat.lower, vmat.upper, vmat.w * 2 + 1)
        self.nuc_cov = nuc_cov.vals
        eng = default_engine()
        eng.set_vmat(vmat.mat, vmat.lower, vmat.upper)
        xc = eng.xcor_dense(mat.get(vmat.lower, vmat.upper, mat.start + offset, mat.end - offset))
        self.vals = xc * self.nuc_cov / self.cov.vals


class SignalDistribution:
    """Distribution of the signal under the bias model at one position (NucleosomeCalling.py:68-88)."""

    def __init__(self, position, vmat, bias_mat, reads):
        self.position, self.reads, self.vmat = position, reads, vmat
        sub = bias_mat.get(vmat.lower, vmat.upper, position - vmat.w, position + vmat.w + 1)
        self.prob_mat = sub / np.sum(sub)
        self.probs = self.prob_mat.flatten()

    def simulateReads(self):
        return np.reshape(np.random.multinomial(self.reads, self.probs), self.vmat.mat.shape)

    def simulateDist(self, numiters=1000):
        self.scores = [np.sum(self.simulateReads() * self.vmat.mat) for _ in range(numiters)]

    def analStd(self):
        return np.sqrt(calculateCov(self.probs, np.ravel(self.vmat.mat), self.reads))

    def analMean(self):
        return np.sum(self.prob_mat * self.vmat.mat * self.reads)


def norm(x, v, w, mean):
    """Normal pdf with variance v rescaled to peak height w (NucleosomeCalling.py:92-97)."""
    n = 1.0 / np.sqrt(2 * np.pi * v) * np.exp(-(x - mean) ** 2 / (2 * v))
    return n * (w / max(n))


class Nucleosome(Chunk):
    def __init__(self, pos, nuctrack):
        self.chrom, self.start, self.end = nuctrack.chrom, pos, pos + 1
        i = pos - nuctrack.start
        self.nfr_cov = nuctrack.nfr_cov.vals[i]
        self.nuc_cov = nuctrack.nuc_cov.vals[i]
        self.nuc_signal = nuctrack.nuc_signal.vals[i]
        self.norm_signal = nuctrack.norm_signal.vals[i]
        self.smoothed = nuctrack.smoothed.vals[i]

    def getLR(self, nuctrack):
        """Log-likelihood ratio of the V-plot model over the bias model (NucleosomeCalling.py:110-122), dense path."""
        p, w = nuctrack.params, nuctrack.params.vmat.w
        mat = nuctrack.mat.get(p.lower, p.upper, self.start - w, self.start + w + 1)
        null_mat = nuctrack.bias_mat.get(p.lower, p.upper, self.start - w, self.start + w + 1)
        bias_mat = nuctrack.bias_mat_prenorm.get(p.lower, p.upper, self.start - w, self.start + w + 1)
        with np.errstate(divide="ignore", invalid="ignore"):
            nuc_model = p.vmat.mat * bias_mat
            nuc_model = nuc_model / np.sum(nuc_model)
            null_model = null_mat / np.sum(null_mat)
            self.lr = np.sum(np.log(nuc_model) * mat) - np.sum(np.log(null_model) * mat)

    def getZScore(self, nuctrack):
        s = SignalDistribution(self.start, nuctrack.params.vmat, nuctrack.bias_mat, self.nuc_cov)
        self.z = self.norm_signal / s.analStd()

    def getOcc(self, nuctrack):
        try:
            i = self.start - nuctrack.start
            self.occ, self.occ_lower, self.occ_upper = nuctrack.occ.vals[i], nuctrack.occ_lower.vals[i], nuctrack.occ_upper.vals[i]
        except Exception:
            self.occ = self.occ_lower = self.occ_upper = np.nan

    def fuzz_job(self, nuctrack):
        """What the fuzziness fit of this call needs (NucleosomeCalling.py:137-194): (left edge, job for fuzz.fit_fuzz)."""
        sep = nuctrack.params.nonredundant_sep
        index = self.start - nuctrack.start
        allnucs = nuctrack.sorted_nuc_keys
        x = bisect_left(allnucs, index)
        if x > 0 and index - allnucs[x - 1] < sep:
            left, means = allnucs[x - 1], (index - allnucs[x - 1], 0)
        else:
            left, means = index - sep // 3, (sep // 3,)
        if x < len(allnucs) - 1 and allnucs[x + 1] - index < sep:
            right = allnucs[x + 1]
            means += (allnucs[x + 1] - left,)
        else:
            right = index + sep // 3 + 1
        sig = nuctrack.smoothed.vals[left:right]
        sig[sig < 0] = 0          # in place, like the reference: later fits of the chunk see the clipped signal
        return int(left), (np.array(sig), tuple(int(m) for m in means), nuctrack.params.smooth_sd)

    def set_fuzz(self, left, res):
        self.fuzz, self.weight, self.fit_pos = res[0], res[1], res[2] + left

    def getFuzz(self, nuctrack):
        """Fit 1-3 gaussians to the smoothed signal around the call (NucleosomeCalling.py:137-194); host scipy."""
        from . import fuzz
        left, job = self.fuzz_job(nuctrack)
        self.set_fuzz(left, fuzz.fit_fuzz(job))

    def asBed(self):
        return "\t".join([self.chrom, str(self.start), str(self.end)] + [fmt12(getattr(self, k, np.nan)) for k in (
            "z", "occ", "occ_lower", "occ_upper", "lr", "norm_signal", "nuc_signal", "nuc_cov", "nfr_cov", "fuzz")])

    def write(self, handle):
        handle.write(self.asBed() + "\n")


class NucParameters:
    """Parameters of the nucleosome-calling run (NucleosomeCalling.py:204-226) + the device engine for them."""

    def __init__(self, vmat, fragmentsizes, bam, fasta, pwm, occ_track=None, atac=True, sd=25, nonredundant_sep=120,
                 redundant_sep=25, min_z=3, min_lr=0, min_reads=1, device=0, xcor_mode=0):
        self.atac, self.vmat = atac, vmat
        self.lower, self.upper = vmat.lower, vmat.upper
        self.window = vmat.mat.shape[1]
        self.fragmentsizes = fragmentsizes
        self.min_reads, self.min_z, self.min_lr = min_reads, min_z, min_lr
        self.smooth_sd, self.redundant_sep, self.nonredundant_sep = sd, redundant_sep, nonredundant_sep
        self.fasta = fasta
        self.pwm = PWM.open(pwm)
        self.chrs = read_chrom_sizes_from_bam(bam)
        self.bam, self.occ_track = bam, occ_track
        self.device, self.xcor_mode = device, xcor_mode

    def engine(self):
        eng = default_engine(self.device)
        if getattr(eng, "_nuc_owner", None) is not self:
            eng.set_pwm(self.pwm.mat, self.pwm.up, self.pwm.down, self.pwm.nucleotides)
            eng.set_vmat(self.vmat.mat, self.vmat.lower, self.vmat.upper)
            eng.set_fragment_sizes(self.fragmentsizes.get(0, self.upper))
            eng.configure_nuc(sd=self.smooth_sd, nonredundant_sep=self.nonredundant_sep, redundant_sep=self.redundant_sep,
                              min_z=self.min_z, min_lr=self.min_lr, min_reads=self.min_reads, atac=self.atac,
                              use_bias=self.fasta is not None, xcor_mode=self.xcor_mode)
            eng._nuc_owner = self
            eng._occ_owner = None
        return eng

    def pad(self):
        return max(self.window, self.upper // 2 + 1)

    def pack(self, chunks):
        items = []
        pad = self.pad()
        reads = fetch_reads_many(self.bam, [(c.chrom, c.start - pad - self.upper, c.end + pad + self.upper) for c in chunks])
        for c, (pos, tlen) in zip(chunks, reads):
            sq, s0 = None, 0
            if self.fasta is not None:
                s0 = c.start - self.window - self.upper // 2 - self.pwm.up
                e0 = c.end + self.window + self.upper // 2 + 1 + self.pwm.down
                if s0 < 0 or e0 > self.chrs[c.chrom]:
                    raise Exception("Insufficient flanking region on chromosome for bias track: " + c.asBed())
                sq = _seq._fasta(self.fasta).fetch(c.chrom, s0, e0).encode()
            items.append((c.start, c.end, pos, tlen, sq, s0))
        return PackedBatch.from_chunks(items)


class NucChunk(Chunk):
    def __init__(self, chunk):
        self.start, self.end, self.chrom = chunk.start, chunk.end, chunk.chrom

    def initialize(self, parameters):
        self.params = parameters

    # ---- step-by-step object API (dense matrices through the primitives)
    def getFragmentMat(self):
        pad = self.params.pad()
        self.mat = FragmentMat2D(self.chrom, self.start - pad, self.end + pad, 0, self.params.upper, atac=self.params.atac)
        self.mat.makeFragmentMat(self.params.bam)

    def makeBiasMat(self):
        p = self.params
        self.bias_mat = BiasMat2D(self.chrom, self.start - p.window, self.end + p.window, 0, p.upper)
        if p.fasta is not None:
            bt = InsertionBiasTrack(self.chrom, self.start - p.window - p.upper // 2, self.end + p.window + p.upper // 2 + 1, log=True)
            bt.computeBias(p.fasta, p.chrs, p.pwm)
            self.bias_mat.makeBiasMat(bt)
        self.bias_mat_prenorm = BiasMat2D(self.chrom, self.start - p.window, self.end + p.window, 0, p.upper)
        self.bias_mat_prenorm.mat = np.array(self.bias_mat.mat)
        self.bias_mat.normByInsertDist(p.fragmentsizes)

    def getNucSignal(self):
        p = self.params
        self.nuc_cov = CoverageTrack(self.chrom, self.start, self.end)
        self.nuc_cov.calculateCoverage(self.mat, p.lower, p.upper, p.window)
        self.bias = BiasTrack(self.chrom, self.start, self.end)
        self.bias.calculateBackgroundSignal(self.bias_mat, p.vmat, self.nuc_cov)
        self.nuc_signal = SignalTrack(self.chrom, self.start, self.end)
        self.nuc_signal.calculateSignal(self.mat, p.vmat)
        self.norm_signal = NormSignalTrack(self.chrom, self.start, self.end)
        self.norm_signal.calculateNormSignal(self.nuc_signal, self.bias)

    def getNFR(self):
        self.nfr_cov = CoverageTrack(self.chrom, self.start, self.end)
        if self.params.lower > 0:
            self.nfr_cov.calculateCoverage(self.mat, 0, self.params.lower, self.params.window)
        else:
            self.nfr_cov.vals = np.zeros(self.length())

    def smoothSignal(self):
        self.smoothed = Track(self.chrom, self.start, self.end, "Smooth Signal")
        self.smoothed.assign_track(np.maximum(np.array(self.norm_signal.vals), 0))
        self.smoothed.smooth_track(6 * self.params.smooth_sd + 1, window="gaussian", sd=self.params.smooth_sd, mode="same", norm=True)

    def getOcc(self):
        """Occupancy tracks written by `nucleoatac occ` (NucleosomeCalling.py:284-293)."""
        base = self.params.occ_track[:-11]
        self.occ = Track(self.chrom, self.start, self.end, "Occupancy")
        self.occ.read_track(self.params.occ_track)
        self.occ_lower = Track(self.chrom, self.start, self.end, "Occupancy")
        self.occ_lower.read_track(base + "lower_bound.bedgraph.gz")
        self.occ_upper = Track(self.chrom, self.start, self.end, "Occupancy")
        self.occ_upper.read_track(base + "upper_bound.bedgraph.gz")

    def findAllNucs(self):
        p = self.params
        self.nuc_collection = {}
        combined = self.norm_signal.vals + self.smoothed.vals
        for i in call_peaks(combined, min_signal=0, sep=p.redundant_sep, boundary=p.nonredundant_sep // 2, order=p.redundant_sep // 2):
            nuc = Nucleosome(int(i) + self.start, self)
            if nuc.nuc_cov > p.min_reads:
                nuc.getLR(self)
                if nuc.lr > p.min_lr:
                    nuc.getZScore(self)
                    if nuc.z >= p.min_z:
                        nuc.getOcc(self)
                        self.nuc_collection[int(i)] = nuc
        self._finish_calls()

    def _finish_calls(self):
        self.sorted_nuc_keys = np.array(sorted(self.nuc_collection.keys()), dtype=np.int64)
        self.nonredundant = reduce_peaks(self.sorted_nuc_keys, [self.nuc_collection[x].z for x in self.sorted_nuc_keys],
                                         self.params.nonredundant_sep)
        self.redundant = np.setdiff1d(self.sorted_nuc_keys, self.nonredundant)

    def fuzz_jobs(self):
        """(left edges, jobs) of all calls of the chunk in key order (the clipping of the signal is order dependent)."""
        lefts, jobs = [], []
        for k in self.sorted_nuc_keys:
            left, job = self.nuc_collection[int(k)].fuzz_job(self)
            lefts.append(left)
            jobs.append(job)
        return lefts, jobs

    def set_fits(self, lefts, results):
        x = np.linspace(0, self.length() - 1, self.length())
        fit = np.zeros(self.length())
        for k, left, res in zip(self.sorted_nuc_keys, lefts, results):
            nuc = self.nuc_collection[int(k)]
            nuc.set_fuzz(left, res)
            fit += norm(x, nuc.fuzz ** 2, nuc.weight, nuc.fit_pos)
        self.fitted = Track(self.chrom, self.start, self.end, "Fitted Nucleosome Signal")
        self.fitted.assign_track(fit)

    def fit(self):
        from . import fuzz
        lefts, jobs = self.fuzz_jobs()
        self.set_fits(lefts, fuzz.fit_many(jobs))

    def makeInsertionTrack(self):
        """Fragment-end counts over the chunk (mat.getIns(), NucleosomeCalling.py:325-327) straight from the reads."""
        p = self.params
        pad = p.pad()
        half = (p.upper + (p.upper - 1) % 2) // 2
        s, e = self.start - pad + half, self.end + pad - half
        pos, tlen = fetch_reads(p.bam, self.chrom, s - p.upper, e + p.upper)
        # only fragments whose centre lies inside the matrix contribute in the reference
        l = pos.astype(np.int64) + (4 if p.atac else 0)
        i = np.abs(tlen.astype(np.int64)) - (8 if p.atac else 0)
        c = l + (i - 1) // 2
        keep = (c >= self.start - pad) & (c < self.end + pad)
        self.ins = InsertionTrack(self.chrom, s, e)
        self.ins.assign_track(default_engine().insertions(pos[keep], tlen[keep], s, e, 0, p.upper, p.atac))

    # ---- fused device path
    def _fill(self, out, pb, j, fit=True):
        a, b = int(pb.out_off[j]), int(pb.out_off[j + 1])

        def trk(cls, vals, *args):
            t = cls(self.chrom, self.start, self.end, *args)
            t.vals = vals
            return t

        self.nuc_cov = trk(CoverageTrack, out["nuc_cov"][a:b])
        self.nfr_cov = trk(CoverageTrack, out["nfr_cov"][a:b])
        self.bias = trk(BiasTrack, out["background"][a:b])
        self.nuc_signal = trk(SignalTrack, out["nuc_signal"][a:b])
        self.norm_signal = trk(NormSignalTrack, out["norm_signal"][a:b])
        self.smoothed = trk(Track, out["smoothed"][a:b], "Smooth Signal")
        if self.params.occ_track is not None:
            self.getOcc()
        co, n = int(out["cand_off"][j]), int(out["cand_count"][j])
        if n < 0:
            raise Exception("candidate capacity exceeded in " + self.asBed())
        self.nuc_collection = {}
        nonred = []
        for q in range(co, co + n):
            if out["cand_flag"][q] & FLAG_Z:
                nuc = Nucleosome(int(out["cand_pos"][q]), self)
                nuc.lr, nuc.z = out["cand_lr"][q], out["cand_z"][q]
                nuc.getOcc(self)
                key = nuc.start - self.start
                self.nuc_collection[key] = nuc
                if out["cand_flag"][q] & FLAG_NONREDUNDANT:
                    nonred.append(key)
        self.sorted_nuc_keys = np.array(sorted(self.nuc_collection.keys()), dtype=np.int64)
        self.nonredundant = np.array(nonred, dtype=np.int64)
        self.redundant = np.setdiff1d(self.sorted_nuc_keys, self.nonredundant)
        if fit:
            self.fit()

    def process(self, params, fit=True, insertions=False):
        """NucChunk.process (NucleosomeCalling.py:328-340) as one device pass (+ the host fuzziness fit).  The
        insertion track is not emitted by run_nuc (run_nuc.py:30-32), so it is only built on request."""
        process_chunks([self], params, fit=fit)
        if insertions:
            self.makeInsertionTrack()

    def removeData(self):
        for name in list(self.__dict__.keys()):
            delattr(self, name)


def process_chunks(nuc_chunks, params, fit=True):
    """Run a list of NucChunk objects through the device in one batch and fill their attributes."""
    eng = params.engine()
    pb = params.pack(nuc_chunks)
    out = eng.process_nuc(pb)
    for j, nc in enumerate(nuc_chunks):
        nc.params = params
        nc._fill(out, pb, j, fit=False)
    if fit:   # the fuzziness fits of the whole batch go to the worker pool in one map
        from . import fuzz
        per_chunk = [nc.fuzz_jobs() for nc in nuc_chunks]
        results = fuzz.fit_many([job for _, jobs in per_chunk for job in jobs])
        at = 0
        for nc, (lefts, jobs) in zip(nuc_chunks, per_chunk):
            nc.set_fits(lefts, results[at:at + len(jobs)])
            at += len(jobs)
    return out
