"""Nucleosome occupancy (nucleoatac/Occupancy.py:21-253): fragment-size mixture model, occupancy MLE tracks,
occupancy peaks.  `OccChunk.process` runs the whole chunk on the device (nb200_occ_run); `process_chunks`
does the same for a list of chunks in one batch, which is what run_occ feeds."""
import numpy as np

from . import seq as _seq
from .bias import PWM, InsertionBiasTrack
from .chunk import Chunk
from .chunkmat2d import BiasMat2D, FragmentMat2D
from .engine import PackedBatch, default_engine
from .fragments import fetch_reads_many
from .fragmentsizes import FragmentSizes
from .tracks import CoverageTrack, Track
from .utils import call_peaks, fmt12, read_chrom_sizes_from_fasta, smooth


class FragmentMixDistribution:
    """Insert-size distribution split into an NFR (gamma) and a nucleosomal part (Occupancy.py:21-87)."""

    def __init__(self, lower=0, upper=2000):
        self.lower, self.upper = lower, upper

    def getFragmentSizes(self, bamfile, chunklist=None):
        self.fragmentsizes = FragmentSizes(self.lower, self.upper)
        self.fragmentsizes.calculateSizes(bamfile, chunks=chunklist)

    def modelNFR(self, boundaries=(35, 115)):
        """Gamma fit of the sub-nucleosomal sizes (scipy brute + fmin, once per run; Occupancy.py:29-66)."""
        from scipy import optimize
        from scipy.special import gamma
        fs = self.fragmentsizes
        head = fs.get(self.lower, boundaries[1])
        b = int(np.where(head == max(head))[0][0]) + self.lower
        boundaries = (min(boundaries[0], b), boundaries[1])
        x = np.arange(boundaries[0], boundaries[1])
        y = fs.get(boundaries[0], boundaries[1])

        def gamma_fit(X, o, p):
            k, theta, a = p[0], p[1], p[2]
            xm = X - o
            res = np.zeros(len(xm))
            nz = xm >= 0 if k >= 1 else xm > 0
            res[nz] = a * xm[nz] ** (k - 1) * np.exp(-xm[nz] / theta) / (theta ** k * gamma(k))
            return res

        score = np.ones(boundaries[0] + 1) * float("inf")
        param = [0] * (boundaries[0] + 1)
        pranges = ((0.01, 10), (0.01, 150), (0.01, 1))
        # optimize.brute(f, pranges, finish=optimize.fmin) as the reference calls it = the 20 x 20 x 20 grid of the ranges
        # (end points included), its first minimum, then Nelder-Mead from there.  The grid's 8000 objective values are the
        # same element-by-element arithmetic evaluated on arrays instead of 8000 Python calls per offset (3 s of a run's
        # fixed cost); the polish is scipy's own.
        axes = [np.linspace(lo, hi, 20) for lo, hi in pranges]
        K, TH, A = (g.reshape(-1, 1) for g in np.meshgrid(*axes, indexing="ij"))
        for i in range(15, boundaries[0] + 1):
            f = lambda p: np.sum((gamma_fit(x, i, p) - y) ** 2)
            xm = (x - i).astype(np.float64)[None, :]
            nz = np.where(K >= 1, xm >= 0, xm > 0)
            with np.errstate(all="ignore"):
                vals = A * np.where(nz, xm, 1.0) ** (K - 1) * np.exp(-np.where(nz, xm, 1.0) / TH) / (TH ** K * gamma(K))
            J = np.array([np.sum((row - y) ** 2) for row in np.where(nz, vals, 0.0)])
            j = int(np.argmin(J))
            res = optimize.fmin(f, np.array([K[j, 0], TH[j, 0], A[j, 0]]), full_output=1, disp=0)
            score[i], param[i] = res[1], res[0]
        which = int(np.argmin(score))
        self.nfr_fit0 = FragmentSizes(self.lower, self.upper, vals=gamma_fit(np.arange(self.lower, self.upper), which, param[which]))
        nfr = np.concatenate((fs.get(self.lower, boundaries[1]), self.nfr_fit0.get(boundaries[1], self.upper)))
        nfr[nfr == 0] = min(nfr[nfr != 0]) * 0.01
        self.nfr_fit = FragmentSizes(self.lower, self.upper, vals=nfr)
        nuc = np.concatenate((np.zeros(boundaries[1] - self.lower),
                              fs.get(boundaries[1], self.upper) - self.nfr_fit.get(boundaries[1], self.upper)))
        nuc[nuc <= 0] = min(min(nfr) * 0.1, min(nuc[nuc > 0]) * 0.001)
        self.nuc_fit = FragmentSizes(self.lower, self.upper, vals=nuc)

    def plotFits(self, filename=None):
        """Plotting is outside the scoring path (matplotlib is optional); the fit table is written instead."""
        if filename:
            np.savetxt(filename[:-4] + ".txt" if filename.endswith(".eps") else filename,
                       np.vstack((self.fragmentsizes.get(), self.nuc_fit.get(), self.nfr_fit.get())))


class OccupancyCalcParams:
    def __init__(self, lower, upper, insert_dist, ci=0.9):
        from scipy import stats
        self.lower, self.upper = lower, upper
        nuc = np.asarray(insert_dist.nuc_fit.get(lower, upper), dtype=np.float64)
        nfr = np.asarray(insert_dist.nfr_fit.get(lower, upper), dtype=np.float64)
        self.nuc_probs = nuc / np.sum(nuc)
        self.nfr_probs = nfr / np.sum(nfr)
        self.alphas = np.linspace(0, 1, 101)
        self.l = len(self.alphas)
        self.cutoff = stats.chi2.ppf(ci, 1)


def calculateOccupancy(inserts, bias, params):
    """Occupancy MLE on the alpha grid with likelihood-ratio bounds (Occupancy.py:104-120) -> (occ, lower, upper)."""
    eng = default_engine()
    eng.set_occ_model(params.nuc_probs, params.nfr_probs, params.alphas, params.cutoff)
    return eng.calculate_occupancy(np.asarray(inserts, dtype=np.float64), np.asarray(bias, dtype=np.float64))


class OccupancyTrack(Track):
    def __init__(self, chrom, start, end):
        Track.__init__(self, chrom, start, end, "occupancy")

    def calculateOccupancyMLE(self, mat, bias_mat, params):
        """Window-by-window MLE from dense matrices (Occupancy.py:128-146): the object-API path; OccChunk.process
        uses the fused device kernels instead."""
        offset = self.start - mat.start
        if offset < params.flank:
            raise Exception("For calculateOccupancyMLE, mat does not have sufficient flanking regions")
        n = self.end - self.start
        self.vals = np.ones(n) * np.nan
        self.lower_bound = np.ones(n) * np.nan
        self.upper_bound = np.ones(n) * np.nan
        for i in range(params.halfstep, n, params.step):
            ins = np.sum(mat.get(lower=0, upper=params.upper, start=self.start + i - params.flank,
                                 end=self.start + i + params.flank + 1), axis=1)
            bias = np.sum(bias_mat.get(lower=0, upper=params.upper, start=self.start + i - params.flank,
                                       end=self.start + i + params.flank + 1), axis=1)
            if sum(ins) > 0:
                lo, hi = i - params.halfstep, min(i + params.halfstep + 1, n)
                self.vals[lo:hi], self.lower_bound[lo:hi], self.upper_bound[lo:hi] = calculateOccupancy(
                    ins, bias, params.occ_calc_params)

    def makeSmoothed(self, window_len=121, sd=20):
        self.smoothed_vals = smooth(self.vals, window_len, window="gaussian", sd=sd, mode="same", norm=True)
        self.smoothed_lower = smooth(self.lower_bound, window_len, window="gaussian", sd=sd, mode="same", norm=True)
        self.smoothed_upper = smooth(self.upper_bound, window_len, window="gaussian", sd=sd, mode="same", norm=True)


class OccPeak(Chunk):
    def __init__(self, pos, chunk):
        self.chrom, self.start, self.end, self.strand = chunk.chrom, pos, pos + 1, "*"
        i = pos - chunk.occ.start
        self.occ = chunk.occ.smoothed_vals[i]
        self.occ_lower = chunk.occ.smoothed_lower[i]
        self.occ_upper = chunk.occ.smoothed_upper[i]
        self.reads = chunk.cov.get(pos=pos)

    def asBed(self):
        return "\t".join([self.chrom, str(self.start), str(self.end), fmt12(self.occ), fmt12(self.occ_lower),
                          fmt12(self.occ_upper), fmt12(self.reads)])

    def write(self, handle):
        handle.write(self.asBed() + "\n")


class OccupancyParameters:
    """Parameters of the occupancy run (Occupancy.py:175-193) + the device engine configured for them."""

    def __init__(self, insert_dist, upper, fasta, pwm, sep=120, min_occ=0.1, flank=60, out=None, bam=None, ci=0.9,
                 step=5, device=0):
        self.sep = sep
        self.fasta = fasta
        self.chrs = read_chrom_sizes_from_fasta(fasta) if fasta is not None else None
        self.pwm = PWM.open(pwm) if fasta is not None else None
        self.window = flank * 2 + 1
        self.min_occ, self.flank, self.bam, self.upper = min_occ, flank, bam, upper
        self.occ_calc_params = OccupancyCalcParams(0, upper, insert_dist, ci=ci)
        if step % 2 == 0:
            step -= 1
        self.step = step
        self.halfstep = (self.step - 1) // 2
        self.device = device

    def engine(self):
        eng = default_engine(self.device)
        if getattr(eng, "_occ_owner", None) is not self:
            cp = self.occ_calc_params
            if self.pwm is not None:
                eng.set_pwm(self.pwm.mat, self.pwm.up, self.pwm.down, self.pwm.nucleotides)
            eng.set_occ_model(cp.nuc_probs, cp.nfr_probs, cp.alphas, cp.cutoff)
            eng.configure_occ(upper=self.upper, flank=self.flank, step=self.step, sep=self.sep, min_occ=self.min_occ,
                              use_bias=self.fasta is not None)
            eng._occ_owner = self
        return eng

    def pack(self, chunks):
        """Host-side gather of a batch: reads from the BAM, sequence from the FASTA (the only host work per chunk)."""
        items = []
        reads = fetch_reads_many(self.bam, [(c.chrom, c.start - self.flank - self.upper, c.end + self.flank + self.upper) for c in chunks])
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


class OccChunk(Chunk):
    def __init__(self, chunk):
        self.start, self.end, self.chrom = chunk.start, chunk.end, chunk.chrom
        self.peaks = {}
        self.nfrs = []

    # ---- step-by-step object API (dense matrices, primitives); kept for drop-in use and the reference's tests
    def getFragmentMat(self):
        self.mat = FragmentMat2D(self.chrom, self.start - self.params.flank, self.end + self.params.flank, 0, self.params.upper)
        self.mat.makeFragmentMat(self.params.bam)

    def makeBiasMat(self):
        p = self.params
        self.bias_mat = BiasMat2D(self.chrom, self.start - p.flank, self.end + p.flank, 0, p.upper)
        if p.fasta is not None:
            bt = InsertionBiasTrack(self.chrom, self.start - p.window - p.upper // 2, self.end + p.window + p.upper // 2 + 1, log=True)
            bt.computeBias(p.fasta, p.chrs, p.pwm)
            self.bias_mat.makeBiasMat(bt)

    def calculateOcc(self):
        self.occ = OccupancyTrack(self.chrom, self.start, self.end)
        self.occ.calculateOccupancyMLE(self.mat, self.bias_mat, self.params)
        self.occ.makeSmoothed(window_len=self.params.window, sd=self.params.flank / 3.0)

    def getCov(self):
        self.cov = CoverageTrack(self.chrom, self.start, self.end)
        self.cov.calculateCoverage(self.mat, 0, self.params.upper, self.params.window)

    def callPeaks(self):
        for peak in call_peaks(self.occ.smoothed_vals, sep=self.params.sep, min_signal=self.params.min_occ):
            tmp = OccPeak(int(peak) + self.start, self)
            if tmp.occ_lower > self.params.min_occ and tmp.reads > 0:
                self.peaks[int(peak)] = tmp

    def getNucDist(self):
        """Sum over peaks of the normalised insert-size histogram of the +-flank window (Occupancy.py:232-240)."""
        if getattr(self, "_nuc_dist", None) is not None:
            return self._nuc_dist
        nuc_dist = np.zeros(self.params.upper)
        for peak in self.peaks:
            sub = np.sum(self.mat.get(start=self.peaks[peak].start - self.params.flank,
                                      end=self.peaks[peak].start + 1 + self.params.flank), axis=1)
            nuc_dist += sub / float(sum(sub))
        return nuc_dist

    # ---- fused device path
    def _fill(self, out, pb, j):
        a, b = int(pb.out_off[j]), int(pb.out_off[j + 1])
        self.occ = OccupancyTrack(self.chrom, self.start, self.end)
        self.occ.vals, self.occ.lower_bound, self.occ.upper_bound = out["vals"][a:b], out["lower_bound"][a:b], out["upper_bound"][a:b]
        self.occ.smoothed_vals = out["smoothed_vals"][a:b]
        self.occ.smoothed_lower = out["smoothed_lower"][a:b]
        self.occ.smoothed_upper = out["smoothed_upper"][a:b]
        self.cov = CoverageTrack(self.chrom, self.start, self.end)
        self.cov.vals = out["cov"][a:b]
        po, n = int(out["peak_off"][j]), int(out["peak_count"][j])
        if n < 0:
            raise Exception("occupancy peak capacity exceeded in " + self.asBed())
        self.peaks = {}
        for q in range(po, po + n):
            self.peaks[int(out["peak_pos"][q]) - self.start] = OccPeak(int(out["peak_pos"][q]), self)
        self._nuc_dist = np.array(out["nuc_dist"][j])

    def process(self, params):
        """OccChunk.process (Occupancy.py:241-248) as one device pass."""
        process_chunks([self], params)

    def removeData(self):
        for name in list(self.__dict__.keys()):
            delattr(self, name)


def process_chunks(occ_chunks, params):
    """Run a list of OccChunk objects through the device in one batch and fill their attributes."""
    eng = params.engine()
    pb = params.pack(occ_chunks)
    out = eng.process_occ(pb, raw=True)
    for j, oc in enumerate(occ_chunks):
        oc.params = params
        oc._fill(out, pb, j)
    return out
