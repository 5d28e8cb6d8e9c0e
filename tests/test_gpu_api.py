"""GPU tests of the reference-facing Python surface: OccChunk / NucChunk / FragmentMat2D / BiasMat2D / tracks /
calculateCov with the reference's names and attributes, and the `nucleoatac occ|nuc` command line on real files
(BAM + FASTA + BED written by tests/synthfiles.py), all checked against the CPU oracle."""
import gzip
import os

import numpy as np
import pytest

from oracle import refalgo as ra, refnuc, refocc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    from tests.synthfiles import make_files
    return make_files(str(tmp_path_factory.mktemp("synth")))


def _bias(files, span, wl):
    g = files["genome"]
    seq = g[span[0] - 10:span[1] + 10].tobytes().decode()
    return ra.log_bias_track(seq, wl.pwm, wl.nucleotides)


class _Dist:  # stands in for FragmentMixDistribution after modelNFR
    def __init__(self, nuc, nfr):
        from nucleoatac_b200.fragmentsizes import FragmentSizes
        self.nuc_fit = FragmentSizes(0, len(nuc), vals=nuc)
        self.nfr_fit = FragmentSizes(0, len(nfr), vals=nfr)


def test_occchunk_object_api(files):
    from nucleoatac_b200.chunk import Chunk
    from nucleoatac_b200.Occupancy import OccChunk, OccupancyParameters
    wl = files["wl"]
    params = OccupancyParameters(_Dist(wl.nuc_probs, wl.nfr_probs), wl.upper, files["fasta"], "Human", bam=files["bam"])
    oparams = refocc.OccParams(wl.nuc_probs, wl.nfr_probs, upper=wl.upper)
    s, e, pos, tlen, seq, s0 = files["chunks"][0]
    oc = OccChunk(Chunk("chrS", s, e))
    oc.process(params)
    span = refocc.occ_bias_track_span(s, e, oparams)
    r = refocc.process_occ_chunk(pos, tlen, s, e, oparams, bias_track=_bias(files, span, wl), bias_track_start=span[0])
    assert np.array_equal(oc.occ.vals, r["vals"], equal_nan=True)
    np.testing.assert_allclose(oc.occ.smoothed_vals, r["smoothed_vals"], rtol=1e-9, equal_nan=True)
    np.testing.assert_allclose(oc.occ.smoothed_lower, r["smoothed_lower"], rtol=1e-9, equal_nan=True)
    assert np.array_equal(oc.cov.vals, r["cov"])
    assert sorted(oc.peaks.keys()) == [p[0] - s for p in r["peaks"]]
    lines = [oc.peaks[k].asBed() for k in sorted(oc.peaks)]
    assert [l.split("\t")[:3] for l in lines] == [["chrS", str(p[0]), str(p[0] + 1)] for p in r["peaks"]]
    np.testing.assert_allclose(oc.getNucDist(), r["nuc_dist"], rtol=1e-9, atol=1e-15)
    # the step-by-step object API (dense matrices through the primitives) agrees with the fused path on a sub-region
    sub = OccChunk(Chunk("chrS", s + 2000, s + 2600))
    sub.params = params
    sub.getFragmentMat()
    sub.makeBiasMat()
    sub.calculateOcc()
    sub.getCov()
    sub.callPeaks()
    fused = OccChunk(Chunk("chrS", s + 2000, s + 2600))
    fused.process(params)
    assert sub.mat.mat.shape == (wl.upper, 600 + 120) and sub.mat.mat.sum() > 0
    assert np.array_equal(sub.occ.vals, fused.occ.vals, equal_nan=True)
    np.testing.assert_allclose(sub.occ.smoothed_vals, fused.occ.smoothed_vals, rtol=1e-9, equal_nan=True)
    assert np.array_equal(sub.cov.vals, fused.cov.vals) and sorted(sub.peaks) == sorted(fused.peaks)
    oc.removeData()
    assert not oc.__dict__


def test_nucchunk_object_api(files):
    from nucleoatac_b200.chunk import Chunk
    from nucleoatac_b200.fragmentsizes import FragmentSizes
    from nucleoatac_b200.NucleosomeCalling import NucChunk, NucParameters
    from nucleoatac_b200.VMat import VMat
    wl = files["wl"]
    vmat = VMat.open(files["vmat"])
    fs = FragmentSizes.open(files["sizes"])
    params = NucParameters(vmat, fs, files["bam"], files["fasta"], "Human", sd=10, xcor_mode=1)
    nparams = refnuc.NucParams((vmat.mat, vmat.lower, vmat.upper), fs.get(0, vmat.upper), sd=10)
    s, e, pos, tlen, seq, s0 = files["chunks"][0]
    nc = NucChunk(Chunk("chrS", s, e))
    nc.process(params, insertions=True)
    _, _, span = refnuc.nuc_geometry(s, e, nparams)
    r = refnuc.process_nuc_chunk(pos, tlen, s, e, nparams, bias_track=_bias(files, span, wl), bias_track_start=span[0],
                                 fit=True, want_ins=True)
    scale = float(np.abs(r["nuc_signal"]).max())
    for attr, key in (("nuc_signal", "nuc_signal"), ("bias", "bias"), ("norm_signal", "norm_signal"), ("smoothed", "smoothed")):
        np.testing.assert_allclose(getattr(nc, attr).vals, r[key], rtol=1e-8, atol=1e-9 * scale)
    assert np.array_equal(nc.nuc_cov.vals, r["nuc_cov"]) and np.array_equal(nc.nfr_cov.vals, r["nfr_cov"])
    assert list(nc.sorted_nuc_keys) == list(r["sorted_nuc_keys"]) and len(nc.sorted_nuc_keys) > 0
    assert sorted(nc.nonredundant) == sorted(r["nonredundant"]) and sorted(nc.redundant) == sorted(r["redundant"])
    for k in nc.sorted_nuc_keys:
        mine, ref = nc.nuc_collection[int(k)], r["nuc_collection"][int(k)]
        for f in ("z", "lr", "norm_signal", "nuc_signal", "nuc_cov", "nfr_cov"):
            np.testing.assert_allclose(getattr(mine, f), ref[f], rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(mine.fuzz, ref["fuzz"], rtol=1e-4)  # host L-BFGS-B on a track equal to 1e-9
        assert len(mine.asBed().split("\t")) == 13
    ins, i0, i1 = r["ins"]
    assert (nc.ins.start, nc.ins.end) == (i0, i1) and np.array_equal(nc.ins.vals, ins)


def test_primitive_classes(files):
    from nucleoatac_b200.bias import PWM, InsertionBiasTrack
    from nucleoatac_b200.chunkmat2d import BiasMat2D, FragmentMat2D
    from nucleoatac_b200.multinomial_cov import calculateCov
    from nucleoatac_b200.tracks import CoverageTrack, InsertionTrack
    from nucleoatac_b200.utils import call_peaks, read_chrom_sizes_from_fasta, reduce_peaks, smooth
    wl = files["wl"]
    s, e, pos, tlen, seq, s0 = files["chunks"][1]
    a, b = s + 1000, s + 1800
    m = FragmentMat2D("chrS", a, b, 0, 251)
    m.makeFragmentMat(files["bam"])
    assert np.array_equal(m.mat, ra.make_fragment_mat(pos, tlen, a, b, 0, 251))
    assert np.array_equal(m.get(lower=100, upper=102, start=a + 5, end=a + 7), m.mat[100:102, 5:7])
    it = InsertionTrack("chrS", a, b)
    it.calculateInsertions(files["bam"])
    assert np.array_equal(it.vals, ra.get_insertions(pos, tlen, a, b, 0, 2000))
    g = m.getIns()
    assert (g.start, g.end) == (a + 125, b - 125)
    chrs = read_chrom_sizes_from_fasta(files["fasta"])
    bt = InsertionBiasTrack("chrS", a - 200, b + 200)
    bt.computeBias(files["fasta"], chrs, PWM.open("Human"))
    np.testing.assert_allclose(bt.vals, _bias(files, (a - 200, b + 200), wl), rtol=1e-12, atol=1e-13)
    bm = BiasMat2D("chrS", a, b, 0, 251)
    bm.makeBiasMat(bt)
    np.testing.assert_allclose(bm.mat, ra.make_bias_mat(bt.get(a - 125, b + 125), 0, 251), rtol=1e-13)
    cov = CoverageTrack("chrS", a + 60, b - 60)
    cov.calculateCoverage(m, 0, 251, 121)
    assert np.array_equal(cov.vals, ra.calculate_coverage(m.mat, a, 0, a + 60, 0, 251, 121))
    x = np.sin(np.arange(400) / 7.0)
    np.testing.assert_allclose(smooth(x, 31, window="gaussian", sd=5, mode="same"), ra.smooth(x, 31, "gaussian", 5, "same"), rtol=1e-12, atol=1e-15)
    assert list(call_peaks(x.copy(), min_signal=0.5, sep=20)) == list(ra.call_peaks(x.copy(), min_signal=0.5, sep=20))
    pk = np.array([10, 40, 55, 200]); sc = [1.0, 3.0, 2.0, 0.5]
    assert list(reduce_peaks(pk, sc, 30)) == list(ra.reduce_peaks(pk, sc, 30))
    p = np.random.RandomState(0).dirichlet(np.ones(500)); v = np.random.RandomState(1).rand(500)
    from oracle import mcov
    assert abs(calculateCov(p, v, 35.7) - mcov.calculate_cov(p, v, 35)) < 1e-12
    with pytest.raises(ValueError):
        calculateCov(p, v[:-1], 3)


def _read(path):
    with gzip.open(path, "rt") as fh:
        return [l.rstrip("\n").split("\t") for l in fh]


def test_cli_nuc_and_occ(files, tmp_path):
    """`nucleoatac nuc` / `occ` end to end on files, compared with the oracle's text output (Python-2 number format)."""
    from nucleoatac_b200.cli import nucleoatac_main
    from nucleoatac_b200.Occupancy import FragmentMixDistribution
    from nucleoatac_b200.fragmentsizes import FragmentSizes
    wl = files["wl"]
    out = str(tmp_path / "run")
    assert nucleoatac_main(["nuc", "--bed", files["bed"], "--bam", files["bam"], "--fasta", files["fasta"], "--vmat", files["vmat"],
                            "--sizes", files["sizes"], "--out", out, "--write_all", "--xcor_mode", "1"]) == 0
    fs = FragmentSizes.open(files["sizes"])
    nparams = refnuc.NucParams((wl.vmat, wl.v_lower, wl.v_upper), fs.get(0, wl.v_upper), sd=10)
    order = sorted(range(len(files["chunks"])), key=lambda i: files["chunks"][i][0])
    exp_sig, exp_pos = "", []
    for i in order:
        s, e, pos, tlen, seq, s0 = files["chunks"][i]
        _, _, span = refnuc.nuc_geometry(s, e, nparams)
        r = refnuc.process_nuc_chunk(pos, tlen, s, e, nparams, bias_track=_bias(files, span, wl), bias_track_start=span[0], fit=True)
        exp_sig += ra.write_track("chrS", s, e, r["norm_signal"])
        exp_pos += [refnuc.nuc_bed("chrS", r["nuc_collection"][int(k)]) for k in sorted(r["nonredundant"])]
    got = _read(out + ".nucleoatac_signal.bedgraph.gz")
    exp = [l.split("\t") for l in exp_sig.splitlines()]
    assert len(got) == len(exp) and [g[:3] for g in got] == [x[:3] for x in exp]
    np.testing.assert_allclose([float(g[3]) for g in got], [float(x[3]) for x in exp], rtol=1e-7, atol=1e-9)
    gp = _read(out + ".nucpos.bed.gz")
    assert [g[:3] for g in gp] == [l.split("\t")[:3] for l in exp_pos] and len(gp) > 0
    for g, l in zip(gp, exp_pos):
        x = l.split("\t")
        np.testing.assert_allclose([float(v) for v in g[3:12]], [float(v) for v in x[3:12]], rtol=1e-7, atol=1e-9, equal_nan=True)
    for n in ("nucpos.redundant.bed.gz", "nucleoatac_signal.smooth.bedgraph.gz", "nucleoatac_background.bedgraph.gz", "nucleoatac_raw.bedgraph.gz"):
        assert os.path.exists(out + "." + n)
    # occ: the gamma NFR model is host scipy (once per run); the oracle scores with the same fits
    assert nucleoatac_main(["occ", "--bed", files["bed"], "--bam", files["bam"], "--fasta", files["fasta"], "--sizes", files["sizes"],
                            "--out", out]) == 0
    fd = FragmentMixDistribution(0, upper=251)
    fd.fragmentsizes = FragmentSizes(0, 251, vals=fs.get(0, 251))
    fd.modelNFR()
    oparams = refocc.OccParams(fd.nuc_fit.get(0, 251), fd.nfr_fit.get(0, 251), upper=251)
    exp_occ, nd = "", np.zeros(251)
    for i in order:
        s, e, pos, tlen, seq, s0 = files["chunks"][i]
        span = refocc.occ_bias_track_span(s, e, oparams)
        r = refocc.process_occ_chunk(pos, tlen, s, e, oparams, bias_track=_bias(files, span, wl), bias_track_start=span[0])
        exp_occ += ra.write_track("chrS", s, e, r["smoothed_vals"])
        nd += r["nuc_dist"]
    got = _read(out + ".occ.bedgraph.gz")
    exp = [l.split("\t") for l in exp_occ.splitlines()]
    assert [g[:3] for g in got] == [x[:3] for x in exp]
    np.testing.assert_allclose([float(g[3]) for g in got], [float(x[3]) for x in exp], rtol=1e-9)
    np.testing.assert_allclose(FragmentSizes.open(out + ".nuc_dist.txt").get(), nd, rtol=1e-9, atol=1e-12)
    assert os.path.exists(out + ".occpeaks.bed.gz") and os.path.exists(out + ".fragmentsizes.txt")


def test_cli_nfr(files, tmp_path):
    """`nucleoatac nfr` on files (after occ -> nuc -> merge wrote its inputs), NFRCalling.py:51-111 / run_nfr.py:72-127:
    insertion track bit-exact, NFR rows equal to the oracle's on the same tracks and calls."""
    from nucleoatac_b200.cli import nucleoatac_main
    from nucleoatac_b200 import hostio
    from oracle import refnfr
    wl = files["wl"]
    out = str(tmp_path / "run")
    base = ["--bed", files["bed"], "--bam", files["bam"], "--fasta", files["fasta"], "--out", out]
    assert nucleoatac_main(["occ"] + base + ["--sizes", files["sizes"]]) == 0
    assert nucleoatac_main(["nuc"] + base + ["--vmat", files["vmat"], "--sizes", files["sizes"], "--occ_track", out + ".occ.bedgraph.gz"]) == 0
    assert nucleoatac_main(["merge", "--occpeaks", out + ".occpeaks.bed.gz", "--nucpos", out + ".nucpos.bed.gz", "--out", out]) == 0
    # loose thresholds so that the synthetic data yields regions
    assert nucleoatac_main(["nfr", "--bed", files["bed"], "--occ_track", out + ".occ.bedgraph.gz", "--calls", out + ".nucmap_combined.bed.gz",
                            "--bam", files["bam"], "--fasta", files["fasta"], "--out", out, "--max_occ", "0.9", "--max_occ_upper", "1.1"]) == 0
    got_nfr = _read(out + ".nfrpos.bed.gz")
    got_ins = _read(out + ".ins.bedgraph.gz")
    assert os.path.exists(out + ".nfrpos.bed.gz.tbi") and os.path.exists(out + ".ins.bedgraph.gz.tbi")
    occ_rows, upp_rows = _read(out + ".occ.bedgraph.gz"), _read(out + ".occ.upper_bound.bedgraph.gz")
    calls = _read(out + ".nucmap_combined.bed.gz")

    def region(rows, s, e):
        v = np.full(e - s, np.nan)
        for r in rows:
            a, b = int(r[1]), int(r[2])
            if b > s and a < e:
                v[max(a - s, 0):min(b - s, e - s)] = float(r[3])
        return v
    exp_nfr, exp_ins = [], ""
    for (s, e, pos, tlen, seq, s0) in sorted(files["chunks"], key=lambda c: c[0]):
        s, e = s + 60, e - 60  # run_nfr reads the BED without slop (run_nfr.py:84-91)
        lb = _bias(files, (s, e), wl)
        dy = [int(r[1]) for r in calls if int(r[1]) < e and int(r[2]) > s]
        recs, ins = refnfr.process_nfr_chunk(pos, tlen, s, e, dy, region(occ_rows, s, e), region(upp_rows, s, e), lb,
                                             max_occ=0.9, max_occ_upper=1.1)
        exp_nfr += [refnfr.nfr_bed("chrS", r).split("\t") for r in recs]
        exp_ins += ra.write_track("chrS", s, e, ins)
    assert got_ins == [l.split("\t") for l in exp_ins.splitlines()]
    assert len(got_nfr) == len(exp_nfr) and len(got_nfr) > 0
    assert [g[:3] for g in got_nfr] == [x[:3] for x in exp_nfr]
    np.testing.assert_allclose([[float(v) for v in g[3:7]] for g in got_nfr], [[float(v) for v in x[3:7]] for x in exp_nfr], rtol=1e-9)
