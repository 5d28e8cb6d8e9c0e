"""GPU parity of the pyatac tools either side of the scoring path (vplot / cov / ins / bias / sizes) against
the oracle's literal restatement (oracle/refpyatac.py): primitives through the C-ABI on seeded reads, then the tools end
to end on files."""
import gzip
import os

import numpy as np
import pytest

from oracle import hostio as ohio, refalgo as ra, refpyatac as rp

pytestmark = pytest.mark.gpu


def _reads(rng, n, lo, hi, max_size=400):
    size = rng.integers(0, max_size, n)
    left = rng.integers(lo, hi, n)
    # stored BAM-like: pos = l - 4, tlen = size + 8 (half of them negative: abs() is taken, fragments.pyx:30)
    sign = np.where(rng.random(n) < 0.5, 1, -1)
    return (left - 4).astype(np.int32), (sign * (size + 8)).astype(np.int32)


@pytest.mark.parametrize("lower,upper,flank,atac", [(0, 250, 250, True), (30, 121, 40, True), (0, 60, 7, False)])
def test_vplot_primitive(lower, upper, flank, atac):
    from nucleoatac_b200.engine import default_engine
    eng = default_engine()
    rng = np.random.default_rng(7)
    sites = [(5000 + 700 * k + int(rng.integers(0, 50)), int(rng.integers(1, 40)), "+-"[k % 2]) for k in range(24)]
    centers, flips, off, ps, ts, exp, exp_scaled = [], [], [0], [], [], 0.0, 0.0
    for s0, length, strand in sites:
        pos, tlen = _reads(rng, 600, s0 - 700, s0 + 700)
        s, e = rp.center(s0, s0 + length, strand)
        centers.append(s)
        flips.append(int(strand == "-"))
        ps.append(pos)
        ts.append(tlen)
        off.append(off[-1] + len(pos))
        exp = exp + rp.vplot_site(pos, tlen, s0, s0 + length, strand, flank, lower, upper, atac)
        exp_scaled = exp_scaled + rp.vplot_site(pos, tlen, s0, s0 + length, strand, flank, lower, upper, atac, scale=True)
    got = eng.vplot(centers, flips, off, np.concatenate(ps), np.concatenate(ts), flank, lower, upper, atac, False)
    assert got.shape == (upper - lower, 2 * flank + 1) and got.sum() > 0
    np.testing.assert_array_equal(got, exp)  # integer counts: bit-exact
    got = eng.vplot(centers, flips, off, np.concatenate(ps), np.concatenate(ts), flank, lower, upper, atac, True)
    np.testing.assert_allclose(got, exp_scaled, rtol=1e-12, atol=1e-15)  # the reference sums float64 terms site by site
    # scaled sums are accumulated as 128-bit fixed-point integers: independent of the order of the atomics, so repeated runs
    # (and the sites in another order) give bit-identical plots
    for trial in range(3):
        perm = np.random.default_rng(trial).permutation(len(centers))
        off_p = np.concatenate(([0], np.cumsum([off[k + 1] - off[k] for k in perm])))
        again = eng.vplot([centers[k] for k in perm], [flips[k] for k in perm], off_p, np.concatenate([ps[k] for k in perm]),
                          np.concatenate([ts[k] for k in perm]), flank, lower, upper, atac, True)
        assert np.array_equal(again, got, equal_nan=True), trial   # (a site without reads turns the whole plot NaN, as in the reference)
    # a site without fragments under --scale: 0/0 in every cell of its matrix, the sum is NaN everywhere (make_vplot.py:34-35)
    off2 = off + [off[-1]]
    got = eng.vplot(centers + [10 ** 6], flips + [0], off2, np.concatenate(ps), np.concatenate(ts), flank, lower, upper, atac, True)
    assert np.isnan(got).all()
    got = eng.vplot(centers + [10 ** 6], flips + [0], off2, np.concatenate(ps), np.concatenate(ts), flank, lower, upper, atac, False)
    np.testing.assert_array_equal(got, exp)
    # no sites at all
    z = eng.vplot([], [], [0], np.zeros(0, np.int32), np.zeros(0, np.int32), flank, lower, upper, atac, False)
    assert z.shape == got.shape and not z.any()


@pytest.mark.parametrize("window,lower,upper", [(121, 0, 2000), (10, 0, 300), (1, 50, 200), (75, 0, 100)])
def test_coverage_primitive(window, lower, upper):
    from nucleoatac_b200.engine import default_engine
    eng = default_engine()
    rng = np.random.default_rng(11)
    start, end = 20000, 21537
    pos, tlen = _reads(rng, 5000, start - 600, end + 300, max_size=500)
    got = eng.coverage(pos, tlen, start, end, lower, upper, window, True)
    exp = rp.cov_chunk(pos, tlen, start, end, lower, upper, window, float(window), True)  # scale = window -> raw counts
    assert len(got) == end - start
    np.testing.assert_array_equal(got, exp)
    empty = eng.coverage(np.zeros(0, np.int32), np.zeros(0, np.int32), start, end, lower, upper, window, True)
    assert not empty.any()


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    from tests.synthfiles import make_files
    return make_files(str(tmp_path_factory.mktemp("pyatac")), ks=(2, 0, 5))


def _read(path):
    with gzip.open(path, "rt") as fh:
        return [l.rstrip("\n").split("\t") for l in fh]


def _check_bedgraph(path, expected_text, rtol=1e-12):
    got, exp = _read(path), [l.split("\t") for l in expected_text.splitlines()]
    assert len(got) == len(exp) and len(got) > 0
    assert [g[:3] for g in got] == [x[:3] for x in exp]
    np.testing.assert_allclose([float(g[3]) for g in got], [float(x[3]) for x in exp], rtol=rtol, atol=0)
    assert os.path.exists(path + ".tbi")


def test_tools_on_files(files, tmp_path):
    from nucleoatac_b200.pyatac_tools import pyatac_main
    from nucleoatac_b200.fragmentsizes import FragmentSizes
    from nucleoatac_b200.VMat import VMat
    _, frags = ohio.read_bam_fragments(files["bam"])
    pos, tlen = frags["chrS"]
    regions = [(int(f[1]), int(f[2])) for f in (l.split("\t") for l in open(files["bed"]))]
    out = str(tmp_path / "t")
    # ---- vplot: sites = the BED regions, alternating strands through a strand column
    sites = str(tmp_path / "sites.bed")
    with open(sites, "w") as fh:
        k = 0
        for s, e in regions:
            for c in range(s + 500, e - 500, 977):
                fh.write("chrS\t%d\t%d\tsite%d\t0\t%s\n" % (c, c + 1 + k % 3, k, "+-"[k % 2]))
                k += 1
    assert pyatac_main(["vplot", "--bed", sites, "--bam", files["bam"], "--out", out, "--strand", "6", "--flank", "120",
                        "--upper", "300"]) == 0
    exp = 0.0
    for l in open(sites):
        f = l.split("\t")
        exp = exp + rp.vplot_site(pos, tlen, int(f[1]), int(f[2]), f[5].strip(), 120, 0, 300)
    vm = VMat.open(out + ".VMat")
    assert vm.lower == 0 and vm.upper == 300 and vm.mat.shape == (300, 241) and vm.mat.sum() > 100
    np.testing.assert_array_equal(vm.mat, exp)
    # ---- cov (default window 121, scale 10) and ins (raw and gaussian-smoothed)
    assert pyatac_main(["cov", "--bam", files["bam"], "--bed", files["bed"], "--out", out, "--upper", "500"]) == 0
    txt = "".join(ra.write_track("chrS", s, e, rp.cov_chunk(pos, tlen, s, e, 0, 500, 121, 10.0)) for s, e in regions)
    _check_bedgraph(out + ".cov.bedgraph.gz", txt)
    assert pyatac_main(["ins", "--bam", files["bam"], "--bed", files["bed"], "--out", out, "--upper", "500"]) == 0
    txt = "".join(ra.write_track("chrS", s, e, rp.ins_chunk(pos, tlen, s, e, 0, 500)[1]) for s, e in regions)
    _check_bedgraph(out + ".ins.bedgraph.gz", txt)
    assert pyatac_main(["ins", "--bam", files["bam"], "--bed", files["bed"], "--out", out + "s", "--upper", "500", "--smooth", "21"]) == 0
    txt = "".join(ra.write_track("chrS", s, e, rp.ins_chunk(pos, tlen, s, e, 0, 500, smooth=21)[1]) for s, e in regions)
    _check_bedgraph(out + "s.ins.bedgraph.gz", txt, rtol=1e-9)
    # ---- bias: log PWM score of the BED regions
    assert pyatac_main(["bias", "--fasta", files["fasta"], "--bed", files["bed"], "--out", out]) == 0
    wl, genome = files["wl"], files["genome"].tobytes().decode()
    txt = "".join(ra.write_track("chrS", s, e, ra.log_bias_track(genome[s - 10:e + 10], wl.pwm, wl.nucleotides)) for s, e in regions)
    _check_bedgraph(out + ".Scores.bedgraph.gz", txt, rtol=1e-9)
    # ---- sizes inside the BED regions
    assert pyatac_main(["sizes", "--bam", files["bam"], "--bed", files["bed"], "--out", out, "--upper", "400", "--no_plot"]) == 0
    counts = ra.fragment_size_counts(pos, tlen, [(s, e) for s, e in regions], 0, 400)
    np.testing.assert_allclose(FragmentSizes.open(out + ".fragmentsizes.txt").get(), ra.normalize_sizes(counts), rtol=1e-11)
