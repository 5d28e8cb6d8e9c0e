"""oracle/ against vectors made by RUNNING THE REFERENCE'S OWN CODE (tests/golden/make_golden_pyref.py: OccChunk.process,
NucChunk.process and ChunkMat2D.get(flip=True) loaded from /root/reference through a Python-2 compatibility loader) on
chunks of the synthetic workload -- inputs other than the shipped example.  Grid values, peaks, calls and everything
integer-valued must be equal; floating-point tracks are the same numpy operations and come out equal to the last bit
(held to 1e-12); the fuzziness fit is scipy's optimiser in both (1e-6)."""
import os

import numpy as np
import pytest

from oracle import refalgo as ra, refnuc, refocc, refpyatac

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pyref_synth.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _case(gold, ci):
    from nucleoatac_b200 import synth
    k, length, density = gold["cases"][ci]
    margin = int(gold["seq_margin"])
    s, e, pos, tlen, seq, s0 = synth.make_chunk(int(k), length=int(length), density=float(density), seq_margin=margin)
    return synth.Workload(251, 251), s, e, pos, tlen, bytes(seq).decode(), s0


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_occ_chunk_equals_reference_run(gold, ci):
    """OccChunk.process + getNucDist (nucleoatac/Occupancy.py:195-253) as the reference itself computed them."""
    wl, s, e, pos, tlen, sq, s0 = _case(gold, ci)
    op = refocc.OccParams(wl.nuc_probs, wl.nfr_probs, upper=wl.upper)
    span = refocc.occ_bias_track_span(s, e, op)
    bt = ra.log_bias_track(sq[span[0] - 10 - s0:span[1] + 10 - s0], wl.pwm, wl.nucleotides)
    r = refocc.process_occ_chunk(pos, tlen, s, e, op, bias_track=bt, bias_track_start=span[0])
    p = "c%d_" % ci
    for mine, ref in (("vals", "occ_vals"), ("lower_bound", "occ_lower"), ("upper_bound", "occ_upper")):
        assert np.array_equal(r[mine], gold[p + ref], equal_nan=True), mine         # grid points: equal or it is a different answer
    assert np.array_equal(r["cov"], gold[p + "occ_cov"])
    for mine, ref in (("smoothed_vals", "occ_smoothed_vals"), ("smoothed_lower", "occ_smoothed_lower"), ("smoothed_upper", "occ_smoothed_upper")):
        assert np.array_equal(np.isnan(r[mine]), np.isnan(gold[p + ref])), mine
        np.testing.assert_allclose(r[mine], gold[p + ref], rtol=1e-12, atol=0, equal_nan=True, err_msg=mine)
    assert [t[0] for t in r["peaks"]] == list(gold[p + "occ_peak_pos"]) and len(r["peaks"]) > 0
    np.testing.assert_allclose(np.array([t[1:] for t in r["peaks"]], dtype=np.float64), gold[p + "occ_peak_stats"], rtol=1e-12)
    np.testing.assert_allclose(r["nuc_dist"], gold[p + "occ_nuc_dist"], rtol=1e-12, atol=1e-15)
    if ci == 2:
        assert np.isnan(gold[p + "occ_vals"]).sum() > 100     # the sparse chunk has windows without fragments


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_nuc_chunk_equals_reference_run(gold, ci):
    """NucChunk.process (nucleoatac/NucleosomeCalling.py:328-340): tracks, calls, z (the reference's own compiled
    calculateCov in both), likelihood ratio, fuzziness."""
    wl, s, e, pos, tlen, sq, s0 = _case(gold, ci)
    par = refnuc.NucParams((wl.vmat, wl.v_lower, wl.v_upper), wl.fragmentsizes, sd=10)
    _, _, span = refnuc.nuc_geometry(s, e, par)
    bt = ra.log_bias_track(sq[span[0] - 10 - s0:span[1] + 10 - s0], wl.pwm, wl.nucleotides)
    r = refnuc.process_nuc_chunk(pos, tlen, s, e, par, bias_track=bt, bias_track_start=span[0], fit=True)
    p = "c%d_" % ci
    assert np.array_equal(r["nuc_cov"], gold[p + "nuc_nuc_cov"]) and np.array_equal(r["nfr_cov"], gold[p + "nuc_nfr_cov"])
    for mine, ref in (("nuc_signal", "nuc_signal"), ("bias", "nuc_background"), ("norm_signal", "nuc_norm_signal"), ("smoothed", "nuc_smoothed")):
        # same numpy operations; the smoothed track is clipped in place by the reference's getFuzz where calls exist
        a, b = r[mine], gold[p + ref]
        if mine == "smoothed":
            a, b = np.maximum(a, 0), np.maximum(b, 0)
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-15, err_msg=mine)
    keys = sorted(r["nuc_collection"].keys())
    assert [k + s for k in keys] == list(gold[p + "nuc_call_pos"])
    assert sorted(int(k) + s for k in r["nonredundant"]) == list(gold[p + "nuc_nonredundant"])
    assert sorted(int(k) + s for k in r["redundant"]) == list(gold[p + "nuc_redundant"])
    cols = ("z", "lr", "norm_signal", "nuc_signal", "nuc_cov", "nfr_cov", "fuzz", "weight", "fit_pos")
    for row, k in zip(gold[p + "nuc_call_stats"], keys):
        rec = r["nuc_collection"][k]
        for c, ref in zip(cols, row):
            mine = rec[c]                      # fit_pos is relative to the chunk start in the reference as well
            tol = 1e-6 if c in ("fuzz", "weight", "fit_pos") else 1e-9
            assert abs(mine - ref) <= tol * max(1.0, abs(ref)), (k, c, mine, ref)
    if ci == 0:
        assert len(keys) >= 10


def test_chunkmat2d_flip_equals_reference_run(gold):
    """ChunkMat2D.get(flip=True), pyatac/chunkmat2d.py:41-54 (the strand flip of `pyatac vplot`), on integer matrices."""
    for fi in range(int(gold["n_flip"])):
        lower, upper, start, g0, g1, r0, r1 = (int(x) for x in gold["flip%d_args" % fi])
        got = refpyatac.mat_get(gold["flip%d_mat" % fi], start, lower, r0, r1, g0, g1, flip=True)
        assert np.array_equal(got, gold["flip%d_out" % fi]), fi


def test_vplot_helper_equals_reference_run(gold):
    """_vplotHelper (pyatac/make_vplot.py:22-43): Chunk.center, the fragment matrix around the site, the strand flip and the
    per-site scaling, summed over 40 sites of mixed strand and width -- as the reference itself computed them."""
    from nucleoatac_b200 import synth
    k, length, density = gold["vplot_chunk"]
    s, e, pos, tlen, seq, s0 = synth.make_chunk(int(k), length=int(length), density=float(density), seq_margin=int(gold["seq_margin"]))
    strands = {1: "+", -1: "-", 0: "*"}
    for name, scale in (("vplot_plain", False), ("vplot_scaled", True)):
        total = np.zeros((250 - 30, 121))
        for a, b, st in gold["vplot_sites"]:
            total += refpyatac.vplot_site(pos, tlen, int(a), int(b), strands[int(st)], 60, 30, 250, atac=True, scale=scale)
        if scale:
            np.testing.assert_allclose(total, gold[name], rtol=1e-13, atol=0)
        else:
            assert np.array_equal(total, gold[name]) and total.sum() > 1000
    assert (gold["vplot_sites"][:, 2] == -1).sum() >= 5


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_other_configurations_equal_reference_run(gold, ci):
    """The same two reference entry points without a bias model (no --fasta: Occupancy.py:209-211, NucleosomeCalling.py:248)
    and with V-plots other than 251 x 251 whose first size is not 0 (BASELINE configs[4] shapes)."""
    from nucleoatac_b200 import synth
    k, length, density, use_bias, R, W, lower = gold["cases2"][ci]
    wl = synth.Workload(int(R), int(W), lower=int(lower))
    s, e, pos, tlen, seq, s0 = synth.make_chunk(int(k), length=int(length), density=float(density), seq_margin=int(gold["seq_margin"]))
    sq = bytes(seq).decode()
    p = "d%d_" % ci
    if int(R) == 251:
        op = refocc.OccParams(wl.nuc_probs, wl.nfr_probs, upper=251)
        bt, b0 = None, None
        if use_bias:
            span = refocc.occ_bias_track_span(s, e, op)
            bt, b0 = ra.log_bias_track(sq[span[0] - 10 - s0:span[1] + 10 - s0], wl.pwm, wl.nucleotides), span[0]
        r = refocc.process_occ_chunk(pos, tlen, s, e, op, bias_track=bt, bias_track_start=b0)
        for mine, ref in (("vals", "occ_vals"), ("lower_bound", "occ_lower"), ("upper_bound", "occ_upper")):
            assert np.array_equal(r[mine], gold[p + ref], equal_nan=True), mine
        assert np.array_equal(r["cov"], gold[p + "occ_cov"])
        np.testing.assert_allclose(r["smoothed_vals"], gold[p + "occ_smoothed_vals"], rtol=1e-12, equal_nan=True)
        assert [t[0] for t in r["peaks"]] == list(gold[p + "occ_peak_pos"])
    par = refnuc.NucParams((wl.vmat, wl.v_lower, wl.v_upper), wl.fragmentsizes, sd=10)
    bt, b0 = None, None
    if use_bias:
        _, _, span = refnuc.nuc_geometry(s, e, par)
        bt, b0 = ra.log_bias_track(sq[span[0] - 10 - s0:span[1] + 10 - s0], wl.pwm, wl.nucleotides), span[0]
    r = refnuc.process_nuc_chunk(pos, tlen, s, e, par, bias_track=bt, bias_track_start=b0, fit=False)
    assert np.array_equal(r["nuc_cov"], gold[p + "nuc_nuc_cov"])
    for mine, ref in (("nuc_signal", "nuc_signal"), ("bias", "nuc_background"), ("norm_signal", "nuc_norm_signal"), ("smoothed", "nuc_smoothed")):
        a, b = r[mine], gold[p + ref]
        if mine == "smoothed":
            a, b = np.maximum(a, 0), np.maximum(b, 0)
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-15, err_msg=mine)
    keys = sorted(r["nuc_collection"].keys())
    assert [k + s for k in keys] == list(gold[p + "nuc_call_pos"])
    for row, k in zip(gold[p + "nuc_call_zlr"], keys):
        assert abs(r["nuc_collection"][k]["z"] - row[0]) <= 1e-9 * max(1.0, abs(row[0]))
        assert abs(r["nuc_collection"][k]["lr"] - row[1]) <= 1e-9 * max(1.0, abs(row[1]))


def test_chunk_list_equals_reference_run(gold, tmp_path):
    """ChunkList.read -> slop -> merge -> split (pyatac/chunk.py:101-207) on a BED with overlaps, an unknown chromosome and
    regions near the chromosome ends: the host mirror and the oracle's chunk helpers against the reference's own run."""
    from nucleoatac_b200.chunk import ChunkList
    names = ["chrA", "chrB", "chrC", "chrUnknown"]
    chroms = {"chrA": 60000, "chrB": 25000, "chrC": 9000}
    bed = tmp_path / "r.bed"
    bed.write_text("".join("%s\t%d\t%d\n" % (names[c], a, b) for c, a, b in gold["chunklist_bed"]))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cl = ChunkList.read(str(bed), chromDict=chroms, min_offset=300, min_length=240)
    cl.slop(chroms, up=60, down=60)
    cl.merge()
    want = [(names[c], int(a), int(b)) for c, a, b in gold["chunklist_merged"]]
    assert [(c.chrom, c.start, c.end) for c in cl] == want and len(want) >= 5
    assert [len(g) for g in cl.split(items=3)] == list(gold["chunklist_split3"])
    mine = ra.merge_chunks(ra.slop_chunks(ra.read_bed_chunks(str(bed), chroms, min_offset=300, min_length=240), chroms, 60, 60))
    assert [tuple(x) for x in mine] == want


@pytest.mark.parametrize("name", ["nfr_synth", "nfr_bumpy"])
def test_model_nfr_equals_reference_run(gold, name):
    """FragmentMixDistribution.modelNFR (nucleoatac/Occupancy.py:29-66) of the host mirror -- the brute-force grid evaluated
    on arrays, scipy's Nelder-Mead polish -- against the reference's own run of optimize.brute on the same sizes."""
    from nucleoatac_b200.fragmentsizes import FragmentSizes
    from nucleoatac_b200.Occupancy import FragmentMixDistribution
    fm = FragmentMixDistribution(0, upper=251)
    fm.fragmentsizes = FragmentSizes(0, 251, vals=gold[name + "_sizes"].copy())
    fm.modelNFR()
    np.testing.assert_allclose(fm.nfr_fit.get(), gold[name + "_nfr_fit"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(fm.nuc_fit.get(), gold[name + "_nuc_fit"], rtol=1e-12, atol=0)
