"""Pin the CPU oracle (oracle/) on the reference's own evidence for this path.

(a) every known-answer test of the reference's test-suite that touches the path
    (tests/test_xcor.py, test_var.py, test_occupancy.py, test_chunkmat2d.py,
    test_tracks.py, test_utils.py under /root/reference), restated on the oracle;
(b) the outputs the reference shipped in example/example_results, for all 19
    regions of example.bed: fragment sizes, the three occupancy tracks, occpeaks,
    nuc_dist, both nucleoatac_signal tracks and the nucpos / nucpos.redundant calls
    with their z / LR / signal statistics -- to the 12 significant digits the
    reference prints (Python-2 str(float)).
CPU only; runs in well under a minute.
"""
import numpy as np
import pytest

from oracle import mcov, refalgo as ra, refnuc, refocc
from tests.fixtures import track_close

UPPER, FLANK, SEP = 251, 60, 120


# --------------------------------------------------------------------------- (a) reference KATs
def test_kat_call_peaks():
    """/root/reference/tests/test_utils.py:8-19"""
    sig = [1, 2, 3, 2, 1, 4, 1, 2, 1, 0, 0]
    assert np.array_equal(ra.call_peaks(np.array(sig, dtype=float), min_signal=1, sep=3), [2, 5])
    assert np.array_equal(ra.call_peaks(np.array(sig, dtype=float), min_signal=1, sep=1), [2, 5, 7])
    assert np.array_equal(ra.call_peaks(np.array(sig, dtype=float), min_signal=3, sep=2), [2, 5])


def test_kat_occupancy():
    """/root/reference/tests/test_occupancy.py:8-57"""
    p = refocc.OccCalcParams(0, 3, nuc_fit=[0.01, 0.49, 0.5], nfr_fit=[0.5, 0.49, 0.01])
    assert refocc.calculate_occupancy(np.array([1, 0, 0]), np.array([1, 1, 1]), p)[0] == 0
    assert refocc.calculate_occupancy(np.array([1, 1, 1]), np.array([1, 1, 1]), p)[0] == 0.5
    rng = np.random.RandomState(7)
    for bias in (np.array([1, 1, 1]), np.array([3, 2, 1])):
        nfrp = p.nfr_probs * bias / np.sum(p.nfr_probs * bias)
        nucp = p.nuc_probs * bias / np.sum(p.nuc_probs * bias)
        res = np.array([refocc.calculate_occupancy(rng.multinomial(10, nfrp) + rng.multinomial(30, nucp), bias, p)
                        for _ in range(100)])
        assert abs(res[:, 0].mean() - 0.75) < 0.1
        assert (res[:, 2] < 0.75).sum() < 85 and (res[:, 1] > 0.75).sum() < 85


def test_kat_chunkmat_get():
    """/root/reference/tests/test_chunkmat2d.py:13-17: coordinate slicing of a 200 x 500 matrix."""
    mat = np.zeros((200, 500))
    mat[100, 5] = 1
    start, lower = 500, 0
    got = mat[(100 - lower):(102 - lower), (505 - start):(507 - start)]
    assert np.array_equal(got, [[1, 0], [0, 0]])


def test_kat_biasmat(example):
    """/root/reference/tests/test_chunkmat2d.py:20-49 on the Scores track of the first bed region."""
    z = example.z
    s0, e0 = int(z["raw0_start"]), int(z["raw0_end"])
    scores = z["raw0_scores"]
    ms, me, lo, up = s0 + 100, e0 - 100, 100, 200
    bias = scores[(ms - up // 2) - s0:(me + up // 2) - s0]
    mat = ra.make_bias_mat(bias, lo, up)
    lit = ra.make_bias_mat_literal(bias, lo, up)
    assert mat.shape == (100, me - ms)
    np.testing.assert_allclose(mat, lit, rtol=1e-13)
    get = lambda pos: scores[pos - s0]
    assert abs(np.exp(get(ms - 49) + get(ms + 50)) - mat[0, 0]) < 1e-12 * mat[0, 0]
    assert abs(np.exp(get(ms + 145) + get(ms + 295)) - mat[51, 220]) < 1e-12 * mat[51, 220]
    normed = ra.norm_by_insert_dist(mat, np.arange(100, 200, dtype=float))
    assert abs(np.exp(get(ms - 50) + get(ms + 50)) * 101 - normed[1, 0]) < 1e-12 * normed[1, 0]


def test_kat_ins_methods(example):
    """/root/reference/tests/test_tracks.py:16-23 on single_read.bam: getIns == getInsertions."""
    z = example.z
    s0, e0 = int(z["raw0_start"]), int(z["raw0_end"])
    pos, tlen = z["single_pos"], z["single_tlen"]
    assert len(pos) == 1
    ins1 = ra.get_insertions(pos, tlen, s0, e0, 0, 2000)
    mat = ra.make_fragment_mat(pos, tlen, s0, e0, 0, 100)
    for literal in (True, False):
        ins2, i0, i1 = ra.get_ins(mat, s0, e0, 0, 100, literal=literal)
        assert np.array_equal(ins1[100:300], ins2[(s0 + 100 - i0):(s0 + 300 - i0)])
    assert mat.sum() == 1 and ins1[100:300].sum() >= 1  # the KAT is not vacuous


def test_kat_track_value(golden):
    """/root/reference/tests/test_tracks.py:31-36"""
    assert abs(1.35994655714 - float(golden["scores_706661"])) < 0.001


def test_kat_xcor(example):
    """/root/reference/tests/test_xcor.py:11-32 (example/example.VMat, 130 x 121, sizes [115,245))."""
    z = example.z
    V, lv, uv = example.vmat_example
    w = V.shape[1] // 2
    s0, e0 = int(z["raw0_start"]), int(z["raw0_end"])
    mat = ra.make_fragment_mat(z["raw0_pos"], z["raw0_tlen"], s0 - w, e0 + w, lv, uv)
    assert mat.sum() > 100
    sig = refnuc.calculate_signal(mat, s0 - w, e0 + w, lv, s0, V, lv, uv)
    sig_direct = refnuc.calculate_signal(mat, s0 - w, e0 + w, lv, s0, V, lv, uv, method="direct")
    np.testing.assert_allclose(sig, sig_direct, atol=1e-12)
    for off in (0, 100):
        a = np.sum(mat[:, off:off + 2 * w + 1] * V)
        assert abs(a - sig_direct[off]) < 1e-10


def test_kat_variance(example):
    """/root/reference/tests/test_var.py:34-43 + the reference's own compiled calculateCov (oracle/_ref)."""
    z = example.z
    V, lv, uv = example.vmat_example
    w = V.shape[1] // 2
    s0, e0 = int(z["raw0_start"]), int(z["raw0_end"])
    scores = z["raw0_scores"]
    ms, me, lo, up = s0 + 200, e0 - 200, 100, 250
    bmat = ra.make_bias_mat(scores[(ms - up // 2) - s0:(me + up // 2) - s0], lo, up)
    pos = s0 + 300
    sub = bmat[(lv - lo):(uv - lo), (pos - w - ms):(pos + w + 1 - ms)]
    prob = sub / np.sum(sub)
    p, v, reads = prob.flatten(), V.flatten(), 35
    var_term = np.sum(prob * (1 - prob) * V ** 2)
    tmp = prob * V
    cov_term = np.sum(np.outer(tmp, tmp)) - np.sum(tmp ** 2)
    sd_alt = np.sqrt(reads * (var_term - cov_term))
    sd_pair = np.sqrt(mcov.calculate_cov(p, v, reads))
    sd_closed = np.sqrt(mcov.calculate_cov(p, v, reads, closed=True))
    assert abs(sd_alt - sd_pair) < 0.001 * sd_alt
    assert abs(sd_pair - sd_closed) < 1e-10 * sd_pair
    ref0 = mcov.reference_calculate_cov()
    if ref0 is not None:  # the reference's own Cython, compiled by oracle/build.py
        def ref(*a):
            # multinomial_cov.pyx:23 never initialises its accumulator: back-to-back calls ADD the previous
            # result (observed here: 1x, 2x, 3x).  The reference's shipped z-scores correspond to a zero start
            # (test_golden_nuc), which is what any other call in between leaves on the stack.
            mcov.calculate_cov(np.zeros(4), np.zeros(4), 1)  # leaves 0.0 in the stale slot
            return ref0(*a)
        assert abs(np.sqrt(ref(p, v, reads)) - sd_pair) < 1e-12 * sd_pair
        assert ref(p, v, 35.7) == ref(p, v, 35)  # C-int truncation of r
    rng = np.random.RandomState(1)  # test_var.py:27-33, 5000-draw simulation within 5 %
    sims = rng.multinomial(reads, p, 5000).dot(v)
    assert abs(np.std(sims) - sd_pair) < 0.05 * np.std(sims)
    with pytest.raises(ValueError):
        mcov.calculate_cov(p, v[:-1], reads)


# --------------------------------------------------------------------------- (b) example_results
def _occ_params(example):
    return refocc.OccParams(example.occ_fit[1], example.occ_fit[2], upper=UPPER, sep=SEP, flank=FLANK)


def _bias_track(example, i, span):
    ts, te = span
    seq = example.seq_slice(i, ts - example.pwm_up, te + example.pwm_down)
    bt = ra.log_bias_track(seq, example.pwm, example.nucleotides)
    assert len(bt) == te - ts
    return bt


def test_golden_fragment_sizes(example, golden):
    """example.fragmentsizes.txt == getFragmentSizesFromChunkList over the 19 merged chunks."""
    cnt = np.zeros(UPPER)
    for i in range(example.n_chunks):
        _, s, e = example.chunk(i)
        cnt += ra.fragment_size_counts(*example.reads(i), [(s, e)], 0, UPPER)
    assert cnt.sum() == 32792
    np.testing.assert_allclose(ra.normalize_sizes(cnt), golden["fragmentsizes"], atol=1e-13)
    np.testing.assert_allclose(example.occ_fit[0], golden["fragmentsizes"], atol=1e-13)


def test_golden_occ(example, golden):
    """occ / lower / upper bedgraphs, occpeaks.bed and nuc_dist.txt for all 19 regions."""
    params = _occ_params(example)
    off = golden["track_off"]
    nuc_dist = np.zeros(UPPER)
    peaks = []
    for i in range(example.n_chunks):
        chrom, s, e = example.chunk(i)
        span = refocc.occ_bias_track_span(s, e, params)
        r = refocc.process_occ_chunk(*example.reads(i), s, e, params,
                                     bias_track=_bias_track(example, i, span), bias_track_start=span[0])
        for key, mine in (("occ", "smoothed_vals"), ("occ_lower", "smoothed_lower"), ("occ_upper", "smoothed_upper")):
            ok, worst = track_close(golden[key][off[i]:off[i + 1]], r[mine], slack=1.6)
            assert ok, (chrom, s, e, key, worst)
        nuc_dist += r["nuc_dist"]
        peaks += [(example.chrom_names.index(chrom),) + p for p in r["peaks"]]
    assert len(peaks) == len(golden["occpeaks_pos"]) == 160
    assert [p[1] for p in peaks] == list(golden["occpeaks_pos"])
    assert [p[0] for p in peaks] == list(golden["occpeaks_chrom"])
    ok, worst = track_close(golden["occpeaks_vals"], np.array([p[2:] for p in peaks]), slack=1.6)
    assert ok, worst
    ok, worst = track_close(golden["nuc_dist"], nuc_dist, slack=1.6)
    assert ok, worst


def test_golden_nuc(example, golden):
    """nucleoatac_signal / .smooth bedgraphs and the nucpos / nucpos.redundant calls (columns 4-12)."""
    params = refnuc.NucParams(example.vmat, example.fragmentsizes, sd=10)
    off = golden["track_off"]
    calls = {"nucpos": [], "redundant": []}
    for i in range(example.n_chunks):
        chrom, s, e = example.chunk(i)
        _, _, span = refnuc.nuc_geometry(s, e, params)
        occ = [golden[k][off[i]:off[i + 1]] for k in ("occ", "occ_lower", "occ_upper")]
        r = refnuc.process_nuc_chunk(*example.reads(i), s, e, params, bias_track=_bias_track(example, i, span),
                                     bias_track_start=span[0], occ_tracks=occ, fit=False)
        for key, mine in (("nuc_signal", "norm_signal"), ("nuc_smooth", "smoothed")):
            g = golden[key][off[i]:off[i + 1]]  # norm = signal - background: allow 1e-13 of the raw signal scale
            ok, worst = track_close(g, r[mine], slack=1.6, atol=1e-13 * max(1.0, np.nanmax(np.abs(r["nuc_signal"]))))
            assert ok, (chrom, s, e, key, worst)
        ci = example.chrom_names.index(chrom)
        for name, keys in (("nucpos", r["nonredundant"]), ("redundant", r["redundant"])):
            for k in sorted(int(x) for x in keys):
                n = r["nuc_collection"][k]
                calls[name].append((ci, n["pos"], n["z"], n["occ"], n["occ_lower"], n["occ_upper"], n["lr"],
                                    n["norm_signal"], n["nuc_signal"], n["nuc_cov"], n["nfr_cov"]))
    for name, n_expected in (("nucpos", 130), ("redundant", 12)):
        mine = calls[name]
        assert len(mine) == n_expected == len(golden[name + "_pos"])
        assert [m[1] for m in mine] == list(golden[name + "_pos"])
        assert [m[0] for m in mine] == list(golden[name + "_chrom"])
        gold_vals = golden[name + "_vals"][:, :9]  # column 13 (fuzz) is the host L-BFGS-B fit, not on the device path
        ok, worst = track_close(gold_vals, np.array([m[2:] for m in mine]), slack=6.0)
        assert ok, (name, worst)


def test_golden_nfr():
    """`nucleoatac nfr` as `nucleoatac run` wires it (cli.py:47-49; --max_occ 0.1, --max_occ_upper 0.25 of cli.py:167-170):
    the oracle reproduces example_results/example.nfrpos.bed.gz row for row as text and example.ins.bedgraph.gz exactly."""
    import os
    from oracle import refnfr
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nfr_golden.npz"))
    names = [str(x) for x in z["chrom_names"]]
    nucs = [str(x) for x in z["pwm_nucleotides"]]
    text, n_chunks = "", len(z["chunk_start"])
    for i in range(n_chunks):
        s, e = int(z["chunk_start"][i]), int(z["chunk_end"][i])
        a, b = z["frag_off"][i:i + 2]
        sa, sb = z["seq_off"][i:i + 2]
        ta, tb = z["track_off"][i:i + 2]
        na, nb = z["nuc_off"][i:i + 2]
        lb = ra.log_bias_track(bytes(z["seq"][sa:sb]).decode(), z["pwm"], nucs)
        assert len(lb) == e - s
        recs, ins = refnfr.process_nfr_chunk(z["frag_pos"][a:b], z["frag_tlen"][a:b], s, e, z["nuc_pos"][na:nb], z["occ"][ta:tb],
                                             z["occ_upper"][ta:tb], lb, max_occ=0.1, max_occ_upper=0.25)
        assert np.array_equal(ins, z["gold_ins"][ta:tb])  # integer counts: bit-exact
        text += "".join(refnfr.nfr_bed(names[int(z["chunk_chrom"][i])], r) + "\n" for r in recs)
    assert text == str(z["gold_nfr_text"]) and len(z["gold_nfr_left"]) == 14
