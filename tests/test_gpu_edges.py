"""GPU parity on the awkward inputs: ragged chunk lengths (not multiples of the window step or of any tile), chunks
with no / one / very many fragments, negative and out-of-range template lengths, non-ATAC mode, a VMat with
lower > 0 and an even number of rows, no-bias mode -- each against the CPU oracle."""
import numpy as np
import pytest

from oracle import refalgo as ra, refnuc, refocc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from nucleoatac_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _chunk(k, length, density, rng_seed=0, margin=420):
    from nucleoatac_b200 import synth
    s, e, pos, tlen, seq, s0 = synth.make_chunk(k, length=length, density=density, seq_margin=margin)
    rng = np.random.RandomState(rng_seed + k)
    if len(pos):
        # sprinkle the nasty cases: reverse-strand style negative tlen, tiny and huge templates, duplicates
        tlen = tlen.copy()
        n = len(tlen)
        tlen[rng.rand(n) < 0.2] *= -1
        tlen[rng.rand(n) < 0.02] = 5          # ATAC size -3 -> dropped
        tlen[rng.rand(n) < 0.02] = 8          # ATAC size 0
        tlen[rng.rand(n) < 0.02] = 9          # ATAC size 1 (single-tap bias cell)
        tlen[rng.rand(n) < 0.02] = 3000
        dup = rng.rand(n) < 0.05
        pos, tlen = np.concatenate([pos, pos[dup]]), np.concatenate([tlen, tlen[dup]])
    return s, e, pos.astype(np.int32), tlen.astype(np.int32), seq, s0


def _cmp_occ(out, pb, j, r):
    a, b = int(pb.out_off[j]), int(pb.out_off[j + 1])
    for key in ("vals", "lower_bound", "upper_bound"):
        assert np.array_equal(out[key][a:b], r[key], equal_nan=True), (j, key)
    for key in ("smoothed_vals", "smoothed_lower", "smoothed_upper"):
        np.testing.assert_allclose(out[key][a:b], r[key], rtol=1e-9, equal_nan=True, err_msg=str((j, key)))
    assert np.array_equal(out["cov"][a:b], r["cov"])
    n, po = int(out["peak_count"][j]), int(out["peak_off"][j])
    assert list(out["peak_pos"][po:po + n]) == [p[0] for p in r["peaks"]], j
    np.testing.assert_allclose(out["nuc_dist"][j], r["nuc_dist"], rtol=1e-9, atol=1e-15)


def _cmp_nuc(out, pb, j, r, s, rtol=1e-9):
    a, b = int(pb.out_off[j]), int(pb.out_off[j + 1])
    assert np.array_equal(out["nuc_cov"][a:b], r["nuc_cov"]) and np.array_equal(out["nfr_cov"][a:b], r["nfr_cov"])
    scale = max(1.0, float(np.abs(r["nuc_signal"]).max()), float(np.nanmax(np.abs(r["bias"]))))
    for key, okey in (("nuc_signal", "nuc_signal"), ("background", "bias"), ("norm_signal", "norm_signal"), ("smoothed", "smoothed")):
        np.testing.assert_allclose(out[key][a:b], r[okey], rtol=rtol, atol=rtol * scale, equal_nan=True, err_msg=str((j, key)))
    n, co = int(out["cand_count"][j]), int(out["cand_off"][j])
    assert list(out["cand_pos"][co:co + n] - s) == list(r["cands"]), j
    kept = [int(p - s) for p, f in zip(out["cand_pos"][co:co + n], out["cand_flag"][co:co + n]) if f & 4]
    assert kept == sorted(r["nuc_collection"].keys())
    nonred = [int(p - s) for p, f in zip(out["cand_pos"][co:co + n], out["cand_flag"][co:co + n]) if f & 8]
    assert nonred == sorted(int(x) for x in r["nonredundant"])


@pytest.mark.parametrize("use_bias", [True, False])
@pytest.mark.parametrize("atac", [True, False])
def test_ragged_empty_dense(eng, use_bias, atac):
    from nucleoatac_b200 import synth
    from nucleoatac_b200.engine import PackedBatch
    wl = synth.Workload(130, 121, upper=251, lower=115)          # even row count, lower > 0 (nfr_cov path)
    wl.configure(eng, use_bias=use_bias, xcor_mode=1)
    eng.configure_nuc(sd=10, use_bias=use_bias, xcor_mode=1, atac=atac)
    eng.configure_occ(upper=251, use_bias=use_bias, atac=atac)
    specs = [(0, 777, 0.25), (1, 1213, 0.0), (2, 2501, 2.0), (3, 400, 0.8), (4, 5000, 0.002), (5, 1001, 0.3)]
    chunks = [_chunk(k, L, d) for k, L, d in specs]
    pb = PackedBatch.from_chunks(chunks)
    h = eng.upload(pb)
    eng.occ_run(h)
    eng.nuc_run(h)
    oout, nout = eng.occ_alloc(pb), eng.nuc_alloc(pb)
    eng.occ_download(h, oout)
    eng.nuc_download(h, nout)
    eng.sync(h)
    eng.free_batch(h)
    oparams = refocc.OccParams(wl.nuc_probs, wl.nfr_probs, upper=251)
    nparams = refnuc.NucParams((wl.vmat, wl.v_lower, wl.v_upper), wl.fragmentsizes, sd=10, atac=atac)
    assert int(oout["peak_count"][1]) == 0 and np.isnan(oout["smoothed_vals"][pb.out_off[1]:pb.out_off[2]]).all()
    for j, (s, e, pos, tlen, seq, s0) in enumerate(chunks):
        sq = bytes(seq).decode()
        span = refocc.occ_bias_track_span(s, e, oparams)
        bt = ra.log_bias_track(sq[span[0] - 10 - s0:span[1] + 10 - s0], wl.pwm, wl.nucleotides) if use_bias else None
        ro = refocc.process_occ_chunk(pos, tlen, s, e, oparams, bias_track=bt, bias_track_start=span[0], atac=atac)
        _cmp_occ(oout, pb, j, ro)
        _, _, span = refnuc.nuc_geometry(s, e, nparams)
        bt = ra.log_bias_track(sq[span[0] - 10 - s0:span[1] + 10 - s0], wl.pwm, wl.nucleotides) if use_bias else None
        # scipy's fft correlate leaves ~1e-17 noise where the exact signal is 0, which the reference turns into
        # spurious zero-coverage candidates (never calls); the direct method gives the exact zeros the device computes
        rn = refnuc.process_nuc_chunk(pos, tlen, s, e, nparams, bias_track=bt, bias_track_start=span[0], fit=False,
                                      xcor_method="direct")
        _cmp_nuc(nout, pb, j, rn, s)


def test_tensor_core_ragged(eng):
    """The tcgen05 path on ragged / sparse chunks (partial x-tiles, zero-read chunks) and a 0-based VMat with size-1 rows."""
    from nucleoatac_b200 import synth
    from nucleoatac_b200.engine import PackedBatch
    wl = synth.Workload(201, 151)
    wl.configure(eng, use_bias=True, xcor_mode=2)
    specs = [(0, 300, 0.3), (1, 1213, 0.0), (2, 2501, 1.0), (3, 257, 0.5), (4, 3333, 0.25)]
    chunks = [_chunk(k, L, d) for k, L, d in specs]
    pb = PackedBatch.from_chunks(chunks)
    out = eng.process_nuc(pb)
    nparams = refnuc.NucParams((wl.vmat, wl.v_lower, wl.v_upper), wl.fragmentsizes, sd=10)
    for j, (s, e, pos, tlen, seq, s0) in enumerate(chunks):
        _, _, span = refnuc.nuc_geometry(s, e, nparams)
        bt = ra.log_bias_track(bytes(seq).decode()[span[0] - 10 - s0:span[1] + 10 - s0], wl.pwm, wl.nucleotides)
        r = refnuc.process_nuc_chunk(pos, tlen, s, e, nparams, bias_track=bt, bias_track_start=span[0], fit=False,
                                     xcor_method="direct")
        a, b = int(pb.out_off[j]), int(pb.out_off[j + 1])
        scale = max(1e-300, float(np.abs(r["nuc_signal"]).max()), float(np.abs(r["bias"]).max()))
        err = float(np.abs(out["background"][a:b] - r["bias"]).max()) / scale
        assert err <= 1e-5, (j, err)
        assert np.array_equal(out["nuc_cov"][a:b], r["nuc_cov"])
        n, co = int(out["cand_count"][j]), int(out["cand_off"][j])
        kept = [int(p - s) for p, f in zip(out["cand_pos"][co:co + n], out["cand_flag"][co:co + n]) if f & 4]
        assert kept == sorted(r["nuc_collection"].keys())


def test_size_independent_properties(eng):
    """Properties that hold at any size (checked on a 64-chunk batch): the dense background track is linear in the
    fragment-size distribution; coverage equals the count of fragment centres; nuc_dist rows sum to the peak count;
    a second identical run is bit-identical (deterministic kernels)."""
    from nucleoatac_b200 import synth
    wl = synth.Workload(251, 251)
    wl.configure(eng, use_bias=True, xcor_mode=2)
    pb = synth.make_batch(100, 64)
    o1 = eng.process_nuc(pb)
    o2 = eng.process_nuc(pb)
    for k in ("background", "norm_signal", "smoothed", "cand_pos", "cand_flag", "cand_z"):
        assert np.array_equal(o1[k], o2[k], equal_nan=True), k
    eng.set_fragment_sizes(wl.fragmentsizes * 0.5)
    eng.configure_nuc(sd=10, use_bias=True, xcor_mode=2)
    o3 = eng.process_nuc(pb)
    np.testing.assert_allclose(o3["background"], o1["background"], rtol=1e-6)   # bx and bcov both scale: ratio unchanged
    wl.configure(eng, use_bias=True)
    oc = eng.process_occ(pb, raw=False)
    counts = oc["peak_count"]
    np.testing.assert_allclose(oc["nuc_dist"].sum(axis=1), counts, rtol=1e-12)
    # coverage at position x = fragments (size < upper) centred within +-60: its sum over the chunk counts every centre
    # inside [start+60, end-60) exactly 121 times
    l = pb.frag_pos.astype(np.int64) + 4
    i = np.abs(pb.frag_tlen.astype(np.int64)) - 8
    c = l + (i - 1) // 2
    j = 7
    s, e = int(pb.starts[j]), int(pb.ends[j])
    f0, f1 = int(pb.frag_off[j]), int(pb.frag_off[j + 1])
    ok = (i[f0:f1] >= 0) & (i[f0:f1] < 251)
    cj = c[f0:f1][ok]
    expect = sum(max(0, min(int(x) + 60, e - 1) - max(int(x) - 60, s) + 1) for x in cj)
    assert oc["cov"][pb.out_off[j]:pb.out_off[j + 1]].sum() == expect
