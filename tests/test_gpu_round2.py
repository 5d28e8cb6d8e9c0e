"""GPU parity of the round-2 kernels through the C-ABI: the guarded occupancy search against the full grid scan, the
block-form smoother against the tap-by-tap one, float32 track delivery, the tcgen05 contraction over the VMat size sweep
of BASELINE.json configs[4] against the fp64 oracle, and first-in first-out ordering of passes across batches."""
import os

import numpy as np
import pytest

from oracle import refalgo as ra, refnuc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from nucleoatac_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _mixed_batch():
    """48 synthetic chunks: the bench's density, sparse ones (windows without fragments), very dense ones (more fragments per
    window than the kernel's cache holds) and lengths that are not multiples of the window step."""
    from nucleoatac_b200 import synth
    from nucleoatac_b200.engine import PackedBatch
    chunks = []
    for k in range(48):
        length = 10000 if k % 4 else (3001 + 7 * k)
        density = (0.25, 0.25, 0.02, 1.5, 0.25, 0.005, 0.25, 0.6)[k % 8]
        chunks.append(synth.make_chunk(k, length=length, density=density))
    return PackedBatch.from_chunks(chunks)


def _occ(eng, pb, env=None):
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        eng.profile_reset()
        out = eng.process_occ(pb)
        return out, eng.profile_report()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("use_bias", [True, False])
def test_occ_search_equals_full_scan(eng, use_bias):
    """The guarded searches over the alpha grid (3 rounds of 16 grid points, Occupancy.py:104-120; NB200_MLE_SEARCH=tw: one
    thread per window, =group: 8 lanes per window) must return what the default scan of all 101 grid points returns -- bit
    for bit, on every window -- and the block-form smoother what the tap-by-tap smoother returns."""
    from nucleoatac_b200 import synth
    wl = synth.Workload(251, 251)
    wl.configure(eng, use_bias=use_bias)
    pb = _mixed_batch()
    new, prof_new = _occ(eng, pb)
    tw, prof_tw = _occ(eng, pb, {"NB200_MLE_SEARCH": "tw"})
    grp, prof_grp = _occ(eng, pb, {"NB200_MLE_SEARCH": "group"})
    full, prof_full = _occ(eng, pb, {"NB200_OCC_SMOOTH_DENSE": "1"})
    ran = lambda prof, k: prof.get(k, (0, 0.0))[0] > 0      # launches since profile_reset
    assert ran(prof_new, "k_occ_mle") and ran(prof_new, "k_occ_smooth_blocks") and not ran(prof_new, "k_occ_mle_tw")
    assert ran(prof_tw, "k_occ_mle_tw") and not ran(prof_tw, "k_occ_mle")
    assert ran(prof_grp, "k_occ_mle_search") and not ran(prof_grp, "k_occ_mle_tw")
    assert ran(prof_full, "k_occ_mle") and ran(prof_full, "k_smooth_same") and not ran(prof_full, "k_occ_smooth_blocks")
    for other, what in ((tw, "thread-per-window search"), (grp, "group search")):
        for key in ("vals", "lower_bound", "upper_bound", "smoothed_vals", "smoothed_lower", "smoothed_upper", "peak_pos", "peak_count"):
            assert np.array_equal(new[key], other[key], equal_nan=True), (what, key)
    nwin = 0
    for key in ("vals", "lower_bound", "upper_bound"):
        assert np.array_equal(new[key], full[key], equal_nan=True), key
        nwin = int((~np.isnan(new[key])).sum()) // 5
    assert nwin > 50000 and np.isnan(new["vals"]).sum() > 1000     # both populated and empty windows were compared
    for key in ("smoothed_vals", "smoothed_lower", "smoothed_upper"):
        assert np.array_equal(np.isnan(new[key]), np.isnan(full[key])), key
        np.testing.assert_allclose(new[key], full[key], rtol=1e-12, atol=1e-15, equal_nan=True, err_msg=key)
    assert np.array_equal(new["cov"], full["cov"])
    # Peaks.  Where no window is missing the two smoothers' tracks differ in the last bits only and must give the same
    # peaks.  Chunks with NaN windows have plateaus (stretches of equal windows, occupancy 1.0 above all) on which
    # call_peaks' 1e-12 jitter picks the maxima and reduce_peaks then ranks exact ties -- by np.argsort's unstable order in
    # the reference, i.e. implementation-defined -- so peaks are not compared there.  What the reference does guarantee on
    # such a plateau is ONE value: utils.smooth sums numerator and denominator by the same routine over identical arrays
    # (x = 1.0 * indicator), so the quotient is exactly 1.0.  The block smoother reproduces that; the tap-by-tap kernel
    # returns 1 - 3e-16 wherever its window holds no NaN (it divides by a separately summed window there).
    def peaks(out, j):
        po, n = int(out["peak_off"][j]), int(out["peak_count"][j])
        return [int(x) for x in out["peak_pos"][po:po + n]]
    n_same = n_plateau = 0
    for j in range(pb.n):
        a, b = int(pb.out_off[j]), int(pb.out_off[j + 1])
        raw = new["vals"][a:b]
        if not np.isnan(raw).any():
            assert peaks(new, j) == peaks(full, j), j
            n_same += 1
        ones = np.where(np.isnan(raw), 1.0, raw) == 1.0      # windows that are 1.0 or missing
        present = ~np.isnan(raw)
        # positions whose whole smoothing window (+-60) holds only 1.0 / missing values and at least one present value
        c_ones = np.concatenate(([0], np.cumsum(ones)))
        c_pres = np.concatenate(([0], np.cumsum(present)))
        x = np.arange(60, (b - a) - 60)
        sel = ((c_ones[x + 61] - c_ones[x - 60]) == 121) & ((c_pres[x + 61] - c_pres[x - 60]) > 0)
        if sel.any():
            assert np.all(new["smoothed_vals"][a:b][x[sel]] == 1.0), j
            n_plateau += int(sel.sum())
    assert n_same >= 20 and n_plateau > 1000
    dense = np.array([not np.isnan(new["vals"][int(pb.out_off[j]):int(pb.out_off[j + 1])]).any() for j in range(pb.n)])
    np.testing.assert_allclose(new["nuc_dist"][dense], full["nuc_dist"][dense], rtol=1e-12, atol=1e-15)


def test_occ_scan_forms_agree(eng):
    """The forms of the grid scan k_occ_mle (Occupancy.py:104-146) -- inputs staged in shared memory or loaded by every warp,
    the short epilogue on plain doubles or the (exponent, mantissa) one, 8 or 16 lanes per window -- return the same grids
    bit for bit on sparse, dense (renormalised products, more fragments under a block than its staging area holds) and ragged
    chunks, with and without the bias model."""
    from nucleoatac_b200 import synth
    pb = _mixed_batch()
    for use_bias in (True, False):
        wl = synth.Workload(251, 251)
        wl.configure(eng, use_bias=use_bias)
        ref, prof = _occ(eng, pb)
        assert prof.get("k_occ_mle", (0, 0.0))[0] > 0
        for env in ({"NB200_MLE_STAGE": "0"}, {"NB200_MLE_EPI": "canonical"}, {"NB200_MLE_STAGE": "0", "NB200_MLE_EPI": "canonical"},
                    {"NB200_MLE_GL": "16", "NB200_MLE_LB": "6"}, {"NB200_MLE_GL": "16", "NB200_MLE_LB": "8", "NB200_MLE_EPI": "canonical"},
                    {"NB200_MLE_LB": "5"}):
            out, prof = _occ(eng, pb, env)
            assert prof.get("k_occ_mle", (0, 0.0))[0] > 0
            for key in ("vals", "lower_bound", "upper_bound", "smoothed_vals", "peak_pos", "peak_count", "nuc_dist"):
                assert np.array_equal(ref[key], out[key], equal_nan=True), (use_bias, env, key)
    wl = synth.Workload(251, 251)
    wl.configure(eng, use_bias=True)


def test_occ_search_other_grids(eng):
    """Grids other than linspace(0, 1, 101): 17, 64 and 121 points go through the search; a cutoff of 0 (nothing but the maximum
    passes) and a huge one (everything passes)."""
    from nucleoatac_b200 import synth
    wl = synth.Workload(251, 251)
    wl.configure(eng, use_bias=True)
    pb = synth.make_batch(3, 6)
    for n_alpha, cutoff in ((17, None), (121, None), (101, 0.0), (101, 1e9), (64, 2.0)):
        eng.set_occ_model(wl.nuc_probs, wl.nfr_probs, alphas=np.linspace(0, 1, n_alpha), cutoff=cutoff)
        eng.configure_occ(upper=wl.upper, use_bias=True)
        full, _ = _occ(eng, pb)
        tw, prof = _occ(eng, pb, {"NB200_MLE_SEARCH": "tw"})
        grp, _ = _occ(eng, pb, {"NB200_MLE_SEARCH": "group"})
        assert prof.get("k_occ_mle_tw", (0, 0.0))[0] > 0
        for key in ("vals", "lower_bound", "upper_bound"):
            assert np.array_equal(tw[key], full[key], equal_nan=True), (n_alpha, cutoff, key, "tw")
            assert np.array_equal(grp[key], full[key], equal_nan=True), (n_alpha, cutoff, key, "group")
    wl.configure(eng, use_bias=True)


def test_download32_is_the_rounded_float64(eng):
    """nb200_{occ,nuc}_download32: every per-position track equals float32(the float64 track); tables are unchanged."""
    from nucleoatac_b200 import synth
    wl = synth.Workload(251, 251)
    wl.configure(eng, use_bias=True, xcor_mode=2)
    pb = synth.make_batch(11, 5)
    h = eng.upload(pb)
    eng.nuc_run(h)
    eng.occ_run(h)
    n64, o64 = eng.nuc_alloc(pb), eng.occ_alloc(pb)
    n32, o32 = eng.nuc_alloc(pb, track_dtype=np.float32), eng.occ_alloc(pb, track_dtype=np.float32)
    b64 = eng.nuc_download(h, n64) + eng.occ_download(h, o64)
    b32 = eng.nuc_download(h, n32) + eng.occ_download(h, o32)
    eng.sync(h)
    eng.free_batch(h)
    tl = pb.total_len
    assert b64 - b32 == 4 * tl * (len(eng.NUC_TRACKS) + len(eng.OCC_TRACKS))
    for out64, out32, names in ((n64, n32, eng.NUC_TRACKS), (o64, o32, eng.OCC_TRACKS)):
        for k, v in out64.items():
            if k in names:
                assert out32[k].dtype == np.float32
                assert np.array_equal(out32[k], v.astype(np.float32), equal_nan=True), k
            else:
                assert np.array_equal(out32[k], v, equal_nan=True), k
    with pytest.raises(TypeError):
        bad = dict(o32)
        bad["cov"] = o64["cov"]
        eng._track_dtype(bad, eng.OCC_TRACKS)


def _shrink_vmat(eng, wl):
    """An earlier test may have left a larger VMat than this workload's fragment sizes cover (the library refuses sizes that
    do not reach vmat.upper): long sizes first, then the smaller VMat, then Workload.configure can set its own sizes."""
    eng.set_fragment_sizes(np.full(1200, 1.0 / 1200))
    eng.set_vmat(wl.vmat, wl.v_lower, wl.v_upper)


@pytest.mark.parametrize("size", [101, 151, 201, 301, 401, 501])
def test_tensor_core_vmat_sweep(eng, size):
    """BASELINE configs[4]: the tcgen05 background cross-correlation (NucleosomeCalling.py:60-63) at VMat sizes 101^2 .. 501^2
    against the fp64 oracle: every track within 1e-5 of the signal scale (north_star: floats within 1e-5), coverage exact."""
    from nucleoatac_b200 import synth
    from nucleoatac_b200.engine import PackedBatch
    wl = synth.Workload(size, size, upper=max(251, size))
    _shrink_vmat(eng, wl)
    wl.configure(eng, use_bias=True, xcor_mode=2)
    params = refnuc.NucParams((wl.vmat, wl.v_lower, wl.v_upper), wl.fragmentsizes, sd=10)
    margin = size + size // 2 + 40
    chunks = [synth.make_chunk(k, length=3000, seq_margin=margin) for k in (2, 9)]
    pb = PackedBatch.from_chunks(chunks)
    eng.profile_reset()
    out = eng.process_nuc(pb)
    rep = eng.profile_report()
    assert rep.get("k_nuc_bx_ts", (0, 0.0))[0] + rep.get("k_nuc_bx_tc", (0, 0.0))[0] > 0
    worst = 0.0
    for j, (s, e, pos, tlen, seq, s0) in enumerate(chunks):
        _, _, span = refnuc.nuc_geometry(s, e, params)
        bt = ra.log_bias_track(bytes(seq).decode()[span[0] - 10 - s0:span[1] + 10 - s0], wl.pwm, wl.nucleotides)
        r = refnuc.process_nuc_chunk(pos, tlen, s, e, params, bias_track=bt, bias_track_start=span[0], fit=False)
        a, b = int(pb.out_off[j]), int(pb.out_off[j + 1])
        scale = max(float(np.abs(r["nuc_signal"]).max()), float(np.abs(r["bias"]).max()), 1e-300)
        for key, okey in (("background", "bias"), ("norm_signal", "norm_signal"), ("smoothed", "smoothed")):
            err = float(np.abs(out[key][a:b] - r[okey]).max()) / scale
            worst = max(worst, err)
            assert err <= 1e-5, (size, j, key, err)
        assert np.array_equal(out["nuc_cov"][a:b], r["nuc_cov"])
    print("VMat %dx%d: worst |error| / signal scale = %.2e (bar 1e-5)" % (size, size, worst))


def test_batches_in_flight_give_identical_results(eng):
    """Three batches in flight (own streams, passes chained first-in first-out, copies overlapping the next pass) return
    what each returns when run alone."""
    from nucleoatac_b200 import synth
    wl = synth.Workload(251, 251)
    _shrink_vmat(eng, wl)
    wl.configure(eng, use_bias=True, xcor_mode=2)
    pbs = [synth.make_batch(20 + 7 * i, 4 + i) for i in range(3)]
    alone = [(eng.process_nuc(pb), eng.process_occ(pb, raw=False)) for pb in pbs]
    hs, outs = [], []
    for rep in range(2):           # second round recycles the handles while the first round's copies may still be running
        for i, pb in enumerate(pbs):
            h = eng.upload(pb, hs[i] if rep else None)
            if not rep:
                hs.append(h)
                outs.append((eng.nuc_alloc(pb, track_dtype=np.float32), eng.occ_alloc(pb, raw=False, track_dtype=np.float32)))
            eng.nuc_run(h)
            eng.nuc_download(h, outs[i][0])
            eng.occ_run(h)
            eng.occ_download(h, outs[i][1])
    for h in hs:
        eng.sync(h)
        eng.free_batch(h)
    def valid(out, off, count, key):   # the filled slots of a capacity-sized table
        return np.concatenate([out[key][int(o):int(o) + int(c)] for o, c in zip(out[off][:-1], out[count])] + [np.zeros(0, out[key].dtype)])
    for (n1, o1), (n2, o2) in zip(alone, outs):
        for k in ("norm_signal", "smoothed"):
            assert np.array_equal(n1[k].astype(np.float32), n2[k], equal_nan=True), k
        assert np.array_equal(n1["cand_count"], n2["cand_count"])
        for k in ("cand_pos", "cand_flag", "cand_z"):
            assert np.array_equal(valid(n1, "cand_off", "cand_count", k), valid(n2, "cand_off", "cand_count", k), equal_nan=True), k
        for k in ("smoothed_vals", "smoothed_lower", "smoothed_upper"):
            assert np.array_equal(o1[k].astype(np.float32), o2[k], equal_nan=True), k
        assert np.array_equal(o1["peak_count"], o2["peak_count"]) and np.array_equal(o1["nuc_dist"], o2["nuc_dist"])
        assert np.array_equal(valid(o1, "peak_off", "peak_count", "peak_pos"), valid(o2, "peak_off", "peak_count", "peak_pos"))


@pytest.mark.parametrize("size", [(251, 251), (201, 151), (146, 121)])
def test_both_tensor_core_kernels_agree(eng, size, monkeypatch):
    """The two tcgen05 forms of the background cross-correlation (NucleosomeCalling.py:60-63) -- hi Hankel operand in tensor
    memory (k_nuc_bx_ts, the default when G fits in shared memory) and both operands in shared memory (k_nuc_bx_tc,
    NB200_TC_TS=0) -- against the exact fp64 CUDA-core kernel on ragged chunks (partial x-tiles, an empty chunk): each within
    1e-5 of the signal scale, each bit-reproducible run to run, the same candidates from all three."""
    from nucleoatac_b200 import synth
    from nucleoatac_b200.engine import PackedBatch
    R, W = size
    wl = synth.Workload(R, W)
    _shrink_vmat(eng, wl)
    margin = W + R // 2 + 40
    specs = [(3, 700), (5, 5000), (8, 129), (11, 2049), (12, 10000)]
    chunks = [synth.make_chunk(k, length=L, seq_margin=margin) for k, L in specs]
    s, e, pos, tlen, seq, s0 = chunks[2]
    chunks[2] = (s, e, pos[:0], tlen[:0], seq, s0)   # a chunk without reads
    pb = PackedBatch.from_chunks(chunks)
    wl.configure(eng, use_bias=True, xcor_mode=1)
    exact = eng.process_nuc(pb)
    scale = float(np.abs(exact["background"]).max())
    wl.configure(eng, use_bias=True, xcor_mode=2)
    outs = {}
    for ts in ("1", "0"):
        monkeypatch.setenv("NB200_TC_TS", ts)
        eng.profile_reset()
        a = eng.process_nuc(pb)
        b = eng.process_nuc(pb)
        assert eng.profile_report().get("k_nuc_bx_ts" if ts == "1" else "k_nuc_bx_tc", (0, 0.0))[0] == 2
        assert np.array_equal(a["background"], b["background"]), "not reproducible (NB200_TC_TS=%s)" % ts
        err = float(np.abs(a["background"] - exact["background"]).max()) / scale
        assert err <= 1e-5, (size, ts, err)
        assert np.array_equal(a["cand_count"], exact["cand_count"])
        for j in range(len(chunks)):
            co, cn = int(a["cand_off"][j]), int(a["cand_count"][j])
            assert np.array_equal(a["cand_pos"][co:co + cn], exact["cand_pos"][co:co + cn])
            assert np.array_equal(a["cand_flag"][co:co + cn] & 4, exact["cand_flag"][co:co + cn] & 4)
        outs[ts] = (a, err)
    d = float(np.abs(outs["1"][0]["background"] - outs["0"][0]["background"]).max()) / scale
    print("VMat %dx%d: |error| / scale: tensor-memory form %.2e, shared-memory form %.2e, difference %.2e" % (R, W, outs["1"][1], outs["0"][1], d))


@pytest.mark.parametrize("shape", [(60, 41, 0), (121, 201, 0), (200, 101, 30), (33, 251, 100), (251, 101, 0), (97, 63, 150), (250, 249, 1),
                                   (16, 17, 120), (280, 121, 0)])
def test_tensor_core_vmat_shapes(eng, shape, monkeypatch):
    """VMat shapes off the beaten path (rows != columns, lower > 0 -- without the single-tap size-1 row --, a few rows only, more
    than two or fewer than two slabs of a taps): both tcgen05 kernels against the exact fp64 kernel, 1e-5 of the signal scale.
    The block plan of k_nuc_bx_ts (slabs, the split of the tensor-memory operand, which slab waits for its second part)
    derives from the shape."""
    from nucleoatac_b200 import synth
    from nucleoatac_b200.engine import PackedBatch
    R, W, lower = shape
    wl = synth.Workload(R, W, upper=lower + R, lower=lower)
    _shrink_vmat(eng, wl)
    margin = W + (lower + R) // 2 + 40
    chunks = [synth.make_chunk(k, length=L, seq_margin=margin) for k, L in ((21, 1500), (22, 513), (23, 3100))]
    pb = PackedBatch.from_chunks(chunks)
    wl.configure(eng, use_bias=True, xcor_mode=1)
    exact = eng.process_nuc(pb)["background"]
    scale = float(np.abs(exact).max())
    assert scale > 0.0, "the fragment-size distribution has no mass on this VMat's rows: nothing to compare"
    wl.configure(eng, use_bias=True, xcor_mode=2)
    errs = []
    for ts in ("1", "0"):
        monkeypatch.setenv("NB200_TC_TS", ts)
        got = eng.process_nuc(pb)["background"]
        errs.append(float(np.abs(got - exact).max()) / scale)
        assert errs[-1] <= 1e-5, (shape, ts, errs[-1])
    print("VMat %dx%d, sizes from %d: |error| / scale %.2e (tensor-memory form), %.2e (shared-memory form)" % (R, W, lower, errs[0], errs[1]))
