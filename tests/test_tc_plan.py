"""Block plan of the tcgen05 background kernel (BiasTrack.calculateBackgroundSignal, nucleoatac/NucleosomeCalling.py:60-63),
computed on the host by nb200_tc_plan_describe -- no GPU.  The kernel's synchronisation protocol leans on properties of the
table (which block frees part 1 of the tensor-memory operand, which one is the first to need part 2, the untrimmed first
block of every slab): they are checked here over VMat shapes, including ones no GPU test runs."""
import ctypes as C

import numpy as np
import pytest

from nucleoatac_b200 import _lib as L
from nucleoatac_b200 import synth

TS_N = 128


def describe(vmat, lower, upper, sizes, max_blocks=256):
    lib = L.load()
    v = np.ascontiguousarray(vmat, dtype=np.float64)
    f = np.ascontiguousarray(sizes, dtype=np.float64)
    stats = np.zeros(16, dtype=np.int32)
    blocks = np.zeros(4 * max_blocks, dtype=np.int32)
    st = lib.nb200_tc_plan_describe(L.ptr(v, C.c_double), lower, upper, v.shape[1], L.ptr(f, C.c_double), len(f), L.ptr(stats, C.c_int32),
                                    L.ptr(blocks, C.c_int32), max_blocks)
    assert st == 0
    names = ("ok", "slabs", "blocks", "c_split", "c_end", "q_need2", "t_need2", "rank_bytes", "NA", "NB", "mma_cols", "nonzero_cols",
             "uncovered", "has_row1")
    d = dict(zip(names, (int(x) for x in stats)))
    d["table"] = blocks[:4 * min(d["blocks"], max_blocks)].reshape(-1, 4)
    return d


def check_plan(d):
    assert d["uncovered"] == 0, "a non-zero of G lies outside every block"
    t = d["table"]
    kb, flags = t[:, 0] & 0xffff, t[:, 0] >> 16
    n_lo, n_t = t[:, 1], ((t[:, 2] >> 17) & 0x3f) * 8
    rows_per_cta = t[:, 3] >> 16
    assert np.all(n_t % 16 == 0) and np.all(n_t >= 16) and np.all(n_lo % 16 == 0) and np.all(n_lo + n_t <= TS_N)
    assert np.array_equal(rows_per_cta, n_t // 2)                       # B is split across the CTA pair
    assert d["c_split"] % 32 == 0 and d["c_end"] % 32 == 0 and 0 <= d["c_split"] <= d["c_end"] <= 256
    assert d["c_end"] * 2 >= d["NB"]                                    # the operand holds every b tap
    assert int(n_t.sum()) == d["mma_cols"] and d["rank_bytes"] == 32 * d["mma_cols"]
    # image offsets: blocks back to back, hi + lo images of the CTA's rows
    off = (t[:, 3] & 0xffff) * 16
    assert np.array_equal(off, np.concatenate(([0], np.cumsum(rows_per_cta[:-1] * 64))))
    # flags: exactly one "part 1 free after this block" and one "first reader of part 2"
    p1 = kb * 8 < d["c_split"]
    assert int((flags & 1).sum()) == 1 and int(((flags >> 2) & 1).sum()) == 1
    i_free = int(np.flatnonzero(flags & 1)[0]); i_need = int(np.flatnonzero((flags >> 2) & 1)[0])
    if p1.any():
        assert i_free == int(np.flatnonzero(p1)[-1]), "part 1 is released before its last reader"
    if (~p1).any():
        assert i_need == int(np.flatnonzero(~p1)[0]), "a block reads part 2 before the wait for it"
        assert not (~p1[:i_need]).any()
    return kb, flags, n_lo, n_t, i_need


def slab_bounds(d, n_t, n_lo):
    """Slab starts = the untrimmed blocks (the first block of a slab initialises all TS_N accumulator columns)."""
    starts = [i for i in range(len(n_t)) if n_lo[i] == 0 and n_t[i] == TS_N]
    return starts


@pytest.mark.parametrize("R,W,lower", [(251, 251, 0), (201, 201, 0), (151, 151, 0), (101, 101, 0), (146, 121, 0), (130, 121, 115), (201, 151, 0),
                                       (60, 41, 0), (121, 201, 0), (200, 101, 30), (33, 251, 100), (251, 101, 0), (97, 63, 150), (250, 249, 1),
                                       (16, 17, 120), (280, 121, 0), (240, 257, 10), (75, 301, 60), (251, 51, 0), (190, 190, 35)])
def test_plan_invariants(R, W, lower):
    wl = synth.Workload(R, W, upper=lower + R, lower=lower)
    d = describe(wl.vmat, wl.v_lower, wl.v_upper, wl.fragmentsizes)
    assert d["blocks"] > 0 and d["slabs"] == -(-d["NA"] // TS_N)
    kb, flags, n_lo, n_t, i_need = check_plan(d)
    # every slab opens with an untrimmed block, and the slab / position of the first part-2 reader are what the kernel is told
    starts = slab_bounds(d, n_t, n_lo)
    assert len(starts) >= d["slabs"] and starts[0] == 0
    if d["ok"]:
        assert d["blocks"] <= 192 and d["slabs"] <= 8 and d["rank_bytes"] < 200 * 1024
    # padding: issued columns never less than the non-zero (row, K block) pairs, and within 1.6x of them
    assert d["mma_cols"] >= d["nonzero_cols"] and d["mma_cols"] <= 1.6 * d["nonzero_cols"] + 2 * TS_N * d["slabs"]


def test_plan_of_the_benchmark_shape():
    """251 x 251 (BASELINE configs[1]): three slabs, 56 blocks, the operand split 64 + 128 columns, part 2 first read in slab 0
    after its eight part-1 blocks, 4592 MMA columns per x-tile and pass for 3938 non-zero ones."""
    wl = synth.Workload(251, 251)
    d = describe(wl.vmat, wl.v_lower, wl.v_upper, wl.fragmentsizes)
    assert (d["ok"], d["slabs"], d["blocks"], d["c_split"], d["c_end"], d["q_need2"], d["t_need2"]) == (1, 3, 56, 64, 192, 0, 8)
    assert d["mma_cols"] == 4592 and d["NA"] == 376 and d["NB"] == 376 and d["has_row1"] == 1
    check_plan(d)


def test_plan_sparse_and_degenerate_vmats():
    """A VMat with a single non-zero row, one whose mass sits in two corners, and one with only the single-tap size 1 (nothing
    for the tensor core: every statistic 0)."""
    sizes = np.full(300, 1.0 / 300)
    v = np.zeros((251, 251)); v[140, :] = 1.0
    check_plan(describe(v, 0, 251, sizes))
    v = np.zeros((251, 251)); v[:8, :8] = 1.0; v[-8:, -8:] = 2.0
    d = describe(v, 0, 251, sizes)
    check_plan(d)
    assert d["mma_cols"] < 1200                                          # the empty middle is neither stored nor multiplied
    d = describe(np.ones((1, 11)), 1, 2, sizes)
    assert d["blocks"] == 0 and d["ok"] == 0
    rng = np.random.default_rng(5)
    for _ in range(6):                                                   # random sparse VMats
        R, W = int(rng.integers(20, 260)), 2 * int(rng.integers(10, 130)) + 1
        v = rng.random((R, W)) * (rng.random((R, W)) < 0.05)
        if not v.any():
            v[R // 2, W // 2] = 1.0
        lower = int(rng.integers(0, 40))
        check_plan(describe(v, lower, lower + R, np.full(lower + R, 1.0 / (lower + R))))
