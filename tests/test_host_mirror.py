"""CPU tests of the host-side mirror of the reference API (no device calls): BED/VMat/sizes IO, the bedgraph
writer against the oracle's literal restatement, BGZF/BAM/FASTA readers and writers, CLI flags."""
import gzip
import io
import os

import numpy as np
import pytest

from nucleoatac_b200 import hostio
from nucleoatac_b200.chunk import Chunk, ChunkList
from oracle import refalgo as ra


def test_chunklist_read_slop_merge_split(tmp_path):
    bed = tmp_path / "a.bed"
    bed.write_text("chr1\t100\t300\nchr1\t350\t500\nchr1\t5000\t5100\nchr2\t10\t90\nchrX\t1\t2\n")
    chrs = {"chr1": 6000, "chr2": 1000}
    with pytest.warns(UserWarning):
        cl = ChunkList.read(str(bed), chromDict=chrs, min_offset=50)
    assert [(c.chrom, c.start, c.end) for c in cl] == [("chr1", 100, 300), ("chr1", 350, 500), ("chr1", 5000, 5100), ("chr2", 50, 90)]
    ref = ra.merge_chunks(ra.slop_chunks(ra.read_bed_chunks(str(bed), chrs, min_offset=50), chrs, 60, 60))
    cl.slop(chrs, up=60, down=60)
    cl.merge()
    assert [[c.chrom, c.start, c.end] for c in cl] == ref
    assert [len(x) for x in cl.split(items=2)] == [2, 1]
    assert Chunk("c", 1, 5).asBed() == "c\t1\t5\t1\tregion\t*"
    c = Chunk("chr1", 10, 31)
    c.center()
    assert (c.start, c.end) == (20, 21)


def test_write_track_matches_reference_semantics():
    from nucleoatac_b200.tracks import Track
    rng = np.random.RandomState(0)
    for trial in range(30):
        n = 60
        vals = np.round(rng.rand(n), 1)
        vals[rng.rand(n) < 0.25] = np.nan
        vals[rng.rand(n) < 0.2] = 0.0
        if trial % 3 == 0:
            vals[-3:] = np.nan
        for wz in (True, False):
            h = io.StringIO()
            Track("chrT", 1000, 1000 + n).write_track(h, vals=vals, write_zero=wz)
            assert h.getvalue() == ra.write_track("chrT", 1000, 1000 + n, vals, write_zero=wz), (trial, wz)


def test_vmat_and_sizes_io(tmp_path, example):
    from nucleoatac_b200.fragmentsizes import FragmentSizes
    from nucleoatac_b200.VMat import VMat, VMat_Error
    mat, lo, hi = example.vmat
    p = str(tmp_path / "x.VMat")
    VMat(mat, lo, hi).save(p)
    v = VMat.open(p)
    assert (v.lower, v.upper, v.w) == (lo, hi, mat.shape[1] // 2)
    np.testing.assert_allclose(v.mat, mat, rtol=1e-11)
    m2, l2, u2 = ra.read_vmat(p)
    assert np.array_equal(m2, v.mat)
    with pytest.raises(VMat_Error):
        VMat(mat, lo, hi + 1)
    # vprocess steps against the oracle restatement of pyatac/VMat.py
    from oracle import refvmat
    raw = example.z["std_vplot"].astype(np.float64)
    v = VMat(np.array(raw), int(example.z["std_vplot_lower"]), int(example.z["std_vplot_upper"]))
    v.trim(105, 251, 60)
    v.symmetrize()
    v.smooth(0.75)
    v.norm()
    np.testing.assert_allclose(v.mat, refvmat.vprocess(raw, 0, 105, 251, 60), rtol=1e-12)
    s = str(tmp_path / "s.txt")
    FragmentSizes(0, 251, vals=example.fragmentsizes).save(s)
    f = FragmentSizes.open(s)
    np.testing.assert_allclose(f.get(0, 251), example.fragmentsizes, rtol=1e-11)
    assert f.get(size=100) == f.vals[100]


def test_bgzf_bam_fasta_roundtrip(tmp_path):
    from tests.synthfiles import write_bam
    rng = np.random.RandomState(1)
    pos = np.sort(rng.randint(100, 200000, 5000))
    tlen = rng.randint(40, 400, 5000)
    path = str(tmp_path / "t.bam")
    write_bam(path, {"chrA": 250000, "chrB": 1000}, [(0, int(p), int(t)) for p, t in zip(pos, tlen)])
    assert gzip.open(path).read(4) == b"BAM\x01"
    bam = hostio.BamFile(path)
    assert bam.references == ["chrA", "chrB"] and bam.lengths == [250000, 1000]
    p, t = bam.fetch_fragments("chrA", 50000, 60000)
    sel = (pos >= 50000) & (pos < 60000)
    keep = p >= 50000
    assert np.array_equal(p[keep], pos[sel]) and np.array_equal(t[keep], tlen[sel])  # forward proper-pair mates only
    assert len(bam.fetch_fragments("chrB", 0, 1000)[0]) == 0
    fa = tmp_path / "g.fa"
    seq = "".join(rng.choice(list("ACGTN"), 1000))
    fa.write_text(">c1 desc\n" + "\n".join(seq[i:i + 70] for i in range(0, 1000, 70)) + "\n>c2\nACGT\n")
    f = hostio.FastaFile(str(fa))
    assert f.references == ["c1", "c2"] and f.lengths == [1000, 4]
    assert f.fetch("c1", 65, 215) == seq[65:215] and f.fetch("c2", 0, 10) == "ACGT"
    bg = tmp_path / "t.bedgraph"
    bg.write_text("c1\t0\t5\t1.5\nc1\t5\t9\t2.0\nc1\t20\t30\t3.0\n")
    hostio.bgzip_file(str(bg), str(bg) + ".gz")
    r = hostio.BedGraphReader(str(bg) + ".gz")
    out = r.read("c1", 3, 25)
    assert out[0] == 1.5 and out[2] == 2.0 and np.isnan(out[10]) and out[-1] == 3.0


def test_cli_flags_match_reference():
    from nucleoatac_b200.cli import build_parser
    a = build_parser().parse_args("occ --bed b --bam m --out o".split())
    assert (a.upper, a.flank, a.min_occ, a.nuc_sep, a.confidence_interval, a.step, a.pwm, a.cores) == (251, 60, 0.1, 120, 0.9, 5, "Human", 1)
    a = build_parser().parse_args("nfr --bed b --occ_track o.occ.bedgraph.gz --calls c.bed.gz --bam m".split())
    assert (a.max_occ, a.max_occ_upper, a.pwm, a.ins_track, a.fasta, a.out) == (0.1, 0.25, "Human", None, None, None)
    a = build_parser().parse_args("nuc --bed b --bam m --out o --vmat v --not_atac --write_all".split())
    assert (a.min_z, a.min_lr, a.nuc_sep, a.redundant_sep, a.sd, a.atac, a.write_all) == (3, 0, 120, 25, 10, False, True)


def test_pwm_bundled(example):
    from nucleoatac_b200.bias import PWM
    p = PWM.open("Human")
    assert (p.up, p.down, p.nucleotides) == (10, 10, ["A", "C", "G", "T"])
    np.testing.assert_allclose(p.mat, example.pwm, rtol=1e-12)


def test_vprocess_reproduces_shipped_vmat(tmp_path, example, golden):
    """`nucleoatac vprocess` on the bundled S. cer V-plot with the example's nuc_dist reproduces the VMat the
    reference shipped in example_results (what `nucleoatac run` wires from occ to nuc, cli.py:38-41)."""
    from nucleoatac_b200.cli import nucleoatac_main
    from nucleoatac_b200.fragmentsizes import FragmentSizes
    from nucleoatac_b200.VMat import VMat
    sizes = str(tmp_path / "nuc_dist.txt")
    FragmentSizes(0, 251, vals=golden["nuc_dist"]).save(sizes)
    out = str(tmp_path / "ex")
    assert nucleoatac_main(["vprocess", "--sizes", sizes, "--out", out]) == 0
    v = VMat.open(out + ".VMat")
    mat, lo, hi = example.vmat
    assert (v.lower, v.upper, v.mat.shape) == (lo, hi, mat.shape)
    np.testing.assert_allclose(v.mat, mat, rtol=2e-11, atol=1e-15)


def test_tabix_index_roundtrip_and_reference_layout(tmp_path):
    """.tbi writer: region queries through our own index return exactly the overlapping rows; when the reference tree
    is present (build container) the index built for its shipped bedgraph has the same bins / chunks / linear
    offsets as the .tbi htslib wrote for that file."""
    import gzip as gz
    import struct
    rng = np.random.RandomState(5)
    rows, pos = [], 0
    for chrom in ("chrA", "chrB"):
        pos = 100
        for _ in range(30000):
            ln = int(rng.randint(1, 40))
            rows.append((chrom, pos, pos + ln, round(float(rng.rand()), 6)))
            pos += ln + int(rng.randint(0, 30))
    plain = tmp_path / "t.bedgraph"
    plain.write_text("".join("%s\t%d\t%d\t%s\n" % r for r in rows))
    hostio.bgzip_file(str(plain), str(plain) + ".gz")
    hostio.tabix_index(str(plain) + ".gz")
    tb = hostio.TabixFile(str(plain) + ".gz")
    for chrom, s, e in (("chrA", 5000, 5600), ("chrB", 400000, 401000), ("chrA", 0, 150), ("chrB", 10 ** 7, 10 ** 7 + 5)):
        got = [(r[0], int(r[1]), int(r[2])) for r in tb.fetch(chrom, s, e)]
        exp = [(c, a, b) for (c, a, b, _v) in rows if c == chrom and b > s and a < e]
        assert got == exp, (chrom, s, e)
    ref = "/root/reference/example/example_results/example.occ.bedgraph.gz"
    if not os.path.exists(ref):
        return
    import shutil
    mine = str(tmp_path / "ref.bedgraph.gz")
    shutil.copy(ref, mine)
    hostio.tabix_index(mine)
    a, b = hostio.TabixFile(mine), hostio.TabixFile.__new__(hostio.TabixFile)
    shutil.copy(ref + ".tbi", mine + ".tbi")
    b.__init__(mine)
    assert a.names == b.names and (a.fmt, a.sc, a.bc, a.ec) == (b.fmt, b.sc, b.bc, b.ec)
    for t in range(len(a.names)):
        ref_bins = {k: v for k, v in b.bins[t].items() if k != 37450}   # htslib's metadata pseudo-bin
        assert a.bins[t] == ref_bins, a.names[t]
        assert a.linear[t] == b.linear[t], a.names[t]


def test_native_bgzip_tabix_equals_python(tmp_path):
    """The native one-pass BGZF + .tbi writer gives the same decompressed bytes and the same index content as the
    Python BgzfWriter + tabix_index pair (which is pinned on the htslib-written index above)."""
    import gzip as gz
    rng = np.random.RandomState(11)
    lines, pos = [], 0
    for chrom in ("chr1", "chr10", "chr2"):
        pos = 5
        for _ in range(60000):
            ln = int(rng.randint(1, 60))
            lines.append("%s\t%d\t%d\t%s\n" % (chrom, pos, pos + ln, repr(round(float(rng.rand()), 9))))
            pos += ln + int(rng.randint(0, 9))
    text = "".join(lines)
    plain = str(tmp_path / "n.bedgraph")
    open(plain, "w").write(text)
    hostio.bgzip_tabix(plain, plain + ".native.gz", threads=4)
    hostio.bgzip_file(plain, plain + ".py.gz")
    hostio.tabix_index(plain + ".py.gz")
    assert gz.open(plain + ".native.gz", "rt").read() == text
    a, b = hostio.TabixFile(plain + ".native.gz"), hostio.TabixFile(plain + ".py.gz")
    assert a.names == b.names and (a.fmt, a.sc, a.bc, a.ec) == (b.fmt, b.sc, b.bc, b.ec)
    # block boundaries coincide (same 0xff00 blocking) but compressed sizes may differ by deflate settings: compare
    # what the index points at rather than raw offsets
    for chrom, s, e in (("chr1", 100000, 100500), ("chr10", 5, 40), ("chr2", 1500000, 1500100), ("chr2", 0, 10 ** 9)):
        assert a.fetch(chrom, s, e) == b.fetch(chrom, s, e), (chrom, s, e)
    for t in range(len(a.names)):
        assert sorted(a.bins[t].keys()) == sorted(b.bins[t].keys())
        assert [len(v) for _, v in sorted(a.bins[t].items())] == [len(v) for _, v in sorted(b.bins[t].items())]
        assert len(a.linear[t]) == len(b.linear[t])
    ref = "/root/reference/example/example_results/example.nucleoatac_signal.bedgraph.gz"
    if os.path.exists(ref):  # build container only: same bins / chunks / linear offsets as the htslib-written index
        import shutil
        p2 = str(tmp_path / "r.bedgraph")
        open(p2, "wb").write(gz.open(ref, "rb").read())
        hostio.bgzip_tabix(p2, p2 + ".gz", threads=3)
        mine = hostio.TabixFile(p2 + ".gz")
        shutil.copy(ref, p2 + ".ref.gz")
        shutil.copy(ref + ".tbi", p2 + ".ref.gz.tbi")
        theirs = hostio.TabixFile(p2 + ".ref.gz")
        assert mine.names == theirs.names
        for t in range(len(mine.names)):
            assert mine.bins[t] == {k: v for k, v in theirs.bins[t].items() if k != 37450}
            assert mine.linear[t] == theirs.linear[t]


def test_merge_reproduces_shipped_nucmap(tmp_path, golden):
    """`nucleoatac merge` on the reference's shipped occpeaks + nucpos gives its shipped nucmap_combined, byte for byte."""
    import gzip as gz
    from nucleoatac_b200.cli import nucleoatac_main
    occ, nuc = str(tmp_path / "e.occpeaks.bed"), str(tmp_path / "e.nucpos.bed")
    open(occ, "w").write(str(golden["occpeaks_text"]))
    open(nuc, "w").write(str(golden["nucpos_text"]))
    out = str(tmp_path / "e")
    assert nucleoatac_main(["merge", "--occpeaks", occ, "--nucpos", nuc, "--out", out]) == 0
    assert gz.open(out + ".nucmap_combined.bed.gz", "rt").read() == str(golden["nucmap_combined_text"])
    assert os.path.exists(out + ".nucmap_combined.bed.gz.tbi")


def test_native_bam_reader_matches_python(tmp_path):
    """nb200_bam_fetch_many (native, threaded) against the pure-Python region reader on an indexed synthetic BAM with
    several references (one without reads): same reads, same order, for regions inside, across and past the data, an
    unknown reference, and every thread count; index_bam's linear index against a scan of the whole file."""
    import numpy as np
    from nucleoatac_b200 import hostio
    from tests.synthfiles import write_bam
    rng = np.random.default_rng(5)
    sizes = {"chrA": 300000, "chrEmpty": 50000, "chrB": 120000}
    reads = []
    for tid, name in ((0, "chrA"), (2, "chrB")):
        n = 20000 if tid == 0 else 3000
        pos = np.sort(rng.integers(0, sizes[name] - 600, n))
        tl = rng.integers(30, 500, n)
        reads += [(tid, int(p), int(t)) for p, t in zip(pos, tl)]
    bam = str(tmp_path / "multi.bam")
    write_bam(bam, sizes, reads)
    hostio.index_bam(bam)
    b = hostio.BamFile(bam)
    assert b._index is not None and len(b._index) == 3 and len(b._index[1]) == 0
    regions = [("chrA", 0, 1000), ("chrA", 16384, 16385), ("chrA", 299000, 400000), ("chrEmpty", 0, 50000), ("chrB", 0, 120000),
               ("chrNope", 5, 10), ("chrB", 119990, 119999), ("chrA", 150000, 150001)]
    for _ in range(40):
        name = ("chrA", "chrB")[int(rng.integers(0, 2))]
        s0 = int(rng.integers(-3000, sizes[name]))
        regions.append((name, s0, s0 + int(rng.integers(1, 60000))))
    expect = [b._fetch_indexed(b._tid[c], max(0, s0), e) if c in b._tid else (np.zeros(0, np.int32), np.zeros(0, np.int32))
              for c, s0, e in regions]
    assert sum(len(p) for p, _ in expect) > 10000
    for threads in (1, 3, 16):
        off, pos, tlen = b.fetch_fragments_many([(c, max(0, s0), e) for c, s0, e in regions], threads=threads)
        assert off[0] == 0 and len(off) == len(regions) + 1 and off[-1] == len(pos) == len(tlen)
        for i, (p, t) in enumerate(expect):
            assert np.array_equal(pos[off[i]:off[i + 1]], p) and np.array_equal(tlen[off[i]:off[i + 1]], t), (threads, regions[i])
    # every forward proper-pair read of a reference is found through the index
    allp, _ = b._fetch_indexed(0, 0, sizes["chrA"])
    assert len(allp) == 20000 and np.all(np.diff(allp) >= 0)
    off, pos, tlen = b.fetch_fragments_many([])
    assert len(off) == 1 and len(pos) == 0
    # a BAM cut in the middle of a block (or of a record) is an error, not a short read list; the untrusted uncompressed
    # size in a block trailer is bounded
    raw = open(bam, "rb").read()
    cut = str(tmp_path / "cut.bam")
    open(cut, "wb").write(raw[:len(raw) // 2 + 13])
    import shutil
    shutil.copy(bam + ".bai", cut + ".bai")
    bc = hostio.BamFile(cut)
    with pytest.raises(IOError, match="truncated"):
        bc.fetch_fragments_many([("chrA", 0, sizes["chrA"]), ("chrB", 0, sizes["chrB"])], threads=2)
    bad = bytearray(raw)
    bsize = (bad[16] | (bad[17] << 8)) + 1                      # first block: patch ISIZE (last 4 bytes) to 1 GiB
    bad[bsize - 4:bsize] = (1 << 30).to_bytes(4, "little")
    evil = str(tmp_path / "evil.bam")
    open(evil, "wb").write(bytes(bad))
    shutil.copy(bam + ".bai", evil + ".bai")
    with pytest.raises(IOError, match="impossible uncompressed size|inflate"):
        hostio.BamFile(evil).fetch_fragments_many([("chrA", 0, 1000)], threads=1)


def test_merge_sorts_unsorted_bed():
    """pyatac/chunk.py:109-125: merge() sorts first when the list is not sorted (an unsorted BED must not lose bases)."""
    cl = ChunkList(Chunk("chr1", 500, 900), Chunk("chr1", 100, 600), Chunk("chr1", 550, 700), Chunk("chr0", 5, 9))
    cl.merge()
    assert [(c.chrom, c.start, c.end) for c in cl] == [("chr0", 5, 9), ("chr1", 100, 900)]
    assert cl.isSorted()


def test_modelNFR_reproduces_shipped_occ_fit(example):
    """FragmentMixDistribution.modelNFR (nucleoatac/Occupancy.py:29-66) against rows 2-3 of the reference's shipped
    example_results/example.occ_fit.txt (nuc_fit, nfr_fit), from the shipped fragment sizes (row 1)."""
    from nucleoatac_b200.fragmentsizes import FragmentSizes
    from nucleoatac_b200.Occupancy import FragmentMixDistribution
    fd = FragmentMixDistribution(0, 251)
    fd.fragmentsizes = FragmentSizes(0, 251, vals=np.array(example.occ_fit[0]))
    fd.modelNFR()
    np.testing.assert_allclose(fd.nuc_fit.get(0, 251), example.occ_fit[1], rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(fd.nfr_fit.get(0, 251), example.occ_fit[2], rtol=1e-9, atol=1e-15)


def test_synthetic_chunks_stay_inside_int32():
    """BASELINE configs[3] (500 000 x 10 kb chunks, 8 GPUs): per-contig coordinates like pyatac/chunk.py:132-175, so chunk
    499 999 (and bench.py's largest index at N=8) is generatable, and an out-of-range coordinate raises instead of wrapping."""
    from nucleoatac_b200 import synth
    from nucleoatac_b200.engine import PackedBatch
    for k in (178955, 178956, 199999, 399999, 499999, 28 * 2000 * 8 + 7):
        s, e, pos, tlen, seq, s0 = synth.make_chunk(k, with_seq=(k == 499999))
        assert 0 < s < e < 2 ** 31 and e - s == synth.CHUNK_LEN
        assert pos.dtype == np.int32 and len(pos) > 1500 and pos.min() > s - 2000 and pos.max() < e + 2000
        assert np.all(np.diff(pos) >= 0)
    assert synth.chunk_contig(499999) == "synth4" and synth.chunk_contig(0) == "synth0"
    # same reads as chunk k on the first contig would have: the generator depends on (seed, k) only through rng and offset
    a, b = synth.make_chunk(5, with_seq=False), synth.make_chunk(100005, with_seq=False)
    assert a[0] == b[0] and len(a[2]) != 0 and not np.array_equal(a[2][:50], b[2][:50])
    pb = PackedBatch.from_chunks([synth.make_chunk(k, with_seq=False) for k in (499998, 499999)])
    assert pb.total_len == 20000
    with pytest.raises(OverflowError):
        PackedBatch(np.array([2 ** 31 + 5]), np.array([2 ** 31 + 10]), [0, 0], np.zeros(0, np.int32), np.zeros(0, np.int32))
    with pytest.raises(OverflowError):
        synth._checked_i32(np.array([2 ** 31]), "x")


def test_sharded_run_without_process_group_is_an_error(monkeypatch):
    """--rank/--world given by hand without a rendezvous: the reductions refuse to return per-shard partial sums."""
    from nucleoatac_b200 import dist
    for k in ("MASTER_ADDR", "RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    x = np.arange(4.0)
    assert dist.allreduce_sum(x, 1) is x and dist.allreduce_sum(x) is x
    with pytest.raises(RuntimeError, match="process group"):
        dist.allreduce_sum(x, 2)
    with pytest.raises(RuntimeError, match="process group"):
        dist.barrier(2)


def test_cli_run_propagates_shard_and_device(monkeypatch, tmp_path):
    """`nucleoatac run` hands rank / world / device / batch / xcor_mode to every sharded step and keeps the host-only
    steps (vprocess, merge) on rank 0."""
    from nucleoatac_b200 import cli, dist
    import nucleoatac_b200.merge as M, nucleoatac_b200.run_nfr as NF, nucleoatac_b200.run_nuc as N, nucleoatac_b200.run_occ as O, \
        nucleoatac_b200.run_vprocess as V
    seen = []
    monkeypatch.setattr(O, "run_occ", lambda a: seen.append(("occ", a.rank, a.world, a.device, a.batch)))
    monkeypatch.setattr(N, "run_nuc", lambda a: seen.append(("nuc", a.rank, a.world, a.device, a.batch, a.xcor_mode)))
    monkeypatch.setattr(NF, "run_nfr", lambda a: seen.append(("nfr", a.rank, a.world, a.device, a.batch)))
    monkeypatch.setattr(V, "run_vprocess", lambda a: seen.append(("vprocess",)))
    monkeypatch.setattr(M, "run_merge", lambda a: seen.append(("merge",)))
    monkeypatch.setattr(dist, "barrier", lambda world=None: seen.append(("barrier", world)))
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    base = ["run", "--bed", "a.bed", "--bam", "a.bam", "--fasta", "a.fa", "--out", str(tmp_path / "o")]
    cli.nucleoatac_main(base + ["--rank", "1", "--world", "4", "--device", "1", "--batch", "64", "--xcor_mode", "1"])
    assert seen == [("occ", 1, 4, 1, 64), ("barrier", 4), ("nuc", 1, 4, 1, 64, 1), ("barrier", 4), ("nfr", 1, 4, 1, 64)]
    seen.clear()
    cli.nucleoatac_main(base)
    assert [s[0] for s in seen] == ["occ", "vprocess", "barrier", "nuc", "merge", "barrier", "nfr"]


def test_bgzip_tabix_streams_in_waves(tmp_path):
    """nb200_bgzip_tabix streams the text in waves of BGZF blocks: a file of several waves (1 thread: 4 MB per wave), with rows
    straddling the wave boundaries and no newline after the last row, gives the same .gz and .tbi as one wave (8 threads)."""
    rng = np.random.RandomState(3)
    rows, pos = [], 0
    for k in range(260000):
        pos += int(rng.randint(1, 40))
        rows.append("chr%d\t%d\t%d\t%s" % (1 + k // 130000, pos % 50000000 + (k // 130000) * 0, pos % 50000000 + 1, repr(float(rng.rand()))))
        if k == 129999:
            pos = 0
    text = "\n".join(rows)          # no trailing newline
    plain = tmp_path / "t.bedgraph"
    plain.write_text(text)
    assert plain.stat().st_size > 2 * 64 * 0xff00
    outs = []
    for threads in (1, 8):
        gz = str(tmp_path / ("t%d.bedgraph.gz" % threads))
        hostio.bgzip_tabix(str(plain), gz, threads=threads)
        outs.append((open(gz, "rb").read(), open(gz + ".tbi", "rb").read()))
    assert outs[0] == outs[1]
    assert gzip.open(str(tmp_path / "t1.bedgraph.gz"), "rt").read() == text
    got = list(hostio.TabixFile(str(tmp_path / "t1.bedgraph.gz")).fetch("chr2", 1000, 1200))
    exp = [r for r in rows[130000:] if 1000 <= int(r.split("\t")[1]) < 1200]
    assert [("\t".join(g) if isinstance(g, (list, tuple)) else g) for g in got] == exp and len(exp) > 3


def test_bgzip_tabix_level(tmp_path, monkeypatch):
    """nb200_bgzip_tabix_level: the default level (-1) writes the file nb200_bgzip_tabix writes (pysam.tabix_compress's level,
    run_occ.py:130-136); level 1 (also through NB200_GZ_LEVEL) writes other bytes that inflate to the same rows and answer
    the same region queries; a level outside -1, 1..9 is an error."""
    rng = np.random.RandomState(5)
    rows = ["chr1\t%d\t%d\t%s" % (100 + k, 101 + k, repr(float(rng.rand()))) for k in range(60000)]
    text = "\n".join(rows) + "\n"
    plain = tmp_path / "l.bedgraph"
    plain.write_text(text)
    gz = lambda name: str(tmp_path / name)
    hostio.bgzip_tabix(str(plain), gz("d.gz"), threads=2)
    hostio.bgzip_tabix(str(plain), gz("m1.gz"), threads=2, level=-1)
    hostio.bgzip_tabix(str(plain), gz("l1.gz"), threads=2, level=1)
    monkeypatch.setenv("NB200_GZ_LEVEL", "1")
    hostio.bgzip_tabix(str(plain), gz("e1.gz"), threads=3)
    monkeypatch.delenv("NB200_GZ_LEVEL")
    rd = lambda name: open(gz(name), "rb").read()
    assert rd("d.gz") == rd("m1.gz") and rd("d.gz.tbi") == rd("m1.gz.tbi")
    assert rd("l1.gz") == rd("e1.gz") and rd("l1.gz") != rd("d.gz") and len(rd("l1.gz")) > len(rd("d.gz"))
    assert gzip.open(gz("l1.gz"), "rt").read() == text
    q = lambda name: list(hostio.TabixFile(gz(name)).fetch("chr1", 30000, 30050))
    assert q("l1.gz") == q("d.gz") and len(q("d.gz")) == 50
    for bad in (0, 10, -2):
        with pytest.raises(IOError):
            hostio.bgzip_tabix(str(plain), gz("bad.gz"), level=bad)


def test_fuzz_fits_pool_equals_serial(monkeypatch):
    """Nucleosome.getFuzz (NucleosomeCalling.py:137-194) as fuzz.fit_fuzz: the fits of a batch on the worker pool are the
    fits made one by one, value for value; a pool that cannot start falls back to the calling process."""
    from concurrent.futures.process import BrokenProcessPool
    from nucleoatac_b200 import fuzz
    rng = np.random.RandomState(11)
    jobs = []
    for k in range(70):
        n = int(rng.randint(70, 200))
        x = np.arange(n)
        m = int(rng.randint(30, n - 30))
        sig = rng.uniform(0.3, 1.2) * np.exp(-(x - m) ** 2 / (2 * rng.uniform(8, 20) ** 2)) + 0.01 * rng.rand(n)
        jobs.append((sig, (m,) if k % 3 else (m, min(n - 5, m + 40)), 10))
    monkeypatch.setenv("NB200_FUZZ_PROCS", "0")
    serial = fuzz.fit_many(jobs)
    assert all(2.0 <= f <= 50.0 and w > 0 for f, w, _ in serial)
    monkeypatch.setenv("NB200_FUZZ_PROCS", "3")
    try:
        assert fuzz.fit_many(jobs) == serial
        assert fuzz._pool is not None
        assert fuzz.fit_many(jobs[:10]) == serial[:10]          # too few jobs for the pool: fitted here
    finally:
        fuzz.close_pool()

    def broken(n):
        raise BrokenProcessPool("workers did not start")
    monkeypatch.setattr(fuzz, "_get_pool", broken)
    monkeypatch.setattr(fuzz, "_broken", False)
    assert fuzz.fit_many(jobs) == serial and fuzz._broken


def test_native_bedgraph_fetch_equals_python_reader(tmp_path):
    """nb200_bedgraph_fetch (BedGraphFile.read, pyatac/bedgraph.py:6-16 = pysam.Tabixfile.fetch into a dense array) against the
    Python tabix reader: run-length rows, gaps (NaN stretches are not written), several chromosomes, regions that start
    before the first row, straddle a chromosome's last rows, hit nothing, or name an unknown chromosome."""
    from nucleoatac_b200.bedgraph import BedGraphFile
    from nucleoatac_b200.tracks import Track
    rng = np.random.RandomState(9)
    plain = tmp_path / "b.bedgraph"
    with open(str(plain), "wb") as fh:
        for chrom, n in (("chr1", 120000), ("chr2", 40000), ("chrX", 9000)):
            v = np.round(rng.rand(n), 2)                     # runs of equal values
            v[rng.rand(n) < 0.02] = np.nan
            v[5000:7000] = np.nan                            # a long gap
            v[20000:20600] = 0.0                             # zeros are rows of their own
            fh.write(Track(chrom, 1000, 1000 + n, vals=v).format_track())
    gz = str(plain) + ".gz"
    hostio.bgzip_tabix(str(plain), gz, threads=2)
    bg = BedGraphFile(gz)
    assert isinstance(bg.reader, hostio.TabixFile)
    queries = [("chr1", 0, 3000), ("chr1", 900, 1100), ("chr1", 5500, 6500), ("chr1", 16000, 33500), ("chr1", 65535, 65537), ("chr1", 120500, 122000),
               ("chr1", 130000, 130100), ("chr2", 1000, 41000), ("chr2", 39990, 41050), ("chrX", 2, 3), ("chrX", 9990, 10010), ("chrNope", 5, 50),
               ("chr2", 777, 777)]
    queries += [("chr1", int(a), int(a + w)) for a, w in zip(rng.randint(0, 121000, 12), rng.randint(1, 12000, 12))]
    n_vals = 0
    for chrom, a, b in queries:
        for empty in (np.nan, 0.0):
            got, want = bg.read(chrom, a, b, empty=empty), bg.read_python(chrom, a, b, empty=empty)
            assert got.shape == want.shape and np.array_equal(got, want, equal_nan=True), (chrom, a, b, empty)
        n_vals += int((~np.isnan(got)).sum())
    assert n_vals > 50000
    with pytest.raises(IOError):
        bad = BedGraphFile(gz)                                                # a missing index is an error, not an empty track
        bad.reader = hostio.TabixFile(gz)
        bad.reader.path = str(tmp_path / "missing.gz")
        bad.read("chr1", 0, 10)
