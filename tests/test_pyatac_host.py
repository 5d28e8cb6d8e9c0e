"""Host-side logic of the pyatac tools (no GPU): command-line defaults as in pyatac/cli.py, the region lists of the track
tools, round-robin sharded track writing merged back into chunk order + bgzip/tabix, and the NUMA helper's contract."""
import gzip
import os

import numpy as np


def test_parser_defaults_match_reference():
    from nucleoatac_b200.pyatac_tools import build_parser
    p = build_parser()
    a = p.parse_args(["vplot", "--bed", "x.bed", "--bam", "x.bam"])      # pyatac/cli.py:202-231
    assert (a.lower, a.upper, a.flank, a.scale, a.atac, a.strand) == (0, 250, 250, False, True, None)
    a = p.parse_args(["cov", "--bam", "x.bam"])                           # pyatac/cli.py:337-353
    assert (a.lower, a.upper, a.window, a.scale, a.bed) == (0, 2000, 121, 10, None)
    a = p.parse_args(["ins", "--bam", "x.bam", "--not_atac"])            # pyatac/cli.py:314-331
    assert (a.lower, a.upper, a.smooth, a.atac) == (0, 2000, None, False)
    a = p.parse_args(["bias", "--fasta", "g.fa"])                         # pyatac/cli.py:137-150
    assert a.pwm == "Human" and a.bed is None
    a = p.parse_args(["sizes", "--bam", "x.bam"])                         # pyatac/cli.py:115-131
    assert (a.lower, a.upper) == (0, 500)


def test_regions_whole_genome_and_bed(tmp_path):
    from nucleoatac_b200.pyatac_tools import _regions

    class A:
        bed = None
    chunks = _regions(A, {"chrB": 2500, "chrA": 1000}, 50000, 1000)      # convertChromSizes(splitsize=1000), get_ins.py:73-75
    assert [(c.chrom, c.start, c.end) for c in chunks] == [("chrA", 0, 1000), ("chrB", 0, 1000), ("chrB", 1000, 2000), ("chrB", 2000, 2500)]
    bed = tmp_path / "r.bed"
    bed.write_text("chrA\t10\t50\nchrA\t40\t90\nchrZ\t1\t5\nchrB\t5\t9\n")
    A.bed = str(bed)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        chunks = _regions(A, {"chrB": 2500, "chrA": 1000}, 50000, 1000)  # read, drop unknown chromosomes, merge overlaps
    assert [(c.chrom, c.start, c.end) for c in chunks] == [("chrA", 10, 90), ("chrB", 5, 9)]


def test_sharded_track_writing_round_trip(tmp_path, monkeypatch):
    """Two ranks write their chunks (k mod 2); rank 0 interleaves the blocks back into chunk order, compresses and indexes."""
    from nucleoatac_b200 import dist as _d, hostio
    monkeypatch.setattr(_d, "barrier", lambda world=None: None)   # both "ranks" run in this process, one after the other
    from nucleoatac_b200.chunk import Chunk, ChunkList
    from nucleoatac_b200.pyatac_tools import _write_tracks
    from nucleoatac_b200.tracks import Track
    chunks = ChunkList(*[Chunk("chr1", 100 * k, 100 * k + 37) for k in range(7)])

    def make(c):
        return Track(c.chrom, c.start, c.end, vals=np.arange(c.start, c.end, dtype=np.float64) / 8.0)

    class Args:
        pass
    single = Args()
    single.out, single.rank, single.world = str(tmp_path / "one"), 0, 1
    _write_tracks(single, ".cov", chunks, make)
    for rank in (1, 0):   # rank 0 last: it merges what both wrote (both shards run in this process; the barrier is patched out)
        a = Args()
        a.out, a.rank, a.world = str(tmp_path / "two"), rank, 2
        if rank == 1:     # rank 1 only writes its shard; emulate by stopping before the merge
            from nucleoatac_b200 import dist
            w = dist.ShardWriter(a.out + ".cov.bedgraph", 1, 2)
            for c in dist.shard(chunks, 1, 2):
                make(c).write_track(w)
                w.end_chunk()
            w.close()
        else:
            _write_tracks(a, ".cov", chunks, make)
    one = gzip.open(str(tmp_path / "one.cov.bedgraph.gz"), "rt").read()
    two = gzip.open(str(tmp_path / "two.cov.bedgraph.gz"), "rt").read()
    assert one == two and one.count("\n") == 7 * 37
    assert os.path.exists(str(tmp_path / "two.cov.bedgraph.gz.tbi"))
    rows = hostio.TabixFile(str(tmp_path / "two.cov.bedgraph.gz")).fetch("chr1", 300, 310)
    assert [(r if isinstance(r, (list, tuple)) else r.split("\t"))[1] for r in rows] == [str(x) for x in range(300, 310)]


def test_numa_helper_never_raises():
    from nucleoatac_b200 import dist
    info = dist.bind_to_gpu_numa(0)   # no NVML / no GPU here: reports why and changes nothing
    assert set(("gpu_numa_node", "cpus_bound", "mempolicy")) <= set(info)
    dist.reset_mempolicy()
