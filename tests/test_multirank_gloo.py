"""world_size-2 test of the N>1 path on CPU (gloo): round-robin sharding of the BED chunks, per-rank outputs,
the end-of-run reductions (fragment-size histogram, nuc_dist) and the merge back into chunk order.  The device
scorer is replaced by the CPU oracle (tests may use it), so this exercises exactly the host-side multi-GPU logic
of run_occ / run_nuc; the result must equal the single-rank run byte for byte."""
import filecmp
import multiprocessing as mp
import os
import socket

import numpy as np
import pytest

from oracle import refalgo as ra, refocc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_occ_scorer(files):
    from nucleoatac_b200.Occupancy import OccPeak, OccupancyTrack
    from nucleoatac_b200.tracks import CoverageTrack
    g = files["genome"]

    def score(occs, params):
        cp = params.occ_calc_params
        op = refocc.OccParams(cp.nuc_probs, cp.nfr_probs, upper=params.upper)
        for oc in occs:
            pos, tlen = files["reads"]
            span = refocc.occ_bias_track_span(oc.start, oc.end, op)
            bt = ra.log_bias_track(g[span[0] - 10:span[1] + 10].tobytes().decode(), files["wl"].pwm, files["wl"].nucleotides)
            r = refocc.process_occ_chunk(pos, tlen, oc.start, oc.end, op, bias_track=bt, bias_track_start=span[0])
            oc.params = params
            oc.occ = OccupancyTrack(oc.chrom, oc.start, oc.end)
            oc.occ.smoothed_vals, oc.occ.smoothed_lower, oc.occ.smoothed_upper = r["smoothed_vals"], r["smoothed_lower"], r["smoothed_upper"]
            oc.cov = CoverageTrack(oc.chrom, oc.start, oc.end)
            oc.cov.vals = r["cov"]
            oc.peaks = {p[0] - oc.start: OccPeak(p[0], oc) for p in r["peaks"]}
            oc._nuc_dist = r["nuc_dist"]
    return score


def _count_sizes(files):
    def count(chunks, bam, lower, upper):
        pos, tlen = files["reads"]
        return ra.fragment_size_counts(pos, tlen, [(c.start, c.end) for c in chunks], lower, upper)
    return count


def _worker(rank, world, port, files, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    if world > 1:
        import torch.distributed as td
        td.init_process_group("gloo", rank=rank, world_size=world)
    from nucleoatac_b200.cli import build_parser
    from nucleoatac_b200.run_occ import run_occ
    args = build_parser().parse_args(["occ", "--bed", files["bed"], "--bam", files["bam"], "--fasta", files["fasta"], "--out", out,
                                      "--rank", str(rank), "--world", str(world), "--batch", "2"])
    # the gamma NFR model is replaced by a fixed split so that the test stays fast and deterministic
    import nucleoatac_b200.Occupancy as O
    from nucleoatac_b200.fragmentsizes import FragmentSizes

    def fake_model(self, boundaries=(35, 115)):
        self.nuc_fit = FragmentSizes(0, 251, vals=files["wl"].nuc_probs)
        self.nfr_fit = FragmentSizes(0, 251, vals=files["wl"].nfr_probs)
    O.FragmentMixDistribution.modelNFR = fake_model
    O.FragmentMixDistribution.plotFits = lambda self, filename=None: None
    run_occ(args, score=_oracle_occ_scorer(files), count_sizes=_count_sizes(files))
    if world > 1:
        import torch.distributed as td
        td.destroy_process_group()


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    from tests.synthfiles import make_files
    d = str(tmp_path_factory.mktemp("mr"))
    f = make_files(d, ks=(0, 1, 2))
    # five short regions inside the three synthetic chunks -> 5 BED chunks over 2 ranks (3 + 2)
    with open(f["bed"], "w") as fh:
        for s in (10500, 13000, 23000, 26000, 35000):
            fh.write("chrS\t%d\t%d\n" % (s, s + 700))
    pos = np.concatenate([c[2] for c in f["chunks"]])
    tlen = np.concatenate([c[3] for c in f["chunks"]])
    f["reads"] = (pos, tlen)
    return f


def test_two_ranks_equal_one_rank(files, tmp_path):
    ctx = mp.get_context("spawn")
    out1, out2 = str(tmp_path / "w1"), str(tmp_path / "w2")
    p = ctx.Process(target=_worker, args=(0, 1, _free_port(), files, out1))
    p.start()
    p.join(300)
    assert p.exitcode == 0
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, files, out2)) for r in range(2)]
    for q in procs:
        q.start()
    for q in procs:
        q.join(300)
        assert q.exitcode == 0
    for suffix in (".occ.bedgraph.gz", ".occ.lower_bound.bedgraph.gz", ".occ.upper_bound.bedgraph.gz", ".occpeaks.bed.gz",
                   ".fragmentsizes.txt"):
        assert filecmp.cmp(out1 + suffix, out2 + suffix, shallow=False), suffix
    from nucleoatac_b200.fragmentsizes import FragmentSizes
    a, b = FragmentSizes.open(out1 + ".nuc_dist.txt").get(), FragmentSizes.open(out2 + ".nuc_dist.txt").get()
    np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-15)  # float sum order differs between 1 and 2 ranks
    assert a.sum() > 0
    assert not [f for f in os.listdir(str(tmp_path)) if ".rank" in f]


def test_shard_writer_merge(tmp_path):
    from nucleoatac_b200 import dist
    n, world = 7, 3
    for r in range(world):
        w = dist.ShardWriter(str(tmp_path / "x.txt"), r, world)
        for k in dist.shard(list(range(n)), r, world):
            w.write("chunk%d\n" % k * (k % 3))  # some chunks write nothing
            w.end_chunk()
        w.close()
    dist.ShardWriter.merge(str(tmp_path / "x.txt"), world, n)
    assert open(str(tmp_path / "x.txt")).read() == "".join("chunk%d\n" % k * (k % 3) for k in range(n))
    assert dist.shard(list(range(7)), 1, 3) == [1, 4]


def test_run_nuc_driver_writes_batches_in_chunk_order(files, tmp_path):
    """The host side of `nucleoatac nuc` (run_nuc.py:141-201) around an injected scorer: batches of two chunks, tracks
    formatted on the writer pool behind the scoring of the next batch, rows and calls in chunk order, all six outputs
    (--write_all) bgzipped + indexed; a scorer that fails surfaces as the run's exception."""
    import gzip
    from nucleoatac_b200.cli import build_parser
    from nucleoatac_b200.NucleosomeCalling import Nucleosome
    from nucleoatac_b200.run_nuc import run_nuc
    from nucleoatac_b200.tracks import Track
    out = str(tmp_path / "n")
    args = build_parser().parse_args(["nuc", "--bed", files["bed"], "--bam", files["bam"], "--fasta", files["fasta"], "--vmat", files["vmat"],
                                      "--sizes", files["sizes"], "--out", out, "--batch", "2", "--write_all"])
    seen = []

    def track(nc, seed):
        rng = np.random.RandomState(seed + nc.start)
        v = np.round(rng.rand(nc.end - nc.start), 1)      # runs of equal values
        v[rng.rand(len(v)) < 0.05] = np.nan
        return Track(nc.chrom, nc.start, nc.end, vals=v)

    def score(nucs, params):
        seen.append([nc.start for nc in nucs])
        for nc in nucs:
            nc.norm_signal, nc.smoothed, nc.bias, nc.nuc_signal = (track(nc, s) for s in (1, 2, 3, 4))
            nc.nuc_cov = nc.nfr_cov = track(nc, 5)
            calls = [nc.start + 100, nc.start + 300, nc.start + 450]
            nc.nuc_collection = {p: Nucleosome(p, nc) for p in calls}
            nc.nonredundant, nc.redundant = np.array(calls[::2]), np.array(calls[1:2])
    run_nuc(args, score=score)
    assert [len(g) for g in seen] == [2, 2, 1]
    starts = [s for g in seen for s in g]
    assert starts == sorted(starts)
    rd = lambda suffix: gzip.open(out + suffix, "rt").read()
    for suffix, seed in ((".nucleoatac_signal.bedgraph.gz", 1), (".nucleoatac_signal.smooth.bedgraph.gz", 2),
                         (".nucleoatac_background.bedgraph.gz", 3), (".nucleoatac_raw.bedgraph.gz", 4)):
        exp = ""
        for s in starts:
            rng = np.random.RandomState(seed + s)
            end = dict(seen_spans(files))[s]
            v = np.round(rng.rand(end - s), 1)
            v[rng.rand(len(v)) < 0.05] = np.nan
            exp += ra.write_track("chrS", s, end, v)
        assert rd(suffix) == exp, suffix
        assert os.path.exists(out + suffix + ".tbi")
    pos = [int(r.split("\t")[1]) for r in rd(".nucpos.bed.gz").splitlines()]
    assert pos == [p for s in starts for p in (s + 100, s + 450)]
    assert [int(r.split("\t")[1]) for r in rd(".nucpos.redundant.bed.gz").splitlines()] == [s + 300 for s in starts]
    assert not [f for f in os.listdir(str(tmp_path)) if f.endswith(".bedgraph") or f.endswith(".bed")]

    def failing(nucs, params):
        if len(seen) >= 4:
            raise RuntimeError("scorer failed")
        score(nucs, params)
    with pytest.raises(RuntimeError, match="scorer failed"):
        run_nuc(args, score=failing)


def seen_spans(files):
    """(start, end) of the chunks the nuc driver makes of the fixture's BED (slop by nuc_sep // 2, merged)."""
    from nucleoatac_b200.cli import build_parser
    from nucleoatac_b200.run_nuc import nuc_chunks
    from nucleoatac_b200.VMat import VMat
    args = build_parser().parse_args(["nuc", "--bed", files["bed"], "--bam", files["bam"], "--fasta", files["fasta"], "--vmat", files["vmat"],
                                      "--sizes", files["sizes"], "--out", "x"])
    return [(c.start, c.end) for c in nuc_chunks(args, VMat.open(files["vmat"]))]
