#!/usr/bin/env python
"""Host side of `nucleoatac occ` / `nuc` without a GPU: the drivers run on synthetic files with a scorer that packs the batch
exactly as the device path does (BAM decode, FASTA fetch, PackedBatch) and fills the chunks with placeholder tracks, so that
what is timed is everything around the kernels: reading, packing, formatting, writing, bgzip + tabix.

    python tools/cli_host_profile.py [n_chunks] [--profile]"""
import cProfile
import os
import pstats
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 200
    from tests.synthfiles import make_files
    from nucleoatac_b200.cli import build_parser
    from nucleoatac_b200.NucleosomeCalling import Nucleosome
    from nucleoatac_b200.Occupancy import OccPeak, OccupancyTrack
    from nucleoatac_b200.run_nuc import run_nuc
    from nucleoatac_b200.run_occ import run_occ
    from nucleoatac_b200.tracks import CoverageTrack, Track
    d = tempfile.mkdtemp(prefix="nb200_host_")
    t = time.time()
    f = make_files(d, ks=tuple(range(n)))
    print("files for %d chunks written in %.1f s" % (n, time.time() - t))
    rng = np.random.default_rng(0)
    pool = rng.random(1 << 20)
    t_pack = [0.0]

    def vals(c, k):
        o = (c.start * 7 + k * 1013) % (len(pool) - (c.end - c.start))
        return pool[o:o + c.end - c.start]

    def score_occ(occs, params):
        t0 = time.perf_counter()
        params.pack(occs)
        t_pack[0] += time.perf_counter() - t0
        for oc in occs:
            oc.params = params
            oc.occ = OccupancyTrack(oc.chrom, oc.start, oc.end)
            oc.occ.smoothed_vals, oc.occ.smoothed_lower, oc.occ.smoothed_upper = vals(oc, 0), vals(oc, 1), vals(oc, 2)
            oc.cov = CoverageTrack(oc.chrom, oc.start, oc.end)
            oc.cov.vals = vals(oc, 3)
            oc.peaks = {p - oc.start: OccPeak(p, oc) for p in range(oc.start + 80, oc.end - 80, 170)}
            oc._nuc_dist = np.zeros(params.upper)

    def score_nuc(nucs, params):
        t0 = time.perf_counter()
        params.pack(nucs)
        t_pack[0] += time.perf_counter() - t0
        for nc in nucs:
            nc.norm_signal, nc.smoothed, nc.nuc_signal = (Track(nc.chrom, nc.start, nc.end, vals=vals(nc, k)) for k in (0, 1, 2))
            nc.nuc_cov = nc.nfr_cov = nc.bias = nc.nuc_signal
            calls = list(range(nc.start + 80, nc.end - 80, 170))
            nc.nuc_collection = {p: Nucleosome(p, nc) for p in calls}
            nc.nonredundant, nc.redundant = np.array(calls), np.array(calls[:3])

    out = os.path.join(d, "run")
    base = ["--bed", f["bed"], "--bam", f["bam"], "--fasta", f["fasta"], "--sizes", f["sizes"], "--out", out]
    for name, argv, fn in (("occ", ["occ"] + base, lambda a: run_occ(a, score=score_occ)),
                           ("nuc", ["nuc"] + base + ["--vmat", f["vmat"]], lambda a: run_nuc(a, score=score_nuc))):
        args = build_parser().parse_args(argv)
        t_pack[0] = 0.0
        pr = cProfile.Profile() if "--profile" in sys.argv else None
        t = time.time()
        if pr:
            pr.enable()
        fn(args)
        if pr:
            pr.disable()
        dt = time.time() - t
        print("== %s host side: %.2f s for %d chunks -> %.2f Mbp/s (reading + packing %.2f s)" % (name, dt, n, n * 0.01 / dt, t_pack[0]))
        if pr:
            pstats.Stats(pr).sort_stats("cumulative").print_stats(18)


if __name__ == "__main__":
    main()
