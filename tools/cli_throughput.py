#!/usr/bin/env python
"""Time the file-based drivers (`nucleoatac occ` / `nuc`) on a synthetic BAM + FASTA: where does host time go?"""
import cProfile
import os
import pstats
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    from tests.synthfiles import make_files
    from nucleoatac_b200.cli import nucleoatac_main
    d = tempfile.mkdtemp(prefix="nb200_cli_")
    t = time.time()
    f = make_files(d, ks=tuple(range(n)))
    print("files for %d chunks written in %.1f s" % (n, time.time() - t))
    out = os.path.join(d, "run")
    for cmd in (["nuc", "--bed", f["bed"], "--bam", f["bam"], "--fasta", f["fasta"], "--vmat", f["vmat"], "--sizes", f["sizes"], "--out", out],
                ["occ", "--bed", f["bed"], "--bam", f["bam"], "--fasta", f["fasta"], "--sizes", f["sizes"], "--out", out]):
        pr = cProfile.Profile()
        t = time.time()
        pr.enable()
        nucleoatac_main(cmd)
        pr.disable()
        dt = time.time() - t
        print("== %s: %.2f s for %d chunks -> %.2f Mbp/s" % (cmd[0], dt, n, n * 0.01 / dt))
        pstats.Stats(pr).sort_stats("cumulative").print_stats(14)


if __name__ == "__main__":
    main()
