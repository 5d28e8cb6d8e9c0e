#!/usr/bin/env python
"""Throughput of the host legs either side of the device path (reported separately from the scoring metric):
BAM region decode, FASTA fetch, bedgraph formatting, BGZF + tabix.  No GPU needed.

    python tools/host_throughput.py [bam] [fasta]      # defaults: a synthetic indexed BAM / genome built in a temp dir
"""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from nucleoatac_b200 import hostio
from nucleoatac_b200.tracks import Track


def synthetic(tmp, length=4_000_000, density=0.25):
    from tests.synthfiles import write_bam
    rng = np.random.default_rng(1)
    n = int(length * density)
    pos = np.sort(rng.integers(0, length - 700, n))
    tl = rng.integers(40, 600, n)
    bam = os.path.join(tmp, "synth.bam")
    write_bam(bam, {"chrS": length}, [(0, int(p), int(t)) for p, t in zip(pos, tl)])
    hostio.index_bam(bam)
    fa = os.path.join(tmp, "synth.fa")
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, length)].tobytes().decode()
    with open(fa, "w") as fh:
        fh.write(">chrS\n")
        for i in range(0, length, 60):
            fh.write(seq[i:i + 60] + "\n")
    hostio.FastaFile(fa).close()
    return bam, fa


def main():
    tmp = tempfile.mkdtemp(prefix="nb200_host_")
    bam_path, fa_path = (sys.argv[1], sys.argv[2]) if len(sys.argv) > 2 else synthetic(tmp)
    bam = hostio.BamFile(bam_path)
    chrom, clen = bam.references[0], bam.lengths[0]
    regions = [(chrom, s - 500, s + 10500) for s in range(1000, clen - 11000, 10000)]
    out = dict(cores=len(os.sched_getaffinity(0)), regions=len(regions), region_bp=11000)
    bam.fetch_fragments_many(regions[:2])                       # load the library
    for th in (1, out["cores"]):
        dt = 1e30
        for _ in range(3):   # best of three: the first calls on shared cores are several times slower than the rest
            t = time.perf_counter()
            off, pos, tlen = bam.fetch_fragments_many(regions, threads=th)
            dt = min(dt, time.perf_counter() - t)
        out["bam_decode_Mfrag_s_%dthr" % th] = round(len(pos) / dt / 1e6, 2)
        out["bam_decode_Mbp_s_%dthr" % th] = round(len(regions) * 10000 / dt / 1e6, 1)
    t = time.perf_counter()
    n = sum(len(bam._fetch_indexed(0, max(0, s), e)[0]) for _c, s, e in regions[:20])
    out["bam_decode_python_Mfrag_s"] = round(n / (time.perf_counter() - t) / 1e6, 3)
    fa = hostio.FastaFile(fa_path)
    t = time.perf_counter()
    nb = sum(len(fa.fetch(chrom, max(0, s), e)) for _c, s, e in regions)
    out["fasta_fetch_MB_s"] = round(nb / (time.perf_counter() - t) / 1e6, 1)
    # bedgraph rows of a continuous track (every position its own row, the worst case) + bgzip / tabix
    vals = np.random.default_rng(0).random(10000)
    plain = os.path.join(tmp, "t.bedgraph")
    t = time.perf_counter()
    with open(plain, "w") as fh:
        for k in range(200):
            Track(chrom, 10000 * k, 10000 * (k + 1), vals=vals).write_track(fh)
    dt = time.perf_counter() - t
    out["bedgraph_format_Mrows_s"] = round(200 * 10000 / dt / 1e6, 2)
    size = os.path.getsize(plain)
    t = time.perf_counter()
    hostio.bgzip_tabix(plain, plain + ".gz")
    dt = time.perf_counter() - t
    out["bgzip_tabix_MB_s"] = round(size / dt / 1e6, 1)
    out["bedgraph_bytes_per_bp"] = round(size / 2e6, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
