"""Test helper: materialise a synthetic workload as real files (BAM, FASTA + .fai, BED, VMat, sizes) so that the
file-based API mirror and the CLI can be exercised on the GPU box, where /root/reference does not exist."""
import os
import struct

import numpy as np

from nucleoatac_b200 import hostio, synth


def write_bam(path, chrom_sizes, reads):
    """Minimal coordinate-sorted paired BAM: `reads` = list of (tid, pos, tlen) for the forward mates; the reverse
    mate of every pair is written too (flag 147) so that the proper-pair/forward filter has something to reject."""
    names = list(chrom_sizes.keys())
    w = hostio.BgzfWriter(path)
    text = "@HD\tVN:1.0\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % (n, chrom_sizes[n]) for n in names)
    hdr = b"BAM\x01" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(names))
    for n in names:
        hdr += struct.pack("<i", len(n) + 1) + n.encode() + b"\x00" + struct.pack("<i", chrom_sizes[n])
    w.write(hdr)
    recs = []
    for i, (tid, pos, tlen) in enumerate(reads):
        recs.append((tid, pos, 99, tlen, pos + abs(tlen) - 36, i))
        recs.append((tid, max(pos + abs(tlen) - 36, 0), 147, -tlen, pos, i))
    recs.sort(key=lambda r: (r[0], r[1]))
    lseq = 36
    for tid, pos, flag, tlen, mpos, i in recs:
        name = ("r%d" % i).encode() + b"\x00"
        body = struct.pack("<iiBBHHHiiii", tid, pos, len(name), 30, 4680, 1, flag, lseq, tid, mpos, tlen)
        body += name + struct.pack("<I", (lseq << 4) | 0) + b"\x11" * ((lseq + 1) // 2) + b"\x28" * lseq
        w.write(struct.pack("<i", len(body)) + body)
    w.close()


def make_files(tmpdir, ks=(2, 0, 5), R=251, W=251):
    """-> dict of paths + the Workload and the in-memory chunks (oracle inputs)."""
    wl = synth.Workload(R, W)
    chunks = [synth.make_chunk(k) for k in ks]
    length = max(c[1] for c in chunks) + 2000
    rng = np.random.default_rng(99)
    genome = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, length)].copy()
    reads = []
    for (s, e, pos, tlen, seq, s0) in chunks:
        genome[s0:s0 + len(seq)] = seq
        reads += [(0, int(p), int(t)) for p, t in zip(pos, tlen)]
    fa = os.path.join(tmpdir, "genome.fa")
    with open(fa, "w") as fh:
        fh.write(">chrS\n")
        txt = genome.tobytes().decode()
        for i in range(0, len(txt), 60):
            fh.write(txt[i:i + 60] + "\n")
    hostio.FastaFile(fa).close()  # builds the .fai
    bam = os.path.join(tmpdir, "reads.bam")
    write_bam(bam, {"chrS": length}, reads)
    hostio.index_bam(bam)  # .bai -> the drivers decode regions with the native reader (nb200_bam_fetch_many)
    bed = os.path.join(tmpdir, "regions.bed")
    with open(bed, "w") as fh:  # the drivers slop by nuc_sep/2 = 60 on both sides
        for (s, e, *_r) in sorted(chunks, key=lambda c: c[0]):  # the reference requires a sorted BED (docs/nucleoatac.md:10)
            fh.write("chrS\t%d\t%d\n" % (s + 60, e - 60))
    from nucleoatac_b200.fragmentsizes import FragmentSizes
    from nucleoatac_b200.VMat import VMat
    vm = os.path.join(tmpdir, "synthetic.VMat")
    VMat(wl.vmat, wl.v_lower, wl.v_upper).save(vm)
    sizes = os.path.join(tmpdir, "sizes.txt")
    FragmentSizes(0, wl.upper, vals=wl.fragmentsizes).save(sizes)
    return dict(fasta=fa, bam=bam, bed=bed, vmat=vm, sizes=sizes, wl=wl, chunks=chunks, genome=genome, length=length)
