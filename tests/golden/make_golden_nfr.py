"""Extract the NFR fixtures (tests/golden/nfr_golden.npz) from the reference tree -- run in the BUILD container only.

Inputs of `nucleoatac nfr` as `nucleoatac run` wires them (nucleoatac/cli.py:47-49): --bed example.bed, --occ_track
example_results/example.occ.bedgraph.gz (+ its upper_bound sibling, NFRCalling.py:66-68), --calls
example_results/example.nucmap_combined.bed.gz, --bam example.bam, --fasta sacCer3.fa, Human PWM.  Golden outputs: the
rows of example_results/example.nfrpos.bed.gz and the values of example_results/example.ins.bedgraph.gz.  Nothing is
computed by the oracle here: inputs are decoded, golden outputs are parsed, both are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import hostio, refalgo as ra  # noqa: E402

EX = "/root/reference/example/"
RES = EX + "example_results/"
PWM_PATH = "/root/reference/pyatac/pwm/Human.PWM.txt"
READ_MARGIN = 2100  # getInsertions fetches [start - 2000, end + 2000)


def main():
    _, frags = hostio.read_bam_fragments(EX + "example.bam")
    fa = hostio.Fasta(EX + "sacCer3.fa")
    chrs = fa.chrom_sizes()
    pwm, up, down, nucs = ra.read_pwm(PWM_PATH)
    # run_nfr.py:84-91: ChunkList.read(bed, chromDict, min_offset = max(pwm.up, pwm.down)); chunks.merge()
    chunks = ra.merge_chunks(ra.read_bed_chunks(EX + "example.bed", chrs, min_offset=max(up, down)))
    chrom_names = sorted({c for c, _, _ in chunks})
    cidx = {c: i for i, c in enumerate(chrom_names)}
    occ_rows = hostio.read_bedgraph_gz(RES + "example.occ.bedgraph.gz")
    upp_rows = hostio.read_bedgraph_gz(RES + "example.occ.upper_bound.bedgraph.gz")
    ins_rows = hostio.read_bedgraph_gz(RES + "example.ins.bedgraph.gz")
    calls = hostio.read_bedgraph_gz(RES + "example.nucmap_combined.bed.gz")
    pos_l, tlen_l, foff = [], [], [0]
    seq_l, soff = [], [0]
    occ_l, upp_l, ins_l, toff = [], [], [], [0]
    nuc_l, noff = [], [0]
    for c, s, e in chunks:
        p, t = frags[c]
        sel = (p >= s - READ_MARGIN) & (p < e + READ_MARGIN)
        pos_l.append(p[sel])
        tlen_l.append(t[sel])
        foff.append(foff[-1] + int(sel.sum()))
        seq = fa.fetch(c, s - up, e + down)
        assert len(seq) == e - s + up + down
        seq_l.append(np.frombuffer(seq.encode(), dtype=np.uint8))
        soff.append(soff[-1] + len(seq))
        occ_l.append(hostio.bedgraph_region(occ_rows, c, s, e))
        upp_l.append(hostio.bedgraph_region(upp_rows, c, s, e))
        ins_l.append(hostio.bedgraph_region(ins_rows, c, s, e))
        toff.append(toff[-1] + e - s)
        # pysam.TabixFile.fetch(chrom, start, end): rows [pos, pos+1) overlapping [start, end), file order
        dy = [int(r[1]) for r in calls if r[0] == c and int(r[1]) < e and int(r[2]) > s]
        nuc_l.append(np.array(dy, dtype=np.int32))
        noff.append(noff[-1] + len(dy))
    nfr = hostio.read_bedgraph_gz(RES + "example.nfrpos.bed.gz")
    np.savez_compressed(
        os.path.join(HERE, "nfr_golden.npz"),
        chrom_names=np.array(chrom_names),
        chunk_chrom=np.array([cidx[c] for c, _, _ in chunks], dtype=np.int32),
        chunk_start=np.array([s for _, s, _ in chunks], dtype=np.int32),
        chunk_end=np.array([e for _, _, e in chunks], dtype=np.int32),
        frag_off=np.array(foff, dtype=np.int64), frag_pos=np.concatenate(pos_l).astype(np.int32),
        frag_tlen=np.concatenate(tlen_l).astype(np.int32),
        seq_off=np.array(soff, dtype=np.int64), seq=np.concatenate(seq_l),
        track_off=np.array(toff, dtype=np.int64), occ=np.concatenate(occ_l), occ_upper=np.concatenate(upp_l),
        nuc_off=np.array(noff, dtype=np.int64), nuc_pos=np.concatenate(nuc_l),
        pwm=pwm, pwm_up=up, pwm_down=down, pwm_nucleotides=np.array(nucs),
        gold_ins=np.concatenate(ins_l),
        gold_nfr_chrom=np.array([cidx[r[0]] for r in nfr], dtype=np.int32),
        gold_nfr_left=np.array([int(r[1]) for r in nfr], dtype=np.int32),
        gold_nfr_right=np.array([int(r[2]) for r in nfr], dtype=np.int32),
        gold_nfr_vals=np.array([[float(x) for x in r[3:7]] for r in nfr], dtype=np.float64),
        gold_nfr_text="".join("\t".join(r) + "\n" for r in nfr),
    )
    print("nfr_golden.npz", os.path.getsize(os.path.join(HERE, "nfr_golden.npz")), "chunks", len(chunks), "nfrs", len(nfr))


if __name__ == "__main__":
    main()
