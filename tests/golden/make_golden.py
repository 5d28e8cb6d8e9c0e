"""Extract the golden fixtures under tests/golden from the reference tree.

Run in the BUILD container only (``python tests/golden/make_golden.py``): it reads
``/root/reference/example`` (inputs and the outputs the reference itself shipped in
``example/example_results``) with the stdlib readers in ``oracle/hostio.py`` -- the
reference cannot be imported here (Python 2, pysam) and does not travel to the GPU
box, so the *bytes it shipped* are the pin.  Nothing is computed by the oracle in
this script: inputs are decoded, golden outputs are parsed, both are stored.

Outputs (np.savez_compressed):
  example_inputs.npz   chunk lists, per-chunk reads (pos, tlen), per-chunk sequence,
                       PWM, VMats, fragment sizes, occ_fit rows, the Scores bedgraph
                       slice and single_read.bam reads used by the reference's KATs
  example_golden.npz   example_results/*: occ x3 + nucleoatac_signal x2 tracks per chunk,
                       occpeaks / nucpos / nucpos.redundant rows, nuc_dist, fragmentsizes
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
STD_VPLOT = "/root/reference/nucleoatac/vplot/standard_vplot.VMat"
READ_MARGIN = 2600  # reads kept around each chunk: > pad (<=501) + upper (2000 for sizes)


def main():
    chrom_sizes_bam, frags = hostio.read_bam_fragments(EX + "example.bam")
    fa = hostio.Fasta(EX + "sacCer3.fa")
    chrs = fa.chrom_sizes()
    pwm, up, down, nucs = ra.read_pwm(PWM_PATH)
    vm_res, vlo, vhi = ra.read_vmat(RES + "example.VMat")  # what `nucleoatac run` feeds to nuc
    vm_ex, vxlo, vxhi = ra.read_vmat(EX + "example.VMat")  # used by tests/test_xcor.py, test_var.py
    std, slo, shi = ra.read_vmat(STD_VPLOT)
    fsz, flo, fhi = ra.read_sizes(RES + "example.fragmentsizes.txt")
    ndist, _, _ = ra.read_sizes(RES + "example.nuc_dist.txt")
    occ_fit = np.loadtxt(RES + "example.occ_fit.txt")

    upper, flank, nuc_sep = 251, 60, 120
    raw = ra.read_bed_chunks(EX + "example.bed")
    occ_chunks = ra.merge_chunks(ra.slop_chunks(ra.read_bed_chunks(
        EX + "example.bed", chrs, min_offset=flank + upper // 2 + max(up, down) + nuc_sep // 2), chrs, 60, 60))
    nuc_chunks = ra.merge_chunks(ra.slop_chunks(ra.read_bed_chunks(
        EX + "example.bed", chrs, min_offset=vm_res.shape[1] + vhi // 2 + max(up, down) + nuc_sep // 2,
        min_length=2 * nuc_sep), chrs, 60, 60))
    assert occ_chunks == nuc_chunks
    chunks = occ_chunks
    chrom_names = sorted({c for c, _, _ in chunks})
    cidx = {c: i for i, c in enumerate(chrom_names)}

    # per-chunk reads and sequence
    pos_l, tlen_l, off = [], [], [0]
    seq_l, seq_off, seq_start = [], [0], []
    smargin = 2 * max(vm_res.shape[1], 121) + upper // 2 + max(up, down) + 8
    for c, s, e in chunks:
        p, t = frags[c]
        sel = (p >= s - READ_MARGIN) & (p < e + READ_MARGIN)
        pos_l.append(p[sel])
        tlen_l.append(t[sel])
        off.append(off[-1] + int(sel.sum()))
        seq = fa.fetch(c, s - smargin, e + smargin)
        assert len(seq) == e - s + 2 * smargin
        seq_l.append(np.frombuffer(seq.encode(), dtype=np.uint8))
        seq_off.append(seq_off[-1] + len(seq))
        seq_start.append(s - smargin)

    # reference KAT inputs: the raw first bed chunk, the Scores track over it, single_read.bam
    c0, s0, e0 = raw[0]
    scores_rows = hostio.read_bedgraph_gz(EX + "example.Scores.bedgraph.gz")
    scores0 = hostio.bedgraph_region(scores_rows, c0, s0, e0)
    _, single = hostio.read_bam_fragments(EX + "single_read.bam")
    p0, t0 = frags[c0]
    sel0 = (p0 >= s0 - READ_MARGIN) & (p0 < e0 + READ_MARGIN)

    np.savez_compressed(
        os.path.join(HERE, "example_inputs.npz"),
        chrom_names=np.array(chrom_names),
        chunk_chrom=np.array([cidx[c] for c, _, _ in chunks], dtype=np.int32),
        chunk_start=np.array([s for _, s, _ in chunks], dtype=np.int32),
        chunk_end=np.array([e for _, _, e in chunks], dtype=np.int32),
        frag_off=np.array(off, dtype=np.int64), frag_pos=np.concatenate(pos_l).astype(np.int32),
        frag_tlen=np.concatenate(tlen_l).astype(np.int32),
        seq_off=np.array(seq_off, dtype=np.int64), seq_start=np.array(seq_start, dtype=np.int32),
        seq=np.concatenate(seq_l),
        pwm=pwm, pwm_up=up, pwm_down=down, pwm_nucleotides=np.array(nucs),
        vmat=vm_res, vmat_lower=vlo, vmat_upper=vhi,
        vmat_example=vm_ex, vmat_example_lower=vxlo, vmat_example_upper=vxhi,
        std_vplot=std.astype(np.float32), std_vplot_lower=slo, std_vplot_upper=shi,
        fragmentsizes=fsz, occ_fit=occ_fit,
        raw0_chrom=c0, raw0_start=s0, raw0_end=e0, raw0_scores=scores0,
        raw0_pos=p0[sel0].astype(np.int32), raw0_tlen=t0[sel0].astype(np.int32),
        single_pos=single[c0][0], single_tlen=single[c0][1],
        n_kept_reads=sum(len(v[0]) for v in frags.values()),
    )

    # golden outputs
    tracks = {}
    for key, fn in [("occ", "occ"), ("occ_lower", "occ.lower_bound"), ("occ_upper", "occ.upper_bound"),
                    ("nuc_signal", "nucleoatac_signal"), ("nuc_smooth", "nucleoatac_signal.smooth")]:
        rows = hostio.read_bedgraph_gz(RES + "example.%s.bedgraph.gz" % fn)
        tracks[key] = np.concatenate([hostio.bedgraph_region(rows, c, s, e) for c, s, e in chunks])

    def bed_table(path, ncol):
        rows = hostio.read_bedgraph_gz(path)
        chrom = np.array([cidx[r[0]] for r in rows], dtype=np.int32)
        pos = np.array([int(r[1]) for r in rows], dtype=np.int32)
        vals = np.array([[float(x) for x in r[3:3 + ncol]] for r in rows], dtype=np.float64)
        return chrom, pos, vals

    oc, op, ov = bed_table(RES + "example.occpeaks.bed.gz", 4)
    nc_, np_, nv = bed_table(RES + "example.nucpos.bed.gz", 10)
    rc, rp, rv = bed_table(RES + "example.nucpos.redundant.bed.gz", 10)
    merged = hostio.read_bedgraph_gz(RES + "example.nucmap_combined.bed.gz")
    occ_text = "".join("\t".join(r) + "\n" for r in hostio.read_bedgraph_gz(RES + "example.occpeaks.bed.gz"))
    nuc_text = "".join("\t".join(r) + "\n" for r in hostio.read_bedgraph_gz(RES + "example.nucpos.bed.gz"))
    merged_text = "".join("\t".join(r) + "\n" for r in merged)
    np.savez_compressed(
        os.path.join(HERE, "example_golden.npz"),
        track_off=np.cumsum([0] + [e - s for _, s, e in chunks]).astype(np.int64),
        occpeaks_chrom=oc, occpeaks_pos=op, occpeaks_vals=ov,
        nucpos_chrom=nc_, nucpos_pos=np_, nucpos_vals=nv,
        redundant_chrom=rc, redundant_pos=rp, redundant_vals=rv,
        occpeaks_text=occ_text, nucpos_text=nuc_text, nucmap_combined_text=merged_text,
        nuc_dist=ndist, fragmentsizes=fsz, scores_706661=hostio.bedgraph_region(scores_rows, "chrII", 706661, 706662)[0],
        **tracks)
    for f in ("example_inputs.npz", "example_golden.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
