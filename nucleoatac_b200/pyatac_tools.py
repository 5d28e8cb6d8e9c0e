"""The pyatac tools that sit either side of the scoring path: `vplot`, `ins`, `cov`, `bias`, `sizes`
(pyatac/make_vplot.py, get_ins.py, get_cov.py, make_bias_track.py, get_sizes.py).  Thin drivers: BED / BAM / FASTA decode
on the host, the binning / scoring arithmetic on the device, outputs written like the reference's (bedgraph rows ->
bgzip + tabix; VMat / fragment-size text files).  Regions shard round-robin over the ranks like `occ` / `nuc`; the only
exchange is the final V-plot / histogram sum.  Plotting (`--no_plot` is implied) is out of scope."""
import os

import numpy as np

from . import dist, hostio
from .bias import InsertionBiasTrack, PWM
from .chunk import ChunkList
from .engine import default_engine
from .fragments import _bam, fetch_reads, getAllFragmentSizes, getFragmentSizesFromChunkList
from .fragmentsizes import FragmentSizes
from .tracks import CoverageTrack, InsertionTrack
from .utils import read_chrom_sizes_from_bam, read_chrom_sizes_from_fasta
from .VMat import VMat


def _basename(path):
    return ".".join(os.path.basename(path).split(".")[0:-1])


def _rank_world(args):
    return getattr(args, "rank", 0), getattr(args, "world", 1)


def _finish(path_plain, path_gz):
    hostio.bgzip_tabix(path_plain, path_gz)  # pysam.tabix_compress + tabix_index(preset="bed")
    os.remove(path_plain)


# ----------------------------------------------------------------------------- pyatac vplot
def vplot_sum(chunks, bam, flank, lower, upper, atac=True, scale=False, device=0, sites_per_call=4096):
    """Sum of `_vplotHelper` over `chunks` (pyatac/make_vplot.py:22-43): every site is centred (Chunk.center), its reads
    are fetched like FragmentMat2D.makeFragmentMat does for [centre - flank - 1, centre + 1 + flank), and the per-site
    insert-size x position matrices (flipped for strand "-") are accumulated on the device in one launch per
    `sites_per_call` sites."""
    eng = default_engine(device)
    total = np.zeros((upper - lower, 2 * flank + 1))
    for i0 in range(0, len(chunks), sites_per_call):
        sites = [ch.center(new=True) for ch in chunks[i0:i0 + sites_per_call]]
        off, pos, tlen = _bam(bam).fetch_fragments_many([(c.chrom, max(0, c.start - flank - 1 - upper), c.end + flank + upper)
                                                         for c in sites])
        centers = [c.start for c in sites]
        flips = [1 if c.strand == "-" else 0 for c in sites]
        total += eng.vplot(centers, flips, off, pos, tlen, flank, lower, upper, atac, scale)
    return total


def make_vplot(args):
    """`pyatac vplot` (pyatac/make_vplot.py:60-89)."""
    if not args.out:
        args.out = _basename(args.bed)
    rank, world = _rank_world(args)
    chunks = ChunkList.read(args.bed, strand_col=args.strand)
    mine = ChunkList(*dist.shard(chunks, rank, world))
    result = vplot_sum(mine, args.bam, args.flank, args.lower, args.upper, args.atac, args.scale,
                       device=getattr(args, "device", 0))
    result = dist.allreduce_sum(result, world)  # make_vplot.py:70-73: sum of the per-set matrices
    vmat = VMat(result, args.lower, args.upper)
    if rank == 0:
        vmat.save(args.out + ".VMat")
    dist.barrier(world)
    return vmat


# ----------------------------------------------------------------------------- pyatac ins / cov / bias
def _regions(args, chrs, bases, splitsize):
    """The chunk list of the track tools (get_ins.py:72-78, get_cov.py:60-66, make_bias_track.py:62-68)."""
    if args.bed is None:
        return ChunkList.convertChromSizes(chrs, splitsize=splitsize)
    chunks = ChunkList.read(args.bed)
    chunks.checkChroms(list(chrs.keys()))
    chunks.merge()
    return chunks


def _write_tracks(args, suffix, chunks, make_track):
    """Per-chunk tracks -> <out><suffix>.bedgraph.gz (+ .tbi), chunk order preserved across the ranks."""
    rank, world = _rank_world(args)
    path = args.out + suffix + ".bedgraph"
    writer = dist.ShardWriter(path, rank, world)
    bw = dist.BatchWriter()

    def write_group(tracks):   # formatted on the pool, written in chunk order, behind the next group's device calls
        for text in bw.map(lambda t: t.format_track(), tracks):
            writer.write_bytes(text)
            writer.end_chunk()

    group = []
    try:
        for chunk in dist.shard(chunks, rank, world):
            try:
                group.append(make_track(chunk))
            except Exception:
                print("Caught exception when processing:\n" + chunk.asBed() + "\n")
                raise
            if len(group) >= 64:
                bw.submit(write_group, group)
                group = []
        if group:
            bw.submit(write_group, group)
    finally:
        bw.close()
    writer.close()
    dist.barrier(world)
    if rank == 0:
        dist.ShardWriter.merge(path, world, len(chunks))
        _finish(path, path + ".gz")
    dist.barrier(world)


def get_ins(args, bases=50000, splitsize=1000):
    """`pyatac ins` (pyatac/get_ins.py:62-101)."""
    if not args.out:
        args.out = _basename(args.bam if args.bed is None else args.bed)
    chunks = _regions(args, read_chrom_sizes_from_bam(args.bam), bases, splitsize)

    def one(chunk):
        if args.smooth:  # _insHelperSmooth, get_ins.py:20-33
            offset = args.smooth // 2
            ins = InsertionTrack(chunk.chrom, chunk.start - offset, chunk.end + offset)
            ins.calculateInsertions(args.bam, lower=args.lower, upper=args.upper, atac=args.atac)
            ins.smooth_track(args.smooth, window="gaussian", mode="valid")
        else:            # _insHelper, get_ins.py:36-46
            ins = InsertionTrack(chunk.chrom, chunk.start, chunk.end)
            ins.calculateInsertions(args.bam, lower=args.lower, upper=args.upper, atac=args.atac)
        return ins

    _write_tracks(args, ".ins", chunks, one)


def coverage_track(chunk, bam, lower, upper, window, scale, atac=True, device=0):
    """`_covHelper` (pyatac/get_cov.py:22-38): windowed count of fragment centres, times scale / window."""
    half = window // 2
    pos, tlen = fetch_reads(bam, chunk.chrom, chunk.start - half - upper, chunk.end + half + upper)
    cov = CoverageTrack(chunk.chrom, chunk.start, chunk.end)
    cov.vals = default_engine(device).coverage(pos, tlen, chunk.start, chunk.end, lower, upper, window, atac)
    cov.vals *= scale / float(window)
    return cov


def get_cov(args, bases=50000, splitsize=1000):
    """`pyatac cov` (pyatac/get_cov.py:54-89)."""
    if not args.out:
        args.out = _basename(args.bam if args.bed is None else args.bed)
    chunks = _regions(args, read_chrom_sizes_from_bam(args.bam), bases, splitsize)
    _write_tracks(args, ".cov", chunks,
                  lambda c: coverage_track(c, args.bam, args.lower, args.upper, args.window, args.scale, args.atac,
                                           getattr(args, "device", 0)))


def make_bias_track(args, bases=500000, splitsize=1000):
    """`pyatac bias` (pyatac/make_bias_track.py:55-90)."""
    if args.out is None:
        args.out = _basename(args.bed if args.bed is not None else args.fasta)
    chrs = read_chrom_sizes_from_fasta(args.fasta)
    pwm = PWM.open(args.pwm)
    chunks = _regions(args, chrs, bases, splitsize)

    def one(chunk):  # _biasHelper, make_bias_track.py:19-30
        bias = InsertionBiasTrack(chunk.chrom, chunk.start, chunk.end)
        bias.computeBias(args.fasta, chrs, pwm)
        return bias

    _write_tracks(args, ".Scores", chunks, one)


# ----------------------------------------------------------------------------- pyatac sizes
def get_sizes(args):
    """`pyatac sizes` (pyatac/get_sizes.py:15-38; FragmentSizes.calculateSizes, fragmentsizes.py:31-37)."""
    if args.out is None:
        args.out = _basename(args.bam)
    rank, world = _rank_world(args)
    if args.bed:
        chunks = ChunkList.read(args.bed)
        chunks.merge()
        counts = np.asarray(getFragmentSizesFromChunkList(dist.shard(chunks, rank, world), args.bam, args.lower, args.upper,
                                                          args.atac), dtype=np.float64)
        counts = dist.allreduce_sum(counts, world)
    else:
        counts = getAllFragmentSizes(args.bam, args.lower, args.upper, args.atac)
    total = np.sum(counts)
    sizes = FragmentSizes(args.lower, args.upper, atac=args.atac, vals=counts / (total + (total == 0)))
    if rank == 0:
        sizes.save(args.out + ".fragmentsizes.txt")
    dist.barrier(world)
    return sizes


# ----------------------------------------------------------------------------- command line
def build_parser():
    """The reference's flags for these five commands (pyatac/cli.py:115-135,137-150,202-231,314-353)."""
    import argparse
    p = argparse.ArgumentParser(prog="pyatac", description="pyatac tools of nucleoatac_b200 (vplot, ins, cov, bias, sizes)")
    sub = p.add_subparsers(dest="command")

    sp = sub.add_parser("sizes")
    sp.add_argument("--bam", required=True)
    sp.add_argument("--bed")
    sp.add_argument("--out")
    sp.add_argument("--not_atac", action="store_false", dest="atac", default=True)
    sp.add_argument("--lower", default=0, type=int)
    sp.add_argument("--upper", default=500, type=int)
    sp.add_argument("--no_plot", action="store_true", default=False)

    sp = sub.add_parser("bias")
    sp.add_argument("--fasta", required=True)
    sp.add_argument("--pwm", default="Human")
    sp.add_argument("--bed")
    sp.add_argument("--out")
    sp.add_argument("--cores", default=1, type=int)

    sp = sub.add_parser("vplot")
    sp.add_argument("--bed", required=True)
    sp.add_argument("--bam", required=True)
    sp.add_argument("--out")
    sp.add_argument("--cores", default=1, type=int)
    sp.add_argument("--lower", default=0, type=int)
    sp.add_argument("--upper", default=250, type=int)
    sp.add_argument("--flank", default=250, type=int)
    sp.add_argument("--scale", action="store_true", default=False)
    sp.add_argument("--weight", type=int)
    sp.add_argument("--strand", type=int)
    sp.add_argument("--not_atac", action="store_false", dest="atac", default=True)
    sp.add_argument("--no_plot", action="store_true", default=False)
    sp.add_argument("--plot_extra", action="store_true", default=False)

    for name in ("ins", "cov"):
        sp = sub.add_parser(name)
        sp.add_argument("--bam", required=True)
        sp.add_argument("--bed")
        sp.add_argument("--out")
        sp.add_argument("--cores", default=1, type=int)
        sp.add_argument("--lower", default=0, type=int)
        sp.add_argument("--upper", default=2000, type=int)
        sp.add_argument("--not_atac", action="store_false", dest="atac", default=True)
        if name == "ins":
            sp.add_argument("--smooth", type=int)
        else:
            sp.add_argument("--window", type=int, default=121)
            sp.add_argument("--scale", type=float, default=10)
    for sp in sub.choices.values():
        sp.add_argument("--device", default=0, type=int, help="CUDA device of this process")
        sp.add_argument("--rank", default=0, type=int, help="shard index: this process takes regions k with k %% world == rank")
        sp.add_argument("--world", default=1, type=int, help="number of shards (one process per GPU)")
    return p


def pyatac_main(argv=None):
    args = build_parser().parse_args(argv)
    if getattr(args, "world", 1) == 1 and int(os.environ.get("WORLD_SIZE", "1")) > 1:  # launched by torchrun
        args.rank, args.world, args.device = dist.init_from_env()   # binds this process to cuda:LOCAL_RANK before NCCL starts
    cmds = dict(sizes=get_sizes, bias=make_bias_track, vplot=make_vplot, ins=get_ins, cov=get_cov)
    if args.command not in cmds:
        build_parser().print_help()
        return 2
    cmds[args.command](args)
    return 0
