"""`nucleoatac nfr` driver (nucleoatac/run_nfr.py:72-127): BED -> merged chunks -> NFRChunk.process per chunk ->
<out>.nfrpos.bed.gz (+ <out>.ins.bedgraph.gz when the insertion track is computed from the BAM), bgzip'd and tabix-indexed.
Chunks shard round-robin over the ranks like `occ` / `nuc`; outputs are merged back in chunk order by rank 0."""
import os

from . import dist, hostio
from .bias import PWM
from .chunk import ChunkList
from .NFRCalling import NFRChunk, NFRParameters
from .utils import read_chrom_sizes_from_bam, read_chrom_sizes_from_fasta


def _finish(path_plain, path_gz):
    hostio.bgzip_tabix(path_plain, path_gz)  # pysam.tabix_compress + tabix_index(preset="bed"), run_nfr.py:120-127
    os.remove(path_plain)


def run_nfr(args):
    if args.bam is None and args.ins_track is None:
        raise Exception("Must supply either bam file or insertion track")
    if not args.out:
        args.out = ".".join(os.path.basename(args.calls).split(".")[0:-3])
    rank, world = getattr(args, "rank", 0), getattr(args, "world", 1)
    if args.fasta is not None:
        chrs_fasta = read_chrom_sizes_from_fasta(args.fasta)
        pwm = PWM.open(args.pwm)
        chunks = ChunkList.read(args.bed, chromDict=chrs_fasta, min_offset=max(pwm.up, pwm.down))
    else:
        chunks = ChunkList.read(args.bed)
    if args.bam is not None:
        chunks.checkChroms(read_chrom_sizes_from_bam(args.bam), chrom_source="BAM file")
    chunks.merge()
    params = NFRParameters(args.occ_track, args.calls, args.ins_track, args.bam, max_occ=args.max_occ,
                           max_occ_upper=args.max_occ_upper, fasta=args.fasta, pwm=args.pwm)
    nfr_writer = dist.ShardWriter(args.out + ".nfrpos.bed", rank, world)
    ins_writer = dist.ShardWriter(args.out + ".ins.bedgraph", rank, world) if params.ins_track is None else None
    for chunk in dist.shard(chunks, rank, world):
        nfr = NFRChunk(chunk)
        try:
            nfr.process(params)
        except Exception:
            print("Caught exception when processing:\n" + chunk.asBed() + "\n")
            raise
        for region in nfr.nfrs:
            region.write(nfr_writer)
        nfr_writer.end_chunk()
        if ins_writer is not None:
            nfr.ins.write_track(ins_writer)
            ins_writer.end_chunk()
        nfr.removeData()
    nfr_writer.close()
    if ins_writer is not None:
        ins_writer.close()
    dist.barrier(world)
    if rank == 0:
        dist.ShardWriter.merge(args.out + ".nfrpos.bed", world, len(chunks))
        _finish(args.out + ".nfrpos.bed", args.out + ".nfrpos.bed.gz")
        if ins_writer is not None:
            dist.ShardWriter.merge(args.out + ".ins.bedgraph", world, len(chunks))
            _finish(args.out + ".ins.bedgraph", args.out + ".ins.bedgraph.gz")
    dist.barrier(world)
