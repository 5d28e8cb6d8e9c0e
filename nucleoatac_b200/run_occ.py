"""`nucleoatac occ` driver (nucleoatac/run_occ.py:77-148): BED -> chunks -> fragment-size model -> batches of chunks
through the device -> bedgraph / bed writers.  The reference's multiprocessing pool + writer processes become: one
process per GPU, chunk k of the list -> GPU k mod N (round-robin), results written in chunk order."""
import os

import numpy as np

from . import dist, hostio
from .bias import PWM
from .chunk import ChunkList
from .fragments import getFragmentSizesFromChunkList
from .fragmentsizes import FragmentSizes
from .Occupancy import FragmentMixDistribution, OccChunk, OccupancyParameters, process_chunks
from .utils import read_chrom_sizes_from_bam, read_chrom_sizes_from_fasta


def _finish(path_plain, path_gz):
    hostio.bgzip_tabix(path_plain, path_gz)  # pysam.tabix_compress + tabix_index(preset="bed"), run_occ.py:130-136
    os.remove(path_plain)


def occ_chunks(args):
    chrs = read_chrom_sizes_from_fasta(args.fasta) if args.fasta else read_chrom_sizes_from_bam(args.bam)
    pwm = PWM.open(args.pwm)
    chunks = ChunkList.read(args.bed, chromDict=chrs,
                            min_offset=args.flank + args.upper // 2 + max(pwm.up, pwm.down) + args.nuc_sep // 2)
    chunks.slop(chrs, up=args.nuc_sep // 2, down=args.nuc_sep // 2)
    chunks.merge()
    return chunks


def run_occ(args, score=process_chunks, count_sizes=getFragmentSizesFromChunkList):
    """`score(list of OccChunk, params)` fills the chunks (default: the device path); `count_sizes(chunks, bam, lower,
    upper)` is the fragment-size histogram of a chunk list (default: host BAM decode + device binning)."""
    rank, world = getattr(args, "rank", 0), getattr(args, "world", 1)
    chunks = occ_chunks(args)
    fragment_dist = FragmentMixDistribution(0, upper=args.upper)
    if args.sizes is not None:
        tmp = FragmentSizes.open(args.sizes)
        fragment_dist.fragmentsizes = FragmentSizes(0, args.upper, vals=tmp.get(0, args.upper))
    else:  # fragments.pyx:122-145, one shard per rank, summed exactly (integer-valued counts)
        counts = dist.allreduce_sum(np.asarray(count_sizes(dist.shard(chunks, rank, world), args.bam, 0, args.upper), dtype=np.float64), world)
        total = np.sum(counts)
        fragment_dist.fragmentsizes = FragmentSizes(0, args.upper, vals=counts / (total + (total == 0)))
    fragment_dist.modelNFR()
    if rank == 0:
        fragment_dist.plotFits(args.out + ".occ_fit.eps")
        fragment_dist.fragmentsizes.save(args.out + ".fragmentsizes.txt")
    params = OccupancyParameters(fragment_dist, args.upper, args.fasta, args.pwm, sep=args.nuc_sep, min_occ=args.min_occ,
                                 flank=args.flank, bam=args.bam, ci=args.confidence_interval, step=args.step,
                                 device=getattr(args, "device", 0))
    mine = ChunkList(*dist.shard(chunks, rank, world))
    names = ("occ", "occ.lower_bound", "occ.upper_bound")
    writers = [dist.ShardWriter(args.out + "." + n + ".bedgraph", rank, world) for n in names]
    peaks_writer = dist.ShardWriter(args.out + ".occpeaks.bed", rank, world)
    nuc_dist = np.zeros(args.upper)
    batch = max(1, getattr(args, "batch", 256))
    bw = dist.BatchWriter()

    def write_batch(occs):   # behind the scoring of the next batch; host data only
        jobs = [(oc.occ, v) for oc in occs for v in (oc.occ.smoothed_vals, oc.occ.smoothed_lower, oc.occ.smoothed_upper)]
        texts = bw.map(lambda j: j[0].format_track(vals=j[1]), jobs)
        for i, oc in enumerate(occs):
            nuc_dist[:] += oc.getNucDist()
            for t in range(3):
                writers[t].write_bytes(texts[3 * i + t])
            for k in sorted(oc.peaks.keys()):
                oc.peaks[k].write(peaks_writer)
            for w in writers + [peaks_writer]:
                w.end_chunk()
            oc.removeData()

    try:
        for group in mine.split(items=batch):
            occs = [OccChunk(c) for c in group]
            try:
                score(occs, params)
            except Exception:
                print("Caught exception when processing:\n" + ChunkList(*group).asBed() + "\n")
                raise
            bw.submit(write_batch, occs)
    finally:
        bw.close()
    for w in writers + [peaks_writer]:
        w.close()
    nuc_dist = dist.allreduce_sum(nuc_dist, world)  # run_occ.py:117-121 summed over the shards
    dist.barrier(world)
    if rank == 0:
        dist.ShardWriter.merge(args.out + ".occpeaks.bed", world, len(chunks))
        _finish(args.out + ".occpeaks.bed", args.out + ".occpeaks.bed.gz")
        for n in names:
            dist.ShardWriter.merge(args.out + "." + n + ".bedgraph", world, len(chunks))
            _finish(args.out + "." + n + ".bedgraph", args.out + "." + n + ".bedgraph.gz")
        FragmentSizes(0, args.upper, vals=nuc_dist).save(args.out + ".nuc_dist.txt")
    dist.barrier(world)
    return nuc_dist
