"""`nucleoatac nuc` driver (nucleoatac/run_nuc.py:141-201): VMat + fragment sizes -> batches of chunks through the
device -> nucpos / signal writers, in chunk order; chunk k -> GPU k mod N."""
import os

from . import dist, hostio
from .bias import PWM
from .chunk import ChunkList
from .fragmentsizes import FragmentSizes
from .NucleosomeCalling import NucChunk, NucParameters, process_chunks
from .utils import read_chrom_sizes_from_bam, read_chrom_sizes_from_fasta
from .VMat import VMat


def nuc_chunks(args, vmat):
    chrs = read_chrom_sizes_from_fasta(args.fasta) if args.fasta else read_chrom_sizes_from_bam(args.bam)
    pwm = PWM.open(args.pwm)
    chunks = ChunkList.read(args.bed, chromDict=chrs,
                            min_offset=vmat.mat.shape[1] + vmat.upper // 2 + max(pwm.up, pwm.down) + args.nuc_sep // 2,
                            min_length=args.nuc_sep * 2)
    chunks.slop(chrs, up=args.nuc_sep // 2, down=args.nuc_sep // 2)
    chunks.merge()
    return chunks


def run_nuc(args, score=process_chunks):
    rank, world = getattr(args, "rank", 0), getattr(args, "world", 1)
    vmat = VMat.open(args.vmat)
    chunks = nuc_chunks(args, vmat)
    if args.sizes is not None:
        fragment_dist = FragmentSizes.open(args.sizes)
    else:
        fragment_dist = FragmentSizes(0, upper=vmat.upper)
        fragment_dist.calculateSizes(args.bam, chunks)
    params = NucParameters(vmat=vmat, fragmentsizes=fragment_dist, bam=args.bam, fasta=args.fasta, pwm=args.pwm,
                           occ_track=args.occ_track, sd=args.sd, nonredundant_sep=args.nuc_sep,
                           redundant_sep=args.redundant_sep, min_z=args.min_z, min_lr=args.min_lr, atac=args.atac,
                           device=getattr(args, "device", 0), xcor_mode=getattr(args, "xcor_mode", 0))
    outputs = ["nucpos", "nucpos.redundant", "nucleoatac_signal", "nucleoatac_signal.smooth"]
    if args.write_all:
        outputs += ["nucleoatac_background", "nucleoatac_raw"]
    ext = lambda n: ".bed" if n.startswith("nucpos") else ".bedgraph"
    handles = {n: dist.ShardWriter(args.out + "." + n + ext(n), rank, world) for n in outputs}
    mine = ChunkList(*dist.shard(chunks, rank, world))
    batch = max(1, getattr(args, "batch", 256))
    bw = dist.BatchWriter()
    tracks = [("nucleoatac_signal", "norm_signal"), ("nucleoatac_signal.smooth", "smoothed")]
    if args.write_all:
        tracks += [("nucleoatac_background", "bias"), ("nucleoatac_raw", "nuc_signal")]

    def write_batch(nucs):   # behind the scoring of the next batch; host data only.  run_nuc.py:30-32: what _nucHelper returns per chunk
        texts = bw.map(lambda j: getattr(j[0], j[1]).format_track(), [(nc, attr) for nc in nucs for _, attr in tracks])
        for i, nc in enumerate(nucs):
            for k in sorted(int(x) for x in nc.nonredundant):
                nc.nuc_collection[k].write(handles["nucpos"])
            for k in sorted(int(x) for x in nc.redundant):
                nc.nuc_collection[k].write(handles["nucpos.redundant"])
            for t, (name, _) in enumerate(tracks):
                handles[name].write_bytes(texts[len(tracks) * i + t])
            for h in handles.values():
                h.end_chunk()
            nc.removeData()

    try:
        for group in mine.split(items=batch):
            nucs = [NucChunk(c) for c in group]
            try:
                score(nucs, params)
            except Exception:
                print("Caught exception when processing:\n" + ChunkList(*group).asBed() + "\n")
                raise
            bw.submit(write_batch, nucs)
    finally:
        bw.close()
    for h in handles.values():
        h.close()
    dist.barrier(world)
    if rank == 0:
        for n in outputs:
            plain = args.out + "." + n + ext(n)
            dist.ShardWriter.merge(plain, world, len(chunks))
            hostio.bgzip_tabix(plain, plain + ".gz")  # tabix_compress + tabix_index, run_nuc.py:194-201
            os.remove(plain)
    dist.barrier(world)
