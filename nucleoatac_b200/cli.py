"""Command line: `python -m nucleoatac_b200 occ|nuc|vprocess|merge|nfr|run ...` with the reference's flags (nucleoatac/cli.py:96-125,
203-240) plus `--gpus`, `--batch` and `--xcor_mode`.  `--cores N` (N > 1) caps the host helper threads / worker processes (the device replaces the scoring pool)."""
import argparse
import os
import time


def build_parser():
    p = argparse.ArgumentParser(prog="nucleoatac", description="B200-native occ / nuc scoring path of NucleoATAC")
    sub = p.add_subparsers(dest="command")
    occ = sub.add_parser("occ", help="nucleoatac function:  Call nucleosome occupancy")
    g = occ.add_argument_group("Required", "Necessary arguments")
    g.add_argument("--bed", metavar="bed_file", required=True, help="Peaks in bed format")
    g.add_argument("--bam", metavar="bam_file", required=True, help="Sorted (and indexed) BAM file")
    g.add_argument("--out", metavar="basename", required=True, help="give output basename")
    g = occ.add_argument_group("Bias calculation information", "Highly recommended. If fasta is not provided, will not calculate bias")
    g.add_argument("--fasta", metavar="genome_seq", help="Indexed fasta file")
    g.add_argument("--pwm", metavar="Tn5_PWM", default="Human", help="PWM descriptor file. Default is Human.PWM.txt included in package")
    g = occ.add_argument_group("General Options", "")
    g.add_argument("--sizes", metavar="fragmentsizes_file", help="File with fragment size distribution.  Use if don't want calculation of fragment size")
    g.add_argument("--cores", metavar="int", default=1, type=int, help="Number of cores to use: chunks are scored on the GPU; a value above 1 caps the host helpers (decode / format / deflate threads, fit workers)")
    g = occ.add_argument_group("Occupancy parameter", "Change with caution")
    g.add_argument("--upper", metavar="int", default=251, type=int, help="upper limit in insert size. default is 251")
    g.add_argument("--flank", metavar="int", default=60, type=int, help="Distance on each side of dyad to include for local occ calculation. Default is 60.")
    g.add_argument("--min_occ", metavar="float", default=0.1, type=float, help="Occupancy cutoff for determining nucleosome distribution. Default is 0.1")
    g.add_argument("--nuc_sep", metavar="int", default=120, type=int, help="minimum separation between occupany peaks. Default is 120.")
    g.add_argument("--confidence_interval", metavar="float", default=0.9, type=float, help="confidence interval level for lower and upper bounds.  default is 0.9, should be between 0 and 1")
    g.add_argument("--step", metavar="int", default=5, type=int, help="step size along genome for comuting occ. Default is 5.  Should be odd, or will be subtracted by 1")
    nuc = sub.add_parser("nuc", help="nucleoatac function:  Call nucleosome positions and make signal tracks")
    g = nuc.add_argument_group("Required", "Necessary arguments")
    g.add_argument("--bed", metavar="bed_file", required=True, help="Regions for which to do stuff.")
    g.add_argument("--vmat", metavar="vdensity_file", required=True, help="VMat object")
    g.add_argument("--bam", metavar="bam_file", required=True, help="Accepts sorted BAM file")
    g.add_argument("--out", metavar="basename", required=True, help="give output basename")
    g = nuc.add_argument_group("Bias options", "If --fasta not provided, bias not calculated")
    g.add_argument("--fasta", metavar="genome_seq", help="Indexed fasta file")
    g.add_argument("--pwm", metavar="Tn5_PWM", default="Human", help="PWM descriptor file. Default is Human.PWM.txt included in package")
    g = nuc.add_argument_group("General options", "")
    g.add_argument("--sizes", metavar="fragmentsizes_file", help="File with fragment size distribution.  Use if don't want calculation of fragment size")
    g.add_argument("--occ_track", metavar="occ_file", help="bgzip compressed bedgraph file with occcupancy track. Otherwise occ not determined for nuc positions.")
    g.add_argument("--cores", metavar="num_cores", default=1, type=int, help="Number of cores to use: chunks are scored on the GPU; a value above 1 caps the host helpers (decode / format / deflate threads, fit workers)")
    g.add_argument("--write_all", action="store_true", default=False, help="write all tracks")
    g.add_argument("--not_atac", dest="atac", action="store_false", default=True, help="data is not atac-seq")
    g = nuc.add_argument_group("Nucleosome calling parameters", "Change with caution")
    g.add_argument("--min_z", metavar="float", default=3, type=float, help="Z-score threshold for nucleosome calls. Default is 3")
    g.add_argument("--min_lr", metavar="float", default=0, type=float, help="Log likelihood ratio threshold for nucleosome calls. Default is 0")
    g.add_argument("--nuc_sep", metavar="int", default=120, type=int, help="Minimum separation between non-redundant nucleosomes. Default is 120")
    g.add_argument("--redundant_sep", metavar="int", default=25, type=int, help="Minimum separation between redundant nucleosomes. Not recommended to be below 15. Default is 25")
    g.add_argument("--sd", metavar="int", default=10, type=int, help="Standard deviation for smoothing. Default is 10")
    g.add_argument("--xcor_mode", default=0, type=int, help="0 auto (tcgen05), 1 fp64 CUDA cores, 2 tcgen05 tensor cores")
    vp = sub.add_parser("vprocess", help="nucleoatac function:  Make processed vplot to use for nucleosome calling")
    g = vp.add_argument_group("Required", "Necessary arguments")
    g.add_argument("--out", metavar="output_basename", required=True)
    g = vp.add_argument_group("VPlot and Insert Size Options", "Optional")
    g.add_argument("--sizes", metavar="file", help="Insert distribution file")
    from .run_vprocess import DEFAULT_VPLOT
    g.add_argument("--vplot", metavar="vmat_file", default=DEFAULT_VPLOT, help="Accepts VMat file.  Default is Vplot from S. Cer.")
    g = vp.add_argument_group("Size parameers", "Use sensible values")
    g.add_argument("--lower", metavar="int", default=105, type=int, help="lower limit (inclusive) in insert size. default is 105")
    g.add_argument("--upper", metavar="int", default=251, type=int, help="upper limit (exclusive) in insert size. default 251")
    g.add_argument("--flank", metavar="int", default=60, type=int, help="distance on each side of dyad to include")
    g = vp.add_argument_group("Options", "")
    g.add_argument("--smooth", metavar="float", default=0.75, type=float, help="SD to use for gaussian smoothing.  Use 0 for no smoothing.")
    g.add_argument("--plot_extra", action="store_true", default=False, help="accepted for compatibility; plotting is not part of this package")
    mg = sub.add_parser("merge", help="nucleoatac function: Merge occ and nuc calls")
    g = mg.add_argument_group("Required", "Necessary arguments")
    g.add_argument("--occpeaks", metavar="occpeaks_file", required=True, help="Output from occ utility")
    g.add_argument("--nucpos", metavar="nucpos_file", required=True, help="Output from nuc utility")
    g = mg.add_argument_group("Options", "optional")
    g.add_argument("--out", metavar="out_basename", help="output file basename")
    g.add_argument("--sep", metavar="min_separation", default=120, help="minimum separation between call")
    g.add_argument("--min_occ", metavar="min_occ", default=0.1, help="minimum lower bound occupancy of nucleosomes to be considered for excluding NFR. default is 0.1")
    nf = sub.add_parser("nfr", help="nucleoatac function: Call NFRs")
    g = nf.add_argument_group("Required", "Necessary arguments")
    g.add_argument("--bed", metavar="bed_file", required=True, help="Peaks in bed format")
    g.add_argument("--occ_track", metavar="occ_file", required=True, help="bgzip compressed, tabix-indexed bedgraph file with occcupancy track.")
    g.add_argument("--calls", metavar="nucpos_file", required=True, help="bed file with nucleosome center calls")
    g = nf.add_argument_group("Insertion track options", "Either input insertion track or bamfile")
    g.add_argument("--ins_track", metavar="ins_file", help="bgzip compressed, tabix-indexed bedgraph file with insertion track. will be generated if not included")
    g.add_argument("--bam", metavar="bam_file", help="Sorted (and indexed) BAM file")
    g = nf.add_argument_group("Bias calculation information", "Highly recommended. If fasta is not provided, will not calculate bias")
    g.add_argument("--fasta", metavar="genome_seq", help="Indexed fasta file")
    g.add_argument("--pwm", metavar="Tn5_PWM", default="Human", help="PWM descriptor file. Default is Human.PWM.txt included in package")
    g = nf.add_argument_group("General options", "optional")
    g.add_argument("--out", metavar="out_basename", help="output file basename")
    g.add_argument("--cores", metavar="num_cores", default=1, type=int, help="Number of cores to use (ignored)")
    g = nf.add_argument_group("NFR determination parameters")
    g.add_argument("--max_occ", metavar="float", default=0.1, type=float, help="Maximum mean occupancy for NFR. Default is 0.1")
    g.add_argument("--max_occ_upper", metavar="float", default=0.25, type=float, help="Maximum for minimum of  upper bound occupancy in NFR. Default is 0.25")
    rn = sub.add_parser("run", help="Main nucleoatac utility: occ, vprocess, nuc, merge and nfr in one go")
    g = rn.add_argument_group("Required", "Necessary arguments")
    g.add_argument("--bed", metavar="bed_file", required=True, help="Regions for which to do stuff.")
    g.add_argument("--bam", metavar="bam_file", required=True, help="Accepts sorted BAM file")
    g.add_argument("--out", metavar="output_basename", required=True, help="give output basename")
    g.add_argument("--fasta", metavar="genome_seq", required=True, help="Indexed fasta file")
    g = rn.add_argument_group("Options", "optional")
    g.add_argument("--pwm", metavar="Tn5_PWM", default="Human", help="PWM descriptor file. Default is Human.PWM.txt included in package")
    g.add_argument("--cores", metavar="num_cores", default=1, type=int, help="Number of cores to use (ignored)")
    g.add_argument("--write_all", action="store_true", default=False, help="write all tracks")
    rn.add_argument("--xcor_mode", default=0, type=int, help="0 auto (tcgen05), 1 fp64 CUDA cores, 2 tcgen05 tensor cores")
    for sp in (occ, nuc, nf, rn):
        g = sp.add_argument_group("Device options", "")
        g.add_argument("--device", default=0, type=int, help="CUDA device of this process")
        g.add_argument("--rank", default=0, type=int, help="shard index: this process scores chunks k with k %% world == rank")
        g.add_argument("--world", default=1, type=int, help="number of shards (one process per GPU)")
        g.add_argument("--batch", default=256, type=int, help="chunks per device batch")
    return p


def nucleoatac_main(argv=None):
    args = build_parser().parse_args(argv)
    t0 = time.time()
    if getattr(args, "cores", 1) > 1:   # the reference's pool size: an upper bound for the host helpers (threads, worker processes)
        os.environ.setdefault("NB200_HOST_WORKERS", str(args.cores))
    if getattr(args, "world", 1) == 1 and int(os.environ.get("WORLD_SIZE", "1")) > 1:  # launched by torchrun
        from . import dist
        args.rank, args.world, args.device = dist.init_from_env()   # binds this process to cuda:LOCAL_RANK before NCCL starts
    if args.command == "occ":
        print("---------Computing Occupancy and Nucleosomal Insert Distribution----")
        from .run_occ import run_occ
        run_occ(args)
    elif args.command == "nuc":
        print("---------Obtaining nucleosome signal and calling positions----------")
        from .run_nuc import run_nuc
        run_nuc(args)
    elif args.command == "merge":
        print("---------Merging----------------------------------------------------")
        from .merge import run_merge
        run_merge(args)
    elif args.command == "nfr":
        print("---------Calling NFR positions--------------------------------------")
        from .run_nfr import run_nfr
        run_nfr(args)
    elif args.command == "run":  # nucleoatac/cli.py:34-64
        from . import dist
        p = build_parser()
        base = ["--bed", args.bed, "--bam", args.bam, "--fasta", args.fasta, "--pwm", args.pwm, "--out", args.out]
        # every sharded step runs on this process's shard and device; the host-only steps (vprocess, merge) run on rank 0
        # between barriers, so that no two ranks write the same file
        shard = ["--rank", str(args.rank), "--world", str(args.world), "--device", str(args.device), "--batch", str(args.batch)]
        print("---------Step1: Computing Occupancy and Nucleosomal Insert Distribution---------")
        from .run_occ import run_occ
        run_occ(p.parse_args(["occ"] + base + shard))
        print("---------Step2: Processing Vplot------------------------------------------------")
        if args.rank == 0:
            from .run_vprocess import run_vprocess
            run_vprocess(p.parse_args(["vprocess", "--sizes", args.out + ".nuc_dist.txt", "--out", args.out]))
        dist.barrier(args.world)
        print("---------Step3: Obtaining nucleosome signal and calling positions---------------")
        from .run_nuc import run_nuc
        run_nuc(p.parse_args(["nuc"] + base + shard + ["--occ_track", args.out + ".occ.bedgraph.gz", "--vmat", args.out + ".VMat",
                                                       "--sizes", args.out + ".fragmentsizes.txt", "--xcor_mode", str(args.xcor_mode)]
                             + (["--write_all"] if args.write_all else [])))
        print("---------Step4: Making combined nucleosome position map ------------------------")
        if args.rank == 0:
            from .merge import run_merge
            run_merge(p.parse_args(["merge", "--occpeaks", args.out + ".occpeaks.bed.gz", "--nucpos", args.out + ".nucpos.bed.gz",
                                    "--out", args.out]))
        dist.barrier(args.world)
        print("---------Step5: Calling NFR positions-------------------------------------------")
        from .run_nfr import run_nfr
        run_nfr(p.parse_args(["nfr", "--bed", args.bed, "--occ_track", args.out + ".occ.bedgraph.gz", "--calls",
                              args.out + ".nucmap_combined.bed.gz", "--out", args.out, "--fasta", args.fasta, "--pwm", args.pwm,
                              "--bam", args.bam] + shard))
    elif args.command == "vprocess":
        print("---------Processing VPlot-----------------------------------------")
        from .run_vprocess import run_vprocess
        run_vprocess(args)
    else:
        build_parser().print_help()
        return 2
    print("nucleoatac %s done in %.1f s" % (args.command, time.time() - t0))
    return 0
