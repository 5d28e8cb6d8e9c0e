/* nucleo_b200.h -- C-ABI of libnucleo_b200.so: the B200-native per-chunk occ/nuc
 * scoring path of NucleoATAC (reference @ a56e741, v0.3.4).
 *
 * The reference has no FFI: its seam is a set of Python/Cython call signatures
 * (SURVEY.md 8b).  Every entry point below names the reference interface it
 * replaces (file:line under /root/reference).  All functions are extern "C",
 * take plain pointers and sizes, return 0 on success and a non-zero status
 * otherwise; nb200_last_error() then holds a message.  The caller owns every host
 * buffer; the library owns device memory behind the opaque handles.  There is no
 * CPU fallback: without a CUDA device nb200_ctx_create fails.
 *
 * Threading: one context per (process, GPU); calls on one context must be
 * serialised by the caller.  Batch calls are asynchronous on the batch's own
 * stream; host output buffers are valid after nb200_batch_sync().
 */
#ifndef NUCLEO_B200_H
#define NUCLEO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nb200_ctx nb200_ctx;
typedef struct nb200_dbatch nb200_dbatch;

#define NB200_OK 0
#define NB200_ERR_CUDA 1
#define NB200_ERR_ARG 2
#define NB200_ERR_STATE 3
#define NB200_ERR_CAPACITY 4
#define NB200_ERR_FLANK 5 /* "Insufficient flanking region..." exceptions of the reference */

/* ---- context --------------------------------------------------------------------------- */
int nb200_ctx_create(int device, nb200_ctx **out);
int nb200_ctx_destroy(nb200_ctx *ctx);
const char *nb200_last_error(nb200_ctx *ctx); /* ctx may be NULL: error of the last failed create */
int nb200_device_info(nb200_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, int64_t *hbm_bytes);
/* pinned host memory for the end-to-end path (numpy arrays are built on top of it) */
int nb200_host_alloc(nb200_ctx *ctx, int64_t bytes, void **out);
int nb200_host_free(nb200_ctx *ctx, void *p);

/* ---- run constants --------------------------------------------------------------------- */
/* PWM.open + np.log(pwm.mat), pyatac/bias.py:47-76,90.  log_pwm is [n_nuc][up+down+1] row-major,
 * nucleotides a string of n_nuc single letters (row order). */
int nb200_set_pwm(nb200_ctx *ctx, const double *log_pwm, int n_nuc, int up, int down, const char *nucleotides);
/* VMat(mat, lower, upper), pyatac/VMat.py:24-37.  mat is [upper-lower][ncol] row-major, ncol odd. */
int nb200_set_vmat(nb200_ctx *ctx, const double *mat, int nrow, int ncol, int lower, int upper);
/* FragmentSizes.get(0, upper), pyatac/fragmentsizes.py:28-43 (normByInsertDist operand). */
int nb200_set_fragment_sizes(nb200_ctx *ctx, const double *freq, int upper);
/* OccupancyCalcParams, nucleoatac/Occupancy.py:89-102: normalised nuc/nfr probabilities over
 * sizes [0,upper), the alpha grid (np.linspace(0,1,101)) and chi2.ppf(ci,1). */
int nb200_set_occ_model(nb200_ctx *ctx, const double *nuc_probs, const double *nfr_probs, int upper,
                        const double *alphas, int n_alpha, double cutoff);
/* RandomState(25).uniform(0,1e-12,n) of pyatac/utils.py:94-97 (prefix property: any n). */
int nb200_set_jitter(nb200_ctx *ctx, const double *jitter, int64_t n);

typedef struct {
    int32_t upper;     /* --upper (251)           nucleoatac/cli.py:113 */
    int32_t flank;     /* --flank (60)            :115 */
    int32_t step;      /* --step (5, forced odd)  :123, Occupancy.py:190-193 */
    int32_t sep;       /* --nuc_sep (120)         :119 */
    double min_occ;    /* --min_occ (0.1)         :117 */
    int32_t atac;      /* 1 = ATAC +4/-8 shift    fragments.pyx:26-31 */
    int32_t use_bias;  /* 0 = no --fasta: bias matrix all ones, Occupancy.py:209-211 */
    const double *smooth_win; /* signal.gaussian(2*flank+1, flank/3.), Occupancy.py:220 */
    int32_t smooth_len;
} nb200_occ_params;

typedef struct {
    int32_t atac;            /* NucParameters, nucleoatac/NucleosomeCalling.py:204-226 */
    int32_t use_bias;        /* 0 = no --fasta (:248) */
    int32_t smooth_sd;       /* --sd */
    int32_t nonredundant_sep;/* --nuc_sep */
    int32_t redundant_sep;   /* --redundant_sep */
    double min_z, min_lr, min_reads;
    const double *smooth_win; /* signal.gaussian(6*sd+1, sd), :276-283 */
    int32_t smooth_len;
    int32_t xcor_mode;       /* 0 = auto, 1 = fp64 CUDA-core xcor, 2 = tcgen05 split-fp16 xcor */
} nb200_nuc_params;

int nb200_occ_configure(nb200_ctx *ctx, const nb200_occ_params *p);
int nb200_nuc_configure(nb200_ctx *ctx, const nb200_nuc_params *p);

/* ---- primitives: the Cython / numpy seams, host buffers in and out ------------------------ */
/* makeFragmentMat(bam, chrom, start, end, lower, upper, atac), pyatac/fragments.pyx:17-40.  pos/tlen are
 * the BAM fields of the reads with is_proper_pair and not is_reverse; out is [upper-lower][end-start]. */
int nb200_fragmat_build(nb200_ctx *ctx, const int32_t *pos, const int32_t *tlen, int64_t n, int32_t start,
                        int32_t end, int32_t lower, int32_t upper, int32_t atac, double *out);
/* getInsertions, pyatac/fragments.pyx:43-67; out is [end-start]. */
int nb200_insertions(nb200_ctx *ctx, const int32_t *pos, const int32_t *tlen, int64_t n, int32_t start,
                     int32_t end, int32_t lower, int32_t upper, int32_t atac, double *out);
/* getFragmentSizesFromChunkList, pyatac/fragments.pyx:122-145, for one chunk list slice: counts[upper-lower]
 * (int64, exact) of fragments whose centre lies in [starts[c], ends[c]) of the chunk they were fetched for.
 * frag_off is [n_chunks+1]. */
int nb200_fragment_sizes(nb200_ctx *ctx, int32_t n_chunks, const int32_t *starts, const int32_t *ends,
                         const int64_t *frag_off, const int32_t *pos, const int32_t *tlen, int32_t lower,
                         int32_t upper, int32_t atac, int64_t *counts);
/* InsertionBiasTrack.computeBias arithmetic, pyatac/bias.py:87-92 + seq.py:37-45: out[len-(up+down)]. */
int nb200_bias_track(nb200_ctx *ctx, const uint8_t *seq, int64_t len, double *out);
/* BiasMat2D.makeBiasMat, pyatac/chunkmat2d.py:140-153: bias = log-bias over [mat_start-upper/2,
 * mat_end+upper/2) (n_bias values); out is [upper-lower][n_bias-(upper+(upper-1)%2)+1]. */
int nb200_biasmat_build(nb200_ctx *ctx, const double *bias, int64_t n_bias, int32_t lower, int32_t upper, double *out);
/* ChunkMat2D.getIns, pyatac/chunkmat2d.py:74-84 on a dense matrix [upper-lower][ncol]; out[ncol-plen+1]. */
int nb200_get_ins(nb200_ctx *ctx, const double *mat, int32_t lower, int32_t upper, int64_t ncol, double *out);
/* scipy.signal.correlate(mat, vmat, 'valid')[0] of NucleosomeCalling.py:34-36,60-63 against the VMat set with
 * nb200_set_vmat: mat is [vmat rows][ncol] dense f64, out[ncol-W+1].  fp64 CUDA-core kernel. */
int nb200_xcor_dense(nb200_ctx *ctx, const double *mat, int64_t ncol, double *out);
/* CoverageTrack.calculateCoverage inner part, pyatac/tracks.py:216-222: column sums of rows [row0,row1) of
 * a dense [nrow][ncol] matrix, flat window window_len, 'valid': out[ncol-window_len+1]. */
int nb200_coverage_dense(nb200_ctx *ctx, const double *mat, int32_t nrow, int64_t ncol, int32_t row0, int32_t row1,
                         int32_t window_len, double *out);
/* pyatac/utils.py:23-52 smooth(): w is the window (flat = ones, gaussian from scipy), mode_same 1/0,
 * norm 1/0 (NaN-aware normalisation).  out has n ('same') or n-wlen+1 ('valid') values. */
int nb200_smooth(nb200_ctx *ctx, const double *sig, int64_t n, const double *w, int32_t wlen, int32_t mode_same,
                 int32_t norm, double *out);
/* call_peaks, pyatac/utils.py:82-102 (needs nb200_set_jitter).  sig is updated in place (NaN -> min) like the
 * reference; peaks ascending in out_idx (capacity cap); *out_n receives the count. */
int nb200_call_peaks(nb200_ctx *ctx, double *sig, int64_t n, double min_signal, int32_t sep, int32_t boundary,
                     int32_t order, int32_t *out_idx, int32_t cap, int32_t *out_n);
/* reduce_peaks, pyatac/utils.py:56-78: keep[i] in {0,1} for ascending peaks[n] with scores sig[n]. */
int nb200_reduce_peaks(nb200_ctx *ctx, const int32_t *peaks, const double *sig, int32_t n, int32_t sep, int32_t *keep);
/* calculateOccupancy(inserts, bias, params), nucleoatac/Occupancy.py:104-120 with the model of
 * nb200_set_occ_model: out = {occ, lower, upper}. */
int nb200_calculate_occupancy(nb200_ctx *ctx, const double *inserts, const double *bias, int32_t n, double *out3);
/* calculateCov(p, v, r), nucleoatac/multinomial_cov.pyx:20-31 (closed form r*(sum p v^2-(sum p v)^2), fp64). */
int nb200_multinomial_cov(nb200_ctx *ctx, const double *p, const double *v, int64_t n, int32_t r, double *out);

/* ---- batched per-chunk paths: OccChunk.process / NucChunk.process ---------------------------- */
typedef struct {
    int32_t n_chunks;
    const int32_t *chunk_start; /* Chunk.start/end after slop+merge (run_occ.py:86-88, run_nuc.py:151-153) */
    const int32_t *chunk_end;
    const int64_t *frag_off;    /* [n_chunks+1] into frag_pos/frag_tlen: reads fetched for the chunk */
    const int32_t *frag_pos;    /* BAM pos of reads with is_proper_pair and not is_reverse */
    const int32_t *frag_tlen;
    const int64_t *seq_off;     /* [n_chunks+1] into seq, or NULL when no --fasta */
    const int32_t *seq_start;   /* genomic coordinate of the first base of each chunk's sequence slice */
    const uint8_t *seq;         /* bases (upper-cased on device, pyatac/seq.py:22) */
} nb200_batch;

/* H2D of a batch (async on the batch stream).  *io may hold a previous batch to recycle its buffers. */
int nb200_batch_upload(nb200_ctx *ctx, const nb200_batch *host, nb200_dbatch **io);
int nb200_batch_free(nb200_ctx *ctx, nb200_dbatch *b);
int nb200_batch_sync(nb200_ctx *ctx, nb200_dbatch *b);
int64_t nb200_batch_total_len(nb200_dbatch *b); /* sum of chunk lengths = length of every packed track */
int64_t nb200_batch_h2d_bytes(nb200_dbatch *b);

/* OccChunk.process + getNucDist on device, nucleoatac/Occupancy.py:241-248,232-240 (run_occ.py:23-39). */
int nb200_occ_run(nb200_ctx *ctx, nb200_dbatch *b);
/* NucChunk.process (without fit/getFuzz), nucleoatac/NucleosomeCalling.py:328-340. */
int nb200_nuc_run(nb200_ctx *ctx, nb200_dbatch *b);

typedef struct {
    /* packed tracks, chunk c at offset sum_{c'<c}(end-start); any pointer may be NULL to skip */
    double *smoothed_vals, *smoothed_lower, *smoothed_upper; /* what run_occ.py:45-49 writes */
    double *vals, *lower_bound, *upper_bound;                /* OccupancyTrack before smoothing */
    double *cov;                                             /* OccChunk.cov */
    double *nuc_dist;     /* [n_chunks][upper]: OccChunk.getNucDist() per chunk */
    int32_t *peak_count;  /* [n_chunks] */
    const int64_t *peak_off; /* [n_chunks+1] caller-chosen capacities (>= len/sep + 2 suffices) */
    int32_t *peak_pos;    /* genomic position, ascending within a chunk (run_occ.py:31) */
    double *peak_occ, *peak_lower, *peak_upper, *peak_reads; /* OccPeak, Occupancy.py:155-168 */
} nb200_occ_out;

typedef struct {
    double *nuc_signal, *background, *norm_signal, *smoothed; /* the 4 tracks of run_nuc.py:30-32 */
    double *nuc_cov, *nfr_cov;                                 /* optional */
    int32_t *cand_count;     /* [n_chunks]: candidates after call_peaks (NucleosomeCalling.py:299-301) */
    const int64_t *cand_off; /* [n_chunks+1] capacities (>= len/redundant_sep + 2 suffices) */
    int32_t *cand_pos;       /* genomic position, ascending */
    int32_t *cand_flag;      /* bit0 nuc_cov>min_reads, bit1 lr>min_lr, bit2 z>=min_z (kept), bit3 nonredundant */
    /* cand_lr / cand_z are the fp64 statistics of NucleosomeCalling.py:110-127 for every candidate with bit1 set (the
     * only ones the reference reports).  For a candidate rejected on the likelihood ratio (bit0 set, bit1 clear) cand_lr
     * holds the value of the stage that rejected it -- an upper bound of, or an fp32-normalised approximation to, its
     * LR, never above min_lr -- and cand_z is NaN (environment NB200_CS_SCREEN=0: exact LR for every candidate). */
    double *cand_z, *cand_lr, *cand_norm_signal, *cand_nuc_signal, *cand_nuc_cov, *cand_nfr_cov, *cand_smoothed;
} nb200_nuc_out;

int nb200_occ_download(nb200_ctx *ctx, nb200_dbatch *b, const nb200_occ_out *out);
int nb200_nuc_download(nb200_ctx *ctx, nb200_dbatch *b, const nb200_nuc_out *out);
int64_t nb200_occ_d2h_bytes(nb200_dbatch *b, const nb200_occ_out *out);
int64_t nb200_nuc_d2h_bytes(nb200_dbatch *b, const nb200_nuc_out *out);

/* The same results with the per-position tracks delivered as float32: they are converted on the device (round to nearest,
 * relative error <= 6e-8, well inside the 1e-5 the scoring path is specified to) and cross the host link at 4 bytes per
 * position and track instead of 8.  What the reference does with these tracks is print them (Track.write_track,
 * pyatac/tracks.py:37-74) and read single positions back (run_nuc.py --occ_track); peak / candidate tables, counts and
 * nuc_dist keep their types.  Fields as in nb200_occ_out / nb200_nuc_out. */
typedef struct {
    float *smoothed_vals, *smoothed_lower, *smoothed_upper;
    float *vals, *lower_bound, *upper_bound;
    float *cov;
    double *nuc_dist;
    int32_t *peak_count;
    const int64_t *peak_off;
    int32_t *peak_pos;
    double *peak_occ, *peak_lower, *peak_upper, *peak_reads;
} nb200_occ_out32;

typedef struct {
    float *nuc_signal, *background, *norm_signal, *smoothed;
    float *nuc_cov, *nfr_cov;
    int32_t *cand_count;
    const int64_t *cand_off;
    int32_t *cand_pos;
    int32_t *cand_flag;
    double *cand_z, *cand_lr, *cand_norm_signal, *cand_nuc_signal, *cand_nuc_cov, *cand_nfr_cov, *cand_smoothed;
} nb200_nuc_out32;

int nb200_occ_download32(nb200_ctx *ctx, nb200_dbatch *b, const nb200_occ_out32 *out);
int nb200_nuc_download32(nb200_ctx *ctx, nb200_dbatch *b, const nb200_nuc_out32 *out);
int64_t nb200_occ_d2h_bytes32(nb200_dbatch *b, const nb200_occ_out32 *out);
int64_t nb200_nuc_d2h_bytes32(nb200_dbatch *b, const nb200_nuc_out32 *out);

/* ---- measurement ------------------------------------------------------------------------ */
/* CUDA-event bracket on the batch stream. */
int nb200_timer_start(nb200_ctx *ctx, nb200_dbatch *b);
int nb200_timer_stop(nb200_ctx *ctx, nb200_dbatch *b);
int nb200_timer_elapsed_ms(nb200_ctx *ctx, nb200_dbatch *b, float *ms); /* syncs on the stop event */
/* Per-kernel accounting: launches and (when enabled) summed CUDA-event time of each named kernel. */
int nb200_profile_enable(nb200_ctx *ctx, int on);
int nb200_profile_reset(nb200_ctx *ctx);
int nb200_profile_count(nb200_ctx *ctx);
int nb200_profile_get(nb200_ctx *ctx, int i, const char **name, int64_t *launches, double *ms);
/* write `bytes` of device memory to evict L2 between timed iterations */
int nb200_flush_l2(nb200_ctx *ctx, nb200_dbatch *b);

/* ---- pyatac tools either side of the scoring path --------------------------- */
/* _vplotHelper summed over sites, pyatac/make_vplot.py:22-43 (+ ChunkMat2D.get(flip=...), chunkmat2d.py:21-54):
 * site s is centred at centers[s] (Chunk.center, chunk.py:41-54), flips[s] != 0 for strand "-", its reads (fetched for
 * [centre - flank - 1 - upper, centre + 1 + flank + upper)) are pos/tlen[frag_off[s] .. frag_off[s+1]).
 * out[upper-lower][2*flank+1] = sum over sites of the site matrix, each divided by its own sum when scale != 0
 * (a site without fragments then makes every cell NaN, as 0/0 does in the reference).  Unscaled sums are exact. */
int nb200_vplot(nb200_ctx *ctx, int32_t n_sites, const int32_t *centers, const int32_t *flips, const int64_t *frag_off,
                const int32_t *pos, const int32_t *tlen, int32_t flank, int32_t lower, int32_t upper, int32_t atac,
                int32_t scale, double *out);
/* _covHelper without the final scaling, pyatac/get_cov.py:22-31 + CoverageTrack.calculateCoverage, tracks.py:209-222:
 * out[x] = number of fragments with size in [lower, upper) whose centre lies within window_len/2 of start + x
 * (an even window is made one longer, pyatac/utils.py:34-36); out has end-start values.  Exact. */
int nb200_coverage(nb200_ctx *ctx, const int32_t *pos, const int32_t *tlen, int64_t n, int32_t start, int32_t end,
                   int32_t lower, int32_t upper, int32_t window_len, int32_t atac, double *out);

/* ---- host-side output formatting --------------------------------------------------------- */
/* Track.write_track, pyatac/tracks.py:37-74: run-length bedgraph rows "chrom\tstart\tend\tvalue\n" with the
 * reference's number format (Python-2 str(float)).  Pure host code, no context.  Returns the bytes needed; the
 * text is written when it fits into cap. */
int64_t nb200_format_track(const char *chrom, int64_t start, const double *vals, int64_t n, int32_t write_zero, char *out,
                           int64_t cap);

/* pysam.tabix_compress + pysam.tabix_index(preset="bed") of nucleoatac/run_occ.py:130-136 / run_nuc.py:194-201 in one
 * pass: plain sorted BED / bedgraph -> BGZF file + .tbi.  Pure host code (zlib, `threads` deflate workers). */
int nb200_bgzip_tabix(const char *path_plain, const char *path_gz, int threads, char *err, int errcap);
/* The same with a deflate level: -1 = zlib's default (what htslib / pysam.tabix_compress write with: byte-identical files),
 * 1..9 as zlib.  The .tbi does not depend on the level. */
int nb200_bgzip_tabix_level(const char *path_plain, const char *path_gz, int threads, int level, char *err, int errcap);

/* BedGraphFile.read, pyatac/bedgraph.py:6-16 (pysam.Tabixfile.fetch over a bgzip'd bedgraph with a .tbi): out[0 .. end-start) =
 * `empty`, then the value of every row overlapping [start, end) over the positions it covers.  Pure host code, no context. */
int nb200_bedgraph_fetch(const char *path_gz, const char *chrom, int64_t start, int64_t end, double empty, double *out, char *err,
                         int errcap);

/* ---- host-side BAM decode ------------------------------------------------------------------ */
/* The reads the path consumes -- pysam AlignmentFile.fetch + `is_proper_pair and not is_reverse` of
 * pyatac/fragments.pyx:21-25,47-50,128-131 -- for n_regions regions at once, decoded by `threads` host threads (zlib).
 * voffset[r] = BGZF virtual offset to start scanning region r from (the .bai linear index entry of its first 16 kb
 * window), tid[r] its reference id, [start[r], end[r]) the region.  frag_off[n_regions+1] receives the CSR offsets,
 * *pos / *tlen library-allocated int32 arrays of frag_off[n_regions] values (release with nb200_free).  Reads are kept when
 * pos < end and pos + max(l_seq,1) + 64 > start: a superset of htslib's overlap test; every consumer re-checks its own
 * cell bounds like fragments.pyx:37 does.  Pure host code, no context. */
int nb200_bam_fetch_many(const char *path, int32_t n_regions, const uint64_t *voffset, const int32_t *tid,
                         const int32_t *start, const int32_t *end, int32_t threads, int64_t *frag_off, int32_t **pos,
                         int32_t **tlen, char *err, int errcap);
void nb200_free(void *p);

/* ---- developer / test aid: the block plan of the tcgen05 background kernel, computed on the host --------------- */
/* What nb200_set_vmat + nb200_set_fragment_sizes would plan for BiasTrack.calculateBackgroundSignal
 * (nucleoatac/NucleosomeCalling.py:60-63) with this VMat (rows = insert sizes [lower, upper), `cols` columns) and
 * fragment-size distribution (n_sizes >= upper values): stats[16] = {eligible for the tensor-memory kernel, slabs, K16
 * blocks, tensor-memory operand columns of part 1, columns in all, slab and position of the first block that reads part 2,
 * bytes of one CTA's image of G, NA, NB, MMA columns per x-tile and precision pass, non-zero (row, K block) pairs of G,
 * non-zeros of G outside every block (must be 0), size-1 term present, 0, 0}; blocks[4 i ..] = the table entry of block i
 * ({K block | flags << 16, first row, instruction descriptor, image offset / 16 | rows per CTA << 16}; at most max_blocks
 * entries are copied, blocks may be NULL).  No context, no device: the CPU tests check the plan's invariants over shapes. */
int nb200_tc_plan_describe(const double *vmat, int32_t lower, int32_t upper, int32_t cols, const double *sizes, int32_t n_sizes,
                           int32_t *stats, int32_t *blocks, int32_t max_blocks);

/* ---- multi-GPU end-of-run reductions (NCCL over NVLink) ----------------------------------- */
/* fragment-size histogram (fragments.pyx:122-145), nuc_dist (run_occ.py:117-121), V-plot sum
 * (pyatac/make_vplot.py:70-73).  unique_id is the 128-byte ncclUniqueId from rank 0. */
int nb200_nccl_unique_id(void *out128);
int nb200_nccl_init(nb200_ctx *ctx, const void *unique_id128, int rank, int world);
int nb200_allreduce_f64(nb200_ctx *ctx, double *host_inout, int64_t n);
int nb200_allreduce_i64(nb200_ctx *ctx, int64_t *host_inout, int64_t n);
int nb200_nccl_finalize(nb200_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* NUCLEO_B200_H */
