// nb200_batch.cu -- batch upload (one H2D of packed reads + sequence), the two preparation stages
// shared by the occ and nuc paths (Tn5 log-bias track; fragment matrix in compressed-column form)
// and the packed D2H of results.
//
// Data layout in HBM (DESIGN.md "layout"): the reference's dense insert-size x position float64
// matrices (FragmentMat2D / BiasMat2D, 21 MB each per 10 kb chunk) are never materialised.
//   * FragmentMat2D  -> CSC: per genomic column the sorted list of insert sizes (rows) of the
//                       fragments centred there + an exclusive prefix of per-column counts, which
//                       doubles as the coverage prefix sum (integer, exact).
//   * BiasMat2D      -> the 1-D track E[p] = exp(log-bias[p]); a cell is E[l]*E[r] (two taps).
#include "nb200_dev.cuh"

// ---------------------------------------------------------------------------------------------
// K1: makeFragmentMat, pyatac/fragments.pyx:17-40, into compressed-column form.  One block per
// chunk: count per column (integer atomics), exclusive scan, fill, per-column sort of the rows
// (=> deterministic layout).  row = ilen - 0, col = (ilen-1)//2 + l_pos - (start - pad).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool frag_cell(int pos, int tlen, int atac, int mat_start, int ncol, int upper, int &row, int &col)
{
    int l, i;
    frag_geometry(pos, tlen, atac, l, i);
    row = i;
    col = floordiv2(i - 1) + l - mat_start;
    return col >= 0 && col < ncol && row < upper && row >= 0;
}

// in-place exclusive scan over a[0..n) by one block; a[n] receives the total.  red: shared int[blockDim.x]
static __device__ void block_exclusive_scan(int *a, int n, int *red)
{
    const int T = blockDim.x, t = threadIdx.x;
    const int per = (n + T - 1) / T;
    const int lo = min(n, t * per), hi = min(n, lo + per);
    int s = 0;
    for (int i = lo; i < hi; i++) s += a[i];
    red[t] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over the T partial sums
    for (int o = 1; o < T; o <<= 1) {
        int v = (t >= o) ? red[t - o] : 0;
        __syncthreads();
        red[t] += v;
        __syncthreads();
    }
    int run = red[t] - s;
    for (int i = lo; i < hi; i++) {
        int v = a[i];
        a[i] = run;
        run += v;
    }
    if (t == T - 1) a[n] = red[T - 1];
    __syncthreads();
}

__global__ void __launch_bounds__(512) k_csc_build(const int32_t *__restrict__ start, const int32_t *__restrict__ end,
                                                   const int64_t *__restrict__ frag_off,
                                                   const int32_t *__restrict__ pos, const int32_t *__restrict__ tlen,
                                                   const int64_t *__restrict__ col_off, int pad, int upper, int atac,
                                                   int lower_split, int32_t *__restrict__ col_ptr,
                                                   int32_t *__restrict__ col_low, int32_t *__restrict__ cursor,
                                                   int2 *__restrict__ ent)
{
    __shared__ int red[512];
    const int c = blockIdx.x;
    const int mat_start = start[c] - pad;
    const int ncol = end[c] - start[c] + 2 * pad;
    const int64_t f0 = frag_off[c], f1 = frag_off[c + 1];
    int32_t *cp = col_ptr + col_off[c];
    int32_t *cl = col_low ? col_low + col_off[c] : nullptr;
    int32_t *cur = cursor + col_off[c];
    int2 *en = ent + f0;
    for (int i = threadIdx.x; i <= ncol; i += blockDim.x) {
        cp[i] = 0;
        if (cl) cl[i] = 0;
    }
    __syncthreads();
    for (int64_t f = f0 + threadIdx.x; f < f1; f += blockDim.x) {
        int row, col;
        if (frag_cell(pos[f], tlen[f], atac, mat_start, ncol, upper, row, col)) {
            atomicAdd(&cp[col], 1);
            if (cl && row < lower_split) atomicAdd(&cl[col], 1);
        }
    }
    __syncthreads();
    block_exclusive_scan(cp, ncol, red);
    if (cl) block_exclusive_scan(cl, ncol, red);
    for (int i = threadIdx.x; i < ncol; i += blockDim.x) cur[i] = cp[i];
    __syncthreads();
    for (int64_t f = f0 + threadIdx.x; f < f1; f += blockDim.x) {
        int row, col;
        if (frag_cell(pos[f], tlen[f], atac, mat_start, ncol, upper, row, col)) {
            int slot = atomicAdd(&cur[col], 1);
            en[slot] = make_int2(col, row);
        }
    }
    __syncthreads();
    // deterministic order inside a column: insertion sort by row (columns hold a handful of entries)
    for (int i = threadIdx.x; i < ncol; i += blockDim.x) {
        int a = cp[i], b = cp[i + 1];
        for (int j = a + 1; j < b; j++) {
            int2 key = en[j];
            int k = j - 1;
            while (k >= a && en[k].y > key.y) {
                en[k + 1] = en[k];
                k--;
            }
            en[k + 1] = key;
        }
    }
}

int nb200_prep_bias(nb200_ctx *ctx, nb200_dbatch *b)
{
    if (b->bias_done) return NB200_OK;
    if (!b->have_seq) return nb200_fail(ctx, NB200_ERR_STATE, "bias requested (use_bias=1) but the batch has no sequence");
    RunConst &r = ctx->rc;
    if (!r.have_pwm) return nb200_fail(ctx, NB200_ERR_STATE, "nb200_set_pwm has not been called");
    const int n = b->n_chunks;
    std::vector<int64_t> boff(n + 1, 0);
    int64_t max_b = 0;
    for (int c = 0; c < n; c++) {
        int64_t slen = b->h_seq_off[c + 1] - b->h_seq_off[c];
        int64_t bl = slen - (r.pwm_width - 1);
        if (bl < 1) return nb200_fail(ctx, NB200_ERR_ARG, "chunk %d: sequence shorter than the PWM", c);
        boff[c + 1] = boff[c] + bl;
        if (bl > max_b) max_b = bl;
    }
    b->n_bias = boff[n];
    NB_CUDA(ctx, b->d_bias_off.reserve(sizeof(int64_t) * (n + 1)));
    memcpy(b->pin_slot(1), boff.data(), sizeof(int64_t) * (n + 1));
    NB_CUDA(ctx, cudaMemcpyAsync(b->d_bias_off.p, b->pin_slot(1), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, b->stream));
    NB_CUDA(ctx, b->d_E.reserve(sizeof(double) * b->n_bias));
    {
        ProfScope ps(ctx, b->stream, "k_bias_track");
        dim3 grid((unsigned)div_up64(max_b, BT_TILE), n);
        k_bias_track<<<grid, 256, 0, b->stream>>>(b->d_seq.as<uint8_t>(), b->d_seq_off.as<int64_t>(),
                                                  b->d_bias_off.as<int64_t>(), r.log_pwm.as<double>(),
                                                  r.nuc_code.as<int8_t>(), r.n_nuc, r.pwm_width, b->d_E.as<double>(), nullptr);
    }
    NB_LAUNCH_CHECK(ctx);
    b->bias_done = true;
    return NB200_OK;
}

int nb200_prep_csc(nb200_ctx *ctx, nb200_dbatch *b, int pad, int upper, int atac, int lower_split)
{
    if (b->csc_pad >= pad && b->csc_upper == upper && b->csc_atac == atac &&
        (lower_split <= 0 || b->csc_lower_split == lower_split))
        return NB200_OK;
    if (b->csc_pad > pad) pad = b->csc_pad;
    if (lower_split <= 0 && b->csc_upper == upper && b->csc_atac == atac) lower_split = b->csc_lower_split;
    const int n = b->n_chunks;
    std::vector<int64_t> coff(n + 1, 0);
    for (int c = 0; c < n; c++) coff[c + 1] = coff[c] + (b->h_end[c] - b->h_start[c]) + 2 * (int64_t)pad + 1;
    b->n_colptr = coff[n];
    NB_CUDA(ctx, b->d_col_off.reserve(sizeof(int64_t) * (n + 1)));
    memcpy(b->pin_slot(2), coff.data(), sizeof(int64_t) * (n + 1));
    NB_CUDA(ctx, cudaMemcpyAsync(b->d_col_off.p, b->pin_slot(2), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, b->stream));
    NB_CUDA(ctx, b->d_col_ptr.reserve(sizeof(int32_t) * b->n_colptr));
    NB_CUDA(ctx, b->d_cursor.reserve(sizeof(int32_t) * b->n_colptr));
    if (lower_split > 0) NB_CUDA(ctx, b->d_col_low.reserve(sizeof(int32_t) * b->n_colptr));
    NB_CUDA(ctx, b->d_ent.reserve(sizeof(int2) * (size_t)(b->n_frag > 0 ? b->n_frag : 1)));
    {
        ProfScope ps(ctx, b->stream, "k_csc_build");
        k_csc_build<<<n, 512, 0, b->stream>>>(b->d_start.as<int32_t>(), b->d_end.as<int32_t>(), b->d_frag_off.as<int64_t>(),
                                              b->d_pos.as<int32_t>(), b->d_tlen.as<int32_t>(), b->d_col_off.as<int64_t>(),
                                              pad, upper, atac, lower_split, b->d_col_ptr.as<int32_t>(),
                                              lower_split > 0 ? b->d_col_low.as<int32_t>() : nullptr,
                                              b->d_cursor.as<int32_t>(), b->d_ent.as<int2>());
    }
    NB_LAUNCH_CHECK(ctx);
    b->csc_pad = pad;
    b->csc_upper = upper;
    b->csc_atac = atac;
    b->csc_lower_split = lower_split > 0 ? lower_split : -1;
    return NB200_OK;
}

extern "C" {

int nb200_batch_upload(nb200_ctx *ctx, const nb200_batch *h, nb200_dbatch **io)
{
    if (!ctx || !h || !io) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_batch_upload: NULL argument");
    if (h->n_chunks < 1 || !h->chunk_start || !h->chunk_end || !h->frag_off)
        return nb200_fail(ctx, NB200_ERR_ARG, "nb200_batch_upload: need n_chunks >= 1, chunk_start, chunk_end, frag_off");
    if (h->n_chunks > 65535)  // chunks ride on gridDim.y
        return nb200_fail(ctx, NB200_ERR_CAPACITY, "nb200_batch_upload: at most 65535 chunks per batch (got %d); split the chunk list", h->n_chunks);
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    nb200_dbatch *b = *io;
    if (!b) {
        b = new nb200_dbatch();
        cudaError_t e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_pass, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_copied_occ, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_copied_nuc, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            nb200_batch_free(nullptr, b);
            return nb200_cuda_fail(ctx, e, "cudaStreamCreate", __FILE__, __LINE__);
        }
        *io = b;
    } else {
        // the arrays of the previous use may still be on their way to the host
        NB_CUDA(ctx, cudaStreamWaitEvent(b->stream, b->ev_copied_occ, 0));
        NB_CUDA(ctx, cudaStreamWaitEvent(b->stream, b->ev_copied_nuc, 0));
    }
    const int n = h->n_chunks;
    if (b->h_pin_cap < (size_t)n + 1) {
        NB_CUDA(ctx, cudaStreamSynchronize(b->stream));
        if (b->h_pin) cudaFreeHost(b->h_pin);
        b->h_pin = nullptr;
        b->h_pin_cap = 0;
        NB_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void **>(&b->h_pin), sizeof(int64_t) * 5 * ((size_t)n + 1), cudaHostAllocDefault));
        b->h_pin_cap = (size_t)n + 1;
    }
    b->n_chunks = n;
    b->h_start.assign(h->chunk_start, h->chunk_start + n);
    b->h_end.assign(h->chunk_end, h->chunk_end + n);
    b->h_frag_off.assign(h->frag_off, h->frag_off + n + 1);
    b->h_out_off.assign(n + 1, 0);
    b->max_len = 0;
    b->min_len = INT32_MAX;
    for (int c = 0; c < n; c++) {
        int64_t len = (int64_t)b->h_end[c] - b->h_start[c];
        if (len < 1 || len > (1 << 28)) return nb200_fail(ctx, NB200_ERR_ARG, "chunk %d: bad length %lld", c, (long long)len);
        if (b->h_frag_off[c + 1] < b->h_frag_off[c]) return nb200_fail(ctx, NB200_ERR_ARG, "frag_off must be non-decreasing");
        b->h_out_off[c + 1] = b->h_out_off[c] + len;
        if (len > b->max_len) b->max_len = (int)len;
        if (len < b->min_len) b->min_len = (int)len;
    }
    if (b->h_frag_off[0] != 0) return nb200_fail(ctx, NB200_ERR_ARG, "frag_off[0] must be 0");
    b->total_len = b->h_out_off[n];
    b->n_frag = b->h_frag_off[n];
    if (b->n_frag > 0 && (!h->frag_pos || !h->frag_tlen)) return nb200_fail(ctx, NB200_ERR_ARG, "frag_pos/frag_tlen are NULL");
    b->have_seq = h->seq_off && h->seq_start && h->seq;
    b->n_seq = 0;
    if (b->have_seq) {
        b->h_seq_off.assign(h->seq_off, h->seq_off + n + 1);
        b->h_seq_start.assign(h->seq_start, h->seq_start + n);
        if (b->h_seq_off[0] != 0) return nb200_fail(ctx, NB200_ERR_ARG, "seq_off[0] must be 0");
        for (int c = 0; c < n; c++)
            if (b->h_seq_off[c + 1] < b->h_seq_off[c]) return nb200_fail(ctx, NB200_ERR_ARG, "seq_off must be non-decreasing");
        b->n_seq = b->h_seq_off[n];
    }
    b->bias_done = false;
    b->csc_pad = b->csc_upper = b->csc_atac = b->csc_lower_split = -1;
    b->occ_done = b->nuc_done = false;
    b->occ_cols_gen = -1;
    b->h2d_bytes = 0;
    auto up = [&](DevBuf &d, const void *src, size_t bytes) -> int {
        NB_CUDA(ctx, d.reserve(bytes ? bytes : 1));
        if (bytes) NB_CUDA(ctx, cudaMemcpyAsync(d.p, src, bytes, cudaMemcpyHostToDevice, b->stream));
        b->h2d_bytes += (int64_t)bytes;
        return NB200_OK;
    };
    NB_CHECK(up(b->d_start, h->chunk_start, sizeof(int32_t) * n));
    NB_CHECK(up(b->d_end, h->chunk_end, sizeof(int32_t) * n));
    NB_CHECK(up(b->d_frag_off, h->frag_off, sizeof(int64_t) * (n + 1)));
    NB_CHECK(up(b->d_pos, h->frag_pos, sizeof(int32_t) * b->n_frag));
    NB_CHECK(up(b->d_tlen, h->frag_tlen, sizeof(int32_t) * b->n_frag));
    memcpy(b->pin_slot(0), b->h_out_off.data(), sizeof(int64_t) * (n + 1));
    NB_CHECK(up(b->d_out_off, b->pin_slot(0), sizeof(int64_t) * (n + 1)));
    if (b->have_seq) {
        NB_CHECK(up(b->d_seq_off, h->seq_off, sizeof(int64_t) * (n + 1)));
        NB_CHECK(up(b->d_seq_start, h->seq_start, sizeof(int32_t) * n));
        NB_CHECK(up(b->d_seq, h->seq, (size_t)b->n_seq));
    }
    return NB200_OK;
}

int nb200_batch_free(nb200_ctx *ctx, nb200_dbatch *b)
{
    if (!b) return NB200_OK;
    if (ctx) cudaSetDevice(ctx->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    if (b->copy_stream) cudaStreamSynchronize(b->copy_stream);
    DevBuf *bufs[] = {&b->d_start, &b->d_end, &b->d_frag_off, &b->d_pos, &b->d_tlen, &b->d_seq_off, &b->d_seq_start,
                      &b->d_seq, &b->d_out_off, &b->d_bias_off, &b->d_E, &b->d_col_off, &b->d_col_ptr, &b->d_col_low,
                      &b->d_cursor, &b->d_ent, &b->o_vals, &b->o_lower, &b->o_upper, &b->o_svals, &b->o_slower,
                      &b->o_supper, &b->o_cov, &b->o_nuc_dist, &b->o_peak_count, &b->o_peak_pos, &b->o_peak_occ,
                      &b->o_peak_lower, &b->o_peak_upper, &b->o_peak_reads, &b->o_cn, &b->o_cf, &b->o_wsn, &b->o_wsf,
                      &b->o_peak_off, &b->n_signal, &b->n_bg, &b->n_norm, &b->n_smooth, &b->n_nuc_cov, &b->n_nfr_cov,
                      &b->n_bx, &b->n_bcov, &b->n_cB, &b->n_comb, &b->n_cand_bcov, &b->n_cand_count, &b->n_cand_pos, &b->n_cand_flag,
                      &b->n_cand_z, &b->n_cand_lr, &b->n_cand_norm, &b->n_cand_sig, &b->n_cand_cov, &b->n_cand_nfr,
                      &b->n_cand_smooth, &b->n_cand_off, &b->n_work, &b->n_work_count, &b->sc_i32, &b->sc_f64, &b->sc_u8,
                      &b->pack32_occ, &b->pack32_nuc, &b->o_wv};
    for (auto d : bufs) d->release();
    if (b->ev_start) cudaEventDestroy(b->ev_start);
    if (b->ev_stop) cudaEventDestroy(b->ev_stop);
    if (b->h_pin) cudaFreeHost(b->h_pin);
    if (b->ev_pass) cudaEventDestroy(b->ev_pass);
    if (b->ev_copied_occ) cudaEventDestroy(b->ev_copied_occ);
    if (b->ev_copied_nuc) cudaEventDestroy(b->ev_copied_nuc);
    if (b->copy_stream) cudaStreamDestroy(b->copy_stream);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
    return NB200_OK;
}

int nb200_batch_sync(nb200_ctx *ctx, nb200_dbatch *b)
{
    if (!ctx || !b) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_batch_sync: NULL argument");
    NB_CUDA(ctx, cudaStreamSynchronize(b->stream));
    NB_CUDA(ctx, cudaStreamSynchronize(b->copy_stream));
    NB_CUDA(ctx, cudaGetLastError());
    return NB200_OK;
}

int64_t nb200_batch_total_len(nb200_dbatch *b) { return b ? b->total_len : 0; }
int64_t nb200_batch_h2d_bytes(nb200_dbatch *b) { return b ? b->h2d_bytes : 0; }

}  // extern "C"
