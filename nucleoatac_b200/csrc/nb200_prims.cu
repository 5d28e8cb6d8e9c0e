// nb200_prims.cu -- single-call primitives behind the reference's Cython / numpy seams:
// host buffers in, host buffers out, dense float64 like the reference's own arrays.  These back the
// Python mirror classes (FragmentMat2D.makeFragmentMat, BiasMat2D.makeBiasMat, SignalTrack.calculateSignal,
// CoverageTrack.calculateCoverage, smooth, call_peaks, calculateOccupancy, calculateCov ...) and the
// reference's known-answer tests; the batched occ/nuc paths never materialise these dense arrays.
#include "nb200_dev.cuh"

static int h2d(nb200_ctx *ctx, DevBuf &d, const void *src, size_t bytes)
{
    NB_CUDA(ctx, d.reserve(bytes ? bytes : 1));
    if (bytes) NB_CUDA(ctx, cudaMemcpyAsync(d.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return NB200_OK;
}
static int d2h_sync(nb200_ctx *ctx, void *dst, const DevBuf &d, size_t bytes)
{
    if (bytes) NB_CUDA(ctx, cudaMemcpyAsync(dst, d.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    NB_CUDA(ctx, cudaGetLastError());
    return NB200_OK;
}
static inline unsigned blocks_for(int64_t n, int t) { return (unsigned)(n > 0 ? (n + t - 1) / t : 1); }

// ---- makeFragmentMat, pyatac/fragments.pyx:17-40 -------------------------------------------------
__global__ void k_fragmat_dense(const int32_t *__restrict__ pos, const int32_t *__restrict__ tlen, int64_t n, int start,
                                int ncol, int lower, int nrow, int atac, double *__restrict__ mat)
{
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    int l, i;
    frag_geometry(pos[f], tlen[f], atac, l, i);
    int row = i - lower, col = floordiv2(i - 1) + l - start;
    if (col >= 0 && col < ncol && row < nrow && row >= 0) atomicAdd(&mat[(size_t)row * ncol + col], 1.0);
}

// ---- getInsertions, pyatac/fragments.pyx:43-67 ---------------------------------------------------
__global__ void k_insertions(const int32_t *__restrict__ pos, const int32_t *__restrict__ tlen, int64_t n, int start, int end,
                             int lower, int upper, int atac, double *__restrict__ out)
{
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    int l, i;
    frag_geometry(pos[f], tlen[f], atac, l, i);
    if (i < lower || i >= upper) return;
    int r = l + i - 1;
    if (l >= start && l < end) atomicAdd(&out[l - start], 1.0);
    if (r >= start && r < end) atomicAdd(&out[r - start], 1.0);
}

// ---- getFragmentSizesFromChunkList, pyatac/fragments.pyx:122-145 ---------------------------------
__global__ void k_fragment_sizes(const int32_t *__restrict__ starts, const int32_t *__restrict__ ends,
                                 const int64_t *__restrict__ frag_off, const int32_t *__restrict__ pos,
                                 const int32_t *__restrict__ tlen, int lower, int upper, int atac,
                                 unsigned long long *__restrict__ counts)
{
    extern __shared__ unsigned int s_cnt[];
    const int c = blockIdx.x, nb = upper - lower;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    for (int64_t f = frag_off[c] + threadIdx.x; f < frag_off[c + 1]; f += blockDim.x) {
        int l, i;
        frag_geometry(pos[f], tlen[f], atac, l, i);
        int center = l + floordiv2(i - 1);
        if (i < upper && i >= lower && center >= starts[c] && center < ends[c]) atomicAdd(&s_cnt[i - lower], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += blockDim.x)
        if (s_cnt[i]) atomicAdd(&counts[i], (unsigned long long)s_cnt[i]);
}

// ---- BiasMat2D.makeBiasMat, pyatac/chunkmat2d.py:140-153 (literal exp(b_a + b_b)) ----------------
__global__ void k_biasmat_dense(const double *__restrict__ b, int lower, int upper, int ncol, double *__restrict__ mat)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = lower + blockIdx.y;
    if (n >= ncol) return;
    const int off = upper / 2;
    const int ia = off + n - floordiv2(i - 1), ib = off + n + (i >> 1);
    mat[(size_t)blockIdx.y * ncol + n] = (ia == ib) ? exp(b[ia]) : exp(b[ib] + b[ia]);
}

// ---- ChunkMat2D.getIns, pyatac/chunkmat2d.py:74-84 ------------------------------------------------
__global__ void k_get_ins(const double *__restrict__ mat, int lower, int upper, int64_t ncol, int64_t nout, double *__restrict__ out)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nout) return;
    const int mid = upper / 2;
    double s = 0.0;
    for (int i = lower; i < upper; i++) {
        const int t1 = mid + floordiv2(i - 1), t2 = mid - (i >> 1);
        const double *row = mat + (size_t)(i - lower) * ncol;
        s += row[t1 + n];
        if (t2 != t1) s += row[t2 + n];
    }
    out[n] = s;
}

// ---- scipy.signal.correlate(mat, vmat, 'valid')[0], NucleosomeCalling.py:34-36 (dense fp64) -------
__global__ void k_xcor_dense(const double *__restrict__ mat, int64_t ncol, const double *__restrict__ V, int R, int W,
                             int64_t nout, double *__restrict__ out)
{
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nout) return;
    double s = 0.0;
    for (int r = 0; r < R; r++) {
        const double *m = mat + (size_t)r * ncol + x;
        const double *v = V + (size_t)r * W;
        for (int k = 0; k < W; k++) s = fma(m[k], v[k], s);
    }
    out[x] = s;
}

// ---- CoverageTrack.calculateCoverage inner part, pyatac/tracks.py:216-222 ---------------------------
__global__ void k_coverage_dense(const double *__restrict__ mat, int64_t ncol, int row0, int row1, int window, int64_t nout,
                                 double *__restrict__ out)
{
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nout) return;
    double s = 0.0;
    for (int k = 0; k < window; k++) {  // flat-window sum of the column sums, left to right like np.convolve
        double col = 0.0;
        for (int r = row0; r < row1; r++) col += mat[(size_t)r * ncol + x + k];
        s += col;
    }
    out[x] = s;
}

// ---- smooth(), pyatac/utils.py:23-52, any window / mode ---------------------------------------------
__global__ void k_smooth_generic(const double *__restrict__ sig, int64_t n, const double *__restrict__ w, int wlen, int same,
                                 int norm, int64_t nout, double *__restrict__ out)
{
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= nout) return;
    // full convolution index: same -> o + (wlen-1)/2 ; valid -> o + wlen - 1
    const int64_t fidx = same ? o + (wlen - 1) / 2 : o + wlen - 1;
    double num = 0.0, den = 0.0;
    for (int m = 0; m < wlen; m++) {
        const int64_t idx = fidx - m;
        if (idx < 0 || idx >= n) continue;
        const double v = sig[idx];
        if (v == v) {
            num += w[m] * v;
            den += w[m];
        }
    }
    out[o] = norm ? (den == 0.0 ? nb_nan() : num / den) : num;
}

// ---- call_peaks / reduce_peaks, pyatac/utils.py:56-102 ----------------------------------------------
__global__ void __launch_bounds__(512) k_call_peaks(double *sig, int n, const double *__restrict__ jitter, double min_signal,
                                                    int sep, int boundary, int order, int *cpos, double *cval,
                                                    unsigned char *cst, int *out_idx, int cap, int *out_n)
{
    __shared__ double red_d[32];
    __shared__ int red_i[32];
    __shared__ int s_base, s_flag;
    const int tid = threadIdx.x;
    double mn = CUDART_INF;
    int nnan = 0;
    for (int x = tid; x < n; x += blockDim.x) {
        double v = sig[x];
        if (v != v) nnan++;
        else mn = fmin(mn, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(NB_FULL, mn, o));
        nnan += __shfl_xor_sync(NB_FULL, nnan, o);
    }
    if ((tid & 31) == 0) {
        red_d[tid >> 5] = mn;
        red_i[tid >> 5] = nnan;
    }
    __syncthreads();
    mn = CUDART_INF;
    nnan = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
        mn = fmin(mn, red_d[w]);
        nnan += red_i[w];
    }
    __syncthreads();
    if (tid == 0) s_base = 0;
    __syncthreads();
    if (nnan == n) {
        if (tid == 0) *out_n = 0;
        return;
    }
    if (nnan > 0)
        for (int x = tid; x < n; x += blockDim.x)
            if (sig[x] != sig[x]) sig[x] = mn;
    __syncthreads();
    const int lo = max(0, boundary), hi = n - boundary;
    for (int x0 = 0; x0 < n; x0 += blockDim.x) {
        const int x = x0 + tid;
        int flag = 0;
        double v = 0.0;
        if (x >= lo && x < hi) {
            v = sig[x];
            const double j0 = v * (1.0 + jitter[x]);
            flag = (v >= min_signal);
            for (int d = 1; d <= order && flag; d++) {
                const int xl = max(x - d, 0), xr = min(x + d, n - 1);
                flag = (j0 > sig[xl] * (1.0 + jitter[xl])) && (j0 > sig[xr] * (1.0 + jitter[xr]));
            }
        }
        int slot = block_compact_slot(flag, &s_base, red_i);
        if (flag) {
            cpos[slot] = x;
            cval[slot] = v;
        }
    }
    __syncthreads();
    const int m = s_base;
    block_nms(cpos, cval, cst, m, sep, &s_flag);
    __syncthreads();
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int j0 = 0; j0 < m; j0 += blockDim.x) {
        const int j = j0 + tid;
        const int flag = (j < m && cst[j] == 1);
        int slot = block_compact_slot(flag, &s_base, red_i);
        if (flag && slot < cap) out_idx[slot] = cpos[j];
    }
    __syncthreads();
    if (tid == 0) *out_n = s_base;
}

__global__ void __launch_bounds__(512) k_reduce_peaks(const int *pos, const double *val, unsigned char *st, int n, int sep, int *keep)
{
    __shared__ int s_flag;
    block_nms(pos, val, st, n, sep, &s_flag);
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) keep[j] = (st[j] == 1);
}

// ---- calculateOccupancy, nucleoatac/Occupancy.py:104-120 (dense, literal NaN semantics) -------------
__global__ void __launch_bounds__(NB200_MAX_ALPHA) k_calc_occupancy(const double *__restrict__ ins, const double *__restrict__ bias,
                                                                    int n, const double *__restrict__ pn,
                                                                    const double *__restrict__ pf,
                                                                    const double *__restrict__ alphas, int n_alpha, double cutoff,
                                                                    double *__restrict__ out3)
{
    __shared__ double red[32];
    __shared__ double s_ll[NB200_MAX_ALPHA];
    const int tid = threadIdx.x;
    double sn = 0.0, sf = 0.0;
    for (int i = tid; i < n; i += blockDim.x) {
        sn += __dmul_rn(pn[i], bias[i]);
        sf += __dmul_rn(pf[i], bias[i]);
    }
    const double SN = block_sum(sn, red), SF = block_sum(sf, red);
    double ll = nb_ninf();
    if (tid < n_alpha) {
        const double al = alphas[tid], om = 1.0 - al;
        double acc = 0.0;
        for (int i = 0; i < n; i++) {
            double nu = __dmul_rn(pn[i], bias[i]) / SN, nf = __dmul_rn(pf[i], bias[i]) / SF;
            double v = __dadd_rn(__dmul_rn(al, nu), __dmul_rn(om, nf));
            acc = __dadd_rn(acc, __dmul_rn(log(v), ins[i]));  // 0 * -inf = NaN, as in numpy
        }
        ll = (acc != acc) ? nb_ninf() : acc;
        s_ll[tid] = ll;
    }
    __syncthreads();
    if (tid == 0) {
        double best = nb_ninf();
        int bi = 0;
        for (int a = 0; a < n_alpha; a++)
            if (s_ll[a] > best) {
                best = s_ll[a];
                bi = a;
            }
        int lo = -1, hi = -1;
        for (int a = 0; a < n_alpha; a++)
            if (2.0 * (best - s_ll[a]) < cutoff) {
                if (lo < 0) lo = a;
                hi = a;
            }
        out3[0] = alphas[bi];
        out3[1] = lo >= 0 ? alphas[lo] : nb_nan();
        out3[2] = hi >= 0 ? alphas[hi] : nb_nan();
    }
}

// ---- calculateCov, nucleoatac/multinomial_cov.pyx:20-31 in closed form ------------------------------
__global__ void __launch_bounds__(1024) k_multinomial_cov(const double *__restrict__ p, const double *__restrict__ v, int64_t n, int r,
                                                          double *__restrict__ out)
{
    __shared__ double red[32];
    double s1 = 0.0, s2 = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double pv = p[i] * v[i];
        s1 += pv;
        s2 = fma(pv, v[i], s2);
    }
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) out[0] = (s2 - s1 * s1) * (double)r;
}

extern "C" {

int nb200_fragmat_build(nb200_ctx *ctx, const int32_t *pos, const int32_t *tlen, int64_t n, int32_t start, int32_t end,
                        int32_t lower, int32_t upper, int32_t atac, double *out)
{
    if (!ctx || !out || n < 0 || (n > 0 && (!pos || !tlen)) || end <= start || upper <= lower)
        return nb200_fail(ctx, NB200_ERR_ARG, "nb200_fragmat_build: bad argument");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int ncol = end - start, nrow = upper - lower;
    const size_t bytes = sizeof(double) * (size_t)ncol * nrow;
    NB_CHECK(h2d(ctx, ctx->s0, pos, sizeof(int32_t) * n));
    NB_CHECK(h2d(ctx, ctx->s1, tlen, sizeof(int32_t) * n));
    NB_CUDA(ctx, ctx->s2.reserve(bytes));
    NB_CUDA(ctx, cudaMemsetAsync(ctx->s2.p, 0, bytes, ctx->stream));
    if (n > 0) {
        ProfScope ps(ctx, ctx->stream, "k_fragmat_dense");
        k_fragmat_dense<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(ctx->s0.as<int32_t>(), ctx->s1.as<int32_t>(), n, start, ncol,
                                                                     lower, nrow, atac, ctx->s2.as<double>());
        NB_LAUNCH_CHECK(ctx);
    }
    return d2h_sync(ctx, out, ctx->s2, bytes);
}

int nb200_insertions(nb200_ctx *ctx, const int32_t *pos, const int32_t *tlen, int64_t n, int32_t start, int32_t end,
                     int32_t lower, int32_t upper, int32_t atac, double *out)
{
    if (!ctx || !out || n < 0 || (n > 0 && (!pos || !tlen)) || end <= start)
        return nb200_fail(ctx, NB200_ERR_ARG, "nb200_insertions: bad argument");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = sizeof(double) * (size_t)(end - start);
    NB_CHECK(h2d(ctx, ctx->s0, pos, sizeof(int32_t) * n));
    NB_CHECK(h2d(ctx, ctx->s1, tlen, sizeof(int32_t) * n));
    NB_CUDA(ctx, ctx->s2.reserve(bytes));
    NB_CUDA(ctx, cudaMemsetAsync(ctx->s2.p, 0, bytes, ctx->stream));
    if (n > 0) {
        ProfScope ps(ctx, ctx->stream, "k_insertions");
        k_insertions<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(ctx->s0.as<int32_t>(), ctx->s1.as<int32_t>(), n, start, end, lower,
                                                                  upper, atac, ctx->s2.as<double>());
        NB_LAUNCH_CHECK(ctx);
    }
    return d2h_sync(ctx, out, ctx->s2, bytes);
}

int nb200_fragment_sizes(nb200_ctx *ctx, int32_t n_chunks, const int32_t *starts, const int32_t *ends, const int64_t *frag_off,
                         const int32_t *pos, const int32_t *tlen, int32_t lower, int32_t upper, int32_t atac, int64_t *counts)
{
    if (!ctx || !counts || n_chunks < 1 || !starts || !ends || !frag_off || upper <= lower || upper - lower > 8192)
        return nb200_fail(ctx, NB200_ERR_ARG, "nb200_fragment_sizes: bad argument");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t n = frag_off[n_chunks];
    const size_t bytes = sizeof(int64_t) * (size_t)(upper - lower);
    NB_CHECK(h2d(ctx, ctx->s0, pos, sizeof(int32_t) * n));
    NB_CHECK(h2d(ctx, ctx->s1, tlen, sizeof(int32_t) * n));
    NB_CHECK(h2d(ctx, ctx->s2, starts, sizeof(int32_t) * n_chunks));
    NB_CHECK(h2d(ctx, ctx->s3, ends, sizeof(int32_t) * n_chunks));
    NB_CHECK(h2d(ctx, ctx->s4, frag_off, sizeof(int64_t) * (n_chunks + 1)));
    NB_CUDA(ctx, ctx->flush.reserve(bytes));
    NB_CUDA(ctx, cudaMemsetAsync(ctx->flush.p, 0, bytes, ctx->stream));
    {
        ProfScope ps(ctx, ctx->stream, "k_fragment_sizes");
        k_fragment_sizes<<<n_chunks, 256, sizeof(unsigned) * (upper - lower), ctx->stream>>>(
            ctx->s2.as<int32_t>(), ctx->s3.as<int32_t>(), ctx->s4.as<int64_t>(), ctx->s0.as<int32_t>(), ctx->s1.as<int32_t>(), lower,
            upper, atac, ctx->flush.as<unsigned long long>());
        NB_LAUNCH_CHECK(ctx);
    }
    return d2h_sync(ctx, counts, ctx->flush, bytes);
}

int nb200_bias_track(nb200_ctx *ctx, const uint8_t *seq, int64_t len, double *out)
{
    if (!ctx || !seq || !out) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_bias_track: NULL argument");
    RunConst &r = ctx->rc;
    if (!r.have_pwm) return nb200_fail(ctx, NB200_ERR_STATE, "nb200_set_pwm has not been called");
    const int64_t blen = len - (r.pwm_width - 1);
    if (blen < 1) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_bias_track: sequence shorter than the PWM");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    int64_t off[2] = {0, len}, boff[2] = {0, blen};
    NB_CHECK(h2d(ctx, ctx->s0, seq, (size_t)len));
    NB_CHECK(h2d(ctx, ctx->s1, off, sizeof(off)));
    NB_CHECK(h2d(ctx, ctx->s2, boff, sizeof(boff)));
    NB_CUDA(ctx, ctx->s3.reserve(sizeof(double) * blen));
    {
        ProfScope ps(ctx, ctx->stream, "k_bias_track");
        dim3 grid((unsigned)div_up64(blen, BT_TILE), 1);
        k_bias_track<<<grid, 256, 0, ctx->stream>>>(ctx->s0.as<uint8_t>(), ctx->s1.as<int64_t>(), ctx->s2.as<int64_t>(),
                                                    r.log_pwm.as<double>(), r.nuc_code.as<int8_t>(), r.n_nuc, r.pwm_width, nullptr,
                                                    ctx->s3.as<double>());
        NB_LAUNCH_CHECK(ctx);
    }
    return d2h_sync(ctx, out, ctx->s3, sizeof(double) * blen);
}

int nb200_biasmat_build(nb200_ctx *ctx, const double *bias, int64_t n_bias, int32_t lower, int32_t upper, double *out)
{
    if (!ctx || !bias || !out || upper <= lower || lower < 0) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_biasmat_build: bad argument");
    const int plen = upper + (upper - 1) % 2;
    const int64_t ncol = n_bias - plen + 1;
    if (ncol < 1) return nb200_fail(ctx, NB200_ERR_FLANK, "nb200_biasmat_build: bias track shorter than the pattern width");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = sizeof(double) * (size_t)ncol * (upper - lower);
    NB_CHECK(h2d(ctx, ctx->s0, bias, sizeof(double) * n_bias));
    NB_CUDA(ctx, ctx->s2.reserve(bytes));
    {
        ProfScope ps(ctx, ctx->stream, "k_biasmat_dense");
        dim3 grid(blocks_for(ncol, 256), upper - lower);
        k_biasmat_dense<<<grid, 256, 0, ctx->stream>>>(ctx->s0.as<double>(), lower, upper, (int)ncol, ctx->s2.as<double>());
        NB_LAUNCH_CHECK(ctx);
    }
    return d2h_sync(ctx, out, ctx->s2, bytes);
}

int nb200_get_ins(nb200_ctx *ctx, const double *mat, int32_t lower, int32_t upper, int64_t ncol, double *out)
{
    if (!ctx || !mat || !out || upper <= lower || lower < 0) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_get_ins: bad argument");
    const int plen = upper + (upper - 1) % 2;
    const int64_t nout = ncol - plen + 1;
    if (nout < 1) return nb200_fail(ctx, NB200_ERR_FLANK, "nb200_get_ins: matrix narrower than the pattern");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CHECK(h2d(ctx, ctx->s2, mat, sizeof(double) * (size_t)ncol * (upper - lower)));
    NB_CUDA(ctx, ctx->s3.reserve(sizeof(double) * nout));
    {
        ProfScope ps(ctx, ctx->stream, "k_get_ins");
        k_get_ins<<<blocks_for(nout, 256), 256, 0, ctx->stream>>>(ctx->s2.as<double>(), lower, upper, ncol, nout, ctx->s3.as<double>());
        NB_LAUNCH_CHECK(ctx);
    }
    return d2h_sync(ctx, out, ctx->s3, sizeof(double) * nout);
}

int nb200_xcor_dense(nb200_ctx *ctx, const double *mat, int64_t ncol, double *out)
{
    if (!ctx || !mat || !out) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_xcor_dense: NULL argument");
    RunConst &r = ctx->rc;
    if (!r.have_vmat) return nb200_fail(ctx, NB200_ERR_STATE, "nb200_set_vmat has not been called");
    const int64_t nout = ncol - r.v_cols + 1;
    if (nout < 1) return nb200_fail(ctx, NB200_ERR_FLANK, "Insufficient flanking region on mat to calculate signal");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CHECK(h2d(ctx, ctx->s2, mat, sizeof(double) * (size_t)ncol * r.v_rows));
    NB_CUDA(ctx, ctx->s3.reserve(sizeof(double) * nout));
    {
        ProfScope ps(ctx, ctx->stream, "k_xcor_dense");
        k_xcor_dense<<<blocks_for(nout, 128), 128, 0, ctx->stream>>>(ctx->s2.as<double>(), ncol, r.vmat.as<double>(), r.v_rows, r.v_cols,
                                                                     nout, ctx->s3.as<double>());
        NB_LAUNCH_CHECK(ctx);
    }
    return d2h_sync(ctx, out, ctx->s3, sizeof(double) * nout);
}

int nb200_coverage_dense(nb200_ctx *ctx, const double *mat, int32_t nrow, int64_t ncol, int32_t row0, int32_t row1, int32_t window_len,
                         double *out)
{
    if (!ctx || !mat || !out || row0 < 0 || row1 > nrow || window_len < 1)
        return nb200_fail(ctx, NB200_ERR_ARG, "nb200_coverage_dense: bad argument");
    const int64_t nout = ncol - window_len + 1;
    if (nout < 1) return nb200_fail(ctx, NB200_ERR_FLANK, "Insufficient flanking region on mat to calculate coverage with desired window");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CHECK(h2d(ctx, ctx->s2, mat, sizeof(double) * (size_t)ncol * nrow));
    NB_CUDA(ctx, ctx->s3.reserve(sizeof(double) * nout));
    {
        ProfScope ps(ctx, ctx->stream, "k_coverage_dense");
        k_coverage_dense<<<blocks_for(nout, 128), 128, 0, ctx->stream>>>(ctx->s2.as<double>(), ncol, row0, row1, window_len, nout,
                                                                         ctx->s3.as<double>());
        NB_LAUNCH_CHECK(ctx);
    }
    return d2h_sync(ctx, out, ctx->s3, sizeof(double) * nout);
}

int nb200_smooth(nb200_ctx *ctx, const double *sig, int64_t n, const double *w, int32_t wlen, int32_t mode_same, int32_t norm, double *out)
{
    if (!ctx || !sig || !w || !out || n < 1 || wlen < 1) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_smooth: bad argument");
    const int64_t nout = mode_same ? (n > wlen ? n : wlen) : (n >= wlen ? n - wlen + 1 : wlen - n + 1);
    if (mode_same && n < wlen) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_smooth: 'same' needs len(sig) >= len(window)");
    if (!mode_same && n < wlen) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_smooth: 'valid' needs len(sig) >= len(window)");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CHECK(h2d(ctx, ctx->s0, sig, sizeof(double) * n));
    NB_CHECK(h2d(ctx, ctx->s1, w, sizeof(double) * wlen));
    NB_CUDA(ctx, ctx->s3.reserve(sizeof(double) * nout));
    {
        ProfScope ps(ctx, ctx->stream, "k_smooth_generic");
        k_smooth_generic<<<blocks_for(nout, 256), 256, 0, ctx->stream>>>(ctx->s0.as<double>(), n, ctx->s1.as<double>(), wlen, mode_same,
                                                                         norm, nout, ctx->s3.as<double>());
        NB_LAUNCH_CHECK(ctx);
    }
    return d2h_sync(ctx, out, ctx->s3, sizeof(double) * nout);
}

int nb200_call_peaks(nb200_ctx *ctx, double *sig, int64_t n, double min_signal, int32_t sep, int32_t boundary, int32_t order,
                     int32_t *out_idx, int32_t cap, int32_t *out_n)
{
    if (!ctx || !sig || !out_idx || !out_n || n < 1 || n > (1 << 28) || sep < 1 || order < 1 || cap < 0)
        return nb200_fail(ctx, NB200_ERR_ARG, "nb200_call_peaks: bad argument");
    RunConst &r = ctx->rc;
    if (r.n_jitter < n) return nb200_fail(ctx, NB200_ERR_STATE, "nb200_set_jitter: need >= %lld values", (long long)n);
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CHECK(h2d(ctx, ctx->s0, sig, sizeof(double) * n));
    NB_CUDA(ctx, ctx->s1.reserve(sizeof(int32_t) * n));
    NB_CUDA(ctx, ctx->s2.reserve(sizeof(double) * n));
    NB_CUDA(ctx, ctx->s3.reserve((size_t)n));
    NB_CUDA(ctx, ctx->s4.reserve(sizeof(int32_t) * ((size_t)cap + 1)));
    {
        ProfScope ps(ctx, ctx->stream, "k_call_peaks");
        k_call_peaks<<<1, 512, 0, ctx->stream>>>(ctx->s0.as<double>(), (int)n, r.jitter.as<double>(), min_signal, sep, boundary, order,
                                                 ctx->s1.as<int32_t>(), ctx->s2.as<double>(), ctx->s3.as<unsigned char>(),
                                                 ctx->s4.as<int32_t>() + 1, cap, ctx->s4.as<int32_t>());
        NB_LAUNCH_CHECK(ctx);
    }
    NB_CUDA(ctx, cudaMemcpyAsync(sig, ctx->s0.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(ctx, cudaMemcpyAsync(out_n, ctx->s4.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*out_n > cap) return nb200_fail(ctx, NB200_ERR_CAPACITY, "nb200_call_peaks: %d peaks exceed the capacity %d", *out_n, cap);
    if (*out_n > 0) NB_CUDA(ctx, cudaMemcpy(out_idx, ctx->s4.as<int32_t>() + 1, sizeof(int32_t) * (*out_n), cudaMemcpyDeviceToHost));
    return NB200_OK;
}

int nb200_reduce_peaks(nb200_ctx *ctx, const int32_t *peaks, const double *sig, int32_t n, int32_t sep, int32_t *keep)
{
    if (!ctx || n < 0 || (n > 0 && (!peaks || !sig || !keep))) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_reduce_peaks: bad argument");
    if (n == 0) return NB200_OK;
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CHECK(h2d(ctx, ctx->s0, peaks, sizeof(int32_t) * n));
    NB_CHECK(h2d(ctx, ctx->s1, sig, sizeof(double) * n));
    NB_CUDA(ctx, ctx->s2.reserve((size_t)n));
    NB_CUDA(ctx, ctx->s3.reserve(sizeof(int32_t) * n));
    {
        ProfScope ps(ctx, ctx->stream, "k_reduce_peaks");
        k_reduce_peaks<<<1, 512, 0, ctx->stream>>>(ctx->s0.as<int32_t>(), ctx->s1.as<double>(), ctx->s2.as<unsigned char>(), n, sep,
                                                   ctx->s3.as<int32_t>());
        NB_LAUNCH_CHECK(ctx);
    }
    return d2h_sync(ctx, keep, ctx->s3, sizeof(int32_t) * n);
}

int nb200_calculate_occupancy(nb200_ctx *ctx, const double *inserts, const double *bias, int32_t n, double *out3)
{
    if (!ctx || !inserts || !bias || !out3) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_calculate_occupancy: NULL argument");
    RunConst &r = ctx->rc;
    if (!r.have_occ_model || r.occ_upper != n)
        return nb200_fail(ctx, NB200_ERR_STATE, "nb200_set_occ_model missing or its length differs from the inserts vector");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CHECK(h2d(ctx, ctx->s0, inserts, sizeof(double) * n));
    NB_CHECK(h2d(ctx, ctx->s1, bias, sizeof(double) * n));
    NB_CUDA(ctx, ctx->s3.reserve(sizeof(double) * 3));
    {
        ProfScope ps(ctx, ctx->stream, "k_calc_occupancy");
        k_calc_occupancy<<<1, NB200_MAX_ALPHA, 0, ctx->stream>>>(ctx->s0.as<double>(), ctx->s1.as<double>(), n, r.nuc_probs.as<double>(),
                                                                 r.nfr_probs.as<double>(), r.alphas.as<double>(), r.n_alpha, r.cutoff,
                                                                 ctx->s3.as<double>());
        NB_LAUNCH_CHECK(ctx);
    }
    return d2h_sync(ctx, out3, ctx->s3, sizeof(double) * 3);
}

int nb200_multinomial_cov(nb200_ctx *ctx, const double *p, const double *v, int64_t n, int32_t r, double *out)
{
    if (!ctx || !p || !v || !out || n < 1) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_multinomial_cov: bad argument");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CHECK(h2d(ctx, ctx->s0, p, sizeof(double) * n));
    NB_CHECK(h2d(ctx, ctx->s1, v, sizeof(double) * n));
    NB_CUDA(ctx, ctx->s3.reserve(sizeof(double)));
    {
        ProfScope ps(ctx, ctx->stream, "k_multinomial_cov");
        k_multinomial_cov<<<1, 1024, 0, ctx->stream>>>(ctx->s0.as<double>(), ctx->s1.as<double>(), n, r, ctx->s3.as<double>());
        NB_LAUNCH_CHECK(ctx);
    }
    return d2h_sync(ctx, out, ctx->s3, sizeof(double));
}

}  // extern "C"
