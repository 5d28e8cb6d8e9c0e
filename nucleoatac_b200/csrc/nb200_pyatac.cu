// nb200_pyatac.cu -- the aggregate / per-region primitives behind the pyatac tools that sit either side of the
// scoring path: `pyatac vplot` (pyatac/make_vplot.py:22-43) and `pyatac cov` (pyatac/get_cov.py:22-38).
// Both reduce to integer scatter work on the packed (pos, tlen) reads; neither materialises the per-site dense
// FragmentMat2D the reference builds (501 x 250 float64 per site for a V-plot, 2000 x (L + 120) per coverage chunk).
#include "nb200_dev.cuh"

static int up(nb200_ctx *ctx, DevBuf &d, const void *src, size_t bytes)
{
    NB_CUDA(ctx, d.reserve(bytes ? bytes : 1));
    if (bytes) NB_CUDA(ctx, cudaMemcpyAsync(d.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return NB200_OK;
}

// Column of a fragment in the site's V-plot, or -1.  _vplotHelper (make_vplot.py:29-33): chunk.center(); the matrix spans
// [c - flank - 1, c + 1 + flank); get(start = c - flank, end = c + 1 + flank, flip = strand == "-").  Unflipped, column t is
// genomic c - flank + t.  Flipped (chunkmat2d.py:41-54): odd sizes are mirrored about c, even sizes (whose centre is the
// left one of the two middle bases) about c - 1/2:  t = c + flank - centre  (odd),  c + flank - 1 - centre  (even).
__device__ __forceinline__ int vplot_col(int centre, int size, int c, int flank, int flip)
{
    const int t = flip ? (c + flank - ((size & 1) ? 0 : 1) - centre) : (centre - (c - flank));
    return (t >= 0 && t <= 2 * flank) ? t : -1;
}

// One block per site.  Pass 1 (scale only): the site's in-plot fragment count; pass 2: add 1 (or 1 / count) per fragment.
// Unscaled sums are integers in float64 and therefore exact whatever the order of the atomics.  Scaled sums (a site adds
// 1 / count per fragment, make_vplot.py:34-35) are accumulated as 128-bit fixed-point integers -- units of 2^-64, a low word
// with its carries counted into a high word -- because integer addition commutes: the result does not depend on the order in
// which the sites' atomics land (a floating-point atomicAdd would make the last bits vary from run to run), and it is
// closer to the exact sum (each term is off by < 2^-64) than a float64 sum in any order.
__global__ void __launch_bounds__(256) k_vplot(const int32_t *__restrict__ centers, const int32_t *__restrict__ flips,
                                               const int64_t *__restrict__ frag_off, const int32_t *__restrict__ pos,
                                               const int32_t *__restrict__ tlen, int flank, int lower, int upper, int atac,
                                               int scale, double *__restrict__ out, unsigned long long *__restrict__ fix,
                                               int32_t *__restrict__ empty_sites)
{
    __shared__ int red[32];
    const int s = blockIdx.x, c = centers[s], flip = flips[s], ncol = 2 * flank + 1;
    const int64_t f0 = frag_off[s], f1 = frag_off[s + 1];
    unsigned long long w_hi = 0, w_lo = 0;   // 1 / count in units of 2^-64
    if (scale) {
        int cnt = 0;
        for (int64_t f = f0 + threadIdx.x; f < f1; f += blockDim.x) {
            int l, i;
            frag_geometry(pos[f], tlen[f], atac, l, i);
            if (i >= lower && i < upper && vplot_col(l + floordiv2(i - 1), i, c, flank, flip) >= 0) cnt++;
        }
        cnt = warp_sum_i(cnt);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
        __syncthreads();
        cnt = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) cnt += red[w];
        if (cnt == 0) {  // add / np.sum(add) is 0 / 0 in every cell (make_vplot.py:34-35): the whole V-plot turns NaN
            if (threadIdx.x == 0) atomicAdd(empty_sites, 1);
            return;
        }
        if (cnt == 1)
            w_hi = 1;
        else
            w_lo = 0xffffffffffffffffull / (unsigned long long)cnt + ((0xffffffffffffffffull % (unsigned long long)cnt) + 1 == (unsigned long long)cnt ? 1 : 0);  // floor(2^64 / cnt)
    }
    for (int64_t f = f0 + threadIdx.x; f < f1; f += blockDim.x) {
        int l, i;
        frag_geometry(pos[f], tlen[f], atac, l, i);
        if (i < lower || i >= upper) continue;
        const int t = vplot_col(l + floordiv2(i - 1), i, c, flank, flip);
        if (t < 0) continue;
        const size_t cell = (size_t)(i - lower) * ncol + t;
        if (!scale) {
            atomicAdd(&out[cell], 1.0);
        } else {
            if (w_lo) {
                const unsigned long long old = atomicAdd(&fix[2 * cell], w_lo);
                if (old + w_lo < old) atomicAdd(&fix[2 * cell + 1], 1ull);   // carry out of the low word
            }
            if (w_hi) atomicAdd(&fix[2 * cell + 1], w_hi);
        }
    }
}

// fixed-point sums -> float64: hi + lo * 2^-64 (hi is a small integer, the conversion of lo rounds once)
__global__ void k_vplot_finish(const unsigned long long *__restrict__ fix, size_t cells, double *__restrict__ out)
{
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < cells) out[c] = (double)fix[2 * c + 1] + (double)fix[2 * c] * 5.421010862427522e-20;
}

// Fragment-centre histogram over the genomic columns [start - half, start - half + ncol), sizes in [lower, upper)
// (the column sums of the FragmentMat2D of _covHelper, get_cov.py:27-28 + tracks.py:216-218).
__global__ void k_centre_hist(const int32_t *__restrict__ pos, const int32_t *__restrict__ tlen, int64_t n, int col0, int ncol,
                              int lower, int upper, int atac, int32_t *__restrict__ hist)
{
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    int l, i;
    frag_geometry(pos[f], tlen[f], atac, l, i);
    if (i < lower || i >= upper) return;
    const int col = l + floordiv2(i - 1) - col0;
    if (col >= 0 && col < ncol) atomicAdd(&hist[col], 1);
}

// np.convolve(ones(window), colsums, 'valid') (tracks.py:219-222) on the integer histogram: exact.
__global__ void k_flat_window(const int32_t *__restrict__ hist, int window, int64_t nout, double *__restrict__ out)
{
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nout) return;
    int64_t s = 0;
    for (int k = 0; k < window; k++) s += hist[x + k];
    out[x] = (double)s;
}

extern "C" {

int nb200_vplot(nb200_ctx *ctx, int32_t n_sites, const int32_t *centers, const int32_t *flips, const int64_t *frag_off,
                const int32_t *pos, const int32_t *tlen, int32_t flank, int32_t lower, int32_t upper, int32_t atac, int32_t scale,
                double *out)
{
    if (!ctx || !out || n_sites < 0 || (n_sites > 0 && (!centers || !flips || !frag_off)) || flank < 0 || upper <= lower)
        return nb200_fail(ctx, NB200_ERR_ARG, "nb200_vplot: bad argument");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t cells = (size_t)(upper - lower) * (2 * (size_t)flank + 1), bytes = sizeof(double) * cells;
    int32_t empty = 0;
    // layout of s4: float64 plot | empty-site counter (padded to 16 bytes) | 128-bit fixed-point sums (scaled mode)
    const size_t fix_off = bytes + 16, total = fix_off + (scale ? 2 * sizeof(unsigned long long) * cells : 0);
    NB_CUDA(ctx, ctx->s4.reserve(total));
    NB_CUDA(ctx, cudaMemsetAsync(ctx->s4.p, 0, total, ctx->stream));
    if (n_sites > 0) {
        const int64_t n = frag_off[n_sites];
        if (n > 0 && (!pos || !tlen)) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_vplot: NULL read arrays");
        NB_CHECK(up(ctx, ctx->s0, pos, sizeof(int32_t) * n));
        NB_CHECK(up(ctx, ctx->s1, tlen, sizeof(int32_t) * n));
        NB_CHECK(up(ctx, ctx->s2, centers, sizeof(int32_t) * n_sites));
        NB_CHECK(up(ctx, ctx->s3, flips, sizeof(int32_t) * n_sites));
        NB_CHECK(up(ctx, ctx->flush, frag_off, sizeof(int64_t) * (n_sites + 1)));
        int32_t *d_empty = reinterpret_cast<int32_t *>(ctx->s4.as<char>() + bytes);
        {
            ProfScope ps(ctx, ctx->stream, "k_vplot");
            k_vplot<<<n_sites, 256, 0, ctx->stream>>>(ctx->s2.as<int32_t>(), ctx->s3.as<int32_t>(), ctx->flush.as<int64_t>(),
                                                      ctx->s0.as<int32_t>(), ctx->s1.as<int32_t>(), flank, lower, upper, atac, scale,
                                                      ctx->s4.as<double>(), reinterpret_cast<unsigned long long *>(ctx->s4.as<char>() + fix_off), d_empty);
            NB_LAUNCH_CHECK(ctx);
            if (scale) {
                k_vplot_finish<<<(unsigned)div_up64((int64_t)cells, 256), 256, 0, ctx->stream>>>(
                    reinterpret_cast<unsigned long long *>(ctx->s4.as<char>() + fix_off), cells, ctx->s4.as<double>());
                NB_LAUNCH_CHECK(ctx);
            }
        }
        NB_CUDA(ctx, cudaMemcpyAsync(&empty, d_empty, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    NB_CUDA(ctx, cudaMemcpyAsync(out, ctx->s4.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    NB_CUDA(ctx, cudaGetLastError());
    if (empty > 0) {
        const double qnan = __builtin_nan("");
        for (size_t i = 0; i < cells; i++) out[i] = qnan;
    }
    return NB200_OK;
}

int nb200_coverage(nb200_ctx *ctx, const int32_t *pos, const int32_t *tlen, int64_t n, int32_t start, int32_t end, int32_t lower,
                   int32_t upper, int32_t window_len, int32_t atac, double *out)
{
    if (!ctx || !out || n < 0 || (n > 0 && (!pos || !tlen)) || end <= start || upper <= lower || window_len < 1)
        return nb200_fail(ctx, NB200_ERR_ARG, "nb200_coverage: bad argument");
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    // the matrix of _covHelper spans [start - half, end + half); smooth() makes an even window one longer (utils.py:34-36)
    const int half = window_len / 2, ncol = (end - start) + 2 * half, weff = window_len + (window_len % 2 == 0);
    const int64_t nout = (int64_t)ncol - weff + 1;                      // = end - start
    if (nout < 1) return nb200_fail(ctx, NB200_ERR_FLANK, "Insufficient flanking region on mat to calculate coverage with desired window");
    NB_CHECK(up(ctx, ctx->s0, pos, sizeof(int32_t) * n));
    NB_CHECK(up(ctx, ctx->s1, tlen, sizeof(int32_t) * n));
    NB_CUDA(ctx, ctx->s2.reserve(sizeof(int32_t) * (size_t)ncol));
    NB_CUDA(ctx, cudaMemsetAsync(ctx->s2.p, 0, sizeof(int32_t) * (size_t)ncol, ctx->stream));
    NB_CUDA(ctx, ctx->s3.reserve(sizeof(double) * nout));
    if (n > 0) {
        ProfScope ps(ctx, ctx->stream, "k_centre_hist");
        k_centre_hist<<<(unsigned)div_up64(n, 256), 256, 0, ctx->stream>>>(ctx->s0.as<int32_t>(), ctx->s1.as<int32_t>(), n, start - half,
                                                                           ncol, lower, upper, atac, ctx->s2.as<int32_t>());
        NB_LAUNCH_CHECK(ctx);
    }
    {
        ProfScope ps(ctx, ctx->stream, "k_flat_window");
        k_flat_window<<<(unsigned)div_up64(nout, 256), 256, 0, ctx->stream>>>(ctx->s2.as<int32_t>(), weff, nout, ctx->s3.as<double>());
        NB_LAUNCH_CHECK(ctx);
    }
    NB_CUDA(ctx, cudaMemcpyAsync(out, ctx->s3.p, sizeof(double) * nout, cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    NB_CUDA(ctx, cudaGetLastError());
    return NB200_OK;
}

}  // extern "C"
