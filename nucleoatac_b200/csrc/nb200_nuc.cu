// nb200_nuc.cu -- NucChunk.process on the device (nucleoatac/NucleosomeCalling.py:230-345, run_nuc.py:22-39),
// without fit/getFuzz (host scipy optimiser).
//
// Stages (per batch, every chunk in parallel):
//   k_pair_colsums  cB[c] = sum_{i in [lv,uv)} f_i * Bp[i,c]                      (bias coverage operand, :56-58)
//   k_nuc_bx_fp64   bx[x] = sum_i sum_k f_i V[i,k] Bp[i, x-w+k]   dense background xcor, fp64 CUDA cores  (:60-63)
//                   (nb200_xcor_tc.cu holds the tcgen05 version of the same contraction)
//   k_nuc_tracks    nuc_cov / nfr_cov from the CSC prefix, bias coverage, sparse signal xcor (:29-36),
//                   background = bx*nuc_cov/bcov (:64), norm = signal - background (:42-43)
//   k_smooth_same   gaussian smoothing of max(norm,0) (:274-283)
//   k_nuc_peaks     candidates = call_peaks(norm+smoothed, order 12, sep 25, boundary 60)   (:294-301)
//   k_cand_stats    per candidate: likelihood ratio (:110-122), closed-form multinomial variance
//                   (multinomial_cov.pyx:20-31) and z-score (:123-127), fp64, one block per candidate
//   k_nuc_reduce    nonredundant set = reduce_peaks by z, sep 120 (:312-315)
#include "nb200_dev.cuh"

// ---------------------------------------------------------------------------------------------
// Dense background cross-correlation, fp64.  One block = BX_NT threads x BX_XT consecutive outputs
// each; for every insert-size row the pre-normalisation bias row Bp_i is generated on the fly from
// the 1-D track E (never materialised in HBM) into shared memory in an XT-way interleaved layout so
// that the sliding-window register tile reads it conflict-free; T = f_i*V rows are zero padded to a
// multiple of XT.  2*R*Wpad flops per output.
// ---------------------------------------------------------------------------------------------
#define BX_NT 64
#define BX_XT 16
#define BX_TX (BX_NT * BX_XT)
__global__ void __launch_bounds__(BX_NT) k_nuc_bx_fp64(const int32_t *__restrict__ start, const int64_t *__restrict__ out_off,
                                                       const int64_t *__restrict__ bias_off,
                                                       const int32_t *__restrict__ seq_start, int pwm_up,
                                                       const double *__restrict__ E, const double *__restrict__ Tfp, int lv,
                                                       int R, int w, int wpad, double *__restrict__ bx)
{
    extern __shared__ double sm_bx[];
    const int c = blockIdx.y;
    const int64_t oo = out_off[c];
    const int L = (int)(out_off[c + 1] - oo);
    const int x0 = blockIdx.x * BX_TX;
    if (x0 >= L) return;
    const int uv = lv + R;
    const int half = uv / 2;
    const int nB = BX_TX + wpad;            // Bp values per row tile (those past Tx+W-1 are multiplied by T == 0)
    const int Q = nB / BX_XT;               // interleave stride
    const int nE = nB + 2 * half + 2;
    double *s_E = sm_bx;                    // [nE]   E over genomic [g0 - w - half, ...)
    double *s_B = s_E + nE;                 // [2][nB]
    double *s_T = s_B + 2 * nB;             // [2][wpad]
    const int g0 = start[c] + x0;
    const int64_t eb = bias_off[c] - (int64_t)(seq_start[c] + pwm_up);
    const int64_t e_lo = bias_off[c], e_hi = bias_off[c + 1];
    for (int i = threadIdx.x; i < nE; i += BX_NT) {
        int64_t idx = eb + g0 - w - half + i;
        s_E[i] = (idx >= e_lo && idx < e_hi) ? E[idx] : 0.0;  // beyond the track only under zero T padding
    }
    __syncthreads();
    auto fill = [&](int r, int buf) {
        const int i = lv + r;
        double *B = s_B + buf * nB;
        for (int n = threadIdx.x; n < nB; n += BX_NT)
            B[(n % BX_XT) * Q + n / BX_XT] = bias_cell(s_E + half + n, i);
        double *T = s_T + buf * wpad;
        for (int k = threadIdx.x; k < wpad; k += BX_NT) T[k] = Tfp[(size_t)r * wpad + k];
    };
    double acc[BX_XT];
#pragma unroll
    for (int u = 0; u < BX_XT; u++) acc[u] = 0.0;
    fill(0, 0);
    __syncthreads();
    for (int r = 0; r < R; r++) {
        if (r + 1 < R) fill(r + 1, (r + 1) & 1);
        const double *B = s_B + (r & 1) * nB + threadIdx.x;   // element n = t*XT + e  ->  B[(e%XT)*Q + e/XT]
        const double *T = s_T + (r & 1) * wpad;
        double win[BX_XT];
#pragma unroll
        for (int u = 0; u < BX_XT; u++) win[u] = B[u * Q];
        for (int k = 0; k < wpad; k += BX_XT) {
            const int q = k / BX_XT + 1;
#pragma unroll
            for (int kk = 0; kk < BX_XT; kk++) {
                const double t = T[k + kk];
#pragma unroll
                for (int u = 0; u < BX_XT; u++) acc[u] = fma(win[(u + kk) % BX_XT], t, acc[u]);
                win[kk] = B[kk * Q + q];   // element k + kk + XT
            }
        }
        __syncthreads();
    }
    const int xb = x0 + threadIdx.x * BX_XT;
#pragma unroll
    for (int u = 0; u < BX_XT; u++)
        if (xb + u < L) bx[oo + xb + u] = acc[u];
}

int nb200_nuc_bx_fp64(nb200_ctx *ctx, nb200_dbatch *b)
{
    RunConst &r = ctx->rc;
    const int nB = BX_TX + r.v_wpad;
    const int half = r.v_upper / 2;
    size_t smem = sizeof(double) * ((size_t)nB + 2 * half + 2 + 2 * (size_t)nB + 2 * (size_t)r.v_wpad);
    if (smem > 200 * 1024) return nb200_fail(ctx, NB200_ERR_ARG, "VMat too large for the fp64 background kernel");
    if (smem > 48 * 1024) NB_CUDA(ctx, cudaFuncSetAttribute(k_nuc_bx_fp64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(ctx, b->stream, "k_nuc_bx_fp64");
    dim3 grid((unsigned)div_up64(b->max_len, BX_TX), b->n_chunks);
    k_nuc_bx_fp64<<<grid, BX_NT, smem, b->stream>>>(b->d_start.as<int32_t>(), b->d_out_off.as<int64_t>(),
                                                    b->d_bias_off.as<int64_t>(), b->d_seq_start.as<int32_t>(), r.pwm_up,
                                                    b->d_E.as<double>(), r.vmat_fp.as<double>(), r.v_lower, r.v_rows, r.v_w,
                                                    r.v_wpad, b->n_bx.as<double>());
    NB_LAUNCH_CHECK(ctx);
    return NB200_OK;
}

// ---------------------------------------------------------------------------------------------
struct NucTrackArgs {
    const int64_t *out_off, *col_off, *frag_off;
    const int32_t *col_ptr, *col_low;
    const int2 *ent;
    const double *V, *cB, *bx;
    double *nuc_cov, *nfr_cov, *bcov, *signal, *bg, *norm;
    int lv, uv, W, w, csc_pad, use_bias;
    double bcov_nobias, bx_nobias;
};

#ifndef NT_TILE
#define NT_TILE 256
#endif
#define NT_FRAG_CAP 512    // fragments of a tile staged in shared memory (denser tiles read them from global memory); kept small:
                           // the VMat gathers are L2 -> L1 traffic and every KB of shared memory is a KB less of L1
__global__ void __launch_bounds__(NT_TILE) k_nuc_tracks(NucTrackArgs a)
{
    extern __shared__ __align__(16) double sm_nt[];  // cB tile [NT_TILE + 2w (+8)], sums of 8 [..], fragment tile
    const int c = blockIdx.y;
    const int64_t oo = a.out_off[c];
    const int L = (int)(a.out_off[c + 1] - oo);
    const int x0 = blockIdx.x * NT_TILE;
    if (x0 >= L) return;
    const int ncb = (NT_TILE + 2 * a.w + 8) & ~7;
    double *s_cb = sm_nt, *s_b8 = sm_nt + ncb;                       // s_b8[m] = sum of s_cb[8m .. 8m+7]
    int2 *s_ent = reinterpret_cast<int2 *>(sm_nt + ncb + ncb / 8);
    const int32_t *cp = a.col_ptr + a.col_off[c];
    const int2 *en = a.ent + a.frag_off[c];
    const int nx = min(NT_TILE, L - x0);
    // fragments of the whole tile: columns [x0 - w, x0 + nx - 1 + w]
    const int te0 = cp[x0 - a.w + a.csc_pad], te1 = cp[x0 + nx - 1 + a.w + 1 + a.csc_pad];
    const bool staged = (te1 - te0) <= NT_FRAG_CAP;
    if (staged)
        for (int i = threadIdx.x; i < te1 - te0; i += NT_TILE) s_ent[i] = en[te0 + i];
    if (a.use_bias) {
        const int64_t co = oo + 2 * (int64_t)a.w * c + x0;
        const int nc = nx + 2 * a.w;
        for (int i = threadIdx.x; i < ncb; i += NT_TILE) s_cb[i] = (i < nc) ? a.cB[co + i] : 0.0;
        __syncthreads();
        for (int m = threadIdx.x; m < ncb / 8; m += NT_TILE) {
            const double *q = s_cb + 8 * m;
            s_b8[m] = ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
        }
    }
    __syncthreads();
    const int x = x0 + threadIdx.x;
    const bool valid = x < L;
    if (!__any_sync(NB_FULL, valid)) return;
    const int xc = valid ? x : L - 1;   // lanes past the chunk end shadow its last position and store nothing
    const int lo = xc - a.w + a.csc_pad, hi = xc + a.w + 1 + a.csc_pad;
    const int e0 = cp[lo], e1 = cp[hi];
    int nlow = 0;
    if (a.col_low) {
        const int32_t *cl = a.col_low + a.col_off[c];
        nlow = cl[hi] - cl[lo];
    }
    const double nuc_cov = (double)(e1 - e0 - nlow), nfr_cov = (double)nlow;
    double bcov, bxv;
    if (a.use_bias) {
        // window [t, t + W) = head up to the next multiple of 8, whole groups of 8, tail
        const int t = xc - x0, t1 = t + a.W;
        const int g0 = (t + 7) >> 3, g1 = t1 >> 3;
        double s = 0.0;
        if (g0 <= g1) {
            for (int k = t; k < 8 * g0; k++) s += s_cb[k];
            for (int m = g0; m < g1; m++) s += s_b8[m];
            for (int k = 8 * g1; k < t1; k++) s += s_cb[k];
        } else
            for (int k = t; k < t1; k++) s += s_cb[k];
        bcov = s;
        bxv = a.bx[oo + xc];
    } else {
        bcov = a.bcov_nobias;
        bxv = a.bx_nobias;
    }
    // sparse signal xcor: every fragment centred within +-w contributes one VMat entry.  The warp walks the union of its
    // lanes' fragment ranges in lockstep: the entry is a broadcast read, its VMat row is read by the lanes whose window
    // holds the fragment at consecutive (descending) columns -- one coalesced segment per fragment instead of one
    // scattered gather per lane.  Every lane still adds its own fragments in ascending order (same sums as a per-lane loop).
    double sig = 0.0;
    const int kb = a.w - (xc + a.csc_pad);
    const int we0 = __shfl_sync(NB_FULL, e0, 0), we1 = __reduce_max_sync(NB_FULL, e1);
    const int2 *se = staged ? (s_ent - te0) : en;
    for (int e = we0; e < we1; e++) {
        const int2 v = se[e];
        if (v.y < a.lv || v.y >= a.uv) continue;   // warp-uniform
        if (e >= e0 && e < e1) sig += __ldg(a.V + (size_t)(v.y - a.lv) * a.W + (v.x + kb));
    }
    if (!valid) return;
    const double bg = bxv * nuc_cov / bcov;  // NucleosomeCalling.py:64
    a.nuc_cov[oo + x] = nuc_cov;
    a.nfr_cov[oo + x] = nfr_cov;
    a.bcov[oo + x] = bcov;
    a.signal[oo + x] = sig;
    a.bg[oo + x] = bg;
    a.norm[oo + x] = sig - bg;
}

// ---------------------------------------------------------------------------------------------
struct NucPeakArgs {
    const int32_t *start;
    const int64_t *out_off, *cand_off;
    const double *jitter, *norm, *smooth, *signal, *nuc_cov, *nfr_cov;
    double *comb;                    // scratch track
    int32_t *sc_pos;
    double *sc_val;
    unsigned char *sc_state;
    int32_t *cand_count, *cand_pos, *cand_flag;
    double *cand_z, *cand_lr, *cand_norm, *cand_sig, *cand_cov, *cand_nfr, *cand_smooth, *cand_bcov;
    const double *bcov;
    int2 *work;
    int32_t *work_count;
    int sep, boundary, order;
    double min_signal, min_reads;
};

#define PK_THREADS_N 512
__global__ void __launch_bounds__(PK_THREADS_N) k_nuc_peaks(NucPeakArgs a)
{
    __shared__ double red_d[32];
    __shared__ int red_i[32];
    __shared__ int s_cnt[32 * (PK_THREADS_N / 32)];
    __shared__ int s_base, s_flag;
    const int c = blockIdx.x;
    const int64_t oo = a.out_off[c];
    const int L = (int)(a.out_off[c + 1] - oo);
    double *cb = a.comb + oo;
    int32_t *cpos = a.sc_pos + oo;
    double *cval = a.sc_val + oo;
    unsigned char *cst = a.sc_state + oo;
    const int tid = threadIdx.x;
    double mn = CUDART_INF;
    int nnan = 0;
    {   // four positions in flight per thread: the pass is load latency
        const double *__restrict__ nr = a.norm + oo, *__restrict__ sr = a.smooth + oo;
        double *__restrict__ cw = cb;
        const int stride = (int)blockDim.x;
        int x = tid;
        for (; x + 3 * stride < L; x += 4 * stride) {
            double v4[4];
#pragma unroll
            for (int u = 0; u < 4; u++) v4[u] = nr[x + u * stride] + sr[x + u * stride];  // NucleosomeCalling.py:297
#pragma unroll
            for (int u = 0; u < 4; u++) {
                cw[x + u * stride] = v4[u];
                if (v4[u] != v4[u]) nnan++;
                else mn = fmin(mn, v4[u]);
            }
        }
        for (; x < L; x += stride) {
            const double v = nr[x] + sr[x];
            cw[x] = v;
            if (v != v) nnan++;
            else mn = fmin(mn, v);
        }
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
    for (int wv = 0; wv < (int)(blockDim.x >> 5); wv++) {
        mn = fmin(mn, red_d[wv]);
        nnan += red_i[wv];
    }
    __syncthreads();
    if (tid == 0) s_base = 0;
    __syncthreads();
    int m = 0;
    if (nnan < L) {
        if (nnan > 0)
            for (int x = tid; x < L; x += blockDim.x)
                if (cb[x] != cb[x]) cb[x] = mn;
        __syncthreads();
        const int lo = max(0, a.boundary), hi = L - a.boundary;
        block_compact_ordered(
            L, s_cnt, red_i, &s_base,
            [&](int x) {
                if (x < lo || x >= hi) return false;
                const double v = cb[x];
                const double j0 = v * (1.0 + a.jitter[x]);
                bool flag = (v >= a.min_signal);
                for (int d = 1; d <= a.order && flag; d++) {  // argrelmax(order), mode='clip'
                    const int xl = max(x - d, 0), xr = min(x + d, L - 1);
                    flag = (j0 > cb[xl] * (1.0 + a.jitter[xl])) && (j0 > cb[xr] * (1.0 + a.jitter[xr]));
                }
                return flag;
            },
            [&](int x, int slot) {
                cpos[slot] = x;
                cval[slot] = cb[x];
            });
        m = s_base;
        block_nms(cpos, cval, cst, m, a.sep, &s_flag);
    }
    __syncthreads();
    if (tid == 0) s_base = 0;
    __syncthreads();
    const int64_t po = a.cand_off[c];
    const int cap = (int)(a.cand_off[c + 1] - po);
    for (int j0 = 0; j0 < m; j0 += blockDim.x) {
        const int j = j0 + tid;
        const int flag = (j < m && cst[j] == 1);
        int slot = block_compact_slot(flag, &s_base, red_i);
        if (flag && slot < cap) {
            const int x = cpos[j];
            const double cov = a.nuc_cov[oo + x];
            a.cand_pos[po + slot] = a.start[c] + x;
            a.cand_norm[po + slot] = a.norm[oo + x];
            a.cand_sig[po + slot] = a.signal[oo + x];
            a.cand_cov[po + slot] = cov;
            a.cand_nfr[po + slot] = a.nfr_cov[oo + x];
            a.cand_smooth[po + slot] = a.smooth[oo + x];
            a.cand_bcov[po + slot] = a.bcov[oo + x];
            a.cand_z[po + slot] = nb_nan();
            a.cand_lr[po + slot] = nb_nan();
            int fl = 0;
            if (cov > a.min_reads) {  // NucleosomeCalling.py:304
                fl = 1;
                int wslot = atomicAdd(&a.work_count[0], 1);
                a.work[wslot] = make_int2(c, slot);
            }
            a.cand_flag[po + slot] = fl;
        }
    }
    __syncthreads();
    if (tid == 0) a.cand_count[c] = (s_base <= cap) ? s_base : -s_base;
}

// ---------------------------------------------------------------------------------------------
struct CandArgs {
    const int32_t *start;
    const int64_t *out_off, *col_off, *frag_off, *bias_off, *cand_off;
    const int32_t *seq_start;
    const int32_t *col_ptr;
    const int2 *ent;
    const double *E, *V, *f;
    const double2 *pair;       // [3][J2][W2] paired templates V, f*V, f*V^2 (RunConst::vp_pair)
    const double *one;         // [3][W2]     their insert-size-1 rows
    const int2 *work;
    const int32_t *work_count;
    const int32_t *cand_pos;
    int32_t *cand_flag;
    const double *cand_norm, *cand_cov, *cand_bcov;  // cand_bcov = S_B = sum f*Bp over the window (bias coverage track)
    double *cand_z, *cand_lr;
    int pwm_up, lv, R, W, w, csc_pad, use_bias, lr_is_nan, J2, W2;
    double min_lr, min_z;
};

#define CS_THREADS 256
#define CS_GROUP_DEFAULT 4   // candidates scored together: every template element is loaded once per group

// Dense window sum  sum_{i,k} T[i,k] * Bp[i, c_k]  (c_k = P - w + k) for NG candidates at once, T given as a paired template.
// Sizes 2j+1 and 2j+2 share their left tap (Bp[2j+1,c] = E[c-j] E[c+j], Bp[2j+2,c] = E[c-j] E[c+j+1]), so a column costs
//   sum_j E[c-j] * (T[2j+1,k] * E[c+j] + T[2j+2,k] * E[c+j+1])          -- 1.5 flops-pairs per cell instead of 2.
// A thread owns two adjacent columns and walks j two steps at a time: the taps of (k, k+1) x (j, j+1) are two sliding
// aligned pairs per side, i.e. one 16-byte shared-memory load per side per candidate for 8 cells.  Thread groups split
// the j range when the window has fewer than 2 * CS_THREADS columns.  s_E: per candidate nEw2 doubles, element off0 + k
// is E[c_k]; both even, so every pair is 16-byte aligned.  Adds the partial sums of this thread to acc[0..NG).
template <int NG>
__device__ __forceinline__ void pair_window_sums(const double2 *__restrict__ T, const double *__restrict__ t1, int J2, int W2,
                                                 const double *s_E, int nEw2, int off0, double *acc)
{
    const int ncp = W2 >> 1, tpc = min(ncp, CS_THREADS), rows_par = CS_THREADS / tpc;
    const int cp0 = threadIdx.x % tpc, rr = threadIdx.x / tpc;
    const int jseg = ((J2 / 2 + rows_par - 1) / rows_par) * 2;
    const int j0 = rr * jseg, j1 = min(J2, j0 + jseg);
    if (rr >= rows_par || j0 >= j1) return;
    for (int cp = cp0; cp < ncp; cp += tpc) {
        const int k = 2 * cp;
        const double *Ec = s_E + off0 + k;
        double2 L[NG], Rt[NG];   // carries: (E[c-j], E[c-j+1]) and (E[c+j], E[c+j+1])
#pragma unroll
        for (int cg = 0; cg < NG; cg++) {
            L[cg] = *reinterpret_cast<const double2 *>(Ec + cg * nEw2 - j0);
            Rt[cg] = *reinterpret_cast<const double2 *>(Ec + cg * nEw2 + j0);
        }
        if (j0 == 0) {  // insert size 1: a single tap, Bp[1,c] = E[c]
            const double2 c1 = *reinterpret_cast<const double2 *>(t1 + k);
#pragma unroll
            for (int cg = 0; cg < NG; cg++) acc[cg] = fma(c1.x, L[cg].x, fma(c1.y, L[cg].y, acc[cg]));
        }
        const double2 *Tp = T + (size_t)j0 * W2 + k;
        double2 a0 = __ldg(Tp), a1 = __ldg(Tp + 1), b0 = __ldg(Tp + W2), b1 = __ldg(Tp + W2 + 1);
#pragma unroll 2
        for (int j = j0; j < j1; j += 2) {
            Tp += 2 * (size_t)W2;
            double2 na0 = a0, na1 = a1, nb0 = b0, nb1 = b1;
            if (j + 2 < j1) {   // next step's template values are in flight while this step is contracted
                na0 = __ldg(Tp);
                na1 = __ldg(Tp + 1);
                nb0 = __ldg(Tp + W2);
                nb1 = __ldg(Tp + W2 + 1);
            }
#pragma unroll
            for (int cg = 0; cg < NG; cg++) {
                const double2 Ln = *reinterpret_cast<const double2 *>(Ec + cg * nEw2 - j - 2);   // (E[c-j-2], E[c-j-1])
                const double2 Rn = *reinterpret_cast<const double2 *>(Ec + cg * nEw2 + j + 2);   // (E[c+j+2], E[c+j+3])
                const double2 Lc = L[cg], Rc = Rt[cg];
                double s = acc[cg];
                s = fma(Lc.x, fma(a0.x, Rc.x, a0.y * Rc.y), s);   // column k,   step j
                s = fma(Lc.y, fma(a1.x, Rc.y, a1.y * Rn.x), s);   // column k+1, step j
                s = fma(Ln.y, fma(b0.x, Rc.y, b0.y * Rn.x), s);   // column k,   step j+1
                s = fma(Lc.x, fma(b1.x, Rn.x, b1.y * Rn.y), s);   // column k+1, step j+1
                acc[cg] = s;
                L[cg] = Ln;
                Rt[cg] = Rn;
            }
            a0 = na0;
            a1 = na1;
            b0 = nb0;
            b1 = nb1;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Candidate screen.  A candidate survives NucleosomeCalling.py:302-310 only if its likelihood ratio exceeds min_lr, and
// the reference keeps nothing of a candidate that does not (its LR is never written).  On real and synthetic data well
// under 1 % of the candidates pass, yet every one of them needs the dense window sum S_VB = sum V*Bp (63 001 cells at
// 251x251).  The screen evaluates S_VB in fp32 (same paired two-tap walk, half the shared-memory / L1 traffic per cell,
// 8 candidates per template load), forms the likelihood ratio with it, and passes on to the exact fp64 kernel
// (k_cand_stats) every candidate whose approximate LR is not below min_lr by more than a rigorous error margin:
// |LR32 - LR| = n * |log(S32 / S_VB)| <= n * CS_EPS32, where n = fragments in the window and CS_EPS32 is ~8x the worst-case
// rounding of the fp32 accumulation chains (<= ~190 FMAs of positive terms per thread, 1.2e-5).  Anything non-finite goes to
// the exact kernel too.  Decisions, LR and z of every kept nucleosome therefore come from the fp64 path; a rejected
// candidate's cand_lr holds the fp32-screen value (within n * CS_EPS32 of the exact one).
#define CS_EPS32 1e-4
#ifndef CS_SCREEN_GROUP
#define CS_SCREEN_GROUP 8
#endif

template <int NG>
__device__ __forceinline__ void pair_window_sums_f32(const float2 *__restrict__ T, const float *__restrict__ t1, int J2, int W2,
                                                     const float *s_E, int nEw2, int off0, float *acc)
{
    const int ncp = W2 >> 1, tpc = min(ncp, CS_THREADS), rows_par = CS_THREADS / tpc;
    const int cp0 = threadIdx.x % tpc, rr = threadIdx.x / tpc;
    const int jseg = ((J2 / 2 + rows_par - 1) / rows_par) * 2;
    const int j0 = rr * jseg, j1 = min(J2, j0 + jseg);
    if (rr >= rows_par || j0 >= j1) return;
    for (int cp = cp0; cp < ncp; cp += tpc) {
        const int k = 2 * cp;
        const float *Ec = s_E + off0 + k;
        float2 L[NG], Rt[NG];   // carries: (E[c-j], E[c-j+1]) and (E[c+j], E[c+j+1])
#pragma unroll
        for (int cg = 0; cg < NG; cg++) {
            L[cg] = *reinterpret_cast<const float2 *>(Ec + cg * nEw2 - j0);
            Rt[cg] = *reinterpret_cast<const float2 *>(Ec + cg * nEw2 + j0);
        }
        if (j0 == 0) {  // insert size 1: a single tap
            const float2 c1 = *reinterpret_cast<const float2 *>(t1 + k);
#pragma unroll
            for (int cg = 0; cg < NG; cg++) acc[cg] = fmaf(c1.x, L[cg].x, fmaf(c1.y, L[cg].y, acc[cg]));
        }
        const float2 *Tp = T + (size_t)j0 * W2 + k;
        float2 a0 = __ldg(Tp), a1 = __ldg(Tp + 1), b0 = __ldg(Tp + W2), b1 = __ldg(Tp + W2 + 1);
#pragma unroll 2
        for (int j = j0; j < j1; j += 2) {
            Tp += 2 * (size_t)W2;
            float2 na0 = a0, na1 = a1, nb0 = b0, nb1 = b1;
            if (j + 2 < j1) {
                na0 = __ldg(Tp);
                na1 = __ldg(Tp + 1);
                nb0 = __ldg(Tp + W2);
                nb1 = __ldg(Tp + W2 + 1);
            }
#pragma unroll
            for (int cg = 0; cg < NG; cg++) {
                const float2 Ln = *reinterpret_cast<const float2 *>(Ec + cg * nEw2 - j - 2);
                const float2 Rn = *reinterpret_cast<const float2 *>(Ec + cg * nEw2 + j + 2);
                const float2 Lc = L[cg], Rc = Rt[cg];
                float s = acc[cg];
                s = fmaf(Lc.x, fmaf(a0.x, Rc.x, a0.y * Rc.y), s);
                s = fmaf(Lc.y, fmaf(a1.x, Rc.y, a1.y * Rn.x), s);
                s = fmaf(Ln.y, fmaf(b0.x, Rc.y, b0.y * Rn.x), s);
                s = fmaf(Lc.x, fmaf(b1.x, Rn.x, b1.y * Rn.y), s);
                acc[cg] = s;
                L[cg] = Ln;
                Rt[cg] = Rn;
            }
            a0 = na0;
            a1 = na1;
            b0 = nb0;
            b1 = nb1;
        }
    }
}

// First stage of the cascade: an upper bound of the likelihood ratio from quantities the pass already has.  With V >= 0
// and f >= 0,  bx = sum f_i V Bp <= f_max * sum V Bp = f_max * S_VB  (bx = the background cross-correlation at the
// candidate, NucleosomeCalling.py:60-63), so log S_VB >= log(bx / f_max) and
//     LR <= sum_frag [ log(V bp f_max / bx) - log(bp f / S_B) ].
// bx may come from the tensor-core path (relative error ~1e-6 of the chunk's scale, all terms positive): it is taken at
// half its value, which costs n*log 2 of slack and tolerates any plausible rounding of bx.  On the synthetic workload
// and the example data the bound alone still rejects ~96 % of the candidates with a warp-sized sparse sum;
// the rest go to the fp32 screen.  A candidate rejected here keeps the bound in cand_lr (the reference drops its LR).
struct BoundArgs {
    CandArgs c;
    const double *bx;          // raw background xcor track (packed per position), or nullptr without bias
    double bx_nobias, f_max;
    int bound_ok;              // V >= 0, f >= 0, f_max > 0
    int2 *next;                // work list of the screen
    int32_t *next_count;
};

__global__ void __launch_bounds__(256) k_cand_bound(BoundArgs ba)
{
    const CandArgs &a = ba.c;
    const int nwork = a.work_count[0];
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int wk = blockIdx.x * wpb + (threadIdx.x >> 5); wk < nwork; wk += gridDim.x * wpb) {
        const int2 it = a.work[wk];
        const int c = it.x;
        const int64_t ci = a.cand_off[c] + it.y;
        if (a.lr_is_nan) {  // 0*log(0) cells make both likelihoods NaN in the reference: nothing passes
            if (lane == 0) a.cand_lr[ci] = nb_nan();
            continue;
        }
        bool pass_on = !ba.bound_ok;
        double lr_ub = 0.0;
        int nfr = 0;
        if (ba.bound_ok) {
            const int P = a.cand_pos[ci], x = P - a.start[c];
            const double bxv = a.use_bias ? ba.bx[a.out_off[c] + x] : ba.bx_nobias;
            const double cB = a.cand_bcov[ci];
            const int32_t *cp = a.col_ptr + a.col_off[c];
            const int2 *en = a.ent + a.frag_off[c];
            const int e0 = cp[x - a.w + a.csc_pad], e1 = cp[x + a.w + 1 + a.csc_pad];
            const int kb = a.w - (x + a.csc_pad);
            const double *Eg = a.use_bias ? a.E + (a.bias_off[c] - (int64_t)(a.seq_start[c] + a.pwm_up) + (P - a.w)) : nullptr;
            const double c1 = ba.f_max / (0.5 * bxv);
            double nl = 0.0, ul = 0.0;
            for (int e = e0 + lane; e < e1; e += 32) {
                const int2 v = en[e];
                const int r = v.y - a.lv;
                if (r >= 0 && r < a.R) {
                    const int k = v.x + kb;
                    const double bp = a.use_bias ? bias_cell(Eg + k, v.y) : 1.0;
                    nl += log(a.V[(size_t)r * a.W + k] * bp * c1);
                    ul += log(__dmul_rn(bp, a.f[a.lv + r]) / cB);
                }
            }
            nl = warp_sum(nl);
            ul = warp_sum(ul);
            lr_ub = nl - ul;
            nfr = e1 - e0;
            const double margin = (double)nfr * CS_EPS32 + 1e-6;
            pass_on = !(bxv > 1e-300 && bxv < 1e300) || !(lr_ub <= a.min_lr - margin);
        }
        if (lane == 0) {
            if (pass_on) {
                const int slot = atomicAdd(ba.next_count, 1);
                ba.next[slot] = it;
            } else
                a.cand_lr[ci] = lr_ub;
        }
    }
}

struct ScreenArgs {
    CandArgs c;
    const float2 *pair32;
    const float *one32;
    int2 *confirm;            // work list of the exact kernel
    int32_t *confirm_count;
};

__global__ void __launch_bounds__(CS_THREADS, 2) k_cand_screen(ScreenArgs sa)
{
    constexpr int NG = CS_SCREEN_GROUP, NWARP = CS_THREADS / 32;
    static_assert(NWARP % NG == 0, "warps must divide evenly over the candidates of a group");
    const CandArgs &a = sa.c;
    extern __shared__ __align__(16) float sm_sc[];
    __shared__ float s_part[NG][NWARP];
    const int off0 = a.J2 + 2, nEw = off0 + a.W2 + a.J2 + 4;   // same window layout as k_cand_stats
    float *s_E = sm_sc;                      // [NG][nEw]
    const int nwork = a.work_count[0];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int g0 = blockIdx.x * NG; g0 < nwork; g0 += gridDim.x * NG) {
        const int ng = min(NG, nwork - g0);
        if (a.lr_is_nan) {  // 0*log(0) cells make both likelihoods NaN in the reference: nothing passes
            if (tid < ng) {
                const int2 it = a.work[g0 + tid];
                a.cand_lr[a.cand_off[it.x] + it.y] = nb_nan();
            }
            continue;
        }
        __syncthreads();
        for (int cg = 0; cg < NG; cg++) {
            if (cg < ng && a.use_bias) {
                const int2 it = a.work[g0 + cg];
                const int c = it.x;
                const int P = a.cand_pos[a.cand_off[c] + it.y];
                const int64_t e_lo = a.bias_off[c], e_hi = a.bias_off[c + 1];
                const int64_t eb = e_lo - (int64_t)(a.seq_start[c] + a.pwm_up) + (P - a.w - off0);
                for (int i = tid; i < nEw; i += CS_THREADS) {
                    const int64_t idx = eb + i;
                    s_E[cg * nEw + i] = (idx >= e_lo && idx < e_hi) ? (float)a.E[idx] : 0.0f;
                }
            } else
                for (int i = tid; i < nEw; i += CS_THREADS) s_E[cg * nEw + i] = a.use_bias ? 0.0f : 1.0f;
        }
        __syncthreads();
        float sVB[NG];
#pragma unroll
        for (int cg = 0; cg < NG; cg++) sVB[cg] = 0.0f;
        pair_window_sums_f32<NG>(sa.pair32, sa.one32, a.J2, a.W2, s_E, nEw, off0, sVB);
#pragma unroll
        for (int cg = 0; cg < NG; cg++) {
            float v = sVB[cg];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(NB_FULL, v, o);
            if (lane == 0) s_part[cg][wid] = v;
        }
        __syncthreads();
        // sparse likelihoods (NucleosomeCalling.py:110-122) with the fp32 normaliser; fp64 bias cells straight from the track
        {
            constexpr int WPC = NWARP / NG;
            const int cg = wid % NG, sub = wid / NG;
            if (cg < ng) {
                double cVB = 0.0;
#pragma unroll
                for (int w2 = 0; w2 < NWARP; w2++) cVB += (double)s_part[cg][w2];
                const int2 it = a.work[g0 + cg];
                const int c = it.x;
                const int64_t ci = a.cand_off[c] + it.y;
                const double cB = a.cand_bcov[ci];
                const int P = a.cand_pos[ci], x = P - a.start[c];
                const int32_t *cp = a.col_ptr + a.col_off[c];
                const int2 *en = a.ent + a.frag_off[c];
                const int e0 = cp[x - a.w + a.csc_pad], e1 = cp[x + a.w + 1 + a.csc_pad];
                const int kb = a.w - (x + a.csc_pad);
                const double *Eg = a.use_bias ? a.E + (a.bias_off[c] - (int64_t)(a.seq_start[c] + a.pwm_up) + (P - a.w)) : nullptr;   // E[P - w + k] = Eg[k]
                double nl = 0.0, ul = 0.0;
                for (int e = e0 + sub * 32 + lane; e < e1; e += 32 * WPC) {
                    const int2 v = en[e];
                    const int r = v.y - a.lv;
                    if (r >= 0 && r < a.R) {
                        const int k = v.x + kb;
                        const double bp = a.use_bias ? bias_cell(Eg + k, v.y) : 1.0;
                        nl += log(__dmul_rn(a.V[(size_t)r * a.W + k], bp) / cVB);
                        ul += log(__dmul_rn(bp, a.f[a.lv + r]) / cB);
                    }
                }
                nl = warp_sum(nl);
                ul = warp_sum(ul);
                static_assert(WPC == 1, "one warp per candidate: the warp's sums are the candidate's");
                if (lane == 0) {
                    const double lr = nl - ul;
                    const double margin = (double)(e1 - e0) * CS_EPS32 + 1e-6;
                    const bool sane = cVB > 1e-30 && cVB < 1e30;
                    if (!sane || !(lr <= a.min_lr - margin)) {
                        const int slot = atomicAdd(sa.confirm_count, 1);
                        sa.confirm[slot] = it;
                    } else
                        a.cand_lr[ci] = lr;
                }
            }
        }
    }
}

template <int CS_GROUP>
__global__ void __launch_bounds__(CS_THREADS, 2) k_cand_stats(CandArgs a)
{
    extern __shared__ __align__(16) double sm_cs[];
    __shared__ double red[32];
    __shared__ double s_sum[2][CS_GROUP];
    __shared__ double s_part[CS_GROUP][CS_THREADS / 32], s_ll[2][CS_THREADS / 32];
    __shared__ int s_need2;
    static_assert((CS_THREADS / 32) % CS_GROUP == 0, "warps must divide evenly over the candidates of a group");
    // E window of a candidate: element off0 + k is E[P - w + k]; J2 + 2 doubles of reach on both sides
    const int off0 = a.J2 + 2, nEw = off0 + a.W2 + a.J2 + 4;
    double *s_E = sm_cs;                      // [CS_GROUP][nEw]
    double *s_f = sm_cs + CS_GROUP * nEw;     // [R]              f_i over the VMat's sizes
    const int nwork = a.work_count[0];
    const int tid = threadIdx.x;
    const size_t tsz = (size_t)a.J2 * a.W2;
    for (int i = tid; i < a.R; i += CS_THREADS) s_f[i] = a.f[a.lv + i];
    for (int g0 = blockIdx.x * CS_GROUP; g0 < nwork; g0 += gridDim.x * CS_GROUP) {
        const int ng = min(CS_GROUP, nwork - g0);
        __syncthreads();
        if (tid == 0) s_need2 = 0;
        for (int cg = 0; cg < CS_GROUP; cg++) {
            if (cg < ng && a.use_bias) {
                const int2 it = a.work[g0 + cg];
                const int c = it.x;
                const int P = a.cand_pos[a.cand_off[c] + it.y];
                const int64_t e_lo = a.bias_off[c], e_hi = a.bias_off[c + 1];
                const int64_t eb = e_lo - (int64_t)(a.seq_start[c] + a.pwm_up) + (P - a.w - off0);
                for (int i = tid; i < nEw; i += CS_THREADS) {   // the few elements of reach past the window may lie off the track
                    const int64_t idx = eb + i;
                    s_E[cg * nEw + i] = (idx >= e_lo && idx < e_hi) ? a.E[idx] : 0.0;
                }
            } else
                for (int i = tid; i < nEw; i += CS_THREADS) s_E[cg * nEw + i] = a.use_bias ? 0.0 : 1.0;
        }
        __syncthreads();
        // ---- phase 1: S_VB = sum V*Bp (normaliser of the nucleosome model in the likelihood ratio)
        double sVB[CS_GROUP];
#pragma unroll
        for (int cg = 0; cg < CS_GROUP; cg++) sVB[cg] = 0.0;
        pair_window_sums<CS_GROUP>(a.pair, a.one, a.J2, a.W2, s_E, nEw, off0, sVB);
        // block totals of the NG sums: one pass of warp shuffles, one barrier
        const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
        for (int cg = 0; cg < CS_GROUP; cg++) {
            const double v = warp_sum(sVB[cg]);
            if (lane == 0) s_part[cg][wid] = v;
        }
        __syncthreads();
        // ---- sparse likelihoods over the fragments of each window, NucleosomeCalling.py:110-122: warp `wid` takes candidate
        // wid % NG (CS_THREADS / 32 / NG warps share a candidate's fragment list)
        {
            constexpr int NWARP = CS_THREADS / 32, WPC = NWARP / CS_GROUP;   // warps per candidate
            const int cg = wid % CS_GROUP, sub = wid / CS_GROUP;
            double nl = 0.0, ul = 0.0;
            if (cg < ng) {
                double cVB = 0.0;
#pragma unroll
                for (int w2 = 0; w2 < NWARP; w2++) cVB += s_part[cg][w2];
                const int2 it = a.work[g0 + cg];
                const int c = it.x;
                const int64_t ci = a.cand_off[c] + it.y;
                // S_B = sum f*Bp over the window = the bias coverage at the candidate (NucleosomeCalling.py:56-58)
                const double cB = a.cand_bcov[ci];
                const int x = a.cand_pos[ci] - a.start[c];
                const int32_t *cp = a.col_ptr + a.col_off[c];
                const int2 *en = a.ent + a.frag_off[c];
                const int e0 = cp[x - a.w + a.csc_pad], e1 = cp[x + a.w + 1 + a.csc_pad];
                const int kb = a.w - (x + a.csc_pad);
                for (int e = e0 + sub * 32 + lane; e < e1; e += 32 * WPC) {
                    const int2 v = en[e];
                    const int r = v.y - a.lv;
                    if (r >= 0 && r < a.R) {
                        const int k = v.x + kb;
                        const double bp = a.use_bias ? bias_cell(s_E + cg * nEw + off0 + k, v.y) : 1.0;
                        nl += log(__dmul_rn(a.V[(size_t)r * a.W + k], bp) / cVB);
                        ul += log(__dmul_rn(bp, s_f[r]) / cB);
                    }
                }
                if (sub == 0 && lane == 0) s_sum[1][cg] = cB;
            }
            nl = warp_sum(nl);
            ul = warp_sum(ul);
            if (lane == 0) {
                s_ll[0][wid] = nl;
                s_ll[1][wid] = ul;
            }
            __syncthreads();
            if (tid < ng) {
                double tn = 0.0, tu = 0.0;
#pragma unroll
                for (int q = 0; q < WPC; q++) {
                    tn += s_ll[0][tid + q * CS_GROUP];
                    tu += s_ll[1][tid + q * CS_GROUP];
                }
                const int2 it = a.work[g0 + tid];
                const int64_t ci = a.cand_off[it.x] + it.y;
                const double lr = a.lr_is_nan ? nb_nan() : tn - tu;  // 0*log(0) cells make both likelihoods NaN in the reference
                a.cand_lr[ci] = lr;
                if (lr > a.min_lr) {
                    a.cand_flag[ci] |= 2;
                    atomicOr(&s_need2, 1 << tid);
                }
            }
        }
        __syncthreads();
        const int need2 = s_need2;
        if (!need2) continue;
        // ---- phase 2 (only where lr > min_lr): S_BV = sum f*V*Bp, S_BV2 = sum f*V^2*Bp  -> variance, z
        for (int cg = 0; cg < ng; cg++) {
            if (!(need2 & (1 << cg))) continue;
            double sBV = 0.0, sBV2 = 0.0;
            pair_window_sums<1>(a.pair + tsz, a.one + a.W2, a.J2, a.W2, s_E + cg * nEw, nEw, off0, &sBV);
            pair_window_sums<1>(a.pair + 2 * tsz, a.one + 2 * a.W2, a.J2, a.W2, s_E + cg * nEw, nEw, off0, &sBV2);
            const double x1 = block_sum(sBV, red), x2 = block_sum(sBV2, red);
            if (tid == 0) {
                const int2 it = a.work[g0 + cg];
                const int64_t ci = a.cand_off[it.x] + it.y;
                const double cB = s_sum[1][cg];
                const double mean = x1 / cB;
                // calculateCov closed form r*(sum p v^2 - (sum p v)^2), r truncated to C int (multinomial_cov.pyx:20)
                const double var = (double)(int)a.cand_cov[ci] * (x2 / cB - mean * mean);
                const double z = a.cand_norm[ci] / sqrt(var);
                a.cand_z[ci] = z;
                if (z >= a.min_z) a.cand_flag[ci] |= 4;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
struct NucReduceArgs {
    const int64_t *cand_off;
    const int32_t *cand_count, *cand_pos;
    int32_t *cand_flag;
    const double *cand_z;
    int32_t *sc_pos;   // scratch sized like the candidate arrays
    double *sc_val;
    unsigned char *sc_state;
    int32_t *sc_idx;
    int sep;
};

__global__ void __launch_bounds__(256) k_nuc_reduce(NucReduceArgs a)
{
    __shared__ int red_i[32];
    __shared__ int s_base, s_flag;
    const int c = blockIdx.x;
    const int64_t po = a.cand_off[c];
    int n = a.cand_count[c];
    if (n < 0) n = (int)(a.cand_off[c + 1] - po);
    const int tid = threadIdx.x;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int j0 = 0; j0 < n; j0 += blockDim.x) {
        const int j = j0 + tid;
        const int flag = (j < n) && (a.cand_flag[po + j] & 4);
        int slot = block_compact_slot(flag, &s_base, red_i);
        if (flag) {
            a.sc_pos[po + slot] = a.cand_pos[po + j];
            a.sc_val[po + slot] = a.cand_z[po + j];
            a.sc_idx[po + slot] = j;
        }
    }
    __syncthreads();
    const int m = s_base;
    block_nms(a.sc_pos + po, a.sc_val + po, a.sc_state + po, m, a.sep, &s_flag);
    __syncthreads();
    for (int j = tid; j < m; j += blockDim.x)
        if (a.sc_state[po + j] == 1) a.cand_flag[po + a.sc_idx[po + j]] |= 8;
}

// ---------------------------------------------------------------------------------------------
extern "C" {

int nb200_nuc_run(nb200_ctx *ctx, nb200_dbatch *b)
{
    if (!ctx || !b) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_nuc_run: NULL argument");
    if (!ctx->nuc_configured) return nb200_fail(ctx, NB200_ERR_STATE, "nb200_nuc_configure has not been called");
    RunConst &r = ctx->rc;
    const nb200_nuc_params &p = ctx->nuc;
    double bx_nobias = 0.0;   // the background xcor without --fasta (a constant), set where the tracks are launched
    if (!r.have_vmat) return nb200_fail(ctx, NB200_ERR_STATE, "nb200_set_vmat has not been called");
    if (!r.have_sizes || r.sizes_upper < r.v_upper)
        return nb200_fail(ctx, NB200_ERR_STATE, "nb200_set_fragment_sizes missing or shorter than the VMat's upper size");
    if (r.v_upper > NB200_MAX_UPPER) return nb200_fail(ctx, NB200_ERR_ARG, "VMat upper > %d unsupported", NB200_MAX_UPPER);
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CUDA(ctx, cudaStreamWaitEvent(b->stream, b->ev_copied_nuc, 0));   // a download of the previous pass may still read the arrays
    NB_CHECK(nb200_fifo_enter(ctx, b->stream));                          // passes run first-in first-out across batches
    const int n = b->n_chunks;
    const int W = r.v_cols, w = r.v_w, lv = r.v_lower, uv = r.v_upper;
    if (b->min_len < p.smooth_len)
        return nb200_fail(ctx, NB200_ERR_ARG, "a chunk is shorter (%d) than the smoothing window (%d)", b->min_len, p.smooth_len);
    if (r.n_jitter < b->max_len)
        return nb200_fail(ctx, NB200_ERR_STATE, "nb200_set_jitter: need >= %d values (longest chunk), have %lld", b->max_len,
                          (long long)r.n_jitter);
    const int pad = W > uv / 2 + 1 ? W : uv / 2 + 1;  // NucleosomeCalling.py:240
    NB_CHECK(nb200_prep_csc(ctx, b, pad, uv, p.atac, lv));
    if (p.use_bias) {
        NB_CHECK(nb200_prep_bias(ctx, b));
        for (int c = 0; c < n; c++) {
            int64_t b0 = (int64_t)b->h_seq_start[c] + r.pwm_up;
            int64_t b1 = b0 + (b->h_seq_off[c + 1] - b->h_seq_off[c]) - (r.pwm_width - 1);
            if (b0 > (int64_t)b->h_start[c] - w - uv / 2 || b1 < (int64_t)b->h_end[c] + w + uv / 2 + 1)
                return nb200_fail(ctx, NB200_ERR_FLANK,
                                  "Insufficient flanking region: chunk %d needs sequence over [%lld, %lld)", c,
                                  (long long)b->h_start[c] - w - uv / 2 - r.pwm_up,
                                  (long long)b->h_end[c] + w + uv / 2 + 1 + r.pwm_down);
        }
    }
    const size_t tl = (size_t)b->total_len;
    DevBuf *tracks[] = {&b->n_signal, &b->n_bg, &b->n_norm, &b->n_smooth, &b->n_nuc_cov, &b->n_nfr_cov, &b->n_bcov, &b->n_comb, &b->sc_f64};
    for (auto t : tracks) NB_CUDA(ctx, t->reserve(sizeof(double) * tl));
    NB_CUDA(ctx, b->sc_i32.reserve(sizeof(int32_t) * tl));
    NB_CUDA(ctx, b->sc_u8.reserve(tl));
    // candidate capacities: kept candidates are >= redundant_sep apart
    b->h_ncand_off.assign(n + 1, 0);
    for (int c = 0; c < n; c++) b->h_ncand_off[c + 1] = b->h_ncand_off[c] + (b->h_end[c] - b->h_start[c]) / p.redundant_sep + 2;
    const size_t nc = (size_t)b->h_ncand_off[n];
    NB_CUDA(ctx, b->n_cand_off.reserve(sizeof(int64_t) * (n + 1)));
    if (b->nuc_done) NB_CUDA(ctx, cudaStreamSynchronize(b->stream));  // re-run: the staging slot may still be in flight
    memcpy(b->pin_slot(3), b->h_ncand_off.data(), sizeof(int64_t) * (n + 1));
    NB_CUDA(ctx, cudaMemcpyAsync(b->n_cand_off.p, b->pin_slot(3), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, b->stream));
    NB_CUDA(ctx, b->n_cand_count.reserve(sizeof(int32_t) * n));
    NB_CUDA(ctx, b->n_cand_pos.reserve(sizeof(int32_t) * nc));
    NB_CUDA(ctx, b->n_cand_flag.reserve(sizeof(int32_t) * nc));
    DevBuf *cd[] = {&b->n_cand_z, &b->n_cand_lr, &b->n_cand_norm, &b->n_cand_sig, &b->n_cand_cov, &b->n_cand_nfr, &b->n_cand_smooth, &b->n_cand_bcov};
    for (auto t : cd) NB_CUDA(ctx, t->reserve(sizeof(double) * nc));
    NB_CUDA(ctx, b->n_work.reserve(sizeof(int2) * 3 * nc));   // [0, nc): candidates with reads, [nc, 2 nc): past the bound, [2 nc, 3 nc): past the fp32 screen
    NB_CUDA(ctx, b->n_work_count.reserve(sizeof(int32_t) * 4));
    NB_CUDA(ctx, cudaMemsetAsync(b->n_work_count.p, 0, sizeof(int32_t) * 4, b->stream));

    if (p.use_bias) {
        NB_CUDA(ctx, b->n_bx.reserve(sizeof(double) * tl));
        NB_CUDA(ctx, b->n_cB.reserve(sizeof(double) * (tl + 2 * (size_t)w * n)));
        // One pass over the columns serves nucleosome calling (cB, weights f over [lv, uv), pad w) and, when the occupancy
        // path is configured with bias too, its two column sums (cn / cf, weights pn / pf over [0, upper), pad flank): the
        // tap walk and the E loads are shared, nb200_occ_run of this batch then skips its own pass.
        const nb200_occ_params &po = ctx->occ;
        const bool merge_occ = ctx->occ_configured && po.use_bias && r.have_occ_model && po.upper <= r.occ_upper &&
                               b->occ_cols_gen != ctx->occ_gen && !getenv("NB200_NO_MERGED_COLSUMS");
        if (merge_occ) {
            const size_t ncs = tl + 2 * (size_t)po.flank * n;
            NB_CUDA(ctx, b->o_cn.reserve(sizeof(double) * ncs));
            NB_CUDA(ctx, b->o_cf.reserve(sizeof(double) * ncs));
            PairColsumArgs<3> pa;
            pa.start = b->d_start.as<int32_t>();
            pa.out_off = b->d_out_off.as<int64_t>();
            pa.bias_off = b->d_bias_off.as<int64_t>();
            pa.seq_start = b->d_seq_start.as<int32_t>();
            pa.E = b->d_E.as<double>();
            pa.pwm_up = r.pwm_up;
            pa.wt[0] = r.sizes.as<double>();
            pa.out[0] = b->n_cB.as<double>();
            pa.lo[0] = lv;
            pa.hi[0] = uv;
            pa.pad[0] = w;
            pa.wt[1] = r.nuc_probs.as<double>();
            pa.wt[2] = r.nfr_probs.as<double>();
            pa.out[1] = b->o_cn.as<double>();
            pa.out[2] = b->o_cf.as<double>();
            for (int t = 1; t < 3; t++) {
                pa.lo[t] = 0;
                pa.hi[t] = po.upper;
                pa.pad[t] = po.flank;
            }
            const int hmax = std::max(uv, (int)po.upper), pmax = std::max(w, (int)po.flank);
            const size_t smem = pair_colsums_smem<3>(hmax);
            if (smem > 48 * 1024) NB_CUDA(ctx, cudaFuncSetAttribute(k_pair_colsums<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ProfScope ps(ctx, b->stream, "k_colsums_merged");
            dim3 grid((unsigned)div_up64(b->max_len + 2 * pmax, 2 * PC_THREADS), n);
            k_pair_colsums<3><<<grid, PC_THREADS, smem, b->stream>>>(pa);
            NB_LAUNCH_CHECK(ctx);
            b->occ_cols_gen = ctx->occ_gen;
        } else {
            PairColsumArgs<1> pa;
            pa.start = b->d_start.as<int32_t>();
            pa.out_off = b->d_out_off.as<int64_t>();
            pa.bias_off = b->d_bias_off.as<int64_t>();
            pa.seq_start = b->d_seq_start.as<int32_t>();
            pa.E = b->d_E.as<double>();
            pa.wt[0] = r.sizes.as<double>();
            pa.out[0] = b->n_cB.as<double>();
            pa.pwm_up = r.pwm_up;
            pa.lo[0] = lv;
            pa.hi[0] = uv;
            pa.pad[0] = w;
            const size_t smem = pair_colsums_smem<1>(uv);
            if (smem > 48 * 1024) NB_CUDA(ctx, cudaFuncSetAttribute(k_pair_colsums<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ProfScope ps(ctx, b->stream, "k_nuc_colsums");
            dim3 grid((unsigned)div_up64(b->max_len + 2 * w, 2 * PC_THREADS), n);
            k_pair_colsums<1><<<grid, PC_THREADS, smem, b->stream>>>(pa);
            NB_LAUNCH_CHECK(ctx);
        }
        int mode = p.xcor_mode;
        if (mode == 0) mode = nb200_tc_available(ctx) ? 2 : 1;  // auto: tcgen05 contraction when its plan exists
        if (mode == 2) {
            if (!nb200_tc_available(ctx)) return nb200_fail(ctx, NB200_ERR_STATE, "xcor_mode 2 (tcgen05) is not available in this build");
            NB_CHECK(nb200_nuc_bx_tc(ctx, b));
        } else
            NB_CHECK(nb200_nuc_bx_fp64(ctx, b));
    }
    {
        NucTrackArgs a;
        a.out_off = b->d_out_off.as<int64_t>();
        a.col_off = b->d_col_off.as<int64_t>();
        a.frag_off = b->d_frag_off.as<int64_t>();
        a.col_ptr = b->d_col_ptr.as<int32_t>();
        a.col_low = (lv > 0) ? b->d_col_low.as<int32_t>() : nullptr;
        a.ent = b->d_ent.as<int2>();
        a.V = r.vmat.as<double>();
        a.cB = b->n_cB.as<double>();
        a.bx = b->n_bx.as<double>();
        a.nuc_cov = b->n_nuc_cov.as<double>();
        a.nfr_cov = b->n_nfr_cov.as<double>();
        a.bcov = b->n_bcov.as<double>();
        a.signal = b->n_signal.as<double>();
        a.bg = b->n_bg.as<double>();
        a.norm = b->n_norm.as<double>();
        a.lv = lv;
        a.uv = uv;
        a.W = W;
        a.w = w;
        a.csc_pad = b->csc_pad;
        a.use_bias = p.use_bias;
        // no --fasta: bias matrix = ones * f_i (NucleosomeCalling.py:248-254) -> both xcors are constants
        a.bcov_nobias = r.f_sum_v * W;
        double s = 0.0;
        for (int i = 0; i < r.v_rows; i++) {
            double rs = 0.0;
            for (int k = 0; k < W; k++) rs += r.h_vmat[(size_t)i * W + k];
            s += rs * r.h_sizes[lv + i];
        }
        a.bx_nobias = s;
        bx_nobias = s;
        const size_t ncb = (NT_TILE + 2 * (size_t)w + 8) & ~(size_t)7;
        size_t smem = sizeof(double) * (ncb + ncb / 8 + NT_FRAG_CAP);
        if (smem > 48 * 1024) NB_CUDA(ctx, cudaFuncSetAttribute(k_nuc_tracks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ProfScope ps(ctx, b->stream, "k_nuc_tracks");
        dim3 grid((unsigned)div_up64(b->max_len, NT_TILE), n);
        k_nuc_tracks<<<grid, NT_TILE, smem, b->stream>>>(a);
        NB_LAUNCH_CHECK(ctx);
    }
    {
        SmoothTracks tr;
        tr.in[0] = tr.in[1] = tr.in[2] = b->n_norm.as<double>();
        tr.out[0] = tr.out[1] = tr.out[2] = b->n_smooth.as<double>();
        size_t smem = smooth_same_smem(p.smooth_len);
        if (smem > 48 * 1024) NB_CUDA(ctx, cudaFuncSetAttribute(k_smooth_same, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ProfScope ps(ctx, b->stream, "k_smooth_same");
        dim3 grid((unsigned)div_up64(b->max_len, SM_TILE), n, 1);
        k_smooth_same<<<grid, SM_THREADS, smem, b->stream>>>(tr, b->d_out_off.as<int64_t>(), r.nuc_win.as<double>(), p.smooth_len, 1);
        NB_LAUNCH_CHECK(ctx);
    }
    {
        NB_CUDA(ctx, b->n_bx.reserve(sizeof(double) * tl));  // reused as the combined-signal scratch when bias is off
        NucPeakArgs a;
        a.start = b->d_start.as<int32_t>();
        a.out_off = b->d_out_off.as<int64_t>();
        a.cand_off = b->n_cand_off.as<int64_t>();
        a.jitter = r.jitter.as<double>();
        a.norm = b->n_norm.as<double>();
        a.smooth = b->n_smooth.as<double>();
        a.signal = b->n_signal.as<double>();
        a.nuc_cov = b->n_nuc_cov.as<double>();
        a.nfr_cov = b->n_nfr_cov.as<double>();
        a.comb = b->n_comb.as<double>();
        a.bcov = b->n_bcov.as<double>();
        a.cand_bcov = b->n_cand_bcov.as<double>();
        a.sc_pos = b->sc_i32.as<int32_t>();
        a.sc_val = b->sc_f64.as<double>();
        a.sc_state = b->sc_u8.as<unsigned char>();
        a.cand_count = b->n_cand_count.as<int32_t>();
        a.cand_pos = b->n_cand_pos.as<int32_t>();
        a.cand_flag = b->n_cand_flag.as<int32_t>();
        a.cand_z = b->n_cand_z.as<double>();
        a.cand_lr = b->n_cand_lr.as<double>();
        a.cand_norm = b->n_cand_norm.as<double>();
        a.cand_sig = b->n_cand_sig.as<double>();
        a.cand_cov = b->n_cand_cov.as<double>();
        a.cand_nfr = b->n_cand_nfr.as<double>();
        a.cand_smooth = b->n_cand_smooth.as<double>();
        a.work = b->n_work.as<int2>();
        a.work_count = b->n_work_count.as<int32_t>();
        a.sep = p.redundant_sep;
        a.boundary = p.nonredundant_sep / 2;
        a.order = p.redundant_sep / 2;
        a.min_signal = 0.0;
        a.min_reads = p.min_reads;
        ProfScope ps(ctx, b->stream, "k_nuc_peaks");
        k_nuc_peaks<<<n, PK_THREADS_N, 0, b->stream>>>(a);
        NB_LAUNCH_CHECK(ctx);
    }
    {
        CandArgs a;
        a.start = b->d_start.as<int32_t>();
        a.out_off = b->d_out_off.as<int64_t>();
        a.col_off = b->d_col_off.as<int64_t>();
        a.frag_off = b->d_frag_off.as<int64_t>();
        a.bias_off = b->d_bias_off.as<int64_t>();
        a.cand_off = b->n_cand_off.as<int64_t>();
        a.seq_start = b->d_seq_start.as<int32_t>();
        a.col_ptr = b->d_col_ptr.as<int32_t>();
        a.ent = b->d_ent.as<int2>();
        a.E = b->d_E.as<double>();
        a.V = r.vmat.as<double>();
        a.f = r.sizes.as<double>();
        a.pair = r.vp_pair.as<double2>();
        a.one = r.vp_one.as<double>();
        a.J2 = r.vp_J2;
        a.W2 = r.vp_W2;
        a.work = b->n_work.as<int2>();
        a.work_count = b->n_work_count.as<int32_t>();
        a.cand_pos = b->n_cand_pos.as<int32_t>();
        a.cand_flag = b->n_cand_flag.as<int32_t>();
        a.cand_norm = b->n_cand_norm.as<double>();
        a.cand_cov = b->n_cand_cov.as<double>();
        a.cand_bcov = b->n_cand_bcov.as<double>();
        a.cand_z = b->n_cand_z.as<double>();
        a.cand_lr = b->n_cand_lr.as<double>();
        a.pwm_up = r.pwm_up;
        a.lv = lv;
        a.R = r.v_rows;
        a.W = W;
        a.w = w;
        a.csc_pad = b->csc_pad;
        a.use_bias = p.use_bias;
        a.lr_is_nan = (r.v_has_zero || r.f_has_zero) ? 1 : 0;
        a.min_lr = p.min_lr;
        a.min_z = p.min_z;
        // cascade: cheap upper bound of LR -> fp32 screen -> exact fp64 statistics (NB200_CS_SCREEN=0: exact for all)
        const int cs_screen = getenv("NB200_CS_SCREEN") ? atoi(getenv("NB200_CS_SCREEN")) : 1;
        int cs_group_exact = 0;
        if (cs_screen) {
            const int64_t ncap = b->h_ncand_off.back();
            int32_t *counts = b->n_work_count.as<int32_t>();
            BoundArgs ba;
            ba.c = a;
            ba.bx = p.use_bias ? b->n_bx.as<double>() : nullptr;
            ba.bx_nobias = bx_nobias;
            ba.f_max = r.f_max_v;
            ba.bound_ok = (r.v_nonneg && r.f_max_v > 0.0 && !getenv("NB200_CS_NOBOUND")) ? 1 : 0;
            ba.next = b->n_work.as<int2>() + ncap;
            ba.next_count = counts + 1;
            {
                ProfScope ps(ctx, b->stream, "k_cand_bound");
                k_cand_bound<<<ctx->sm_count * 8, 256, 0, b->stream>>>(ba);
                NB_LAUNCH_CHECK(ctx);
            }
            ScreenArgs sa;
            sa.c = a;
            sa.c.work = ba.next;
            sa.c.work_count = ba.next_count;
            sa.pair32 = r.vp_pair32.as<float2>();
            sa.one32 = r.vp_one32.as<float>();
            sa.confirm = b->n_work.as<int2>() + 2 * ncap;
            sa.confirm_count = counts + 2;
            const size_t smem32 = sizeof(float) * (CS_SCREEN_GROUP * ((size_t)r.vp_W2 + 2 * r.vp_J2 + 6) + 8);
            if (smem32 > 48 * 1024)
                NB_CUDA(ctx, cudaFuncSetAttribute(k_cand_screen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32));
            {
                ProfScope ps(ctx, b->stream, "k_cand_screen");
                k_cand_screen<<<ctx->sm_count * 4, CS_THREADS, smem32, b->stream>>>(sa);
                NB_LAUNCH_CHECK(ctx);
            }
            a.work = sa.confirm;
            a.work_count = sa.confirm_count;
            cs_group_exact = 4;
        }
        ProfScope ps(ctx, b->stream, "k_cand_stats");
        static const int cs_group = getenv("NB200_CS_GROUP") ? atoi(getenv("NB200_CS_GROUP")) : CS_GROUP_DEFAULT;
        const int G = cs_group_exact ? cs_group_exact : (cs_group <= 4 ? 4 : 8);
        size_t smem = sizeof(double) * (G * ((size_t)r.vp_W2 + 2 * r.vp_J2 + 6) + r.v_rows);
        auto kern = G == 1 ? k_cand_stats<1> : (G == 4 ? k_cand_stats<4> : k_cand_stats<8>);
        if (smem > 48 * 1024) NB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<ctx->sm_count * 4, CS_THREADS, smem, b->stream>>>(a);
        NB_LAUNCH_CHECK(ctx);
    }
    {
        NucReduceArgs a;
        a.cand_off = b->n_cand_off.as<int64_t>();
        a.cand_count = b->n_cand_count.as<int32_t>();
        a.cand_pos = b->n_cand_pos.as<int32_t>();
        a.cand_flag = b->n_cand_flag.as<int32_t>();
        a.cand_z = b->n_cand_z.as<double>();
        // the per-position scratch arrays are larger than the candidate arrays (cand capacity <= len/sep+2 <= len)
        a.sc_pos = b->sc_i32.as<int32_t>();
        a.sc_val = b->sc_f64.as<double>();
        a.sc_state = b->sc_u8.as<unsigned char>();
        a.sc_idx = reinterpret_cast<int32_t *>(b->n_comb.p);
        a.sep = p.nonredundant_sep;
        ProfScope ps(ctx, b->stream, "k_nuc_reduce");
        k_nuc_reduce<<<n, 256, 0, b->stream>>>(a);
        NB_LAUNCH_CHECK(ctx);
    }
    NB_CHECK(nb200_fifo_leave(ctx, b->stream));
    b->nuc_done = true;
    return NB200_OK;
}

static int d2h(nb200_ctx *ctx, nb200_dbatch *b, void *dst, const DevBuf &src, size_t bytes)
{
    if (!dst || !bytes) return NB200_OK;
    NB_CUDA(ctx, cudaMemcpyAsync(dst, src.p, bytes, cudaMemcpyDeviceToHost, b->copy_stream));
    return NB200_OK;
}

}  // extern "C"

template <typename T, typename Out>
static int nuc_download_impl(nb200_ctx *ctx, nb200_dbatch *b, const Out *o, const char *who)
{
    if (!ctx || !b || !o) return nb200_fail(ctx, NB200_ERR_ARG, "%s: NULL argument", who);
    if (!b->nuc_done) return nb200_fail(ctx, NB200_ERR_STATE, "%s: nb200_nuc_run has not been called on this batch", who);
    const size_t tl = (size_t)b->total_len, tb = sizeof(T) * tl;
    const int n = b->n_chunks;
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CUDA(ctx, cudaEventRecord(b->ev_pass, b->stream));            // copies start once the pass has finished ...
    NB_CUDA(ctx, cudaStreamWaitEvent(b->copy_stream, b->ev_pass, 0));
    struct Copied {                                                  // ... and the next pass over these arrays waits for them
        nb200_dbatch *b;
        ~Copied() { cudaEventRecord(b->ev_copied_nuc, b->copy_stream); }
    } copied{b};
    T *host[6] = {o->nuc_signal, o->background, o->norm_signal, o->smoothed, o->nuc_cov, o->nfr_cov};
    const DevBuf *dev[6] = {&b->n_signal, &b->n_bg, &b->n_norm, &b->n_smooth, &b->n_nuc_cov, &b->n_nfr_cov};
    if (sizeof(T) == sizeof(double)) {
        for (int t = 0; t < 6; t++) NB_CHECK(d2h(ctx, b, host[t], *dev[t], tb));
    } else {   // float32 tracks: converted on the device (k_pack_f32 on the copy stream, after the pass), then copied
        Pack32Args pa;
        int nt = 0, which[6];
        for (int t = 0; t < 6; t++)
            if (host[t]) which[nt++] = t;
        if (nt && tl) {
            const size_t slab = (tl + 3) & ~(size_t)3;
            NB_CUDA(ctx, b->pack32_nuc.reserve(sizeof(float) * slab * nt));
            for (int k = 0; k < nt; k++) {
                pa.src[k] = dev[which[k]]->template as<double>();
                pa.dst[k] = b->pack32_nuc.as<float>() + slab * k;
            }
            pa.n = (int64_t)tl;
            ProfScope ps(ctx, b->copy_stream, "k_pack_f32");
            k_pack_f32<<<dim3((unsigned)std::min<int64_t>(div_up64((int64_t)tl, 4 * 256), ctx->sm_count * 8), nt), 256, 0, b->copy_stream>>>(pa);
            NB_LAUNCH_CHECK(ctx);
        }
        for (int k = 0; k < nt; k++)
            NB_CUDA(ctx, cudaMemcpyAsync(host[which[k]], pa.dst[k], tb, cudaMemcpyDeviceToHost, b->copy_stream));
    }
    NB_CHECK(d2h(ctx, b, o->cand_count, b->n_cand_count, sizeof(int32_t) * n));
    if (o->cand_pos) {
        if (!o->cand_off) return nb200_fail(ctx, NB200_ERR_ARG, "%s: cand_off is required with cand_pos", who);
        for (int c = 0; c <= n; c++)
            if (o->cand_off[c] != b->h_ncand_off[c])
                return nb200_fail(ctx, NB200_ERR_CAPACITY, "%s: cand_off must equal len/redundant_sep+2 capacities (chunk %d)", who, c);
        const size_t nc = (size_t)b->h_ncand_off[n];
        NB_CHECK(d2h(ctx, b, o->cand_pos, b->n_cand_pos, sizeof(int32_t) * nc));
        NB_CHECK(d2h(ctx, b, o->cand_flag, b->n_cand_flag, sizeof(int32_t) * nc));
        NB_CHECK(d2h(ctx, b, o->cand_z, b->n_cand_z, sizeof(double) * nc));
        NB_CHECK(d2h(ctx, b, o->cand_lr, b->n_cand_lr, sizeof(double) * nc));
        NB_CHECK(d2h(ctx, b, o->cand_norm_signal, b->n_cand_norm, sizeof(double) * nc));
        NB_CHECK(d2h(ctx, b, o->cand_nuc_signal, b->n_cand_sig, sizeof(double) * nc));
        NB_CHECK(d2h(ctx, b, o->cand_nuc_cov, b->n_cand_cov, sizeof(double) * nc));
        NB_CHECK(d2h(ctx, b, o->cand_nfr_cov, b->n_cand_nfr, sizeof(double) * nc));
        NB_CHECK(d2h(ctx, b, o->cand_smoothed, b->n_cand_smooth, sizeof(double) * nc));
    }
    return NB200_OK;
}

template <typename Out>
static int64_t nuc_d2h_bytes_impl(nb200_dbatch *b, const Out *o, int64_t elem)
{
    if (!b || !o) return 0;
    const int64_t tb = elem * b->total_len;
    int64_t s = 0;
    const void *tr[] = {o->nuc_signal, o->background, o->norm_signal, o->smoothed, o->nuc_cov, o->nfr_cov};
    for (auto p : tr)
        if (p) s += tb;
    if (o->cand_count) s += 4LL * b->n_chunks;
    if (o->cand_pos && !b->h_ncand_off.empty()) {
        int64_t nc = b->h_ncand_off[b->n_chunks];
        s += 4 * nc;
        if (o->cand_flag) s += 4 * nc;
        const void *cd[] = {o->cand_z, o->cand_lr, o->cand_norm_signal, o->cand_nuc_signal, o->cand_nuc_cov, o->cand_nfr_cov, o->cand_smoothed};
        for (auto p : cd)
            if (p) s += 8 * nc;
    }
    return s;
}

extern "C" {

int nb200_nuc_download(nb200_ctx *ctx, nb200_dbatch *b, const nb200_nuc_out *o)
{
    return nuc_download_impl<double>(ctx, b, o, "nb200_nuc_download");
}
int nb200_nuc_download32(nb200_ctx *ctx, nb200_dbatch *b, const nb200_nuc_out32 *o)
{
    return nuc_download_impl<float>(ctx, b, o, "nb200_nuc_download32");
}
int64_t nb200_nuc_d2h_bytes(nb200_dbatch *b, const nb200_nuc_out *o) { return nuc_d2h_bytes_impl(b, o, 8); }
int64_t nb200_nuc_d2h_bytes32(nb200_dbatch *b, const nb200_nuc_out32 *o) { return nuc_d2h_bytes_impl(b, o, 4); }

}  // extern "C"
