// nb200_occ.cu -- OccChunk.process on the device (nucleoatac/Occupancy.py:195-253, run_occ.py:23-39).
//
// Stages (per batch, every chunk in parallel):
//   k_pair_colsums per-column sums  cn[c] = sum_i pn[i]*Bp[i,c],  cf[c] = sum_i pf[i]*Bp[i,c]   (bias only; nb200_dev.cuh,
//                  profiled as "k_occ_colsums")
//   k_occ_mle      one window per 8 lanes: the window's fragments from the CSC fragment matrix, 101-point alpha
//                  log-likelihood grid in fp64 (as the log of a product), first-max argmax and the
//                  likelihood-ratio confidence bounds                                   (Occupancy.py:104-146)
//   k_smooth_same  NaN-aware gaussian smoothing of the three tracks                     (Occupancy.py:147-153)
//   k_occ_peaks    coverage, call_peaks + OccPeak filter + getNucDist, one block per chunk (Occupancy.py:221-240)
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cmath>

#include "nb200_dev.cuh"

// ---------------------------------------------------------------------------------------------
struct OccMleArgs {
    const int32_t *start;
    const int64_t *out_off, *col_off, *frag_off, *bias_off;
    const int32_t *seq_start;
    const int32_t *col_ptr;
    const int2 *ent;
    const double *E, *cn, *cf, *pn, *pf, *alphas;
    const double *wsn, *wsf;   // per window: sums of cn / cf over its 2*flank+1 columns (k_occ_winsums); window k of chunk c at out_off[c]/step + c + k
    double *vals, *lower, *upper_b;
    int pwm_up, upper, flank, step, halfstep, csc_pad, n_alpha, use_bias;
    int pn_has_zero, pf_has_zero, both_zero;
    double cutoff, sn_nobias, sf_nobias;
    double thr_m;   // exp(-cutoff/2) = thr_m * 2^thr_e, thr_m in [1,2) (NaN for a NaN cutoff); thr_zero: it underflows to 0
    int thr_e, thr_zero;
};

// Window sums of the per-column sums: SN[k] = sum of cn over the 2*flank+1 columns of window k (t = halfstep + k*step), SF
// likewise from cf -- the bias model's normalisers of Occupancy.py:106-109.  Block = WS_WIN consecutive windows of a chunk,
// their columns staged in shared memory; a lane's columns are `step` doubles apart (conflict free for odd steps).
#ifndef WS_WIN
#define WS_WIN 128
#endif
__global__ void __launch_bounds__(WS_WIN) k_occ_winsums(const int64_t *__restrict__ out_off, const double *__restrict__ cn,
                                                        const double *__restrict__ cf, int flank, int step, int halfstep,
                                                        double *__restrict__ wsn, double *__restrict__ wsf)
{
    extern __shared__ double sm_ws[];
    const int c = blockIdx.y;
    const int64_t oo = out_off[c];
    const int L = (int)(out_off[c + 1] - oo);
    const int nwin = (L - halfstep + step - 1) / step;
    const int k0 = blockIdx.x * WS_WIN;
    if (k0 >= nwin) return;
    const int window = 2 * flank + 1;
    const int nk = min(WS_WIN, nwin - k0);
    const int ncol = (nk - 1) * step + window;
    double *s_n = sm_ws, *s_f = sm_ws + (WS_WIN - 1) * step + window;
    const int64_t co = oo + 2 * (int64_t)flank * c + halfstep + (int64_t)k0 * step;  // colsum index of the first column of window k0
    for (int i = threadIdx.x; i < ncol; i += WS_WIN) {
        s_n[i] = cn[co + i];
        s_f[i] = cf[co + i];
    }
    __syncthreads();
    if ((int)threadIdx.x >= nk) return;
    const double *pn = s_n + threadIdx.x * step, *pf = s_f + threadIdx.x * step;
    double n4[4] = {0.0, 0.0, 0.0, 0.0}, f4[4] = {0.0, 0.0, 0.0, 0.0};
    int k = 0;
    for (; k + 4 <= window; k += 4) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            n4[u] += pn[k + u];
            f4[u] += pf[k + u];
        }
    }
    for (; k < window; k++) {
        n4[0] += pn[k];
        f4[0] += pf[k];
    }
    const int64_t wo = oo / step + c + k0 + threadIdx.x;
    wsn[wo] = (n4[0] + n4[1]) + (n4[2] + n4[3]);
    wsf[wo] = (f4[0] + f4[1]) + (f4[2] + f4[3]);
}

// One window per 8-lane group, 4 windows per warp (Occupancy.py:104-146).  The bias of a fragment's insert size over the
// window multiplies both mixture components (nuc[s] = pn[s]*bias[s]/SN, nfr[s] = pf[s]*bias[s]/SF, Occupancy.py:106-109),
// so it is a factor of the likelihood that does not depend on alpha: it drops out of the argmax and of the likelihood
// ratios 2*(max - ll).  What a window needs from the bias model is only SN and SF (sums of the per-column sums cn, cf).
// ll[a] = sum_f log(v_f(a)) is evaluated as log(prod_f v_f) with every factor pre-scaled by a power of two (again constant
// in alpha) and the running product renormalised every 32 factors: one log per alpha instead of one per (alpha, fragment).
#define MLE_WARPS 4
#define MLE_GROUPS 4
#ifndef MLE_ITERS
#define MLE_ITERS 8
#endif
template <int NQ, int LB>  // NQ alphas per lane: lane r of a group owns alphas r, r + 8, ...; LB resident blocks per SM
__global__ void __launch_bounds__(MLE_WARPS * 32, LB) k_occ_mle(OccMleArgs a)
{
    extern __shared__ double sm_mle[];  // pn[upper], pf[upper]
    __shared__ double s_d[MLE_WARPS][32], s_q[MLE_WARPS][32], s_p[MLE_WARPS][32];
    double *s_pn = sm_mle, *s_pf = sm_mle + a.upper;
    for (int i = threadIdx.x; i < a.upper; i += blockDim.x) {
        s_pn[i] = a.pn[i];
        s_pf[i] = a.pf[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3, r = lane & 7;
    const int c = blockIdx.y;
    const int64_t oo = a.out_off[c];
    const int L = (int)(a.out_off[c + 1] - oo);
    const int nwin = (L - a.halfstep + a.step - 1) / a.step;  // windows at t = halfstep + k*step < L
    const int32_t *cp = a.col_ptr + a.col_off[c];
    const int2 *en = a.ent + a.frag_off[c];
    // grid constants of this lane: its alphas and which of them are dead (0 * log 0 = NaN -> -inf, Occupancy.py:112-114)
    double al[NQ];
    unsigned dead = 0;
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const int ai = r + 8 * q;
        al[q] = (ai < a.n_alpha) ? a.alphas[ai] : 0.5;
        if ((ai >= a.n_alpha) || a.both_zero || (al[q] == 0.0 && a.pf_has_zero) || (al[q] == 1.0 && a.pn_has_zero)) dead |= 1u << q;
    }
    const bool last_is_one = (al[NQ - 1] == 1.0);  // alpha == 1 (only ever the last grid value, checked on the host)
  for (int it = 0; it < MLE_ITERS; it++) {   // a warp scores MLE_ITERS groups of 4 windows
    const int wbase = ((blockIdx.x * MLE_ITERS + it) * MLE_WARPS + warp) * MLE_GROUPS;
    if (wbase >= nwin) break;  // whole warp past the last window
    const int wi = wbase + g;
    const bool valid = wi < nwin;
    const int t = a.halfstep + wi * a.step;
    const int e0 = valid ? cp[t - a.flank + a.csc_pad] : 0, e1 = valid ? cp[t + a.flank + 1 + a.csc_pad] : 0;
    const int n = e1 - e0;
    int nmax = n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(NB_FULL, nmax, o));
    double SN = a.sn_nobias, SF = a.sf_nobias;
    if (a.use_bias && valid) {
        const int64_t wo = oo / a.step + c + wi;
        SN = a.wsn[wo];
        SF = a.wsf[wo];
    }
    const double rSN = 1.0 / SN, rSF = 1.0 / SF;
    double mant[NQ];
    int ex[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        mant[q] = 1.0;
        ex[q] = 0;
    }
    for (int base = 0; base < nmax; base += 8) {
        {   // lane (g, r) prepares fragment base + r of window g: nuc_probs / sum, nfr_probs / sum, Occupancy.py:106-109
            const int idx = base + r;
            double pv = 1.0, qv = 1.0;  // padding fragments contribute the factor 1
            if (idx < n) {
                const int sz = en[e0 + idx].y;
                pv = s_pn[sz] * rSN;
                qv = s_pf[sz] * rSF;
                const long long mb = __double_as_longlong(fmax(pv, qv));
                const int e2 = (int)((mb >> 52) & 0x7ff);
                if (e2 > 0 && e2 < 0x7fe) {  // scale the pair so that max(p, q) is in [1, 2): exact, constant in alpha
                    const double sc = __longlong_as_double((long long)(2046 - e2) << 52);
                    pv *= sc;
                    qv *= sc;
                }
            }
            s_d[warp][lane] = pv - qv;  // alpha*p + (1-alpha)*q is evaluated as q + alpha*(p-q): one FMA per (fragment, alpha)
            s_q[warp][lane] = qv;
            s_p[warp][lane] = pv;       // alpha == 1 uses p itself (q + (p-q) would lose p when p << q)
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const double dj = s_d[warp][8 * g + j], qj = s_q[warp][8 * g + j];
#pragma unroll
            for (int q = 0; q < NQ - 1; q++) mant[q] *= fma(al[q], dj, qj);
            double v = fma(al[NQ - 1], dj, qj);
            if (last_is_one) v = s_p[warp][8 * g + j];
            mant[NQ - 1] *= v;
        }
        __syncwarp();
        if ((base & 24) == 24) {  // every 32 factors (each in (2^-7, 2) away from the grid ends): exponent -> ex
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                const long long bits = __double_as_longlong(mant[q]);
                const int e2 = (int)((bits >> 52) & 0x7ff);
                if (e2 != 0 && e2 != 0x7ff) {  // positive normal
                    ex[q] += e2 - 1023;
                    mant[q] = __longlong_as_double(bits - ((long long)(e2 - 1023) << 52));
                }
            }
        }
    }
    double occ = nb_nan(), lo = nb_nan(), hi = nb_nan();
    {
        // Everything the reference does with the log-likelihoods is a comparison (np.argmax; 2*(max - ll) < cutoff,
        // Occupancy.py:115-119), and log is monotone: compare the products themselves, as exact (binary exponent,
        // mantissa in [1,2)) pairs -- no logarithm.  Dead grid points (0*log 0 = NaN -> -inf) and zero / NaN products
        // (log 0 = -inf) get the key (INT_MIN, 0).  The interval test ll > max - cutoff/2 is prod > max_prod * exp(-cutoff/2).
        int ke[NQ];
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            double mq = mant[q];
            int e = ex[q];
            const bool alive = !((dead >> q) & 1) && mq > 0.0;
            if (alive && mq < 2.2250738585072014e-308) {  // subnormal: make it normal first
                mq *= 18446744073709551616.0;            // 2^64
                e -= 64;
            }
            const long long bits = __double_as_longlong(mq);
            const int e2 = (int)((bits >> 52) & 0x7ff);
            ke[q] = alive ? e + e2 - 1023 : INT_MIN;
            mant[q] = alive ? __longlong_as_double((bits & 0x800fffffffffffffLL) | 0x3ff0000000000000LL) : 0.0;
        }
        int beste = INT_MIN, besti = 1 << 30;
        double bestm = 0.0;
#pragma unroll
        for (int q = 0; q < NQ; q++) {  // first maximum (np.argmax): strictly greater while walking up the grid
            const int ai = r + 8 * q;
            if (ai < a.n_alpha && (ke[q] > beste || (ke[q] == beste && mant[q] > bestm))) {
                beste = ke[q];
                bestm = mant[q];
                besti = ai;
            }
        }
        if (besti == (1 << 30) && r < a.n_alpha) besti = r;  // all -inf on this lane: its first alpha ties with the others
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const int oe = __shfl_xor_sync(NB_FULL, beste, o);
            const double om = __shfl_xor_sync(NB_FULL, bestm, o);
            const int oi = __shfl_xor_sync(NB_FULL, besti, o);
            const bool gt = oe > beste || (oe == beste && om > bestm), eq = oe == beste && om == bestm;
            if (gt || (eq && oi < besti)) {
                beste = oe;
                bestm = om;
                besti = oi;
            }
        }
        // threshold = max_prod * exp(-cutoff/2), canonical again
        int te = INT_MIN;       // no grid point passes when the maximum is -inf (2*(-inf - -inf) = NaN) or the cutoff is NaN
        double tm = 0.0;
        bool none = (beste == INT_MIN) || !(a.thr_m == a.thr_m);
        if (!none) {
            if (a.thr_zero) {   // exp(-cutoff/2) underflows: every finite log-likelihood passes
                te = INT_MIN + 1;
            } else {
                tm = bestm * a.thr_m;           // [1, 4)
                te = beste + a.thr_e;
                if (tm >= 2.0) {
                    tm *= 0.5;
                    te += 1;
                }
            }
        }
        int okmin = 1 << 30, okmax = -1;
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const int ai = r + 8 * q;
            if (ai < a.n_alpha && !none && (ke[q] > te || (ke[q] == te && mant[q] > tm))) {  // Occupancy.py:116-119
                okmin = min(okmin, ai);
                okmax = max(okmax, ai);
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            okmin = min(okmin, __shfl_xor_sync(NB_FULL, okmin, o));
            okmax = max(okmax, __shfl_xor_sync(NB_FULL, okmax, o));
        }
        if (n > 0 && okmax >= 0 && besti < a.n_alpha) {  // Occupancy.py:141 `if sum(new_inserts)>0`
            occ = a.alphas[besti];
            lo = a.alphas[okmin];
            hi = a.alphas[okmax];
        }
    }
    if (valid) {
        const int left = t - a.halfstep, right = min(t + a.halfstep + 1, L);
        for (int x = left + r; x < right; x += 8) {
            a.vals[oo + x] = occ;
            a.lower[oo + x] = lo;
            a.upper_b[oo + x] = hi;
        }
        if (wi == nwin - 1)  // positions past the last window stay NaN (np.ones(n)*nan, Occupancy.py:133-135)
            for (int x = right + r; x < L; x += 8) {
                a.vals[oo + x] = nb_nan();
                a.lower[oo + x] = nb_nan();
                a.upper_b[oo + x] = nb_nan();
            }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
struct OccPeakArgs {
    const int32_t *start;
    const int64_t *out_off, *col_off, *frag_off, *peak_off;
    const int32_t *col_ptr;
    const int2 *ent;
    const double *jitter;
    double *svals;                   // NaN -> min in place, like utils.py:86-91
    const double *slower, *supper;
    double *cov;
    int32_t *sc_pos;                 // scratch, packed like a track
    double *sc_val;
    unsigned char *sc_state;
    int32_t *peak_count, *peak_pos;
    double *peak_occ, *peak_lower, *peak_upper, *peak_reads, *nuc_dist;
    int upper, flank, sep, csc_pad;
    double min_occ;
};

#define PK_THREADS 512
__global__ void __launch_bounds__(PK_THREADS) k_occ_peaks(OccPeakArgs a)
{
    extern __shared__ unsigned char sm_pk[];
    double *s_nd = reinterpret_cast<double *>(sm_pk);          // [upper]
    int *s_hist = reinterpret_cast<int *>(s_nd + a.upper);      // [upper]
    __shared__ double red_d[32];
    __shared__ int red_i[32];
    __shared__ int s_base, s_flag;
    const int c = blockIdx.x;
    const int64_t oo = a.out_off[c];
    const int L = (int)(a.out_off[c + 1] - oo);
    const int32_t *cp = a.col_ptr + a.col_off[c];
    double *sv = a.svals + oo;
    double *cov = a.cov + oo;
    int32_t *cpos = a.sc_pos + oo;
    double *cval = a.sc_val + oo;
    unsigned char *cst = a.sc_state + oo;
    const int tid = threadIdx.x;
    // coverage: flat window over fragment centres = difference of the CSC prefix (tracks.py:209-222)
    for (int x = tid; x < L; x += blockDim.x)
        cov[x] = (double)(cp[x + a.flank + 1 + a.csc_pad] - cp[x - a.flank + a.csc_pad]);
    // NaN -> min (utils.py:86-91)
    double mn = CUDART_INF;
    int nnan = 0;
    for (int x = tid; x < L; x += blockDim.x) {
        double v = sv[x];
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
    for (int i = tid; i < a.upper; i += blockDim.x) s_nd[i] = 0.0;
    if (tid == 0) s_base = 0;
    __syncthreads();
    int m = 0;
    if (nnan < L) {
        if (nnan > 0)
            for (int x = tid; x < L; x += blockDim.x)
                if (sv[x] != sv[x]) sv[x] = mn;
        __syncthreads();
        // strict local maxima of sig*(1+jitter), order 1 (argrelmax, clip mode), utils.py:94-100
        const int boundary = a.sep / 2;
        const int lo = max(1, boundary), hi = min(L - 1, L - boundary);
        for (int x0 = 0; x0 < L; x0 += blockDim.x) {
            const int x = x0 + tid;
            int flag = 0;
            double v = 0.0;
            if (x >= lo && x < hi) {
                v = sv[x];
                double j0 = v * (1.0 + a.jitter[x]);
                double jl = sv[x - 1] * (1.0 + a.jitter[x - 1]);
                double jr = sv[x + 1] * (1.0 + a.jitter[x + 1]);
                flag = (j0 > jl) && (j0 > jr) && (v >= a.min_occ);
            }
            int slot = block_compact_slot(flag, &s_base, red_i);
            if (flag) {
                cpos[slot] = x;
                cval[slot] = v;
            }
        }
        __syncthreads();
        m = s_base;
        block_nms(cpos, cval, cst, m, a.sep, &s_flag);
    }
    __syncthreads();
    // OccPeak filter (Occupancy.py:228-231) and ordered output
    if (tid == 0) s_base = 0;
    __syncthreads();
    const int64_t po = a.peak_off[c];
    const int cap = (int)(a.peak_off[c + 1] - po);
    for (int j0 = 0; j0 < m; j0 += blockDim.x) {
        const int j = j0 + tid;
        int flag = 0, p = 0;
        if (j < m && cst[j] == 1) {
            p = cpos[j];
            flag = (a.slower[oo + p] > a.min_occ) && (cov[p] > 0);
        }
        int slot = block_compact_slot(flag, &s_base, red_i);
        if (flag && slot < cap) {
            a.peak_pos[po + slot] = a.start[c] + p;
            a.peak_occ[po + slot] = sv[p];
            a.peak_lower[po + slot] = a.slower[oo + p];
            a.peak_upper[po + slot] = a.supper[oo + p];
            a.peak_reads[po + slot] = cov[p];
        }
    }
    __syncthreads();
    const int npk = min(s_base, cap);
    if (tid == 0) a.peak_count[c] = (s_base <= cap) ? s_base : -s_base;  // negative: capacity exceeded
    // getNucDist, Occupancy.py:232-240: sum over peaks of the window's insert-size histogram / its total
    const int2 *en = a.ent + a.frag_off[c];
    for (int k = 0; k < npk; k++) {
        const int p = a.peak_pos[po + k] - a.start[c];
        const int e0 = cp[p - a.flank + a.csc_pad], e1 = cp[p + a.flank + 1 + a.csc_pad];
        for (int i = tid; i < a.upper; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        for (int e = e0 + tid; e < e1; e += blockDim.x) atomicAdd(&s_hist[en[e].y], 1);
        __syncthreads();
        const double tot = (double)(e1 - e0);
        for (int i = tid; i < a.upper; i += blockDim.x) s_nd[i] += (double)s_hist[i] / tot;
        __syncthreads();
    }
    for (int i = tid; i < a.upper; i += blockDim.x) a.nuc_dist[(int64_t)c * a.upper + i] = s_nd[i];
}

// ---------------------------------------------------------------------------------------------
extern "C" {

int nb200_occ_run(nb200_ctx *ctx, nb200_dbatch *b)
{
    if (!ctx || !b) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_occ_run: NULL argument");
    if (!ctx->occ_configured) return nb200_fail(ctx, NB200_ERR_STATE, "nb200_occ_configure has not been called");
    RunConst &r = ctx->rc;
    const nb200_occ_params &p = ctx->occ;
    if (!r.have_occ_model || r.occ_upper != p.upper)
        return nb200_fail(ctx, NB200_ERR_STATE, "nb200_set_occ_model missing or its size range differs from --upper");
    if (p.upper > NB200_MAX_UPPER) return nb200_fail(ctx, NB200_ERR_ARG, "upper > %d unsupported", NB200_MAX_UPPER);
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CUDA(ctx, cudaStreamWaitEvent(b->stream, b->ev_copied_occ, 0));   // a download of the previous pass may still read the arrays
    const int n = b->n_chunks;
    const int window = 2 * p.flank + 1;
    const int halfstep = (p.step - 1) / 2;
    if (b->min_len < p.smooth_len)
        return nb200_fail(ctx, NB200_ERR_ARG, "a chunk is shorter (%d) than the smoothing window (%d)", b->min_len, p.smooth_len);
    if (r.n_jitter < b->max_len)
        return nb200_fail(ctx, NB200_ERR_STATE, "nb200_set_jitter: need >= %d values (longest chunk), have %lld", b->max_len,
                          (long long)r.n_jitter);
    int pad = p.flank;
    if (ctx->nuc_configured && r.have_vmat && r.v_upper == p.upper) {  // share one CSC with the nuc path
        int npad = r.v_cols > r.v_upper / 2 + 1 ? r.v_cols : r.v_upper / 2 + 1;
        if (npad > pad) pad = npad;
    }
    int lower_split = (ctx->nuc_configured && r.have_vmat && r.v_upper == p.upper && r.v_lower > 0) ? r.v_lower : 0;
    NB_CHECK(nb200_prep_csc(ctx, b, pad, p.upper, p.atac, lower_split));
    if (p.use_bias) {
        NB_CHECK(nb200_prep_bias(ctx, b));
        for (int c = 0; c < n; c++) {  // Occupancy.py:131-132 / bias.py:93-107 flank checks
            int64_t b0 = (int64_t)b->h_seq_start[c] + r.pwm_up;
            int64_t b1 = b0 + (b->h_seq_off[c + 1] - b->h_seq_off[c]) - (r.pwm_width - 1);
            if (b0 > (int64_t)b->h_start[c] - p.flank - p.upper / 2 || b1 < (int64_t)b->h_end[c] + p.flank + p.upper / 2 + 1)
                return nb200_fail(ctx, NB200_ERR_FLANK,
                                  "Insufficient flanking region: chunk %d needs sequence over [%lld, %lld)", c,
                                  (long long)b->h_start[c] - p.flank - p.upper / 2 - r.pwm_up,
                                  (long long)b->h_end[c] + p.flank + p.upper / 2 + 1 + r.pwm_down);
        }
    }
    const size_t tl = (size_t)b->total_len;
    DevBuf *tracks[] = {&b->o_vals, &b->o_lower, &b->o_upper, &b->o_svals, &b->o_slower, &b->o_supper, &b->o_cov, &b->sc_f64};
    for (auto t : tracks) NB_CUDA(ctx, t->reserve(sizeof(double) * tl));
    NB_CUDA(ctx, b->sc_i32.reserve(sizeof(int32_t) * tl));
    NB_CUDA(ctx, b->sc_u8.reserve(tl));
    NB_CUDA(ctx, b->o_nuc_dist.reserve(sizeof(double) * (size_t)n * p.upper));
    // peak capacities: kept peaks are >= sep apart
    b->h_opeak_off.assign(n + 1, 0);
    for (int c = 0; c < n; c++) b->h_opeak_off[c + 1] = b->h_opeak_off[c] + (b->h_end[c] - b->h_start[c]) / p.sep + 2;
    const size_t np = (size_t)b->h_opeak_off[n];
    NB_CUDA(ctx, b->o_peak_off.reserve(sizeof(int64_t) * (n + 1)));
    if (b->occ_done) NB_CUDA(ctx, cudaStreamSynchronize(b->stream));  // re-run: the staging slot may still be in flight
    memcpy(b->pin_slot(4), b->h_opeak_off.data(), sizeof(int64_t) * (n + 1));
    NB_CUDA(ctx, cudaMemcpyAsync(b->o_peak_off.p, b->pin_slot(4), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, b->stream));
    NB_CUDA(ctx, b->o_peak_count.reserve(sizeof(int32_t) * n));
    NB_CUDA(ctx, b->o_peak_pos.reserve(sizeof(int32_t) * np));
    DevBuf *pk[] = {&b->o_peak_occ, &b->o_peak_lower, &b->o_peak_upper, &b->o_peak_reads};
    for (auto t : pk) NB_CUDA(ctx, t->reserve(sizeof(double) * np));

    if (p.use_bias && b->occ_cols_gen != ctx->occ_gen) {   // else: nb200_nuc_run of this batch already made them (one merged pass)
        const size_t ncs = tl + 2 * (size_t)p.flank * n;
        NB_CUDA(ctx, b->o_cn.reserve(sizeof(double) * ncs));
        NB_CUDA(ctx, b->o_cf.reserve(sizeof(double) * ncs));
        PairColsumArgs<2> pa;
        pa.start = b->d_start.as<int32_t>();
        pa.out_off = b->d_out_off.as<int64_t>();
        pa.bias_off = b->d_bias_off.as<int64_t>();
        pa.seq_start = b->d_seq_start.as<int32_t>();
        pa.E = b->d_E.as<double>();
        pa.wt[0] = r.nuc_probs.as<double>();
        pa.wt[1] = r.nfr_probs.as<double>();
        pa.out[0] = b->o_cn.as<double>();
        pa.out[1] = b->o_cf.as<double>();
        pa.pwm_up = r.pwm_up;
        for (int t = 0; t < 2; t++) {
            pa.lo[t] = 0;
            pa.hi[t] = p.upper;
            pa.pad[t] = p.flank;
        }
        const size_t smem = pair_colsums_smem<2>(p.upper);
        if (smem > 48 * 1024) NB_CUDA(ctx, cudaFuncSetAttribute(k_pair_colsums<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ProfScope ps(ctx, b->stream, "k_occ_colsums");
        dim3 grid((unsigned)div_up64(b->max_len + 2 * p.flank, 2 * PC_THREADS), n);
        k_pair_colsums<2><<<grid, PC_THREADS, smem, b->stream>>>(pa);
        NB_LAUNCH_CHECK(ctx);
        b->occ_cols_gen = ctx->occ_gen;
    }
    {
        OccMleArgs a;
        a.start = b->d_start.as<int32_t>();
        a.out_off = b->d_out_off.as<int64_t>();
        a.col_off = b->d_col_off.as<int64_t>();
        a.frag_off = b->d_frag_off.as<int64_t>();
        a.bias_off = b->d_bias_off.as<int64_t>();
        a.seq_start = b->d_seq_start.as<int32_t>();
        a.col_ptr = b->d_col_ptr.as<int32_t>();
        a.ent = b->d_ent.as<int2>();
        a.E = b->d_E.as<double>();
        a.cn = b->o_cn.as<double>();
        a.cf = b->o_cf.as<double>();
        a.pn = r.nuc_probs.as<double>();
        a.pf = r.nfr_probs.as<double>();
        a.alphas = r.alphas.as<double>();
        a.vals = b->o_vals.as<double>();
        a.lower = b->o_lower.as<double>();
        a.upper_b = b->o_upper.as<double>();
        a.pwm_up = r.pwm_up;
        a.upper = p.upper;
        a.flank = p.flank;
        a.step = p.step;
        a.halfstep = halfstep;
        a.csc_pad = b->csc_pad;
        a.n_alpha = r.n_alpha;
        a.use_bias = p.use_bias;
        a.pn_has_zero = r.pn_has_zero;
        a.pf_has_zero = r.pf_has_zero;
        a.both_zero = r.both_zero;
        a.cutoff = r.cutoff;
        {
            const double k = exp(-0.5 * r.cutoff);
            int e = 0;
            const double m = frexp(k, &e);   // k = m * 2^e, m in [0.5, 1)
            a.thr_zero = (k == 0.0) ? 1 : 0;
            a.thr_m = (k > 0.0 && std::isfinite(k)) ? 2.0 * m : (k == 0.0 ? 1.0 : (double)NAN);  // NaN: nothing passes
            a.thr_e = e - 1;
        }
        a.sn_nobias = r.pn_sum * window;
        a.sf_nobias = r.pf_sum * window;
        int max_win = (b->max_len - halfstep + p.step - 1) / p.step;
        if (max_win < 1) max_win = 1;
        const size_t smem = sizeof(double) * 2 * (size_t)p.upper;
        a.wsn = a.wsf = nullptr;
        if (p.use_bias) {
            const size_t nws = tl / p.step + n + 2;
            NB_CUDA(ctx, b->o_wsn.reserve(sizeof(double) * nws));
            NB_CUDA(ctx, b->o_wsf.reserve(sizeof(double) * nws));
            a.wsn = b->o_wsn.as<double>();
            a.wsf = b->o_wsf.as<double>();
            const size_t wsm = sizeof(double) * 2 * ((size_t)(WS_WIN - 1) * p.step + window);
            if (wsm > 48 * 1024) NB_CUDA(ctx, cudaFuncSetAttribute(k_occ_winsums, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsm));
            ProfScope ps(ctx, b->stream, "k_occ_winsums");
            dim3 wgrid((unsigned)div_up64(max_win, WS_WIN), n);
            k_occ_winsums<<<wgrid, WS_WIN, wsm, b->stream>>>(a.out_off, a.cn, a.cf, p.flank, p.step, halfstep, b->o_wsn.as<double>(),
                                                            b->o_wsf.as<double>());
            NB_LAUNCH_CHECK(ctx);
        }
        ProfScope ps(ctx, b->stream, "k_occ_mle");
        dim3 grid((unsigned)div_up64(max_win, MLE_WARPS * MLE_GROUPS * MLE_ITERS), n);
        static const int mle_lb = getenv("NB200_MLE_LB") ? atoi(getenv("NB200_MLE_LB")) : 4;
        if (r.n_alpha <= 104) {
            if (mle_lb >= 5)
                k_occ_mle<13, 5><<<grid, MLE_WARPS * 32, smem, b->stream>>>(a);
            else
                k_occ_mle<13, 4><<<grid, MLE_WARPS * 32, smem, b->stream>>>(a);
        } else
            k_occ_mle<16, 4><<<grid, MLE_WARPS * 32, smem, b->stream>>>(a);
        NB_LAUNCH_CHECK(ctx);
    }
    {
        SmoothTracks tr;
        tr.in[0] = b->o_vals.as<double>();
        tr.in[1] = b->o_lower.as<double>();
        tr.in[2] = b->o_upper.as<double>();
        tr.out[0] = b->o_svals.as<double>();
        tr.out[1] = b->o_slower.as<double>();
        tr.out[2] = b->o_supper.as<double>();
        size_t smem = smooth_same_smem(p.smooth_len);
        if (smem > 48 * 1024) NB_CUDA(ctx, cudaFuncSetAttribute(k_smooth_same, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ProfScope ps(ctx, b->stream, "k_smooth_same");
        dim3 grid((unsigned)div_up64(b->max_len, SM_TILE), n, 3);
        k_smooth_same<<<grid, SM_THREADS, smem, b->stream>>>(tr, b->d_out_off.as<int64_t>(), r.occ_win.as<double>(), p.smooth_len, 0);
        NB_LAUNCH_CHECK(ctx);
    }
    {
        OccPeakArgs a;
        a.start = b->d_start.as<int32_t>();
        a.out_off = b->d_out_off.as<int64_t>();
        a.col_off = b->d_col_off.as<int64_t>();
        a.frag_off = b->d_frag_off.as<int64_t>();
        a.peak_off = b->o_peak_off.as<int64_t>();
        a.col_ptr = b->d_col_ptr.as<int32_t>();
        a.ent = b->d_ent.as<int2>();
        a.jitter = r.jitter.as<double>();
        a.svals = b->o_svals.as<double>();
        a.slower = b->o_slower.as<double>();
        a.supper = b->o_supper.as<double>();
        a.cov = b->o_cov.as<double>();
        a.sc_pos = b->sc_i32.as<int32_t>();
        a.sc_val = b->sc_f64.as<double>();
        a.sc_state = b->sc_u8.as<unsigned char>();
        a.peak_count = b->o_peak_count.as<int32_t>();
        a.peak_pos = b->o_peak_pos.as<int32_t>();
        a.peak_occ = b->o_peak_occ.as<double>();
        a.peak_lower = b->o_peak_lower.as<double>();
        a.peak_upper = b->o_peak_upper.as<double>();
        a.peak_reads = b->o_peak_reads.as<double>();
        a.nuc_dist = b->o_nuc_dist.as<double>();
        a.upper = p.upper;
        a.flank = p.flank;
        a.sep = p.sep;
        a.csc_pad = b->csc_pad;
        a.min_occ = p.min_occ;
        size_t smem = (size_t)p.upper * 12;
        ProfScope ps(ctx, b->stream, "k_occ_peaks");
        k_occ_peaks<<<n, PK_THREADS, smem, b->stream>>>(a);
        NB_LAUNCH_CHECK(ctx);
    }
    b->occ_upper = p.upper;
    b->occ_done = true;
    return NB200_OK;
}

static int d2h(nb200_ctx *ctx, nb200_dbatch *b, void *dst, const DevBuf &src, size_t bytes)
{
    if (!dst || !bytes) return NB200_OK;
    NB_CUDA(ctx, cudaMemcpyAsync(dst, src.p, bytes, cudaMemcpyDeviceToHost, b->copy_stream));
    return NB200_OK;
}

int nb200_occ_download(nb200_ctx *ctx, nb200_dbatch *b, const nb200_occ_out *o)
{
    if (!ctx || !b || !o) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_occ_download: NULL argument");
    if (!b->occ_done) return nb200_fail(ctx, NB200_ERR_STATE, "nb200_occ_download: nb200_occ_run has not been called on this batch");
    const size_t tb = sizeof(double) * (size_t)b->total_len;
    const int n = b->n_chunks;
    NB_CUDA(ctx, cudaEventRecord(b->ev_pass, b->stream));            // copies start once the pass has finished ...
    NB_CUDA(ctx, cudaStreamWaitEvent(b->copy_stream, b->ev_pass, 0));
    struct Copied {                                                  // ... and the next pass over these arrays waits for them
        nb200_dbatch *b;
        ~Copied() { cudaEventRecord(b->ev_copied_occ, b->copy_stream); }
    } copied{b};
    NB_CHECK(d2h(ctx, b, o->smoothed_vals, b->o_svals, tb));
    NB_CHECK(d2h(ctx, b, o->smoothed_lower, b->o_slower, tb));
    NB_CHECK(d2h(ctx, b, o->smoothed_upper, b->o_supper, tb));
    NB_CHECK(d2h(ctx, b, o->vals, b->o_vals, tb));
    NB_CHECK(d2h(ctx, b, o->lower_bound, b->o_lower, tb));
    NB_CHECK(d2h(ctx, b, o->upper_bound, b->o_upper, tb));
    NB_CHECK(d2h(ctx, b, o->cov, b->o_cov, tb));
    NB_CHECK(d2h(ctx, b, o->nuc_dist, b->o_nuc_dist, sizeof(double) * (size_t)n * ctx->occ.upper));
    NB_CHECK(d2h(ctx, b, o->peak_count, b->o_peak_count, sizeof(int32_t) * n));
    if (o->peak_pos) {
        if (!o->peak_off) return nb200_fail(ctx, NB200_ERR_ARG, "nb200_occ_download: peak_off is required with peak_pos");
        for (int c = 0; c <= n; c++)
            if (o->peak_off[c] != b->h_opeak_off[c])
                return nb200_fail(ctx, NB200_ERR_CAPACITY, "nb200_occ_download: peak_off must equal len/sep+2 capacities (chunk %d)", c);
        const size_t np = (size_t)b->h_opeak_off[n];
        NB_CHECK(d2h(ctx, b, o->peak_pos, b->o_peak_pos, sizeof(int32_t) * np));
        NB_CHECK(d2h(ctx, b, o->peak_occ, b->o_peak_occ, sizeof(double) * np));
        NB_CHECK(d2h(ctx, b, o->peak_lower, b->o_peak_lower, sizeof(double) * np));
        NB_CHECK(d2h(ctx, b, o->peak_upper, b->o_peak_upper, sizeof(double) * np));
        NB_CHECK(d2h(ctx, b, o->peak_reads, b->o_peak_reads, sizeof(double) * np));
    }
    return NB200_OK;
}

int64_t nb200_occ_d2h_bytes(nb200_dbatch *b, const nb200_occ_out *o)
{
    if (!b || !o) return 0;
    const int64_t tb = 8 * b->total_len;
    int64_t s = 0;
    const void *tr[] = {o->smoothed_vals, o->smoothed_lower, o->smoothed_upper, o->vals, o->lower_bound, o->upper_bound, o->cov};
    for (auto p : tr)
        if (p) s += tb;
    if (o->nuc_dist) s += 8LL * b->n_chunks * b->occ_upper;
    if (o->peak_count) s += 4LL * b->n_chunks;
    if (o->peak_pos && !b->h_opeak_off.empty()) {
        int64_t np = b->h_opeak_off[b->n_chunks];
        s += 4 * np;
        const void *pk[] = {o->peak_occ, o->peak_lower, o->peak_upper, o->peak_reads};
        for (auto p : pk)
            if (p) s += 8 * np;
    }
    return s;
}

}  // extern "C"
