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
    const double *wsn, *wsf;   // per window: RECIPROCALS of the sums of cn / cf over its 2*flank+1 columns (k_occ_winsums); window k of chunk c at out_off[c]/step + c + k
    double *vals, *lower, *upper_b;
    double *wv;            // per-window occ / lower / upper (3 slabs of wv_stride), or null
    int64_t wv_stride;
    int pwm_up, upper, flank, step, halfstep, csc_pad, n_alpha, use_bias;
    int pn_has_zero, pf_has_zero, both_zero;
    double cutoff, sn_nobias, sf_nobias;   // the last two: 1 / (sum of the model over a window) without a bias model
    double thr_m;   // exp(-cutoff/2) = thr_m * 2^thr_e, thr_m in [1,2) (NaN for a NaN cutoff); thr_zero: it underflows to 0
    int thr_e, thr_zero;
    double thr_k;   // exp(-cutoff/2) itself; fast_epi: it is a normal number well inside the double range (the direct comparisons of
    int fast_epi;   // k_occ_mle's short epilogue round exactly like the canonical ones)
    int stage;      // k_occ_mle stages its block's column pointers / fragment sizes / window normalisers in shared memory
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
    // stored as reciprocals: every lane of the likelihood kernels would otherwise divide by the same two sums (the division
    // is the same correctly rounded IEEE operation here as there, so the grids are bit-identical)
    wsn[wo] = 1.0 / ((n4[0] + n4[1]) + (n4[2] + n4[3]));
    wsf[wo] = 1.0 / ((f4[0] + f4[1]) + (f4[2] + f4[3]));
}

// One window per 8-lane group, 4 windows per warp (Occupancy.py:104-146).  The bias of a fragment's insert size over the
// window multiplies both mixture components (nuc[s] = pn[s]*bias[s]/SN, nfr[s] = pf[s]*bias[s]/SF, Occupancy.py:106-109),
// so it is a factor of the likelihood that does not depend on alpha: it drops out of the argmax and of the likelihood
// ratios 2*(max - ll).  What a window needs from the bias model is only SN and SF (sums of the per-column sums cn, cf).
// ll[a] = sum_f log(v_f(a)) is evaluated as log(prod_f v_f) with every factor pre-scaled by a power of two (again constant
// in alpha) and the running product renormalised every 64 factors: no logarithm at all, the argmax and the interval test
// compare the products (epilogues below).
#define MLE_WARPS 4
#define MLE_GROUPS 4
#ifndef MLE_ITERS
#define MLE_ITERS 8
#endif
#ifndef MLE_GL_DEFAULT
#define MLE_GL_DEFAULT 8
#endif
#ifndef MLE_LB_DEFAULT
#define MLE_LB_DEFAULT 4
#endif
#define MLE_WPB (MLE_WARPS * MLE_GROUPS * MLE_ITERS)   // windows per block (consecutive windows of one chunk)
#define MLE_CPK (((MLE_WPB - 1) * 5 + 2 * 60 + 2 + MLE_WARPS * 32 - 1) / (MLE_WARPS * 32))   // staged column pointers per thread at the default step / flank
#define MLE_WPK ((MLE_WPB + MLE_WARPS * 32 - 1) / (MLE_WARPS * 32))                             // staged window normalisers per thread
#ifndef MLE_CAPF
#define MLE_CAPF 1024                                   // fragment sizes staged per block (the rest is read from global memory)
#endif
// NQ alphas per lane: lane r of a window's GL-lane group owns alphas r, r + GL, ...; LB resident blocks per SM; a warp scores
// ITERS groups of 32 / GL windows.  GL = 16 (7 alphas per lane, ~64 registers, twice the resident warps) and GL = 8 (13 alphas
// per lane) return identical grids: a grid point's product takes the same factors in the same order either way.
template <int NQ, int LB, int GL, int ITERS>
__global__ void __launch_bounds__(MLE_WARPS * 32, LB) k_occ_mle(OccMleArgs a)
{
    // dynamic: pn[upper], pf[upper] | staged inputs of the block's MLE_WPB consecutive windows (a.stage): wsn[WPB], wsf[WPB],
    // the column pointers its windows' ends read, the sizes of the first MLE_CAPF fragments under them
    extern __shared__ double sm_mle[];
    __shared__ double s_d[MLE_WARPS][32], s_q[MLE_WARPS][32], s_p[MLE_WARPS][32];
    static_assert(MLE_WARPS * (32 / GL) * ITERS == MLE_WPB, "windows per block");
    double *s_pn = sm_mle, *s_pf = sm_mle + a.upper;
    double *s_wsn = s_pf + a.upper, *s_wsf = s_wsn + MLE_WPB;
    int *s_cp = reinterpret_cast<int *>(s_wsf + MLE_WPB), *s_sz = s_cp + (((MLE_WPB - 1) * a.step + 2 * a.flank + 2 + 3) & ~3);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane / GL, r = lane % GL;
    const int c = blockIdx.y;
    const int64_t oo = a.out_off[c];
    const int L = (int)(a.out_off[c + 1] - oo);
    const int nwin = (L - a.halfstep + a.step - 1) / a.step;  // windows at t = halfstep + k*step < L
    const int32_t *cp = a.col_ptr + a.col_off[c];
    const int2 *en = a.ent + a.frag_off[c];
    const int wb0 = blockIdx.x * MLE_WPB;                       // first window of the block
    if (wb0 >= nwin) return;
    const int c_lo = a.halfstep + wb0 * a.step - a.flank + a.csc_pad;   // column pointer the first window starts at
    int eA = 0;
    for (int i = threadIdx.x; i < a.upper; i += blockDim.x) {
        s_pn[i] = a.pn[i];
        s_pf[i] = a.pf[i];
    }
    if (a.stage) {
        // Every load the warps' loops would otherwise wait for (long-scoreboard stalls were 12 % of the kernel's samples with
        // four warps per scheduler to hide them) is made once per block, coalesced.
        const int nw = min(MLE_WPB, nwin - wb0);
        const int ncp = (nw - 1) * a.step + 2 * a.flank + 2;
        eA = cp[c_lo];
        const int eB = cp[c_lo + ncp - 1];
        // all the loads of a thread are issued before its first store (one exposed global-memory latency per block, not one per pass)
        int vcp[MLE_CPK];
#pragma unroll
        for (int k = 0; k < MLE_CPK; k++) {
            const int i = threadIdx.x + k * (MLE_WARPS * 32);
            vcp[k] = (i < ncp) ? cp[c_lo + i] : 0;
        }
        double vn[MLE_WPK], vf[MLE_WPK];
#pragma unroll
        for (int k = 0; k < MLE_WPK; k++) {
            const int i = threadIdx.x + k * (MLE_WARPS * 32);
            vn[k] = vf[k] = 0.0;
            if (a.use_bias && i < nw) {
                const int64_t wo0 = oo / a.step + c + wb0;
                vn[k] = a.wsn[wo0 + i];
                vf[k] = a.wsf[wo0 + i];
            }
        }
        const int nf = min(eB - eA, MLE_CAPF);
        int vsz[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = threadIdx.x + k * (MLE_WARPS * 32);
            vsz[k] = (i < nf) ? en[eA + i].y : 0;
        }
#pragma unroll
        for (int k = 0; k < MLE_CPK; k++) {
            const int i = threadIdx.x + k * (MLE_WARPS * 32);
            if (i < ncp) s_cp[i] = vcp[k];
        }
        for (int i = threadIdx.x + MLE_CPK * (MLE_WARPS * 32); i < ncp; i += blockDim.x) s_cp[i] = cp[c_lo + i];   // steps / flanks beyond the default
#pragma unroll
        for (int k = 0; k < MLE_WPK; k++) {   // slots past the chunk's last window are never read
            const int i = threadIdx.x + k * (MLE_WARPS * 32);
            if (i < MLE_WPB) {
                s_wsn[i] = vn[k];
                s_wsf[i] = vf[k];
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = threadIdx.x + k * (MLE_WARPS * 32);
            if (i < nf) s_sz[i] = vsz[k];
        }
        for (int i = threadIdx.x + 4 * (MLE_WARPS * 32); i < nf; i += blockDim.x) s_sz[i] = en[eA + i].y;
    }
    __syncthreads();
    // grid constants of this lane: its alphas and which of them are dead (0 * log 0 = NaN -> -inf, Occupancy.py:112-114)
    double al[NQ];
    unsigned dead = 0;
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const int ai = r + GL * q;
        al[q] = (ai < a.n_alpha) ? a.alphas[ai] : 0.5;
        if ((ai >= a.n_alpha) || a.both_zero || (al[q] == 0.0 && a.pf_has_zero) || (al[q] == 1.0 && a.pn_has_zero)) dead |= 1u << q;
    }
    const bool last_is_one = (al[NQ - 1] == 1.0);  // alpha == 1 (only ever the last grid value, checked on the host)
  for (int it = 0; it < ITERS; it++) {   // a warp scores ITERS groups of 32 / GL windows
    const int wbase = ((blockIdx.x * ITERS + it) * MLE_WARPS + warp) * (32 / GL);
    if (wbase >= nwin) break;  // whole warp past the last window
    const int wi = wbase + g;
    const bool valid = wi < nwin;
    const int t = a.halfstep + wi * a.step;
    int e0 = 0, e1 = 0;
    if (valid) {
        if (a.stage) {
            e0 = s_cp[(wi - wb0) * a.step];
            e1 = s_cp[(wi - wb0) * a.step + 2 * a.flank + 1];
        } else {
            e0 = cp[t - a.flank + a.csc_pad];
            e1 = cp[t + a.flank + 1 + a.csc_pad];
        }
    }
    const int n = e1 - e0;
    const int so = e0 - eA;   // slot of the window's first fragment in s_sz
    int nmax = n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(NB_FULL, nmax, o));
    double rSN = a.sn_nobias, rSF = a.sf_nobias;   // reciprocals of the window's normalisers
    if (a.use_bias && valid) {
        if (a.stage) {
            rSN = s_wsn[wi - wb0];
            rSF = s_wsf[wi - wb0];
        } else {
            const int64_t wo = oo / a.step + c + wi;
            rSN = a.wsn[wo];
            rSF = a.wsf[wo];
        }
    }
    double mant[NQ];
    int ex[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        mant[q] = ((dead >> q) & 1) ? 0.0 : 1.0;   // a dead grid point's product is 0 from the start (-inf either way below)
        ex[q] = 0;
    }
    for (int base = 0; base < nmax; base += GL) {
        {   // lane (g, r) prepares fragment base + r of window g: nuc_probs / sum, nfr_probs / sum, Occupancy.py:106-109
            const int idx = base + r;
            double pv = 1.0, qv = 1.0;  // padding fragments contribute the factor 1
            if (idx < n) {
                const int sz = (a.stage && so + idx < MLE_CAPF) ? s_sz[so + idx] : en[e0 + idx].y;
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
        auto factor = [&](int j) {   // fragment j of the round into the NQ products of this lane
            const double dj = s_d[warp][GL * g + j], qj = s_q[warp][GL * g + j];
#pragma unroll
            for (int q = 0; q < NQ - 1; q++) mant[q] *= fma(al[q], dj, qj);
            double v = fma(al[NQ - 1], dj, qj);
            if (last_is_one) v = s_p[warp][GL * g + j];
            mant[NQ - 1] *= v;
        };
        if (nmax - base >= GL) {
#pragma unroll
            for (int j = 0; j < GL; j++) factor(j);
        } else {   // last round: only its first nmax - base slots hold a fragment of any window (the rest are factors of exactly 1.0)
#pragma unroll 1
            for (int j = 0; j < nmax - base; j++) factor(j);
        }
        __syncwarp();
        if (((base + GL) & 63) == 0) {  // every 64 factors (each in (2^-7, 2) away from the grid ends, so the product stays above 2^-448): exponent -> ex
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
    bool done = false;
    if (nmax <= 64 - GL && a.fast_epi) {
        // Short epilogue (warp-uniform condition): no product of this warp has been renormalised yet (all ex[] are 0), so the
        // products are plain doubles and every comparison of the canonical epilogue below is the same comparison made on them
        // directly -- exact for any pair of non-negative doubles, subnormal or not; dead points and zero / NaN products hold
        // 0 / NaN and never compare greater.  The one rounded operation, threshold = max * exp(-cutoff/2), rounds like the
        // canonical mantissa product as long as the result is a normal number: a warp with a maximum below 1e-250 takes the
        // canonical epilogue instead.
        double best = 0.0;
        int besti = 1 << 30;
#pragma unroll
        for (int q = 0; q < NQ; q++)
            if (mant[q] > best) {   // first maximum (np.argmax): strictly greater while walking up the grid
                best = mant[q];
                besti = r + GL * q;
            }
        if (besti == (1 << 30) && r < a.n_alpha) besti = r;
#pragma unroll
        for (int o = GL / 2; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(NB_FULL, best, o);
            const int oi = __shfl_xor_sync(NB_FULL, besti, o);
            if (ob > best || (ob == best && oi < besti)) {
                best = ob;
                besti = oi;
            }
        }
        if (!__any_sync(NB_FULL, best > 0.0 && best < 1e-250)) {
            done = true;
            const double thr = best * a.thr_k;   // Occupancy.py:116-119: ll > max - cutoff/2
            int okmin = 1 << 30, okmax = -1;
#pragma unroll
            for (int q = 0; q < NQ; q++)
                if (mant[q] > thr) {
                    okmin = min(okmin, r + GL * q);
                    okmax = max(okmax, r + GL * q);
                }
#pragma unroll
            for (int o = GL / 2; o > 0; o >>= 1) {
                okmin = min(okmin, __shfl_xor_sync(NB_FULL, okmin, o));
                okmax = max(okmax, __shfl_xor_sync(NB_FULL, okmax, o));
            }
            if (n > 0 && okmax >= 0 && besti < a.n_alpha) {  // Occupancy.py:141 `if sum(new_inserts)>0`
                occ = a.alphas[besti];
                lo = a.alphas[okmin];
                hi = a.alphas[okmax];
            }
        }
    }
    if (!done) {
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
            const int ai = r + GL * q;
            if (ai < a.n_alpha && (ke[q] > beste || (ke[q] == beste && mant[q] > bestm))) {
                beste = ke[q];
                bestm = mant[q];
                besti = ai;
            }
        }
        if (besti == (1 << 30) && r < a.n_alpha) besti = r;  // all -inf on this lane: its first alpha ties with the others
#pragma unroll
        for (int o = GL / 2; o > 0; o >>= 1) {
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
            const int ai = r + GL * q;
            if (ai < a.n_alpha && !none && (ke[q] > te || (ke[q] == te && mant[q] > tm))) {  // Occupancy.py:116-119
                okmin = min(okmin, ai);
                okmax = max(okmax, ai);
            }
        }
#pragma unroll
        for (int o = GL / 2; o > 0; o >>= 1) {
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
        if (r == 0 && a.wv) {   // per-window values (the block smoother's input)
            const int64_t wo = oo / a.step + c + wi;
            a.wv[wo] = occ;
            a.wv[a.wv_stride + wo] = lo;
            a.wv[2 * a.wv_stride + wo] = hi;
        }
        const int left = t - a.halfstep, right = min(t + a.halfstep + 1, L);
#pragma unroll 1
        for (int x = left + r; x < right; x += GL) {
            a.vals[oo + x] = occ;
            a.lower[oo + x] = lo;
            a.upper_b[oo + x] = hi;
        }
        if (wi == nwin - 1)  // positions past the last window stay NaN (np.ones(n)*nan, Occupancy.py:133-135)
#pragma unroll 1
            for (int x = right + r; x < L; x += GL) {
                a.vals[oo + x] = nb_nan();
                a.lower[oo + x] = nb_nan();
                a.upper_b[oo + x] = nb_nan();
            }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// Guarded search over the alpha grid (NB200_MLE_SEARCH=group; the scan k_occ_mle above is the default: it is faster).
//
// log L(alpha) = sum_f log(q_f + alpha (p_f - q_f)) is concave in alpha, so (a) the first maximum over the grid lies strictly
// between the neighbours of the best point of any coarser sub-grid and (b) the grid points passing the likelihood-ratio
// test 2 (max - ll) < cutoff form one interval around it.  Three rounds of 16 grid points per window (8 lanes x 2) replace
// the scan of all 101:
//   round 1  16 evenly spaced points c_0 .. c_15 (spacing <= 8)
//   round 2  every grid point strictly between the neighbours of the best coarse point -> the first maximum, exactly as
//            np.argmax over the whole grid finds it
//   round 3  the grid points inside the two coarse intervals in which the pass / fail status changes -> both bounds
// Every comparison is the exact (exponent, mantissa) comparison of k_occ_mle on identically computed products, so the
// result is the full scan's whenever the computed values are ordered like the true ones.  That can only fail at ties at
// rounding level; a window whose decisive values come within 1e-11 (relative) of each other -- the best coarse point
// against its neighbours, any evaluated point against the interval threshold -- is re-scored by a scan of the whole grid
// (the warp does it together; windows without fragments never tie: they are NaN by Occupancy.py:141).
// The window's fragments are prepared once (scaled p, q, p - q) into a per-window shared-memory cache; fragments beyond
// MLS_CAP are prepared again in every round.
#ifndef MLS_CAP
#define MLS_CAP 48
#endif
#define MLS_GROUP_STRIDE (2 * MLS_CAP + 2)
#define MLS_TIE 1e-11
struct MleKey {   // canonical likelihood: 2^e * m, m in [1, 2); e = INT_MIN: -inf (dead grid point, zero / NaN product)
    int e;
    double m;
};
__device__ __forceinline__ bool key_gt(const MleKey &a, const MleKey &b) { return a.e > b.e || (a.e == b.e && a.m > b.m); }
// a >= b * (1 - tol): the two likelihoods are equal at rounding level or a is the larger one
__device__ __forceinline__ bool key_close_or_above(const MleKey &a, const MleKey &b, double tol)
{
    if (b.e == INT_MIN) return true;
    if (a.e == INT_MIN) return false;
    if (a.e > b.e) return true;
    if (a.e == b.e) return a.m >= b.m * (1.0 - tol);
    if (a.e == b.e - 1) return a.m * 0.5 >= b.m * (1.0 - tol);
    return false;
}
__device__ __forceinline__ bool key_close(const MleKey &a, const MleKey &b, double tol)
{
    return key_close_or_above(a, b, tol) && key_close_or_above(b, a, tol);
}

template <int LB>
__global__ void __launch_bounds__(MLE_WARPS * 32, LB) k_occ_mle_search(OccMleArgs a)
{
    extern __shared__ __align__(16) double sm_mle[];  // pn[upper], pf[upper] | per window group: (d, q)[MLS_CAP], (0, p)[MLS_CAP]
    __shared__ double s_d[MLE_WARPS][32], s_q[MLE_WARPS][32], s_p[MLE_WARPS][32];   // overflow fragments (index >= MLS_CAP)
    const int up2 = (a.upper + 1) & ~1;
    double *s_pn = sm_mle, *s_pf = sm_mle + up2;
    for (int i = threadIdx.x; i < a.upper; i += blockDim.x) {
        s_pn[i] = a.pn[i];
        s_pf[i] = a.pf[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3, r = lane & 7;
    // group regions are 2 * MLS_CAP + 2 entries apart (32 bytes modulo a 128-byte row) and the (0, p) copy sits one entry
    // past MLS_CAP: the four groups of a warp and the two arrays read eight different 16-byte bank slots
    double2 *c_dq = reinterpret_cast<double2 *>(sm_mle + 2 * up2) + (size_t)(warp * MLE_GROUPS + g) * MLS_GROUP_STRIDE;
    double2 *c_0p = c_dq + MLS_CAP + 1;
    const int c = blockIdx.y;
    const int64_t oo = a.out_off[c];
    const int L = (int)(a.out_off[c + 1] - oo);
    const int nwin = (L - a.halfstep + a.step - 1) / a.step;
    const int32_t *cp = a.col_ptr + a.col_off[c];
    const int2 *en = a.ent + a.frag_off[c];
    const int NA = a.n_alpha;
    const int last = NA - 1;
    const unsigned gmask = 0xffu << (8 * g);
  for (int it = 0; it < MLE_ITERS; it++) {
    const int wbase = ((blockIdx.x * MLE_ITERS + it) * MLE_WARPS + warp) * MLE_GROUPS;
    if (wbase >= nwin) break;
    const int wi = wbase + g;
    const bool valid = wi < nwin;
    const int t = a.halfstep + wi * a.step;
    const int e0 = valid ? cp[t - a.flank + a.csc_pad] : 0, e1 = valid ? cp[t + a.flank + 1 + a.csc_pad] : 0;
    const int n = e1 - e0;
    int nmax = n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(NB_FULL, nmax, o));
    double rSN = a.sn_nobias, rSF = a.sf_nobias;   // reciprocals of the window's normalisers
    if (a.use_bias && valid) {
        const int64_t wo = oo / a.step + c + wi;
        rSN = a.wsn[wo];
        rSF = a.wsf[wo];
    }
    // lane (g, r) prepares fragment idx of window g: nuc_probs / sum, nfr_probs / sum (Occupancy.py:106-109), the pair
    // scaled by a power of two so that max(p, q) is in [1, 2) (exact, constant in alpha); padding fragments are the factor 1
    auto prep = [&](int idx, double &pv, double &qv) {
        pv = 1.0;
        qv = 1.0;
        if (idx < n) {
            const int sz = en[e0 + idx].y;
            pv = s_pn[sz] * rSN;
            qv = s_pf[sz] * rSF;
            const long long mb = __double_as_longlong(fmax(pv, qv));
            const int e2 = (int)((mb >> 52) & 0x7ff);
            if (e2 > 0 && e2 < 0x7fe) {
                const double sc = __longlong_as_double((long long)(2046 - e2) << 52);
                pv *= sc;
                qv *= sc;
            }
        }
    };
    const int ncache = min((nmax + 7) & ~7, MLS_CAP);
    __syncwarp();   // the previous iteration's readers of the cache are done
    for (int base = 0; base < ncache; base += 8) {
        double pv, qv;
        prep(base + r, pv, qv);
        c_dq[base + r] = make_double2(pv - qv, qv);   // alpha p + (1 - alpha) q = q + alpha (p - q): one FMA per (fragment, alpha)
        c_0p[base + r] = make_double2(0.0, pv);       // alpha == 1 takes p itself (q + (p - q) would lose p when p << q)
    }
    __syncwarp();
    // ---- one round: this lane's two grid points i0, i1 (-1: none) -> canonical keys
    // may_one (warp uniform): some lane of the warp may hold alpha == 1, which reads the (0, p) copy instead of (d, q);
    // only the coarse round and the full scan touch the last grid point, the other rounds share one load per fragment
    auto eval2 = [&](int i0, int i1, MleKey &k0, MleKey &k1, const bool may_one) {
        const double al0 = (i0 >= 0) ? a.alphas[i0] : 0.5, al1 = (i1 >= 0) ? a.alphas[i1] : 0.5;
        const bool one0 = al0 == 1.0, one1 = al1 == 1.0;
        const double2 *src0 = one0 ? c_0p : c_dq, *src1 = one1 ? c_0p : c_dq;
        double m0 = 1.0, m1 = 1.0;
        int x0 = 0, x1 = 0;
        auto renorm = [](double &m, int &x) {
            const long long bits = __double_as_longlong(m);
            const int e2 = (int)((bits >> 52) & 0x7ff);
            if (e2 != 0 && e2 != 0x7ff) {
                x += e2 - 1023;
                m = __longlong_as_double(bits - ((long long)(e2 - 1023) << 52));
            }
        };
        if (may_one) {
            for (int j0 = 0; j0 < ncache; j0 += 8) {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const double2 u = src0[j0 + j], v = src1[j0 + j];
                    m0 *= fma(al0, u.x, u.y);
                    m1 *= fma(al1, v.x, v.y);
                }
                if ((j0 & 24) == 24) {   // every 32 factors (each in (2^-7, 2) away from the grid ends)
                    renorm(m0, x0);
                    renorm(m1, x1);
                }
            }
        } else {
            for (int j0 = 0; j0 < ncache; j0 += 8) {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const double2 u = c_dq[j0 + j];
                    m0 *= fma(al0, u.x, u.y);
                    m1 *= fma(al1, u.x, u.y);
                }
                if ((j0 & 24) == 24) {
                    renorm(m0, x0);
                    renorm(m1, x1);
                }
            }
        }
        for (int base = MLS_CAP; base < nmax; base += 8) {   // windows with more fragments than the cache holds
            double pv, qv;
            prep(base + r, pv, qv);
            s_d[warp][lane] = pv - qv;
            s_q[warp][lane] = qv;
            s_p[warp][lane] = pv;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const double dj = s_d[warp][8 * g + j], qj = s_q[warp][8 * g + j], pj = s_p[warp][8 * g + j];
                m0 *= one0 ? pj : fma(al0, dj, qj);
                m1 *= one1 ? pj : fma(al1, dj, qj);
            }
            __syncwarp();
            if ((base & 24) == 24) {
                renorm(m0, x0);
                renorm(m1, x1);
            }
        }
        auto canon = [&](int i, double al, double m, int x) {
            MleKey k;
            const bool dead = (i < 0) || a.both_zero || (al == 0.0 && a.pf_has_zero) || (al == 1.0 && a.pn_has_zero);
            const bool alive = !dead && m > 0.0;   // zero / NaN products: log = -inf / NaN -> -inf (Occupancy.py:112-114)
            if (alive && m < 2.2250738585072014e-308) {   // subnormal: make it normal first
                m *= 18446744073709551616.0;             // 2^64
                x -= 64;
            }
            const long long bits = __double_as_longlong(m);
            const int e2 = (int)((bits >> 52) & 0x7ff);
            k.e = alive ? x + e2 - 1023 : INT_MIN;
            k.m = alive ? __longlong_as_double((bits & 0x800fffffffffffffLL) | 0x3ff0000000000000LL) : 0.0;
            return k;
        };
        k0 = canon(i0, al0, m0, x0);
        k1 = canon(i1, al1, m1, x1);
    };
    // first maximum over the group: larger key, ties to the smaller grid index
    auto group_best = [&](MleKey &bk, int &bi) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            MleKey ok;
            ok.e = __shfl_xor_sync(NB_FULL, bk.e, o);
            ok.m = __shfl_xor_sync(NB_FULL, bk.m, o);
            const int oi = __shfl_xor_sync(NB_FULL, bi, o);
            if (key_gt(ok, bk) || (ok.e == bk.e && ok.m == bk.m && oi < bi)) {
                bk = ok;
                bi = oi;
            }
        }
    };
    auto take = [&](MleKey &bk, int &bi, const MleKey &k, int i) {   // i >= 0 evaluated
        if (i >= 0 && (bi < 0 || key_gt(k, bk) || (k.e == bk.e && k.m == bk.m && i < bi))) {
            bk = k;
            bi = i;
        }
    };
    const int BIG = 1 << 30;
    // ---- round 1: coarse points c_j = (j * last) / 15, j = r and r + 8
    const int j0c = r, j1c = r + 8;
    const int ci0 = (j0c * last) / 15, ci1 = (j1c * last) / 15;
    MleKey ck0, ck1;
    eval2(ci0, ci1, ck0, ck1, true);
    MleKey bk;
    bk.e = INT_MIN;
    bk.m = 0.0;
    int bi = -1;
    take(bk, bi, ck0, ci0);
    take(bk, bi, ck1, ci1);
    if (bi < 0) bi = BIG;
    group_best(bk, bi);
    // coarse ordinal of the best point and the keys of its coarse neighbours (lane j & 7 holds ordinal j in slot j >> 3)
    int bj = (ci0 == bi) ? j0c : ((ci1 == bi) ? j1c : -1);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) bj = max(bj, __shfl_xor_sync(NB_FULL, bj, o));
    auto coarse_key = [&](int j, MleKey &k, int &idx) {   // every lane of the group calls it with the same j in [0, 15]
        const int src = (lane & 24) | (j & 7);
        const int e_lo = __shfl_sync(NB_FULL, ck0.e, src), e_hi = __shfl_sync(NB_FULL, ck1.e, src);
        const double m_lo = __shfl_sync(NB_FULL, ck0.m, src), m_hi = __shfl_sync(NB_FULL, ck1.m, src);
        k.e = (j >> 3) ? e_hi : e_lo;
        k.m = (j >> 3) ? m_hi : m_lo;
        idx = (j * last) / 15;
    };
    bool tie = false;
    const bool all_dead = (bk.e == INT_MIN);
    const int bjc = max(bj, 0);
    MleKey kl, kr;
    int il, ir;
    coarse_key(max(bjc - 1, 0), kl, il);
    coarse_key(min(bjc + 1, 15), kr, ir);
    int LO = (bjc > 0) ? il : 0, HI = (bjc < 15) ? ir : last;     // evaluated contiguous range after round 2: [LO, HI]
    if (!all_dead) {
        if (bjc > 0 && il != bi && key_close_or_above(kl, bk, MLS_TIE)) tie = true;
        if (bjc < 15 && ir != bi && key_close_or_above(kr, bk, MLS_TIE)) tie = true;
    }
    // ---- round 2: the unknown points of (LO, HI): every index strictly inside except the best coarse point itself
    // (coarse spacing <= 8: at most 14 of them)
    MleKey rk0, rk1;
    int ri0 = -1, ri1 = -1;
    {
        const int lo_u = LO + ((bjc > 0) ? 1 : 0), hi_u = HI - ((bjc < 15) ? 1 : 0);   // the ends are coarse points unless at the grid ends
        auto kth = [&](int k) {   // k-th unknown index, -1 past the end
            int idx = lo_u + k;
            if (idx >= bi) idx++;                       // skip the best coarse point
            return (idx <= hi_u) ? idx : -1;
        };
        ri0 = kth(r);
        ri1 = kth(r + 8);
        if (all_dead) ri0 = ri1 = -1;
        eval2(ri0, ri1, rk0, rk1, false);
        int nbi = bi;
        MleKey nbk = bk;
        take(nbk, nbi, rk0, ri0);
        take(nbk, nbi, rk1, ri1);
        group_best(nbk, nbi);
        bk = nbk;
        bi = nbi;
    }
    // ---- threshold = max * exp(-cutoff / 2), canonical
    MleKey th;
    th.e = INT_MIN;
    th.m = 0.0;
    bool none = all_dead || !(a.thr_m == a.thr_m);
    if (!none) {
        if (a.thr_zero) {
            th.e = INT_MIN + 1;
        } else {
            th.m = bk.m * a.thr_m;
            th.e = bk.e + a.thr_e;
            if (th.m >= 2.0) {
                th.m *= 0.5;
                th.e += 1;
            }
        }
    }
    auto passes = [&](const MleKey &k) { return !none && key_gt(k, th); };
    auto near_thr = [&](const MleKey &k, int i) { return i >= 0 && !none && !a.thr_zero && k.e != INT_MIN && key_close(k, th, MLS_TIE); };
    // pass / fail over everything evaluated so far: the smallest / largest passing index, and per coarse ordinal a bit
    int okmin = BIG, okmax = -1;
    auto note = [&](const MleKey &k, int i) {
        if (i >= 0 && passes(k)) {
            okmin = min(okmin, i);
            okmax = max(okmax, i);
        }
        if (near_thr(k, i)) tie = true;
    };
    note(ck0, ci0);
    note(ck1, ci1);
    note(rk0, ri0);
    note(rk1, ri1);
    const unsigned pass_lo = __ballot_sync(NB_FULL, passes(ck0)) >> (8 * g) & 0xffu;    // coarse ordinals 0..7
    const unsigned pass_hi = __ballot_sync(NB_FULL, passes(ck1)) >> (8 * g) & 0xffu;    // coarse ordinals 8..15
    const unsigned cpass = pass_lo | (pass_hi << 8);
    // ---- round 3: the coarse intervals in which the status changes.  Left: the last failing coarse point below the
    // evaluated range [LO, HI] (if LO itself passes); right: the first failing coarse point above it (if HI passes).
    int li0 = -1, li1 = -1;
    {
        int l_from = -1, l_to = -2, r_from = -1, r_to = -2;   // unknown index ranges [from, to]
        if (!none && bjc > 1) {
            const int jlo = bjc - 1;                     // ordinal of LO
            if ((cpass >> jlo) & 1) {                    // LO passes: look further left
                int jf = -1;
                for (int j = jlo - 1; j >= 0; j--)
                    if (!((cpass >> j) & 1)) {
                        jf = j;
                        break;
                    }
                if (jf >= 0) {
                    l_from = (jf * last) / 15 + 1;
                    l_to = ((jf + 1) * last) / 15 - 1;
                }
            }
        }
        if (!none && bjc < 14) {
            const int jhi = bjc + 1;
            if ((cpass >> jhi) & 1) {
                int jf = -1;
                for (int j = jhi + 1; j <= 15; j++)
                    if (!((cpass >> j) & 1)) {
                        jf = j;
                        break;
                    }
                if (jf >= 0) {
                    r_from = ((jf - 1) * last) / 15 + 1;
                    r_to = (jf * last) / 15 - 1;
                }
            }
        }
        // slots 0..7 (first grid point of a lane) serve the left interval, 8..15 the right one (each holds <= 7 points)
        li0 = (l_from + r <= l_to) ? l_from + r : -1;
        li1 = (r_from + r <= r_to) ? r_from + r : -1;
        if (__ballot_sync(NB_FULL, li0 >= 0 || li1 >= 0)) {   // some window of the warp has an interval left to resolve
            MleKey k0, k1;
            eval2(li0, li1, k0, k1, false);
            note(k0, li0);
            note(k1, li1);
        }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        okmin = min(okmin, __shfl_xor_sync(NB_FULL, okmin, o));
        okmax = max(okmax, __shfl_xor_sync(NB_FULL, okmax, o));
    }
    tie = (__ballot_sync(NB_FULL, tie && valid && n > 0) & gmask) != 0;
    // ---- windows with rounding-level ties: scan of the whole grid, the warp together (rare)
    if (__ballot_sync(NB_FULL, tie)) {
        MleKey fk;
        fk.e = INT_MIN;
        fk.m = 0.0;
        int fi = -1;
        for (int b0 = 0; b0 < NA; b0 += 16) {
            const int i0 = (b0 + r < NA) ? b0 + r : -1, i1 = (b0 + 8 + r < NA) ? b0 + 8 + r : -1;
            MleKey k0, k1;
            eval2(i0, i1, k0, k1, true);
            take(fk, fi, k0, i0);
            take(fk, fi, k1, i1);
        }
        if (fi < 0) fi = BIG;
        group_best(fk, fi);
        MleKey fth;
        fth.e = INT_MIN;
        fth.m = 0.0;
        const bool fnone = (fk.e == INT_MIN) || !(a.thr_m == a.thr_m);
        if (!fnone) {
            if (a.thr_zero) {
                fth.e = INT_MIN + 1;
            } else {
                fth.m = fk.m * a.thr_m;
                fth.e = fk.e + a.thr_e;
                if (fth.m >= 2.0) {
                    fth.m *= 0.5;
                    fth.e += 1;
                }
            }
        }
        int fmin = BIG, fmax = -1;
        for (int b0 = 0; b0 < NA; b0 += 16) {
            const int i0 = (b0 + r < NA) ? b0 + r : -1, i1 = (b0 + 8 + r < NA) ? b0 + 8 + r : -1;
            MleKey k0, k1;
            eval2(i0, i1, k0, k1, true);
            if (i0 >= 0 && !fnone && key_gt(k0, fth)) {
                fmin = min(fmin, i0);
                fmax = max(fmax, i0);
            }
            if (i1 >= 0 && !fnone && key_gt(k1, fth)) {
                fmin = min(fmin, i1);
                fmax = max(fmax, i1);
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            fmin = min(fmin, __shfl_xor_sync(NB_FULL, fmin, o));
            fmax = max(fmax, __shfl_xor_sync(NB_FULL, fmax, o));
        }
        if (tie) {
            bi = fi;
            okmin = fmin;
            okmax = fmax;
        }
    }
    double occ = nb_nan(), lo = nb_nan(), hi = nb_nan();
    if (n > 0 && okmax >= 0 && bi >= 0 && bi < NA) {  // Occupancy.py:141 `if sum(new_inserts)>0`
        occ = a.alphas[bi];
        lo = a.alphas[okmin];
        hi = a.alphas[okmax];
    }
    if (valid) {
        if (r == 0 && a.wv) {   // per-window values (the block smoother's input): window wi of chunk c at oo / step + c + wi
            const int64_t wo = oo / a.step + c + wi;
            a.wv[wo] = occ;
            a.wv[a.wv_stride + wo] = lo;
            a.wv[2 * a.wv_stride + wo] = hi;
        }
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
  }
}

// ---------------------------------------------------------------------------------------------
// The same guarded search with ONE THREAD PER WINDOW (the default).  In the group form above the search logic -- a few
// hundred instructions of index arithmetic, shuffles and votes per window -- is executed by all 32 lanes for 4 windows, and
// ncu shows it, not the fp64 pipe, bounding the kernel (19 % of the instructions were DFMA / DMUL).  Here a lane owns a
// window outright: it prepares the window's fragments once into its own column of a shared-memory cache, evaluates 16
// grid points per round as 16 independent product chains (one 16-byte load per fragment and round), and runs the search
// logic for itself -- no shuffles, no votes, no barriers; the logic is paid once per window by one lane.  The arithmetic
// per (window, grid point) is the group kernel's, operation for operation, so the results are bit-identical.
// Rounds: 0 coarse (16 evenly spaced points), 1 the unknown points between the neighbours of the best coarse point,
// 2 the two coarse intervals in which pass / fail changes; a window with a rounding-level tie scans the whole grid
// (stages 3.. for the maximum, then for the interval) while the other lanes of its warp wait.
#define MTW_THREADS 128
#ifndef MTW_CAP
#define MTW_CAP 48
#endif
// branch-free forms for the per-lane search logic (no short-circuit evaluation: the logic is straight-line selects)
__device__ __forceinline__ bool keyx_gt(int ae, double am, int be, double bm) { return (ae > be) | ((ae == be) & (am > bm)); }
__device__ __forceinline__ MleKey keyx_scale(const MleKey &k, double f)   // k * f for f in (0.5, 2), canonical; -inf stays -inf
{
    MleKey r;
    double m = k.m * f;
    int e = k.e;
    const bool lo = m < 1.0, hi = m >= 2.0;
    m = lo ? m * 2.0 : (hi ? m * 0.5 : m);
    e += hi ? 1 : (lo ? -1 : 0);
    r.e = (k.e == INT_MIN) ? INT_MIN : e;
    r.m = (k.e == INT_MIN) ? 0.0 : m;
    return r;
}
template <int LB>
__global__ void __launch_bounds__(MTW_THREADS, LB) k_occ_mle_tw(OccMleArgs a)
{
    extern __shared__ __align__(16) double sm_mle[];   // pn[upper], pf[upper], alphas[n_alpha] | cache (d, q)[MTW_CAP][MTW_THREADS]
    const int up2 = (a.upper + 1) & ~1, na2 = (a.n_alpha + 1) & ~1;
    double *s_pn = sm_mle, *s_pf = sm_mle + up2, *s_al = sm_mle + 2 * up2;
    double2 *cache = reinterpret_cast<double2 *>(sm_mle + 2 * up2 + na2) + threadIdx.x;   // entry f of this lane at cache[f * MTW_THREADS]
    for (int i = threadIdx.x; i < a.upper; i += MTW_THREADS) {
        s_pn[i] = a.pn[i];
        s_pf[i] = a.pf[i];
    }
    for (int i = threadIdx.x; i < a.n_alpha; i += MTW_THREADS) s_al[i] = a.alphas[i];
    __syncthreads();
    const int c = blockIdx.y;
    const int64_t oo = a.out_off[c];
    const int L = (int)(a.out_off[c + 1] - oo);
    const int nwin = (L - a.halfstep + a.step - 1) / a.step;
    const int wi = blockIdx.x * MTW_THREADS + threadIdx.x;
    if (wi >= nwin) return;
    const int32_t *cp = a.col_ptr + a.col_off[c];
    const int2 *en = a.ent + a.frag_off[c];
    const int NA = a.n_alpha, last = NA - 1;
    const int t = a.halfstep + wi * a.step;
    const int e0 = cp[t - a.flank + a.csc_pad], e1 = cp[t + a.flank + 1 + a.csc_pad];
    const int n = e1 - e0;
    double rSN = a.sn_nobias, rSF = a.sf_nobias;   // reciprocals of the window's normalisers
    if (a.use_bias) {
        const int64_t wo = oo / a.step + c + wi;
        rSN = a.wsn[wo];
        rSF = a.wsf[wo];
    }
    // fragment f: nuc_probs / sum, nfr_probs / sum (Occupancy.py:106-109), the pair scaled by a power of two so that
    // max(p, q) is in [1, 2) (exact, constant in alpha)
    auto prep = [&](int f, double &pv, double &qv) {
        const int sz = en[e0 + f].y;
        pv = s_pn[sz] * rSN;
        qv = s_pf[sz] * rSF;
        const long long mb = __double_as_longlong(fmax(pv, qv));
        const int e2 = (int)((mb >> 52) & 0x7ff);
        if (e2 > 0 && e2 < 0x7fe) {
            const double sc = __longlong_as_double((long long)(2046 - e2) << 52);
            pv *= sc;
            qv *= sc;
        }
    };
    auto renorm = [](double &m, int &x) {
        const long long bits = __double_as_longlong(m);
        const int e2 = (int)((bits >> 52) & 0x7ff);
        if (e2 != 0 && e2 != 0x7ff) {
            x += e2 - 1023;
            m = __longlong_as_double(bits - ((long long)(e2 - 1023) << 52));
        }
    };
    auto canon = [&](bool dead, double m, int x) {
        MleKey k;
        const bool alive = !dead & (m > 0.0);   // zero / NaN products: log = -inf / NaN -> -inf (Occupancy.py:112-114)
        const bool sub = m < 2.2250738585072014e-308;   // subnormal: make it normal first
        m = sub ? m * 18446744073709551616.0 : m;      // 2^64
        x = sub ? x - 64 : x;
        const long long bits = __double_as_longlong(m);
        const int e2 = (int)((bits >> 52) & 0x7ff);
        k.e = alive ? x + e2 - 1023 : INT_MIN;
        k.m = alive ? __longlong_as_double((bits & 0x800fffffffffffffLL) | 0x3ff0000000000000LL) : 0.0;
        return k;
    };
    // ---- the window's fragments, once: cache (p - q, q); the product of the p's is L(alpha = 1) (q + (p - q) would lose p << q)
    MleKey key_one;
    {
        double m1 = 1.0;
        int x1 = 0;
        for (int f0 = 0; f0 < n; f0 += 4) {   // four fragments at a time: their size loads are in flight together
            int sz[4];
#pragma unroll
            for (int u = 0; u < 4; u++) sz[u] = (f0 + u < n) ? en[e0 + f0 + u].y : 0;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int f = f0 + u;
                if (f < n) {
                    double pv = s_pn[sz[u]] * rSN, qv = s_pf[sz[u]] * rSF;
                    const long long mb = __double_as_longlong(fmax(pv, qv));
                    const int e2 = (int)((mb >> 52) & 0x7ff);
                    if (e2 > 0 && e2 < 0x7fe) {
                        const double sc = __longlong_as_double((long long)(2046 - e2) << 52);
                        pv *= sc;
                        qv *= sc;
                    }
                    if (f < MTW_CAP) cache[f * MTW_THREADS] = make_double2(pv - qv, qv);
                    m1 *= pv;
                    if ((f & 31) == 31) renorm(m1, x1);
                }
            }
        }
        key_one = canon(a.both_zero || a.pn_has_zero, m1, x1);
    }
    // ---- 16 grid points (index -1: none) -> canonical keys
    // ---- 16 grid points idx[] (-1: none) -> canonical keys Ke[], Km[] (straight-line code, used by every round below)
#define MTW_EVAL16()                                                                                                       \
    {                                                                                                                      \
        double al[16];                                                                                                     \
        _Pragma("unroll") for (int k = 0; k < 16; k++)                                                                     \
        {                                                                                                                  \
            al[k] = (idx[k] >= 0) ? s_al[idx[k]] : 0.5;                                                                    \
            Km[k] = 1.0;                                                                                                   \
            Ke[k] = 0;                                                                                                     \
        }                                                                                                                  \
        const int nc = min(n, MTW_CAP);                                                                                    \
        for (int f0 = 0; f0 < nc; f0 += 32) { /* 32 factors (each in (2^-7, 2) away from the grid ends), then exponents out */ \
            const int f1 = min(f0 + 32, nc);                                                                               \
            _Pragma("unroll 2") for (int f = f0; f < f1; f++)                                                              \
            {                                                                                                              \
                const double2 u = cache[f * MTW_THREADS];                                                                  \
                _Pragma("unroll") for (int k = 0; k < 16; k++) Km[k] *= fma(al[k], u.x, u.y);                              \
            }                                                                                                              \
            _Pragma("unroll") for (int k = 0; k < 16; k++) renorm(Km[k], Ke[k]);                                           \
        }                                                                                                                  \
        _Pragma("unroll 1") for (int f = MTW_CAP; f < n; f++) { /* more fragments than the cache holds: prepared again */  \
            double pv, qv;                                                                                                 \
            prep(f, pv, qv);                                                                                               \
            const double dv = pv - qv;                                                                                     \
            _Pragma("unroll") for (int k = 0; k < 16; k++) Km[k] *= fma(al[k], dv, qv);                                    \
            if ((f & 31) == 31) {                                                                                          \
                _Pragma("unroll") for (int k = 0; k < 16; k++) renorm(Km[k], Ke[k]);                                       \
            }                                                                                                              \
        }                                                                                                                  \
        _Pragma("unroll") for (int k = 0; k < 16; k++)                                                                     \
        {                                                                                                                  \
            const bool dead = (idx[k] < 0) || a.both_zero || (al[k] == 0.0 && a.pf_has_zero) || (al[k] == 1.0 && a.pn_has_zero); \
            const MleKey kk = (al[k] == 1.0 && idx[k] >= 0) ? key_one : canon(dead, Km[k], Ke[k]);                         \
            Ke[k] = kk.e;                                                                                                  \
            Km[k] = kk.m;                                                                                                  \
        }                                                                                                                  \
    }
    const int BIG = 1 << 30;
    auto mk = [](int e, double m) {
        MleKey k;
        k.e = e;
        k.m = m;
        return k;
    };
    MleKey bk = mk(INT_MIN, 0.0), th = mk(INT_MIN, 0.0), th_lo = th, th_hi = th;
    int bi = BIG, okmin = BIG, okmax = -1;
    bool tie = false, none = true;
    auto take = [&](int ke, double km, int i) {   // first maximum: larger key, ties to the smaller grid index
        const bool better = (i >= 0) & (keyx_gt(ke, km, bk.e, bk.m) | ((ke == bk.e) & (km == bk.m) & (i < bi)));
        bk.e = better ? ke : bk.e;
        bk.m = better ? km : bk.m;
        bi = better ? i : bi;
    };
    auto set_threshold = [&]() {   // max * exp(-cutoff / 2), canonical; nothing passes when the maximum is -inf or the cutoff NaN
        none = (bk.e == INT_MIN) || !(a.thr_m == a.thr_m);
        th = mk(INT_MIN, 0.0);
        if (!none) {
            if (a.thr_zero) {
                th.e = INT_MIN + 1;
            } else {
                th.m = bk.m * a.thr_m;
                th.e = bk.e + a.thr_e;
                if (th.m >= 2.0) {
                    th.m *= 0.5;
                    th.e += 1;
                }
            }
        }
        // a likelihood within MLS_TIE (relative) of the threshold is a rounding-level tie: band (th_lo, th_hi)
        const bool band = !none && !a.thr_zero;
        th_lo = band ? keyx_scale(th, 1.0 - MLS_TIE) : mk(INT_MAX, 0.0);
        th_hi = band ? keyx_scale(th, 1.0 + MLS_TIE) : mk(INT_MIN, 0.0);
    };
    auto note = [&](int ke, double km, int i, bool check_tie) {
        const bool p = (i >= 0) & !none & keyx_gt(ke, km, th.e, th.m);
        okmin = p ? min(okmin, i) : okmin;
        okmax = p ? max(okmax, i) : okmax;
        if (check_tie) tie = tie | ((i >= 0) & (ke != INT_MIN) & keyx_gt(ke, km, th_lo.e, th_lo.m) & keyx_gt(th_hi.e, th_hi.m, ke, km));
        return p;
    };
    if (n > 0) {
        int idx[16], Ke[16];
        double Km[16];
        // ---- round 0: coarse points c_k = (k * last) / 15
#pragma unroll
        for (int k = 0; k < 16; k++) idx[k] = (k * last) / 15;
        MTW_EVAL16();
        int cke[16];
        double ckm[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            cke[k] = Ke[k];
            ckm[k] = Km[k];
            take(Ke[k], Km[k], idx[k]);
        }
        int bj = 0;
#pragma unroll
        for (int k = 0; k < 16; k++)
            if (idx[k] == bi) bj = k;
        {   // a coarse neighbour of the best coarse point within MLS_TIE of it (or above): rounding-level tie
            const MleKey b_lo = keyx_scale(bk, 1.0 - MLS_TIE);
#pragma unroll
            for (int k = 0; k < 16; k++)
                tie = tie | (((k == bj - 1) | (k == bj + 1)) & (bk.e != INT_MIN) & (cke[k] != INT_MIN) & !keyx_gt(b_lo.e, b_lo.m, cke[k], ckm[k]));
        }
        const int LO = (bj > 0) ? ((bj - 1) * last) / 15 : 0, HI = (bj < 15) ? ((bj + 1) * last) / 15 : last;
        // ---- round 1: the unknown points strictly inside (LO, HI) but the best coarse point (coarse spacing <= 8: at most 14)
        {
            const int lo_u = LO + ((bj > 0) ? 1 : 0), hi_u = HI - ((bj < 15) ? 1 : 0);
            bool any = false;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                int i = lo_u + k;
                if (i >= bi) i++;                       // skip the best coarse point
                idx[k] = (i <= hi_u && bk.e != INT_MIN) ? i : -1;
                any = any || idx[k] >= 0;
            }
            if (any) {
                MTW_EVAL16();
#pragma unroll
                for (int k = 0; k < 16; k++) take(Ke[k], Km[k], idx[k]);
            }
            set_threshold();
            if (any) {
#pragma unroll
                for (int k = 0; k < 16; k++) note(Ke[k], Km[k], idx[k], true);
            }
        }
        unsigned cpass = 0;
#pragma unroll
        for (int k = 0; k < 16; k++)
            if (note(cke[k], ckm[k], (k * last) / 15, true)) cpass |= 1u << k;
        // ---- round 2: the coarse intervals in which pass / fail changes
        {
            int l_from = -1, l_to = -2, r_from = -1, r_to = -2;
            if (!none && bj > 1 && ((cpass >> (bj - 1)) & 1)) {          // LO passes: the last failing coarse point below it
                const unsigned failing = ~cpass & ((1u << (bj - 1)) - 1u);
                if (failing) {
                    const int jf = 31 - __clz(failing);
                    l_from = (jf * last) / 15 + 1;
                    l_to = ((jf + 1) * last) / 15 - 1;
                }
            }
            if (!none && bj < 14 && ((cpass >> (bj + 1)) & 1)) {         // HI passes: the first failing coarse point above it
                const unsigned failing = ~cpass & 0xffffu & ~((2u << (bj + 1)) - 1u);
                if (failing) {
                    const int jf = __ffs(failing) - 1;
                    r_from = ((jf - 1) * last) / 15 + 1;
                    r_to = (jf * last) / 15 - 1;
                }
            }
            bool any = false;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                idx[k] = (l_from + k <= l_to) ? l_from + k : -1;
                idx[8 + k] = (r_from + k <= r_to) ? r_from + k : -1;
                any = any || idx[k] >= 0 || idx[8 + k] >= 0;
            }
            if (any) {
                MTW_EVAL16();
#pragma unroll
                for (int k = 0; k < 16; k++) note(Ke[k], Km[k], idx[k], true);
            }
        }
        // ---- rounding-level tie somewhere decisive: scan the whole grid, one grid point at a time, twice (maximum, then
        // interval).  Rare, so small code matters more than speed here.
        if (tie) {
            auto eval1 = [&](int i) {
                const double al1 = s_al[i];
                double m = 1.0;
                int x = 0;
#pragma unroll 1
                for (int f = 0; f < n; f++) {
                    double dv, qv;
                    if (f < MTW_CAP) {
                        const double2 u = cache[f * MTW_THREADS];
                        dv = u.x;
                        qv = u.y;
                    } else {
                        double pv;
                        prep(f, pv, qv);
                        dv = pv - qv;
                    }
                    m *= fma(al1, dv, qv);
                    if ((f & 31) == 31) renorm(m, x);
                }
                const bool dead = a.both_zero || (al1 == 0.0 && a.pf_has_zero) || (al1 == 1.0 && a.pn_has_zero);
                return (al1 == 1.0) ? key_one : canon(dead, m, x);
            };
            bk = mk(INT_MIN, 0.0);
            bi = BIG;
#pragma unroll 1
            for (int i = 0; i < NA; i++) {
                const MleKey k1 = eval1(i);
                take(k1.e, k1.m, i);
            }
            set_threshold();
            okmin = BIG;
            okmax = -1;
#pragma unroll 1
            for (int i = 0; i < NA; i++) {
                const MleKey k1 = eval1(i);
                note(k1.e, k1.m, i, false);
            }
        }
    }
#undef MTW_EVAL16
    double occ = nb_nan(), lo = nb_nan(), hi = nb_nan();
    if (n > 0 && okmax >= 0 && bi >= 0 && bi < NA) {  // Occupancy.py:141 `if sum(new_inserts)>0`
        occ = s_al[bi];
        lo = s_al[okmin];
        hi = s_al[okmax];
    }
    if (a.wv) {   // per-window values (the block smoother's input): window wi of chunk c at oo / step + c + wi
        const int64_t wo = oo / a.step + c + wi;
        a.wv[wo] = occ;
        a.wv[a.wv_stride + wo] = lo;
        a.wv[2 * a.wv_stride + wo] = hi;
    }
    const int left = t - a.halfstep, right = min(t + a.halfstep + 1, L);
    for (int x = left; x < right; x++) {
        a.vals[oo + x] = occ;
        a.lower[oo + x] = lo;
        a.upper_b[oo + x] = hi;
    }
    if (wi == nwin - 1)  // positions past the last window stay NaN (np.ones(n)*nan, Occupancy.py:133-135)
        for (int x = right; x < L; x++) {
            a.vals[oo + x] = nb_nan();
            a.lower[oo + x] = nb_nan();
            a.upper_b[oo + x] = nb_nan();
        }
}

// ---------------------------------------------------------------------------------------------
// OccupancyTrack.makeSmoothed (Occupancy.py:147-153 -> pyatac/utils.py:23-52, mode 'same', norm) on the per-window values.
// The unsmoothed tracks are constant over blocks of `step` positions (block k = window k covers [k S, k S + S), Occupancy.py
// :142-146), so the 2 flank + 1 taps of an output collapse to ~(2 flank + 1) / S + 1 block taps:
//     out[n] = sum_b V[b] T[n - b S + h] / sum_b I[b] T[n - b S + h],     T[i] = sum_{t < S} w[i - t]
// with V[b] the window's value (0 where it is NaN), I[b] its presence (0 for NaN windows and for blocks off the chunk, which
// is utils.smooth's zero padding) and h = (wlen - 1) / 2.  A thread owns the S outputs of one block and the three tracks
// (they are NaN together); block values are staged in shared memory as (V0, V1), (V2, I) pairs.  The only case the block
// form does not cover -- a last window cut short by the chunk end -- goes tap by tap for the outputs that reach it.
#define SB_THREADS 128
template <int S>
__global__ void __launch_bounds__(SB_THREADS) k_occ_smooth_blocks(const double *__restrict__ wv, int64_t wv_stride,
                                                                  const int64_t *__restrict__ out_off, const double *__restrict__ win,
                                                                  int wlen, int halfstep, double *__restrict__ o0, double *__restrict__ o1,
                                                                  double *__restrict__ o2)
{
    extern __shared__ __align__(16) double sm_sb[];
    __shared__ double s_den;                         // sum of the window, summed like k_smooth_same does
    const int h = (wlen - 1) / 2;
    const int R = (h + S - 1) / S;                   // block reach on either side
    const int nT = wlen + S - 1;
    const int padL = R * S - h;                      // T is stored from index -padL on: no bounds checks in the tap loop
    const int nTs = (2 * R + 1) * S;                 // stored entries: indices h - R S .. h + R S + S - 1
    const int nTp = (nTs + 1) & ~1;
    double *s_T = sm_sb;                             // [nTp], s_T[i + padL] = T[i]
    const int nWd = (wlen + 5) & ~1;
    double *s_wd = sm_sb + nTp;                      // [nWd] the taps in the tap-by-tap path's order, zero padded: s_wd[k] = w[h - (dlo + k)]
    double2 *s_vp = reinterpret_cast<double2 *>(sm_sb + nTp + nWd);  // [SB_THREADS + 2 R] (V0, V1) of every block in reach of the tile
    double2 *s_vq = s_vp + (SB_THREADS + 2 * R);               // [SB_THREADS + 2 R] (V2, I): two arrays, so a warp's 16-byte loads are contiguous
    int *s_cnt = reinterpret_cast<int *>(s_vq + (SB_THREADS + 2 * R));   // [SB_THREADS + 2 R + 1] prefix count of missing blocks
    const int c = blockIdx.y;
    const int64_t oo = out_off[c];
    const int L = (int)(out_off[c + 1] - oo);
    const int nwin = (L - halfstep + S - 1) / S;
    const int nblk = (L + S - 1) / S;                // blocks that hold an output
    const int bt0 = blockIdx.x * SB_THREADS;
    if (bt0 >= nblk) return;
    const int64_t wo0 = oo / S + c;                  // window 0 of this chunk
    if (threadIdx.x < 32) {
        double d = 0.0;
        for (int m = threadIdx.x; m < wlen; m += 32) d += win[m];
        d = warp_sum(d);
        if (threadIdx.x == 0) s_den = d;
    }
    for (int k = threadIdx.x; k < nTs; k += SB_THREADS) {
        const int i = k - padL;
        double t = 0.0;
#pragma unroll
        for (int u = 0; u < S; u++) {
            const int m = i - u;
            if (m >= 0 && m < wlen) t += win[m];
        }
        s_T[k] = t;
    }
    {
        const int dlo = -((wlen - h) & ~1);
        for (int k = threadIdx.x; k < nWd; k += SB_THREADS) {
            const int m = h - (dlo + k);
            s_wd[k] = (m >= 0 && m < wlen) ? win[m] : 0.0;
        }
    }
    for (int i = threadIdx.x; i < SB_THREADS + 2 * R; i += SB_THREADS) {
        const int b = bt0 - R + i;
        double v0 = 0.0, v1 = 0.0, v2 = 0.0, pres = 0.0;
        if (b >= 0 && b < nwin) {   // three independent loads (the bounds of a NaN window are NaN as well and are dropped with it)
            v0 = wv[wo0 + b];
            v1 = wv[wv_stride + wo0 + b];
            v2 = wv[2 * wv_stride + wo0 + b];
            if (v0 == v0)
                pres = 1.0;
            else
                v0 = v1 = v2 = 0.0;
        }
        s_vp[i] = make_double2(v0, v1);
        s_vq[i] = make_double2(v2, pres);
    }
    __syncthreads();
    if (threadIdx.x < 32) {   // s_cnt[i] = missing blocks among tile slots [0, i)
        int run = 0;
        for (int i0 = 0; i0 < SB_THREADS + 2 * R; i0 += 32) {
            const int i = i0 + threadIdx.x;
            const bool bad = i < SB_THREADS + 2 * R && s_vq[i].y == 0.0;
            const unsigned mk = __ballot_sync(NB_FULL, bad);
            if (i <= SB_THREADS + 2 * R) s_cnt[i] = run + __popc(mk & ((1u << threadIdx.x) - 1u));
            run += __popc(mk);
        }
        if (threadIdx.x == 0) s_cnt[SB_THREADS + 2 * R] = run;
    }
    __syncthreads();
    const int b0 = bt0 + threadIdx.x;
    if (b0 >= nblk) return;
    const int n0 = b0 * S;
    // Outputs whose window holds a missing value (a NaN window, the chunk edges) or reaches a last block cut short by the
    // chunk end go tap by tap, in the order and association of k_smooth_same (taps from dlo on in pairs, the odd tap of a
    // pair first), so that they are bit-identical to the tap-by-tap smoother: where few windows are present the smoothed
    // track has plateaus on which call_peaks' 1e-12 jitter (utils.py:94-97) decides, and the last bit must not depend
    // on which smoother ran.
    const bool partial = (int64_t)nwin * S > L;
    const bool slow = (partial && b0 + R >= nwin - 1) || s_cnt[threadIdx.x + 2 * R + 1] != s_cnt[threadIdx.x];   // a missing block in [b0 - R, b0 + R]
    if (slow) {
        const int dlo = -((wlen - h) & ~1);
        const int T2 = (h - dlo + 2) & ~1;
        const int nvalid = min(L, nwin * S);              // positions with a window value
        auto tapw = [&](int k) { return s_wd[k]; };       // wd[k] = w[h - (dlo + k)], zero padded
        for (int u = 0; u < S; u++) {
            const int n = n0 + u;
            if (n >= L) break;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, den = 0.0;
            bool miss = false;                            // a real tap of this output is missing
            for (int k = 0; k < T2; k += 2) {
                const double wx = tapw(k), wy = tapw(k + 1);
                double x0[2], x1[2], x2[2], pi[2];
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int j = n + dlo + k + e;
                    x0[e] = x1[e] = x2[e] = pi[e] = 0.0;
                    if (j >= 0 && j < nvalid) {
                        const int i = j / S - (bt0 - R);      // tile slot of the block
                        if (i >= 0 && i < SB_THREADS + 2 * R) {
                            const double2 p = s_vp[i], q = s_vq[i];
                            x0[e] = p.x;
                            x1[e] = p.y;
                            x2[e] = q.x;
                            pi[e] = q.y;
                        }
                    }
                }
                a0 = fma(wx, x0[0], fma(wy, x0[1], a0));
                a1 = fma(wx, x1[0], fma(wy, x1[1], a1));
                a2 = fma(wx, x2[0], fma(wy, x2[1], a2));
                den = fma(wx, pi[0], fma(wy, pi[1], den));
                miss = miss || (wx != 0.0 && pi[0] == 0.0) || (wy != 0.0 && pi[1] == 0.0);
            }
            if (!miss) den = s_den;                       // k_smooth_same's fast path divides by the window's sum
            o0[oo + n] = (den == 0.0) ? nb_nan() : a0 / den;
            o1[oo + n] = (den == 0.0) ? nb_nan() : a1 / den;
            o2[oo + n] = (den == 0.0) ? nb_nan() : a2 / den;
        }
        return;
    }
    // Every block in reach is present.  The sums run over the differences to the centre block's values: a stretch of
    // equal windows then smooths to exactly that value whatever the output's phase inside its block, like the tap-by-tap
    // sum (and numpy's) gives one constant there -- on such plateaus call_peaks' jitter alone must pick the maxima.
    double a0[S], a1[S], a2[S], dn[S];
#pragma unroll
    for (int u = 0; u < S; u++) a0[u] = a1[u] = a2[u] = dn[u] = 0.0;
    const double2 cp = s_vp[threadIdx.x + R], cq = s_vq[threadIdx.x + R];   // centre block (V0, V1), (V2, 1)
    // block b0 + d contributes to output u through T[u - d S + h]; d runs over [-R, R]
    for (int d = -R; d <= R; d++) {
        const int i = threadIdx.x + R + d;
        double2 p = s_vp[i], q = s_vq[i];
        p.x -= cp.x;
        p.y -= cp.y;
        q.x -= cq.x;
        const double *Tb = s_T + (R - d) * S;         // T[h - d S + u] at Tb[u]
#pragma unroll
        for (int u = 0; u < S; u++) {
            const double w = Tb[u];
            a0[u] = fma(w, p.x, a0[u]);
            a1[u] = fma(w, p.y, a1[u]);
            a2[u] = fma(w, q.x, a2[u]);
            dn[u] = fma(w, q.y, dn[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < S; u++) {
        const int n = n0 + u;
        if (n < L) {
            o0[oo + n] = cp.x + a0[u] / dn[u];   // dn = sum of the window over present taps > 0 here
            o1[oo + n] = cp.y + a1[u] / dn[u];
            o2[oo + n] = cq.x + a2[u] / dn[u];
        }
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
    int n_hist;                      // insert-size histograms in shared memory (getNucDist: that many peaks per round)
    double min_occ;
};

#define PK_THREADS 512
__global__ void __launch_bounds__(PK_THREADS) k_occ_peaks(OccPeakArgs a)
{
    extern __shared__ unsigned char sm_pk[];
    double *s_nd = reinterpret_cast<double *>(sm_pk);          // [upper]
    int *s_hist = reinterpret_cast<int *>(s_nd + a.upper);      // [n_hist][upper]
    __shared__ double red_d[32];
    __shared__ int red_i[32];
    __shared__ int s_cnt[32 * (PK_THREADS / 32)];
    __shared__ double s_tot[PK_THREADS / 32];
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
#ifdef PK_TIMING
    long long pk_t[8];
    pk_t[0] = clock64();
#endif
    // coverage: flat window over fragment centres = difference of the CSC prefix (tracks.py:209-222); in the same pass the
    // minimum and the NaN count for NaN -> min (utils.py:86-91).  One loop, four positions in flight per thread: the two
    // separate loops were a quarter of the kernel, all of it load latency.
    double mn = CUDART_INF;
    int nnan = 0;
    {
        const int32_t *__restrict__ cpr = cp + a.csc_pad;
        const double *__restrict__ svr = sv;
        double *__restrict__ covw = cov;
        const int stride = (int)blockDim.x;
        int x = tid;
        for (; x + 3 * stride < L; x += 4 * stride) {
            int hi4[4], lo4[4];
            double v4[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                hi4[u] = cpr[x + u * stride + a.flank + 1];
                lo4[u] = cpr[x + u * stride - a.flank];
                v4[u] = svr[x + u * stride];
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                covw[x + u * stride] = (double)(hi4[u] - lo4[u]);
                if (v4[u] != v4[u]) nnan++;
                else mn = fmin(mn, v4[u]);
            }
        }
        for (; x < L; x += stride) {
            covw[x] = (double)(cpr[x + a.flank + 1] - cpr[x - a.flank]);
            const double v = svr[x];
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
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
        mn = fmin(mn, red_d[w]);
        nnan += red_i[w];
    }
    __syncthreads();
    for (int i = tid; i < a.upper; i += blockDim.x) s_nd[i] = 0.0;
    if (tid == 0) s_base = 0;
    __syncthreads();
#ifdef PK_TIMING
    pk_t[1] = clock64();
#endif
    int m = 0;
    if (nnan < L) {
        if (nnan > 0)
            for (int x = tid; x < L; x += blockDim.x)
                if (sv[x] != sv[x]) sv[x] = mn;
        __syncthreads();
        // strict local maxima of sig*(1+jitter), order 1 (argrelmax, clip mode), utils.py:94-100
        const int boundary = a.sep / 2;
        const int lo = max(1, boundary), hi = min(L - 1, L - boundary);
        block_compact_ordered(
            L, s_cnt, red_i, &s_base,
            [&](int x) {
                if (x < lo || x >= hi) return false;
                const double v = sv[x];
                const double j0 = v * (1.0 + a.jitter[x]);
                const double jl = sv[x - 1] * (1.0 + a.jitter[x - 1]);
                const double jr = sv[x + 1] * (1.0 + a.jitter[x + 1]);
                return (j0 > jl) && (j0 > jr) && (v >= a.min_occ);
            },
            [&](int x, int slot) {
                cpos[slot] = x;
                cval[slot] = sv[x];
            });
        m = s_base;
#ifdef PK_TIMING
        pk_t[2] = clock64();
#endif
        block_nms(cpos, cval, cst, m, a.sep, &s_flag);
    }
    __syncthreads();
#ifdef PK_TIMING
    pk_t[3] = clock64();
#endif
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
#ifdef PK_TIMING
    pk_t[4] = clock64();
#endif
    if (tid == 0) a.peak_count[c] = (s_base <= cap) ? s_base : -s_base;  // negative: capacity exceeded
    // getNucDist, Occupancy.py:232-240: sum over peaks of the window's insert-size histogram / its total
    // n_hist peaks per round: a warp builds the histogram of one peak's window (a few dozen fragments), then the sizes are
    // accumulated peak by peak in the reference's order (a whole-block round per peak was half of this kernel's time)
    const int2 *en = a.ent + a.frag_off[c];
    const int wid = tid >> 5, lane = tid & 31, nwarp = (int)(blockDim.x >> 5);
    for (int k0 = 0; k0 < npk; k0 += a.n_hist) {
        const int nk = min(a.n_hist, npk - k0);
        for (int kk = wid; kk < nk; kk += nwarp) {
            int *h = s_hist + (size_t)kk * a.upper;
            for (int i = lane; i < a.upper; i += 32) h[i] = 0;
            __syncwarp();
            const int p = a.peak_pos[po + k0 + kk] - a.start[c];
            const int e0 = cp[p - a.flank + a.csc_pad], e1 = cp[p + a.flank + 1 + a.csc_pad];
            for (int e = e0 + lane; e < e1; e += 32) atomicAdd(&h[en[e].y], 1);
            if (lane == 0) s_tot[kk] = (double)(e1 - e0);
        }
        __syncthreads();
        for (int i = tid; i < a.upper; i += blockDim.x) {
            double acc = s_nd[i];
            for (int kk = 0; kk < nk; kk++) acc += (double)s_hist[(size_t)kk * a.upper + i] / s_tot[kk];
            s_nd[i] = acc;
        }
        __syncthreads();
    }
    for (int i = tid; i < a.upper; i += blockDim.x) a.nuc_dist[(int64_t)c * a.upper + i] = s_nd[i];
#ifdef PK_TIMING
    pk_t[5] = clock64();
    if (tid == 0 && (c % 500) == 7)
        printf("k_occ_peaks chunk %d: L %d m %d npk %d | cov+nan %lld maxima %lld nms %lld filter %lld nuc_dist %lld (clk)\n", c, L, m, npk,
               pk_t[1] - pk_t[0], pk_t[2] - pk_t[1], pk_t[3] - pk_t[2], pk_t[4] - pk_t[3], pk_t[5] - pk_t[4]);
#endif
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
    NB_CHECK(nb200_fifo_enter(ctx, b->stream));                          // passes run first-in first-out across batches
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
            a.thr_k = k;
            const char *epi = getenv("NB200_MLE_EPI");   // developer switch: "canonical" = always the (exponent, mantissa) epilogue
            a.fast_epi = (std::isfinite(k) && k >= 1e-30 && k <= 1e30 && !(epi && !strcmp(epi, "canonical"))) ? 1 : 0;
        }
        a.sn_nobias = 1.0 / (r.pn_sum * window);
        a.sf_nobias = 1.0 / (r.pf_sum * window);
        int max_win = (b->max_len - halfstep + p.step - 1) / p.step;
        if (max_win < 1) max_win = 1;
        size_t smem = sizeof(double) * 2 * (size_t)p.upper;
        {
            const size_t smem_staged = smem + sizeof(double) * 2 * MLE_WPB +
                                       sizeof(int) * ((((size_t)(MLE_WPB - 1) * p.step + 2 * (size_t)p.flank + 2 + 3) & ~(size_t)3) + MLE_CAPF);
            const char *st = getenv("NB200_MLE_STAGE");   // developer switch: "0" = every warp loads its own inputs from global memory
            a.stage = (smem_staged <= 40 * 1024 && !(st && !strcmp(st, "0"))) ? 1 : 0;   // unusually wide windows / steps: unstaged
            if (a.stage) smem = smem_staged;
        }
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
        // per-window values for the block smoother (3 slabs)
        a.wv_stride = (int64_t)(tl / p.step + n + 2);
        NB_CUDA(ctx, b->o_wv.reserve(sizeof(double) * 3 * (size_t)a.wv_stride));
        a.wv = b->o_wv.as<double>();
        // Three kernels return bit-identical grids.  The scan of all grid points (k_occ_mle) is the default: the guarded
        // searches do 48 instead of 104 evaluations per window, but their search logic costs more instructions than the
        // evaluations it saves (measured per 20 Mbp: scan 3.29 ms, one thread per window 3.61 ms, 8 lanes per window 3.72 ms).
        // NB200_MLE_SEARCH=tw | group selects a search (grids of 17..121 points: its 16 coarse points need a spacing <= 8).
        const char *mle_env = getenv("NB200_MLE_SEARCH");
        const bool mle_search = mle_env && (!strcmp(mle_env, "tw") || !strcmp(mle_env, "group")) && r.n_alpha >= 17 && r.n_alpha <= 121;
        const bool mle_group = mle_search && !strcmp(mle_env, "group");
        ProfScope ps(ctx, b->stream, mle_search ? (mle_group ? "k_occ_mle_search" : "k_occ_mle_tw") : "k_occ_mle");
        dim3 grid((unsigned)div_up64(max_win, MLE_WARPS * MLE_GROUPS * MLE_ITERS), n);   // 128 windows per block in every form
        const int mle_lb = getenv("NB200_MLE_LB") ? atoi(getenv("NB200_MLE_LB")) : MLE_LB_DEFAULT;   // read per call: tools/mle_ab.py switches forms inside one process
        const int mle_gl = getenv("NB200_MLE_GL") ? atoi(getenv("NB200_MLE_GL")) : MLE_GL_DEFAULT;   // lanes per window: 8 | 16
        if (mle_search && !mle_group) {   // one thread per window
            const size_t smem_t = sizeof(double) * (2 * (size_t)((p.upper + 1) & ~1) + (size_t)((r.n_alpha + 1) & ~1)) + sizeof(double2) * (size_t)MTW_CAP * MTW_THREADS;
            static const int mtw_lb = getenv("NB200_MTW_LB") ? atoi(getenv("NB200_MTW_LB")) : 2;
            auto kern = mtw_lb >= 3 ? k_occ_mle_tw<3> : (mtw_lb == 2 ? k_occ_mle_tw<2> : k_occ_mle_tw<1>);
            NB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
            dim3 tgrid((unsigned)div_up64(max_win, MTW_THREADS), n);
            kern<<<tgrid, MTW_THREADS, smem_t, b->stream>>>(a);
        } else if (mle_search) {
            const size_t smem_s = sizeof(double) * 2 * (size_t)((p.upper + 1) & ~1) + sizeof(double2) * (size_t)MLE_WARPS * MLE_GROUPS * MLS_GROUP_STRIDE;
            static const int mls_lb = getenv("NB200_MLS_LB") ? atoi(getenv("NB200_MLS_LB")) : 4;
            auto kern = mls_lb >= 6 ? k_occ_mle_search<6> : (mls_lb == 5 ? k_occ_mle_search<5> : k_occ_mle_search<4>);
            if (smem_s > 48 * 1024) NB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
            kern<<<grid, MLE_WARPS * 32, smem_s, b->stream>>>(a);
        } else if (mle_gl == 16) {   // 16 lanes per window
            if (r.n_alpha > 112)
                k_occ_mle<8, 6, 16, 2 * MLE_ITERS><<<grid, MLE_WARPS * 32, smem, b->stream>>>(a);
            else if (mle_lb >= 8)
                k_occ_mle<7, 8, 16, 2 * MLE_ITERS><<<grid, MLE_WARPS * 32, smem, b->stream>>>(a);
            else if (mle_lb == 7)
                k_occ_mle<7, 7, 16, 2 * MLE_ITERS><<<grid, MLE_WARPS * 32, smem, b->stream>>>(a);
            else
                k_occ_mle<7, 6, 16, 2 * MLE_ITERS><<<grid, MLE_WARPS * 32, smem, b->stream>>>(a);
        } else if (r.n_alpha <= 104) {
            if (mle_lb >= 5)
                k_occ_mle<13, 5, 8, MLE_ITERS><<<grid, MLE_WARPS * 32, smem, b->stream>>>(a);
            else
                k_occ_mle<13, 4, 8, MLE_ITERS><<<grid, MLE_WARPS * 32, smem, b->stream>>>(a);
        } else
            k_occ_mle<16, 4, 8, MLE_ITERS><<<grid, MLE_WARPS * 32, smem, b->stream>>>(a);
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
        const bool dense_smooth = getenv("NB200_OCC_SMOOTH_DENSE") != nullptr;   // developer switch: tap-by-tap smoothing
        if (p.step == 5 && !dense_smooth) {   // block form on the per-window values (the default step)
            const int R = ((p.smooth_len - 1) / 2 + 4) / 5;
            const size_t smem = sizeof(double) * ((size_t)(((2 * R + 1) * 5 + 1) & ~1) + (size_t)((p.smooth_len + 5) & ~1)) + sizeof(double2) * 2 * (size_t)(SB_THREADS + 2 * R) +
                                sizeof(int) * (size_t)(SB_THREADS + 2 * R + 4);
            if (smem > 48 * 1024) NB_CUDA(ctx, cudaFuncSetAttribute(k_occ_smooth_blocks<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ProfScope ps(ctx, b->stream, "k_occ_smooth_blocks");
            dim3 grid((unsigned)div_up64(div_up64(b->max_len, 5), SB_THREADS), n);
            k_occ_smooth_blocks<5><<<grid, SB_THREADS, smem, b->stream>>>(b->o_wv.as<double>(), (int64_t)(tl / p.step + n + 2), b->d_out_off.as<int64_t>(),
                                                                         r.occ_win.as<double>(), p.smooth_len, halfstep, tr.out[0], tr.out[1], tr.out[2]);
            NB_LAUNCH_CHECK(ctx);
        } else {
            size_t smem = smooth_same_smem(p.smooth_len);
            if (smem > 48 * 1024) NB_CUDA(ctx, cudaFuncSetAttribute(k_smooth_same, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ProfScope ps(ctx, b->stream, "k_smooth_same");
            dim3 grid((unsigned)div_up64(b->max_len, SM_TILE), n, 3);
            k_smooth_same<<<grid, SM_THREADS, smem, b->stream>>>(tr, b->d_out_off.as<int64_t>(), r.occ_win.as<double>(), p.smooth_len, 0);
            NB_LAUNCH_CHECK(ctx);
        }
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
        a.n_hist = std::max(1, std::min(PK_THREADS / 32, (int)((40 * 1024 - (size_t)p.upper * 8) / ((size_t)p.upper * 4))));
        size_t smem = (size_t)p.upper * 8 + (size_t)a.n_hist * p.upper * 4;
        ProfScope ps(ctx, b->stream, "k_occ_peaks");
        k_occ_peaks<<<n, PK_THREADS, smem, b->stream>>>(a);
        NB_LAUNCH_CHECK(ctx);
    }
    b->occ_upper = p.upper;
    NB_CHECK(nb200_fifo_leave(ctx, b->stream));
    b->occ_done = true;
    return NB200_OK;
}

static int d2h(nb200_ctx *ctx, nb200_dbatch *b, void *dst, const DevBuf &src, size_t bytes)
{
    if (!dst || !bytes) return NB200_OK;
    NB_CUDA(ctx, cudaMemcpyAsync(dst, src.p, bytes, cudaMemcpyDeviceToHost, b->copy_stream));
    return NB200_OK;
}

}  // extern "C"

// Shared body of nb200_occ_download / nb200_occ_download32: T = double copies the tracks as they are, T = float converts
// them on the device first (k_pack_f32 on the copy stream, after the pass) and copies the float32 slabs.
template <typename T, typename Out>
static int occ_download_impl(nb200_ctx *ctx, nb200_dbatch *b, const Out *o, const char *who)
{
    if (!ctx || !b || !o) return nb200_fail(ctx, NB200_ERR_ARG, "%s: NULL argument", who);
    if (!b->occ_done) return nb200_fail(ctx, NB200_ERR_STATE, "%s: nb200_occ_run has not been called on this batch", who);
    const size_t tl = (size_t)b->total_len, tb = sizeof(T) * tl;
    const int n = b->n_chunks;
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    NB_CUDA(ctx, cudaEventRecord(b->ev_pass, b->stream));            // copies start once the pass has finished ...
    NB_CUDA(ctx, cudaStreamWaitEvent(b->copy_stream, b->ev_pass, 0));
    struct Copied {                                                  // ... and the next pass over these arrays waits for them
        nb200_dbatch *b;
        ~Copied() { cudaEventRecord(b->ev_copied_occ, b->copy_stream); }
    } copied{b};
    T *host[7] = {o->smoothed_vals, o->smoothed_lower, o->smoothed_upper, o->vals, o->lower_bound, o->upper_bound, o->cov};
    const DevBuf *dev[7] = {&b->o_svals, &b->o_slower, &b->o_supper, &b->o_vals, &b->o_lower, &b->o_upper, &b->o_cov};
    if (sizeof(T) == sizeof(double)) {
        for (int t = 0; t < 7; t++) NB_CHECK(d2h(ctx, b, host[t], *dev[t], tb));
    } else {
        Pack32Args pa;
        int nt = 0, which[7];
        for (int t = 0; t < 7; t++)
            if (host[t]) which[nt++] = t;
        if (nt && tl) {
            const size_t slab = (tl + 3) & ~(size_t)3;
            NB_CUDA(ctx, b->pack32_occ.reserve(sizeof(float) * slab * nt));
            for (int k = 0; k < nt; k++) {
                pa.src[k] = dev[which[k]]->template as<double>();
                pa.dst[k] = b->pack32_occ.as<float>() + slab * k;
            }
            pa.n = (int64_t)tl;
            ProfScope ps(ctx, b->copy_stream, "k_pack_f32");
            k_pack_f32<<<dim3((unsigned)std::min<int64_t>(div_up64((int64_t)tl, 4 * 256), ctx->sm_count * 8), nt), 256, 0, b->copy_stream>>>(pa);
            NB_LAUNCH_CHECK(ctx);
        }
        for (int k = 0; k < nt; k++)
            NB_CUDA(ctx, cudaMemcpyAsync(host[which[k]], pa.dst[k], tb, cudaMemcpyDeviceToHost, b->copy_stream));
    }
    NB_CHECK(d2h(ctx, b, o->nuc_dist, b->o_nuc_dist, sizeof(double) * (size_t)n * ctx->occ.upper));
    NB_CHECK(d2h(ctx, b, o->peak_count, b->o_peak_count, sizeof(int32_t) * n));
    if (o->peak_pos) {
        if (!o->peak_off) return nb200_fail(ctx, NB200_ERR_ARG, "%s: peak_off is required with peak_pos", who);
        for (int c = 0; c <= n; c++)
            if (o->peak_off[c] != b->h_opeak_off[c])
                return nb200_fail(ctx, NB200_ERR_CAPACITY, "%s: peak_off must equal len/sep+2 capacities (chunk %d)", who, c);
        const size_t np = (size_t)b->h_opeak_off[n];
        NB_CHECK(d2h(ctx, b, o->peak_pos, b->o_peak_pos, sizeof(int32_t) * np));
        NB_CHECK(d2h(ctx, b, o->peak_occ, b->o_peak_occ, sizeof(double) * np));
        NB_CHECK(d2h(ctx, b, o->peak_lower, b->o_peak_lower, sizeof(double) * np));
        NB_CHECK(d2h(ctx, b, o->peak_upper, b->o_peak_upper, sizeof(double) * np));
        NB_CHECK(d2h(ctx, b, o->peak_reads, b->o_peak_reads, sizeof(double) * np));
    }
    return NB200_OK;
}

template <typename Out>
static int64_t occ_d2h_bytes_impl(nb200_dbatch *b, const Out *o, int64_t elem)
{
    if (!b || !o) return 0;
    const int64_t tb = elem * b->total_len;
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

extern "C" {

int nb200_occ_download(nb200_ctx *ctx, nb200_dbatch *b, const nb200_occ_out *o)
{
    return occ_download_impl<double>(ctx, b, o, "nb200_occ_download");
}
int nb200_occ_download32(nb200_ctx *ctx, nb200_dbatch *b, const nb200_occ_out32 *o)
{
    return occ_download_impl<float>(ctx, b, o, "nb200_occ_download32");
}
int64_t nb200_occ_d2h_bytes(nb200_dbatch *b, const nb200_occ_out *o) { return occ_d2h_bytes_impl(b, o, 8); }
int64_t nb200_occ_d2h_bytes32(nb200_dbatch *b, const nb200_occ_out32 *o) { return occ_d2h_bytes_impl(b, o, 4); }

}  // extern "C"
