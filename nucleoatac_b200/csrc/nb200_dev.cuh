// nb200_dev.cuh -- device-side helpers shared by the kernels of libnucleo_b200 (sm_100a).
#pragma once
#include <algorithm>

#include "nb200_common.cuh"

#define NB_FULL 0xffffffffu

__device__ __forceinline__ double nb_nan() { return __longlong_as_double(0x7ff8000000000000LL); }
__device__ __forceinline__ double nb_ninf() { return __longlong_as_double(0xfff0000000000000LL); }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(NB_FULL, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(NB_FULL, v, o);
    return v;
}
__device__ __forceinline__ int warp_min_i(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(NB_FULL, v, o));
    return v;
}
__device__ __forceinline__ int warp_max_i(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(NB_FULL, v, o));
    return v;
}

// block-wide sum of doubles; `red` is shared scratch of >= 32 doubles.  All threads get the result.
__device__ __forceinline__ double block_sum(double v, double *red)
{
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double r = (lane < nw) ? red[lane] : 0.0;
    r = warp_sum(r);
    return r;
}

// Cell of the (pre-normalisation) Tn5 bias matrix, pyatac/chunkmat2d.py:140-153 in the derived two-tap
// form: Bp[i, c] = exp(b[c-(i-1)//2] + b[c+i//2]) = E[l] * E[r]; i == 1 has a single tap.
// `Ec` points at E[c] (the caller guarantees the taps are inside the uploaded track).
__device__ __forceinline__ double bias_cell(const double *Ec, int i)
{
    if (i == 1) return Ec[0];
    int a = (i - 1) >> 1;  // arithmetic shift == Python floor division, (0-1)//2 = -1
    int b = i >> 1;
    return Ec[-a] * Ec[b];
}

// ---------------------------------------------------------------------------------------------
// Per-column sums of the bias matrix against NW weight vectors over the insert sizes [lo, hi):
//     out_t[col] = sum_i wt_t[i] * Bp[i, col]                 (the operands of every sliding window sum of BiasMat2D:
//                                                              Occupancy.py:136-140, tracks.py:209-222 on chunkmat2d.py:140-156)
// Sizes 2j+1 and 2j+2 share their left tap, so a column costs sum_j E[c-j] (w[2j+1] E[c+j] + w[2j+2] E[c+j+1]);
// size 0 has the taps of size 2, size 1 a single tap.  A thread owns two adjacent columns and walks j two steps at a
// time: both tap windows slide as aligned pairs (one 16-byte shared load per side for 8 cells), the paired weights are
// broadcast loads.  Block = PC_THREADS threads = 2 * PC_THREADS columns of chunk blockIdx.y.  Every weight vector t has its
// own column range [start - pad_t, end + pad_t), whose first column lands at out_t[out_off[c] + 2 * pad_t * c]; the tile
// grid runs over the widest range, so one pass serves occupancy (pad = flank) and nucleosome calling (pad = w) together.
#ifndef PC_THREADS
#define PC_THREADS 128
#endif
#ifndef PC_SHARED_PRODUCTS
#define PC_SHARED_PRODUCTS 1
#endif
template <int NW>
struct PairColsumArgs {
    const int32_t *start;
    const int64_t *out_off, *bias_off;
    const int32_t *seq_start;
    const double *E;
    const double *wt[NW];
    double *out[NW];
    int pwm_up;
    int lo[NW], hi[NW], pad[NW];   // per weight vector: its insert sizes [lo, hi) and the pad of its output columns
};

template <int NW>
static __global__ void __launch_bounds__(PC_THREADS) k_pair_colsums(PairColsumArgs<NW> a)
{
    extern __shared__ __align__(16) double sm_pc[];
    int hmax = a.hi[0], pmax = a.pad[0];
#pragma unroll
    for (int t = 1; t < NW; t++) {
        hmax = max(hmax, a.hi[t]);
        pmax = max(pmax, a.pad[t]);
    }
    const int J2 = (max(1, hmax / 2) + 1) & ~1;
    const int off0 = J2 + 2, nE = off0 + 2 * PC_THREADS + J2 + 4;
    double2 *s_wp = reinterpret_cast<double2 *>(sm_pc);            // [NW][J2] (w[2j+1], w[2j+2])
    double *s_E = sm_pc + 2 * (size_t)NW * J2;                     // element off0 + t is E at the tile's column t
    const int c = blockIdx.y;
    const int L = (int)(a.out_off[c + 1] - a.out_off[c]);
    const int ncol = L + 2 * pmax;
    const int col0 = blockIdx.x * (2 * PC_THREADS);
    if (col0 >= ncol) return;
    double w1[NW];
#pragma unroll
    for (int t = 0; t < NW; t++) {
        const double *w = a.wt[t];
        const int lo = a.lo[t], hi = a.hi[t];
        auto W = [&](int i) { return (i >= lo && i < hi) ? w[i] : 0.0; };
        for (int j = threadIdx.x; j < J2; j += PC_THREADS)
            s_wp[t * J2 + j] = make_double2(j == 0 ? 0.0 : W(2 * j + 1), W(2 * j + 2) + (j == 0 ? W(0) : 0.0));
        w1[t] = W(1);
    }
    const int64_t e_lo = a.bias_off[c], e_hi = a.bias_off[c + 1];
    const int64_t eb = e_lo - (int64_t)(a.seq_start[c] + a.pwm_up) + (a.start[c] - pmax + col0 - off0);
    for (int i = threadIdx.x; i < nE; i += PC_THREADS) {   // the last elements of reach only meet zero weights and may lie off the track
        const int64_t idx = eb + i;
        s_E[i] = (idx >= e_lo && idx < e_hi) ? a.E[idx] : 0.0;
    }
    __syncthreads();
    const int col = col0 + 2 * threadIdx.x;
    if (col >= ncol) return;
    const double *Ec = s_E + off0 + 2 * threadIdx.x;
    double2 Lc = *reinterpret_cast<const double2 *>(Ec), Rc = Lc;   // (E[c-j], E[c-j+1]) and (E[c+j], E[c+j+1]) at j = 0
    double s0[NW], s1[NW];
#pragma unroll
    for (int t = 0; t < NW; t++) {
        s0[t] = w1[t] * Lc.x;
        s1[t] = w1[t] * Lc.y;
    }
#pragma unroll 2
    for (int j = 0; j < J2; j += 2) {
        const double2 Ln = *reinterpret_cast<const double2 *>(Ec - j - 2);   // (E[c-j-2], E[c-j-1])
        const double2 Rn = *reinterpret_cast<const double2 *>(Ec + j + 2);   // (E[c+j+2], E[c+j+3])
#if PC_SHARED_PRODUCTS
        // the two-tap products E[l] E[r] of the four cells (column, step) are the same for every weight vector: 8 multiplies +
        // 8 FMAs per weight instead of 12 operations per weight (NW = 3: 32 instead of 36); also the reference's own order
        // (Bp = E[l] E[r] first, chunkmat2d.py:140-153, then the weighted sum)
        const double p00 = Lc.x * Rc.x, p01 = Lc.x * Rc.y;   // column c,   step j:   sizes 2j+1, 2j+2
        const double p10 = Lc.y * Rc.y, p11 = Lc.y * Rn.x;   // column c+1, step j
        const double q00 = Ln.y * Rc.y, q01 = Ln.y * Rn.x;   // column c,   step j+1
        const double q10 = Lc.x * Rn.x, q11 = Lc.x * Rn.y;   // column c+1, step j+1
#pragma unroll
        for (int t = 0; t < NW; t++) {
            const double2 wa = s_wp[t * J2 + j], wb = s_wp[t * J2 + j + 1];
            s0[t] = fma(wa.y, p01, fma(wa.x, p00, s0[t]));
            s1[t] = fma(wa.y, p11, fma(wa.x, p10, s1[t]));
            s0[t] = fma(wb.y, q01, fma(wb.x, q00, s0[t]));
            s1[t] = fma(wb.y, q11, fma(wb.x, q10, s1[t]));
        }
#else
#pragma unroll
        for (int t = 0; t < NW; t++) {
            const double2 wa = s_wp[t * J2 + j], wb = s_wp[t * J2 + j + 1];
            s0[t] = fma(Lc.x, fma(wa.x, Rc.x, wa.y * Rc.y), s0[t]);   // column c,   step j
            s1[t] = fma(Lc.y, fma(wa.x, Rc.y, wa.y * Rn.x), s1[t]);   // column c+1, step j
            s0[t] = fma(Ln.y, fma(wb.x, Rc.y, wb.y * Rn.x), s0[t]);   // column c,   step j+1
            s1[t] = fma(Lc.x, fma(wb.x, Rn.x, wb.y * Rn.y), s1[t]);   // column c+1, step j+1
        }
#endif
        Lc = Ln;
        Rc = Rn;
    }
#pragma unroll
    for (int t = 0; t < NW; t++) {
        const int ct = col - (pmax - a.pad[t]), nt = L + 2 * a.pad[t];   // column within this weight's own range
        const int64_t o = a.out_off[c] + 2 * (int64_t)a.pad[t] * c + ct;
        if (ct >= 0 && ct < nt) a.out[t][o] = s0[t];
        if (ct + 1 >= 0 && ct + 1 < nt) a.out[t][o + 1] = s1[t];
    }
}

template <int NW>
static inline size_t pair_colsums_smem(int hi)
{
    const int J2 = (std::max(1, hi / 2) + 1) & ~1;
    return sizeof(double) * (2 * (size_t)NW * J2 + (J2 + 2) + 2 * PC_THREADS + J2 + 4);
}

// ATAC shift + centre of a fragment, pyatac/fragments.pyx:26-36.  Returns false when the row is outside
// [0, upper - lower).
__device__ __forceinline__ void frag_geometry(int pos, int tlen, int atac, int &l_pos, int &ilen)
{
    int t = tlen < 0 ? -tlen : tlen;
    if (atac) {
        l_pos = pos + 4;
        ilen = t - 8;
    } else {
        l_pos = pos;
        ilen = t;
    }
}
__device__ __forceinline__ int floordiv2(int a) { return a >> 1; }

// ---------------------------------------------------------------------------------------------
// Greedy non-maximum suppression of pyatac/utils.py:56-78 (reduce_peaks), run by one thread block
// on m candidates sorted by position.  Equivalent to walking the candidates by descending score:
// a candidate that outranks every undecided neighbour within `sep` is kept, then its neighbours are
// excluded; repeat.  Ties: the later candidate outranks (np.argsort stable order walked backwards).
// state: 0 undecided, 1 kept, 2 excluded.  NaN scores rank lowest (np.argsort puts NaN last -> the
// reference would walk them FIRST; callers never pass NaN scores).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool nms_outranks(double va, int ia, double vb, int ib)
{
    return (va > vb) || (va == vb && ia > ib);
}

static __device__ void block_nms(const int *pos, const double *val, unsigned char *state, int m, int sep, int *flag_sh)
{
    for (int j = threadIdx.x; j < m; j += blockDim.x) state[j] = 0;
    __syncthreads();
    for (int it = 0; it < m + 1; it++) {
        if (threadIdx.x == 0) *flag_sh = 0;
        __syncthreads();
        // phase 1: winners
        for (int j = threadIdx.x; j < m; j += blockDim.x) {
            if (state[j] != 0) continue;
            bool win = true;
            for (int k = j - 1; k >= 0 && pos[j] - pos[k] < sep && win; k--)
                if ((state[k] == 0 || state[k] == 3) && nms_outranks(val[k], k, val[j], j)) win = false;
            for (int k = j + 1; k < m && pos[k] - pos[j] < sep && win; k++)
                if ((state[k] == 0 || state[k] == 3) && nms_outranks(val[k], k, val[j], j)) win = false;
            if (win) state[j] = 3;  // provisional keep; 3 still counts as undecided for the others in this phase
        }
        __syncthreads();
        // phase 2: exclusion around winners
        for (int j = threadIdx.x; j < m; j += blockDim.x) {
            if (state[j] != 0) continue;
            bool ex = false;
            for (int k = j - 1; k >= 0 && pos[j] - pos[k] < sep && !ex; k--)
                if (state[k] == 3) ex = true;
            for (int k = j + 1; k < m && pos[k] - pos[j] < sep && !ex; k++)
                if (state[k] == 3) ex = true;
            if (ex) state[j] = 4;  // provisional exclude
        }
        __syncthreads();
        for (int j = threadIdx.x; j < m; j += blockDim.x) {
            if (state[j] == 3) state[j] = 1;
            else if (state[j] == 4) state[j] = 2;
            if (state[j] == 0) *flag_sh = 1;
        }
        __syncthreads();
        int more = *flag_sh;
        __syncthreads();
        if (!more) break;
    }
}

// ---------------------------------------------------------------------------------------------
// Ordered block compaction: every thread passes flag (0/1) for the element it holds in this round
// (elements are visited in rounds of blockDim.x consecutive indices); returns the output slot of a
// flagged element and advances *base_sh (shared) by the round's total.  red: shared int[32].
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_compact_slot(int flag, int *base_sh, int *red)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    unsigned m = __ballot_sync(NB_FULL, flag);
    int within = __popc(m & ((1u << lane) - 1));
    if (lane == 0) red[wid] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < nw; w++) {
        int v = red[w];
        if (w < wid) before += v;
        total += v;
    }
    int base = *base_sh;
    __syncthreads();
    if (threadIdx.x == 0) *base_sh = base + total;
    __syncthreads();
    return base + before + within;
}

// ---------------------------------------------------------------------------------------------
// Ordered block compaction of the positions x in [0, L) with flag_of(x) != 0, 32 rounds of blockDim.x positions at a time:
// the per-warp counts of every round first (no block barrier between rounds, so their loads overlap), ONE block scan over
// the (round, warp) counts, then emit(x, slot) in position order.  block_compact_slot per round costs three barriers per
// blockDim.x positions, which was a sixth of k_occ_peaks / k_nuc_peaks.  s_cnt: shared int[32 * warps], red: shared int[32],
// *base_sh: running output count (shared).  blockDim.x <= 1024, a multiple of 32.
// ---------------------------------------------------------------------------------------------
template <typename FlagFn, typename EmitFn>
__device__ __forceinline__ void block_compact_ordered(int L, int *s_cnt, int *red, int *base_sh, FlagFn flag_of, EmitFn emit)
{
    const int tid = threadIdx.x, nthr = (int)blockDim.x, nwarp = nthr >> 5, wid = tid >> 5, lane = tid & 31;
    for (int g0 = 0; g0 < L; g0 += 32 * nthr) {
        const int nr = min(32, (L - g0 + nthr - 1) / nthr);
        unsigned mask = 0;
        for (int r = 0; r < nr; r++) {
            const int x = g0 + r * nthr + tid;
            const int flag = (x < L) ? (flag_of(x) ? 1 : 0) : 0;
            const unsigned bal = __ballot_sync(NB_FULL, flag);
            if (flag) mask |= 1u << r;
            if (lane == 0) s_cnt[r * nwarp + wid] = __popc(bal);
        }
        __syncthreads();
        const int n_ent = nr * nwarp;   // entries in (round, warp) = position order
        int run = 0;                    // entries before this pass of the scan
        for (int e0 = 0; e0 < n_ent; e0 += nthr) {
            const int cnt = e0 + tid < n_ent ? s_cnt[e0 + tid] : 0;
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(NB_FULL, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) red[wid] = incl;
            __syncthreads();
            int woff = 0, total = 0;
            for (int w = 0; w < nwarp; w++) {
                const int t = red[w];
                if (w < wid) woff += t;
                total += t;
            }
            __syncthreads();
            if (e0 + tid < n_ent) s_cnt[e0 + tid] = run + woff + incl - cnt;   // exclusive prefix inside this group of rounds
            run += total;
        }
        const int base = *base_sh;
        __syncthreads();
        if (tid == 0) *base_sh = base + run;
        for (int r = 0; r < nr; r++) {
            const int f = (mask >> r) & 1;
            const unsigned bal = __ballot_sync(NB_FULL, f);
            if (f) emit(g0 + r * nthr + tid, base + s_cnt[r * nwarp + wid] + __popc(bal & ((1u << lane) - 1)));
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// pyatac/utils.py:23-52 smooth(mode='same', norm=True) on packed tracks: out[n] = sum_m w[m]*x[n+h-m]
// over non-NaN x (zero padded), divided by the same sum over the non-NaN indicator; 0 -> NaN.
// clip_neg: values < 0 are taken as 0 first (NucChunk.smoothSignal, NucleosomeCalling.py:278-280).
// grid (tiles, chunks, tracks); block SM_TILE threads; dynamic smem (SM_TILE + 2*wlen) doubles.
// ---------------------------------------------------------------------------------------------
#ifndef SM_TILE
#define SM_TILE 512
#endif
#define SM_XT 4
#define SM_THREADS (SM_TILE / SM_XT)
struct SmoothTracks {
    const double *in[3];
    double *out[3];
};
// out[n] = sum_d wd[d] x[n + d], wd[d] = w[h - d].  Each thread owns SM_XT = 4 consecutive outputs and walks the taps two
// at a time: the inputs of (n .. n+3) x (d, d+1) are three aligned pairs that slide by one pair per step, so a step is one
// 16-byte load of x, one broadcast 16-byte load of the tap pair and 8 FMAs.  Tiles without NaN / zero padding take this
// fast path (denominator = sum of the window, as np.convolve of the window with an all-ones indicator gives).  A thread
// whose window holds a missing value (chunk edges, NaN stretches) runs the same walk on two arrays -- the values with the
// missing ones zeroed and a 1/0 presence indicator -- i.e. utils.smooth's two convolutions, 16 FMAs per step.
// The staged tile is stored with 16 bytes of padding after every 128 (SM_PHYS): a thread's pairs are 32 bytes apart, which
// would put lanes i and i + 4 of a quarter warp on the same banks; with the padding the 16-byte loads are conflict free.
#define SM_PHYS(j) ((j) + (((j) >> 4) << 1))
static inline size_t smooth_same_smem(int wlen)   // taps (padded) + staged tile + NaN prefix counts
{
    const size_t T2 = (size_t)wlen + 4, nX = SM_TILE + T2 + 4, nXp = nX + 2 * (nX / 16 + 1);   // staged tile incl. row padding
    return sizeof(double) * (T2 + 2 * nXp) + sizeof(int) * (nX + 2) + 16;   // taps, tile, presence indicator, NaN prefix counts
}
static __global__ void __launch_bounds__(SM_THREADS) k_smooth_same(SmoothTracks tr, const int64_t *__restrict__ out_off,
                                                                 const double *__restrict__ win, int wlen, int clip_neg)
{
    extern __shared__ __align__(16) double sm_s[];
    __shared__ int s_slow;
    __shared__ double s_den;
    const int h = (wlen - 1) / 2;
    const int dlo = -((wlen - h) & ~1);                  // first (even) tap offset: d in [h - wlen + 1, h] is covered by
    const int T2 = (h - dlo + 2) & ~1;                   //   T2 (even) taps from dlo on, zero padded
    const int nX = SM_TILE + T2 + 4;
    double *s_w = sm_s;                 // [T2] wd, zero padded
    double *s_x = sm_s + T2;            // [nX, padded] s_x[SM_PHYS(j)] = x[x0 + dlo + j]
    const int nXp = nX + 2 * (nX / 16 + 1);
    double *s_i = s_x + nXp;            // [nX, padded] 1.0 where s_x holds a value, 0.0 where it is missing (NaN / off the chunk)
    const int c = blockIdx.y;
    const int64_t o = out_off[c];
    const int L = (int)(out_off[c + 1] - o);
    const int x0 = blockIdx.x * SM_TILE;
    if (x0 >= L) return;
    const double *in = tr.in[blockIdx.z] + o;
    double *out = tr.out[blockIdx.z] + o;
    if (threadIdx.x == 0) s_slow = 0;
    for (int k = threadIdx.x; k < T2; k += blockDim.x) {
        const int m = h - (dlo + k);
        s_w[k] = (m >= 0 && m < wlen) ? win[m] : 0.0;
    }
    if (threadIdx.x < 32) {             // sum of the window
        double d = 0.0;
        for (int m = threadIdx.x; m < wlen; m += 32) d += win[m];
        d = warp_sum(d);
        if (threadIdx.x == 0) s_den = d;
    }
    __syncthreads();
    const int lo = x0 + dlo;
    const int n_end = min(x0 + SM_TILE, L);
    int slow = 0;
    for (int j = threadIdx.x; j < nX; j += blockDim.x) {
        const int idx = lo + j;
        double v = nb_nan();
        if (idx >= 0 && idx < L) {
            v = in[idx];
            if (clip_neg && v < 0) v = 0.0;
        }
        const bool missing = v != v;
        if (missing) {
            if (idx >= x0 + h - wlen + 1 && idx <= n_end - 1 + h) slow = 1;  // a real tap of this tile is missing
            v = 0.0;
        }
        s_x[SM_PHYS(j)] = v;
        s_i[SM_PHYS(j)] = missing ? 0.0 : 1.0;
    }
    if (slow) s_slow = 1;
    __syncthreads();
    // In a tile with missing taps (chunk edges, NaN stretches) only the threads whose own window holds one go tap by tap:
    // s_cnt[j] = number of NaN among s_x[0 .. j)
    bool my_slow = false;
    if (s_slow) {
        int *s_cnt = reinterpret_cast<int *>(s_i + nXp);
        if (threadIdx.x < 32) {
            int run = 0;
            for (int j0 = 0; j0 < nX; j0 += 32) {
                const int j = j0 + threadIdx.x;
                const bool bad = j < nX && s_i[SM_PHYS(j)] == 0.0;
                const unsigned m = __ballot_sync(NB_FULL, bad);
                if (j < nX) s_cnt[j] = run + __popc(m & ((1u << threadIdx.x) - 1u));
                run += __popc(m);
            }
            if (threadIdx.x == 0) s_cnt[nX] = run;
        }
        __syncthreads();
        const int a = threadIdx.x * SM_XT, b = min(a + T2 + SM_XT, nX);
        my_slow = s_cnt[b] != s_cnt[a];
    }
    const int n0 = x0 + threadIdx.x * SM_XT;
    if (n0 >= L) return;
    // out[n0 .. n0+3] of the correlation of `src` (padded tile layout) with the taps: one 16-byte load of the tile, one
    // broadcast 16-byte load of the tap pair and 8 FMAs per step
    auto walk = [&](const double *src, double &a0, double &a1, double &a2, double &a3) {
        const double2 *px = reinterpret_cast<const double2 *>(src);   // pair q of the tile sits at px[q + (q >> 3)]
        const double2 *pw = reinterpret_cast<const double2 *>(s_w);
        const int q0 = threadIdx.x * (SM_XT / 2);
        a0 = a1 = a2 = a3 = 0.0;
        double2 A = px[q0 + (q0 >> 3)], B = px[q0 + 1 + ((q0 + 1) >> 3)];
#pragma unroll 4
        for (int k2 = 0; k2 < T2 / 2; k2++) {
            const int q = q0 + k2 + 2;
            const double2 C = px[q + (q >> 3)], w = pw[k2];
            a0 = fma(w.x, A.x, fma(w.y, A.y, a0));
            a1 = fma(w.x, A.y, fma(w.y, B.x, a1));
            a2 = fma(w.x, B.x, fma(w.y, B.y, a2));
            a3 = fma(w.x, B.y, fma(w.y, C.x, a3));
            A = B;
            B = C;
        }
    };
    double a0, a1, a2, a3;
    walk(s_x, a0, a1, a2, a3);
    if (!my_slow) {
        const double den = s_den;
        if (n0 < L) out[n0] = a0 / den;
        if (n0 + 1 < L) out[n0 + 1] = a1 / den;
        if (n0 + 2 < L) out[n0 + 2] = a2 / den;
        if (n0 + 3 < L) out[n0 + 3] = a3 / den;
    } else {   // the same walk over the presence indicator gives the denominators; smoothed_norm == 0 -> NaN (utils.py:49)
        double d0, d1, d2, d3;
        walk(s_i, d0, d1, d2, d3);
        if (n0 < L) out[n0] = (d0 == 0.0) ? nb_nan() : a0 / d0;
        if (n0 + 1 < L) out[n0 + 1] = (d1 == 0.0) ? nb_nan() : a1 / d1;
        if (n0 + 2 < L) out[n0 + 2] = (d2 == 0.0) ? nb_nan() : a2 / d2;
        if (n0 + 3 < L) out[n0 + 3] = (d3 == 0.0) ? nb_nan() : a3 / d3;
    }
}

// ---------------------------------------------------------------------------------------------
// float64 -> float32 conversion of up to 8 packed tracks in one launch (the *_download32 staging): grid-stride over
// quads, 2 x 16-byte loads and one 16-byte store per thread and step; blockIdx.y = track.
// ---------------------------------------------------------------------------------------------
struct Pack32Args {
    const double *src[8];
    float *dst[8];
    int64_t n;
};
static __global__ void __launch_bounds__(256) k_pack_f32(Pack32Args a)
{
    const double *__restrict__ src = a.src[blockIdx.y];
    float *__restrict__ dst = a.dst[blockIdx.y];
    const int64_t nq = a.n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride) {
        const double2 u = reinterpret_cast<const double2 *>(src)[2 * q], v = reinterpret_cast<const double2 *>(src)[2 * q + 1];
        reinterpret_cast<float4 *>(dst)[q] = make_float4((float)u.x, (float)u.y, (float)v.x, (float)v.y);
    }
    if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) dst[4 * nq + threadIdx.x] = (float)src[4 * nq + threadIdx.x];
}

// ---------------------------------------------------------------------------------------------
// K2: InsertionBiasTrack.computeBias, pyatac/bias.py:85-92 + seq.py:37-45.
//   b[p] = sum_j logPWM[nuc(seq[p - up + j]), j]   (non-ACGT contributes 0: all-zero one-hot column)
// written as E[p] = exp(b[p]) and / or b[p].  grid (tiles, chunks); block 256; the sequence tile is
// staged in shared memory as PWM row codes.
// ---------------------------------------------------------------------------------------------
#define BT_TILE 1024
static __global__ void __launch_bounds__(256) k_bias_track(const uint8_t *__restrict__ seq, const int64_t *__restrict__ seq_off,
                                                           const int64_t *__restrict__ bias_off,
                                                           const double *__restrict__ log_pwm,
                                                           const int8_t *__restrict__ nuc_code, int n_nuc, int width,
                                                           double *__restrict__ E, double *__restrict__ b_out)
{
    __shared__ double s_pwm[NB200_MAX_NUC * NB200_MAX_PWM_WIDTH];
    __shared__ int8_t s_code[256];
    __shared__ int8_t s_seq[BT_TILE + NB200_MAX_PWM_WIDTH];
    const int c = blockIdx.y;
    const int64_t so = seq_off[c];
    const int64_t slen = seq_off[c + 1] - so;
    const int64_t blen = slen - (width - 1);
    const int64_t x0 = (int64_t)blockIdx.x * BT_TILE;
    if (x0 >= blen) return;
    for (int i = threadIdx.x; i < n_nuc * width; i += blockDim.x) s_pwm[i] = log_pwm[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_code[i] = nuc_code[i];
    __syncthreads();
    const int nload = (int)min((int64_t)(BT_TILE + width - 1), slen - x0);
    for (int i = threadIdx.x; i < BT_TILE + width + 3 && i < BT_TILE + NB200_MAX_PWM_WIDTH; i += blockDim.x)
        s_seq[i] = i < nload ? s_code[seq[so + x0 + i]] : (int8_t)-1;
    __syncthreads();
    const int n = (int)min((int64_t)BT_TILE, blen - x0);
    // four consecutive positions per thread: four independent chains of `width` dependent additions (each position still adds
    // its terms in ascending j), and four exp() whose instruction streams interleave
    for (int t4 = 4 * threadIdx.x; t4 < n; t4 += 4 * blockDim.x) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int j = 0; j < width; j++) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int code = s_seq[t4 + u + j];   // (positions past the end of the tile read the -1 padding and are not stored)
                if (code >= 0) acc[u] += s_pwm[code * width + j];
            }
        }
        const int64_t o = bias_off[c] + x0 + t4;
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (t4 + u < n) {
                if (E) E[o + u] = exp(acc[u]);
                if (b_out) b_out[o + u] = acc[u];
            }
    }
}
