// nb200_common.cuh -- shared declarations of libnucleo_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/nucleo_b200.h"

#define NB200_MAX_PWM_WIDTH 64
#define NB200_MAX_NUC 8
#define NB200_MAX_UPPER 2048
#define NB200_MAX_ALPHA 128

// ---------------------------------------------------------------------------------------------
// device buffer that only ever grows (no per-batch cudaMalloc on the steady-state path)
// ---------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct ProfEntry {
    std::string name;
    int64_t launches = 0;
    double ms = 0.0;
};

// run constants living on the device
struct RunConst {
    // PWM
    int pwm_up = 0, pwm_down = 0, pwm_width = 0, n_nuc = 0;
    bool have_pwm = false;
    DevBuf log_pwm;   // f64 [n_nuc][width]
    DevBuf nuc_code;  // int8 [256]: byte -> PWM row or -1
    // VMat
    bool have_vmat = false;
    int v_rows = 0, v_cols = 0, v_lower = 0, v_upper = 0, v_w = 0;
    int v_wpad = 0;   // v_cols rounded up to a multiple of 16 (zero padded rows of vmat_fp)
    bool v_has_zero = false;
    bool v_nonneg = false;    // every VMat entry is >= 0 and finite (precondition of the candidates' likelihood-ratio bound)
    DevBuf vmat;      // f64 [R][W]
    DevBuf vmat_fp;   // f64 [R][Wpad]  f_i * V, zero padded (operand of the dense background xcor)
    // "paired" templates of the candidate statistics (k_cand_stats): for T in {V, f*V, f*V^2}
    //   pair[t][j][k] = (T[2j+1,k], T[2j+2,k]) (j = 0: (0, T[0,k] + T[2,k])), one[t][k] = T[1,k]; zero outside [lower, upper),
    //   J2 x W2 (both even) double2 per template
    DevBuf vp_pair, vp_one;
    DevBuf vp_pair32, vp_one32;   // template V of the same layout in fp32 (k_cand_screen)
    int vp_J2 = 0, vp_W2 = 0;
    std::vector<double> h_vmat;
    // fragment sizes
    bool have_sizes = false;
    int sizes_upper = 0;
    bool f_has_zero = false;  // a zero frequency inside [v_lower, v_upper)
    double f_sum_v = 0.0;     // sum of f over [v_lower, v_upper)
    double f_max_v = 0.0;     // max of f over [v_lower, v_upper); < 0 when some f is negative / not finite
    DevBuf sizes;     // f64 [upper]
    std::vector<double> h_sizes;
    // occupancy model
    bool have_occ_model = false;
    int occ_upper = 0, n_alpha = 0;
    double cutoff = 0.0;
    int pn_has_zero = 0, pf_has_zero = 0, both_zero = 0;
    double pn_sum = 0.0, pf_sum = 0.0;
    DevBuf nuc_probs, nfr_probs, alphas;
    // jitter
    int64_t n_jitter = 0;
    DevBuf jitter;
    // smoothing windows
    DevBuf occ_win, nuc_win;
};

struct nb200_ctx {
    int device = 0;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    size_t hbm_bytes = 0;
    cudaStream_t stream = nullptr;  // stream of the primitive calls
    std::string err;
    RunConst rc;
    nb200_occ_params occ{};
    nb200_nuc_params nuc{};
    bool occ_configured = false, nuc_configured = false;
    int occ_gen = 0;   // bumped whenever something the occupancy column sums depend on changes (model, parameters, PWM)
    // scratch for the primitive (single-call) paths
    DevBuf s0, s1, s2, s3, s4;
    // profiling
    bool prof_on = false;
    std::vector<ProfEntry> prof;
    std::vector<cudaEvent_t> ev_pool;
    struct PendingEv {
        int idx;
        cudaEvent_t a, b;
    };
    std::vector<PendingEv> pending;
    DevBuf flush;  // L2 flush target
    // nccl
    void *nccl_lib = nullptr;
    void *nccl_comm = nullptr;
    int nccl_rank = 0, nccl_world = 1;
    void *tc_plan = nullptr;  // tensor-core xcor plan (nb200_xcor_tc.cu)
    // Passes of different batches run first-in first-out: every nb200_{occ,nuc}_run waits for the pass enqueued before it
    // (of whichever batch) and records this event at its end.  Batches keep their own streams for the H2D of their inputs
    // and the D2H of their results, which therefore overlap the passes of the neighbouring batches; without the chain the
    // passes of the batches in flight time-slice the SMs and all finish late together, so that no copy has anything left
    // to overlap with (measured: 26 ms per step end to end against 17 ms of compute and 15.6 ms of copies).
    cudaEvent_t ev_fifo = nullptr;
    bool fifo = true;
};
int nb200_fifo_enter(nb200_ctx *ctx, cudaStream_t st);   // order this pass after the previously enqueued one
int nb200_fifo_leave(nb200_ctx *ctx, cudaStream_t st);

int nb200_fail(nb200_ctx *ctx, int code, const char *fmt, ...);
int nb200_cuda_fail(nb200_ctx *ctx, cudaError_t e, const char *what, const char *file, int line);

#define NB_CUDA(ctx, call)                                                                    \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess) return nb200_cuda_fail((ctx), _e, #call, __FILE__, __LINE__);  \
    } while (0)

#define NB_CHECK(call)                 \
    do {                               \
        int _s = (call);               \
        if (_s != NB200_OK) return _s; \
    } while (0)

#define NB_LAUNCH_CHECK(ctx) NB_CUDA(ctx, cudaGetLastError())

// profiling bracket around a kernel launch
struct ProfScope {
    nb200_ctx *ctx;
    cudaStream_t st;
    int idx;
    cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(nb200_ctx *c, cudaStream_t s, const char *name);
    ~ProfScope();
};

void nb200_prof_collect(nb200_ctx *ctx);  // drain pending events (sync)

// ---------------------------------------------------------------------------------------------
// batch on the device.  Every packed per-position track of a batch has total_len values; chunk c
// occupies [out_off[c], out_off[c+1]).
// ---------------------------------------------------------------------------------------------
struct nb200_dbatch {
    cudaStream_t stream = nullptr;
    // results go back on their own stream, ordered after the producing pass by an event, so the next pass (of this or
    // another batch) computes while they are on the wire; a pass / upload waits for the copies that read what it rewrites
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_pass = nullptr, ev_copied_occ = nullptr, ev_copied_nuc = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    // pinned staging for the small per-chunk offset tables a pass uploads (5 slots of n_chunks + 1 int64): a copy from
    // pageable memory would synchronise the stream with the host in the middle of a pass
    int64_t *h_pin = nullptr;
    size_t h_pin_cap = 0;   // elements per slot
    int64_t *pin_slot(int slot) { return h_pin + (size_t)slot * h_pin_cap; }
    int n_chunks = 0;
    int64_t total_len = 0;   // sum of chunk lengths
    int64_t n_frag = 0;
    int64_t n_seq = 0;
    int64_t h2d_bytes = 0;
    bool have_seq = false;
    int max_len = 0, min_len = 0;
    // geometry (host copies for launch configuration / validation)
    std::vector<int32_t> h_start, h_end, h_seq_start;
    std::vector<int64_t> h_out_off, h_frag_off, h_seq_off;
    // device inputs
    DevBuf d_start, d_end, d_frag_off, d_pos, d_tlen, d_seq_off, d_seq_start, d_seq, d_out_off;
    // bias track: E = exp(log-bias) packed per chunk; chunk c covers genomic [bias0[c], bias0[c]+bias_len[c])
    bool bias_done = false;
    DevBuf d_bias_off;  // int64 [n+1]
    DevBuf d_E;         // f64 packed
    int64_t n_bias = 0;
    // fragment matrix in compressed-column form over genomic columns [start-pad, end+pad)
    int csc_pad = -1, csc_upper = -1, csc_atac = -1, csc_lower_split = -1;
    DevBuf d_col_off;   // int64 [n+1] offsets into col_ptr / col_low (chunk c has ncol_c+1 entries)
    DevBuf d_col_ptr;   // int32 packed: exclusive prefix of per-column counts (all rows < upper)
    DevBuf d_col_low;   // int32 packed: same for rows < lower_split (only when lower_split > 0)
    DevBuf d_cursor;    // int32 scratch, same shape
    DevBuf d_ent;       // int2 {col (relative to start-pad), row} packed at frag_off
    int64_t n_colptr = 0;
    // occ outputs (device)
    DevBuf o_vals, o_lower, o_upper, o_svals, o_slower, o_supper, o_cov, o_nuc_dist;
    DevBuf o_peak_count, o_peak_pos, o_peak_occ, o_peak_lower, o_peak_upper, o_peak_reads;
    DevBuf o_cn, o_cf;          // per-column sums of pn*Bp, pf*Bp over [start-flank, end+flank)
    DevBuf o_wsn, o_wsf;        // their sums over every occupancy window (k_occ_winsums)
    DevBuf o_wv;                // per-window occ / lower / upper values (3 slabs), the block smoother's input
    DevBuf o_peak_off;          // int64 [n+1]
    std::vector<int64_t> h_opeak_off;
    bool occ_done = false;
    int occ_cols_gen = -1;      // ctx->occ_gen the column sums o_cn / o_cf of this batch were computed for (-1: not computed)
    int occ_upper = 0;
    // nuc outputs (device)
    DevBuf n_signal, n_bg, n_norm, n_smooth, n_nuc_cov, n_nfr_cov, n_bx, n_bcov, n_cB, n_comb, n_cand_bcov;
    DevBuf n_cand_count, n_cand_pos, n_cand_flag, n_cand_z, n_cand_lr, n_cand_norm, n_cand_sig, n_cand_cov,
        n_cand_nfr, n_cand_smooth;
    DevBuf n_cand_off;          // int64 [n+1]
    std::vector<int64_t> h_ncand_off;
    DevBuf n_work;              // candidate work list {chunk, slot}
    DevBuf n_work_count;        // int32[2]: appended, consumed
    bool nuc_done = false;
    // scratch shared by the peak callers
    DevBuf sc_i32, sc_f64, sc_u8;
    // float32 staging of the tracks a *_download32 call converts on the device before the copy (one slab per track)
    DevBuf pack32_occ, pack32_nuc;
};

// stage drivers implemented across the .cu files
int nb200_prep_bias(nb200_ctx *ctx, nb200_dbatch *b);
int nb200_prep_csc(nb200_ctx *ctx, nb200_dbatch *b, int pad, int upper, int atac, int lower_split);
int nb200_nuc_bx_fp64(nb200_ctx *ctx, nb200_dbatch *b);   // dense background xcor, fp64 CUDA cores
int nb200_nuc_bx_tc(nb200_ctx *ctx, nb200_dbatch *b);     // dense background xcor, tcgen05
int nb200_tc_setup(nb200_ctx *ctx);                       // (re)build tensor-core operands after vmat / sizes change
int nb200_tc_available(nb200_ctx *ctx);
void nb200_tc_release(nb200_ctx *ctx);

static inline int64_t div_up64(int64_t a, int64_t b) { return (a + b - 1) / b; }
